// forge2d_b200 — world queries between steps: ray casts, AABB overlap, explosion. They run on the host image of the
// device-resident world (brought up to date on demand) because their protocol is a synchronous host callback per
// candidate shape whose return value steers the traversal (B2/src/world.c:2040-2310, types.h:1202-1220); candidate
// order and clipping follow B2/src/dynamic_tree.c:1114-1290 and the shape ray casts B2/src/geometry.c:506-760.
#pragma once
#include "f2d_distance.h"
#include "f2d_tree.h"

namespace f2d
{

struct RayInput
{
	V2 origin, translation;
	float maxFraction;
};
struct CastOutput
{
	V2 normal, point;
	float fraction;
	int32_t iterations;
	bool hit;
};
inline CastOutput noHit()
{
	CastOutput o;
	memset( &o, 0, sizeof( o ) );
	return o;
}

// geometry.c:506-558
inline CastOutput rayCastCircle( const RayInput& in, V2 center, float radius )
{
	CastOutput out = noHit();
	V2 s = sub( in.origin, center );
	float rr = radius * radius;
	float len;
	V2 d = lengthAndNormalize( &len, in.translation );
	auto originInside = [&]() {
		if ( lengthSq( s ) < rr )
		{
			out.point = in.origin;
			out.hit = true;
		}
	};
	if ( len == 0.0f )
	{
		originInside();
		return out;
	}
	float t = -dot( s, d );
	V2 c = mulAdd( s, t, d );
	float cc = dot( c, c );
	if ( cc > rr )
		return out;
	float h = sqrtf( rr - cc );
	float fraction = t - h;
	if ( fraction < 0.0f || in.maxFraction * len < fraction )
	{
		originInside();
		return out;
	}
	V2 hitPoint = mulAdd( s, fraction, d );
	out.fraction = fraction / len;
	out.normal = normalize( hitPoint );
	out.point = mulAdd( center, radius, out.normal );
	out.hit = true;
	return out;
}

// geometry.c:560-673
inline CastOutput rayCastCapsule( const RayInput& in, const Capsule& shape )
{
	CastOutput out = noHit();
	V2 v1 = shape.c1, v2 = shape.c2;
	float capsuleLength;
	V2 a = lengthAndNormalize( &capsuleLength, sub( v2, v1 ) );
	if ( capsuleLength < FLT_EPSILON )
		return rayCastCircle( in, v1, shape.radius );
	V2 p1 = in.origin;
	V2 d = in.translation;
	V2 q = sub( p1, v1 );
	float qa = dot( q, a );
	V2 qp = mulAdd( q, -qa, a );
	float radius = shape.radius;
	if ( dot( qp, qp ) < radius * radius )
	{
		if ( qa < 0.0f )
			return rayCastCircle( in, v1, radius );
		if ( qa > capsuleLength )
			return rayCastCircle( in, v2, radius );
		out.point = in.origin;
		out.hit = true;
		return out;
	}
	V2 n = { a.y, -a.x };
	float rayLength;
	V2 u = lengthAndNormalize( &rayLength, d );
	float den = -a.x * u.y + u.x * a.y;
	if ( -FLT_EPSILON < den && den < FLT_EPSILON )
		return out;
	V2 b1 = mulSub( q, radius, n );
	V2 b2 = mulAdd( q, radius, n );
	float invDen = 1.0f / den;
	float s21 = ( a.x * b1.y - b1.x * a.y ) * invDen;
	float s22 = ( a.x * b2.y - b2.x * a.y ) * invDen;
	float s2;
	V2 b;
	if ( s21 < s22 )
	{
		s2 = s21;
		b = b1;
	}
	else
	{
		s2 = s22;
		b = b2;
		n = neg( n );
	}
	if ( s2 < 0.0f || in.maxFraction * rayLength < s2 )
		return out;
	float s1 = ( -b.x * u.y + u.x * b.y ) * invDen;
	if ( s1 < 0.0f )
		return rayCastCircle( in, v1, radius );
	if ( capsuleLength < s1 )
		return rayCastCircle( in, v2, radius );
	out.fraction = s2 / rayLength;
	out.point = add( lerp( v1, v2, s1 / capsuleLength ), mulSV( radius, n ) );
	out.normal = n;
	out.hit = true;
	return out;
}

// geometry.c:675-749
inline CastOutput rayCastSegment( const RayInput& in, const Segment& shape, bool oneSided )
{
	CastOutput out = noHit();
	if ( oneSided )
	{
		float offset = cross( sub( in.origin, shape.p1 ), sub( shape.p2, shape.p1 ) );
		if ( offset < 0.0f )
			return out;
	}
	V2 p1 = in.origin, d = in.translation;
	V2 v1 = shape.p1, v2 = shape.p2;
	float len;
	V2 eUnit = lengthAndNormalize( &len, sub( v2, v1 ) );
	if ( len == 0.0f )
		return out;
	V2 normal = rightPerp( eUnit );
	float numerator = dot( normal, sub( v1, p1 ) );
	float denominator = dot( normal, d );
	if ( denominator == 0.0f )
		return out;
	float t = numerator / denominator;
	if ( t < 0.0f || in.maxFraction < t )
		return out;
	V2 p = mulAdd( p1, t, d );
	float s = dot( sub( p, v1 ), eUnit );
	if ( s < 0.0f || len < s )
		return out;
	if ( numerator > 0.0f )
		normal = neg( normal );
	out.fraction = t;
	out.point = p;
	out.normal = normal;
	out.hit = true;
	return out;
}

// b2ShapeCast, distance.c:608-697: conservative advancement of proxy B along `translationB` against proxy A with the
// GJK distance (core shapes, no radii) until the surfaces are `target` apart; canEncroach is false on the only path that
// uses it here (the rounded-polygon ray cast below).
inline CastOutput shapeCast( const ShapeProxy& proxyA, const ShapeProxy& proxyB, Xf transformA, Xf transformB, V2 translationB,
							 float maxFraction )
{
	float linearSlop = kLinearSlop;
	float totalRadius = proxyA.radius + proxyB.radius;
	float target = maxf( linearSlop, totalRadius - linearSlop );
	float tolerance = 0.25f * linearSlop;
	SimplexCache cache;
	memset( &cache, 0, sizeof( cache ) );
	float fraction = 0.0f;
	V2 delta2 = translationB;
	Xf xfB = transformB;
	CastOutput output = noHit();
	const int maxIterations = 20;
	for ( int iteration = 0; iteration < maxIterations; ++iteration )
	{
		output.iterations += 1;
		DistanceOutput distanceOutput = shapeDistance( proxyA, proxyB, transformA, xfB, false, &cache );
		if ( distanceOutput.distance < target + tolerance )
		{
			if ( iteration == 0 )
			{
				// initial overlap: a common point
				output.hit = true;
				V2 c1 = mulAdd( distanceOutput.pointA, proxyA.radius, distanceOutput.normal );
				V2 c2 = mulAdd( distanceOutput.pointB, -proxyB.radius, distanceOutput.normal );
				output.point = lerp( c1, c2, 0.5f );
				return output;
			}
			output.fraction = fraction;
			output.point = mulAdd( distanceOutput.pointA, proxyA.radius, distanceOutput.normal );
			output.normal = distanceOutput.normal;
			output.hit = true;
			return output;
		}
		// approaching?
		float denominator = dot( delta2, distanceOutput.normal );
		if ( denominator >= 0.0f )
			return output;
		fraction += ( target - distanceOutput.distance ) / denominator;
		if ( fraction >= maxFraction )
			return output;
		xfB.p = mulAdd( transformB.p, fraction, delta2 );
	}
	return output;
}

// geometry.c:799-887: sharp polygons by slab clipping, rounded ones through b2ShapeCast of the ray origin as a point
inline CastOutput rayCastPolygon( const RayInput& in, const Poly& shape )
{
	CastOutput out = noHit();
	if ( shape.radius != 0.0f )
	{
		const Xf identity = { { 0.0f, 0.0f }, { 1.0f, 0.0f } };
		return shapeCast( makeProxy( shape.v, shape.count, shape.radius ), makeProxy( &in.origin, 1, 0.0f ), identity, identity,
						  in.translation, in.maxFraction );
	}
	V2 base = shape.v[0];
	V2 p1 = sub( in.origin, base );
	V2 d = in.translation;
	float lower = 0.0f, upper = in.maxFraction;
	int index = -1;
	for ( int i = 0; i < shape.count; ++i )
	{
		V2 vertex = sub( shape.v[i], base );
		float numerator = dot( shape.n[i], sub( vertex, p1 ) );
		float denominator = dot( shape.n[i], d );
		if ( denominator == 0.0f )
		{
			if ( numerator < 0.0f )
				return out;
		}
		else
		{
			if ( denominator < 0.0f && numerator < lower * denominator )
			{
				lower = numerator / denominator;
				index = i;
			}
			else if ( denominator > 0.0f && numerator < upper * denominator )
			{
				upper = numerator / denominator;
			}
		}
		if ( upper < lower )
			return out;
	}
	if ( index >= 0 )
	{
		out.fraction = lower;
		out.normal = shape.n[index];
		out.point = mulAdd( in.origin, lower, d );
		out.hit = true;
	}
	else
	{
		out.point = in.origin;
		out.hit = true;
	}
	return out;
}

// shape.c:792-823 b2RayCastShape
inline CastOutput rayCastShape( const RayInput& in, const Shape& shape, Xf transform )
{
	RayInput local = in;
	local.origin = invRotate( transform.q, sub( in.origin, transform.p ) );
	local.translation = invRotate( transform.q, in.translation );
	CastOutput out = noHit();
	switch ( shape.type )
	{
		case kCapsule:
			out = rayCastCapsule( local, shape.capsule );
			break;
		case kCircle:
			out = rayCastCircle( local, shape.circle.center, shape.circle.radius );
			break;
		case kPolygon:
			out = rayCastPolygon( local, shape.polygon );
			break;
		case kSegment:
			out = rayCastSegment( local, shape.segment, false );
			break;
		case kChainSegment:
			out = rayCastSegment( local, shape.chainSegment.segment, true );
			break;
		default:
			return out;
	}
	out.point = xfPoint( transform, out.point );
	out.normal = rotate( transform.q, out.normal );
	return out;
}

struct TreeStats
{
	int nodeVisits, leafVisits;
};

// dynamic_tree.c:1114-1170 with visit counts (treeQuery in f2d_tree.h is the step's copy without them)
template <class F> inline TreeStats treeQueryStats( World* w, const Tree& t, Box box, uint64_t maskBits, F&& visit )
{
	TreeStats stats = { 0, 0 };
	if ( t.nodeCount == 0 )
		return stats;
	const TreeNode* nodes = ptr( w, t.nodes );
	int32_t stack[kTreeStack];
	int sp = 0;
	stack[sp++] = t.root;
	while ( sp > 0 )
	{
		int id = stack[--sp];
		if ( id == kNull )
			continue;
		const TreeNode& n = nodes[id];
		stats.nodeVisits += 1;
		if ( boxOverlaps( n.box, box ) && ( n.category & maskBits ) != 0 )
		{
			if ( n.flags & kNodeLeaf )
			{
				bool proceed = visit( id, n.userData );
				stats.leafVisits += 1;
				if ( proceed == false )
					return stats;
			}
			else if ( sp < kTreeStack - 1 )
			{
				stack[sp++] = n.child1;
				stack[sp++] = n.child2;
			}
		}
	}
	return stats;
}

// dynamic_tree.c:1172-1290. `visit(subInput, proxyId, userData)` returns the new clip fraction (0 stops, < 0 or
// > maxFraction ignores).
template <class F> inline TreeStats treeRayCast( World* w, const Tree& t, const RayInput& input, uint64_t maskBits, F&& visit )
{
	TreeStats stats = { 0, 0 };
	if ( t.nodeCount == 0 )
		return stats;
	V2 p1 = input.origin;
	V2 d = input.translation;
	V2 r = normalize( d );
	V2 v = crossSV( 1.0f, r );
	V2 abs_v = { absf( v.x ), absf( v.y ) };
	float maxFraction = input.maxFraction;
	V2 p2 = mulAdd( p1, maxFraction, d );
	Box segmentBox = { vmin( p1, p2 ), vmax( p1, p2 ) };
	const TreeNode* nodes = ptr( w, t.nodes );
	int32_t stack[kTreeStack];
	int sp = 0;
	stack[sp++] = t.root;
	RayInput subInput = input;
	while ( sp > 0 )
	{
		int id = stack[--sp];
		if ( id == kNull )
			continue;
		const TreeNode& n = nodes[id];
		stats.nodeVisits += 1;
		Box nodeBox = n.box;
		if ( ( n.category & maskBits ) == 0 || boxOverlaps( nodeBox, segmentBox ) == false )
			continue;
		V2 c = boxCenter( nodeBox );
		V2 h = { 0.5f * ( nodeBox.hi.x - nodeBox.lo.x ), 0.5f * ( nodeBox.hi.y - nodeBox.lo.y ) };
		float term1 = absf( dot( v, sub( p1, c ) ) );
		float term2 = dot( abs_v, h );
		if ( term2 < term1 )
			continue;
		if ( n.flags & kNodeLeaf )
		{
			subInput.maxFraction = maxFraction;
			float value = visit( subInput, id, n.userData );
			stats.leafVisits += 1;
			if ( value == 0.0f )
				return stats;
			if ( 0.0f < value && value <= maxFraction )
			{
				maxFraction = value;
				p2 = mulAdd( p1, maxFraction, d );
				segmentBox.lo = vmin( p1, p2 );
				segmentBox.hi = vmax( p1, p2 );
			}
		}
		else if ( sp < kTreeStack - 1 )
		{
			V2 c1 = boxCenter( nodes[n.child1].box );
			V2 c2 = boxCenter( nodes[n.child2].box );
			if ( distanceSq( c1, p1 ) < distanceSq( c2, p1 ) )
			{
				stack[sp++] = n.child2;
				stack[sp++] = n.child1;
			}
			else
			{
				stack[sp++] = n.child1;
				stack[sp++] = n.child2;
			}
		}
	}
	return stats;
}

// shape.h:132-135 b2ShouldQueryCollide
inline bool shouldQueryCollide( const Filter& shapeFilter, uint64_t queryCategory, uint64_t queryMask )
{
	return ( shapeFilter.category & queryMask ) != 0 && ( shapeFilter.mask & queryCategory ) != 0;
}

// shape.c:656-705 b2GetShapeProjectedPerimeter
inline float shapeProjectedPerimeter( const Shape& shape, V2 line )
{
	switch ( shape.type )
	{
		case kCapsule:
		{
			V2 axis = sub( shape.capsule.c2, shape.capsule.c1 );
			return absf( dot( axis, line ) ) + 2.0f * shape.capsule.radius;
		}
		case kCircle:
			return 2.0f * shape.circle.radius;
		case kPolygon:
		{
			float value = dot( shape.polygon.v[0], line );
			float lower = value, upper = value;
			for ( int i = 1; i < shape.polygon.count; ++i )
			{
				value = dot( shape.polygon.v[i], line );
				lower = minf( lower, value );
				upper = maxf( upper, value );
			}
			return ( upper - lower ) + 2.0f * shape.polygon.radius;
		}
		case kSegment:
			return absf( dot( shape.segment.p2, line ) - dot( shape.segment.p1, line ) );
		case kChainSegment:
			return absf( dot( shape.chainSegment.segment.p2, line ) - dot( shape.chainSegment.segment.p1, line ) );
		default:
			return 0.0f;
	}
}

} // namespace f2d
