// forge2d_b200 — GJK distance and conservative-advancement time of impact for continuous collision.
// Only fast bodies reach this code (a few per step) but its result moves bodies, so it follows
// B2/src/distance.c decision for decision (cited per function).
#pragma once
#include "f2d_types.h"

namespace f2d
{

struct ShapeProxy // collision.h b2ShapeProxy
{
	V2 points[kMaxPolyVerts];
	int32_t count;
	float radius;
};
struct Sweep // collision.h b2Sweep
{
	V2 localCenter, c1, c2;
	Rot q1, q2;
};

// distance.c:108-122
F2D_HD ShapeProxy makeProxy( const V2* points, int count, float radius )
{
	count = mini( count, kMaxPolyVerts );
	ShapeProxy p;
	for ( int i = 0; i < count; ++i )
		p.points[i] = points[i];
	p.count = count;
	p.radius = radius;
	return p;
}

// shape.c:922-943
F2D_HD ShapeProxy makeShapeProxy( const Shape& s )
{
	switch ( s.type )
	{
		case kCapsule:
			return makeProxy( &s.capsule.c1, 2, s.capsule.radius );
		case kCircle:
			return makeProxy( &s.circle.center, 1, s.circle.radius );
		case kPolygon:
			return makeProxy( s.polygon.v, s.polygon.count, s.polygon.radius );
		case kSegment:
			return makeProxy( &s.segment.p1, 2, 0.0f );
		default:
			return makeProxy( &s.chainSegment.segment.p1, 2, 0.0f );
	}
}

// distance.c:13-31
F2D_HD Xf sweepTransform( const Sweep& sw, float time )
{
	Xf xf;
	xf.p = add( mulSV( 1.0f - time, sw.c1 ), mulSV( time, sw.c2 ) );
	Rot q = { ( 1.0f - time ) * sw.q1.c + time * sw.q2.c, ( 1.0f - time ) * sw.q1.s + time * sw.q2.s };
	xf.q = normalizeRot( q );
	xf.p = sub( xf.p, rotate( xf.q, sw.localCenter ) );
	return xf;
}

struct SimplexVertex
{
	V2 wA, wB, w;
	float a;
	int32_t indexA, indexB;
};
struct Simplex
{
	SimplexVertex v[3];
	int32_t count;
};

F2D_HD V2 weight2( float a1, V2 w1, float a2, V2 w2 ) { return V2{ a1 * w1.x + a2 * w2.x, a1 * w1.y + a2 * w2.y }; }
F2D_HD V2 weight3( float a1, V2 w1, float a2, V2 w2, float a3, V2 w3 )
{
	return V2{ a1 * w1.x + a2 * w2.x + a3 * w3.x, a1 * w1.y + a2 * w2.y + a3 * w3.y };
}

// distance.c:152-171
F2D_HD int findSupport( const ShapeProxy& proxy, V2 direction )
{
	int bestIndex = 0;
	float bestValue = dot( proxy.points[0], direction );
	for ( int i = 1; i < proxy.count; ++i )
	{
		float value = dot( proxy.points[i], direction );
		if ( value > bestValue )
		{
			bestIndex = i;
			bestValue = value;
		}
	}
	return bestIndex;
}

// distance.c:173-208
F2D_HD Simplex makeSimplexFromCache( const SimplexCache& cache, const ShapeProxy& proxyA, const ShapeProxy& proxyB )
{
	Simplex s;
	s.count = cache.count;
	for ( int i = 0; i < s.count; ++i )
	{
		SimplexVertex& v = s.v[i];
		v.indexA = cache.indexA[i];
		v.indexB = cache.indexB[i];
		v.wA = proxyA.points[v.indexA];
		v.wB = proxyB.points[v.indexB];
		v.w = sub( v.wA, v.wB );
		v.a = -1.0f;
	}
	if ( s.count == 0 )
	{
		SimplexVertex& v = s.v[0];
		v.indexA = 0;
		v.indexB = 0;
		v.wA = proxyA.points[0];
		v.wB = proxyB.points[0];
		v.w = sub( v.wA, v.wB );
		v.a = 1.0f;
		s.count = 1;
	}
	return s;
}

F2D_HD void makeSimplexCache( SimplexCache& cache, const Simplex& s )
{
	cache.count = (uint16_t)s.count;
	for ( int i = 0; i < s.count; ++i )
	{
		cache.indexA[i] = (uint8_t)s.v[i].indexA;
		cache.indexB[i] = (uint8_t)s.v[i].indexB;
	}
}

// distance.c:221-246
F2D_HD void simplexWitness( V2* a, V2* b, const Simplex& s )
{
	switch ( s.count )
	{
		case 1:
			*a = s.v[0].wA;
			*b = s.v[0].wB;
			break;
		case 2:
			*a = weight2( s.v[0].a, s.v[0].wA, s.v[1].a, s.v[1].wA );
			*b = weight2( s.v[0].a, s.v[0].wB, s.v[1].a, s.v[1].wB );
			break;
		case 3:
			*a = weight3( s.v[0].a, s.v[0].wA, s.v[1].a, s.v[1].wA, s.v[2].a, s.v[2].wA );
			*b = *a;
			break;
		default:
			*a = V2{ 0.0f, 0.0f };
			*b = V2{ 0.0f, 0.0f };
			break;
	}
}

// distance.c:266-300
F2D_HD V2 solveSimplex2( Simplex& s )
{
	V2 w1 = s.v[0].w, w2 = s.v[1].w;
	V2 e12 = sub( w2, w1 );
	float d12_2 = -dot( w1, e12 );
	if ( d12_2 <= 0.0f )
	{
		s.v[0].a = 1.0f;
		s.count = 1;
		return neg( w1 );
	}
	float d12_1 = dot( w2, e12 );
	if ( d12_1 <= 0.0f )
	{
		s.v[1].a = 1.0f;
		s.count = 1;
		s.v[0] = s.v[1];
		return neg( w2 );
	}
	float inv_d12 = 1.0f / ( d12_1 + d12_2 );
	s.v[0].a = d12_1 * inv_d12;
	s.v[1].a = d12_2 * inv_d12;
	s.count = 2;
	return crossSV( cross( add( w1, w2 ), e12 ), e12 );
}

// distance.c:302-419
F2D_HDF inline V2 solveSimplex3( Simplex& s )
{
	V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
	V2 e12 = sub( w2, w1 );
	float w1e12 = dot( w1, e12 );
	float w2e12 = dot( w2, e12 );
	float d12_1 = w2e12;
	float d12_2 = -w1e12;
	V2 e13 = sub( w3, w1 );
	float w1e13 = dot( w1, e13 );
	float w3e13 = dot( w3, e13 );
	float d13_1 = w3e13;
	float d13_2 = -w1e13;
	V2 e23 = sub( w3, w2 );
	float w2e23 = dot( w2, e23 );
	float w3e23 = dot( w3, e23 );
	float d23_1 = w3e23;
	float d23_2 = -w2e23;
	float n123 = cross( e12, e13 );
	float d123_1 = n123 * cross( w2, w3 );
	float d123_2 = n123 * cross( w3, w1 );
	float d123_3 = n123 * cross( w1, w2 );

	if ( d12_2 <= 0.0f && d13_2 <= 0.0f )
	{
		s.v[0].a = 1.0f;
		s.count = 1;
		return neg( w1 );
	}
	if ( d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f )
	{
		float inv = 1.0f / ( d12_1 + d12_2 );
		s.v[0].a = d12_1 * inv;
		s.v[1].a = d12_2 * inv;
		s.count = 2;
		return crossSV( cross( add( w1, w2 ), e12 ), e12 );
	}
	if ( d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f )
	{
		float inv = 1.0f / ( d13_1 + d13_2 );
		s.v[0].a = d13_1 * inv;
		s.v[2].a = d13_2 * inv;
		s.count = 2;
		s.v[1] = s.v[2];
		return crossSV( cross( add( w1, w3 ), e13 ), e13 );
	}
	if ( d12_1 <= 0.0f && d23_2 <= 0.0f )
	{
		s.v[1].a = 1.0f;
		s.count = 1;
		s.v[0] = s.v[1];
		return neg( w2 );
	}
	if ( d13_1 <= 0.0f && d23_1 <= 0.0f )
	{
		s.v[2].a = 1.0f;
		s.count = 1;
		s.v[0] = s.v[2];
		return neg( w3 );
	}
	if ( d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f )
	{
		float inv = 1.0f / ( d23_1 + d23_2 );
		s.v[1].a = d23_1 * inv;
		s.v[2].a = d23_2 * inv;
		s.count = 2;
		s.v[0] = s.v[2];
		return crossSV( cross( add( w2, w3 ), e23 ), e23 );
	}
	float inv = 1.0f / ( d123_1 + d123_2 + d123_3 );
	s.v[0].a = d123_1 * inv;
	s.v[1].a = d123_2 * inv;
	s.v[2].a = d123_3 * inv;
	s.count = 3;
	return V2{ 0.0f, 0.0f };
}

struct DistanceOutput
{
	V2 pointA, pointB, normal;
	float distance;
	int32_t iterations;
};

// GJK: distance.c:421-605 (simplex recording omitted: debug-only in the reference)
F2D_HDF inline DistanceOutput shapeDistance( const ShapeProxy& proxyA, const ShapeProxy& proxyBIn, Xf xfA, Xf xfB, bool useRadii,
											 SimplexCache* cache )
{
	DistanceOutput out;
	memset( &out, 0, sizeof( out ) );
	ShapeProxy localB;
	{
		Xf t = invMulXf( xfA, xfB );
		localB.count = proxyBIn.count;
		localB.radius = proxyBIn.radius;
		for ( int i = 0; i < localB.count; ++i )
			localB.points[i] = xfPoint( t, proxyBIn.points[i] );
	}
	Simplex simplex = makeSimplexFromCache( *cache, proxyA, localB );
	V2 nonUnitNormal = { 0.0f, 0.0f };
	int saveA[3], saveB[3];
	const int maxIterations = 20;
	int iteration = 0;
	while ( iteration < maxIterations )
	{
		int saveCount = simplex.count;
		for ( int i = 0; i < saveCount; ++i )
		{
			saveA[i] = simplex.v[i].indexA;
			saveB[i] = simplex.v[i].indexB;
		}
		V2 d = { 0.0f, 0.0f };
		switch ( simplex.count )
		{
			case 1:
				d = neg( simplex.v[0].w );
				break;
			case 2:
				d = solveSimplex2( simplex );
				break;
			case 3:
				d = solveSimplex3( simplex );
				break;
			default:
				break;
		}
		if ( simplex.count == 3 )
		{
			V2 lA, lB;
			simplexWitness( &lA, &lB, simplex );
			out.pointA = xfPoint( xfA, lA );
			out.pointB = xfPoint( xfA, lB );
			return out;
		}
		if ( dot( d, d ) < FLT_EPSILON * FLT_EPSILON )
		{
			V2 lA, lB;
			simplexWitness( &lA, &lB, simplex );
			out.pointA = xfPoint( xfA, lA );
			out.pointB = xfPoint( xfA, lB );
			return out;
		}
		nonUnitNormal = d;
		SimplexVertex& vertex = simplex.v[simplex.count];
		vertex.indexA = findSupport( proxyA, d );
		vertex.wA = proxyA.points[vertex.indexA];
		vertex.indexB = findSupport( localB, neg( d ) );
		vertex.wB = localB.points[vertex.indexB];
		vertex.w = sub( vertex.wA, vertex.wB );
		++iteration;
		bool duplicate = false;
		for ( int i = 0; i < saveCount; ++i )
		{
			if ( vertex.indexA == saveA[i] && vertex.indexB == saveB[i] )
			{
				duplicate = true;
				break;
			}
		}
		if ( duplicate )
			break;
		simplex.count += 1;
	}
	V2 normal = normalize( nonUnitNormal );
	normal = rotate( xfA.q, normal );
	V2 lA, lB;
	simplexWitness( &lA, &lB, simplex );
	out.normal = normal;
	out.distance = distance( lA, lB );
	out.pointA = xfPoint( xfA, lA );
	out.pointB = xfPoint( xfA, lB );
	out.iterations = iteration;
	makeSimplexCache( *cache, simplex );
	if ( useRadii && out.distance > 0.1f * kLinearSlop )
	{
		float rA = proxyA.radius, rB = proxyBIn.radius;
		out.distance = maxf( 0.0f, out.distance - rA - rB );
		out.pointA = mulAdd( out.pointA, rA, normal );
		out.pointB = mulSub( out.pointB, rB, normal );
	}
	return out;
}

// Separating-axis function for the TOI root finder: distance.c:936-1135
enum : int
{
	kSepPoints = 0,
	kSepFaceA = 1,
	kSepFaceB = 2
};
struct SeparationFcn
{
	const ShapeProxy* proxyA;
	const ShapeProxy* proxyB;
	Sweep sweepA, sweepB;
	V2 localPoint, axis;
	int32_t type;
};

F2D_HDF inline SeparationFcn makeSeparationFcn( const SimplexCache& cache, const ShapeProxy* proxyA, const Sweep& sweepA,
												const ShapeProxy* proxyB, const Sweep& sweepB, float t1 )
{
	SeparationFcn f;
	f.proxyA = proxyA;
	f.proxyB = proxyB;
	int count = cache.count;
	f.sweepA = sweepA;
	f.sweepB = sweepB;
	Xf xfA = sweepTransform( sweepA, t1 );
	Xf xfB = sweepTransform( sweepB, t1 );
	if ( count == 1 )
	{
		f.type = kSepPoints;
		V2 lA = proxyA->points[cache.indexA[0]];
		V2 lB = proxyB->points[cache.indexB[0]];
		V2 pA = xfPoint( xfA, lA );
		V2 pB = xfPoint( xfB, lB );
		f.axis = normalize( sub( pB, pA ) );
		f.localPoint = V2{ 0.0f, 0.0f };
		return f;
	}
	if ( cache.indexA[0] == cache.indexA[1] )
	{
		f.type = kSepFaceB;
		V2 lB1 = proxyB->points[cache.indexB[0]];
		V2 lB2 = proxyB->points[cache.indexB[1]];
		f.axis = crossVS( sub( lB2, lB1 ), 1.0f );
		f.axis = normalize( f.axis );
		V2 normal = rotate( xfB.q, f.axis );
		f.localPoint = V2{ 0.5f * ( lB1.x + lB2.x ), 0.5f * ( lB1.y + lB2.y ) };
		V2 pB = xfPoint( xfB, f.localPoint );
		V2 lA = proxyA->points[cache.indexA[0]];
		V2 pA = xfPoint( xfA, lA );
		float s = dot( sub( pA, pB ), normal );
		if ( s < 0.0f )
			f.axis = neg( f.axis );
		return f;
	}
	f.type = kSepFaceA;
	V2 lA1 = proxyA->points[cache.indexA[0]];
	V2 lA2 = proxyA->points[cache.indexA[1]];
	f.axis = crossVS( sub( lA2, lA1 ), 1.0f );
	f.axis = normalize( f.axis );
	V2 normal = rotate( xfA.q, f.axis );
	f.localPoint = V2{ 0.5f * ( lA1.x + lA2.x ), 0.5f * ( lA1.y + lA2.y ) };
	V2 pA = xfPoint( xfA, f.localPoint );
	V2 lB = proxyB->points[cache.indexB[0]];
	V2 pB = xfPoint( xfB, lB );
	float s = dot( sub( pB, pA ), normal );
	if ( s < 0.0f )
		f.axis = neg( f.axis );
	return f;
}

F2D_HDF inline float findMinSeparation( const SeparationFcn& f, int* indexA, int* indexB, float t )
{
	Xf xfA = sweepTransform( f.sweepA, t );
	Xf xfB = sweepTransform( f.sweepB, t );
	switch ( f.type )
	{
		case kSepPoints:
		{
			V2 axisA = invRotate( xfA.q, f.axis );
			V2 axisB = invRotate( xfB.q, neg( f.axis ) );
			*indexA = findSupport( *f.proxyA, axisA );
			*indexB = findSupport( *f.proxyB, axisB );
			V2 pA = xfPoint( xfA, f.proxyA->points[*indexA] );
			V2 pB = xfPoint( xfB, f.proxyB->points[*indexB] );
			return dot( sub( pB, pA ), f.axis );
		}
		case kSepFaceA:
		{
			V2 normal = rotate( xfA.q, f.axis );
			V2 pA = xfPoint( xfA, f.localPoint );
			V2 axisB = invRotate( xfB.q, neg( normal ) );
			*indexA = -1;
			*indexB = findSupport( *f.proxyB, axisB );
			V2 pB = xfPoint( xfB, f.proxyB->points[*indexB] );
			return dot( sub( pB, pA ), normal );
		}
		case kSepFaceB:
		{
			V2 normal = rotate( xfB.q, f.axis );
			V2 pB = xfPoint( xfB, f.localPoint );
			V2 axisA = invRotate( xfA.q, neg( normal ) );
			*indexB = -1;
			*indexA = findSupport( *f.proxyA, axisA );
			V2 pA = xfPoint( xfA, f.proxyA->points[*indexA] );
			return dot( sub( pA, pB ), normal );
		}
		default:
			*indexA = -1;
			*indexB = -1;
			return 0.0f;
	}
}

F2D_HDF inline float evaluateSeparation( const SeparationFcn& f, int indexA, int indexB, float t )
{
	Xf xfA = sweepTransform( f.sweepA, t );
	Xf xfB = sweepTransform( f.sweepB, t );
	switch ( f.type )
	{
		case kSepPoints:
		{
			V2 pA = xfPoint( xfA, f.proxyA->points[indexA] );
			V2 pB = xfPoint( xfB, f.proxyB->points[indexB] );
			return dot( sub( pB, pA ), f.axis );
		}
		case kSepFaceA:
		{
			V2 normal = rotate( xfA.q, f.axis );
			V2 pA = xfPoint( xfA, f.localPoint );
			V2 pB = xfPoint( xfB, f.proxyB->points[indexB] );
			return dot( sub( pB, pA ), normal );
		}
		case kSepFaceB:
		{
			V2 normal = rotate( xfB.q, f.axis );
			V2 pB = xfPoint( xfB, f.localPoint );
			V2 pA = xfPoint( xfA, f.proxyA->points[indexA] );
			return dot( sub( pA, pB ), normal );
		}
		default:
			return 0.0f;
	}
}

// Conservative advancement: distance.c:1137-1414. Returns the impact fraction (output.fraction).
F2D_HDF inline float timeOfImpact( const ShapeProxy& proxyA, const ShapeProxy& proxyB, const Sweep& sweepA, const Sweep& sweepB,
								   float maxFraction )
{
	float fraction = maxFraction;
	float tMax = maxFraction;
	float totalRadius = proxyA.radius + proxyB.radius;
	float target = maxf( kLinearSlop, totalRadius - kLinearSlop );
	float tolerance = 0.25f * kLinearSlop;
	float t1 = 0.0f;
	const int kMaxIterations = 20;
	int distanceIterations = 0;
	SimplexCache cache;
	memset( &cache, 0, sizeof( cache ) );

	for ( ;; )
	{
		Xf xfA = sweepTransform( sweepA, t1 );
		Xf xfB = sweepTransform( sweepB, t1 );
		DistanceOutput dout = shapeDistance( proxyA, proxyB, xfA, xfB, false, &cache );
		distanceIterations += 1;
		if ( dout.distance <= 0.0f )
		{
			fraction = 0.0f; // overlapped
			break;
		}
		if ( dout.distance <= target + tolerance )
		{
			fraction = t1; // hit
			break;
		}
		SeparationFcn fcn = makeSeparationFcn( cache, &proxyA, sweepA, &proxyB, sweepB, t1 );
		bool done = false;
		float t2 = tMax;
		int pushBackIterations = 0;
		for ( ;; )
		{
			int indexA, indexB;
			float s2 = findMinSeparation( fcn, &indexA, &indexB, t2 );
			if ( s2 > target + tolerance )
			{
				fraction = tMax; // separated
				done = true;
				break;
			}
			if ( s2 > target - tolerance )
			{
				t1 = t2;
				break;
			}
			float s1 = evaluateSeparation( fcn, indexA, indexB, t1 );
			if ( s1 < target - tolerance )
			{
				fraction = t1; // failed
				done = true;
				break;
			}
			if ( s1 <= target + tolerance )
			{
				fraction = t1; // hit
				done = true;
				break;
			}
			int rootIterationCount = 0;
			float a1 = t1, a2 = t2;
			for ( ;; )
			{
				float t;
				if ( rootIterationCount & 1 )
					t = a1 + ( target - s1 ) * ( a2 - a1 ) / ( s2 - s1 );
				else
					t = 0.5f * ( a1 + a2 );
				rootIterationCount += 1;
				float s = evaluateSeparation( fcn, indexA, indexB, t );
				if ( absf( s - target ) < tolerance )
				{
					t2 = t;
					break;
				}
				if ( s > target )
				{
					a1 = t;
					s1 = s;
				}
				else
				{
					a2 = t;
					s2 = s;
				}
				if ( rootIterationCount == 50 )
					break;
			}
			pushBackIterations += 1;
			if ( pushBackIterations == kMaxPolyVerts )
				break;
		}
		if ( done )
			break;
		if ( distanceIterations == kMaxIterations )
		{
			fraction = t1; // failed
			break;
		}
	}
	return fraction;
}

} // namespace f2d
