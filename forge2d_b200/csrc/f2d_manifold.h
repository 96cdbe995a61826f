// forge2d_b200 — narrowphase manifolds (one function per shape-pair class).
//
// Same SAT / clipping / closest-feature decisions and the same floating-point expression order as
// B2/src/manifold.c (cited per function) so point counts, feature ids and anchors are bit-identical.
#pragma once
#include "f2d_distance.h"
#include "f2d_types.h"

namespace f2d
{

F2D_HD uint16_t makeFeatureId( int a, int b ) { return (uint16_t)( ( (uint8_t)a << 8 ) | (uint8_t)b ); } // manifold.c:14

F2D_HD Manifold emptyManifold()
{
	Manifold m;
	memset( &m, 0, sizeof( m ) );
	return m;
}

// manifold.c:16-34
F2D_HD Poly makeCapsulePoly( V2 p1, V2 p2, float radius )
{
	Poly s;
	memset( &s, 0, sizeof( s ) );
	s.v[0] = p1;
	s.v[1] = p2;
	s.centroid = lerp( p1, p2, 0.5f );
	V2 d = sub( p2, p1 );
	V2 axis = normalize( d );
	V2 normal = rightPerp( axis );
	s.n[0] = normal;
	s.n[1] = neg( normal );
	s.count = 2;
	s.radius = radius;
	return s;
}

// Writes the single world-space point shared by the circle-like manifolds (manifold.c:61-73, 126-138, ...)
F2D_HD void finishSinglePoint( Manifold& m, Xf xfA, Xf xfB, V2 normalLocal, V2 contactPointA, float separation, bool pointOrderAnchorFirst )
{
	m.normal = rotate( xfA.q, normalLocal );
	ManifoldPoint& mp = m.points[0];
	mp.anchorA = rotate( xfA.q, contactPointA );
	mp.anchorB = add( mp.anchorA, sub( xfA.p, xfB.p ) );
	mp.point = pointOrderAnchorFirst ? add( mp.anchorA, xfA.p ) : add( xfA.p, mp.anchorA );
	mp.separation = separation;
	mp.id = 0;
	m.pointCount = 1;
}

// manifold.c:40-74
F2D_HDF inline Manifold collideCircles( const Circle& a, Xf xfA, const Circle& b, Xf xfB )
{
	Manifold m = emptyManifold();
	Xf xf = invMulXf( xfA, xfB );
	V2 pA = a.center;
	V2 pB = xfPoint( xf, b.center );
	float dist;
	V2 normal = lengthAndNormalize( &dist, sub( pB, pA ) );
	float rA = a.radius, rB = b.radius;
	float separation = dist - rA - rB;
	if ( separation > kSpeculative )
		return m;
	V2 cA = mulAdd( pA, rA, normal );
	V2 cB = mulAdd( pB, -rB, normal );
	V2 contact = lerp( cA, cB, 0.5f );
	finishSinglePoint( m, xfA, xfB, normal, contact, separation, true );
	return m;
}

// manifold.c:77-139
F2D_HDF inline Manifold collideCapsuleAndCircle( const Capsule& a, Xf xfA, const Circle& b, Xf xfB )
{
	Manifold m = emptyManifold();
	Xf xf = invMulXf( xfA, xfB );
	V2 pB = xfPoint( xf, b.center );
	V2 p1 = a.c1, p2 = a.c2;
	V2 e = sub( p2, p1 );
	V2 pA;
	float s1 = dot( sub( pB, p1 ), e );
	float s2 = dot( sub( p2, pB ), e );
	if ( s1 < 0.0f )
		pA = p1;
	else if ( s2 < 0.0f )
		pA = p2;
	else
	{
		float s = s1 / dot( e, e );
		pA = mulAdd( p1, s, e );
	}
	float dist;
	V2 normal = lengthAndNormalize( &dist, sub( pB, pA ) );
	float rA = a.radius, rB = b.radius;
	float separation = dist - rA - rB;
	if ( separation > kSpeculative )
		return m;
	V2 cA = mulAdd( pA, rA, normal );
	V2 cB = mulAdd( pB, -rB, normal );
	V2 contact = lerp( cA, cB, 0.5f );
	finishSinglePoint( m, xfA, xfB, normal, contact, separation, false );
	return m;
}

// manifold.c:141-257
F2D_HDF inline Manifold collidePolygonAndCircle( const Poly& a, Xf xfA, const Circle& b, Xf xfB )
{
	Manifold m = emptyManifold();
	Xf xf = invMulXf( xfA, xfB );
	V2 center = xfPoint( xf, b.center );
	float rA = a.radius, rB = b.radius;
	float radius = rA + rB;

	int normalIndex = 0;
	float separation = -FLT_MAX;
	int count = a.count;
	for ( int i = 0; i < count; ++i )
	{
		float s = dot( a.n[i], sub( center, a.v[i] ) );
		if ( s > separation )
		{
			separation = s;
			normalIndex = i;
		}
	}
	if ( separation > radius + kSpeculative )
		return m;

	int i1 = normalIndex;
	int i2 = i1 + 1 < count ? i1 + 1 : 0;
	V2 v1 = a.v[i1], v2 = a.v[i2];
	float u1 = dot( sub( center, v1 ), sub( v2, v1 ) );
	float u2 = dot( sub( center, v2 ), sub( v1, v2 ) );

	if ( ( u1 < 0.0f && separation > FLT_EPSILON ) || ( u2 < 0.0f && separation > FLT_EPSILON ) )
	{
		// vertex region (v1 tested first, as in the reference's if / else-if)
		V2 v = ( u1 < 0.0f && separation > FLT_EPSILON ) ? v1 : v2;
		V2 normal = normalize( sub( center, v ) );
		separation = dot( sub( center, v ), normal );
		if ( separation > radius + kSpeculative )
			return m;
		V2 cA = mulAdd( v, rA, normal );
		V2 cB = mulSub( center, rB, normal );
		V2 contact = lerp( cA, cB, 0.5f );
		finishSinglePoint( m, xfA, xfB, normal, contact, dot( sub( cB, cA ), normal ), false );
	}
	else
	{
		V2 normal = a.n[normalIndex];
		V2 cA = mulAdd( center, rA - dot( sub( center, v1 ), normal ), normal );
		V2 cB = mulSub( center, rB, normal );
		V2 contact = lerp( cA, cB, 0.5f );
		finishSinglePoint( m, xfA, xfB, normal, contact, separation - radius, false );
	}
	return m;
}

// Capsule-capsule with optional two-point clipping: manifold.c:261-530
F2D_HDF inline Manifold collideCapsules( const Capsule& a, Xf xfA, const Capsule& b, Xf xfB )
{
	V2 origin = a.c1;
	Xf sfA = { add( xfA.p, rotate( xfA.q, origin ) ), xfA.q };
	Xf xf = invMulXf( sfA, xfB );

	V2 p1 = { 0.0f, 0.0f };
	V2 q1 = sub( a.c2, origin );
	V2 p2 = xfPoint( xf, b.c1 );
	V2 q2 = xfPoint( xf, b.c2 );

	V2 d1 = sub( q1, p1 );
	V2 d2 = sub( q2, p2 );
	float dd1 = dot( d1, d1 );
	float dd2 = dot( d2, d2 );
	const float epsSqr = FLT_EPSILON * FLT_EPSILON;

	V2 r = sub( p1, p2 );
	float rd1 = dot( r, d1 );
	float rd2 = dot( r, d2 );
	float d12 = dot( d1, d2 );
	float denom = dd1 * dd2 - d12 * d12;

	float f1 = 0.0f;
	if ( denom != 0.0f )
		f1 = clampf( ( d12 * rd2 - rd1 * dd2 ) / denom, 0.0f, 1.0f );
	float f2 = ( d12 * f1 + rd2 ) / dd2;
	if ( f2 < 0.0f )
	{
		f2 = 0.0f;
		f1 = clampf( -rd1 / dd1, 0.0f, 1.0f );
	}
	else if ( f2 > 1.0f )
	{
		f2 = 1.0f;
		f1 = clampf( ( d12 - rd1 ) / dd1, 0.0f, 1.0f );
	}

	V2 closest1 = mulAdd( p1, f1, d1 );
	V2 closest2 = mulAdd( p2, f2, d2 );
	float distSq = distanceSq( closest1, closest2 );

	Manifold m = emptyManifold();
	float rA = a.radius, rB = b.radius;
	float radius = rA + rB;
	float maxDistance = radius + kSpeculative;
	if ( distSq > maxDistance * maxDistance )
		return m;

	float dist = sqrtf( distSq );
	float length1, length2;
	V2 u1 = lengthAndNormalize( &length1, d1 );
	V2 u2 = lengthAndNormalize( &length2, d2 );

	float fp2 = dot( sub( p2, p1 ), u1 );
	float fq2 = dot( sub( q2, p1 ), u1 );
	bool outsideA = ( fp2 <= 0.0f && fq2 <= 0.0f ) || ( fp2 >= length1 && fq2 >= length1 );
	float fp1 = dot( sub( p1, p2 ), u2 );
	float fq1 = dot( sub( q1, p2 ), u2 );
	bool outsideB = ( fp1 <= 0.0f && fq1 <= 0.0f ) || ( fp1 >= length2 && fq1 >= length2 );

	if ( outsideA == false && outsideB == false )
	{
		V2 normalA;
		float separationA;
		{
			normalA = leftPerp( u1 );
			float ss1 = dot( sub( p2, p1 ), normalA );
			float ss2 = dot( sub( q2, p1 ), normalA );
			float s1p = ss1 < ss2 ? ss1 : ss2;
			float s1n = -ss1 < -ss2 ? -ss1 : -ss2;
			if ( s1p > s1n )
				separationA = s1p;
			else
			{
				separationA = s1n;
				normalA = neg( normalA );
			}
		}
		V2 normalB;
		float separationB;
		{
			normalB = leftPerp( u2 );
			float ss1 = dot( sub( p1, p2 ), normalB );
			float ss2 = dot( sub( q1, p2 ), normalB );
			float s1p = ss1 < ss2 ? ss1 : ss2;
			float s1n = -ss1 < -ss2 ? -ss1 : -ss2;
			if ( s1p > s1n )
				separationB = s1p;
			else
			{
				separationB = s1n;
				normalB = neg( normalB );
			}
		}

		if ( separationA + 0.1f * kLinearSlop >= separationB )
		{
			m.normal = normalA;
			V2 cp = p2, cq = q2;
			if ( fp2 < 0.0f && fq2 > 0.0f )
				cp = lerp( p2, q2, ( 0.0f - fp2 ) / ( fq2 - fp2 ) );
			else if ( fq2 < 0.0f && fp2 > 0.0f )
				cq = lerp( q2, p2, ( 0.0f - fq2 ) / ( fp2 - fq2 ) );
			if ( fp2 > length1 && fq2 < length1 )
				cp = lerp( p2, q2, ( fp2 - length1 ) / ( fp2 - fq2 ) );
			else if ( fq2 > length1 && fp2 < length1 )
				cq = lerp( q2, p2, ( fq2 - length1 ) / ( fq2 - fp2 ) );
			float sp = dot( sub( cp, p1 ), normalA );
			float sq = dot( sub( cq, p1 ), normalA );
			if ( sp <= dist + kLinearSlop || sq <= dist + kLinearSlop )
			{
				ManifoldPoint* mp = m.points + 0;
				mp->anchorA = mulAdd( cp, 0.5f * ( rA - rB - sp ), normalA );
				mp->separation = sp - radius;
				mp->id = makeFeatureId( 0, 0 );
				mp = m.points + 1;
				mp->anchorA = mulAdd( cq, 0.5f * ( rA - rB - sq ), normalA );
				mp->separation = sq - radius;
				mp->id = makeFeatureId( 0, 1 );
				m.pointCount = 2;
			}
		}
		else
		{
			m.normal = neg( normalB );
			V2 cp = p1, cq = q1;
			if ( fp1 < 0.0f && fq1 > 0.0f )
				cp = lerp( p1, q1, ( 0.0f - fp1 ) / ( fq1 - fp1 ) );
			else if ( fq1 < 0.0f && fp1 > 0.0f )
				cq = lerp( q1, p1, ( 0.0f - fq1 ) / ( fp1 - fq1 ) );
			if ( fp1 > length2 && fq1 < length2 )
				cp = lerp( p1, q1, ( fp1 - length2 ) / ( fp1 - fq1 ) );
			else if ( fq1 > length2 && fp1 < length2 )
				cq = lerp( q1, p1, ( fq1 - length2 ) / ( fq1 - fp1 ) );
			float sp = dot( sub( cp, p2 ), normalB );
			float sq = dot( sub( cq, p2 ), normalB );
			if ( sp <= dist + kLinearSlop || sq <= dist + kLinearSlop )
			{
				ManifoldPoint* mp = m.points + 0;
				mp->anchorA = mulAdd( cp, 0.5f * ( rB - rA - sp ), normalB );
				mp->separation = sp - radius;
				mp->id = makeFeatureId( 0, 0 );
				mp = m.points + 1;
				mp->anchorA = mulAdd( cq, 0.5f * ( rB - rA - sq ), normalB );
				mp->separation = sq - radius;
				mp->id = makeFeatureId( 1, 0 );
				m.pointCount = 2;
			}
		}
	}

	if ( m.pointCount == 0 )
	{
		V2 normal = sub( closest2, closest1 );
		if ( dot( normal, normal ) > epsSqr )
			normal = normalize( normal );
		else
			normal = leftPerp( u1 );
		V2 c1 = mulAdd( closest1, rA, normal );
		V2 c2 = mulAdd( closest2, -rB, normal );
		int i1 = f1 == 0.0f ? 0 : 1;
		int i2 = f2 == 0.0f ? 0 : 1;
		m.normal = normal;
		m.points[0].anchorA = lerp( c1, c2, 0.5f );
		m.points[0].separation = sqrtf( distSq ) - radius;
		m.points[0].id = makeFeatureId( i1, i2 );
		m.pointCount = 1;
	}

	m.normal = rotate( xfA.q, m.normal );
	for ( int i = 0; i < m.pointCount; ++i )
	{
		ManifoldPoint* mp = m.points + i;
		mp->anchorA = rotate( xfA.q, add( mp->anchorA, origin ) );
		mp->anchorB = add( mp->anchorA, sub( xfA.p, xfB.p ) );
		mp->point = add( xfA.p, mp->anchorA );
	}
	return m;
}

// Segment distance (Ericson 5.1.9): B2/src/distance.c:33-106
struct SegmentDistance
{
	V2 closest1, closest2;
	float fraction1, fraction2, distanceSquared;
};
F2D_HDF inline SegmentDistance segmentDistance( V2 p1, V2 q1, V2 p2, V2 q2 )
{
	SegmentDistance res;
	V2 d1 = sub( q1, p1 );
	V2 d2 = sub( q2, p2 );
	V2 r = sub( p1, p2 );
	float dd1 = dot( d1, d1 );
	float dd2 = dot( d2, d2 );
	float rd1 = dot( r, d1 );
	float rd2 = dot( r, d2 );
	const float epsSqr = FLT_EPSILON * FLT_EPSILON;
	if ( dd1 < epsSqr || dd2 < epsSqr )
	{
		if ( dd1 >= epsSqr )
		{
			res.fraction1 = clampf( -rd1 / dd1, 0.0f, 1.0f );
			res.fraction2 = 0.0f;
		}
		else if ( dd2 >= epsSqr )
		{
			res.fraction1 = 0.0f;
			res.fraction2 = clampf( rd2 / dd2, 0.0f, 1.0f );
		}
		else
		{
			res.fraction1 = 0.0f;
			res.fraction2 = 0.0f;
		}
	}
	else
	{
		float d12 = dot( d1, d2 );
		float denominator = dd1 * dd2 - d12 * d12;
		float f1 = 0.0f;
		if ( denominator != 0.0f )
			f1 = clampf( ( d12 * rd2 - rd1 * dd2 ) / denominator, 0.0f, 1.0f );
		float f2 = ( d12 * f1 + rd2 ) / dd2;
		if ( f2 < 0.0f )
		{
			f2 = 0.0f;
			f1 = clampf( -rd1 / dd1, 0.0f, 1.0f );
		}
		else if ( f2 > 1.0f )
		{
			f2 = 1.0f;
			f1 = clampf( ( d12 - rd1 ) / dd1, 0.0f, 1.0f );
		}
		res.fraction1 = f1;
		res.fraction2 = f2;
	}
	res.closest1 = mulAdd( p1, res.fraction1, d1 );
	res.closest2 = mulAdd( p2, res.fraction2, d2 );
	res.distanceSquared = distanceSq( res.closest1, res.closest2 );
	return res;
}

// Polygon with a compile-time vertex capacity. Every loop over it is fully unrolled and guarded by `count`, and every
// run-time index goes through pickV (a select chain), so the vertices and normals stay in registers: a per-thread
// array indexed at run time would live in local memory, which at ~1000 resident threads per SM spills out of L1.
template <int N> struct PolyR
{
	V2 v[N];
	V2 n[N];
	float radius;
	int count;
};

template <int N> F2D_HD V2 pickV( const V2 ( &a )[N], int index )
{
	V2 r = a[0];
#pragma unroll
	for ( int k = 1; k < N; ++k )
	{
		if ( index == k )
			r = a[k];
	}
	return r;
}

// Reference-edge / incident-edge clipping: manifold.c:545-677. The four end points, the reference normal and the two
// radii are already selected by the caller (poly1 = reference polygon, poly2 = incident polygon).
F2D_HD Manifold clipSegments( V2 normal, V2 v11, V2 v12, V2 v21, V2 v22, float r1, float r2, int i11, int i12, int i21, int i22, bool flip )
{
	Manifold m = emptyManifold();
	V2 tangent = crossSV( 1.0f, normal );

	float lower1 = 0.0f;
	float upper1 = dot( sub( v12, v11 ), tangent );
	float upper2 = dot( sub( v21, v11 ), tangent );
	float lower2 = dot( sub( v22, v11 ), tangent );
	if ( upper2 < lower1 || upper1 < lower2 )
		return m;

	V2 vLower;
	if ( lower2 < lower1 && upper2 - lower2 > FLT_EPSILON )
		vLower = lerp( v22, v21, ( lower1 - lower2 ) / ( upper2 - lower2 ) );
	else
		vLower = v22;
	V2 vUpper;
	if ( upper2 > upper1 && upper2 - lower2 > FLT_EPSILON )
		vUpper = lerp( v22, v21, ( upper1 - lower2 ) / ( upper2 - lower2 ) );
	else
		vUpper = v21;

	float separationLower = dot( sub( vLower, v11 ), normal );
	float separationUpper = dot( sub( vUpper, v11 ), normal );
	vLower = mulAdd( vLower, 0.5f * ( r1 - r2 - separationLower ), normal );
	vUpper = mulAdd( vUpper, 0.5f * ( r1 - r2 - separationUpper ), normal );
	float radius = r1 + r2;

	if ( flip == false )
	{
		m.normal = normal;
		m.points[0].anchorA = vLower;
		m.points[0].separation = separationLower - radius;
		m.points[0].id = makeFeatureId( i11, i22 );
		m.points[1].anchorA = vUpper;
		m.points[1].separation = separationUpper - radius;
		m.points[1].id = makeFeatureId( i12, i21 );
		m.pointCount = 2;
	}
	else
	{
		m.normal = neg( normal );
		m.points[0].anchorA = vUpper;
		m.points[0].separation = separationUpper - radius;
		m.points[0].id = makeFeatureId( i21, i12 );
		m.points[1].anchorA = vLower;
		m.points[1].separation = separationLower - radius;
		m.points[1].id = makeFeatureId( i22, i11 );
		m.pointCount = 2;
	}
	return m;
}

// manifold.c:545-585: which polygon is the reference one depends on `flip`
template <int NA, int NB>
F2D_HD Manifold clipPolygons( const PolyR<NA>& polyA, const PolyR<NB>& polyB, int edgeA, int edgeB, bool flip )
{
	int nextA = edgeA + 1 < polyA.count ? edgeA + 1 : 0;
	int nextB = edgeB + 1 < polyB.count ? edgeB + 1 : 0;
	V2 a1 = pickV( polyA.v, edgeA ), a2 = pickV( polyA.v, nextA );
	V2 b1 = pickV( polyB.v, edgeB ), b2 = pickV( polyB.v, nextB );
	if ( flip )
		return clipSegments( pickV( polyB.n, edgeB ), b1, b2, a1, a2, polyB.radius, polyA.radius, edgeB, nextB, edgeA, nextA, true );
	return clipSegments( pickV( polyA.n, edgeA ), a1, a2, b1, b2, polyA.radius, polyB.radius, edgeA, nextA, edgeB, nextB, false );
}

// manifold.c:680-716
template <int N1, int N2> F2D_HD float findMaxSeparation( int* edgeIndex, const PolyR<N1>& poly1, const PolyR<N2>& poly2 )
{
	int count1 = poly1.count, count2 = poly2.count;
	int bestIndex = 0;
	float maxSeparation = -FLT_MAX;
#pragma unroll
	for ( int i = 0; i < N1; ++i )
	{
		if ( i < count1 )
		{
			V2 n = poly1.n[i];
			V2 v1 = poly1.v[i];
			float si = FLT_MAX;
#pragma unroll
			for ( int j = 0; j < N2; ++j )
			{
				if ( j < count2 )
				{
					float sij = dot( n, sub( poly2.v[j], v1 ) );
					if ( sij < si )
						si = sij;
				}
			}
			if ( si > maxSeparation )
			{
				maxSeparation = si;
				bestIndex = i;
			}
		}
	}
	*edgeIndex = bestIndex;
	return maxSeparation;
}

// Edge of `poly` whose normal is most anti-parallel to `searchDirection` (manifold.c:790-830)
template <int N> F2D_HD int findIncidentEdge( const PolyR<N>& poly, V2 searchDirection )
{
	int edge = 0;
	float minDot = FLT_MAX;
#pragma unroll
	for ( int i = 0; i < N; ++i )
	{
		if ( i < poly.count )
		{
			float d = dot( searchDirection, poly.n[i] );
			if ( d < minDot )
			{
				minDot = d;
				edge = i;
			}
		}
	}
	return edge;
}

// manifold.c:736-1075 on register-resident polygons. `polygonA`/`polygonB` are in their bodies' frames.
template <int NA, int NB> F2D_HDF inline Manifold collidePolygonsR( const PolyR<NA>& polygonA, Xf xfA, const PolyR<NB>& polygonB, Xf xfB )
{
	V2 origin = polygonA.v[0];
	Xf sfA = { add( xfA.p, rotate( xfA.q, origin ) ), xfA.q };
	Xf xf = invMulXf( sfA, xfB );

	PolyR<NA> localA;
	localA.count = polygonA.count;
	localA.radius = polygonA.radius;
	localA.v[0] = V2{ 0.0f, 0.0f };
	localA.n[0] = polygonA.n[0];
#pragma unroll
	for ( int i = 1; i < NA; ++i )
	{
		localA.v[i] = sub( polygonA.v[i], origin );
		localA.n[i] = polygonA.n[i];
	}
	PolyR<NB> localB;
	localB.count = polygonB.count;
	localB.radius = polygonB.radius;
#pragma unroll
	for ( int i = 0; i < NB; ++i )
	{
		localB.v[i] = xfPoint( xf, polygonB.v[i] );
		localB.n[i] = rotate( xf.q, polygonB.n[i] );
	}

	int edgeA = 0;
	float separationA = findMaxSeparation( &edgeA, localA, localB );
	int edgeB = 0;
	float separationB = findMaxSeparation( &edgeB, localB, localA );
	float radius = localA.radius + localB.radius;
	if ( separationA > kSpeculative + radius || separationB > kSpeculative + radius )
		return emptyManifold();

	bool flip;
	if ( separationA >= separationB )
	{
		flip = false;
		edgeB = findIncidentEdge( localB, pickV( localA.n, edgeA ) );
	}
	else
	{
		flip = true;
		edgeA = findIncidentEdge( localA, pickV( localB.n, edgeB ) );
	}

	Manifold m = emptyManifold();
	if ( separationA > 0.1f * kLinearSlop || separationB > 0.1f * kLinearSlop )
	{
		int i11 = edgeA;
		int i12 = edgeA + 1 < localA.count ? edgeA + 1 : 0;
		int i21 = edgeB;
		int i22 = edgeB + 1 < localB.count ? edgeB + 1 : 0;
		V2 v11 = pickV( localA.v, i11 ), v12 = pickV( localA.v, i12 );
		V2 v21 = pickV( localB.v, i21 ), v22 = pickV( localB.v, i22 );

		SegmentDistance result = segmentDistance( v11, v12, v21, v22 );
		float dist = sqrtf( result.distanceSquared );
		float separation = dist - radius;
		if ( dist - radius > kSpeculative )
			return m;

		m = clipPolygons( localA, localB, edgeA, edgeB, flip );
		float minSeparation = FLT_MAX;
		if ( m.pointCount > 0 )
			minSeparation = minf( minSeparation, m.points[0].separation );
		if ( m.pointCount > 1 )
			minSeparation = minf( minSeparation, m.points[1].separation );

		if ( separation + 0.1f * kLinearSlop < minSeparation )
		{
			// vertex-vertex: pick the pair of end points the segment distance landed on (manifold.c:870-937)
			bool f10 = result.fraction1 == 0.0f, f11 = result.fraction1 == 1.0f;
			bool f20 = result.fraction2 == 0.0f, f21 = result.fraction2 == 1.0f;
			if ( ( f10 || f11 ) && ( f20 || f21 ) )
			{
				V2 va = f10 ? v11 : v12;
				V2 vb = f20 ? v21 : v22;
				int ia = f10 ? i11 : i12;
				int ib = f20 ? i21 : i22;
				V2 normal = sub( vb, va );
				float invDistance = 1.0f / dist;
				normal.x *= invDistance;
				normal.y *= invDistance;
				V2 c1 = mulAdd( va, localA.radius, normal );
				V2 c2 = mulAdd( vb, -localB.radius, normal );
				m.normal = normal;
				m.points[0].anchorA = lerp( c1, c2, 0.5f );
				m.points[0].separation = dist - radius;
				m.points[0].id = makeFeatureId( ia, ib );
				m.pointCount = 1;
			}
		}
	}
	else
	{
		m = clipPolygons( localA, localB, edgeA, edgeB, flip );
	}

	if ( m.pointCount > 0 )
	{
		m.normal = rotate( xfA.q, m.normal );
#pragma unroll
		for ( int i = 0; i < 2; ++i )
		{
			if ( i < m.pointCount )
			{
				ManifoldPoint* mp = m.points + i;
				mp->anchorA = rotate( xfA.q, add( mp->anchorA, origin ) );
				mp->anchorB = add( mp->anchorA, sub( xfA.p, xfB.p ) );
				mp->point = add( xfA.p, mp->anchorA );
			}
		}
	}
	return m;
}

// Loads the first N vertices / normals of a stored polygon (entries beyond `count` are never used)
template <int N> F2D_HD PolyR<N> loadPoly( const Poly& p )
{
	PolyR<N> r;
	r.count = p.count;
	r.radius = p.radius;
#pragma unroll
	for ( int i = 0; i < N; ++i )
	{
		bool live = i < p.count;
		r.v[i] = live ? p.v[i] : V2{ 0.0f, 0.0f };
		r.n[i] = live ? p.n[i] : V2{ 0.0f, 0.0f };
	}
	return r;
}

// manifold.c:16-34 b2MakeCapsule as a two-vertex polygon
template <int N> F2D_HD PolyR<N> capsulePolyR( V2 p1, V2 p2, float radius )
{
	PolyR<N> s;
#pragma unroll
	for ( int i = 0; i < N; ++i )
	{
		s.v[i] = V2{ 0.0f, 0.0f };
		s.n[i] = V2{ 0.0f, 0.0f };
	}
	s.v[0] = p1;
	s.v[1] = p2;
	V2 d = sub( p2, p1 );
	V2 axis = normalize( d );
	V2 normal = rightPerp( axis );
	s.n[0] = normal;
	s.n[1] = neg( normal );
	s.count = 2;
	s.radius = radius;
	return s;
}

// Boxes and capsule-polygons (<= 4 vertices) take the small instantiation; anything larger the 8-vertex one.
F2D_HDF inline Manifold collidePolygons( const Poly& polygonA, Xf xfA, const Poly& polygonB, Xf xfB )
{
	if ( polygonA.count <= 4 && polygonB.count <= 4 )
		return collidePolygonsR( loadPoly<4>( polygonA ), xfA, loadPoly<4>( polygonB ), xfB );
	return collidePolygonsR( loadPoly<8>( polygonA ), xfA, loadPoly<8>( polygonB ), xfB );
}
F2D_HDF inline Manifold collideSegmentAndPolygon( V2 p1, V2 p2, float radius, Xf xfA, const Poly& polygonB, Xf xfB )
{
	if ( polygonB.count <= 4 )
		return collidePolygonsR( capsulePolyR<2>( p1, p2, radius ), xfA, loadPoly<4>( polygonB ), xfB );
	return collidePolygonsR( capsulePolyR<2>( p1, p2, radius ), xfA, loadPoly<8>( polygonB ), xfB );
}
F2D_HDF inline Manifold collidePolygonAndCapsule( const Poly& polygonA, Xf xfA, const Capsule& b, Xf xfB )
{
	if ( polygonA.count <= 4 )
		return collidePolygonsR( loadPoly<4>( polygonA ), xfA, capsulePolyR<2>( b.c1, b.c2, b.radius ), xfB );
	return collidePolygonsR( loadPoly<8>( polygonA ), xfA, capsulePolyR<2>( b.c1, b.c2, b.radius ), xfB );
}

// ------------------------------------------------------------------------------------------------ chain segments
// manifold.c:16-34 b2MakeCapsule as a stored polygon (the chain-segment path indexes vertices dynamically)
F2D_HD Poly capsuleAsPoly( V2 p1, V2 p2, float radius )
{
	Poly shape;
	memset( &shape, 0, sizeof( shape ) );
	shape.v[0] = p1;
	shape.v[1] = p2;
	shape.centroid = lerp( p1, p2, 0.5f );
	V2 axis = normalize( sub( p2, p1 ) );
	V2 normal = rightPerp( axis );
	shape.n[0] = normal;
	shape.n[1] = neg( normal );
	shape.count = 2;
	shape.radius = radius;
	return shape;
}

// manifold.c:1089-1168: one-sided segment with ghost vertices against a circle
F2D_HDF inline Manifold collideChainSegmentAndCircle( const ChainSegment& segmentA, Xf xfA, const Circle& circleB, Xf xfB )
{
	Manifold m = emptyManifold();
	Xf xf = invMulXf( xfA, xfB );
	V2 pB = xfPoint( xf, circleB.center );
	V2 p1 = segmentA.segment.p1, p2 = segmentA.segment.p2;
	V2 e = sub( p2, p1 );
	float offset = dot( rightPerp( e ), sub( pB, p1 ) );
	if ( offset < 0.0f )
		return m; // behind the one-sided segment
	float u = dot( e, sub( p2, pB ) );
	float v = dot( e, sub( pB, p1 ) );
	V2 pA;
	if ( v <= 0.0f )
	{
		// region of p1: the previous segment owns the contact when the circle is on its side
		V2 prevEdge = sub( p1, segmentA.ghost1 );
		float uPrev = dot( prevEdge, sub( pB, p1 ) );
		if ( uPrev <= 0.0f )
			return m;
		pA = p1;
	}
	else if ( u <= 0.0f )
	{
		V2 nextEdge = sub( segmentA.ghost2, p2 );
		float vNext = dot( nextEdge, sub( pB, p2 ) );
		if ( vNext > 0.0f )
			return m;
		pA = p2;
	}
	else
	{
		float ee = dot( e, e );
		pA = V2{ u * p1.x + v * p2.x, u * p1.y + v * p2.y };
		pA = ee > 0.0f ? mulSV( 1.0f / ee, pA ) : p1;
	}
	float distance;
	V2 normal = lengthAndNormalize( &distance, sub( pB, pA ) );
	float radius = circleB.radius;
	float separation = distance - radius;
	if ( separation > kSpeculative )
		return m;
	V2 cA = pA;
	V2 cB = mulAdd( pB, -radius, normal );
	V2 contactPointA = lerp( cA, cB, 0.5f );
	m.normal = rotate( xfA.q, normal );
	ManifoldPoint* mp = m.points + 0;
	mp->anchorA = rotate( xfA.q, contactPointA );
	mp->anchorB = add( mp->anchorA, sub( xfA.p, xfB.p ) );
	mp->point = add( xfA.p, mp->anchorA );
	mp->separation = separation;
	mp->id = 0;
	m.pointCount = 1;
	return m;
}

// manifold.c:1177-1249 b2ClipSegments: segment b clipped to the extent of segment a along the tangent of `normal`
F2D_HDF inline Manifold clipSegments( V2 a1, V2 a2, V2 b1, V2 b2, V2 normal, float ra, float rb, uint16_t id1, uint16_t id2 )
{
	Manifold m = emptyManifold();
	V2 tangent = leftPerp( normal );
	float lower1 = 0.0f;
	float upper1 = dot( sub( a2, a1 ), tangent );
	float upper2 = dot( sub( b1, a1 ), tangent );
	float lower2 = dot( sub( b2, a1 ), tangent );
	if ( upper2 < lower1 || upper1 < lower2 )
		return m;
	V2 vLower;
	if ( lower2 < lower1 && upper2 - lower2 > FLT_EPSILON )
		vLower = lerp( b2, b1, ( lower1 - lower2 ) / ( upper2 - lower2 ) );
	else
		vLower = b2;
	V2 vUpper;
	if ( upper2 > upper1 && upper2 - lower2 > FLT_EPSILON )
		vUpper = lerp( b2, b1, ( upper1 - lower2 ) / ( upper2 - lower2 ) );
	else
		vUpper = b1;
	float separationLower = dot( sub( vLower, a1 ), normal );
	float separationUpper = dot( sub( vUpper, a1 ), normal );
	vLower = mulAdd( vLower, 0.5f * ( ra - rb - separationLower ), normal );
	vUpper = mulAdd( vUpper, 0.5f * ( ra - rb - separationUpper ), normal );
	float radius = ra + rb;
	m.normal = normal;
	m.points[0].anchorA = vLower;
	m.points[0].separation = separationLower - radius;
	m.points[0].id = id1;
	m.points[1].anchorA = vUpper;
	m.points[1].separation = separationUpper - radius;
	m.points[1].id = id2;
	m.pointCount = 2;
	return m;
}

// manifold.c:1251-1304: may a contact normal be used, given the neighbouring segments (smooth collision)?
enum NormalType
{
	kNormalSkip,
	kNormalAdmit,
	kNormalSnap
};
struct ChainSegmentParams
{
	V2 edge1, normal0, normal2;
	bool convex1, convex2;
};
F2D_HD NormalType classifyNormal( const ChainSegmentParams& params, V2 normal )
{
	const float sinTol = 0.01f;
	if ( dot( normal, params.edge1 ) <= 0.0f )
	{
		if ( params.convex1 )
			return cross( normal, params.normal0 ) > sinTol ? kNormalSkip : kNormalAdmit;
		return kNormalSnap;
	}
	if ( params.convex2 )
		return cross( params.normal2, normal ) > sinTol ? kNormalSkip : kNormalAdmit;
	return kNormalSnap;
}

// two-point manifolds of this family are produced in A's local frame: rotate / translate them out (manifold.c:1480-1492)
F2D_HD void chainManifoldToWorld( Manifold& m, V2 localNormal, Xf xfA, Xf xfB )
{
	if ( m.pointCount != 2 )
		return;
	m.normal = rotate( xfA.q, localNormal );
	m.points[0].anchorA = rotate( xfA.q, m.points[0].anchorA );
	m.points[1].anchorA = rotate( xfA.q, m.points[1].anchorA );
	V2 pAB = sub( xfA.p, xfB.p );
	m.points[0].anchorB = add( m.points[0].anchorA, pAB );
	m.points[1].anchorB = add( m.points[1].anchorA, pAB );
	m.points[0].point = add( xfA.p, m.points[0].anchorA );
	m.points[1].point = add( xfA.p, m.points[1].anchorA );
}

// manifold.c:1306-1726
F2D_HDF inline Manifold collideChainSegmentAndPolygon( const ChainSegment& segmentA, Xf xfA, const Poly& polygonB, Xf xfB, SimplexCache* cache )
{
	Manifold m = emptyManifold();
	Xf xf = invMulXf( xfA, xfB );
	V2 centroidB = xfPoint( xf, polygonB.centroid );
	float radiusB = polygonB.radius;
	V2 p1 = segmentA.segment.p1, p2 = segmentA.segment.p2;
	V2 edge1 = normalize( sub( p2, p1 ) );
	ChainSegmentParams smooth;
	smooth.edge1 = edge1;
	const float convexTol = 0.01f;
	V2 edge0 = normalize( sub( p1, segmentA.ghost1 ) );
	smooth.normal0 = rightPerp( edge0 );
	smooth.convex1 = cross( edge0, edge1 ) >= convexTol;
	V2 edge2 = normalize( sub( segmentA.ghost2, p2 ) );
	smooth.normal2 = rightPerp( edge2 );
	smooth.convex2 = cross( edge1, edge2 ) >= convexTol;
	V2 normal1 = rightPerp( edge1 );
	bool behind1 = dot( normal1, sub( centroidB, p1 ) ) < 0.0f;
	bool behind0 = true, behind2 = true;
	if ( smooth.convex1 )
		behind0 = dot( smooth.normal0, sub( centroidB, p1 ) ) < 0.0f;
	if ( smooth.convex2 )
		behind2 = dot( smooth.normal2, sub( centroidB, p2 ) ) < 0.0f;
	if ( behind1 && behind0 && behind2 )
		return m;

	const int count = polygonB.count;
	V2 vertices[kMaxPolyVerts], normals[kMaxPolyVerts];
	for ( int i = 0; i < count; ++i )
	{
		vertices[i] = xfPoint( xf, polygonB.v[i] );
		normals[i] = rotate( xf.q, polygonB.n[i] );
	}
	const Xf identity = { { 0.0f, 0.0f }, { 1.0f, 0.0f } };
	DistanceOutput output = shapeDistance( makeProxy( &segmentA.segment.p1, 2, 0.0f ), makeProxy( vertices, count, 0.0f ), identity, identity,
										   false, cache );
	if ( output.distance > radiusB + kSpeculative )
		return m;

	V2 n0 = smooth.convex1 ? smooth.normal0 : normal1;
	V2 n2 = smooth.convex2 ? smooth.normal2 : normal1;
	int incidentIndex = -1;
	int incidentNormal = -1;
	// a polygon edge chosen as reference face must not face away from the neighbouring segment it hangs over
	auto neighbourRejects = [&]( V2 n, V2 a1 ) {
		float dot1 = dot( n, sub( p1, a1 ) );
		float dot2 = dot( n, sub( p2, a1 ) );
		if ( dot1 < dot2 )
			return dot( n0, n ) < dot( normal1, n );
		return dot( n2, n ) < dot( normal1, n );
	};

	if ( behind1 == false && output.distance > 0.1f * kLinearSlop )
	{
		// separated: the GJK witness features decide
		if ( cache->count == 1 )
		{
			V2 pA = output.pointA, pB = output.pointB;
			V2 normal = normalize( sub( pB, pA ) );
			NormalType type = classifyNormal( smooth, normal );
			if ( type == kNormalSkip )
				return m;
			if ( type == kNormalAdmit )
			{
				m.normal = rotate( xfA.q, normal );
				ManifoldPoint* cp = m.points + 0;
				cp->anchorA = rotate( xfA.q, pA );
				cp->anchorB = add( cp->anchorA, sub( xfA.p, xfB.p ) );
				cp->point = add( xfA.p, cp->anchorA );
				cp->separation = output.distance - radiusB;
				cp->id = makeFeatureId( cache->indexA[0], cache->indexB[0] );
				m.pointCount = 1;
				return m;
			}
			incidentIndex = cache->indexB[0];
		}
		else
		{
			int ia1 = cache->indexA[0], ia2 = cache->indexA[1];
			int ib1 = cache->indexB[0], ib2 = cache->indexB[1];
			if ( ia1 == ia2 )
			{
				// one segment vertex against a polygon edge: the better aligned of the two edge normals is the reference
				V2 normalB = sub( output.pointA, output.pointB );
				float dot1 = dot( normalB, normals[ib1] );
				float dot2 = dot( normalB, normals[ib2] );
				int ib = dot1 > dot2 ? ib1 : ib2;
				normalB = normals[ib];
				NormalType type = classifyNormal( smooth, neg( normalB ) );
				if ( type == kNormalSkip )
					return m;
				if ( type == kNormalAdmit )
				{
					ib1 = ib;
					ib2 = ib < count - 1 ? ib + 1 : 0;
					V2 b1 = vertices[ib1], b2 = vertices[ib2];
					if ( neighbourRejects( normalB, b1 ) )
						return m;
					m = clipSegments( b1, b2, p1, p2, normalB, radiusB, 0.0f, makeFeatureId( ib1, 1 ), makeFeatureId( ib2, 0 ) );
					chainManifoldToWorld( m, neg( normalB ), xfA, xfB );
					return m;
				}
				incidentNormal = ib;
			}
			else
			{
				// segment edge against one polygon vertex region
				float dot1 = dot( normal1, sub( vertices[ib1], p1 ) );
				float dot2 = dot( normal1, sub( vertices[ib2], p2 ) );
				incidentIndex = dot1 < dot2 ? ib1 : ib2;
			}
		}
	}
	else
	{
		// overlapping (or the polygon centre is behind the segment): separating-axis search
		float edgeSeparation = FLT_MAX;
		for ( int i = 0; i < count; ++i )
		{
			float s = dot( normal1, sub( vertices[i], p1 ) );
			if ( s < edgeSeparation )
			{
				edgeSeparation = s;
				incidentIndex = i;
			}
		}
		if ( smooth.convex1 )
		{
			float s0 = FLT_MAX;
			for ( int i = 0; i < count; ++i )
			{
				float s = dot( smooth.normal0, sub( vertices[i], p1 ) );
				if ( s < s0 )
					s0 = s;
			}
			if ( s0 > edgeSeparation )
			{
				edgeSeparation = s0;
				incidentIndex = -1;
			}
		}
		if ( smooth.convex2 )
		{
			float s2 = FLT_MAX;
			for ( int i = 0; i < count; ++i )
			{
				float s = dot( smooth.normal2, sub( vertices[i], p2 ) );
				if ( s < s2 )
					s2 = s;
			}
			if ( s2 > edgeSeparation )
			{
				edgeSeparation = s2;
				incidentIndex = -1;
			}
		}
		float polygonSeparation = -FLT_MAX;
		int referenceIndex = -1;
		for ( int i = 0; i < count; ++i )
		{
			V2 n = normals[i];
			if ( classifyNormal( smooth, neg( n ) ) != kNormalAdmit )
				continue;
			V2 p = vertices[i];
			float s = minf( dot( n, sub( p2, p ) ), dot( n, sub( p1, p ) ) );
			if ( s > polygonSeparation )
			{
				polygonSeparation = s;
				referenceIndex = i;
			}
		}
		if ( polygonSeparation > edgeSeparation )
		{
			int ia1 = referenceIndex;
			int ia2 = ia1 < count - 1 ? ia1 + 1 : 0;
			V2 a1 = vertices[ia1], a2 = vertices[ia2];
			V2 n = normals[ia1];
			if ( neighbourRejects( n, a1 ) )
				return m;
			m = clipSegments( a1, a2, p1, p2, normals[ia1], radiusB, 0.0f, makeFeatureId( ia1, 1 ), makeFeatureId( ia2, 0 ) );
			chainManifoldToWorld( m, neg( normals[ia1] ), xfA, xfB );
			return m;
		}
		if ( incidentIndex == -1 )
			return m;
	}

	// the segment is the reference face: pick the incident polygon edge and clip it
	V2 b1, b2;
	int ib1, ib2;
	if ( incidentNormal != -1 )
	{
		ib1 = incidentNormal;
		ib2 = ib1 < count - 1 ? ib1 + 1 : 0;
		b1 = vertices[ib1];
		b2 = vertices[ib2];
	}
	else
	{
		int i2 = incidentIndex;
		int i1 = i2 > 0 ? i2 - 1 : count - 1;
		float d1 = dot( normal1, normals[i1] );
		float d2 = dot( normal1, normals[i2] );
		if ( d1 < d2 )
		{
			ib1 = i1;
			ib2 = i2;
		}
		else
		{
			ib1 = i2;
			ib2 = i2 < count - 1 ? i2 + 1 : 0;
		}
		b1 = vertices[ib1];
		b2 = vertices[ib2];
	}
	m = clipSegments( p1, p2, b1, b2, normal1, 0.0f, radiusB, makeFeatureId( 0, ib2 ), makeFeatureId( 1, ib1 ) );
	chainManifoldToWorld( m, m.normal, xfA, xfB );
	return m;
}

// manifold.c:1170-1175
F2D_HDF inline Manifold collideChainSegmentAndCapsule( const ChainSegment& segmentA, Xf xfA, const Capsule& capsuleB, Xf xfB, SimplexCache* cache )
{
	Poly polyB = capsuleAsPoly( capsuleB.c1, capsuleB.c2, capsuleB.radius );
	return collideChainSegmentAndPolygon( segmentA, xfA, polyB, xfB, cache );
}

// Dispatch on the (primary-ordered) shape-type pair: B2/src/contact.c:166-184 registers.
// Returns false when the pair has no manifold function (e.g. segment vs segment).
F2D_HD bool pairHasManifold( int typeA, int typeB )
{
	if ( typeA == kChainSegment || typeB == kChainSegment )
	{
		int other = typeA == kChainSegment ? typeB : typeA;
		return other == kCircle || other == kCapsule || other == kPolygon;
	}
	if ( typeA == kSegment && typeB == kSegment )
		return false;
	return true;
}
// primary[type1][type2] == true when (type1,type2) is the registered order (contact.c:151-164)
F2D_HD bool pairIsPrimary( int t1, int t2 )
{
	if ( t1 == t2 )
		return true;
	// registered primaries: capsule-circle, polygon-circle, polygon-capsule, segment-{circle,capsule,polygon},
	// chainSegment-{circle,capsule,polygon}
	if ( t1 == kCapsule && t2 == kCircle )
		return true;
	if ( t1 == kPolygon && ( t2 == kCircle || t2 == kCapsule ) )
		return true;
	if ( t1 == kSegment && ( t2 == kCircle || t2 == kCapsule || t2 == kPolygon ) )
		return true;
	if ( t1 == kChainSegment && ( t2 == kCircle || t2 == kCapsule || t2 == kPolygon ) )
		return true;
	return false;
}

F2D_HDF inline Manifold computeManifold( World* w, const Shape& a, Xf xfA, const Shape& b, Xf xfB, SimplexCache* cache )
{
	switch ( a.type )
	{
		case kCircle:
			return collideCircles( a.circle, xfA, b.circle, xfB );
		case kCapsule:
			if ( b.type == kCircle )
				return collideCapsuleAndCircle( a.capsule, xfA, b.circle, xfB );
			return collideCapsules( a.capsule, xfA, b.capsule, xfB );
		case kPolygon:
			if ( b.type == kCircle )
				return collidePolygonAndCircle( a.polygon, xfA, b.circle, xfB );
			if ( b.type == kCapsule )
				return collidePolygonAndCapsule( a.polygon, xfA, b.capsule, xfB ); // manifold.c:538-542
			return collidePolygons( a.polygon, xfA, b.polygon, xfB );
		case kSegment:
		{
			if ( b.type == kCircle )
			{
				Capsule capA = { a.segment.p1, a.segment.p2, 0.0f }; // manifold.c:1077-1081
				return collideCapsuleAndCircle( capA, xfA, b.circle, xfB );
			}
			if ( b.type == kCapsule )
			{
				Capsule capA = { a.segment.p1, a.segment.p2, 0.0f }; // manifold.c:532-536
				return collideCapsules( capA, xfA, b.capsule, xfB );
			}
			return collideSegmentAndPolygon( a.segment.p1, a.segment.p2, 0.0f, xfA, b.polygon, xfB ); // manifold.c:1083-1087
		}
		case kChainSegment:
			if ( b.type == kCircle )
				return collideChainSegmentAndCircle( a.chainSegment, xfA, b.circle, xfB );
			if ( b.type == kCapsule )
				return collideChainSegmentAndCapsule( a.chainSegment, xfA, b.capsule, xfB, cache );
			return collideChainSegmentAndPolygon( a.chainSegment, xfA, b.polygon, xfB, cache );
		default:
			setError( w, kErrUnsupported, __LINE__ );
			return emptyManifold();
	}
}

} // namespace f2d
