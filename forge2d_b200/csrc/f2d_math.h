// forge2d_b200 — fp32 math with the reference's exact expression trees.
//
// Every function here reproduces the operation ORDER of the corresponding inline in
// B2/include/box2d/math_functions.h (cited per function) so that, compiled without FMA contraction
// (nvcc -fmad=false / g++ -ffp-contract=off), results are bit-identical to the reference's x86-64 SSE2 build.
// Min/max use the reference's `a<b?a:b` / `a>b?a:b` forms (NaN and signed-zero behaviour differs from fminf/fmaxf).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined( __CUDACC__ )
#define F2D_HD __host__ __device__ __forceinline__
#define F2D_HDN __host__ __device__ __noinline__
#define F2D_HDF __host__ __device__
// Parts of the step with large locals (traversal stacks, candidate lists, GJK / TOI state): real calls in the kernels that
// give a world 128 registers per thread - the frames of the parts then overlap instead of adding up (5.2 -> 2.1 KB of
// local memory per thread) and the one-block bench2d frame gets 3 % faster - inlined in the batch kernels, where a call
// at 64-72 registers spills around itself (measured: -3 % batch throughput).
#if defined( F2D_SMALL_TEAM_KERNELS )
#define F2D_HDC __host__ __device__
#else
#define F2D_HDC __host__ __device__ __noinline__
#endif
#else
#define F2D_HD inline
#define F2D_HDN
#define F2D_HDF
#define F2D_HDC
#endif

namespace f2d
{

// Non-binding hint: start fetching the 128-byte line that holds `p` (device only)
F2D_HD void prefetchLine( const void* p )
{
#if defined( __CUDA_ARCH__ )
	asm volatile( "prefetch.global.L1 [%0];" ::"l"( p ) );
#else
	(void)p;
#endif
}

// Same hint, but only as far as L2: used one loop iteration ahead, where a line parked in the (small, shared by eight
// worlds) L1 would be gone before it is needed, while an L2 hit still replaces a DRAM round trip.
F2D_HD void prefetchL2( const void* p )
{
#if defined( __CUDA_ARCH__ )
	asm volatile( "prefetch.global.L2 [%0];" ::"l"( p ) );
#else
	(void)p;
#endif
}

constexpr float kPi = 3.14159265359f;		  // math_functions.h:79
constexpr float kLinearSlop = 0.005f;		  // constants.h:21 (lengthUnitsPerMeter == 1, forge2d never changes it)
constexpr float kSpeculative = 4.0f * 0.005f; // constants.h:34
constexpr float kAabbMargin = 0.05f;		  // constants.h:39
constexpr float kHuge = 100000.0f;			  // constants.h:10
constexpr float kTimeToSleep = 0.5f;		  // constants.h:47
constexpr float kMaxRotation = 0.25f * kPi;	  // constants.h:31
constexpr int kColorCount = 12;				  // constants.h:17
constexpr int kOverflow = kColorCount - 1;
constexpr int kNull = -1;
constexpr int kMaxPolyVerts = 8;

struct V2
{
	float x, y;
};
struct Rot
{
	float c, s;
};
struct Xf
{
	V2 p;
	Rot q;
};
struct Box
{
	V2 lo, hi;
};

// A 16-byte chunk moved with ONE memory instruction (LDG.128 / STG.128 on the device). The records of the world image
// are laid out in such chunks (f2d_types.h), and every request to L2 costs the same whether it carries 4 or 16 bytes,
// so the hot gathers read and write whole chunks instead of one field at a time. `p` must be 16-byte aligned.
#if defined( __GNUC__ ) || defined( __CUDACC__ )
#define F2D_MAY_ALIAS __attribute__( ( may_alias ) )
#else
#define F2D_MAY_ALIAS
#endif
struct alignas( 16 ) F2D_MAY_ALIAS Q4
{
	float x, y, z, w;
};
F2D_HD Q4 load16( const void* p ) { return *static_cast<const Q4*>( p ); }
F2D_HD void store16( void* p, Q4 v ) { *static_cast<Q4*>( p ) = v; }

F2D_HD uint32_t floatBits( float f )
{
#if defined( __CUDA_ARCH__ )
	return __float_as_uint( f );
#else
	uint32_t u;
	__builtin_memcpy( &u, &f, 4 );
	return u;
#endif
}
F2D_HD float floatFromBits( uint32_t u )
{
#if defined( __CUDA_ARCH__ )
	return __uint_as_float( u );
#else
	float f;
	__builtin_memcpy( &f, &u, 4 );
	return f;
#endif
}
F2D_HD float minf( float a, float b ) { return a < b ? a : b; }				 // :126
F2D_HD float maxf( float a, float b ) { return a > b ? a : b; }				 // :132
F2D_HD float absf( float a ) { return a < 0 ? -a : a; }						 // :138
F2D_HD float clampf( float a, float lo, float hi ) { return a < lo ? lo : ( a > hi ? hi : a ); } // :144
F2D_HD int mini( int a, int b ) { return a < b ? a : b; }
F2D_HD int maxi( int a, int b ) { return a > b ? a : b; }

F2D_HD V2 v2( float x, float y ) { return V2{ x, y }; }
F2D_HD float dot( V2 a, V2 b ) { return a.x * b.x + a.y * b.y; }			 // :157
F2D_HD float cross( V2 a, V2 b ) { return a.x * b.y - a.y * b.x; }			 // :163
F2D_HD V2 crossVS( V2 v, float s ) { return V2{ s * v.y, -s * v.x }; }		 // :169
F2D_HD V2 crossSV( float s, V2 v ) { return V2{ -s * v.y, s * v.x }; }		 // :175
F2D_HD V2 leftPerp( V2 v ) { return V2{ -v.y, v.x }; }						 // :181
F2D_HD V2 rightPerp( V2 v ) { return V2{ v.y, -v.x }; }						 // :187
F2D_HD V2 add( V2 a, V2 b ) { return V2{ a.x + b.x, a.y + b.y }; }
F2D_HD V2 sub( V2 a, V2 b ) { return V2{ a.x - b.x, a.y - b.y }; }
F2D_HD V2 neg( V2 a ) { return V2{ -a.x, -a.y }; }
F2D_HD V2 lerp( V2 a, V2 b, float t ) { return V2{ ( 1.0f - t ) * a.x + t * b.x, ( 1.0f - t ) * a.y + t * b.y }; } // :215
F2D_HD V2 mulSV( float s, V2 v ) { return V2{ s * v.x, s * v.y }; }
F2D_HD V2 mulAdd( V2 a, float s, V2 b ) { return V2{ a.x + s * b.x, a.y + s * b.y }; } // :233
F2D_HD V2 mulSub( V2 a, float s, V2 b ) { return V2{ a.x - s * b.x, a.y - s * b.y }; } // :239
F2D_HD V2 vmin( V2 a, V2 b ) { return V2{ minf( a.x, b.x ), minf( a.y, b.y ) }; }
F2D_HD V2 vmax( V2 a, V2 b ) { return V2{ maxf( a.x, b.x ), maxf( a.y, b.y ) }; }
F2D_HD float length( V2 v ) { return sqrtf( v.x * v.x + v.y * v.y ); }
F2D_HD float lengthSq( V2 v ) { return v.x * v.x + v.y * v.y; }
F2D_HD float distanceSq( V2 a, V2 b )
{
	V2 c = { b.x - a.x, b.y - a.y };
	return c.x * c.x + c.y * c.y;
}
F2D_HD float distance( V2 a, V2 b )
{
	float dx = b.x - a.x, dy = b.y - a.y;
	return sqrtf( dx * dx + dy * dy );
}
// :294 b2Normalize
F2D_HD V2 normalize( V2 v )
{
	float len = sqrtf( v.x * v.x + v.y * v.y );
	if ( len < FLT_EPSILON )
		return V2{ 0.0f, 0.0f };
	float inv = 1.0f / len;
	return V2{ inv * v.x, inv * v.y };
}
// :315 b2GetLengthAndNormalize
F2D_HD V2 lengthAndNormalize( float* len, V2 v )
{
	*len = sqrtf( v.x * v.x + v.y * v.y );
	if ( *len < FLT_EPSILON )
		return V2{ 0.0f, 0.0f };
	float inv = 1.0f / *len;
	return V2{ inv * v.x, inv * v.y };
}
// :332 b2NormalizeRot
F2D_HD Rot normalizeRot( Rot q )
{
	float mag = sqrtf( q.s * q.s + q.c * q.c );
	float inv = mag > 0.0f ? 1.0f / mag : 0.0f;
	return Rot{ q.c * inv, q.s * inv };
}
// :343 b2IntegrateRotation
F2D_HD Rot integrateRot( Rot q1, float dAngle )
{
	Rot q2 = { q1.c - dAngle * q1.s, q1.s + dAngle * q1.c };
	float mag = sqrtf( q2.s * q2.s + q2.c * q2.c );
	float inv = mag > 0.0f ? 1.0f / mag : 0.0f;
	return Rot{ q2.c * inv, q2.s * inv };
}
// :390 b2NLerp
F2D_HD Rot nlerp( Rot q1, Rot q2, float t )
{
	float omt = 1.0f - t;
	Rot q = { omt * q1.c + t * q2.c, omt * q1.s + t * q2.s };
	float mag = sqrtf( q.s * q.s + q.c * q.c );
	float inv = mag > 0.0f ? 1.0f / mag : 0.0f;
	return Rot{ q.c * inv, q.s * inv };
}
// :440 b2MulRot
F2D_HD Rot mulRot( Rot q, Rot r )
{
	Rot o;
	o.s = q.s * r.c + q.c * r.s;
	o.c = q.c * r.c - q.s * r.s;
	return o;
}
// :452 b2InvMulRot
F2D_HD Rot invMulRot( Rot q, Rot r )
{
	Rot o;
	o.s = q.c * r.s - q.s * r.c;
	o.c = q.c * r.c + q.s * r.s;
	return o;
}
F2D_HD V2 rotate( Rot q, V2 v ) { return V2{ q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y }; }	   // :488
F2D_HD V2 invRotate( Rot q, V2 v ) { return V2{ q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y }; } // :494
// :500 b2TransformPoint (note the parenthesised rotation before the translation)
F2D_HD V2 xfPoint( Xf t, V2 p )
{
	float x = ( t.q.c * p.x - t.q.s * p.y ) + t.p.x;
	float y = ( t.q.s * p.x + t.q.c * p.y ) + t.p.y;
	return V2{ x, y };
}
// :532 b2InvMulTransforms
F2D_HD Xf invMulXf( Xf A, Xf B )
{
	Xf C;
	C.q = invMulRot( A.q, B.q );
	C.p = invRotate( A.q, sub( B.p, A.p ) );
	return C;
}
// :569 b2Solve22 on a column-major 2x2 (cx, cy)
struct M22
{
	V2 cx, cy;
};
F2D_HD V2 solve22( M22 A, V2 b )
{
	float a11 = A.cx.x, a12 = A.cy.x, a21 = A.cx.y, a22 = A.cy.y;
	float det = a11 * a22 - a12 * a21;
	if ( det != 0.0f )
		det = 1.0f / det;
	return V2{ det * ( a22 * b.x - a12 * b.y ), det * ( a11 * b.y - a21 * b.x ) };
}
F2D_HD V2 mulMV( M22 A, V2 v ) { return V2{ A.cx.x * v.x + A.cy.x * v.y, A.cx.y * v.x + A.cy.y * v.y }; }
// b2GetInverse22 (math_functions.h)
F2D_HD M22 inverse22( M22 A )
{
	float a = A.cx.x, b = A.cy.x, c = A.cx.y, d = A.cy.y;
	float det = a * d - b * c;
	if ( det != 0.0f )
		det = 1.0f / det;
	M22 B = { { det * d, -det * c }, { -det * b, det * a } };
	return B;
}

// "This value is needed HERE": pins the load that produces it before the next branch. Without it the compiler sinks a
// load below every branch that does not use it, and a gather written as one round of independent loads becomes a chain
// of rounds, one per early-out.
#if defined( __CUDA_ARCH__ )
#define F2D_ISSUE_F( x ) asm volatile( "" ::"f"( x ) )
#define F2D_ISSUE_I( x ) asm volatile( "" ::"r"( x ) )
#else
#define F2D_ISSUE_F( x ) (void)( x )
#define F2D_ISSUE_I( x ) (void)( x )
#endif

// AABB helpers: math_functions.h:582-640, B2/src/aabb.h
// (bitwise on purpose, here and in boxOverlaps: with && / || the compiler fetches the boxes piece by piece, one load
// and one branch per comparison - a memory round trip each when the boxes come straight from records)
F2D_HD bool boxContains( Box a, Box b )
{
	return ( a.lo.x <= b.lo.x ) & ( a.lo.y <= b.lo.y ) & ( b.hi.x <= a.hi.x ) & ( b.hi.y <= a.hi.y );
}
F2D_HD V2 boxCenter( Box a ) { return V2{ 0.5f * ( a.lo.x + a.hi.x ), 0.5f * ( a.lo.y + a.hi.y ) }; }
F2D_HD Box boxUnion( Box a, Box b )
{
	Box c;
	c.lo.x = minf( a.lo.x, b.lo.x );
	c.lo.y = minf( a.lo.y, b.lo.y );
	c.hi.x = maxf( a.hi.x, b.hi.x );
	c.hi.y = maxf( a.hi.y, b.hi.y );
	return c;
}
F2D_HD bool boxOverlaps( Box a, Box b ) { return ( ( b.lo.x > a.hi.x ) | ( b.lo.y > a.hi.y ) | ( a.lo.x > b.hi.x ) | ( a.lo.y > b.hi.y ) ) == false; }
F2D_HD float boxPerimeter( Box a )
{
	float wx = a.hi.x - a.lo.x;
	float wy = a.hi.y - a.lo.y;
	return 2.0f * ( wx + wy );
}
// B2/src/aabb.h b2EnlargeAABB
F2D_HD bool boxEnlarge( Box* a, Box b )
{
	bool changed = false;
	if ( b.lo.x < a->lo.x )
	{
		a->lo.x = b.lo.x;
		changed = true;
	}
	if ( b.lo.y < a->lo.y )
	{
		a->lo.y = b.lo.y;
		changed = true;
	}
	if ( a->hi.x < b.hi.x )
	{
		a->hi.x = b.hi.x;
		changed = true;
	}
	if ( a->hi.y < b.hi.y )
	{
		a->hi.y = b.hi.y;
		changed = true;
	}
	return changed;
}

// Deterministic trig, never libm: B2/src/math_functions.c:61-148
F2D_HD float atan2Poly( float y, float x )
{
	if ( x == 0.0f && y == 0.0f )
		return 0.0f;
	float ax = absf( x );
	float ay = absf( y );
	float mx = maxf( ay, ax );
	float mn = minf( ay, ax );
	float a = mn / mx;

	// Minimax polynomial approximation to atan(a) on [0,1]
	float s = a * a;
	float c = s * a;
	float q = s * s;
	float r = 0.024840285f * q + 0.18681418f;
	float t = -0.094097948f * q - 0.33213072f;
	r = r * s + t;
	r = r * c + a;

	// Map to full circle
	if ( ay > ax )
		r = 1.57079637f - r;
	if ( x < 0 )
		r = 3.14159274f - r;
	if ( y < 0 )
		r = -r;
	return r;
}
F2D_HD float relativeAngle( Rot b, Rot a )
{
	float s = b.s * a.c - b.c * a.s;
	float c = b.c * a.c + b.s * a.s;
	return atan2Poly( s, c );
}
F2D_HD float unwindAngle( float radians ) { return remainderf( radians, 2.0f * kPi ); }
// b2MakeRot( radians ) = b2ComputeCosSin: rational approximations of cosine / sine, then normalised
// (math_functions.c:107-148, math_functions.h:370-374)
F2D_HD Rot makeRotAngle( float radians )
{
	float x = unwindAngle( radians );
	float pi2 = kPi * kPi;
	float c;
	if ( x < -0.5f * kPi )
	{
		float y = x + kPi;
		float y2 = y * y;
		c = -( pi2 - 4.0f * y2 ) / ( pi2 + y2 );
	}
	else if ( x > 0.5f * kPi )
	{
		float y = x - kPi;
		float y2 = y * y;
		c = -( pi2 - 4.0f * y2 ) / ( pi2 + y2 );
	}
	else
	{
		float y2 = x * x;
		c = ( pi2 - 4.0f * y2 ) / ( pi2 + y2 );
	}
	float s;
	if ( x < 0.0f )
	{
		float y = x + kPi;
		s = -16.0f * y * ( kPi - y ) / ( 5.0f * pi2 - 4.0f * y * ( kPi - y ) );
	}
	else
	{
		s = 16.0f * x * ( kPi - x ) / ( 5.0f * pi2 - 4.0f * x * ( kPi - x ) );
	}
	float mag = sqrtf( s * s + c * c );
	float invMag = mag > 0.0 ? 1.0f / mag : 0.0f;
	return Rot{ c * invMag, s * invMag };
}

// Softness: B2/src/solver.h:140-182
struct Soft
{
	float biasRate, massScale, impulseScale;
};
F2D_HD Soft makeSoft( float hertz, float zeta, float h )
{
	if ( hertz == 0.0f )
		return Soft{ 0.0f, 0.0f, 0.0f };
	float omega = 2.0f * kPi * hertz;
	float a1 = 2.0f * zeta + h * omega;
	float a2 = h * omega * a1;
	float a3 = 1.0f / ( 1.0f + a2 );
	return Soft{ omega / a1, a2 * a3, a3 };
}

} // namespace f2d
