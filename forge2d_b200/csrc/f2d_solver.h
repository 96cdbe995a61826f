// forge2d_b200 — soft-step contact solver on structure-of-arrays constraints.
//
// One constraint slot per touching contact; one thread per slot inside a colour (no two slots of a colour share a
// non-static body, so the colour is embarrassingly parallel). Per-slot arithmetic is the per-lane arithmetic of
// the reference's SSE2 kernels, operation for operation (mul and add never fused):
//   prepare   B2/src/contact_solver.c:1438-1669      warm start :1671-1727      solve/relax :1729-1967
//   restitution :1969-2076                            store      :2078-2120
// The overflow colour (index 11) is solved by rank 0 in array order with the reference's *scalar* formulas,
// which differ subtly: contact_solver.c:24-509.
#pragma once
#include "f2d_team.h"
#include "f2d_world.h"

namespace f2d
{

struct ConView
{
	float* base;
	int stride;
	F2D_HD float& f( int field, int slot ) const { return base[field * stride + slot]; }
	F2D_HD int32_t& i( int field, int slot ) const { return reinterpret_cast<int32_t*>( base )[field * stride + slot]; }
};

F2D_HD ConView conView( World* w ) { return ConView{ ptr( w, w->cons ), w->consStride }; }

// Gathered solver body: identity for a static body (contact_solver.c:1280-1329)
struct SolverBody
{
	V2 v;
	float w;
	float flagBits; // BodyState::flags as loaded: travels with v and w so that they go back in one 16-byte store
	V2 dp;
	Rot dq;
};
// The 32-byte state moves as two 16-byte chunks (v, w, flags | dp, dq). Inside a colour a non-static body belongs to
// exactly one constraint, so the thread that gathers a state is the only one that writes it.
F2D_HD SolverBody gatherBody( const BodyState* states, int index )
{
	SolverBody b;
	if ( index == kNull )
	{
		b.v = V2{ 0.0f, 0.0f };
		b.w = 0.0f;
		b.flagBits = 0.0f;
		b.dp = V2{ 0.0f, 0.0f };
		b.dq = Rot{ 1.0f, 0.0f };
	}
	else
	{
		const BodyState& s = states[index];
		const Q4 lo = load16( &s.v ), hi = load16( &s.dp );
		b.v = V2{ lo.x, lo.y };
		b.w = lo.z;
		b.flagBits = lo.w;
		b.dp = V2{ hi.x, hi.y };
		b.dq = Rot{ hi.z, hi.w };
	}
	return b;
}
// warm start and restitution do not read dp / dq
F2D_HD SolverBody gatherBodyVelocity( const BodyState* states, int index )
{
	SolverBody b;
	b.dp = V2{ 0.0f, 0.0f };
	b.dq = Rot{ 1.0f, 0.0f };
	if ( index == kNull )
	{
		b.v = V2{ 0.0f, 0.0f };
		b.w = 0.0f;
		b.flagBits = 0.0f;
	}
	else
	{
		const Q4 lo = load16( &states[index].v );
		b.v = V2{ lo.x, lo.y };
		b.w = lo.z;
		b.flagBits = lo.w;
	}
	return b;
}
F2D_HD void scatterBody( BodyState* states, int index, const SolverBody& b )
{
	if ( index != kNull )
		store16( &states[index].v, Q4{ b.v.x, b.v.y, b.w, b.flagBits } );
}

// contact_solver.c:1455-1616 (per lane) == :46-155 (overflow): identical formulas
F2D_HDF inline void prepareContactSlot( World* w, const ConView& c, int slot, int contactId, const BodyState* states, float warmStartScale )
{
	const ContactSim& sim = ptr( w, w->contactSims )[contactId];
	// eight 16-byte chunks of the record: manifold M0-M4, then indices / material / masses
	const Q4 m0 = load16( &sim.manifold.pointCount ), m1 = load16( &sim.manifold.normalImpulse0 ), m2 = load16( &sim.manifold.normal ),
			 m3 = load16( &sim.manifold.anchorA0 ), m4 = load16( &sim.manifold.anchorA1 );
	const Q4 k9 = load16( &sim.bodySimIndexA ), k10 = load16( &sim.invMassA ), k11 = load16( &sim.rollingResistance );
	const int pointCount = (int)floatBits( m0.x );
	const int indexA = (int)floatBits( k9.x ), indexB = (int)floatBits( k9.y );
	V2 vA = { 0.0f, 0.0f };
	float wA = 0.0f;
	float mA = k10.x, iA = k10.y;
	if ( indexA != kNull )
	{
		const Q4 vw = load16( &states[indexA].v );
		vA = V2{ vw.x, vw.y };
		wA = vw.z;
	}
	V2 vB = { 0.0f, 0.0f };
	float wB = 0.0f;
	float mB = k10.z, iB = k10.w;
	if ( indexB != kNull )
	{
		const Q4 vw = load16( &states[indexB].v );
		vB = V2{ vw.x, vw.y };
		wB = vw.z;
	}
	const Soft soft = ( indexA == kNull || indexB == kNull ) ? w->step.staticSoftness : w->step.contactSoftness;
	// (stores only from here on: every load above is in flight before the first of them)
	c.i( cfIndexA, slot ) = indexA;
	c.i( cfIndexB, slot ) = indexB;
	c.i( cfPointCount, slot ) = pointCount;
	c.f( cfInvMassA, slot ) = mA;
	c.f( cfInvMassB, slot ) = mB;
	c.f( cfInvIA, slot ) = iA;
	c.f( cfInvIB, slot ) = iB;
	{
		float k = iA + iB;
		c.f( cfRollingMass, slot ) = k > 0.0f ? 1.0f / k : 0.0f;
	}
	V2 normal = { m2.x, m2.y };
	c.f( cfNormalX, slot ) = normal.x;
	c.f( cfNormalY, slot ) = normal.y;
	c.f( cfFriction, slot ) = k9.z;
	c.f( cfTangentSpeed, slot ) = k11.y;
	c.f( cfRestitution, slot ) = k9.w;
	c.f( cfRollingResistance, slot ) = k11.x;
	c.f( cfRollingImpulse, slot ) = warmStartScale * m0.z;
	c.f( cfBiasRate, slot ) = soft.biasRate;
	c.f( cfMassScale, slot ) = soft.massScale;
	c.f( cfImpulseScale, slot ) = soft.impulseScale;

	V2 tangent = rightPerp( normal );
	{
		V2 rA = { m3.x, m3.y }, rB = { m3.z, m3.w };
		c.f( cfAnchorA1X, slot ) = rA.x;
		c.f( cfAnchorA1Y, slot ) = rA.y;
		c.f( cfAnchorB1X, slot ) = rB.x;
		c.f( cfAnchorB1Y, slot ) = rB.y;
		c.f( cfBaseSeparation1, slot ) = m2.z - dot( sub( rB, rA ), normal );
		c.f( cfNormalImpulse1, slot ) = warmStartScale * m1.x;
		c.f( cfTangentImpulse1, slot ) = warmStartScale * m1.y;
		c.f( cfTotalNormalImpulse1, slot ) = 0.0f;
		float rnA = cross( rA, normal );
		float rnB = cross( rB, normal );
		float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
		c.f( cfNormalMass1, slot ) = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
		float rtA = cross( rA, tangent );
		float rtB = cross( rB, tangent );
		float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
		c.f( cfTangentMass1, slot ) = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
		V2 vrA = add( vA, crossSV( wA, rA ) );
		V2 vrB = add( vB, crossSV( wB, rB ) );
		c.f( cfRelativeVelocity1, slot ) = dot( normal, sub( vrB, vrA ) );
	}
	if ( pointCount == 2 )
	{
		V2 rA = { m4.x, m4.y }, rB = { m4.z, m4.w };
		c.f( cfAnchorA2X, slot ) = rA.x;
		c.f( cfAnchorA2Y, slot ) = rA.y;
		c.f( cfAnchorB2X, slot ) = rB.x;
		c.f( cfAnchorB2Y, slot ) = rB.y;
		c.f( cfBaseSeparation2, slot ) = m2.w - dot( sub( rB, rA ), normal );
		c.f( cfNormalImpulse2, slot ) = warmStartScale * m1.z;
		c.f( cfTangentImpulse2, slot ) = warmStartScale * m1.w;
		c.f( cfTotalNormalImpulse2, slot ) = 0.0f;
		float rnA = cross( rA, normal );
		float rnB = cross( rB, normal );
		float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
		c.f( cfNormalMass2, slot ) = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
		float rtA = cross( rA, tangent );
		float rtB = cross( rB, tangent );
		float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
		c.f( cfTangentMass2, slot ) = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
		V2 vrA = add( vA, crossSV( wA, rA ) );
		V2 vrB = add( vB, crossSV( wB, rB ) );
		c.f( cfRelativeVelocity2, slot ) = dot( normal, sub( vrB, vrA ) );
	}
	else
	{
		c.f( cfBaseSeparation2, slot ) = 0.0f;
		c.f( cfNormalImpulse2, slot ) = 0.0f;
		c.f( cfTangentImpulse2, slot ) = 0.0f;
		c.f( cfTotalNormalImpulse2, slot ) = 0.0f;
		c.f( cfAnchorA2X, slot ) = 0.0f;
		c.f( cfAnchorA2Y, slot ) = 0.0f;
		c.f( cfAnchorB2X, slot ) = 0.0f;
		c.f( cfAnchorB2Y, slot ) = 0.0f;
		c.f( cfNormalMass2, slot ) = 0.0f;
		c.f( cfTangentMass2, slot ) = 0.0f;
		c.f( cfRelativeVelocity2, slot ) = 0.0f;
	}
}

// Applies impulse P at anchors: the shared "apply contact impulse" tail of contact_solver.c:1808-1818 etc.
F2D_HD void applyImpulse( SolverBody& bA, SolverBody& bB, float mA, float iA, float mB, float iB, V2 rA, V2 rB, float Px, float Py )
{
	bA.v.x = bA.v.x - mA * Px;
	bA.v.y = bA.v.y - mA * Py;
	bA.w = bA.w - iA * ( rA.x * Py - rA.y * Px );
	bB.v.x = bB.v.x + mB * Px;
	bB.v.y = bB.v.y + mB * Py;
	bB.w = bB.w + iB * ( rB.x * Py - rB.y * Px );
}

// contact_solver.c:1671-1727
F2D_HDF inline void warmStartSlot( const ConView& c, int slot, BodyState* states )
{
	int indexA = c.i( cfIndexA, slot ), indexB = c.i( cfIndexB, slot );
	SolverBody bA = gatherBodyVelocity( states, indexA );
	SolverBody bB = gatherBodyVelocity( states, indexB );
	float nx = c.f( cfNormalX, slot ), ny = c.f( cfNormalY, slot );
	float tx = ny;
	float ty = 0.0f - nx;
	float mA = c.f( cfInvMassA, slot ), iA = c.f( cfInvIA, slot );
	float mB = c.f( cfInvMassB, slot ), iB = c.f( cfInvIB, slot );
	{
		V2 rA = { c.f( cfAnchorA1X, slot ), c.f( cfAnchorA1Y, slot ) };
		V2 rB = { c.f( cfAnchorB1X, slot ), c.f( cfAnchorB1Y, slot ) };
		float ni = c.f( cfNormalImpulse1, slot ), ti = c.f( cfTangentImpulse1, slot );
		float Px = ni * nx + ti * tx;
		float Py = ni * ny + ti * ty;
		bA.w = bA.w - iA * ( rA.x * Py - rA.y * Px );
		bA.v.x = bA.v.x - mA * Px;
		bA.v.y = bA.v.y - mA * Py;
		bB.w = bB.w + iB * ( rB.x * Py - rB.y * Px );
		bB.v.x = bB.v.x + mB * Px;
		bB.v.y = bB.v.y + mB * Py;
	}
	{
		V2 rA = { c.f( cfAnchorA2X, slot ), c.f( cfAnchorA2Y, slot ) };
		V2 rB = { c.f( cfAnchorB2X, slot ), c.f( cfAnchorB2Y, slot ) };
		float ni = c.f( cfNormalImpulse2, slot ), ti = c.f( cfTangentImpulse2, slot );
		float Px = ni * nx + ti * tx;
		float Py = ni * ny + ti * ty;
		bA.w = bA.w - iA * ( rA.x * Py - rA.y * Px );
		bA.v.x = bA.v.x - mA * Px;
		bA.v.y = bA.v.y - mA * Py;
		bB.w = bB.w + iB * ( rB.x * Py - rB.y * Px );
		bB.v.x = bB.v.x + mB * Px;
		bB.v.y = bB.v.y + mB * Py;
	}
	float ri = c.f( cfRollingImpulse, slot );
	bA.w = bA.w - iA * ri;
	bB.w = bB.w + iB * ri;
	scatterBody( states, indexA, bA );
	scatterBody( states, indexB, bB );
}

// One non-penetration row: contact_solver.c:1764-1819 (point 1) / :1821-1871 (point 2)
F2D_HD float solveNormalRow( SolverBody& bA, SolverBody& bB, V2 dp, V2 rA, V2 rB, float nx, float ny, float baseSeparation,
							 float normalMass, float& normalImpulse, float& totalNormalImpulse, float biasRate, float massScale,
							 float impulseScale, float inv_h, float minBiasVel, float mA, float iA, float mB, float iB )
{
	V2 rsA = rotate( bA.dq, rA );
	V2 rsB = rotate( bB.dq, rB );
	float dsx = dp.x + ( rsB.x - rsA.x );
	float dsy = dp.y + ( rsB.y - rsA.y );
	float s = ( nx * dsx + ny * dsy ) + baseSeparation;
	bool speculative = s > 0.0f;
	float specBias = s * inv_h;
	float softBias = maxf( biasRate * s, minBiasVel );
	float bias = speculative ? specBias : softBias;
	float pointMassScale = speculative ? 1.0f : massScale;
	float pointImpulseScale = speculative ? 0.0f : impulseScale;

	float dvx = ( bB.v.x - bB.w * rB.y ) - ( bA.v.x - bA.w * rA.y );
	float dvy = ( bB.v.y + bB.w * rB.x ) - ( bA.v.y + bA.w * rA.x );
	float vn = dvx * nx + dvy * ny;
	float negImpulse = normalMass * ( pointMassScale * ( vn + bias ) ) + pointImpulseScale * normalImpulse;
	float newImpulse = maxf( normalImpulse - negImpulse, 0.0f );
	float impulse = newImpulse - normalImpulse;
	normalImpulse = newImpulse;
	totalNormalImpulse = totalNormalImpulse + newImpulse;
	float Px = impulse * nx;
	float Py = impulse * ny;
	applyImpulse( bA, bB, mA, iA, mB, iB, rA, rB, Px, Py );
	return newImpulse;
}

// One friction row: contact_solver.c:1876-1911
F2D_HD void solveFrictionRow( SolverBody& bA, SolverBody& bB, V2 rA, V2 rB, float tx, float ty, float tangentSpeed, float tangentMass,
							  float friction, float normalImpulse, float& tangentImpulse, float mA, float iA, float mB, float iB )
{
	float dvx = ( bB.v.x - bB.w * rB.y ) - ( bA.v.x - bA.w * rA.y );
	float dvy = ( bB.v.y + bB.w * rB.x ) - ( bA.v.y + bA.w * rA.x );
	float vt = dvx * tx + dvy * ty;
	vt = vt - tangentSpeed;
	float negImpulse = tangentMass * vt;
	float maxFriction = friction * normalImpulse;
	float newImpulse = tangentImpulse - negImpulse;
	newImpulse = maxf( 0.0f - maxFriction, minf( newImpulse, maxFriction ) );
	float impulse = newImpulse - tangentImpulse;
	tangentImpulse = newImpulse;
	float Px = impulse * tx;
	float Py = impulse * ty;
	applyImpulse( bA, bB, mA, iA, mB, iB, rA, rB, Px, Py );
}

// contact_solver.c:1729-1967
F2D_HDF inline void solveSlot( const ConView& c, int slot, BodyState* states, bool useBias, float inv_h, float contactSpeed )
{
	int indexA = c.i( cfIndexA, slot ), indexB = c.i( cfIndexB, slot );
	SolverBody bA = gatherBody( states, indexA );
	SolverBody bB = gatherBody( states, indexB );
	float biasRate, massScale, impulseScale;
	if ( useBias )
	{
		biasRate = c.f( cfBiasRate, slot );
		massScale = c.f( cfMassScale, slot );
		impulseScale = c.f( cfImpulseScale, slot );
	}
	else
	{
		biasRate = 0.0f;
		massScale = 1.0f;
		impulseScale = 0.0f;
	}
	float minBiasVel = -contactSpeed;
	float totalNormalImpulse = 0.0f;
	V2 dp = { bB.dp.x - bA.dp.x, bB.dp.y - bA.dp.y };
	// Every field is loaded before the first store: the compiler cannot move a load above a store to another array of
	// the same view (they might alias), so loads placed where the reference reads them come in three rounds - one per
	// group of stores - each a memory round trip in the middle of the pass.
	const float nx = c.f( cfNormalX, slot ), ny = c.f( cfNormalY, slot );
	const float mA = c.f( cfInvMassA, slot ), iA = c.f( cfInvIA, slot );
	const float mB = c.f( cfInvMassB, slot ), iB = c.f( cfInvIB, slot );
	const V2 rA1 = { c.f( cfAnchorA1X, slot ), c.f( cfAnchorA1Y, slot ) };
	const V2 rB1 = { c.f( cfAnchorB1X, slot ), c.f( cfAnchorB1Y, slot ) };
	const V2 rA2 = { c.f( cfAnchorA2X, slot ), c.f( cfAnchorA2Y, slot ) };
	const V2 rB2 = { c.f( cfAnchorB2X, slot ), c.f( cfAnchorB2Y, slot ) };
	float ni1 = c.f( cfNormalImpulse1, slot ), tni1 = c.f( cfTotalNormalImpulse1, slot );
	float ni2 = c.f( cfNormalImpulse2, slot ), tni2 = c.f( cfTotalNormalImpulse2, slot );
	const float baseSeparation1 = c.f( cfBaseSeparation1, slot ), normalMass1 = c.f( cfNormalMass1, slot );
	const float baseSeparation2 = c.f( cfBaseSeparation2, slot ), normalMass2 = c.f( cfNormalMass2, slot );
	const float friction = c.f( cfFriction, slot ), tangentSpeed = c.f( cfTangentSpeed, slot );
	float ti1 = c.f( cfTangentImpulse1, slot ), ti2 = c.f( cfTangentImpulse2, slot );
	const float tangentMass1 = c.f( cfTangentMass1, slot ), tangentMass2 = c.f( cfTangentMass2, slot );
	const float rollingMass = c.f( cfRollingMass, slot ), rollingResistance = c.f( cfRollingResistance, slot );
	const float lambda = c.f( cfRollingImpulse, slot );

	float new1 = solveNormalRow( bA, bB, dp, rA1, rB1, nx, ny, baseSeparation1, normalMass1, ni1, tni1, biasRate, massScale, impulseScale,
								 inv_h, minBiasVel, mA, iA, mB, iB );
	totalNormalImpulse = totalNormalImpulse + new1;
	float new2 = solveNormalRow( bA, bB, dp, rA2, rB2, nx, ny, baseSeparation2, normalMass2, ni2, tni2, biasRate, massScale, impulseScale,
								 inv_h, minBiasVel, mA, iA, mB, iB );
	totalNormalImpulse = totalNormalImpulse + new2;

	float tx = ny;
	float ty = 0.0f - nx;
	solveFrictionRow( bA, bB, rA1, rB1, tx, ty, tangentSpeed, tangentMass1, friction, ni1, ti1, mA, iA, mB, iB );
	solveFrictionRow( bA, bB, rA2, rB2, tx, ty, tangentSpeed, tangentMass2, friction, ni2, ti2, mA, iA, mB, iB );

	float newLambda;
	{
		// rolling resistance: contact_solver.c:1950-1960
		float deltaLambda = rollingMass * ( bA.w - bB.w );
		float maxLambda = rollingResistance * totalNormalImpulse;
		float nb = -maxLambda; // sign-bit flip, as the reference's xor with -0.0f
		float sum = lambda + deltaLambda;
		newLambda = maxf( nb, minf( sum, maxLambda ) );
		deltaLambda = newLambda - lambda;
		bA.w = bA.w - iA * deltaLambda;
		bB.w = bB.w + iB * deltaLambda;
	}
	c.f( cfNormalImpulse1, slot ) = ni1;
	c.f( cfTotalNormalImpulse1, slot ) = tni1;
	c.f( cfNormalImpulse2, slot ) = ni2;
	c.f( cfTotalNormalImpulse2, slot ) = tni2;
	c.f( cfTangentImpulse1, slot ) = ti1;
	c.f( cfTangentImpulse2, slot ) = ti2;
	c.f( cfRollingImpulse, slot ) = newLambda;
	scatterBody( states, indexA, bA );
	scatterBody( states, indexB, bB );
}

// contact_solver.c:1969-2076. `groupHasRestitution`: any of the 4 slots of this slot's SIMD lane group has
// restitution != 0 (the reference skips whole groups, :1982-1986).
F2D_HDF inline void restitutionSlot( const ConView& c, int slot, BodyState* states, float threshold )
{
	float restitution = c.f( cfRestitution, slot );
	bool restitutionMask = restitution == 0.0f;
	int indexA = c.i( cfIndexA, slot ), indexB = c.i( cfIndexB, slot );
	SolverBody bA = gatherBodyVelocity( states, indexA );
	SolverBody bB = gatherBodyVelocity( states, indexB );
	float nx = c.f( cfNormalX, slot ), ny = c.f( cfNormalY, slot );
	float mA = c.f( cfInvMassA, slot ), iA = c.f( cfInvIA, slot );
	float mB = c.f( cfInvMassB, slot ), iB = c.f( cfInvIB, slot );
	for ( int p = 0; p < 2; ++p )
	{
		float relativeVelocity = c.f( p == 0 ? cfRelativeVelocity1 : cfRelativeVelocity2, slot );
		float totalNormalImpulse = c.f( p == 0 ? cfTotalNormalImpulse1 : cfTotalNormalImpulse2, slot );
		bool mask1 = ( relativeVelocity + threshold ) > 0.0f;
		bool mask2 = totalNormalImpulse == 0.0f;
		bool mask = mask1 || mask2 || restitutionMask;
		float mass = mask ? 0.0f : c.f( p == 0 ? cfNormalMass1 : cfNormalMass2, slot );
		V2 rA = { c.f( p == 0 ? cfAnchorA1X : cfAnchorA2X, slot ), c.f( p == 0 ? cfAnchorA1Y : cfAnchorA2Y, slot ) };
		V2 rB = { c.f( p == 0 ? cfAnchorB1X : cfAnchorB2X, slot ), c.f( p == 0 ? cfAnchorB1Y : cfAnchorB2Y, slot ) };
		float dvx = ( bB.v.x - bB.w * rB.y ) - ( bA.v.x - bA.w * rA.y );
		float dvy = ( bB.v.y + bB.w * rB.x ) - ( bA.v.y + bA.w * rA.x );
		float vn = dvx * nx + dvy * ny;
		float negImpulse = mass * ( vn + restitution * relativeVelocity );
		const int impulseField = p == 0 ? cfNormalImpulse1 : cfNormalImpulse2;
		float normalImpulse = c.f( impulseField, slot );
		float newImpulse = maxf( normalImpulse - negImpulse, 0.0f );
		float impulse = newImpulse - normalImpulse;
		c.f( impulseField, slot ) = newImpulse;
		float Px = impulse * nx;
		float Py = impulse * ny;
		applyImpulse( bA, bB, mA, iA, mB, iB, rA, rB, Px, Py );
	}
	scatterBody( states, indexA, bA );
	scatterBody( states, indexB, bB );
}

// contact_solver.c:2078-2120 (and :480-509 for overflow: only the first pointCount points are written there)
F2D_HD void storeSlot( World* w, const ConView& c, int slot, int contactId, bool overflow )
{
	StoredManifold& m = ptr( w, w->contactSims )[contactId].manifold;
	if ( overflow == false )
	{
		// both points, whatever the point count: two whole chunks (M1: impulses, M6: total impulses and normal velocities);
		// all nine loads before the first store
		const float rolling = c.f( cfRollingImpulse, slot );
		const Q4 impulses = { c.f( cfNormalImpulse1, slot ), c.f( cfTangentImpulse1, slot ), c.f( cfNormalImpulse2, slot ),
							  c.f( cfTangentImpulse2, slot ) };
		const Q4 totals = { c.f( cfTotalNormalImpulse1, slot ), c.f( cfTotalNormalImpulse2, slot ), c.f( cfRelativeVelocity1, slot ),
							c.f( cfRelativeVelocity2, slot ) };
		m.rollingImpulse = rolling;
		store16( &m.normalImpulse0, impulses );
		store16( &m.totalNormalImpulse0, totals );
		return;
	}
	m.rollingImpulse = c.f( cfRollingImpulse, slot );
	int n = m.pointCount;
	if ( n > 0 )
	{
		m.normalImpulse0 = c.f( cfNormalImpulse1, slot );
		m.tangentImpulse0 = c.f( cfTangentImpulse1, slot );
		m.totalNormalImpulse0 = c.f( cfTotalNormalImpulse1, slot );
		m.normalVelocity0 = c.f( cfRelativeVelocity1, slot );
	}
	if ( n > 1 )
	{
		m.normalImpulse1 = c.f( cfNormalImpulse2, slot );
		m.tangentImpulse1 = c.f( cfTangentImpulse2, slot );
		m.totalNormalImpulse1 = c.f( cfTotalNormalImpulse2, slot );
		m.normalVelocity1 = c.f( cfRelativeVelocity2, slot );
	}
}

// ---- overflow colour: scalar formulas of contact_solver.c:160-478, executed in array order by one thread ----
struct OverflowBody
{
	BodyState* state;
	BodyState dummy;
};

F2D_HDF inline void overflowWarmStart( const ConView& c, int slot, BodyState* states )
{
	BodyState dummy = identityState();
	int indexA = c.i( cfIndexA, slot ), indexB = c.i( cfIndexB, slot );
	BodyState* sA = indexA == kNull ? &dummy : states + indexA;
	BodyState* sB = indexB == kNull ? &dummy : states + indexB;
	V2 vA = sA->v;
	float wA = sA->w;
	V2 vB = sB->v;
	float wB = sB->w;
	float mA = c.f( cfInvMassA, slot ), iA = c.f( cfInvIA, slot );
	float mB = c.f( cfInvMassB, slot ), iB = c.f( cfInvIB, slot );
	V2 normal = { c.f( cfNormalX, slot ), c.f( cfNormalY, slot ) };
	V2 tangent = rightPerp( normal );
	int pointCount = c.i( cfPointCount, slot );
	for ( int j = 0; j < pointCount; ++j )
	{
		V2 rA = { c.f( j == 0 ? cfAnchorA1X : cfAnchorA2X, slot ), c.f( j == 0 ? cfAnchorA1Y : cfAnchorA2Y, slot ) };
		V2 rB = { c.f( j == 0 ? cfAnchorB1X : cfAnchorB2X, slot ), c.f( j == 0 ? cfAnchorB1Y : cfAnchorB2Y, slot ) };
		float ni = c.f( j == 0 ? cfNormalImpulse1 : cfNormalImpulse2, slot );
		float ti = c.f( j == 0 ? cfTangentImpulse1 : cfTangentImpulse2, slot );
		V2 P = add( mulSV( ni, normal ), mulSV( ti, tangent ) );
		wA -= iA * cross( rA, P );
		vA = mulAdd( vA, -mA, P );
		wB += iB * cross( rB, P );
		vB = mulAdd( vB, mB, P );
	}
	float ri = c.f( cfRollingImpulse, slot );
	wA -= iA * ri;
	wB += iB * ri;
	sA->v = vA;
	sA->w = wA;
	sB->v = vB;
	sB->w = wB;
}

F2D_HDF inline void overflowSolve( const ConView& c, int slot, BodyState* states, bool useBias, float inv_h, float pushout )
{
	BodyState dummy = identityState();
	int indexA = c.i( cfIndexA, slot ), indexB = c.i( cfIndexB, slot );
	float mA = c.f( cfInvMassA, slot ), iA = c.f( cfInvIA, slot );
	float mB = c.f( cfInvMassB, slot ), iB = c.f( cfInvIB, slot );
	BodyState* sA = indexA == kNull ? &dummy : states + indexA;
	V2 vA = sA->v;
	float wA = sA->w;
	Rot dqA = sA->dq;
	BodyState* sB = indexB == kNull ? &dummy : states + indexB;
	V2 vB = sB->v;
	float wB = sB->w;
	Rot dqB = sB->dq;
	V2 dp = sub( sB->dp, sA->dp );
	V2 normal = { c.f( cfNormalX, slot ), c.f( cfNormalY, slot ) };
	V2 tangent = rightPerp( normal );
	float friction = c.f( cfFriction, slot );
	float sBiasRate = c.f( cfBiasRate, slot ), sMassScale = c.f( cfMassScale, slot ), sImpulseScale = c.f( cfImpulseScale, slot );
	int pointCount = c.i( cfPointCount, slot );
	float totalNormalImpulse = 0.0f;

	for ( int j = 0; j < pointCount; ++j )
	{
		V2 rA = { c.f( j == 0 ? cfAnchorA1X : cfAnchorA2X, slot ), c.f( j == 0 ? cfAnchorA1Y : cfAnchorA2Y, slot ) };
		V2 rB = { c.f( j == 0 ? cfAnchorB1X : cfAnchorB2X, slot ), c.f( j == 0 ? cfAnchorB1Y : cfAnchorB2Y, slot ) };
		float& normalImpulse = c.f( j == 0 ? cfNormalImpulse1 : cfNormalImpulse2, slot );
		float& totalNI = c.f( j == 0 ? cfTotalNormalImpulse1 : cfTotalNormalImpulse2, slot );
		float normalMass = c.f( j == 0 ? cfNormalMass1 : cfNormalMass2, slot );
		float baseSeparation = c.f( j == 0 ? cfBaseSeparation1 : cfBaseSeparation2, slot );

		V2 ds = add( dp, sub( rotate( dqB, rB ), rotate( dqA, rA ) ) );
		float s = baseSeparation + dot( ds, normal );
		float velocityBias = 0.0f;
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( s > 0.0f )
		{
			velocityBias = s * inv_h;
		}
		else if ( useBias )
		{
			velocityBias = maxf( sBiasRate * s, -pushout );
			massScale = sMassScale;
			impulseScale = sImpulseScale;
		}
		V2 vrA = add( vA, crossSV( wA, rA ) );
		V2 vrB = add( vB, crossSV( wB, rB ) );
		float vn = dot( sub( vrB, vrA ), normal );
		float impulse = -normalMass * massScale * ( vn + velocityBias ) - impulseScale * normalImpulse;
		float newImpulse = maxf( normalImpulse + impulse, 0.0f );
		impulse = newImpulse - normalImpulse;
		normalImpulse = newImpulse;
		totalNI += newImpulse;
		totalNormalImpulse += newImpulse;
		V2 P = mulSV( impulse, normal );
		vA = mulSub( vA, mA, P );
		wA -= iA * cross( rA, P );
		vB = mulAdd( vB, mB, P );
		wB += iB * cross( rB, P );
	}
	float tangentSpeed = c.f( cfTangentSpeed, slot );
	for ( int j = 0; j < pointCount; ++j )
	{
		V2 rA = { c.f( j == 0 ? cfAnchorA1X : cfAnchorA2X, slot ), c.f( j == 0 ? cfAnchorA1Y : cfAnchorA2Y, slot ) };
		V2 rB = { c.f( j == 0 ? cfAnchorB1X : cfAnchorB2X, slot ), c.f( j == 0 ? cfAnchorB1Y : cfAnchorB2Y, slot ) };
		float normalImpulse = c.f( j == 0 ? cfNormalImpulse1 : cfNormalImpulse2, slot );
		float& tangentImpulse = c.f( j == 0 ? cfTangentImpulse1 : cfTangentImpulse2, slot );
		float tangentMass = c.f( j == 0 ? cfTangentMass1 : cfTangentMass2, slot );
		V2 vrB = add( vB, crossSV( wB, rB ) );
		V2 vrA = add( vA, crossSV( wA, rA ) );
		float vt = dot( sub( vrB, vrA ), tangent ) - tangentSpeed;
		float impulse = tangentMass * ( -vt );
		float maxFriction = friction * normalImpulse;
		float newImpulse = clampf( tangentImpulse + impulse, -maxFriction, maxFriction );
		impulse = newImpulse - tangentImpulse;
		tangentImpulse = newImpulse;
		V2 P = mulSV( impulse, tangent );
		vA = mulSub( vA, mA, P );
		wA -= iA * cross( rA, P );
		vB = mulAdd( vB, mB, P );
		wB += iB * cross( rB, P );
	}
	{
		float deltaLambda = -c.f( cfRollingMass, slot ) * ( wB - wA );
		float lambda = c.f( cfRollingImpulse, slot );
		float maxLambda = c.f( cfRollingResistance, slot ) * totalNormalImpulse;
		float newLambda = clampf( lambda + deltaLambda, -maxLambda, maxLambda );
		c.f( cfRollingImpulse, slot ) = newLambda;
		deltaLambda = newLambda - lambda;
		wA -= iA * deltaLambda;
		wB += iB * deltaLambda;
	}
	sA->v = vA;
	sA->w = wA;
	sB->v = vB;
	sB->w = wB;
}

F2D_HDF inline void overflowRestitution( const ConView& c, int slot, BodyState* states, float threshold )
{
	float restitution = c.f( cfRestitution, slot );
	if ( restitution == 0.0f )
		return;
	BodyState dummy = identityState();
	int indexA = c.i( cfIndexA, slot ), indexB = c.i( cfIndexB, slot );
	float mA = c.f( cfInvMassA, slot ), iA = c.f( cfInvIA, slot );
	float mB = c.f( cfInvMassB, slot ), iB = c.f( cfInvIB, slot );
	BodyState* sA = indexA == kNull ? &dummy : states + indexA;
	V2 vA = sA->v;
	float wA = sA->w;
	BodyState* sB = indexB == kNull ? &dummy : states + indexB;
	V2 vB = sB->v;
	float wB = sB->w;
	V2 normal = { c.f( cfNormalX, slot ), c.f( cfNormalY, slot ) };
	int pointCount = c.i( cfPointCount, slot );
	for ( int j = 0; j < pointCount; ++j )
	{
		float relativeVelocity = c.f( j == 0 ? cfRelativeVelocity1 : cfRelativeVelocity2, slot );
		float& totalNI = c.f( j == 0 ? cfTotalNormalImpulse1 : cfTotalNormalImpulse2, slot );
		if ( relativeVelocity > -threshold || totalNI == 0.0f )
			continue;
		V2 rA = { c.f( j == 0 ? cfAnchorA1X : cfAnchorA2X, slot ), c.f( j == 0 ? cfAnchorA1Y : cfAnchorA2Y, slot ) };
		V2 rB = { c.f( j == 0 ? cfAnchorB1X : cfAnchorB2X, slot ), c.f( j == 0 ? cfAnchorB1Y : cfAnchorB2Y, slot ) };
		float& normalImpulse = c.f( j == 0 ? cfNormalImpulse1 : cfNormalImpulse2, slot );
		float normalMass = c.f( j == 0 ? cfNormalMass1 : cfNormalMass2, slot );
		V2 vrB = add( vB, crossSV( wB, rB ) );
		V2 vrA = add( vA, crossSV( wA, rA ) );
		float vn = dot( sub( vrB, vrA ), normal );
		float impulse = -normalMass * ( vn + restitution * relativeVelocity );
		float newImpulse = maxf( normalImpulse + impulse, 0.0f );
		impulse = newImpulse - normalImpulse;
		normalImpulse = newImpulse;
		totalNI += impulse;
		V2 P = mulSV( impulse, normal );
		vA = mulSub( vA, mA, P );
		wA -= iA * cross( rA, P );
		vB = mulAdd( vB, mB, P );
		wB += iB * cross( rB, P );
	}
	sA->v = vA;
	sA->w = wA;
	sB->v = vB;
	sB->w = wB;
}

} // namespace f2d
