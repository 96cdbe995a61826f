// forge2d_b200 — joint constraints (prepare / warm start / solve), one thread per joint inside a colour.
// Scalar formulas of B2/src/joint.c:1348-1470 and the per-type files (revolute_joint.c, distance_joint.c,
// prismatic_joint.c, weld_joint.c, wheel_joint.c, motor_joint.c, mouse_joint.c): every floating-point expression keeps
// the reference's operand order, so a joint stepped here carries bit-identical impulses and body velocities.
#pragma once
#include "f2d_team.h"
#include "f2d_world.h"

namespace f2d
{

struct JointBodies
{
	BodyState* a;
	BodyState* b;
};

// revolute_joint.c:202-259
F2D_HDF inline void prepareRevolute( World* w, JointSim& base )
{
	const Body* bodies = ptr( w, w->bodies );
	const BodySim* sims = ptr( w, w->sims );
	const Body& bodyA = bodies[base.bodyIdA];
	const Body& bodyB = bodies[base.bodyIdB];
	const BodySim& simA = sims[base.bodyIdA];
	const BodySim& simB = sims[base.bodyIdB];
	float mA = simA.invMass, iA = simA.invInertia;
	float mB = simB.invMass, iB = simB.invInertia;
	base.invMassA = mA;
	base.invMassB = mB;
	base.invIA = iA;
	base.invIB = iB;
	RevoluteJointData& j = base.revolute;
	j.indexA = bodyA.setIndex == kAwakeSet ? bodyA.localIndex : kNull;
	j.indexB = bodyB.setIndex == kAwakeSet ? bodyB.localIndex : kNull;
	j.anchorA = rotate( simA.transform.q, sub( base.localOriginAnchorA, simA.localCenter ) );
	j.anchorB = rotate( simB.transform.q, sub( base.localOriginAnchorB, simB.localCenter ) );
	j.deltaCenter = sub( simB.center, simA.center );
	j.deltaAngle = relativeAngle( simB.transform.q, simA.transform.q );
	float k = iA + iB;
	j.axialMass = k > 0.0f ? 1.0f / k : 0.0f;
	j.springSoftness = makeSoft( j.hertz, j.dampingRatio, w->step.h );
	if ( w->step.enableWarmStarting == 0 )
	{
		j.linearImpulse = V2{ 0.0f, 0.0f };
		j.springImpulse = 0.0f;
		j.motorImpulse = 0.0f;
		j.lowerImpulse = 0.0f;
		j.upperImpulse = 0.0f;
	}
}

// revolute_joint.c:261-287
F2D_HDF inline void warmStartRevolute( JointSim& base, BodyState* states )
{
	float mA = base.invMassA, mB = base.invMassB, iA = base.invIA, iB = base.invIB;
	BodyState dummy = identityState();
	RevoluteJointData& j = base.revolute;
	BodyState* sA = j.indexA == kNull ? &dummy : states + j.indexA;
	BodyState* sB = j.indexB == kNull ? &dummy : states + j.indexB;
	// both states and every joint field first, the stores last: a load cannot move above a store that might alias it
	// (sA->w after sA->v, body B after body A), and each such load is a memory round trip of its own
	const Q4 aLo = load16( &sA->v ), aHi = load16( &sA->dp ), bLo = load16( &sB->v ), bHi = load16( &sB->dp );
	const V2 anchorA = j.anchorA, anchorB = j.anchorB, linearImpulse = j.linearImpulse;
	const float axialImpulse = j.springImpulse + j.motorImpulse + j.lowerImpulse - j.upperImpulse;
	V2 rA = rotate( Rot{ aHi.z, aHi.w }, anchorA );
	V2 rB = rotate( Rot{ bHi.z, bHi.w }, anchorB );
	const V2 vA = mulSub( V2{ aLo.x, aLo.y }, mA, linearImpulse );
	const float wA = aLo.z - iA * ( cross( rA, linearImpulse ) + axialImpulse );
	const V2 vB = mulAdd( V2{ bLo.x, bLo.y }, mB, linearImpulse );
	const float wB = bLo.z + iB * ( cross( rB, linearImpulse ) + axialImpulse );
	store16( &sA->v, Q4{ vA.x, vA.y, wA, aLo.w } );
	store16( &sB->v, Q4{ vB.x, vB.y, wB, bLo.w } );
}

// revolute_joint.c:289-465
F2D_HDF inline void solveRevolute( JointSim& base, BodyState* states, bool useBias, float h, float inv_h )
{
	float mA = base.invMassA, mB = base.invMassB, iA = base.invIA, iB = base.invIB;
	BodyState dummy = identityState();
	RevoluteJointData& j = base.revolute;
	BodyState* sA = j.indexA == kNull ? &dummy : states + j.indexA;
	BodyState* sB = j.indexB == kNull ? &dummy : states + j.indexB;
	// one round of loads for the states and for what every revolute joint needs (the point constraint below), pinned
	// before the first branch: left where the reference reads them they follow one another, a round trip per branch
	const Q4 aLo = load16( &sA->v ), aHi = load16( &sA->dp ), bLo = load16( &sB->v ), bHi = load16( &sB->dp );
	const bool enableSpring = j.enableSpring, enableMotor = j.enableMotor, enableLimit = j.enableLimit;
	const V2 anchorA = j.anchorA, anchorB = j.anchorB, deltaCenter = j.deltaCenter, linearImpulse0 = j.linearImpulse;
	const Soft softness = base.constraintSoftness;
	F2D_ISSUE_F( aLo.x );
	F2D_ISSUE_F( aHi.x );
	F2D_ISSUE_F( bLo.x );
	F2D_ISSUE_F( bHi.x );
	F2D_ISSUE_F( anchorA.x );
	F2D_ISSUE_F( anchorB.x );
	F2D_ISSUE_F( deltaCenter.x );
	F2D_ISSUE_F( linearImpulse0.x );
	F2D_ISSUE_F( softness.biasRate );
	F2D_ISSUE_I( (int)enableSpring | (int)enableMotor << 1 | (int)enableLimit << 2 );
	V2 vA = { aLo.x, aLo.y };
	float wA = aLo.z;
	V2 vB = { bLo.x, bLo.y };
	float wB = bLo.z;
	const Rot dqA = { aHi.z, aHi.w };
	const Rot dqB = { bHi.z, bHi.w };
	bool fixedRotation = ( iA + iB == 0.0f );

	if ( enableSpring && fixedRotation == false )
	{
		float jointAngle = relativeAngle( dqB, dqA ) + j.deltaAngle;
		float jointAngleDelta = unwindAngle( jointAngle - j.targetAngle );
		float C = jointAngleDelta;
		float bias = j.springSoftness.biasRate * C;
		float massScale = j.springSoftness.massScale;
		float impulseScale = j.springSoftness.impulseScale;
		float Cdot = wB - wA;
		float impulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * j.springImpulse;
		j.springImpulse += impulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}
	if ( enableMotor && fixedRotation == false )
	{
		float Cdot = wB - wA - j.motorSpeed;
		float impulse = -j.axialMass * Cdot;
		float oldImpulse = j.motorImpulse;
		float maxImpulse = h * j.maxMotorTorque;
		j.motorImpulse = clampf( j.motorImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j.motorImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}
	if ( enableLimit && fixedRotation == false )
	{
		float jointAngle = relativeAngle( dqB, dqA ) + j.deltaAngle - j.referenceAngle;
		jointAngle = unwindAngle( jointAngle );
		{
			float C = jointAngle - j.lowerAngle;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if ( C > 0.0f )
				bias = C * inv_h;
			else if ( useBias )
			{
				bias = base.constraintSoftness.biasRate * C;
				massScale = base.constraintSoftness.massScale;
				impulseScale = base.constraintSoftness.impulseScale;
			}
			float Cdot = wB - wA;
			float oldImpulse = j.lowerImpulse;
			float impulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * oldImpulse;
			j.lowerImpulse = maxf( oldImpulse + impulse, 0.0f );
			impulse = j.lowerImpulse - oldImpulse;
			wA -= iA * impulse;
			wB += iB * impulse;
		}
		{
			float C = j.upperAngle - jointAngle;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if ( C > 0.0f )
				bias = C * inv_h;
			else if ( useBias )
			{
				bias = base.constraintSoftness.biasRate * C;
				massScale = base.constraintSoftness.massScale;
				impulseScale = base.constraintSoftness.impulseScale;
			}
			float Cdot = wA - wB;
			float oldImpulse = j.upperImpulse;
			float impulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * oldImpulse;
			j.upperImpulse = maxf( oldImpulse + impulse, 0.0f );
			impulse = j.upperImpulse - oldImpulse;
			wA += iA * impulse;
			wB -= iB * impulse;
		}
	}
	{
		V2 rA = rotate( dqA, anchorA );
		V2 rB = rotate( dqB, anchorB );
		V2 Cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );
		V2 bias = { 0.0f, 0.0f };
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			V2 dcA = { aHi.x, aHi.y };
			V2 dcB = { bHi.x, bHi.y };
			V2 separation = add( add( sub( dcB, dcA ), sub( rB, rA ) ), deltaCenter );
			bias = mulSV( softness.biasRate, separation );
			massScale = softness.massScale;
			impulseScale = softness.impulseScale;
		}
		M22 K;
		K.cx.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		K.cy.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		K.cx.y = K.cy.x;
		K.cy.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		V2 b = solve22( K, add( Cdot, bias ) );
		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * linearImpulse0.x;
		impulse.y = -massScale * b.y - impulseScale * linearImpulse0.y;
		j.linearImpulse.x = linearImpulse0.x + impulse.x;
		j.linearImpulse.y = linearImpulse0.y + impulse.y;
		vA = mulSub( vA, mA, impulse );
		wA -= iA * cross( rA, impulse );
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}
	store16( &sA->v, Q4{ vA.x, vA.y, wA, aLo.w } );
	store16( &sB->v, Q4{ vB.x, vB.y, wB, bLo.w } );
}

// ------------------------------------------------------------------------------------------------ shared pieces
// The two solver bodies of a joint while one joint routine runs: velocities in registers, written back by finish().
// A body that is not awake (static, or asleep on the far side of a disabled link) is the identity state.
struct JointPair
{
	float mA, mB, iA, iB;
	BodyState dummyA, dummyB;
	BodyState *sA, *sB;
	V2 vA, vB;
	float wA, wB;
	F2D_HD void begin( const JointSim& base, int indexA, int indexB, BodyState* states )
	{
		mA = base.invMassA;
		mB = base.invMassB;
		iA = base.invIA;
		iB = base.invIB;
		dummyA = identityState();
		dummyB = identityState();
		sA = indexA == kNull ? &dummyA : states + indexA;
		sB = indexB == kNull ? &dummyB : states + indexB;
		vA = sA->v;
		wA = sA->w;
		vB = sB->v;
		wB = sB->w;
	}
	// equal and opposite impulse P at the anchors rA / rB
	F2D_HD void pushAt( V2 rA, V2 rB, V2 P )
	{
		vA = mulSub( vA, mA, P );
		wA -= iA * cross( rA, P );
		vB = mulAdd( vB, mB, P );
		wB += iB * cross( rB, P );
	}
	// linear impulse P with precomputed angular arms LA / LB
	F2D_HD void pushArms( V2 P, float LA, float LB )
	{
		vA = mulSub( vA, mA, P );
		wA -= iA * LA;
		vB = mulAdd( vB, mB, P );
		wB += iB * LB;
	}
	// the same with the sign of an upper limit (body A pushed forward, body B back)
	F2D_HD void pullArms( V2 P, float LA, float LB )
	{
		vA = mulAdd( vA, mA, P );
		wA += iA * LA;
		vB = mulSub( vB, mB, P );
		wB -= iB * LB;
	}
	F2D_HD void finish()
	{
		sA->v = vA;
		sA->w = wA;
		sB->v = vB;
		sB->w = wB;
	}
};

// what every prepare routine starts with: inverse masses into the joint, awake indices, anchors about the centres of mass
struct JointFrame
{
	int indexA, indexB;
	V2 anchorA, anchorB, deltaCenter;
	Rot qA, qB;
};
F2D_HDF inline JointFrame prepareFrame( World* w, JointSim& base )
{
	const Body* bodies = ptr( w, w->bodies );
	const BodySim* sims = ptr( w, w->sims );
	const Body& bodyA = bodies[base.bodyIdA];
	const Body& bodyB = bodies[base.bodyIdB];
	const BodySim& simA = sims[base.bodyIdA];
	const BodySim& simB = sims[base.bodyIdB];
	base.invMassA = simA.invMass;
	base.invMassB = simB.invMass;
	base.invIA = simA.invInertia;
	base.invIB = simB.invInertia;
	JointFrame f;
	f.indexA = bodyA.setIndex == kAwakeSet ? bodyA.localIndex : kNull;
	f.indexB = bodyB.setIndex == kAwakeSet ? bodyB.localIndex : kNull;
	f.qA = simA.transform.q;
	f.qB = simB.transform.q;
	f.anchorA = rotate( f.qA, sub( base.localOriginAnchorA, simA.localCenter ) );
	f.anchorB = rotate( f.qB, sub( base.localOriginAnchorB, simB.localCenter ) );
	f.deltaCenter = sub( simB.center, simA.center );
	return f;
}

// limit rows share one soft-constraint choice: speculative when open, soft when violated and biased, rigid otherwise
struct RowSoftness
{
	float bias, massScale, impulseScale;
};
F2D_HD RowSoftness limitRow( float C, bool useBias, const Soft& soft, float inv_h )
{
	RowSoftness r = { 0.0f, 1.0f, 0.0f };
	if ( C > 0.0f )
		r.bias = C * inv_h;
	else if ( useBias )
	{
		r.bias = soft.biasRate * C;
		r.massScale = soft.massScale;
		r.impulseScale = soft.impulseScale;
	}
	return r;
}

// ------------------------------------------------------------------------------------------------ distance joint
// distance_joint.c:245-293
F2D_HDF inline void prepareDistance( World* w, JointSim& base )
{
	JointFrame f = prepareFrame( w, base );
	DistanceJointData& j = base.distance;
	j.indexA = f.indexA;
	j.indexB = f.indexB;
	j.anchorA = f.anchorA;
	j.anchorB = f.anchorB;
	j.deltaCenter = f.deltaCenter;
	V2 rA = j.anchorA, rB = j.anchorB;
	V2 separation = add( sub( rB, rA ), j.deltaCenter );
	V2 axis = normalize( separation );
	float crA = cross( rA, axis );
	float crB = cross( rB, axis );
	float k = base.invMassA + base.invMassB + base.invIA * crA * crA + base.invIB * crB * crB;
	j.axialMass = k > 0.0f ? 1.0f / k : 0.0f;
	j.distanceSoftness = makeSoft( j.hertz, j.dampingRatio, w->step.h );
	if ( w->step.enableWarmStarting == 0 )
	{
		j.impulse = 0.0f;
		j.lowerImpulse = 0.0f;
		j.upperImpulse = 0.0f;
		j.motorImpulse = 0.0f;
	}
}
// distance_joint.c:295-325
F2D_HDF inline void warmStartDistance( JointSim& base, BodyState* states )
{
	DistanceJointData& j = base.distance;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	V2 rA = rotate( p.sA->dq, j.anchorA );
	V2 rB = rotate( p.sB->dq, j.anchorB );
	V2 ds = add( sub( p.sB->dp, p.sA->dp ), sub( rB, rA ) );
	V2 separation = add( j.deltaCenter, ds );
	V2 axis = normalize( separation );
	float axialImpulse = j.impulse + j.lowerImpulse - j.upperImpulse + j.motorImpulse;
	p.pushAt( rA, rB, mulSV( axialImpulse, axis ) );
	p.finish();
}
// distance_joint.c:327-500
F2D_HDF inline void solveDistance( JointSim& base, BodyState* states, bool useBias, float h, float inv_h )
{
	DistanceJointData& j = base.distance;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	V2 rA = rotate( p.sA->dq, j.anchorA );
	V2 rB = rotate( p.sB->dq, j.anchorB );
	V2 ds = add( sub( p.sB->dp, p.sA->dp ), sub( rB, rA ) );
	V2 separation = add( j.deltaCenter, ds );
	float len = length( separation );
	V2 axis = normalize( separation );
	// relative velocity of the anchors along the axis, B relative to A
	auto axialSpeed = [&]() {
		V2 vr = add( sub( p.vB, p.vA ), sub( crossSV( p.wB, rB ), crossSV( p.wA, rA ) ) );
		return dot( axis, vr );
	};
	if ( j.enableSpring && ( j.minLength < j.maxLength || j.enableLimit == false ) )
	{
		if ( j.hertz > 0.0f )
		{
			float Cdot = axialSpeed();
			float C = len - j.length;
			float bias = j.distanceSoftness.biasRate * C;
			float m = j.distanceSoftness.massScale * j.axialMass;
			float impulse = -m * ( Cdot + bias ) - j.distanceSoftness.impulseScale * j.impulse;
			j.impulse += impulse;
			p.pushAt( rA, rB, mulSV( impulse, axis ) );
		}
		if ( j.enableLimit )
		{
			{
				float Cdot = axialSpeed();
				RowSoftness r = limitRow( len - j.minLength, useBias, base.constraintSoftness, inv_h );
				float impulse = -r.massScale * j.axialMass * ( Cdot + r.bias ) - r.impulseScale * j.lowerImpulse;
				float newImpulse = maxf( 0.0f, j.lowerImpulse + impulse );
				impulse = newImpulse - j.lowerImpulse;
				j.lowerImpulse = newImpulse;
				p.pushAt( rA, rB, mulSV( impulse, axis ) );
			}
			{
				V2 vr = add( sub( p.vA, p.vB ), sub( crossSV( p.wA, rA ), crossSV( p.wB, rB ) ) );
				float Cdot = dot( axis, vr );
				RowSoftness r = limitRow( j.maxLength - len, useBias, base.constraintSoftness, inv_h );
				float impulse = -r.massScale * j.axialMass * ( Cdot + r.bias ) - r.impulseScale * j.upperImpulse;
				float newImpulse = maxf( 0.0f, j.upperImpulse + impulse );
				impulse = newImpulse - j.upperImpulse;
				j.upperImpulse = newImpulse;
				p.pushAt( rA, rB, mulSV( -impulse, axis ) );
			}
		}
		if ( j.enableMotor )
		{
			float Cdot = axialSpeed();
			float impulse = j.axialMass * ( j.motorSpeed - Cdot );
			float oldImpulse = j.motorImpulse;
			float maxImpulse = h * j.maxMotorForce;
			j.motorImpulse = clampf( j.motorImpulse + impulse, -maxImpulse, maxImpulse );
			impulse = j.motorImpulse - oldImpulse;
			p.pushAt( rA, rB, mulSV( impulse, axis ) );
		}
	}
	else
	{
		// rigid rod
		float Cdot = axialSpeed();
		float C = len - j.length;
		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			bias = base.constraintSoftness.biasRate * C;
			massScale = base.constraintSoftness.massScale;
			impulseScale = base.constraintSoftness.impulseScale;
		}
		float impulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * j.impulse;
		j.impulse += impulse;
		p.pushAt( rA, rB, mulSV( impulse, axis ) );
	}
	p.finish();
}

// ------------------------------------------------------------------------------------------------ prismatic joint
// prismatic_joint.c:253-316
F2D_HDF inline void preparePrismatic( World* w, JointSim& base )
{
	JointFrame f = prepareFrame( w, base );
	PrismaticJointData& j = base.prismatic;
	j.indexA = f.indexA;
	j.indexB = f.indexB;
	j.anchorA = f.anchorA;
	j.anchorB = f.anchorB;
	j.axisA = rotate( f.qA, j.localAxisA );
	j.deltaCenter = f.deltaCenter;
	j.deltaAngle = relativeAngle( f.qB, f.qA ) - j.referenceAngle;
	j.deltaAngle = unwindAngle( j.deltaAngle );
	V2 rA = j.anchorA, rB = j.anchorB;
	V2 d = add( j.deltaCenter, sub( rB, rA ) );
	float a1 = cross( add( d, rA ), j.axisA );
	float a2 = cross( rB, j.axisA );
	float k = base.invMassA + base.invMassB + base.invIA * a1 * a1 + base.invIB * a2 * a2;
	j.axialMass = k > 0.0f ? 1.0f / k : 0.0f;
	j.springSoftness = makeSoft( j.hertz, j.dampingRatio, w->step.h );
	if ( w->step.enableWarmStarting == 0 )
	{
		j.impulse = V2{ 0.0f, 0.0f };
		j.springImpulse = 0.0f;
		j.motorImpulse = 0.0f;
		j.lowerImpulse = 0.0f;
		j.upperImpulse = 0.0f;
	}
}
// prismatic_joint.c:318-361
F2D_HDF inline void warmStartPrismatic( JointSim& base, BodyState* states )
{
	PrismaticJointData& j = base.prismatic;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	V2 rA = rotate( p.sA->dq, j.anchorA );
	V2 rB = rotate( p.sB->dq, j.anchorB );
	V2 d = add( add( sub( p.sB->dp, p.sA->dp ), j.deltaCenter ), sub( rB, rA ) );
	V2 axisA = rotate( p.sA->dq, j.axisA );
	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );
	float axialImpulse = j.springImpulse + j.motorImpulse + j.lowerImpulse - j.upperImpulse;
	V2 perpA = leftPerp( axisA );
	float s1 = cross( add( d, rA ), perpA );
	float s2 = cross( rB, perpA );
	float perpImpulse = j.impulse.x;
	float angleImpulse = j.impulse.y;
	V2 P = add( mulSV( axialImpulse, axisA ), mulSV( perpImpulse, perpA ) );
	float LA = axialImpulse * a1 + perpImpulse * s1 + angleImpulse;
	float LB = axialImpulse * a2 + perpImpulse * s2 + angleImpulse;
	p.pushArms( P, LA, LB );
	p.finish();
}
// prismatic_joint.c:363-575
F2D_HDF inline void solvePrismatic( JointSim& base, BodyState* states, bool useBias, float h, float inv_h )
{
	PrismaticJointData& j = base.prismatic;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	V2 rA = rotate( p.sA->dq, j.anchorA );
	V2 rB = rotate( p.sB->dq, j.anchorB );
	V2 d = add( add( sub( p.sB->dp, p.sA->dp ), j.deltaCenter ), sub( rB, rA ) );
	V2 axisA = rotate( p.sA->dq, j.axisA );
	float translation = dot( axisA, d );
	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );
	auto axialSpeed = [&]() { return dot( axisA, sub( p.vB, p.vA ) ) + a2 * p.wB - a1 * p.wA; };
	if ( j.enableSpring )
	{
		float C = translation - j.targetTranslation;
		float bias = j.springSoftness.biasRate * C;
		float massScale = j.springSoftness.massScale;
		float impulseScale = j.springSoftness.impulseScale;
		float Cdot = axialSpeed();
		float deltaImpulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * j.springImpulse;
		j.springImpulse += deltaImpulse;
		p.pushArms( mulSV( deltaImpulse, axisA ), deltaImpulse * a1, deltaImpulse * a2 );
	}
	if ( j.enableMotor )
	{
		float Cdot = axialSpeed();
		float impulse = j.axialMass * ( j.motorSpeed - Cdot );
		float oldImpulse = j.motorImpulse;
		float maxImpulse = h * j.maxMotorForce;
		j.motorImpulse = clampf( j.motorImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j.motorImpulse - oldImpulse;
		p.pushArms( mulSV( impulse, axisA ), impulse * a1, impulse * a2 );
	}
	if ( j.enableLimit )
	{
		{
			RowSoftness r = limitRow( translation - j.lowerTranslation, useBias, base.constraintSoftness, inv_h );
			float oldImpulse = j.lowerImpulse;
			float Cdot = axialSpeed();
			float impulse = -j.axialMass * r.massScale * ( Cdot + r.bias ) - r.impulseScale * oldImpulse;
			j.lowerImpulse = maxf( oldImpulse + impulse, 0.0f );
			impulse = j.lowerImpulse - oldImpulse;
			p.pushArms( mulSV( impulse, axisA ), impulse * a1, impulse * a2 );
		}
		{
			RowSoftness r = limitRow( j.upperTranslation - translation, useBias, base.constraintSoftness, inv_h );
			float oldImpulse = j.upperImpulse;
			float Cdot = dot( axisA, sub( p.vA, p.vB ) ) + a1 * p.wA - a2 * p.wB;
			float impulse = -j.axialMass * r.massScale * ( Cdot + r.bias ) - r.impulseScale * oldImpulse;
			j.upperImpulse = maxf( oldImpulse + impulse, 0.0f );
			impulse = j.upperImpulse - oldImpulse;
			p.pullArms( mulSV( impulse, axisA ), impulse * a1, impulse * a2 );
		}
	}
	{
		// the 2x2 block: no motion across the axis, no relative rotation
		V2 perpA = leftPerp( axisA );
		float s1 = cross( add( d, rA ), perpA );
		float s2 = cross( rB, perpA );
		V2 Cdot;
		Cdot.x = dot( perpA, sub( p.vB, p.vA ) ) + s2 * p.wB - s1 * p.wA;
		Cdot.y = p.wB - p.wA;
		V2 bias = { 0.0f, 0.0f };
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			V2 C;
			C.x = dot( perpA, d );
			C.y = relativeAngle( p.sB->dq, p.sA->dq ) + j.deltaAngle;
			bias = mulSV( base.constraintSoftness.biasRate, C );
			massScale = base.constraintSoftness.massScale;
			impulseScale = base.constraintSoftness.impulseScale;
		}
		float k11 = p.mA + p.mB + p.iA * s1 * s1 + p.iB * s2 * s2;
		float k12 = p.iA * s1 + p.iB * s2;
		float k22 = p.iA + p.iB;
		if ( k22 == 0.0f )
			k22 = 1.0f;
		M22 K;
		K.cx = V2{ k11, k12 };
		K.cy = V2{ k12, k22 };
		V2 b = solve22( K, add( Cdot, bias ) );
		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j.impulse.x;
		impulse.y = -massScale * b.y - impulseScale * j.impulse.y;
		j.impulse.x += impulse.x;
		j.impulse.y += impulse.y;
		p.pushArms( mulSV( impulse.x, perpA ), impulse.x * s1 + impulse.y, impulse.x * s2 + impulse.y );
	}
	p.finish();
}

// ------------------------------------------------------------------------------------------------ weld joint
// weld_joint.c:92-154
F2D_HDF inline void prepareWeld( World* w, JointSim& base )
{
	JointFrame f = prepareFrame( w, base );
	WeldJointData& j = base.weld;
	j.indexA = f.indexA;
	j.indexB = f.indexB;
	j.anchorA = f.anchorA;
	j.anchorB = f.anchorB;
	j.deltaCenter = f.deltaCenter;
	j.deltaAngle = relativeAngle( f.qB, f.qA ) - j.referenceAngle;
	j.deltaAngle = unwindAngle( j.deltaAngle );
	float ka = base.invIA + base.invIB;
	j.axialMass = ka > 0.0f ? 1.0f / ka : 0.0f;
	j.linearSoftness = j.linearHertz == 0.0f ? base.constraintSoftness : makeSoft( j.linearHertz, j.linearDampingRatio, w->step.h );
	j.angularSoftness = j.angularHertz == 0.0f ? base.constraintSoftness : makeSoft( j.angularHertz, j.angularDampingRatio, w->step.h );
	if ( w->step.enableWarmStarting == 0 )
	{
		j.linearImpulse = V2{ 0.0f, 0.0f };
		j.angularImpulse = 0.0f;
	}
}
// shared by weld and motor joints: weld_joint.c:156-180, motor_joint.c:130-152
F2D_HD void warmStartLinearAngular( JointPair& p, V2 anchorA, V2 anchorB, V2 linearImpulse, float angularImpulse )
{
	V2 rA = rotate( p.sA->dq, anchorA );
	V2 rB = rotate( p.sB->dq, anchorB );
	p.sA->v = mulSub( p.sA->v, p.mA, linearImpulse );
	p.sA->w -= p.iA * ( cross( rA, linearImpulse ) + angularImpulse );
	p.sB->v = mulAdd( p.sB->v, p.mB, linearImpulse );
	p.sB->w += p.iB * ( cross( rB, linearImpulse ) + angularImpulse );
}
F2D_HDF inline void warmStartWeld( JointSim& base, BodyState* states )
{
	WeldJointData& j = base.weld;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	warmStartLinearAngular( p, j.anchorA, j.anchorB, j.linearImpulse, j.angularImpulse );
}
// the point-to-point block of revolute and weld joints: K = sum of the two bodies' anchor mobilities
F2D_HD M22 anchorMobility( const JointPair& p, V2 rA, V2 rB )
{
	M22 K;
	K.cx.x = p.mA + p.mB + rA.y * rA.y * p.iA + rB.y * rB.y * p.iB;
	K.cy.x = -rA.y * rA.x * p.iA - rB.y * rB.x * p.iB;
	K.cx.y = K.cy.x;
	K.cy.y = p.mA + p.mB + rA.x * rA.x * p.iA + rB.x * rB.x * p.iB;
	return K;
}
// weld_joint.c:182-262
F2D_HDF inline void solveWeld( JointSim& base, BodyState* states, bool useBias )
{
	WeldJointData& j = base.weld;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	{
		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias || j.angularHertz > 0.0f )
		{
			float C = relativeAngle( p.sB->dq, p.sA->dq ) + j.deltaAngle;
			bias = j.angularSoftness.biasRate * C;
			massScale = j.angularSoftness.massScale;
			impulseScale = j.angularSoftness.impulseScale;
		}
		float Cdot = p.wB - p.wA;
		float impulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * j.angularImpulse;
		j.angularImpulse += impulse;
		p.wA -= p.iA * impulse;
		p.wB += p.iB * impulse;
	}
	{
		V2 rA = rotate( p.sA->dq, j.anchorA );
		V2 rB = rotate( p.sB->dq, j.anchorB );
		V2 bias = { 0.0f, 0.0f };
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias || j.linearHertz > 0.0f )
		{
			V2 C = add( add( sub( p.sB->dp, p.sA->dp ), sub( rB, rA ) ), j.deltaCenter );
			bias = mulSV( j.linearSoftness.biasRate, C );
			massScale = j.linearSoftness.massScale;
			impulseScale = j.linearSoftness.impulseScale;
		}
		V2 Cdot = sub( add( p.vB, crossSV( p.wB, rB ) ), add( p.vA, crossSV( p.wA, rA ) ) );
		V2 b = solve22( anchorMobility( p, rA, rB ), add( Cdot, bias ) );
		V2 impulse = { -massScale * b.x - impulseScale * j.linearImpulse.x, -massScale * b.y - impulseScale * j.linearImpulse.y };
		j.linearImpulse = add( j.linearImpulse, impulse );
		p.pushAt( rA, rB, impulse );
	}
	p.finish();
}

// ------------------------------------------------------------------------------------------------ wheel joint
// wheel_joint.c:239-312
F2D_HDF inline void prepareWheel( World* w, JointSim& base )
{
	JointFrame f = prepareFrame( w, base );
	WheelJointData& j = base.wheel;
	j.indexA = f.indexA;
	j.indexB = f.indexB;
	j.anchorA = f.anchorA;
	j.anchorB = f.anchorB;
	j.axisA = rotate( f.qA, j.localAxisA );
	j.deltaCenter = f.deltaCenter;
	const float mA = base.invMassA, mB = base.invMassB, iA = base.invIA, iB = base.invIB;
	V2 rA = j.anchorA, rB = j.anchorB;
	V2 d = add( j.deltaCenter, sub( rB, rA ) );
	V2 axisA = j.axisA;
	V2 perpA = leftPerp( axisA );
	float s1 = cross( add( d, rA ), perpA );
	float s2 = cross( rB, perpA );
	float kp = mA + mB + iA * s1 * s1 + iB * s2 * s2;
	j.perpMass = kp > 0.0f ? 1.0f / kp : 0.0f;
	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );
	float ka = mA + mB + iA * a1 * a1 + iB * a2 * a2;
	j.axialMass = ka > 0.0f ? 1.0f / ka : 0.0f;
	j.springSoftness = makeSoft( j.hertz, j.dampingRatio, w->step.h );
	float km = iA + iB;
	j.motorMass = km > 0.0f ? 1.0f / km : 0.0f;
	if ( w->step.enableWarmStarting == 0 )
	{
		j.perpImpulse = 0.0f;
		j.springImpulse = 0.0f;
		j.motorImpulse = 0.0f;
		j.lowerImpulse = 0.0f;
		j.upperImpulse = 0.0f;
	}
}
// wheel_joint.c:314-352
F2D_HDF inline void warmStartWheel( JointSim& base, BodyState* states )
{
	WheelJointData& j = base.wheel;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	V2 rA = rotate( p.sA->dq, j.anchorA );
	V2 rB = rotate( p.sB->dq, j.anchorB );
	V2 d = add( add( sub( p.sB->dp, p.sA->dp ), j.deltaCenter ), sub( rB, rA ) );
	V2 axisA = rotate( p.sA->dq, j.axisA );
	V2 perpA = leftPerp( axisA );
	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );
	float s1 = cross( add( d, rA ), perpA );
	float s2 = cross( rB, perpA );
	float axialImpulse = j.springImpulse + j.lowerImpulse - j.upperImpulse;
	V2 P = add( mulSV( axialImpulse, axisA ), mulSV( j.perpImpulse, perpA ) );
	float LA = axialImpulse * a1 + j.perpImpulse * s1 + j.motorImpulse;
	float LB = axialImpulse * a2 + j.perpImpulse * s2 + j.motorImpulse;
	p.pushArms( P, LA, LB );
	p.finish();
}
// wheel_joint.c:354-520
F2D_HDF inline void solveWheel( JointSim& base, BodyState* states, bool useBias, float h, float inv_h )
{
	WheelJointData& j = base.wheel;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	bool fixedRotation = ( p.iA + p.iB == 0.0f );
	V2 rA = rotate( p.sA->dq, j.anchorA );
	V2 rB = rotate( p.sB->dq, j.anchorB );
	V2 d = add( add( sub( p.sB->dp, p.sA->dp ), j.deltaCenter ), sub( rB, rA ) );
	V2 axisA = rotate( p.sA->dq, j.axisA );
	float translation = dot( axisA, d );
	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );
	auto axialSpeed = [&]() { return dot( axisA, sub( p.vB, p.vA ) ) + a2 * p.wB - a1 * p.wA; };
	if ( j.enableMotor && fixedRotation == false )
	{
		float Cdot = p.wB - p.wA - j.motorSpeed;
		float impulse = -j.motorMass * Cdot;
		float oldImpulse = j.motorImpulse;
		float maxImpulse = h * j.maxMotorTorque;
		j.motorImpulse = clampf( j.motorImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j.motorImpulse - oldImpulse;
		p.wA -= p.iA * impulse;
		p.wB += p.iB * impulse;
	}
	if ( j.enableSpring )
	{
		float C = translation;
		float bias = j.springSoftness.biasRate * C;
		float massScale = j.springSoftness.massScale;
		float impulseScale = j.springSoftness.impulseScale;
		float Cdot = axialSpeed();
		float impulse = -massScale * j.axialMass * ( Cdot + bias ) - impulseScale * j.springImpulse;
		j.springImpulse += impulse;
		p.pushArms( mulSV( impulse, axisA ), impulse * a1, impulse * a2 );
	}
	if ( j.enableLimit )
	{
		{
			RowSoftness r = limitRow( translation - j.lowerTranslation, useBias, base.constraintSoftness, inv_h );
			float Cdot = axialSpeed();
			float impulse = -r.massScale * j.axialMass * ( Cdot + r.bias ) - r.impulseScale * j.lowerImpulse;
			float oldImpulse = j.lowerImpulse;
			j.lowerImpulse = maxf( oldImpulse + impulse, 0.0f );
			impulse = j.lowerImpulse - oldImpulse;
			p.pushArms( mulSV( impulse, axisA ), impulse * a1, impulse * a2 );
		}
		{
			RowSoftness r = limitRow( j.upperTranslation - translation, useBias, base.constraintSoftness, inv_h );
			float Cdot = dot( axisA, sub( p.vA, p.vB ) ) + a1 * p.wA - a2 * p.wB;
			float impulse = -r.massScale * j.axialMass * ( Cdot + r.bias ) - r.impulseScale * j.upperImpulse;
			float oldImpulse = j.upperImpulse;
			j.upperImpulse = maxf( oldImpulse + impulse, 0.0f );
			impulse = j.upperImpulse - oldImpulse;
			p.pullArms( mulSV( impulse, axisA ), impulse * a1, impulse * a2 );
		}
	}
	{
		// no motion across the axis
		V2 perpA = leftPerp( axisA );
		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			float C = dot( perpA, d );
			bias = base.constraintSoftness.biasRate * C;
			massScale = base.constraintSoftness.massScale;
			impulseScale = base.constraintSoftness.impulseScale;
		}
		float s1 = cross( add( d, rA ), perpA );
		float s2 = cross( rB, perpA );
		float Cdot = dot( perpA, sub( p.vB, p.vA ) ) + s2 * p.wB - s1 * p.wA;
		float impulse = -massScale * j.perpMass * ( Cdot + bias ) - impulseScale * j.perpImpulse;
		j.perpImpulse += impulse;
		p.pushArms( mulSV( impulse, perpA ), impulse * s1, impulse * s2 );
	}
	p.finish();
}

// ------------------------------------------------------------------------------------------------ motor joint
// motor_joint.c:80-128
F2D_HDF inline void prepareMotor( World* w, JointSim& base )
{
	JointFrame f = prepareFrame( w, base );
	MotorJointData& j = base.motor;
	j.indexA = f.indexA;
	j.indexB = f.indexB;
	j.anchorA = f.anchorA;
	j.anchorB = f.anchorB;
	j.deltaCenter = sub( f.deltaCenter, j.linearOffset );
	j.deltaAngle = relativeAngle( f.qB, f.qA ) - j.angularOffset;
	const float mA = base.invMassA, mB = base.invMassB, iA = base.invIA, iB = base.invIB;
	V2 rA = j.anchorA, rB = j.anchorB;
	M22 K;
	K.cx.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
	K.cx.y = -rA.y * rA.x * iA - rB.y * rB.x * iB;
	K.cy.x = K.cx.y;
	K.cy.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
	j.linearMass = inverse22( K );
	float ka = iA + iB;
	j.angularMass = ka > 0.0f ? 1.0f / ka : 0.0f;
	if ( w->step.enableWarmStarting == 0 )
	{
		j.linearImpulse = V2{ 0.0f, 0.0f };
		j.angularImpulse = 0.0f;
	}
}
F2D_HDF inline void warmStartMotor( JointSim& base, BodyState* states )
{
	MotorJointData& j = base.motor;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	warmStartLinearAngular( p, j.anchorA, j.anchorB, j.linearImpulse, j.angularImpulse );
}
// motor_joint.c:154-238
F2D_HDF inline void solveMotor( JointSim& base, BodyState* states, float h, float inv_h )
{
	MotorJointData& j = base.motor;
	JointPair p;
	p.begin( base, j.indexA, j.indexB, states );
	{
		float angularSeparation = relativeAngle( p.sB->dq, p.sA->dq ) + j.deltaAngle;
		angularSeparation = unwindAngle( angularSeparation );
		float angularBias = inv_h * j.correctionFactor * angularSeparation;
		float Cdot = p.wB - p.wA;
		float impulse = -j.angularMass * ( Cdot + angularBias );
		float oldImpulse = j.angularImpulse;
		float maxImpulse = h * j.maxTorque;
		j.angularImpulse = clampf( j.angularImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j.angularImpulse - oldImpulse;
		p.wA -= p.iA * impulse;
		p.wB += p.iB * impulse;
	}
	{
		V2 rA = rotate( p.sA->dq, j.anchorA );
		V2 rB = rotate( p.sB->dq, j.anchorB );
		V2 ds = add( sub( p.sB->dp, p.sA->dp ), sub( rB, rA ) );
		V2 linearSeparation = add( j.deltaCenter, ds );
		V2 linearBias = mulSV( inv_h * j.correctionFactor, linearSeparation );
		V2 Cdot = sub( add( p.vB, crossSV( p.wB, rB ) ), add( p.vA, crossSV( p.wA, rA ) ) );
		V2 b = mulMV( j.linearMass, add( Cdot, linearBias ) );
		V2 impulse = { -b.x, -b.y };
		V2 oldImpulse = j.linearImpulse;
		float maxImpulse = h * j.maxForce;
		j.linearImpulse = add( j.linearImpulse, impulse );
		if ( lengthSq( j.linearImpulse ) > maxImpulse * maxImpulse )
		{
			j.linearImpulse = normalize( j.linearImpulse );
			j.linearImpulse.x *= maxImpulse;
			j.linearImpulse.y *= maxImpulse;
		}
		impulse = sub( j.linearImpulse, oldImpulse );
		p.pushAt( rA, rB, impulse );
	}
	p.finish();
}

// ------------------------------------------------------------------------------------------------ mouse joint
// mouse_joint.c:80-126: only body B takes part (body A is the anchor the target is expressed in)
F2D_HDF inline void prepareMouse( World* w, JointSim& base )
{
	const Body& bodyB = ptr( w, w->bodies )[base.bodyIdB];
	const BodySim& simB = ptr( w, w->sims )[base.bodyIdB];
	base.invMassB = simB.invMass;
	base.invIB = simB.invInertia;
	MouseJointData& j = base.mouse;
	j.indexB = bodyB.setIndex == kAwakeSet ? bodyB.localIndex : kNull;
	j.anchorB = rotate( simB.transform.q, sub( base.localOriginAnchorB, simB.localCenter ) );
	j.linearSoftness = makeSoft( j.hertz, j.dampingRatio, w->step.h );
	float angularHertz = 0.5f;
	float angularDampingRatio = 0.1f;
	j.angularSoftness = makeSoft( angularHertz, angularDampingRatio, w->step.h );
	V2 rB = j.anchorB;
	float mB = simB.invMass, iB = simB.invInertia;
	M22 K;
	K.cx.x = mB + iB * rB.y * rB.y;
	K.cx.y = -iB * rB.x * rB.y;
	K.cy.x = K.cx.y;
	K.cy.y = mB + iB * rB.x * rB.x;
	j.linearMass = inverse22( K );
	j.deltaCenter = sub( simB.center, j.targetA );
	if ( w->step.enableWarmStarting == 0 )
	{
		j.linearImpulse = V2{ 0.0f, 0.0f };
		j.angularImpulse = 0.0f;
	}
}
// mouse_joint.c:128-148
F2D_HDF inline void warmStartMouse( JointSim& base, BodyState* states )
{
	MouseJointData& j = base.mouse;
	if ( j.indexB == kNull )
		return; // the reference indexes the awake states unconditionally; a mouse joint on a non-awake body is inert here
	BodyState& sB = states[j.indexB];
	V2 rB = rotate( sB.dq, j.anchorB );
	V2 vB = mulAdd( sB.v, base.invMassB, j.linearImpulse );
	float wB = sB.w + base.invIB * ( cross( rB, j.linearImpulse ) + j.angularImpulse );
	sB.v = vB;
	sB.w = wB;
}
// mouse_joint.c:150-212
F2D_HDF inline void solveMouse( JointSim& base, BodyState* states, float h )
{
	MouseJointData& j = base.mouse;
	if ( j.indexB == kNull )
		return;
	float mB = base.invMassB, iB = base.invIB;
	BodyState& sB = states[j.indexB];
	V2 vB = sB.v;
	float wB = sB.w;
	{
		float massScale = j.angularSoftness.massScale;
		float impulseScale = j.angularSoftness.impulseScale;
		float impulse = iB > 0.0f ? -wB / iB : 0.0f;
		impulse = massScale * impulse - impulseScale * j.angularImpulse;
		j.angularImpulse += impulse;
		wB += iB * impulse;
	}
	float maxImpulse = j.maxForce * h;
	{
		V2 rB = rotate( sB.dq, j.anchorB );
		V2 Cdot = add( vB, crossSV( wB, rB ) );
		V2 separation = add( add( sB.dp, rB ), j.deltaCenter );
		V2 bias = mulSV( j.linearSoftness.biasRate, separation );
		float massScale = j.linearSoftness.massScale;
		float impulseScale = j.linearSoftness.impulseScale;
		V2 b = mulMV( j.linearMass, add( Cdot, bias ) );
		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j.linearImpulse.x;
		impulse.y = -massScale * b.y - impulseScale * j.linearImpulse.y;
		V2 oldImpulse = j.linearImpulse;
		j.linearImpulse.x += impulse.x;
		j.linearImpulse.y += impulse.y;
		float mag = length( j.linearImpulse );
		if ( mag > maxImpulse )
			j.linearImpulse = mulSV( maxImpulse, normalize( j.linearImpulse ) );
		impulse.x = j.linearImpulse.x - oldImpulse.x;
		impulse.y = j.linearImpulse.y - oldImpulse.y;
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}
	sB.v = vB;
	sB.w = wB;
}

// joint.c:1348-1390
F2D_HDF inline void prepareJoint( World* w, JointSim& joint )
{
	float hertz = minf( joint.constraintHertz, 0.25f * w->step.inv_h );
	joint.constraintSoftness = makeSoft( hertz, joint.constraintDampingRatio, w->step.h );
	switch ( joint.type )
	{
		case kDistanceJoint:
			prepareDistance( w, joint );
			break;
		case kMotorJoint:
			prepareMotor( w, joint );
			break;
		case kMouseJoint:
			prepareMouse( w, joint );
			break;
		case kFilterJoint:
			break;
		case kPrismaticJoint:
			preparePrismatic( w, joint );
			break;
		case kRevoluteJoint:
			prepareRevolute( w, joint );
			break;
		case kWeldJoint:
			prepareWeld( w, joint );
			break;
		case kWheelJoint:
			prepareWheel( w, joint );
			break;
		default:
			setError( w, kErrUnsupported, __LINE__ );
			break;
	}
}
// joint.c:1392-1430
F2D_HDF inline void warmStartJoint( World* w, JointSim& joint, BodyState* states )
{
	(void)w;
	switch ( joint.type )
	{
		case kDistanceJoint:
			warmStartDistance( joint, states );
			break;
		case kMotorJoint:
			warmStartMotor( joint, states );
			break;
		case kMouseJoint:
			warmStartMouse( joint, states );
			break;
		case kPrismaticJoint:
			warmStartPrismatic( joint, states );
			break;
		case kRevoluteJoint:
			warmStartRevolute( joint, states );
			break;
		case kWeldJoint:
			warmStartWeld( joint, states );
			break;
		case kWheelJoint:
			warmStartWheel( joint, states );
			break;
		default:
			break;
	}
}
// joint.c:1432-1470
F2D_HDF inline void solveJoint( World* w, JointSim& joint, BodyState* states, bool useBias )
{
	const float h = w->step.h, inv_h = w->step.inv_h;
	switch ( joint.type )
	{
		case kDistanceJoint:
			solveDistance( joint, states, useBias, h, inv_h );
			break;
		case kMotorJoint:
			solveMotor( joint, states, h, inv_h );
			break;
		case kMouseJoint:
			solveMouse( joint, states, h );
			break;
		case kPrismaticJoint:
			solvePrismatic( joint, states, useBias, h, inv_h );
			break;
		case kRevoluteJoint:
			solveRevolute( joint, states, useBias, h, inv_h );
			break;
		case kWeldJoint:
			solveWeld( joint, states, useBias );
			break;
		case kWheelJoint:
			solveWheel( joint, states, useBias, h, inv_h );
			break;
		default:
			break;
	}
}

} // namespace f2d
