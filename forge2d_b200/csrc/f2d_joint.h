// forge2d_b200 — joint constraints (prepare / warm start / solve), one thread per joint inside a colour.
// Scalar formulas of B2/src/joint.c:1348-1470 and B2/src/revolute_joint.c:202-465, expression for expression.
// Joint types other than revolute and filter raise kErrUnsupported on the device path (next scope row).
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
	V2 rA = rotate( sA->dq, j.anchorA );
	V2 rB = rotate( sB->dq, j.anchorB );
	float axialImpulse = j.springImpulse + j.motorImpulse + j.lowerImpulse - j.upperImpulse;
	sA->v = mulSub( sA->v, mA, j.linearImpulse );
	sA->w -= iA * ( cross( rA, j.linearImpulse ) + axialImpulse );
	sB->v = mulAdd( sB->v, mB, j.linearImpulse );
	sB->w += iB * ( cross( rB, j.linearImpulse ) + axialImpulse );
}

// revolute_joint.c:289-465
F2D_HDF inline void solveRevolute( JointSim& base, BodyState* states, bool useBias, float h, float inv_h )
{
	float mA = base.invMassA, mB = base.invMassB, iA = base.invIA, iB = base.invIB;
	BodyState dummy = identityState();
	RevoluteJointData& j = base.revolute;
	BodyState* sA = j.indexA == kNull ? &dummy : states + j.indexA;
	BodyState* sB = j.indexB == kNull ? &dummy : states + j.indexB;
	V2 vA = sA->v;
	float wA = sA->w;
	V2 vB = sB->v;
	float wB = sB->w;
	const Rot dqA = sA->dq;
	const Rot dqB = sB->dq;
	bool fixedRotation = ( iA + iB == 0.0f );

	if ( j.enableSpring && fixedRotation == false )
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
	if ( j.enableMotor && fixedRotation == false )
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
	if ( j.enableLimit && fixedRotation == false )
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
		V2 rA = rotate( dqA, j.anchorA );
		V2 rB = rotate( dqB, j.anchorB );
		V2 Cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );
		V2 bias = { 0.0f, 0.0f };
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			V2 dcA = sA->dp;
			V2 dcB = sB->dp;
			V2 separation = add( add( sub( dcB, dcA ), sub( rB, rA ) ), j.deltaCenter );
			bias = mulSV( base.constraintSoftness.biasRate, separation );
			massScale = base.constraintSoftness.massScale;
			impulseScale = base.constraintSoftness.impulseScale;
		}
		M22 K;
		K.cx.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		K.cy.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		K.cx.y = K.cy.x;
		K.cy.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		V2 b = solve22( K, add( Cdot, bias ) );
		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j.linearImpulse.x;
		impulse.y = -massScale * b.y - impulseScale * j.linearImpulse.y;
		j.linearImpulse.x += impulse.x;
		j.linearImpulse.y += impulse.y;
		vA = mulSub( vA, mA, impulse );
		wA -= iA * cross( rA, impulse );
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}
	sA->v = vA;
	sA->w = wA;
	sB->v = vB;
	sB->w = wB;
}

// joint.c:1348-1382
F2D_HDF inline void prepareJoint( World* w, JointSim& joint )
{
	float hertz = minf( joint.constraintHertz, 0.25f * w->step.inv_h );
	joint.constraintSoftness = makeSoft( hertz, joint.constraintDampingRatio, w->step.h );
	switch ( joint.type )
	{
		case kRevoluteJoint:
			prepareRevolute( w, joint );
			break;
		case kFilterJoint:
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
	if ( joint.type == kRevoluteJoint )
		warmStartRevolute( joint, states );
}
// joint.c:1432-1470
F2D_HDF inline void solveJoint( World* w, JointSim& joint, BodyState* states, bool useBias )
{
	if ( joint.type == kRevoluteJoint )
		solveRevolute( joint, states, useBias, w->step.h, w->step.inv_h );
}

} // namespace f2d
