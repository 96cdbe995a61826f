// forge2d_b200 — API-time world mutations on a world image: create / destroy bodies, shapes and joints, mass data.
// Executed on the HOST image (the C-ABI layer syncs host <-> device around them); identical ordering semantics to
// B2/src/body.c, shape.c, joint.c because the state they build (id pools, shape lists, proxy ids, move buffer,
// islands, colours) decides the order of everything the step does afterwards.
#pragma once
#include "f2d_world.h"
#include <vector>

namespace f2d
{

struct BodyParams // mirrors b2BodyDef (types.h) without pointers
{
	int type;
	V2 position;
	Rot rotation;
	V2 linearVelocity;
	float angularVelocity, linearDamping, angularDamping, gravityScale, sleepThreshold;
	const char* name;
	uint64_t userData;
	bool enableSleep, isAwake, fixedRotation, isBullet, isEnabled, allowFastRotation;
};

struct ShapeParams // mirrors b2ShapeDef (types.h)
{
	uint64_t userData;
	float friction, restitution, rollingResistance, tangentSpeed;
	int userMaterialId;
	uint32_t customColor;
	float density;
	Filter filter;
	bool isSensor, enableSensorEvents, enableContactEvents, enableHitEvents, enablePreSolveEvents, invokeContactCreation,
		updateBodyMass;
};

// body.c:84-97
inline void createIslandForBody( World* w, int setIndex, Body& body )
{
	int islandId = createIsland( w, setIndex );
	Island& is = ptr( w, w->islands )[islandId];
	body.islandId = islandId;
	is.headBody = body.id;
	is.tailBody = body.id;
	is.bodyCount = 1;
}

// body.c:177-330. Returns the body id (0-based).
inline int createBody( World* w, const BodyParams& def )
{
	bool isAwake = ( def.isAwake || def.enableSleep == false ) && def.isEnabled;
	int setId;
	if ( def.isEnabled == false )
		setId = kDisabledSet;
	else if ( def.type == kStaticBody )
		setId = kStaticSet;
	else if ( isAwake )
		setId = kAwakeSet;
	else
	{
		// new sleeping set holding just this body and its island
		setId = allocSolverSet( w );
		SolverSet& s = ptr( w, w->sets )[setId];
		memset( &s, 0, sizeof( s ) );
		s.setIndex = setId;
		sleepPoolAlloc( w, s, setId, 1, 0, 0, 1 );
	}

	int bodyId = allocId( w, w->bodyIds );
	if ( bodyId == w->bodies.count )
	{
		Body empty;
		memset( &empty, 0, sizeof( empty ) );
		F2D_PUSH( w, w->bodies, empty );
		BodySim emptySim;
		memset( &emptySim, 0, sizeof( emptySim ) );
		F2D_PUSH( w, w->sims, emptySim );
	}
	BodySim& sim = ptr( w, w->sims )[bodyId];
	memset( &sim, 0, sizeof( sim ) );
	sim.transform.p = def.position;
	sim.transform.q = def.rotation;
	sim.center = def.position;
	sim.rotation0 = sim.transform.q;
	sim.center0 = sim.center;
	sim.minExtent = kHuge;
	sim.maxExtent = 0.0f;
	sim.linearDamping = def.linearDamping;
	sim.angularDamping = def.angularDamping;
	sim.gravityScale = def.gravityScale;
	sim.isBullet = def.isBullet;
	sim.allowFastRotation = def.allowFastRotation;

	int localIndex;
	if ( setId == kAwakeSet )
	{
		localIndex = w->awakeBodies.count;
		F2D_PUSH( w, w->awakeBodies, bodyId );
		BodyState st = identityState();
		st.v = def.linearVelocity;
		st.w = def.angularVelocity;
		F2D_PUSH( w, w->states, st );
	}
	else if ( setId == kStaticSet )
	{
		localIndex = w->staticBodies.count;
		F2D_PUSH( w, w->staticBodies, bodyId );
	}
	else if ( setId == kDisabledSet )
	{
		localIndex = w->disabledBodies.count;
		F2D_PUSH( w, w->disabledBodies, bodyId );
	}
	else
	{
		SolverSet& s = ptr( w, w->sets )[setId];
		localIndex = s.bodyCount;
		setBodyList( w, s )[s.bodyCount++] = bodyId;
	}

	Body& body = ptr( w, w->bodies )[bodyId];
	memset( body.name, 0, sizeof( body.name ) );
	if ( def.name )
	{
		int i = 0;
		while ( i < 31 && def.name[i] != 0 )
		{
			body.name[i] = def.name[i];
			i += 1;
		}
	}
	body.userData = def.userData;
	body.setIndex = setId;
	body.localIndex = localIndex;
	body.generation += 1;
	body.headShapeId = kNull;
	body.shapeCount = 0;
	body.headChainId = kNull;
	body.headContactKey = kNull;
	body.contactCount = 0;
	body.headJointKey = kNull;
	body.jointCount = 0;
	body.islandId = kNull;
	body.islandPrev = kNull;
	body.islandNext = kNull;
	body.bodyMoveIndex = kNull;
	body.id = bodyId;
	body.mass = 0.0f;
	body.inertia = 0.0f;
	body.sleepThreshold = def.sleepThreshold;
	body.sleepTime = 0.0f;
	body.type = def.type;
	body.enableSleep = def.enableSleep;
	body.fixedRotation = def.fixedRotation;
	body.isSpeedCapped = false;
	body.isMarked = false;
	body.colorMask = 0;
	if ( setId >= kAwakeSet )
		createIslandForBody( w, setId, body );
	return bodyId;
}

// body.c:528-631
inline void updateBodyMassData( World* w, Body& body )
{
	BodySim& sim = ptr( w, w->sims )[body.id];
	const Shape* shapes = ptr( w, w->shapes );
	body.mass = 0.0f;
	body.inertia = 0.0f;
	sim.invMass = 0.0f;
	sim.invInertia = 0.0f;
	sim.localCenter = V2{ 0.0f, 0.0f };
	sim.minExtent = kHuge;
	sim.maxExtent = 0.0f;
	if ( body.type != kDynamicBody )
	{
		sim.center = sim.transform.p;
		if ( body.type == kKinematicBody )
		{
			for ( int s = body.headShapeId; s != kNull; s = shapes[s].nextShapeId )
			{
				float mn, mx;
				shapeExtent( shapes[s], V2{ 0.0f, 0.0f }, &mn, &mx );
				sim.minExtent = minf( sim.minExtent, mn );
				sim.maxExtent = maxf( sim.maxExtent, mx );
			}
		}
		return;
	}
	V2 localCenter = { 0.0f, 0.0f };
	int shapeId = body.headShapeId;
	while ( shapeId != kNull )
	{
		const Shape& s = shapes[shapeId];
		shapeId = s.nextShapeId;
		if ( s.density == 0.0f )
			continue;
		MassData md = shapeMass( s );
		body.mass += md.mass;
		localCenter = mulAdd( localCenter, md.mass, md.center );
		body.inertia += md.inertia;
	}
	if ( body.mass > 0.0f )
	{
		sim.invMass = 1.0f / body.mass;
		localCenter = mulSV( sim.invMass, localCenter );
	}
	if ( body.inertia > 0.0f && body.fixedRotation == false )
	{
		body.inertia -= body.mass * dot( localCenter, localCenter );
		sim.invInertia = 1.0f / body.inertia;
	}
	else
	{
		body.inertia = 0.0f;
		sim.invInertia = 0.0f;
	}
	V2 oldCenter = sim.center;
	sim.localCenter = localCenter;
	sim.center = xfPoint( sim.transform, sim.localCenter );
	sim.center0 = sim.center;
	if ( body.setIndex == kAwakeSet )
	{
		BodyState& state = ptr( w, w->states )[body.localIndex];
		V2 deltaLinear = crossSV( state.w, sub( sim.center, oldCenter ) );
		state.v = add( state.v, deltaLinear );
	}
	for ( int s = body.headShapeId; s != kNull; s = shapes[s].nextShapeId )
	{
		float mn, mx;
		shapeExtent( shapes[s], localCenter, &mn, &mx );
		sim.minExtent = minf( sim.minExtent, mn );
		sim.maxExtent = maxf( sim.maxExtent, mx );
	}
}

// shape.c:36-57 + :901-911
inline void createShapeProxy( World* w, Shape& shape, int type, Xf transform, bool forcePairCreation )
{
	Box aabb = inflate( shapeAABB( shape, transform ), kSpeculative );
	shape.aabb = aabb;
	float margin = type == kStaticBody ? kSpeculative : kAabbMargin;
	shape.fatAABB = inflate( aabb, margin );
	shape.proxyKey = bpCreateProxy( w, type, shape.fatAABB, shape.filter.category, shape.id, forcePairCreation );
}

// shape.c:59-161 + :163-186. `geometry` points at the Circle/Capsule/Poly/Segment matching `type`.
inline int createShape( World* w, int bodyId, const ShapeParams& def, const void* geometry, int type )
{
	int shapeId = allocId( w, w->shapeIds );
	if ( shapeId == w->shapes.count )
	{
		Shape empty;
		memset( &empty, 0, sizeof( empty ) );
		F2D_PUSH( w, w->shapes, empty );
	}
	Shape& shape = ptr( w, w->shapes )[shapeId];
	Body& body = ptr( w, w->bodies )[bodyId];
	switch ( type )
	{
		case kCapsule:
			shape.capsule = *static_cast<const Capsule*>( geometry );
			break;
		case kCircle:
			shape.circle = *static_cast<const Circle*>( geometry );
			break;
		case kPolygon:
			shape.polygon = *static_cast<const Poly*>( geometry );
			break;
		case kSegment:
			shape.segment = *static_cast<const Segment*>( geometry );
			break;
		case kChainSegment:
			shape.chainSegment = *static_cast<const ChainSegment*>( geometry );
			break;
		default:
			break;
	}
	shape.id = shapeId;
	shape.bodyId = body.id;
	shape.type = type;
	w->shapeTypeMask |= 1u << type;
	shape.density = def.density;
	shape.friction = def.friction;
	shape.restitution = def.restitution;
	shape.rollingResistance = def.rollingResistance;
	shape.tangentSpeed = def.tangentSpeed;
	shape.userMaterialId = def.userMaterialId;
	shape.filter = def.filter;
	shape.userData = def.userData;
	shape.customColor = def.customColor;
	shape.enlargedAABB = false;
	shape.enableSensorEvents = def.enableSensorEvents;
	shape.enableContactEvents = def.enableContactEvents;
	shape.enableHitEvents = def.enableHitEvents;
	shape.enablePreSolveEvents = def.enablePreSolveEvents;
	shape.proxyKey = kNull;
	shape.localCentroid = shapeCentroid( shape );
	shape.aabb = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
	shape.fatAABB = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
	shape.generation += 1;
	if ( def.enableHitEvents )
		w->hitEventCapable += 1;
	if ( def.enableContactEvents )
		w->contactEventCapable += 1;

	if ( body.setIndex != kDisabledSet )
	{
		Xf transform = ptr( w, w->sims )[bodyId].transform;
		createShapeProxy( w, shape, body.type, transform, def.invokeContactCreation || def.isSensor );
	}
	if ( body.headShapeId != kNull )
		ptr( w, w->shapes )[body.headShapeId].prevShapeId = shapeId;
	shape.prevShapeId = kNull;
	shape.nextShapeId = body.headShapeId;
	body.headShapeId = shapeId;
	body.shapeCount += 1;
	shape.sensorIndex = kNull;
	if ( def.isSensor )
	{
		// shape.c:84-95: a sensor record with two empty overlap lists
		if ( w->sensors.count >= w->sensors.cap )
		{
			setError( w, kErrCapacity, __LINE__ );
		}
		else
		{
			shape.sensorIndex = w->sensors.count;
			Sensor sensor = { shapeId, 0, 0, 0 };
			F2D_PUSH( w, w->sensors, sensor );
			w->sensorRefs.count = 2 * w->sensorOverlapCap * w->sensors.count; // slots in use (kept for re-layout copies)
		}
	}
	if ( def.updateBodyMass )
		updateBodyMassData( w, body );
	return shapeId;
}

struct RevoluteParams // mirrors b2RevoluteJointDef (types.h)
{
	int bodyIdA, bodyIdB;
	V2 localAnchorA, localAnchorB;
	float referenceAngle, targetAngle;
	bool enableSpring;
	float hertz, dampingRatio;
	bool enableLimit;
	float lowerAngle, upperAngle;
	bool enableMotor;
	float maxMotorTorque, motorSpeed, drawSize;
	bool collideConnected;
	uint64_t userData;
};


// ---- joints inside sleeping solver sets (host API paths only) --------------------------------------------------
// A sleeping set is four id lists carved from one block of World::sleepPool with fixed capacities, sized when the island
// fell asleep. The reference's set arrays grow on demand (joint.c:252-287 adds a joint sim to a sleeping set,
// solver_set.c:427-519 merges two sleeping sets); here the block is re-allocated with the room asked for, lists in order.
inline bool growSolverSet( World* w, int setId, int extraBodies, int extraContacts, int extraJoints, int extraIslands )
{
	SolverSet& s = ptr( w, w->sets )[setId];
	const int counts[4] = { s.bodyCount, s.contactCount, s.jointCount, s.islandCount };
	int32_t* lists[4] = { setBodyList( w, s ), setContactList( w, s ), setJointList( w, s ), setIslandList( w, s ) };
	std::vector<int32_t> saved[4];
	for ( int k = 0; k < 4; ++k )
		saved[k].assign( lists[k], lists[k] + counts[k] );
	sleepPoolFree( w, s, setId );
	if ( sleepPoolAlloc( w, s, setId, counts[0] + extraBodies, counts[1] + extraContacts, counts[2] + extraJoints, counts[3] + extraIslands ) ==
		 false )
		return false;
	s.bodyCount = counts[0], s.contactCount = counts[1], s.jointCount = counts[2], s.islandCount = counts[3];
	int32_t* fresh[4] = { setBodyList( w, s ), setContactList( w, s ), setJointList( w, s ), setIslandList( w, s ) };
	for ( int k = 0; k < 4; ++k )
		for ( int i = 0; i < counts[k]; ++i )
			fresh[k][i] = saved[k][i];
	return true;
}

// joint.c:257-268: the joint becomes the last entry of the sleeping set's joint list
inline bool addJointToSleepingSet( World* w, int setId, Joint& joint )
{
	SolverSet& s = ptr( w, w->sets )[setId];
	if ( s.jointCount >= s.jointCap && growSolverSet( w, setId, 0, 0, s.jointCount + 4, 0 ) == false )
		return false;
	joint.setIndex = setId;
	joint.colorIndex = kNull;
	joint.localIndex = s.jointCount;
	setJointList( w, s )[s.jointCount++] = joint.jointId;
	return true;
}

// solver_set.c:427-519 b2MergeSolverSets: the set with fewer bodies is appended to the other one and destroyed
inline bool mergeSolverSets( World* w, int setId1, int setId2 )
{
	SolverSet* sets = ptr( w, w->sets );
	if ( sets[setId1].bodyCount < sets[setId2].bodyCount )
	{
		int t = setId1;
		setId1 = setId2;
		setId2 = t;
	}
	const SolverSet from = sets[setId2];
	if ( growSolverSet( w, setId1, from.bodyCount, from.contactCount, from.jointCount, from.islandCount ) == false )
		return false;
	SolverSet& s1 = sets[setId1];
	SolverSet& s2 = sets[setId2];
	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	Island* islands = ptr( w, w->islands );
	const int32_t* src = setBodyList( w, s2 );
	int32_t* dst = setBodyList( w, s1 );
	for ( int i = 0; i < s2.bodyCount; ++i )
	{
		bodies[src[i]].setIndex = setId1;
		bodies[src[i]].localIndex = s1.bodyCount;
		dst[s1.bodyCount++] = src[i];
	}
	src = setContactList( w, s2 );
	dst = setContactList( w, s1 );
	for ( int i = 0; i < s2.contactCount; ++i )
	{
		contacts[src[i]].setIndex = setId1;
		contacts[src[i]].localIndex = s1.contactCount;
		dst[s1.contactCount++] = src[i];
	}
	src = setJointList( w, s2 );
	dst = setJointList( w, s1 );
	for ( int i = 0; i < s2.jointCount; ++i )
	{
		joints[src[i]].setIndex = setId1;
		joints[src[i]].localIndex = s1.jointCount;
		dst[s1.jointCount++] = src[i];
	}
	src = setIslandList( w, s2 );
	dst = setIslandList( w, s1 );
	for ( int i = 0; i < s2.islandCount; ++i )
	{
		islands[src[i]].setIndex = setId1;
		islands[src[i]].localIndex = s1.islandCount;
		dst[s1.islandCount++] = src[i];
	}
	destroySolverSet( w, setId2 );
	return true;
}

// joint.c:142-300 (b2CreateJoint): id, edge lists, owning set, island link. Returns the joint id.
inline int createJointBase( World* w, int bodyIdA, int bodyIdB, uint64_t userData, float drawSize, int type, bool collideConnected )
{
	Body* bodies = ptr( w, w->bodies );
	Body& bodyA = bodies[bodyIdA];
	Body& bodyB = bodies[bodyIdB];
	int maxSetIndex = maxi( bodyA.setIndex, bodyB.setIndex );

	int jointId = allocId( w, w->jointIds );
	if ( jointId == w->joints.count )
	{
		Joint empty;
		memset( &empty, 0, sizeof( empty ) );
		F2D_PUSH( w, w->joints, empty );
		JointSim emptySim;
		memset( &emptySim, 0, sizeof( emptySim ) );
		F2D_PUSH( w, w->jointSims, emptySim );
	}
	Joint* joints = ptr( w, w->joints );
	Joint& joint = joints[jointId];
	joint.jointId = jointId;
	joint.userData = userData;
	joint.generation += 1;
	joint.setIndex = kNull;
	joint.colorIndex = kNull;
	joint.localIndex = kNull;
	joint.islandId = kNull;
	joint.islandPrev = kNull;
	joint.islandNext = kNull;
	joint.drawSize = drawSize;
	joint.type = type;
	joint.collideConnected = collideConnected;
	joint.isMarked = false;

	joint.edges[0].bodyId = bodyIdA;
	joint.edges[0].prevKey = kNull;
	joint.edges[0].nextKey = bodyA.headJointKey;
	int keyA = ( jointId << 1 ) | 0;
	if ( bodyA.headJointKey != kNull )
		joints[bodyA.headJointKey >> 1].edges[bodyA.headJointKey & 1].prevKey = keyA;
	bodyA.headJointKey = keyA;
	bodyA.jointCount += 1;

	joint.edges[1].bodyId = bodyIdB;
	joint.edges[1].prevKey = kNull;
	joint.edges[1].nextKey = bodyB.headJointKey;
	int keyB = ( jointId << 1 ) | 1;
	if ( bodyB.headJointKey != kNull )
		joints[bodyB.headJointKey >> 1].edges[bodyB.headJointKey & 1].prevKey = keyB;
	bodyB.headJointKey = keyB;
	bodyB.jointCount += 1;

	JointSim& sim = ptr( w, w->jointSims )[jointId];
	memset( &sim, 0, sizeof( sim ) );
	sim.jointId = jointId;
	sim.bodyIdA = bodyIdA;
	sim.bodyIdB = bodyIdB;

	if ( bodyA.setIndex == kDisabledSet || bodyB.setIndex == kDisabledSet )
	{
		joint.setIndex = kDisabledSet;
		joint.localIndex = w->disabledJoints.count;
		F2D_PUSH( w, w->disabledJoints, jointId );
	}
	else if ( bodyA.setIndex == kStaticSet && bodyB.setIndex == kStaticSet )
	{
		joint.setIndex = kStaticSet;
		joint.localIndex = w->staticJoints.count;
		F2D_PUSH( w, w->staticJoints, jointId );
	}
	else if ( bodyA.setIndex == kAwakeSet || bodyB.setIndex == kAwakeSet )
	{
		if ( maxSetIndex >= kFirstSleepingSet )
			wakeSolverSet( w, maxSetIndex );
		joint.setIndex = kAwakeSet;
		addJointToGraph( w, jointId );
	}
	else
	{
		// joint between sleeping and / or static bodies: it goes into the sleeping set, and two different sleeping sets
		// become one (joint.c:250-287)
		bool ok = addJointToSleepingSet( w, maxSetIndex, joint );
		if ( ok && bodyA.setIndex != bodyB.setIndex && bodyA.setIndex >= kFirstSleepingSet && bodyB.setIndex >= kFirstSleepingSet )
			ok = mergeSolverSets( w, bodyA.setIndex, bodyB.setIndex );
		if ( ok == false )
			setError( w, kErrSleepPool, __LINE__ );
	}
	sim.constraintHertz = 60.0f;		  // B2_JOINT_CONSTRAINT_HERTZ constants.h:50
	sim.constraintDampingRatio = 2.0f; // B2_JOINT_CONSTRAINT_DAMPING_RATIO constants.h:53 (value 2 in v3.1.1: joint.c)
	sim.constraintSoftness = Soft{ 0.0f, 1.0f, 0.0f };
	if ( joint.setIndex > kDisabledSet )
		linkJoint( w, joint, true );
	return jointId;
}

// joint.c:302-340 b2DestroyContactsBetweenBodies
inline void destroyContactsBetweenBodies( World* w, int bodyIdA, int bodyIdB )
{
	Body* bodies = ptr( w, w->bodies );
	int contactKey, otherBodyId;
	if ( bodies[bodyIdA].contactCount < bodies[bodyIdB].contactCount )
	{
		contactKey = bodies[bodyIdA].headContactKey;
		otherBodyId = bodyIdB;
	}
	else
	{
		contactKey = bodies[bodyIdB].headContactKey;
		otherBodyId = bodyIdA;
	}
	Contact* contacts = ptr( w, w->contacts );
	while ( contactKey != kNull )
	{
		int contactId = contactKey >> 1;
		int edgeIndex = contactKey & 1;
		Contact& c = contacts[contactId];
		contactKey = c.edges[edgeIndex].nextKey;
		if ( c.edges[edgeIndex ^ 1].bodyId == otherBodyId )
			destroyContact( w, contactId, false );
	}
}

// joint.c b2CreateRevoluteJoint
inline int createRevoluteJoint( World* w, const RevoluteParams& def )
{
	int jointId = createJointBase( w, def.bodyIdA, def.bodyIdB, def.userData, def.drawSize, kRevoluteJoint, def.collideConnected );
	JointSim& sim = ptr( w, w->jointSims )[jointId];
	sim.type = kRevoluteJoint;
	sim.localOriginAnchorA = def.localAnchorA;
	sim.localOriginAnchorB = def.localAnchorB;
	RevoluteJointData& r = sim.revolute;
	memset( &r, 0, sizeof( r ) );
	r.referenceAngle = clampf( def.referenceAngle, -kPi, kPi );
	r.targetAngle = clampf( def.targetAngle, -kPi, kPi );
	r.hertz = def.hertz;
	r.dampingRatio = def.dampingRatio;
	r.lowerAngle = def.lowerAngle;
	r.upperAngle = def.upperAngle;
	r.maxMotorTorque = def.maxMotorTorque;
	r.motorSpeed = def.motorSpeed;
	r.enableSpring = def.enableSpring;
	r.enableLimit = def.enableLimit;
	r.enableMotor = def.enableMotor;
	if ( def.collideConnected == false )
		destroyContactsBetweenBodies( w, def.bodyIdA, def.bodyIdB );
	return jointId;
}

} // namespace f2d
