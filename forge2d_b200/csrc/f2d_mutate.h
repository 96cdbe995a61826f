// forge2d_b200 — host-side structural edits of a world image between steps: destroying bodies / shapes / joints,
// teleporting bodies, re-filtering shapes. Same ordering rules as the creation paths in f2d_create.h: every list edit
// is an append or a swap-remove exactly where the reference appends / swap-removes, so the iteration orders the step
// depends on stay identical to the reference's (B2 = packages/forge2d/third_party/box2d/src).
#pragma once
#include "f2d_create.h"
#include "f2d_distance.h"

namespace f2d
{

// body.c:99-158 b2RemoveBodyFromIsland
inline void removeBodyFromIsland( World* w, Body& body )
{
	if ( body.islandId == kNull )
		return;
	Body* bodies = ptr( w, w->bodies );
	Island& island = ptr( w, w->islands )[body.islandId];
	if ( body.islandPrev != kNull )
		bodies[body.islandPrev].islandNext = body.islandNext;
	if ( body.islandNext != kNull )
		bodies[body.islandNext].islandPrev = body.islandPrev;
	island.bodyCount -= 1;
	if ( island.headBody == body.id )
	{
		island.headBody = body.islandNext;
		if ( island.headBody == kNull )
			destroyIsland( w, island.islandId );
	}
	else if ( island.tailBody == body.id )
	{
		island.tailBody = body.islandPrev;
	}
	body.islandId = kNull;
	body.islandPrev = kNull;
	body.islandNext = kNull;
}

// Swap-removes the body from the id list of the set that owns it (body.c:408-433: b2BodySimArray_RemoveSwap plus the
// moved body's index fix; the awake set also drops the aligned state; an emptied sleeping set is destroyed).
inline void removeBodyFromSet( World* w, Body& body, bool destroyOrphanSet = true )
{
	Body* bodies = ptr( w, w->bodies );
	const int setIndex = body.setIndex, localIndex = body.localIndex;
	if ( setIndex == kAwakeSet )
	{
		int moved = removeSwap( w, w->awakeBodies, localIndex );
		removeSwap( w, w->states, localIndex );
		if ( moved != kNull )
			bodies[ptr( w, w->awakeBodies )[localIndex]].localIndex = localIndex;
	}
	else if ( setIndex == kStaticSet )
	{
		int moved = removeSwap( w, w->staticBodies, localIndex );
		if ( moved != kNull )
			bodies[ptr( w, w->staticBodies )[localIndex]].localIndex = localIndex;
	}
	else if ( setIndex == kDisabledSet )
	{
		int moved = removeSwap( w, w->disabledBodies, localIndex );
		if ( moved != kNull )
			bodies[ptr( w, w->disabledBodies )[localIndex]].localIndex = localIndex;
	}
	else
	{
		SolverSet& set = ptr( w, w->sets )[setIndex];
		int32_t* list = setBodyList( w, set );
		int last = set.bodyCount - 1;
		if ( localIndex != last )
		{
			list[localIndex] = list[last];
			bodies[list[localIndex]].localIndex = localIndex;
		}
		set.bodyCount -= 1;
		if ( set.bodyCount == 0 && destroyOrphanSet )
			destroySolverSet( w, setIndex );
	}
}

// joint.c:702-807 b2DestroyJointInternal
inline void destroyJointInternal( World* w, int jointId, bool wakeBodies )
{
	Joint* joints = ptr( w, w->joints );
	Joint& joint = joints[jointId];
	Body* bodies = ptr( w, w->bodies );
	Edge& edgeA = joint.edges[0];
	Edge& edgeB = joint.edges[1];
	const int idA = edgeA.bodyId, idB = edgeB.bodyId;
	Body& bodyA = bodies[idA];
	Body& bodyB = bodies[idB];

	if ( edgeA.prevKey != kNull )
		joints[edgeA.prevKey >> 1].edges[edgeA.prevKey & 1].nextKey = edgeA.nextKey;
	if ( edgeA.nextKey != kNull )
		joints[edgeA.nextKey >> 1].edges[edgeA.nextKey & 1].prevKey = edgeA.prevKey;
	if ( bodyA.headJointKey == ( ( jointId << 1 ) | 0 ) )
		bodyA.headJointKey = edgeA.nextKey;
	bodyA.jointCount -= 1;

	if ( edgeB.prevKey != kNull )
		joints[edgeB.prevKey >> 1].edges[edgeB.prevKey & 1].nextKey = edgeB.nextKey;
	if ( edgeB.nextKey != kNull )
		joints[edgeB.nextKey >> 1].edges[edgeB.nextKey & 1].prevKey = edgeB.prevKey;
	if ( bodyB.headJointKey == ( ( jointId << 1 ) | 1 ) )
		bodyB.headJointKey = edgeB.nextKey;
	bodyB.jointCount -= 1;

	if ( joint.islandId != kNull )
		unlinkJoint( w, joint );

	const int setIndex = joint.setIndex, localIndex = joint.localIndex;
	if ( setIndex == kAwakeSet )
	{
		removeJointFromGraph( w, idA, idB, joint.colorIndex, localIndex );
	}
	else if ( setIndex == kStaticSet || setIndex == kDisabledSet )
	{
		Arr<int32_t>& list = setIndex == kStaticSet ? w->staticJoints : w->disabledJoints;
		int moved = removeSwap( w, list, localIndex );
		if ( moved != kNull )
			joints[ptr( w, list )[localIndex]].localIndex = localIndex;
	}
	else
	{
		SolverSet& set = ptr( w, w->sets )[setIndex];
		int32_t* list = setJointList( w, set );
		int last = set.jointCount - 1;
		if ( localIndex != last )
		{
			list[localIndex] = list[last];
			joints[list[localIndex]].localIndex = localIndex;
		}
		set.jointCount -= 1;
	}

	joint.setIndex = kNull;
	joint.localIndex = kNull;
	joint.colorIndex = kNull;
	joint.jointId = kNull;
	freeId( w, w->jointIds, jointId );

	if ( wakeBodies )
	{
		wakeBody( w, bodyA );
		wakeBody( w, bodyB );
	}
}

// body.c:160-175 b2DestroyBodyContacts
inline void destroyBodyContacts( World* w, Body& body, bool wakeBodies )
{
	Contact* contacts = ptr( w, w->contacts );
	int edgeKey = body.headContactKey;
	while ( edgeKey != kNull )
	{
		int contactId = edgeKey >> 1;
		int edgeIndex = edgeKey & 1;
		edgeKey = contacts[contactId].edges[edgeIndex].nextKey;
		destroyContact( w, contactId, wakeBodies );
	}
}

// shape.c:913-920 b2DestroyShapeProxy
inline void destroyShapeProxy( World* w, Shape& shape )
{
	if ( shape.proxyKey != kNull )
	{
		bpDestroyProxy( w, shape.proxyKey );
		shape.proxyKey = kNull;
	}
}

// sensor.c:353-390 b2DestroySensor: end events for the current overlaps, then swap-remove the sensor slot (the moved
// sensor's two list blocks travel with it)
inline void destroySensor( World* w, Shape& sensorShape )
{
	const int index = sensorShape.sensorIndex;
	Sensor* sensors = ptr( w, w->sensors );
	ShapeRef* refs = ptr( w, w->sensorRefs );
	const Sensor& sensor = sensors[index];
	const ShapeRef* refs2 = refs + (size_t)( 2 * index + sensor.flip ) * w->sensorOverlapCap;
	for ( int i = 0; i < sensor.count2; ++i )
	{
		SensorEvent ev = { ShapeId{ sensorShape.id + 1, w->worldId, sensorShape.generation },
						   ShapeId{ refs2[i].shapeId + 1, w->worldId, refs2[i].generation } };
		F2D_PUSH_EVENT( w, w->sensorEndEvents[w->endEventArrayIndex], ev );
	}
	const int last = w->sensors.count - 1;
	if ( index != last )
	{
		sensors[index] = sensors[last];
		memcpy( refs + (size_t)( 2 * index ) * w->sensorOverlapCap, refs + (size_t)( 2 * last ) * w->sensorOverlapCap,
				sizeof( ShapeRef ) * 2 * w->sensorOverlapCap );
		ptr( w, w->shapes )[sensors[index].shapeId].sensorIndex = index;
	}
	w->sensors.count -= 1;
	w->sensorRefs.count = 2 * w->sensorOverlapCap * w->sensors.count;
	sensorShape.sensorIndex = kNull;
}

// body.c:343-444 b2DestroyBody (chains cannot be created on this path)
inline void destroyBody( World* w, int bodyId )
{
	Body& body = ptr( w, w->bodies )[bodyId];
	Joint* joints = ptr( w, w->joints );
	const bool wakeBodies = true;
	int edgeKey = body.headJointKey;
	while ( edgeKey != kNull )
	{
		int jointId = edgeKey >> 1;
		int edgeIndex = edgeKey & 1;
		edgeKey = joints[jointId].edges[edgeIndex].nextKey;
		destroyJointInternal( w, jointId, wakeBodies );
	}
	destroyBodyContacts( w, body, wakeBodies );
	Shape* shapes = ptr( w, w->shapes );
	int shapeId = body.headShapeId;
	while ( shapeId != kNull )
	{
		Shape& shape = shapes[shapeId];
		if ( shape.sensorIndex != kNull )
			destroySensor( w, shape );
		destroyShapeProxy( w, shape );
		freeId( w, w->shapeIds, shapeId );
		shape.id = kNull;
		shapeId = shape.nextShapeId;
	}
	removeBodyFromIsland( w, body );
	removeBodyFromSet( w, body );
	freeId( w, w->bodyIds, body.id );
	body.setIndex = kNull;
	body.localIndex = kNull;
	body.id = kNull;
}

// solver_set.c b2TransferBody: the body's id moves from its set's list to the end of the target list (the sim record
// itself lives in a stable slot here). Only the static, disabled and awake sets are targets on this path.
inline void transferBody( World* w, int targetSet, Body& body )
{
	const int sourceSet = body.setIndex;
	removeBodyFromSet( w, body, false );
	if ( targetSet == kAwakeSet )
	{
		body.localIndex = w->awakeBodies.count;
		F2D_PUSH( w, w->awakeBodies, body.id );
		F2D_PUSH( w, w->states, identityState() );
	}
	else
	{
		Arr<int32_t>& list = targetSet == kStaticSet ? w->staticBodies : w->disabledBodies;
		body.localIndex = list.count;
		F2D_PUSH( w, list, body.id );
	}
	(void)sourceSet;
	body.setIndex = targetSet;
}

// solver_set.c b2TransferJoint: add to the target first (the graph picks a colour from the bodies' current types), then
// remove from the source. Returns false only when the sleep pool cannot hold the grown target set.
inline bool transferJoint( World* w, int targetSet, Joint& joint )
{
	Joint* joints = ptr( w, w->joints );
	const int sourceSet = joint.setIndex, localIndex = joint.localIndex, colorIndex = joint.colorIndex;
	if ( targetSet >= kFirstSleepingSet )
	{
		if ( addJointToSleepingSet( w, targetSet, joint ) == false )
			return false;
	}
	else if ( targetSet == kAwakeSet )
	{
		addJointToGraph( w, joint.jointId );
		joint.setIndex = kAwakeSet;
	}
	else
	{
		Arr<int32_t>& list = targetSet == kStaticSet ? w->staticJoints : w->disabledJoints;
		joint.setIndex = targetSet;
		joint.localIndex = list.count;
		joint.colorIndex = kNull;
		F2D_PUSH( w, list, joint.jointId );
	}
	if ( sourceSet == kAwakeSet )
	{
		removeJointFromGraph( w, joint.edges[0].bodyId, joint.edges[1].bodyId, colorIndex, localIndex );
	}
	else if ( sourceSet == kStaticSet || sourceSet == kDisabledSet )
	{
		Arr<int32_t>& list = sourceSet == kStaticSet ? w->staticJoints : w->disabledJoints;
		int moved = removeSwap( w, list, localIndex );
		if ( moved != kNull )
			joints[ptr( w, list )[localIndex]].localIndex = localIndex;
	}
	else
	{
		SolverSet& set = ptr( w, w->sets )[sourceSet];
		int32_t* list = setJointList( w, set );
		int last = set.jointCount - 1;
		if ( localIndex != last )
		{
			list[localIndex] = list[last];
			joints[list[localIndex]].localIndex = localIndex;
		}
		set.jointCount -= 1;
	}
	return true;
}

// body.c:1557-1626 b2Body_Disable
inline bool disableBody( World* w, Body& body )
{
	if ( body.setIndex == kDisabledSet )
		return true;
	destroyBodyContacts( w, body, true );
	removeBodyFromIsland( w, body );
	Shape* shapes = ptr( w, w->shapes );
	for ( int s = body.headShapeId; s != kNull; s = shapes[s].nextShapeId )
		destroyShapeProxy( w, shapes[s] );
	transferBody( w, kDisabledSet, body );
	Joint* joints = ptr( w, w->joints );
	int jointKey = body.headJointKey;
	bool ok = true;
	while ( jointKey != kNull )
	{
		Joint& joint = joints[jointKey >> 1];
		jointKey = joint.edges[jointKey & 1].nextKey;
		if ( joint.setIndex == kDisabledSet )
			continue;
		if ( joint.islandId != kNull )
			unlinkJoint( w, joint );
		ok = transferJoint( w, kDisabledSet, joint ) && ok;
	}
	return ok;
}

// body.c:1628-1720 b2Body_Enable
inline bool enableBody( World* w, Body& body )
{
	if ( body.setIndex != kDisabledSet )
		return true;
	const int setId = body.type == kStaticBody ? kStaticSet : kAwakeSet;
	transferBody( w, setId, body );
	Xf transform = ptr( w, w->sims )[body.id].transform;
	Shape* shapes = ptr( w, w->shapes );
	for ( int s = body.headShapeId; s != kNull; s = shapes[s].nextShapeId )
		createShapeProxy( w, shapes[s], body.type, transform, true );
	if ( setId != kStaticSet )
		createIslandForBody( w, setId, body );
	Joint* joints = ptr( w, w->joints );
	Body* bodies = ptr( w, w->bodies );
	bool ok = true;
	int jointKey = body.headJointKey;
	while ( jointKey != kNull )
	{
		Joint& joint = joints[jointKey >> 1];
		jointKey = joint.edges[jointKey & 1].nextKey;
		const Body& bodyA = bodies[joint.edges[0].bodyId];
		const Body& bodyB = bodies[joint.edges[1].bodyId];
		if ( bodyA.setIndex == kDisabledSet || bodyB.setIndex == kDisabledSet )
			continue;
		int jointSetId;
		if ( bodyA.setIndex == kStaticSet && bodyB.setIndex == kStaticSet )
			jointSetId = kStaticSet;
		else if ( bodyA.setIndex == kStaticSet )
			jointSetId = bodyB.setIndex;
		else
			jointSetId = bodyA.setIndex;
		if ( transferJoint( w, jointSetId, joint ) == false )
		{
			ok = false;
			continue;
		}
		if ( jointSetId != kStaticSet )
			linkJoint( w, joint, false );
	}
	mergeAwakeIslands( w );
	return ok;
}

// body.c:1036-1284 b2Body_SetType
inline bool setBodyType( World* w, Body& body, int type )
{
	const int originalType = body.type;
	if ( originalType == type )
		return true;
	if ( body.setIndex == kDisabledSet )
	{
		body.type = type;
		updateBodyMassData( w, body );
		return true;
	}
	destroyBodyContacts( w, body, false );
	wakeBody( w, body );
	Joint* joints = ptr( w, w->joints );
	Body* bodies = ptr( w, w->bodies );
	Shape* shapes = ptr( w, w->shapes );
	bool ok = true;
	{
		int jointKey = body.headJointKey;
		while ( jointKey != kNull )
		{
			Joint& joint = joints[jointKey >> 1];
			int edgeIndex = jointKey & 1;
			if ( joint.islandId != kNull )
				unlinkJoint( w, joint );
			wakeBody( w, bodies[joint.edges[0].bodyId] );
			wakeBody( w, bodies[joint.edges[1].bodyId] );
			jointKey = joint.edges[edgeIndex].nextKey;
		}
	}
	body.type = type;
	auto recreateProxies = [&]( int proxyType ) {
		Xf transform = ptr( w, w->sims )[body.id].transform;
		int shapeId = body.headShapeId;
		while ( shapeId != kNull )
		{
			Shape& shape = shapes[shapeId];
			shapeId = shape.nextShapeId;
			destroyShapeProxy( w, shape );
			createShapeProxy( w, shape, proxyType, transform, true );
		}
	};
	if ( originalType == kStaticBody )
	{
		transferBody( w, kAwakeSet, body );
		createIslandForBody( w, kAwakeSet, body );
		int jointKey = body.headJointKey;
		while ( jointKey != kNull )
		{
			Joint& joint = joints[jointKey >> 1];
			int edgeIndex = jointKey & 1;
			if ( joint.setIndex == kStaticSet )
			{
				ok = transferJoint( w, kAwakeSet, joint ) && ok;
			}
			else if ( joint.setIndex == kAwakeSet )
			{
				// through the static list and back, so the graph colours it for the bodies' new types
				transferJoint( w, kStaticSet, joint );
				transferJoint( w, kAwakeSet, joint );
			}
			jointKey = joint.edges[edgeIndex].nextKey;
		}
		recreateProxies( type );
	}
	else if ( type == kStaticBody )
	{
		transferBody( w, kStaticSet, body );
		removeBodyFromIsland( w, body );
		ptr( w, w->sims )[body.id].isFast = false;
		int jointKey = body.headJointKey;
		while ( jointKey != kNull )
		{
			Joint& joint = joints[jointKey >> 1];
			int edgeIndex = jointKey & 1;
			jointKey = joint.edges[edgeIndex].nextKey;
			const Body& other = bodies[joint.edges[edgeIndex ^ 1].bodyId];
			if ( joint.setIndex == kDisabledSet )
				continue;
			if ( other.setIndex == kStaticSet )
			{
				transferJoint( w, kStaticSet, joint );
			}
			else
			{
				transferJoint( w, kStaticSet, joint );
				transferJoint( w, kAwakeSet, joint );
			}
		}
		recreateProxies( kStaticBody );
	}
	else
	{
		recreateProxies( type );
	}
	{
		int jointKey = body.headJointKey;
		while ( jointKey != kNull )
		{
			Joint& joint = joints[jointKey >> 1];
			int edgeIndex = jointKey & 1;
			jointKey = joint.edges[edgeIndex].nextKey;
			const Body& other = bodies[joint.edges[edgeIndex ^ 1].bodyId];
			if ( other.setIndex == kDisabledSet )
				continue;
			if ( body.type == kStaticBody && other.type == kStaticBody )
				continue;
			linkJoint( w, joint, false );
		}
		mergeAwakeIslands( w );
	}
	updateBodyMassData( w, body );
	return ok;
}

// shape.c:230-316 b2DestroyShapeInternal + :318-337 b2DestroyShape
inline void destroyShape( World* w, int shapeId, bool updateBodyMass )
{
	Shape* shapes = ptr( w, w->shapes );
	Shape& shape = shapes[shapeId];
	Body& body = ptr( w, w->bodies )[shape.bodyId];
	if ( shape.prevShapeId != kNull )
		shapes[shape.prevShapeId].nextShapeId = shape.nextShapeId;
	if ( shape.nextShapeId != kNull )
		shapes[shape.nextShapeId].prevShapeId = shape.prevShapeId;
	if ( shapeId == body.headShapeId )
		body.headShapeId = shape.nextShapeId;
	body.shapeCount -= 1;
	destroyShapeProxy( w, shape );
	Contact* contacts = ptr( w, w->contacts );
	int contactKey = body.headContactKey;
	while ( contactKey != kNull )
	{
		int contactId = contactKey >> 1;
		int edgeIndex = contactKey & 1;
		const Contact& c = contacts[contactId];
		contactKey = c.edges[edgeIndex].nextKey;
		if ( c.shapeIdA == shapeId || c.shapeIdB == shapeId )
			destroyContact( w, contactId, true );
	}
	if ( shape.sensorIndex != kNull )
		destroySensor( w, shape );
	freeId( w, w->shapeIds, shapeId );
	shape.id = kNull;
	if ( updateBodyMass )
		updateBodyMassData( w, body );
}

// shape.c:36-57 b2UpdateShapeAABBs
inline void updateShapeAABBs( Shape& shape, Xf transform, int proxyType )
{
	Box aabb = inflate( shapeAABB( shape, transform ), kSpeculative );
	shape.aabb = aabb;
	float margin = proxyType == kStaticBody ? kSpeculative : kAabbMargin;
	shape.fatAABB = inflate( aabb, margin );
}

// shape.c:1185-1233 b2ResetProxy
inline void resetProxy( World* w, Shape& shape, bool wakeBodies, bool destroyProxy )
{
	Body& body = ptr( w, w->bodies )[shape.bodyId];
	const int shapeId = shape.id;
	Contact* contacts = ptr( w, w->contacts );
	int contactKey = body.headContactKey;
	while ( contactKey != kNull )
	{
		int contactId = contactKey >> 1;
		int edgeIndex = contactKey & 1;
		const Contact& c = contacts[contactId];
		contactKey = c.edges[edgeIndex].nextKey;
		if ( c.shapeIdA == shapeId || c.shapeIdB == shapeId )
			destroyContact( w, contactId, wakeBodies );
	}
	Xf transform = ptr( w, w->sims )[body.id].transform;
	if ( shape.proxyKey != kNull )
	{
		int type = proxyType( shape.proxyKey );
		updateShapeAABBs( shape, transform, type );
		if ( destroyProxy )
		{
			bpDestroyProxy( w, shape.proxyKey );
			shape.proxyKey = bpCreateProxy( w, type, shape.fatAABB, shape.filter.category, shapeId, true );
		}
		else
		{
			bpMoveProxy( w, shape.proxyKey, shape.fatAABB );
		}
	}
	else
	{
		updateShapeAABBs( shape, transform, body.type );
	}
}

// body.c:688-741 b2Body_SetTransform
inline void setBodyTransform( World* w, Body& body, V2 position, Rot rotation )
{
	BodySim& sim = ptr( w, w->sims )[body.id];
	sim.transform.p = position;
	sim.transform.q = rotation;
	sim.center = xfPoint( sim.transform, sim.localCenter );
	sim.rotation0 = sim.transform.q;
	sim.center0 = sim.center;
	Xf transform = sim.transform;
	Shape* shapes = ptr( w, w->shapes );
	int shapeId = body.headShapeId;
	while ( shapeId != kNull )
	{
		Shape& shape = shapes[shapeId];
		Box aabb = inflate( shapeAABB( shape, transform ), kSpeculative );
		shape.aabb = aabb;
		if ( boxContains( shape.fatAABB, aabb ) == false )
		{
			shape.fatAABB = inflate( aabb, kAabbMargin );
			if ( shape.proxyKey != kNull )
				bpMoveProxy( w, shape.proxyKey, shape.fatAABB );
		}
		shapeId = shape.nextShapeId;
	}
}

// body.c:28-36 b2LimitVelocity
inline void limitVelocity( BodyState& state, float maxLinearSpeed )
{
	float v2 = dot( state.v, state.v );
	if ( v2 > maxLinearSpeed * maxLinearSpeed )
		state.v = mulSV( maxLinearSpeed / sqrtf( v2 ), state.v );
}

// geometry.c:456-504 point tests
inline bool pointInShape( const Shape& shape, V2 localPoint )
{
	switch ( shape.type )
	{
		case kCircle:
			return distanceSq( localPoint, shape.circle.center ) <= shape.circle.radius * shape.circle.radius;
		case kCapsule:
		{
			float rr = shape.capsule.radius * shape.capsule.radius;
			V2 p1 = shape.capsule.c1, p2 = shape.capsule.c2;
			V2 d = sub( p2, p1 );
			float dd = dot( d, d );
			if ( dd == 0.0f )
				return distanceSq( localPoint, p1 ) <= rr;
			float t = dot( sub( localPoint, p1 ), d ) / dd;
			t = clampf( t, 0.0f, 1.0f );
			V2 c = mulAdd( p1, t, d );
			return distanceSq( localPoint, c ) <= rr;
		}
		case kPolygon:
		{
			ShapeProxy proxyA = makeProxy( shape.polygon.v, shape.polygon.count, 0.0f );
			ShapeProxy proxyB = makeProxy( &localPoint, 1, 0.0f );
			SimplexCache cache;
			memset( &cache, 0, sizeof( cache ) );
			Xf identity = { { 0.0f, 0.0f }, { 1.0f, 0.0f } };
			DistanceOutput out = shapeDistance( proxyA, proxyB, identity, identity, false, &cache );
			return out.distance <= shape.polygon.radius;
		}
		default:
			return false;
	}
}

} // namespace f2d
