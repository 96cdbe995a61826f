// forge2d_b200 — the rest of the b2* surface forge2d's FFI backend binds: body / shape / joint accessors
// and mutators, destruction, the seven other joint types. Included by f2d_capi.inl inside its extern "C" block.
// Everything here edits or reads the HOST image (see hostImage / mutableImage): the next b2World_Step uploads it.
// Reference behaviour is cited per function (B2 = packages/forge2d/third_party/box2d/src).

// ---------------------------------------------------------------------------------------------------------------- bodies
static BodySim* simOf( HostWorld* hw, const Body* b ) { return ptr( hw->img, hw->img->sims ) + b->id; }
static BodyState* stateOf( HostWorld* hw, const Body* b )
{
	return b->setIndex == kAwakeSet ? ptr( hw->img, hw->img->states ) + b->localIndex : nullptr;
}
// Mutators refuse to run while a step is in flight (world.c:62-74 b2GetWorldLocked)
static Body* writableBody( b2BodyId id, HostWorld** hw )
{
	Body* b = bodyFromId( id, hw, true );
	if ( b == nullptr || ( *hw )->img->locked )
		return nullptr;
	return b;
}

void b2DestroyBody( b2BodyId bodyId ) // body.c:343-444
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	// body.c:392-405: the attached chain records go with the body (their segment shapes are destroyed as shapes)
	for ( int chainId = b->headChainId; chainId != kNull; )
	{
		HostWorld::Chain& chain = hw->chains[chainId];
		freeId( hw->img, hw->img->chainIds, chainId );
		chain.id = kNull;
		chain.shapeIndices.clear();
		chain.materials.clear();
		chainId = chain.nextChainId;
	}
	destroyBody( hw->img, b->id );
}
b2Vec2 b2Body_GetLocalPoint( b2BodyId bodyId, b2Vec2 worldPoint ) // body.c:656-662
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	if ( b == nullptr )
		return b2Vec2{ 0, 0 };
	Xf t = simOf( hw, b )->transform;
	V2 r = invRotate( t.q, sub( V2{ worldPoint.x, worldPoint.y }, t.p ) );
	return b2Vec2{ r.x, r.y };
}
b2Vec2 b2Body_GetWorldPoint( b2BodyId bodyId, b2Vec2 localPoint ) // body.c:664-670
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	if ( b == nullptr )
		return b2Vec2{ 0, 0 };
	V2 r = xfPoint( simOf( hw, b )->transform, V2{ localPoint.x, localPoint.y } );
	return b2Vec2{ r.x, r.y };
}
b2Vec2 b2Body_GetLocalVector( b2BodyId bodyId, b2Vec2 worldVector ) // body.c:672-678
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	if ( b == nullptr )
		return b2Vec2{ 0, 0 };
	V2 r = invRotate( simOf( hw, b )->transform.q, V2{ worldVector.x, worldVector.y } );
	return b2Vec2{ r.x, r.y };
}
b2Vec2 b2Body_GetWorldVector( b2BodyId bodyId, b2Vec2 localVector ) // body.c:680-686
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	if ( b == nullptr )
		return b2Vec2{ 0, 0 };
	V2 r = rotate( simOf( hw, b )->transform.q, V2{ localVector.x, localVector.y } );
	return b2Vec2{ r.x, r.y };
}
void b2Body_SetTransform( b2BodyId bodyId, b2Vec2 position, b2Rot rotation ) // body.c:688-741
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	setBodyTransform( hw->img, *b, V2{ position.x, position.y }, Rot{ rotation.c, rotation.s } );
}
void b2Body_ApplyForce( b2BodyId bodyId, b2Vec2 force, b2Vec2 point, bool wake ) // body.c:900-916
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( bodyId, &hw, wake );
	if ( b == nullptr )
		return;
	if ( wake && b->setIndex >= kFirstSleepingSet )
		wakeBody( hw->img, *b );
	if ( b->setIndex == kAwakeSet )
	{
		BodySim* s = simOf( hw, b );
		V2 f = { force.x, force.y };
		s->force = add( s->force, f );
		s->torque += cross( sub( V2{ point.x, point.y }, s->center ), f );
		touchRange( *hw, &s->force, 12 ); // force, torque
	}
}
void b2Body_ApplyForceToCenter( b2BodyId bodyId, b2Vec2 force, bool wake ) // body.c:918-933
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( bodyId, &hw, wake );
	if ( b == nullptr )
		return;
	if ( wake && b->setIndex >= kFirstSleepingSet )
		wakeBody( hw->img, *b );
	if ( b->setIndex == kAwakeSet )
	{
		BodySim* s = simOf( hw, b );
		s->force = add( s->force, V2{ force.x, force.y } );
		touchRange( *hw, &s->force, 8 );
	}
}
void b2Body_ApplyTorque( b2BodyId bodyId, float torque, bool wake ) // body.c:935-950
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( bodyId, &hw, wake );
	if ( b == nullptr )
		return;
	if ( wake && b->setIndex >= kFirstSleepingSet )
		wakeBody( hw->img, *b );
	if ( b->setIndex == kAwakeSet )
	{
		simOf( hw, b )->torque += torque;
		touchRange( *hw, &simOf( hw, b )->torque, 4 );
	}
}
void b2Body_ApplyLinearImpulse( b2BodyId bodyId, b2Vec2 impulse, b2Vec2 point, bool wake ) // body.c:952-973
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( bodyId, &hw, wake );
	if ( b == nullptr )
		return;
	if ( wake && b->setIndex >= kFirstSleepingSet )
		wakeBody( hw->img, *b );
	if ( b->setIndex == kAwakeSet )
	{
		BodyState* st = stateOf( hw, b );
		BodySim* s = simOf( hw, b );
		V2 P = { impulse.x, impulse.y };
		st->v = mulAdd( st->v, s->invMass, P );
		st->w += s->invInertia * cross( sub( V2{ point.x, point.y }, s->center ), P );
		limitVelocity( *st, hw->img->maxLinearSpeed );
		touchRange( *hw, st, sizeof( BodyState ) );
	}
}
void b2Body_ApplyLinearImpulseToCenter( b2BodyId bodyId, b2Vec2 impulse, bool wake ) // body.c:975-995
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( bodyId, &hw, wake );
	if ( b == nullptr )
		return;
	if ( wake && b->setIndex >= kFirstSleepingSet )
		wakeBody( hw->img, *b );
	if ( b->setIndex == kAwakeSet )
	{
		BodyState* st = stateOf( hw, b );
		st->v = mulAdd( st->v, simOf( hw, b )->invMass, V2{ impulse.x, impulse.y } );
		limitVelocity( *st, hw->img->maxLinearSpeed );
		touchRange( *hw, st, sizeof( BodyState ) );
	}
}
void b2Body_ApplyAngularImpulse( b2BodyId bodyId, float impulse, bool wake ) // body.c:997-1020
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( bodyId, &hw, wake );
	if ( b == nullptr )
		return;
	if ( wake && b->setIndex >= kFirstSleepingSet )
		wakeBody( hw->img, *b );
	if ( b->setIndex == kAwakeSet )
	{
		stateOf( hw, b )->w += simOf( hw, b )->invInertia * impulse;
		touchRange( *hw, stateOf( hw, b ), sizeof( BodyState ) );
	}
}
void b2Body_SetName( b2BodyId bodyId, const char* name ) // body.c:1286-1304
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, true );
	if ( b == nullptr )
		return;
	memset( b->name, 0, sizeof( b->name ) );
	if ( name != nullptr )
	{
		for ( int i = 0; i < 31 && name[i] != 0; ++i )
			b->name[i] = name[i];
	}
}
const char* b2Body_GetName( b2BodyId bodyId ) // body.c:1306-1311 (pointer into the world's memory, valid until the next call)
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? b->name : "";
}
void b2Body_SetUserData( b2BodyId bodyId, void* userData )
{
	Body* b = bodyFromId( bodyId, nullptr, true );
	if ( b )
		b->userData = (uint64_t)(uintptr_t)userData;
}
void* b2Body_GetUserData( b2BodyId bodyId )
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? (void*)(uintptr_t)b->userData : nullptr;
}
void b2Body_SetMassData( b2BodyId bodyId, b2MassData massData ) // body.c:1357-1382
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	BodySim* s = simOf( hw, b );
	b->mass = massData.mass;
	b->inertia = massData.rotationalInertia;
	s->localCenter = V2{ massData.center.x, massData.center.y };
	V2 center = xfPoint( s->transform, s->localCenter );
	s->center = center;
	s->center0 = center;
	s->invMass = b->mass > 0.0f ? 1.0f / b->mass : 0.0f;
	s->invInertia = b->inertia > 0.0f ? 1.0f / b->inertia : 0.0f;
}
b2MassData b2Body_GetMassData( b2BodyId bodyId ) // body.c:1384-1391
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	b2MassData m = { 0.0f, { 0.0f, 0.0f }, 0.0f };
	if ( b )
	{
		const BodySim* s = simOf( hw, b );
		m.mass = b->mass;
		m.center = b2Vec2{ s->localCenter.x, s->localCenter.y };
		m.rotationalInertia = b->inertia;
	}
	return m;
}
void b2Body_ApplyMassFromShapes( b2BodyId bodyId ) // body.c:1393-1403
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b )
		updateBodyMassData( hw->img, *b );
}
void b2Body_SetLinearDamping( b2BodyId bodyId, float linearDamping ) // body.c:1405-1418
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b )
		simOf( hw, b )->linearDamping = linearDamping;
}
float b2Body_GetLinearDamping( b2BodyId bodyId )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	return b ? simOf( hw, b )->linearDamping : 0.0f;
}
void b2Body_SetAngularDamping( b2BodyId bodyId, float angularDamping ) // body.c:1428-1441
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b )
		simOf( hw, b )->angularDamping = angularDamping;
}
float b2Body_GetAngularDamping( b2BodyId bodyId )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	return b ? simOf( hw, b )->angularDamping : 0.0f;
}
void b2Body_SetGravityScale( b2BodyId bodyId, float gravityScale ) // body.c:1451-1465
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b )
		simOf( hw, b )->gravityScale = gravityScale;
}
float b2Body_GetGravityScale( b2BodyId bodyId )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	return b ? simOf( hw, b )->gravityScale : 0.0f;
}
void b2Body_SetAwake( b2BodyId bodyId, bool awake ) // body.c:1483-1508
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	World* w = hw->img;
	if ( awake && b->setIndex >= kFirstSleepingSet )
	{
		wakeBody( w, *b );
	}
	else if ( awake == false && b->setIndex == kAwakeSet )
	{
		int islandId = b->islandId;
		if ( ptr( w, w->islands )[islandId].constraintRemoveCount > 0 )
			splitIsland( w, islandId );
		// splitting relabels the body (the island it started in is gone)
		trySleepIsland( w, ptr( w, w->bodies )[bodyId.index1 - 1].islandId );
	}
}
bool b2Body_IsEnabled( b2BodyId bodyId )
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? b->setIndex != kDisabledSet : false;
}
bool b2Body_IsSleepEnabled( b2BodyId bodyId )
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? b->enableSleep : false;
}
void b2Body_SetSleepThreshold( b2BodyId bodyId, float sleepThreshold )
{
	Body* b = bodyFromId( bodyId, nullptr, true );
	if ( b )
		b->sleepThreshold = sleepThreshold;
}
float b2Body_GetSleepThreshold( b2BodyId bodyId )
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? b->sleepThreshold : 0.0f;
}
void b2Body_EnableSleep( b2BodyId bodyId, bool enableSleep ) // body.c:1538-1553
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	b->enableSleep = enableSleep;
	if ( enableSleep == false )
		wakeBody( hw->img, *b );
}
void b2Body_SetFixedRotation( b2BodyId bodyId, bool flag ) // body.c:1722-1742
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr || b->fixedRotation == flag )
		return;
	b->fixedRotation = flag;
	BodyState* st = stateOf( hw, b );
	if ( st )
		st->w = 0.0f;
	updateBodyMassData( hw->img, *b );
}
bool b2Body_IsFixedRotation( b2BodyId bodyId )
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? b->fixedRotation : false;
}
void b2Body_SetBullet( b2BodyId bodyId, bool flag ) // body.c:1751-1762
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b )
		simOf( hw, b )->isBullet = flag;
}
bool b2Body_IsBullet( b2BodyId bodyId )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	return b ? simOf( hw, b )->isBullet : false;
}
static void growEventArrays( HostWorld* hw )
{
	reserve( *hw, 0, 0, 0, 0 ); // event arrays grow with the first event-enabled shape
}
void b2Body_EnableContactEvents( b2BodyId bodyId, bool flag ) // body.c:1772-1783
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, true );
	if ( b == nullptr )
		return;
	World* w = hw->img;
	for ( int s = b->headShapeId; s != kNull; s = ptr( w, w->shapes )[s].nextShapeId )
	{
		Shape& shape = ptr( w, w->shapes )[s];
		if ( flag && shape.enableContactEvents == false )
			w->contactEventCapable += 1;
		shape.enableContactEvents = flag;
	}
	growEventArrays( hw );
}
void b2Body_EnableHitEvents( b2BodyId bodyId, bool flag ) // body.c:1785-1796
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, true );
	if ( b == nullptr )
		return;
	World* w = hw->img;
	for ( int s = b->headShapeId; s != kNull; s = ptr( w, w->shapes )[s].nextShapeId )
	{
		Shape& shape = ptr( w, w->shapes )[s];
		if ( flag && shape.enableHitEvents == false )
			w->hitEventCapable += 1;
		shape.enableHitEvents = flag;
	}
	growEventArrays( hw );
}
b2WorldId b2Body_GetWorld( b2BodyId bodyId ) // body.c:1798-1802
{
	HostWorld* hw = worldFromIndex0( bodyId.world0 );
	return hw ? b2WorldId{ (uint16_t)( bodyId.world0 + 1 ), hw->generation } : b2WorldId{ 0, 0 };
}
int b2Body_GetShapes( b2BodyId bodyId, b2ShapeId* shapeArray, int capacity ) // body.c:1811-1828
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	if ( b == nullptr )
		return 0;
	World* w = hw->img;
	int n = 0;
	for ( int s = b->headShapeId; s != kNull && n < capacity; s = ptr( w, w->shapes )[s].nextShapeId )
	{
		const Shape& shape = ptr( w, w->shapes )[s];
		shapeArray[n++] = b2ShapeId{ shape.id + 1, bodyId.world0, shape.generation };
	}
	return n;
}
int b2Body_GetJointCount( b2BodyId bodyId )
{
	Body* b = bodyFromId( bodyId, nullptr, false );
	return b ? b->jointCount : 0;
}
int b2Body_GetJoints( b2BodyId bodyId, b2JointId* jointArray, int capacity ) // body.c:1837-1859
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, false );
	if ( b == nullptr )
		return 0;
	World* w = hw->img;
	int n = 0;
	int key = b->headJointKey;
	while ( key != kNull && n < capacity )
	{
		const Joint& j = ptr( w, w->joints )[key >> 1];
		jointArray[n++] = b2JointId{ ( key >> 1 ) + 1, bodyId.world0, j.generation };
		key = j.edges[key & 1].nextKey;
	}
	return n;
}
static void reportSleepingJointTarget( World* w, bool ok )
{
	if ( ok )
		return;
	reportError( "forge2d_b200: the sleep pool cannot hold a sleeping solver set that has to grow for a joint" );
	setError( w, kErrSleepPool, __LINE__ );
}
void b2Body_SetType( b2BodyId bodyId, b2BodyType type ) // body.c:1036-1284
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	reportSleepingJointTarget( hw->img, setBodyType( hw->img, *b, (int)type ) );
}
void b2Body_Disable( b2BodyId bodyId ) // body.c:1557-1626
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	reportSleepingJointTarget( hw->img, disableBody( hw->img, *b ) );
}
void b2Body_Enable( b2BodyId bodyId ) // body.c:1628-1720
{
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return;
	reportSleepingJointTarget( hw->img, enableBody( hw->img, *b ) );
}

// ---------------------------------------------------------------------------------------------------------------- shapes
static Shape* writableShape( b2ShapeId id, HostWorld** hw )
{
	Shape* s = shapeFromId( id, hw );
	if ( s == nullptr || ( *hw )->img->locked )
		return nullptr;
	int index = s->id;
	mutableImage( **hw );
	return ptr( ( *hw )->img, ( *hw )->img->shapes ) + index;
}
void b2DestroyShape( b2ShapeId shapeId, bool updateBodyMass ) // shape.c:318-337
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s )
		destroyShape( hw->img, s->id, updateBodyMass );
}
b2WorldId b2Shape_GetWorld( b2ShapeId shapeId )
{
	HostWorld* hw = worldFromIndex0( shapeId.world0 );
	return hw ? b2WorldId{ (uint16_t)( shapeId.world0 + 1 ), hw->generation } : b2WorldId{ 0, 0 };
}
void b2Shape_SetUserData( b2ShapeId shapeId, void* userData )
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s )
		s->userData = (uint64_t)(uintptr_t)userData;
}
void* b2Shape_GetUserData( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? (void*)(uintptr_t)s->userData : nullptr;
}
bool b2Shape_IsSensor( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->sensorIndex != kNull : false;
}
bool b2Shape_TestPoint( b2ShapeId shapeId, b2Vec2 point ) // shape.c:979-1002
{
	HostWorld* hw = nullptr;
	Shape* s = shapeFromId( shapeId, &hw );
	if ( s == nullptr )
		return false;
	Xf t = ptr( hw->img, hw->img->sims )[s->bodyId].transform;
	V2 local = invRotate( t.q, sub( V2{ point.x, point.y }, t.p ) );
	return pointInShape( *s, local );
}
void b2Shape_SetDensity( b2ShapeId shapeId, float density, bool updateBodyMass ) // shape.c:1055-1079
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s == nullptr || density == s->density )
		return;
	s->density = density;
	if ( updateBodyMass )
		updateBodyMassData( hw->img, ptr( hw->img, hw->img->bodies )[s->bodyId] );
}
float b2Shape_GetDensity( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->density : 0.0f;
}
void b2Shape_SetFriction( b2ShapeId shapeId, float friction ) // shape.c:1088-1101
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s )
		s->friction = friction;
}
float b2Shape_GetFriction( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->friction : 0.0f;
}
void b2Shape_SetRestitution( b2ShapeId shapeId, float restitution ) // shape.c:1110-1123
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s )
		s->restitution = restitution;
}
float b2Shape_GetRestitution( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->restitution : 0.0f;
}
b2Filter b2Shape_GetFilter( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? b2Filter{ s->filter.category, s->filter.mask, s->filter.group } : b2Filter{ 0, 0, 0 };
}
void b2Shape_SetFilter( b2ShapeId shapeId, b2Filter filter ) // shape.c:1235-1261
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s == nullptr )
		return;
	if ( filter.maskBits == s->filter.mask && filter.categoryBits == s->filter.category && filter.groupIndex == s->filter.group )
		return;
	bool destroyProxy = filter.categoryBits != s->filter.category;
	s->filter = Filter{ filter.categoryBits, filter.maskBits, filter.groupIndex };
	resetProxy( hw->img, *s, true, destroyProxy );
}
void b2Shape_EnableSensorEvents( b2ShapeId shapeId, bool flag ) // shape.c:1263-1273
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s )
		s->enableSensorEvents = flag;
}
bool b2Shape_AreSensorEventsEnabled( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->enableSensorEvents : false;
}
void b2Shape_EnableContactEvents( b2ShapeId shapeId, bool flag ) // shape.c:1282-1292
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s == nullptr )
		return;
	if ( flag && s->enableContactEvents == false )
		hw->img->contactEventCapable += 1;
	s->enableContactEvents = flag;
	growEventArrays( hw );
}
bool b2Shape_AreContactEventsEnabled( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->enableContactEvents : false;
}
void b2Shape_EnablePreSolveEvents( b2ShapeId shapeId, bool flag ) // shape.c:1301-1311
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s )
		s->enablePreSolveEvents = flag;
}
bool b2Shape_ArePreSolveEventsEnabled( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->enablePreSolveEvents : false;
}
void b2Shape_EnableHitEvents( b2ShapeId shapeId, bool flag ) // shape.c:1320-1330
{
	HostWorld* hw = nullptr;
	Shape* s = writableShape( shapeId, &hw );
	if ( s == nullptr )
		return;
	if ( flag && s->enableHitEvents == false )
		hw->img->hitEventCapable += 1;
	s->enableHitEvents = flag;
	growEventArrays( hw );
}
bool b2Shape_AreHitEventsEnabled( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? s->enableHitEvents : false;
}
b2ShapeType b2Shape_GetType( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	return s ? (b2ShapeType)s->type : b2_circleShape;
}
b2Circle b2Shape_GetCircle( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	b2Circle c = { { 0, 0 }, 0 };
	if ( s && s->type == kCircle )
		c = b2Circle{ { s->circle.center.x, s->circle.center.y }, s->circle.radius };
	return c;
}
b2Segment b2Shape_GetSegment( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	b2Segment g = { { 0, 0 }, { 0, 0 } };
	if ( s && s->type == kSegment )
		g = b2Segment{ { s->segment.p1.x, s->segment.p1.y }, { s->segment.p2.x, s->segment.p2.y } };
	return g;
}
b2ChainSegment b2Shape_GetChainSegment( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	b2ChainSegment g;
	memset( &g, 0, sizeof( g ) );
	if ( s && s->type == kChainSegment )
	{
		const ChainSegment& c = s->chainSegment;
		g.ghost1 = b2Vec2{ c.ghost1.x, c.ghost1.y };
		g.segment = b2Segment{ { c.segment.p1.x, c.segment.p1.y }, { c.segment.p2.x, c.segment.p2.y } };
		g.ghost2 = b2Vec2{ c.ghost2.x, c.ghost2.y };
		g.chainId = c.chainId;
	}
	return g;
}
b2Capsule b2Shape_GetCapsule( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	b2Capsule c = { { 0, 0 }, { 0, 0 }, 0 };
	if ( s && s->type == kCapsule )
		c = b2Capsule{ { s->capsule.c1.x, s->capsule.c1.y }, { s->capsule.c2.x, s->capsule.c2.y }, s->capsule.radius };
	return c;
}
b2Polygon b2Shape_GetPolygon( b2ShapeId shapeId )
{
	Shape* s = shapeFromId( shapeId, nullptr );
	b2Polygon p;
	memset( &p, 0, sizeof( p ) );
	if ( s && s->type == kPolygon )
		memcpy( &p, &s->polygon, sizeof( p ) );
	return p;
}

// ---------------------------------------------------------------------------------------------------------------- joints
static Joint* jointFromId( b2JointId id, HostWorld** outWorld, bool forWrite )
{
	HostWorld* hw = worldFromIndex0( id.world0 );
	if ( hw == nullptr )
		return nullptr;
	World* w = forWrite ? mutableImage( *hw ) : hostImage( *hw );
	int index = id.index1 - 1;
	if ( index < 0 || index >= w->joints.count )
		return nullptr;
	Joint* j = ptr( w, w->joints ) + index;
	if ( j->jointId != index || j->generation != id.generation )
		return nullptr;
	if ( outWorld )
		*outWorld = hw;
	return j;
}
// joint.c:122-144 b2GetJointSimCheckType (a type mismatch is an assert there; here the accessor becomes a no-op)
static JointSim* jointSimOfType( b2JointId id, int type, HostWorld** outWorld, bool forWrite )
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( id, &hw, forWrite );
	if ( j == nullptr || j->type != type )
		return nullptr;
	if ( outWorld )
		*outWorld = hw;
	return ptr( hw->img, hw->img->jointSims ) + j->jointId;
}
static b2BodyId makeBodyId( World* w, int bodyId )
{
	return b2BodyId{ bodyId + 1, w->worldId, ptr( w, w->bodies )[bodyId].generation };
}
void b2DestroyJoint( b2JointId jointId ) // joint.c:809-822
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, true );
	if ( j == nullptr || hw->img->locked )
		return;
	destroyJointInternal( hw->img, j->jointId, true );
}
b2JointType b2Joint_GetType( b2JointId jointId )
{
	Joint* j = jointFromId( jointId, nullptr, false );
	return j ? (b2JointType)j->type : b2_distanceJoint;
}
b2BodyId b2Joint_GetBodyA( b2JointId jointId )
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, false );
	return j ? makeBodyId( hw->img, j->edges[0].bodyId ) : b2BodyId{ 0, 0, 0 };
}
b2BodyId b2Joint_GetBodyB( b2JointId jointId )
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, false );
	return j ? makeBodyId( hw->img, j->edges[1].bodyId ) : b2BodyId{ 0, 0, 0 };
}
b2WorldId b2Joint_GetWorld( b2JointId jointId )
{
	HostWorld* hw = worldFromIndex0( jointId.world0 );
	return hw ? b2WorldId{ (uint16_t)( jointId.world0 + 1 ), hw->generation } : b2WorldId{ 0, 0 };
}
b2Vec2 b2Joint_GetLocalAnchorA( b2JointId jointId )
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, false );
	if ( j == nullptr )
		return b2Vec2{ 0, 0 };
	V2 a = ptr( hw->img, hw->img->jointSims )[j->jointId].localOriginAnchorA;
	return b2Vec2{ a.x, a.y };
}
b2Vec2 b2Joint_GetLocalAnchorB( b2JointId jointId )
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, false );
	if ( j == nullptr )
		return b2Vec2{ 0, 0 };
	V2 a = ptr( hw->img, hw->img->jointSims )[j->jointId].localOriginAnchorB;
	return b2Vec2{ a.x, a.y };
}
void b2Joint_SetLocalAnchorA( b2JointId jointId, b2Vec2 localAnchor ) // joint.c:851-859
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, true );
	if ( j )
		ptr( hw->img, hw->img->jointSims )[j->jointId].localOriginAnchorA = V2{ localAnchor.x, localAnchor.y };
}
void b2Joint_SetLocalAnchorB( b2JointId jointId, b2Vec2 localAnchor ) // joint.c:869-877
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, true );
	if ( j )
		ptr( hw->img, hw->img->jointSims )[j->jointId].localOriginAnchorB = V2{ localAnchor.x, localAnchor.y };
}
void b2Joint_SetCollideConnected( b2JointId jointId, bool shouldCollide ) // joint.c:979-1022
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, true );
	if ( j == nullptr || hw->img->locked || j->collideConnected == shouldCollide )
		return;
	World* w = hw->img;
	j->collideConnected = shouldCollide;
	const Body& bodyA = ptr( w, w->bodies )[j->edges[0].bodyId];
	const Body& bodyB = ptr( w, w->bodies )[j->edges[1].bodyId];
	if ( shouldCollide )
	{
		// the broadphase must look for pairs again: re-buffer the proxies of the body with fewer shapes
		int shapeId = bodyA.shapeCount < bodyB.shapeCount ? bodyA.headShapeId : bodyB.headShapeId;
		while ( shapeId != kNull )
		{
			const Shape& shape = ptr( w, w->shapes )[shapeId];
			if ( shape.proxyKey != kNull )
				bufferMove( w, shape.proxyKey );
			shapeId = shape.nextShapeId;
		}
	}
	else
	{
		destroyContactsBetweenBodies( w, j->edges[0].bodyId, j->edges[1].bodyId );
	}
}
bool b2Joint_GetCollideConnected( b2JointId jointId )
{
	Joint* j = jointFromId( jointId, nullptr, false );
	return j ? j->collideConnected : false;
}
void b2Joint_SetUserData( b2JointId jointId, void* userData )
{
	Joint* j = jointFromId( jointId, nullptr, true );
	if ( j )
		j->userData = (uint64_t)(uintptr_t)userData;
}
void* b2Joint_GetUserData( b2JointId jointId )
{
	Joint* j = jointFromId( jointId, nullptr, false );
	return j ? (void*)(uintptr_t)j->userData : nullptr;
}
void b2Joint_WakeBodies( b2JointId jointId ) // joint.c:1045-1059
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, true );
	if ( j == nullptr || hw->img->locked )
		return;
	World* w = hw->img;
	int idA = j->edges[0].bodyId, idB = j->edges[1].bodyId;
	wakeBody( w, ptr( w, w->bodies )[idA] );
	wakeBody( w, ptr( w, w->bodies )[idB] );
}
b2Vec2 b2Joint_GetConstraintForce( b2JointId jointId ) // joint.c:1061-1097 and the b2Get*JointForce of each type
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, false );
	if ( j == nullptr )
		return b2Vec2{ 0, 0 };
	World* w = hw->img;
	const JointSim& s = ptr( w, w->jointSims )[j->jointId];
	const BodySim* sims = ptr( w, w->sims );
	const float inv_h = w->inv_h;
	V2 f = { 0.0f, 0.0f };
	switch ( j->type )
	{
		case kDistanceJoint:
		{
			V2 pA = xfPoint( sims[s.bodyIdA].transform, s.localOriginAnchorA );
			V2 pB = xfPoint( sims[s.bodyIdB].transform, s.localOriginAnchorB );
			V2 axis = normalize( sub( pB, pA ) );
			const DistanceJointData& d = s.distance;
			float force = ( d.impulse + d.lowerImpulse - d.upperImpulse + d.motorImpulse ) * inv_h;
			f = mulSV( force, axis );
			break;
		}
		case kMotorJoint:
			f = mulSV( inv_h, s.motor.linearImpulse );
			break;
		case kMouseJoint:
			f = mulSV( inv_h, s.mouse.linearImpulse );
			break;
		case kPrismaticJoint:
		{
			V2 axisA = rotate( sims[s.bodyIdA].transform.q, s.prismatic.localAxisA );
			V2 perpA = leftPerp( axisA );
			float perpForce = inv_h * s.prismatic.impulse.x;
			float axialForce = inv_h * ( s.prismatic.motorImpulse + s.prismatic.lowerImpulse - s.prismatic.upperImpulse );
			f = add( mulSV( perpForce, perpA ), mulSV( axialForce, axisA ) );
			break;
		}
		case kRevoluteJoint:
			f = mulSV( inv_h, s.revolute.linearImpulse );
			break;
		case kWeldJoint:
			f = mulSV( inv_h, s.weld.linearImpulse );
			break;
		case kWheelJoint:
		{
			V2 axisA = s.wheel.axisA;
			V2 perpA = leftPerp( axisA );
			float perpForce = inv_h * s.wheel.perpImpulse;
			float axialForce = inv_h * ( s.wheel.springImpulse + s.wheel.lowerImpulse - s.wheel.upperImpulse );
			f = add( mulSV( perpForce, perpA ), mulSV( axialForce, axisA ) );
			break;
		}
		default:
			break;
	}
	return b2Vec2{ f.x, f.y };
}
float b2Joint_GetConstraintTorque( b2JointId jointId ) // joint.c:1099-1135
{
	HostWorld* hw = nullptr;
	Joint* j = jointFromId( jointId, &hw, false );
	if ( j == nullptr )
		return 0.0f;
	World* w = hw->img;
	const JointSim& s = ptr( w, w->jointSims )[j->jointId];
	const float inv_h = w->inv_h;
	switch ( j->type )
	{
		case kMotorJoint:
			return inv_h * s.motor.angularImpulse;
		case kMouseJoint:
			return inv_h * s.mouse.angularImpulse;
		case kPrismaticJoint:
			return inv_h * s.prismatic.impulse.y;
		case kRevoluteJoint:
			return inv_h * ( s.revolute.motorImpulse + s.revolute.lowerImpulse - s.revolute.upperImpulse );
		case kWeldJoint:
			return inv_h * s.weld.angularImpulse;
		case kWheelJoint:
			return inv_h * s.wheel.motorImpulse;
		default:
			return 0.0f;
	}
}

// ---- creation of the other joint types: joint.c:353-700 ------------------------------------------------------------
b2DistanceJointDef b2DefaultDistanceJointDef( void ) // joint.c:24-31
{
	b2DistanceJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.length = 1.0f;
	def.maxLength = kHuge;
	def.internalValue = kSecretCookie;
	return def;
}
b2MotorJointDef b2DefaultMotorJointDef( void ) // joint.c:33-41
{
	b2MotorJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.maxForce = 1.0f;
	def.maxTorque = 1.0f;
	def.correctionFactor = 0.3f;
	def.internalValue = kSecretCookie;
	return def;
}
b2MouseJointDef b2DefaultMouseJointDef( void ) // joint.c:43-51
{
	b2MouseJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.hertz = 4.0f;
	def.dampingRatio = 1.0f;
	def.maxForce = 1.0f;
	def.internalValue = kSecretCookie;
	return def;
}
b2FilterJointDef b2DefaultFilterJointDef( void ) // joint.c:53-58
{
	b2FilterJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.internalValue = kSecretCookie;
	return def;
}
b2PrismaticJointDef b2DefaultPrismaticJointDef( void ) // joint.c:60-66
{
	b2PrismaticJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.localAxisA = b2Vec2{ 1.0f, 0.0f };
	def.internalValue = kSecretCookie;
	return def;
}
b2WeldJointDef b2DefaultWeldJointDef( void ) // joint.c:76-81
{
	b2WeldJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.internalValue = kSecretCookie;
	return def;
}
b2WheelJointDef b2DefaultWheelJointDef( void ) // joint.c:83-92
{
	b2WheelJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.localAxisA.y = 1.0f;
	def.enableSpring = true;
	def.hertz = 1.0f;
	def.dampingRatio = 0.7f;
	def.internalValue = kSecretCookie;
	return def;
}
b2ExplosionDef b2DefaultExplosionDef( void ) // joint.c:94-99
{
	b2ExplosionDef def;
	memset( &def, 0, sizeof( def ) );
	def.maskBits = UINT64_MAX; // B2_DEFAULT_MASK_BITS
	return def;
}

// Common head of every b2Create*Joint: validates, grows the image, creates the base joint (joint.c:146-311)
struct JointCreation
{
	World* w;
	int jointId;
	JointSim* sim;
};
static bool beginJoint( b2WorldId worldId, int cookie, b2BodyId bodyIdA, b2BodyId bodyIdB, void* userData, float drawSize, int type,
						bool collideConnected, JointCreation* out )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr || cookie != kSecretCookie )
		return false;
	mutableImage( *hw );
	reserve( *hw, 0, 0, 0, 2 );
	World* w = hw->img;
	if ( w->locked )
		return false;
	int a = bodyIdA.index1 - 1, b = bodyIdB.index1 - 1;
	if ( a < 0 || b < 0 || a >= w->bodies.count || b >= w->bodies.count )
		return false;
	out->w = w;
	out->jointId = createJointBase( w, a, b, (uint64_t)(uintptr_t)userData, drawSize, type, collideConnected );
	out->sim = ptr( w, w->jointSims ) + out->jointId;
	out->sim->type = type;
	return true;
}
static b2JointId finishJoint( const JointCreation& c, bool collideConnected )
{
	if ( collideConnected == false )
		destroyContactsBetweenBodies( c.w, c.sim->bodyIdA, c.sim->bodyIdB );
	return b2JointId{ c.jointId + 1, c.w->worldId, ptr( c.w, c.w->joints )[c.jointId].generation };
}
b2JointId b2CreateDistanceJoint( b2WorldId worldId, const b2DistanceJointDef* def ) // joint.c:353-404
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kDistanceJoint, def->collideConnected, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	c.sim->localOriginAnchorA = V2{ def->localAnchorA.x, def->localAnchorA.y };
	c.sim->localOriginAnchorB = V2{ def->localAnchorB.x, def->localAnchorB.y };
	DistanceJointData& j = c.sim->distance;
	memset( &j, 0, sizeof( j ) );
	j.length = maxf( def->length, kLinearSlop );
	j.hertz = def->hertz;
	j.dampingRatio = def->dampingRatio;
	j.minLength = maxf( def->minLength, kLinearSlop );
	j.maxLength = maxf( def->minLength, def->maxLength );
	j.maxMotorForce = def->maxMotorForce;
	j.motorSpeed = def->motorSpeed;
	j.enableSpring = def->enableSpring;
	j.enableLimit = def->enableLimit;
	j.enableMotor = def->enableMotor;
	return finishJoint( c, def->collideConnected );
}
b2JointId b2CreateMotorJoint( b2WorldId worldId, const b2MotorJointDef* def ) // joint.c:406-442
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kMotorJoint, def->collideConnected, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	c.sim->localOriginAnchorA = V2{ 0.0f, 0.0f };
	c.sim->localOriginAnchorB = V2{ 0.0f, 0.0f };
	MotorJointData& j = c.sim->motor;
	memset( &j, 0, sizeof( j ) );
	j.linearOffset = V2{ def->linearOffset.x, def->linearOffset.y };
	j.angularOffset = def->angularOffset;
	j.maxForce = def->maxForce;
	j.maxTorque = def->maxTorque;
	j.correctionFactor = clampf( def->correctionFactor, 0.0f, 1.0f );
	return finishJoint( c, def->collideConnected );
}
b2JointId b2CreateMouseJoint( b2WorldId worldId, const b2MouseJointDef* def ) // joint.c:444-478
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kMouseJoint, def->collideConnected, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	const BodySim* sims = ptr( c.w, c.w->sims );
	Xf xfA = sims[c.sim->bodyIdA].transform, xfB = sims[c.sim->bodyIdB].transform;
	V2 target = { def->target.x, def->target.y };
	c.sim->localOriginAnchorA = invRotate( xfA.q, sub( target, xfA.p ) );
	c.sim->localOriginAnchorB = invRotate( xfB.q, sub( target, xfB.p ) );
	MouseJointData& j = c.sim->mouse;
	memset( &j, 0, sizeof( j ) );
	j.targetA = target;
	j.hertz = def->hertz;
	j.dampingRatio = def->dampingRatio;
	j.maxForce = def->maxForce;
	return finishJoint( c, true ); // the reference never destroys contacts for a mouse joint (joint.c:444-478)
}
b2JointId b2CreateFilterJoint( b2WorldId worldId, const b2FilterJointDef* def ) // joint.c:480-505
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kFilterJoint, false, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	c.sim->localOriginAnchorA = V2{ 0.0f, 0.0f };
	c.sim->localOriginAnchorB = V2{ 0.0f, 0.0f };
	return finishJoint( c, true ); // joint.c:480-505: no contact destruction here either
}
b2JointId b2CreatePrismaticJoint( b2WorldId worldId, const b2PrismaticJointDef* def ) // joint.c:559-607
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kPrismaticJoint, def->collideConnected, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	c.sim->localOriginAnchorA = V2{ def->localAnchorA.x, def->localAnchorA.y };
	c.sim->localOriginAnchorB = V2{ def->localAnchorB.x, def->localAnchorB.y };
	PrismaticJointData& j = c.sim->prismatic;
	memset( &j, 0, sizeof( j ) );
	j.localAxisA = normalize( V2{ def->localAxisA.x, def->localAxisA.y } );
	j.referenceAngle = def->referenceAngle;
	j.targetTranslation = def->targetTranslation;
	j.hertz = def->hertz;
	j.dampingRatio = def->dampingRatio;
	j.lowerTranslation = def->lowerTranslation;
	j.upperTranslation = def->upperTranslation;
	j.maxMotorForce = def->maxMotorForce;
	j.motorSpeed = def->motorSpeed;
	j.enableSpring = def->enableSpring;
	j.enableLimit = def->enableLimit;
	j.enableMotor = def->enableMotor;
	return finishJoint( c, def->collideConnected );
}
b2JointId b2CreateWeldJoint( b2WorldId worldId, const b2WeldJointDef* def ) // joint.c:609-649
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kWeldJoint, def->collideConnected, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	c.sim->localOriginAnchorA = V2{ def->localAnchorA.x, def->localAnchorA.y };
	c.sim->localOriginAnchorB = V2{ def->localAnchorB.x, def->localAnchorB.y };
	WeldJointData& j = c.sim->weld;
	memset( &j, 0, sizeof( j ) );
	j.referenceAngle = def->referenceAngle;
	j.linearHertz = def->linearHertz;
	j.linearDampingRatio = def->linearDampingRatio;
	j.angularHertz = def->angularHertz;
	j.angularDampingRatio = def->angularDampingRatio;
	return finishJoint( c, def->collideConnected );
}
b2JointId b2CreateWheelJoint( b2WorldId worldId, const b2WheelJointDef* def ) // joint.c:651-700
{
	JointCreation c;
	if ( beginJoint( worldId, def->internalValue, def->bodyIdA, def->bodyIdB, def->userData, 1.0f, kWheelJoint, def->collideConnected, &c ) == false )
		return b2JointId{ 0, 0, 0 };
	c.sim->localOriginAnchorA = V2{ def->localAnchorA.x, def->localAnchorA.y };
	c.sim->localOriginAnchorB = V2{ def->localAnchorB.x, def->localAnchorB.y };
	WheelJointData& j = c.sim->wheel;
	memset( &j, 0, sizeof( j ) );
	j.localAxisA = normalize( V2{ def->localAxisA.x, def->localAxisA.y } );
	j.lowerTranslation = def->lowerTranslation;
	j.upperTranslation = def->upperTranslation;
	j.maxMotorTorque = def->maxMotorTorque;
	j.motorSpeed = def->motorSpeed;
	j.hertz = def->hertz;
	j.dampingRatio = def->dampingRatio;
	j.enableSpring = def->enableSpring;
	j.enableLimit = def->enableLimit;
	j.enableMotor = def->enableMotor;
	return finishJoint( c, def->collideConnected );
}

// ---------------------------------------------------------------------------------------------------------------- chains
// shape.c:339-577, :1476-1576. A chain is a run of one-sided chain-segment shapes with ghost vertices; the record that
// ties them together is only needed by these API calls and stays on the host (HostWorld::chains).
b2ChainDef b2DefaultChainDef( void ) // types.c:76-88
{
	static b2SurfaceMaterial defaultMaterial = { 0.6f, 0.0f, 0.0f, 0.0f, 0, 0 };
	b2ChainDef def;
	memset( &def, 0, sizeof( def ) );
	def.materials = &defaultMaterial;
	def.materialCount = 1;
	def.filter = b2DefaultFilter();
	def.internalValue = kSecretCookie;
	return def;
}
static HostWorld::Chain* chainFromId( b2ChainId id, HostWorld** outWorld )
{
	HostWorld* hw = worldFromIndex0( id.world0 );
	if ( hw == nullptr )
		return nullptr;
	int index = id.index1 - 1;
	if ( index < 0 || index >= (int)hw->chains.size() )
		return nullptr;
	HostWorld::Chain& c = hw->chains[index];
	if ( c.id != index || c.generation != id.generation )
		return nullptr;
	if ( outWorld )
		*outWorld = hw;
	return &c;
}
b2ChainId b2CreateChain( b2BodyId bodyId, const b2ChainDef* def ) // shape.c:339-477
{
	if ( def->internalValue != kSecretCookie || def->count < 4 || def->materialCount < 1 )
		return b2ChainId{ 0, 0, 0 };
	HostWorld* hw = nullptr;
	Body* b = writableBody( bodyId, &hw );
	if ( b == nullptr )
		return b2ChainId{ 0, 0, 0 };
	const int bodyIndex = b->id;
	const int n = def->count;
	reserve( *hw, 0, n + 2, 0, 0 );
	World* w = hw->img;
	int chainId = allocId( w, w->chainIds );
	if ( chainId == (int)hw->chains.size() )
		hw->chains.push_back( HostWorld::Chain() );
	HostWorld::Chain& chain = hw->chains[chainId];
	Body& body = ptr( w, w->bodies )[bodyIndex];
	chain.id = chainId;
	chain.bodyId = bodyIndex;
	chain.nextChainId = body.headChainId;
	chain.generation += 1;
	chain.materials.assign( def->materials, def->materials + def->materialCount );
	chain.shapeIndices.clear();
	body.headChainId = chainId;

	ShapeParams p;
	{
		b2ShapeDef sd = b2DefaultShapeDef();
		p.userData = (uint64_t)(uintptr_t)def->userData;
		p.density = sd.density;
		p.filter = Filter{ def->filter.categoryBits, def->filter.maskBits, def->filter.groupIndex };
		p.isSensor = false;
		p.enableSensorEvents = def->enableSensorEvents;
		p.enableContactEvents = false;
		p.enableHitEvents = false;
		p.enablePreSolveEvents = sd.enablePreSolveEvents;
		p.invokeContactCreation = sd.invokeContactCreation;
		p.updateBodyMass = sd.updateBodyMass;
	}
	const int materialCount = def->materialCount;
	const b2Vec2* pts = def->points;
	auto segment = [&]( int g1, int a, int c, int g2, int materialIndex ) {
		const b2SurfaceMaterial& mat = def->materials[materialCount == 1 ? 0 : materialIndex];
		p.friction = mat.friction;
		p.restitution = mat.restitution;
		p.rollingResistance = mat.rollingResistance;
		p.tangentSpeed = mat.tangentSpeed;
		p.userMaterialId = mat.userMaterialId;
		p.customColor = mat.customColor;
		ChainSegment cs;
		cs.ghost1 = V2{ pts[g1].x, pts[g1].y };
		cs.segment.p1 = V2{ pts[a].x, pts[a].y };
		cs.segment.p2 = V2{ pts[c].x, pts[c].y };
		cs.ghost2 = V2{ pts[g2].x, pts[g2].y };
		cs.chainId = chainId;
		chain.shapeIndices.push_back( createShape( w, bodyIndex, p, &cs, kChainSegment ) );
	};
	if ( def->isLoop )
	{
		int prev = n - 1;
		for ( int i = 0; i < n - 2; ++i )
		{
			segment( prev, i, i + 1, i + 2, i );
			prev = i;
		}
		segment( n - 3, n - 2, n - 1, 0, n - 2 );
		segment( n - 2, n - 1, 0, 1, n - 1 );
	}
	else
	{
		for ( int i = 0; i < n - 3; ++i )
			segment( i, i + 1, i + 2, i + 3, i + 1 );
	}
	return b2ChainId{ chainId + 1, w->worldId, chain.generation };
}
void b2DestroyChain( b2ChainId chainId ) // shape.c:488-537
{
	HostWorld* hw = nullptr;
	HostWorld::Chain* chain = chainFromId( chainId, &hw );
	if ( chain == nullptr )
		return;
	World* w = mutableImage( *hw );
	if ( w->locked )
		return;
	Body& body = ptr( w, w->bodies )[chain->bodyId];
	// unlink from the body's singly linked chain list
	int* link = &body.headChainId;
	bool found = false;
	while ( *link != kNull )
	{
		if ( *link == chain->id )
		{
			*link = chain->nextChainId;
			found = true;
			break;
		}
		link = &hw->chains[*link].nextChainId;
	}
	if ( found == false )
		return;
	for ( int shapeId : chain->shapeIndices )
		destroyShape( w, shapeId, false );
	chain->shapeIndices.clear();
	chain->materials.clear();
	freeId( w, w->chainIds, chain->id );
	chain->id = kNull;
}
bool b2Chain_IsValid( b2ChainId id )
{
	return chainFromId( id, nullptr ) != nullptr;
}
b2WorldId b2Chain_GetWorld( b2ChainId chainId )
{
	HostWorld* hw = worldFromIndex0( chainId.world0 );
	return hw ? b2WorldId{ (uint16_t)( chainId.world0 + 1 ), hw->generation } : b2WorldId{ 0, 0 };
}
int b2Chain_GetSegmentCount( b2ChainId chainId )
{
	HostWorld::Chain* chain = chainFromId( chainId, nullptr );
	return chain ? (int)chain->shapeIndices.size() : 0;
}
int b2Chain_GetSegments( b2ChainId chainId, b2ShapeId* segmentArray, int capacity ) // shape.c:557-576
{
	HostWorld* hw = nullptr;
	HostWorld::Chain* chain = chainFromId( chainId, &hw );
	if ( chain == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int count = mini( (int)chain->shapeIndices.size(), capacity );
	for ( int i = 0; i < count; ++i )
	{
		int shapeId = chain->shapeIndices[i];
		segmentArray[i] = b2ShapeId{ shapeId + 1, chainId.world0, ptr( w, w->shapes )[shapeId].generation };
	}
	return count;
}
void b2Chain_SetFriction( b2ChainId chainId, float friction ) // shape.c:1476-1502
{
	HostWorld* hw = nullptr;
	HostWorld::Chain* chain = chainFromId( chainId, &hw );
	if ( chain == nullptr )
		return;
	World* w = mutableImage( *hw );
	if ( w->locked )
		return;
	for ( b2SurfaceMaterial& m : chain->materials )
		m.friction = friction;
	for ( int shapeId : chain->shapeIndices )
		ptr( w, w->shapes )[shapeId].friction = friction;
}
float b2Chain_GetFriction( b2ChainId chainId )
{
	HostWorld::Chain* chain = chainFromId( chainId, nullptr );
	return chain && chain->materials.empty() == false ? chain->materials[0].friction : 0.0f;
}
void b2Chain_SetRestitution( b2ChainId chainId, float restitution ) // shape.c:1511-1537
{
	HostWorld* hw = nullptr;
	HostWorld::Chain* chain = chainFromId( chainId, &hw );
	if ( chain == nullptr )
		return;
	World* w = mutableImage( *hw );
	if ( w->locked )
		return;
	for ( b2SurfaceMaterial& m : chain->materials )
		m.restitution = restitution;
	for ( int shapeId : chain->shapeIndices )
		ptr( w, w->shapes )[shapeId].restitution = restitution;
}
float b2Chain_GetRestitution( b2ChainId chainId )
{
	HostWorld::Chain* chain = chainFromId( chainId, nullptr );
	return chain && chain->materials.empty() == false ? chain->materials[0].restitution : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------- queries
// world.c:2040-2310. The world's state lives on the device; a query first makes the host image current (one download
// after a step, none between consecutive queries), then walks the trees on the host because every candidate is handed
// to a synchronous host callback whose answer clips or stops the walk.
static b2ShapeId publicShapeId( World* w, int shapeId )
{
	return b2ShapeId{ shapeId + 1, w->worldId, ptr( w, w->shapes )[shapeId].generation };
}
b2QueryFilter b2DefaultQueryFilter( void ) // types.c
{
	return b2QueryFilter{ 1ull, UINT64_MAX };
}
b2TreeStats b2World_OverlapAABB( b2WorldId worldId, b2AABB aabb, b2QueryFilter filter, b2OverlapResultFcn* fcn, void* context )
{
	b2TreeStats total = { 0, 0 };
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return total;
	World* w = hostImage( *hw );
	if ( w->locked )
		return total;
	Box box = { { aabb.lowerBound.x, aabb.lowerBound.y }, { aabb.upperBound.x, aabb.upperBound.y } };
	for ( int i = 0; i < 3; ++i )
	{
		TreeStats st = treeQueryStats( w, w->trees[i], box, filter.maskBits, [&]( int, uint64_t userData ) -> bool {
			int shapeId = (int)userData;
			const Shape& shape = ptr( w, w->shapes )[shapeId];
			if ( shouldQueryCollide( shape.filter, filter.categoryBits, filter.maskBits ) == false )
				return true;
			return fcn( publicShapeId( w, shapeId ), context );
		} );
		total.nodeVisits += st.nodeVisits;
		total.leafVisits += st.leafVisits;
	}
	return total;
}
static b2TreeStats castRayCommon( HostWorld* hw, b2Vec2 origin, b2Vec2 translation, b2QueryFilter filter, b2CastResultFcn* fcn, void* context )
{
	b2TreeStats total = { 0, 0 };
	World* w = hostImage( *hw );
	if ( w->locked )
		return total;
	RayInput input = { { origin.x, origin.y }, { translation.x, translation.y }, 1.0f };
	float worldFraction = 1.0f;
	for ( int i = 0; i < 3; ++i )
	{
		TreeStats st = treeRayCast( w, w->trees[i], input, filter.maskBits, [&]( const RayInput& sub, int, uint64_t userData ) -> float {
			int shapeId = (int)userData;
			const Shape& shape = ptr( w, w->shapes )[shapeId];
			if ( shouldQueryCollide( shape.filter, filter.categoryBits, filter.maskBits ) == false )
				return sub.maxFraction;
			Xf transform = ptr( w, w->sims )[shape.bodyId].transform;
			CastOutput out = rayCastShape( sub, shape, transform );
			if ( out.hit )
			{
				float fraction = fcn( publicShapeId( w, shapeId ), b2Vec2{ out.point.x, out.point.y }, b2Vec2{ out.normal.x, out.normal.y },
									  out.fraction, context );
				if ( 0.0f <= fraction && fraction <= 1.0f )
					worldFraction = fraction;
				return fraction;
			}
			return sub.maxFraction;
		} );
		total.nodeVisits += st.nodeVisits;
		total.leafVisits += st.leafVisits;
		if ( worldFraction == 0.0f )
			return total;
		input.maxFraction = worldFraction;
	}
	return total;
}
b2TreeStats b2World_CastRay( b2WorldId worldId, b2Vec2 origin, b2Vec2 translation, b2QueryFilter filter, b2CastResultFcn* fcn, void* context )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return b2TreeStats{ 0, 0 };
	return castRayCommon( hw, origin, translation, filter, fcn, context );
}
static float closestHit( b2ShapeId shapeId, b2Vec2 point, b2Vec2 normal, float fraction, void* context ) // world.c:2260-2275
{
	if ( fraction == 0.0f )
		return -1.0f;
	b2RayResult* r = (b2RayResult*)context;
	r->shapeId = shapeId;
	r->point = point;
	r->normal = normal;
	r->fraction = fraction;
	r->hit = true;
	return fraction;
}
b2RayResult b2World_CastRayClosest( b2WorldId worldId, b2Vec2 origin, b2Vec2 translation, b2QueryFilter filter ) // world.c:2277-2308
{
	b2RayResult result;
	memset( &result, 0, sizeof( result ) );
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return result;
	b2TreeStats st = castRayCommon( hw, origin, translation, filter, closestHit, &result );
	result.nodeVisits = st.nodeVisits;
	result.leafVisits = st.leafVisits;
	return result;
}
void b2World_Explode( b2WorldId worldId, const b2ExplosionDef* def ) // world.c:2640-2745
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	World* w = mutableImage( *hw );
	if ( w->locked )
		return;
	V2 position = { def->position.x, def->position.y };
	float radius = def->radius, falloff = def->falloff, impulsePerLength = def->impulsePerLength;
	Box box = { { position.x - ( radius + falloff ), position.y - ( radius + falloff ) },
				{ position.x + ( radius + falloff ), position.y + ( radius + falloff ) } };
	treeQueryStats( w, w->trees[kDynamicBody], box, def->maskBits, [&]( int, uint64_t userData ) -> bool {
		const Shape& shape = ptr( w, w->shapes )[(int)userData];
		Body& body = ptr( w, w->bodies )[shape.bodyId];
		BodySim& sim = ptr( w, w->sims )[shape.bodyId];
		Xf transform = sim.transform;
		Xf identity = { { 0.0f, 0.0f }, { 1.0f, 0.0f } };
		SimplexCache cache;
		memset( &cache, 0, sizeof( cache ) );
		DistanceOutput out = shapeDistance( makeShapeProxy( shape ), makeProxy( &position, 1, 0.0f ), transform, identity, true, &cache );
		if ( out.distance > radius + falloff )
			return true;
		wakeBody( w, body );
		if ( body.setIndex != kAwakeSet )
			return true;
		V2 closestPoint = out.pointA;
		if ( out.distance == 0.0f )
			closestPoint = xfPoint( transform, shapeCentroid( shape ) );
		V2 direction = sub( closestPoint, position );
		if ( lengthSq( direction ) > 100.0f * FLT_EPSILON * FLT_EPSILON )
			direction = normalize( direction );
		else
			direction = V2{ 1.0f, 0.0f };
		V2 localLine = invRotate( transform.q, leftPerp( direction ) );
		float perimeter = shapeProjectedPerimeter( shape, localLine );
		float scale = 1.0f;
		if ( out.distance > radius && falloff > 0.0f )
			scale = clampf( ( radius + falloff - out.distance ) / falloff, 0.0f, 1.0f );
		float magnitude = impulsePerLength * perimeter * scale;
		V2 impulse = mulSV( magnitude, direction );
		BodyState& state = ptr( w, w->states )[body.localIndex];
		state.v = mulAdd( state.v, sim.invMass, impulse );
		state.w += sim.invInertia * cross( sub( closestPoint, sim.center ), impulse );
		return true;
	} );
}
// world.c:1710-1740. A registered callback switches b2World_Step to the callback-mediated schedule
// (stepWithHostCallbacks, f2d_capi.inl); passing NULL switches back to the single-launch step.
void b2World_SetCustomFilterCallback( b2WorldId worldId, b2CustomFilterFcn* fcn, void* context )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	hw->customFilterFcn = fcn;
	hw->customFilterContext = context;
	World* w = mutableImage( *hw );
	w->hostCallbacks = (uint8_t)( fcn != nullptr ? ( w->hostCallbacks | kHostCustomFilter ) : ( w->hostCallbacks & ~kHostCustomFilter ) );
}
void b2World_SetPreSolveCallback( b2WorldId worldId, b2PreSolveFcn* fcn, void* context )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	hw->preSolveFcn = fcn;
	hw->preSolveContext = context;
	World* w = mutableImage( *hw );
	w->hostCallbacks = (uint8_t)( fcn != nullptr ? ( w->hostCallbacks | kHostPreSolve ) : ( w->hostCallbacks & ~kHostPreSolve ) );
}

#include "f2d_capi_draw.inl"
#include "f2d_capi_joints.inl"
