// forge2d_b200 — C ABI of the B200-native world step.
//
// Drop-in boundary: forge2d's Dart layer reaches its physics through ffigen-generated `@Native` bindings to the
// Box2D v3.1.1 C API (packages/forge2d/lib/src/ffi/box2d.g.dart, called from
// packages/forge2d/lib/src/backend/raw_box2d_ffi.dart). This library exports the SAME symbol names with the same
// by-value struct layouts for the World.step path and the calls needed to build and observe a world, so the
// native-assets hook (packages/forge2d/hook/build.dart:62-81) can register this .so under the same asset id.
// Each declaration cites the reference declaration it replaces (B2 = packages/forge2d/third_party/box2d).
//
// Plain C types only: no torch / CUDA types cross this boundary. Everything behind b2World_Step runs as CUDA
// kernels on the current device; there is no CPU fallback (b2World_Step reports an error if no GPU is present).
#pragma once
#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F2D_API __attribute__( ( visibility( "default" ) ) )

// ---- value types: B2/include/box2d/math_functions.h:14-62, id.h:37-73 --------------------------------------
typedef struct b2Vec2 { float x, y; } b2Vec2;
typedef struct b2Rot { float c, s; } b2Rot;
typedef struct b2Transform { b2Vec2 p; b2Rot q; } b2Transform;
typedef struct b2AABB { b2Vec2 lowerBound, upperBound; } b2AABB;
typedef struct b2WorldId { uint16_t index1, generation; } b2WorldId;
typedef struct b2BodyId { int32_t index1; uint16_t world0, generation; } b2BodyId;
typedef struct b2ShapeId { int32_t index1; uint16_t world0, generation; } b2ShapeId;
typedef struct b2JointId { int32_t index1; uint16_t world0, generation; } b2JointId;
typedef struct b2ChainId { int32_t index1; uint16_t world0, generation; } b2ChainId;

// ---- geometry: B2/include/box2d/collision.h:24-186 ------------------------------------------------------------
#define B2_MAX_POLYGON_VERTICES 8
typedef struct b2Circle { b2Vec2 center; float radius; } b2Circle;
typedef struct b2Capsule { b2Vec2 center1, center2; float radius; } b2Capsule;
typedef struct b2Polygon
{
	b2Vec2 vertices[B2_MAX_POLYGON_VERTICES];
	b2Vec2 normals[B2_MAX_POLYGON_VERTICES];
	b2Vec2 centroid;
	float radius;
	int count;
} b2Polygon;
typedef struct b2Segment { b2Vec2 point1, point2; } b2Segment;
typedef struct b2Hull { b2Vec2 points[B2_MAX_POLYGON_VERTICES]; int count; } b2Hull;

// collision.h:498-552
typedef struct b2ManifoldPoint
{
	b2Vec2 point, anchorA, anchorB;
	float separation, normalImpulse, tangentImpulse, totalNormalImpulse, normalVelocity;
	uint16_t id;
	bool persisted;
} b2ManifoldPoint;
typedef struct b2Manifold
{
	b2Vec2 normal;
	float rollingImpulse;
	b2ManifoldPoint points[2];
	int pointCount;
} b2Manifold;

// ---- definitions: B2/include/box2d/types.h:66-460 -------------------------------------------------------------
typedef void b2TaskCallback( int startIndex, int endIndex, uint32_t workerIndex, void* taskContext );
typedef void* b2EnqueueTaskCallback( b2TaskCallback* task, int itemCount, int minRange, void* taskContext, void* userContext );
typedef void b2FinishTaskCallback( void* userTask, void* userContext );
typedef float b2FrictionCallback( float frictionA, int userMaterialIdA, float frictionB, int userMaterialIdB );
typedef float b2RestitutionCallback( float restitutionA, int userMaterialIdA, float restitutionB, int userMaterialIdB );

typedef struct b2WorldDef // types.h:66-143
{
	b2Vec2 gravity;
	float restitutionThreshold, hitEventThreshold, contactHertz, contactDampingRatio, maxContactPushSpeed, maximumLinearSpeed;
	b2FrictionCallback* frictionCallback;
	b2RestitutionCallback* restitutionCallback;
	bool enableSleep, enableContinuous;
	int workerCount;
	b2EnqueueTaskCallback* enqueueTask;
	b2FinishTaskCallback* finishTask;
	void* userTaskContext;
	void* userData;
	int internalValue;
} b2WorldDef;

typedef enum b2BodyType { b2_staticBody = 0, b2_kinematicBody = 1, b2_dynamicBody = 2, b2_bodyTypeCount } b2BodyType;

typedef struct b2BodyDef // types.h:166-232
{
	b2BodyType type;
	b2Vec2 position;
	b2Rot rotation;
	b2Vec2 linearVelocity;
	float angularVelocity, linearDamping, angularDamping, gravityScale, sleepThreshold;
	const char* name;
	void* userData;
	bool enableSleep, isAwake, fixedRotation, isBullet, isEnabled, allowFastRotation;
	int internalValue;
} b2BodyDef;

typedef struct b2Filter { uint64_t categoryBits, maskBits; int groupIndex; } b2Filter; // types.h:240-272
typedef struct b2SurfaceMaterial															 // types.h:308-330
{
	float friction, restitution, rollingResistance, tangentSpeed;
	int userMaterialId;
	uint32_t customColor;
} b2SurfaceMaterial;
typedef struct b2ShapeDef // types.h:340-385
{
	void* userData;
	b2SurfaceMaterial material;
	float density;
	b2Filter filter;
	bool isSensor, enableSensorEvents, enableContactEvents, enableHitEvents, enablePreSolveEvents, invokeContactCreation,
		updateBodyMass;
	int internalValue;
} b2ShapeDef;

typedef struct b2RevoluteJointDef // types.h:760-826
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 localAnchorA, localAnchorB;
	float referenceAngle, targetAngle;
	bool enableSpring;
	float hertz, dampingRatio;
	bool enableLimit;
	float lowerAngle, upperAngle;
	bool enableMotor;
	float maxMotorTorque, motorSpeed, drawSize;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2RevoluteJointDef;

// collision.h:91-101, types.h:514-524, collision.h:56-70 + shape types (collision.h / types.h b2ShapeType)
typedef struct b2MassData { float mass; b2Vec2 center; float rotationalInertia; } b2MassData;
typedef enum b2ShapeType { b2_circleShape, b2_capsuleShape, b2_segmentShape, b2_polygonShape, b2_chainSegmentShape, b2_shapeTypeCount } b2ShapeType;
typedef enum b2JointType
{
	b2_distanceJoint, b2_filterJoint, b2_motorJoint, b2_mouseJoint, b2_prismaticJoint, b2_revoluteJoint, b2_weldJoint, b2_wheelJoint
} b2JointType;
typedef struct b2ChainSegment { b2Vec2 ghost1; b2Segment segment; b2Vec2 ghost2; int chainId; } b2ChainSegment;

// ---- joint definitions: types.h:526-900 ------------------------------------------------------------------------
typedef struct b2DistanceJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 localAnchorA, localAnchorB;
	float length;
	bool enableSpring;
	float hertz, dampingRatio;
	bool enableLimit;
	float minLength, maxLength;
	bool enableMotor;
	float maxMotorForce, motorSpeed;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2DistanceJointDef;
typedef struct b2MotorJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 linearOffset;
	float angularOffset, maxForce, maxTorque, correctionFactor;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2MotorJointDef;
typedef struct b2MouseJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 target;
	float hertz, dampingRatio, maxForce;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2MouseJointDef;
typedef struct b2FilterJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	void* userData;
	int internalValue;
} b2FilterJointDef;
typedef struct b2PrismaticJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 localAnchorA, localAnchorB, localAxisA;
	float referenceAngle, targetTranslation;
	bool enableSpring;
	float hertz, dampingRatio;
	bool enableLimit;
	float lowerTranslation, upperTranslation;
	bool enableMotor;
	float maxMotorForce, motorSpeed;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2PrismaticJointDef;
typedef struct b2WeldJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 localAnchorA, localAnchorB;
	float referenceAngle, linearHertz, angularHertz, linearDampingRatio, angularDampingRatio;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2WeldJointDef;
typedef struct b2WheelJointDef
{
	b2BodyId bodyIdA, bodyIdB;
	b2Vec2 localAnchorA, localAnchorB, localAxisA;
	bool enableSpring;
	float hertz, dampingRatio;
	bool enableLimit;
	float lowerTranslation, upperTranslation;
	bool enableMotor;
	float maxMotorTorque, motorSpeed;
	bool collideConnected;
	void* userData;
	int internalValue;
} b2WheelJointDef;
typedef struct b2ChainDef // types.h:429-458
{
	void* userData;
	const b2Vec2* points;
	int count;
	const b2SurfaceMaterial* materials;
	int materialCount;
	b2Filter filter;
	bool isLoop, enableSensorEvents;
	int internalValue;
} b2ChainDef;
typedef struct b2ExplosionDef { uint64_t maskBits; b2Vec2 position; float radius, falloff, impulsePerLength; } b2ExplosionDef;

// ---- queries: types.h:291-305 (b2QueryFilter), :67-76 (b2RayResult), collision.h:658-665 (b2TreeStats), callbacks types.h:1180-1220
typedef struct b2QueryFilter { uint64_t categoryBits, maskBits; } b2QueryFilter;
typedef struct b2RayResult { b2ShapeId shapeId; b2Vec2 point, normal; float fraction; int nodeVisits, leafVisits; bool hit; } b2RayResult;
typedef struct b2TreeStats { int nodeVisits, leafVisits; } b2TreeStats;
typedef bool b2OverlapResultFcn( b2ShapeId shapeId, void* context );
typedef float b2CastResultFcn( b2ShapeId shapeId, b2Vec2 point, b2Vec2 normal, float fraction, void* context );
typedef bool b2CustomFilterFcn( b2ShapeId shapeIdA, b2ShapeId shapeIdB, void* context );
typedef bool b2PreSolveFcn( b2ShapeId shapeIdA, b2ShapeId shapeIdB, b2Manifold* manifold, void* context );

// ---- debug draw: types.h:1228-1460 (b2HexColor is an int-sized enum; only the colours b2World_Draw emits are named here)
typedef enum b2HexColor
{
	b2_colorBlack = 0x000000, b2_colorBlue = 0x0000FF, b2_colorBlueViolet = 0x8A2BE2, b2_colorChocolate = 0xD2691E, b2_colorCoral = 0xFF7F50,
	b2_colorCyan = 0x00FFFF, b2_colorDarkSeaGreen = 0x8FBC8F, b2_colorDimGray = 0x696969, b2_colorGainsboro = 0xDCDCDC, b2_colorGold = 0xFFD700,
	b2_colorGoldenRod = 0xDAA520, b2_colorGray = 0x808080, b2_colorGreen = 0x008000, b2_colorLightGray = 0xD3D3D3, b2_colorLightGreen = 0x90EE90,
	b2_colorMagenta = 0xFF00FF, b2_colorOrange = 0xFFA500, b2_colorOrangeRed = 0xFF4500, b2_colorPaleGreen = 0x98FB98, b2_colorPink = 0xFFC0CB,
	b2_colorRed = 0xFF0000, b2_colorRoyalBlue = 0x4169E1, b2_colorSalmon = 0xFA8072, b2_colorSlateGray = 0x708090, b2_colorTurquoise = 0x40E0D0,
	b2_colorViolet = 0xEE82EE, b2_colorWheat = 0xF5DEB3, b2_colorWhite = 0xFFFFFF, b2_colorYellow = 0xFFFF00
} b2HexColor;
typedef struct b2DebugDraw // types.h:1383-1460: nine callbacks, the bounds, fourteen option flags, the user context
{
	void ( *DrawPolygonFcn )( const b2Vec2* vertices, int vertexCount, b2HexColor color, void* context );
	void ( *DrawSolidPolygonFcn )( b2Transform transform, const b2Vec2* vertices, int vertexCount, float radius, b2HexColor color, void* context );
	void ( *DrawCircleFcn )( b2Vec2 center, float radius, b2HexColor color, void* context );
	void ( *DrawSolidCircleFcn )( b2Transform transform, float radius, b2HexColor color, void* context );
	void ( *DrawSolidCapsuleFcn )( b2Vec2 p1, b2Vec2 p2, float radius, b2HexColor color, void* context );
	void ( *DrawSegmentFcn )( b2Vec2 p1, b2Vec2 p2, b2HexColor color, void* context );
	void ( *DrawTransformFcn )( b2Transform transform, void* context );
	void ( *DrawPointFcn )( b2Vec2 p, float size, b2HexColor color, void* context );
	void ( *DrawStringFcn )( b2Vec2 p, const char* s, b2HexColor color, void* context );
	b2AABB drawingBounds;
	bool useDrawingBounds, drawShapes, drawJoints, drawJointExtras, drawBounds, drawMass, drawBodyNames, drawContacts, drawGraphColors,
		drawContactNormals, drawContactImpulses, drawContactFeatures, drawFrictionImpulses, drawIslands;
	void* context;
} b2DebugDraw;

typedef struct b2Counters // types.h:492-505
{
	int bodyCount, shapeCount, contactCount, jointCount, islandCount, stackUsed, staticTreeHeight, treeHeight, byteCount, taskCount;
	int colorCounts[12];
} b2Counters;

// ---- events: B2/include/box2d/types.h:1000-1160 ---------------------------------------------------------------
typedef struct b2SensorBeginTouchEvent { b2ShapeId sensorShapeId, visitorShapeId; } b2SensorBeginTouchEvent;
typedef struct b2SensorEndTouchEvent { b2ShapeId sensorShapeId, visitorShapeId; } b2SensorEndTouchEvent;
typedef struct b2SensorEvents
{
	b2SensorBeginTouchEvent* beginEvents;
	b2SensorEndTouchEvent* endEvents;
	int beginCount, endCount;
} b2SensorEvents;
typedef struct b2ContactBeginTouchEvent { b2ShapeId shapeIdA, shapeIdB; b2Manifold manifold; } b2ContactBeginTouchEvent;
typedef struct b2ContactEndTouchEvent { b2ShapeId shapeIdA, shapeIdB; } b2ContactEndTouchEvent;
typedef struct b2ContactHitEvent { b2ShapeId shapeIdA, shapeIdB; b2Vec2 point, normal; float approachSpeed; } b2ContactHitEvent;
typedef struct b2ContactEvents
{
	b2ContactBeginTouchEvent* beginEvents;
	b2ContactEndTouchEvent* endEvents;
	b2ContactHitEvent* hitEvents;
	int beginCount, endCount, hitCount;
} b2ContactEvents;
typedef struct b2BodyMoveEvent { b2Transform transform; b2BodyId bodyId; void* userData; bool fellAsleep; } b2BodyMoveEvent;
typedef struct b2BodyEvents { b2BodyMoveEvent* moveEvents; int moveCount; } b2BodyEvents;

// ---- defaults: B2/src/types.c:9-88, joint.c:68-74 -------------------------------------------------------------
F2D_API b2WorldDef b2DefaultWorldDef( void );			  // box2d.g.dart b2DefaultWorldDef / types.h:146
F2D_API b2BodyDef b2DefaultBodyDef( void );				  // types.h:236
F2D_API b2Filter b2DefaultFilter( void );				  // types.h:276
F2D_API b2ShapeDef b2DefaultShapeDef( void );			  // types.h:389
F2D_API b2SurfaceMaterial b2DefaultSurfaceMaterial( void ); // types.h:334
F2D_API b2RevoluteJointDef b2DefaultRevoluteJointDef( void ); // types.h:830

// ---- geometry helpers used by the Dart shape constructors (raw_box2d_ffi.dart:682-724) -------------------------
F2D_API b2Polygon b2MakeBox( float halfWidth, float halfHeight ); // collision.h:148
F2D_API b2Polygon b2MakeSquare( float halfWidth );				  // collision.h:143
F2D_API b2Polygon b2MakeOffsetRoundedBox( float halfWidth, float halfHeight, b2Vec2 center, b2Rot rotation, float radius ); // :166
F2D_API b2Polygon b2MakePolygon( const b2Hull* hull, float radius );	  // collision.h:124
F2D_API b2Hull b2ComputeHull( const b2Vec2* points, int count );		  // collision.h:228 (B2/src/hull.c)

// ---- world: B2/include/box2d/box2d.h:24-160 -------------------------------------------------------------------
F2D_API b2WorldId b2CreateWorld( const b2WorldDef* def ); // box2d.h:27
F2D_API void b2DestroyWorld( b2WorldId worldId );		  // box2d.h:30
F2D_API bool b2World_IsValid( b2WorldId id );			  // box2d.h:33
/// THE hot path. Replaces B2/src/world.c:695-812 with the CUDA phases in forge2d_b200/csrc/f2d_step.h.
F2D_API void b2World_Step( b2WorldId worldId, float timeStep, int subStepCount ); // box2d.h:39
F2D_API b2BodyEvents b2World_GetBodyEvents( b2WorldId worldId );				  // box2d.h:45
F2D_API b2SensorEvents b2World_GetSensorEvents( b2WorldId worldId );			  // box2d.h:48
F2D_API b2ContactEvents b2World_GetContactEvents( b2WorldId worldId );			  // box2d.h:51
F2D_API void b2World_EnableSleeping( b2WorldId worldId, bool flag );			  // box2d.h:83
F2D_API bool b2World_IsSleepingEnabled( b2WorldId worldId );					  // box2d.h:86
F2D_API void b2World_EnableContinuous( b2WorldId worldId, bool flag );			  // box2d.h:91
F2D_API bool b2World_IsContinuousEnabled( b2WorldId worldId );					  // box2d.h:94
F2D_API void b2World_SetGravity( b2WorldId worldId, b2Vec2 gravity );			  // box2d.h:135
F2D_API b2Vec2 b2World_GetGravity( b2WorldId worldId );							  // box2d.h:138
F2D_API void b2World_EnableWarmStarting( b2WorldId worldId, bool flag );		  // box2d.h:~170
/// Milliseconds per phase of the step, averaged over the steps since the previous call (in-kernel phase marks; the first
/// call switches them on and returns zeros). Replaces box2d.h:169 b2World_GetProfile, layout types.h:466-490.
typedef struct b2Profile
{
	float step, pairs, collide, solve, mergeIslands, prepareStages, solveConstraints, prepareConstraints, integrateVelocities,
		warmStart, solveImpulses, integratePositions, relaxImpulses, applyRestitution, storeImpulses, splitIslands, transforms,
		hitEvents, refit, bullets, sleepIslands, sensors;
} b2Profile;
F2D_API b2Profile b2World_GetProfile( b2WorldId worldId );
F2D_API b2Counters b2World_GetCounters( b2WorldId worldId );					  // box2d.h:~190 (world.c:1865-1891)
F2D_API int b2World_GetAwakeBodyCount( b2WorldId worldId );						  // box2d.h (world.c)

// ---- bodies: B2/include/box2d/box2d.h:200-480 -----------------------------------------------------------------
F2D_API b2BodyId b2CreateBody( b2WorldId worldId, const b2BodyDef* def ); // box2d.h:210
F2D_API bool b2Body_IsValid( b2BodyId id );
F2D_API b2BodyType b2Body_GetType( b2BodyId bodyId );
F2D_API b2Vec2 b2Body_GetPosition( b2BodyId bodyId );
F2D_API b2Rot b2Body_GetRotation( b2BodyId bodyId );
F2D_API b2Transform b2Body_GetTransform( b2BodyId bodyId );
F2D_API b2Vec2 b2Body_GetLinearVelocity( b2BodyId bodyId );
F2D_API float b2Body_GetAngularVelocity( b2BodyId bodyId );
F2D_API void b2Body_SetLinearVelocity( b2BodyId bodyId, b2Vec2 linearVelocity );
F2D_API void b2Body_SetAngularVelocity( b2BodyId bodyId, float angularVelocity );
F2D_API float b2Body_GetMass( b2BodyId bodyId );
F2D_API float b2Body_GetRotationalInertia( b2BodyId bodyId );
F2D_API b2Vec2 b2Body_GetLocalCenterOfMass( b2BodyId bodyId );
F2D_API b2Vec2 b2Body_GetWorldCenterOfMass( b2BodyId bodyId );
F2D_API bool b2Body_IsAwake( b2BodyId bodyId );
F2D_API int b2Body_GetShapeCount( b2BodyId bodyId );
F2D_API int b2Body_GetContactCapacity( b2BodyId bodyId );

// ---- shapes: B2/include/box2d/box2d.h:490-700 -----------------------------------------------------------------
F2D_API b2ShapeId b2CreateCircleShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Circle* circle );	  // box2d.h:497
F2D_API b2ShapeId b2CreateSegmentShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Segment* segment );  // box2d.h:502
F2D_API b2ShapeId b2CreateCapsuleShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Capsule* capsule );  // box2d.h:507
F2D_API b2ShapeId b2CreatePolygonShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Polygon* polygon );  // box2d.h:512
F2D_API bool b2Shape_IsValid( b2ShapeId id );
F2D_API b2BodyId b2Shape_GetBody( b2ShapeId shapeId );
F2D_API b2AABB b2Shape_GetAABB( b2ShapeId shapeId );

// ---- joints: B2/include/box2d/box2d.h:760-1250 ----------------------------------------------------------------
F2D_API b2JointId b2CreateRevoluteJoint( b2WorldId worldId, const b2RevoluteJointDef* def ); // box2d.h:~1050
F2D_API bool b2Joint_IsValid( b2JointId id );

// ---- the rest of the body / shape / joint surface (forge2d_b200/csrc/f2d_capi_ext.inl; reference lines cited there) --
F2D_API void b2DestroyBody( b2BodyId bodyId );													  // body.c:343
F2D_API b2Vec2 b2Body_GetLocalPoint( b2BodyId bodyId, b2Vec2 worldPoint );						  // body.c:656
F2D_API b2Vec2 b2Body_GetWorldPoint( b2BodyId bodyId, b2Vec2 localPoint );						  // body.c:664
F2D_API b2Vec2 b2Body_GetLocalVector( b2BodyId bodyId, b2Vec2 worldVector );						  // body.c:672
F2D_API b2Vec2 b2Body_GetWorldVector( b2BodyId bodyId, b2Vec2 localVector );						  // body.c:680
F2D_API void b2Body_SetTransform( b2BodyId bodyId, b2Vec2 position, b2Rot rotation );				  // body.c:688
F2D_API void b2Body_ApplyForce( b2BodyId bodyId, b2Vec2 force, b2Vec2 point, bool wake );			  // body.c:900
F2D_API void b2Body_ApplyForceToCenter( b2BodyId bodyId, b2Vec2 force, bool wake );				  // body.c:918
F2D_API void b2Body_ApplyTorque( b2BodyId bodyId, float torque, bool wake );						  // body.c:935
F2D_API void b2Body_ApplyLinearImpulse( b2BodyId bodyId, b2Vec2 impulse, b2Vec2 point, bool wake ); // body.c:952
F2D_API void b2Body_ApplyLinearImpulseToCenter( b2BodyId bodyId, b2Vec2 impulse, bool wake );		  // body.c:975
F2D_API void b2Body_ApplyAngularImpulse( b2BodyId bodyId, float impulse, bool wake );				  // body.c:997
F2D_API void b2Body_SetType( b2BodyId bodyId, b2BodyType type );									  // body.c:1036
F2D_API void b2Body_SetName( b2BodyId bodyId, const char* name );									  // body.c:1286
F2D_API const char* b2Body_GetName( b2BodyId bodyId );											  // body.c:1306
F2D_API void b2Body_SetUserData( b2BodyId bodyId, void* userData );
F2D_API void* b2Body_GetUserData( b2BodyId bodyId );
F2D_API void b2Body_SetMassData( b2BodyId bodyId, b2MassData massData );							  // body.c:1357
F2D_API b2MassData b2Body_GetMassData( b2BodyId bodyId );											  // body.c:1384
F2D_API void b2Body_ApplyMassFromShapes( b2BodyId bodyId );										  // body.c:1393
F2D_API void b2Body_SetLinearDamping( b2BodyId bodyId, float linearDamping );
F2D_API float b2Body_GetLinearDamping( b2BodyId bodyId );
F2D_API void b2Body_SetAngularDamping( b2BodyId bodyId, float angularDamping );
F2D_API float b2Body_GetAngularDamping( b2BodyId bodyId );
F2D_API void b2Body_SetGravityScale( b2BodyId bodyId, float gravityScale );
F2D_API float b2Body_GetGravityScale( b2BodyId bodyId );
F2D_API void b2Body_SetAwake( b2BodyId bodyId, bool awake );										  // body.c:1483
F2D_API bool b2Body_IsEnabled( b2BodyId bodyId );
F2D_API bool b2Body_IsSleepEnabled( b2BodyId bodyId );
F2D_API void b2Body_SetSleepThreshold( b2BodyId bodyId, float sleepThreshold );
F2D_API float b2Body_GetSleepThreshold( b2BodyId bodyId );
F2D_API void b2Body_EnableSleep( b2BodyId bodyId, bool enableSleep );								  // body.c:1538
F2D_API void b2Body_Disable( b2BodyId bodyId );													  // body.c:1557
F2D_API void b2Body_Enable( b2BodyId bodyId );													  // body.c:1628
F2D_API void b2Body_SetFixedRotation( b2BodyId bodyId, bool flag );								  // body.c:1722
F2D_API bool b2Body_IsFixedRotation( b2BodyId bodyId );
F2D_API void b2Body_SetBullet( b2BodyId bodyId, bool flag );
F2D_API bool b2Body_IsBullet( b2BodyId bodyId );
F2D_API void b2Body_EnableContactEvents( b2BodyId bodyId, bool flag );
F2D_API void b2Body_EnableHitEvents( b2BodyId bodyId, bool flag );
F2D_API b2WorldId b2Body_GetWorld( b2BodyId bodyId );
F2D_API int b2Body_GetShapes( b2BodyId bodyId, b2ShapeId* shapeArray, int capacity );				  // body.c:1811
F2D_API int b2Body_GetJointCount( b2BodyId bodyId );
F2D_API int b2Body_GetJoints( b2BodyId bodyId, b2JointId* jointArray, int capacity );				  // body.c:1837

F2D_API void b2DestroyShape( b2ShapeId shapeId, bool updateBodyMass );							  // shape.c:318
F2D_API b2WorldId b2Shape_GetWorld( b2ShapeId shapeId );
F2D_API void b2Shape_SetUserData( b2ShapeId shapeId, void* userData );
F2D_API void* b2Shape_GetUserData( b2ShapeId shapeId );
F2D_API bool b2Shape_IsSensor( b2ShapeId shapeId );
F2D_API bool b2Shape_TestPoint( b2ShapeId shapeId, b2Vec2 point );								  // shape.c:979
F2D_API void b2Shape_SetDensity( b2ShapeId shapeId, float density, bool updateBodyMass );			  // shape.c:1055
F2D_API float b2Shape_GetDensity( b2ShapeId shapeId );
F2D_API void b2Shape_SetFriction( b2ShapeId shapeId, float friction );
F2D_API float b2Shape_GetFriction( b2ShapeId shapeId );
F2D_API void b2Shape_SetRestitution( b2ShapeId shapeId, float restitution );
F2D_API float b2Shape_GetRestitution( b2ShapeId shapeId );
F2D_API b2Filter b2Shape_GetFilter( b2ShapeId shapeId );
F2D_API void b2Shape_SetFilter( b2ShapeId shapeId, b2Filter filter );								  // shape.c:1235
F2D_API void b2Shape_EnableSensorEvents( b2ShapeId shapeId, bool flag );
F2D_API bool b2Shape_AreSensorEventsEnabled( b2ShapeId shapeId );
F2D_API void b2Shape_EnableContactEvents( b2ShapeId shapeId, bool flag );
F2D_API bool b2Shape_AreContactEventsEnabled( b2ShapeId shapeId );
F2D_API void b2Shape_EnablePreSolveEvents( b2ShapeId shapeId, bool flag );
F2D_API bool b2Shape_ArePreSolveEventsEnabled( b2ShapeId shapeId );
F2D_API void b2Shape_EnableHitEvents( b2ShapeId shapeId, bool flag );
F2D_API bool b2Shape_AreHitEventsEnabled( b2ShapeId shapeId );
F2D_API b2ShapeType b2Shape_GetType( b2ShapeId shapeId );
F2D_API b2Circle b2Shape_GetCircle( b2ShapeId shapeId );
F2D_API b2Segment b2Shape_GetSegment( b2ShapeId shapeId );
F2D_API b2ChainSegment b2Shape_GetChainSegment( b2ShapeId shapeId );
F2D_API b2Capsule b2Shape_GetCapsule( b2ShapeId shapeId );
F2D_API b2Polygon b2Shape_GetPolygon( b2ShapeId shapeId );

F2D_API b2DistanceJointDef b2DefaultDistanceJointDef( void );   // joint.c:24
F2D_API b2MotorJointDef b2DefaultMotorJointDef( void );		   // joint.c:33
F2D_API b2MouseJointDef b2DefaultMouseJointDef( void );		   // joint.c:43
F2D_API b2FilterJointDef b2DefaultFilterJointDef( void );	   // joint.c:53
F2D_API b2PrismaticJointDef b2DefaultPrismaticJointDef( void ); // joint.c:60
F2D_API b2WeldJointDef b2DefaultWeldJointDef( void );		   // joint.c:76
F2D_API b2WheelJointDef b2DefaultWheelJointDef( void );		   // joint.c:83
F2D_API b2ExplosionDef b2DefaultExplosionDef( void );		   // joint.c:94
F2D_API b2JointId b2CreateDistanceJoint( b2WorldId worldId, const b2DistanceJointDef* def );	 // joint.c:353
F2D_API b2JointId b2CreateMotorJoint( b2WorldId worldId, const b2MotorJointDef* def );		 // joint.c:406
F2D_API b2JointId b2CreateMouseJoint( b2WorldId worldId, const b2MouseJointDef* def );		 // joint.c:444
F2D_API b2JointId b2CreateFilterJoint( b2WorldId worldId, const b2FilterJointDef* def );		 // joint.c:480
F2D_API b2JointId b2CreatePrismaticJoint( b2WorldId worldId, const b2PrismaticJointDef* def ); // joint.c:559
F2D_API b2JointId b2CreateWeldJoint( b2WorldId worldId, const b2WeldJointDef* def );			 // joint.c:609
F2D_API b2JointId b2CreateWheelJoint( b2WorldId worldId, const b2WheelJointDef* def );		 // joint.c:651
F2D_API void b2DestroyJoint( b2JointId jointId );												 // joint.c:809
F2D_API b2JointType b2Joint_GetType( b2JointId jointId );
F2D_API b2BodyId b2Joint_GetBodyA( b2JointId jointId );
F2D_API b2BodyId b2Joint_GetBodyB( b2JointId jointId );
F2D_API b2WorldId b2Joint_GetWorld( b2JointId jointId );
F2D_API b2Vec2 b2Joint_GetLocalAnchorA( b2JointId jointId );
F2D_API b2Vec2 b2Joint_GetLocalAnchorB( b2JointId jointId );
F2D_API void b2Joint_SetLocalAnchorA( b2JointId jointId, b2Vec2 localAnchor );
F2D_API void b2Joint_SetLocalAnchorB( b2JointId jointId, b2Vec2 localAnchor );
F2D_API void b2Joint_SetCollideConnected( b2JointId jointId, bool shouldCollide );			 // joint.c:979
F2D_API bool b2Joint_GetCollideConnected( b2JointId jointId );
F2D_API void b2Joint_SetUserData( b2JointId jointId, void* userData );
F2D_API void* b2Joint_GetUserData( b2JointId jointId );
F2D_API void b2Joint_WakeBodies( b2JointId jointId );											 // joint.c:1045
F2D_API b2Vec2 b2Joint_GetConstraintForce( b2JointId jointId );								 // joint.c:1061
F2D_API float b2Joint_GetConstraintTorque( b2JointId jointId );								 // joint.c:1099
// ---- chains: shape.c:339-577, :1476-1576 -------------------------------------------------------------------------
F2D_API b2ChainDef b2DefaultChainDef( void );													 // types.c:76
F2D_API b2ChainId b2CreateChain( b2BodyId bodyId, const b2ChainDef* def );					 // shape.c:339
F2D_API void b2DestroyChain( b2ChainId chainId );												 // shape.c:488
F2D_API bool b2Chain_IsValid( b2ChainId id );
F2D_API b2WorldId b2Chain_GetWorld( b2ChainId chainId );
F2D_API int b2Chain_GetSegmentCount( b2ChainId chainId );
F2D_API int b2Chain_GetSegments( b2ChainId chainId, b2ShapeId* segmentArray, int capacity );	 // shape.c:557
F2D_API void b2Chain_SetFriction( b2ChainId chainId, float friction );						 // shape.c:1476
F2D_API float b2Chain_GetFriction( b2ChainId chainId );
F2D_API void b2Chain_SetRestitution( b2ChainId chainId, float restitution );					 // shape.c:1511
F2D_API float b2Chain_GetRestitution( b2ChainId chainId );

// ---- queries and explosion (run on the host image of the device state: forge2d_b200/csrc/f2d_query.h) ---------------
F2D_API b2QueryFilter b2DefaultQueryFilter( void );
F2D_API b2TreeStats b2World_OverlapAABB( b2WorldId worldId, b2AABB aabb, b2QueryFilter filter, b2OverlapResultFcn* fcn, void* context ); // world.c:2071
F2D_API b2TreeStats b2World_CastRay( b2WorldId worldId, b2Vec2 origin, b2Vec2 translation, b2QueryFilter filter, b2CastResultFcn* fcn,
									 void* context );																				 // world.c:2222
F2D_API b2RayResult b2World_CastRayClosest( b2WorldId worldId, b2Vec2 origin, b2Vec2 translation, b2QueryFilter filter );			 // world.c:2277
F2D_API void b2World_Explode( b2WorldId worldId, const b2ExplosionDef* explosionDef );												 // world.c:2718
/// Host callbacks from inside the step cannot run on the device path: registering one reports an error (loudly).
F2D_API void b2World_SetCustomFilterCallback( b2WorldId worldId, b2CustomFilterFcn* fcn, void* context );
/// types.c:136-151: a b2DebugDraw whose callbacks are all no-ops and whose options are all off.
F2D_API b2DebugDraw b2DefaultDebugDraw( void );
/// box2d.h:92 / world.c:1161-1489: emits the world's debug geometry through the callbacks of `draw` (host callbacks, on
/// the calling thread, during the call), in the reference's order. Reads the device-resident state.
F2D_API void b2World_Draw( b2WorldId worldId, b2DebugDraw* draw );
F2D_API void b2World_SetPreSolveCallback( b2WorldId worldId, b2PreSolveFcn* fcn, void* context );
#include "forge2d_b200_joints.h"

// ---------------------------------------------------------------------------------------------------------------
// Extension (additive, `f2d` prefix): batches of independent worlds, sharded by world — config 5 of BASELINE.json.
// Stock b2WorldId cannot address 8192 worlds (B2_MAX_WORLDS = 128, constants.h:26-28), hence a separate handle.
typedef struct f2dBatch f2dBatch;
/// Replicates the current state of `templateWorld` into `count` device-resident worlds (one image each).
/// Bytes moved between host and device by single-world calls since the library was loaded (what a frame costs: after
/// a step, position / velocity getters fetch the body arrays once; forces, impulses and velocities of awake bodies go
/// up as the few bytes they changed).
F2D_API void f2dGetTransferBytes( unsigned long long* hostToDevice, unsigned long long* deviceToHost );
F2D_API f2dBatch* f2dBatch_Create( b2WorldId templateWorld, int count );
/// A batch of DIFFERENT worlds (ids may repeat): all are brought to one common image layout and uploaded; the host
/// worlds stay usable on their own. Worlds with host callbacks registered are refused (no host in a batch step).
F2D_API f2dBatch* f2dBatch_CreateFromWorlds( const b2WorldId* worlds, int count );
F2D_API void f2dBatch_Destroy( f2dBatch* batch );
/// One b2World_Step for every world of the batch; a single kernel sequence, one thread block per world.
F2D_API void f2dBatch_Step( f2dBatch* batch, float timeStep, int subStepCount );
/// Queues `steps` steps without host synchronisation in between (bench / headless simulation).
F2D_API void f2dBatch_StepN( f2dBatch* batch, float timeStep, int subStepCount, int steps );
F2D_API void f2dBatch_Synchronize( f2dBatch* batch );
/// Threads per world (block size) and resident blocks per SM the batch kernel is compiled for; returns 0 if unknown.
/// Available: 128x8 (default), 64x16, 32x32, 256x4, 256x2.
F2D_API int f2dBatch_SetLaunchConfig( f2dBatch* batch, int threadsPerWorld, int blocksPerSM );
/// 1 (the default): several worlds per thread block, kept in the same phase of the step (instruction-cache friendly);
/// 0: one world per block in the configuration of f2dBatch_SetLaunchConfig (which also switches to that mode).
F2D_API void f2dBatch_SetGangMode( f2dBatch* batch, int on );
F2D_API int f2dBatch_GetWorldCount( f2dBatch* batch );
/// Body move events of every world -> host buffer: `out` receives count*maxBodies records, `counts[w]` valid ones.
F2D_API int f2dBatch_GetBodyEvents( f2dBatch* batch, b2BodyMoveEvent* out, int maxBodiesPerWorld, int* counts );
/// Same records through a pinned staging buffer owned by the batch; pointers valid until the next call.
F2D_API int f2dBatch_ReadBodyEvents( f2dBatch* batch, int maxBodiesPerWorld, const b2BodyMoveEvent** outEvents, const int** outCounts );
/// f2dBatch_Step + f2dBatch_ReadBodyEvents as one call (same results): the worlds are stepped in slices on separate
/// streams and each slice's events cross PCIe while later slices are still being stepped. Returns the event total.
F2D_API int f2dBatch_StepAndReadBodyEvents( f2dBatch* batch, float timeStep, int subStepCount, int maxBodiesPerWorld,
											const b2BodyMoveEvent** outEvents, const int** outCounts );
/// Pipelined step + read-back: call k queues step k and the copy of its results to pinned host memory, and returns the
/// results of step k-1 (nothing on the first call), whose copy overlapped the computation of step k.
/// f2dBatch_FlushPipelined returns the results of the last queued step. format 0: b2BodyMoveEvent records (40 bytes per
/// body); format 1: b2Transform only (16 bytes per body, awake order). Pointers stay valid until the next call.
F2D_API int f2dBatch_StepPipelined( f2dBatch* batch, float timeStep, int subStepCount, int maxBodiesPerWorld, int format,
									const void** outRecords, const int** outCounts );
F2D_API int f2dBatch_FlushPipelined( f2dBatch* batch, const void** outRecords, const int** outCounts );
/// Per-world gravity: the batch counterpart of b2World_SetGravity (box2d.h:135); `gravity` holds `count` vectors.
F2D_API void f2dBatch_SetGravity( f2dBatch* batch, const b2Vec2* gravity, int count );
/// Translates every body, shape box and broadphase box of world k by offsets[k]: replicas placed side by side, or
/// decorrelated (same physics, different floating-point rounding).
F2D_API void f2dBatch_TranslateWorlds( f2dBatch* batch, const b2Vec2* offsets, int count );
/// CUDA events on the batch's stream (slots 0..7) so callers can time device work without a torch stream.
F2D_API void f2dBatch_EventRecord( f2dBatch* batch, int slot );
F2D_API float f2dBatch_EventElapsedMs( f2dBatch* batch, int fromSlot, int toSlot );
F2D_API unsigned long long f2dBatch_GetWorldBytes( f2dBatch* batch ); ///< HBM bytes per world image
/// Copies world `index` of the batch back into an ordinary world (inspection, parity tests).
F2D_API void f2dBatch_DownloadWorld( f2dBatch* batch, int index, b2WorldId into );
/// Per-world x-translation of every body by `index * dx` (decorrelates otherwise identical worlds).
F2D_API uint32_t f2dBatch_GetErrorFlags( f2dBatch* batch );
/// Error flags per world (0 = fine); returns how many worlds have a flag up. A world whose new contacts exceed its
/// capacity never stays behind: the batch grows every image on the device and that world repeats the step
/// (f2dBatch_GetGrowthCount counts how often that happened).
F2D_API int f2dBatch_GetWorldErrors( f2dBatch* batch, uint32_t* out, int cap );
F2D_API int f2dBatch_GetGrowthCount( f2dBatch* batch );

/// Library diagnostics
F2D_API int f2dHasDevice( void );			///< 1 when a CUDA device is usable
F2D_API int f2dSetDevice( int device );		///< selects the CUDA device of the calling thread (one process per GPU)
F2D_API void* f2dHostAlloc( unsigned long long bytes ); ///< pinned host memory for batch inputs
F2D_API void f2dHostFree( void* p );
F2D_API const char* f2dGetLastError( void ); ///< last error message ("" if none)
F2D_API void f2dClearLastError( void );
/// 0 = one block per world (default for small worlds), 1 = cooperative grid per world, -1 = automatic
F2D_API void f2dWorld_SetLaunchMode( b2WorldId worldId, int mode );
F2D_API uint32_t f2dWorld_GetErrorFlags( b2WorldId worldId );
F2D_API long long f2dWorld_GetKernelLaunchCount( void ); ///< kernels launched by this library so far (bench accounting)
/// Times of the last step's phases in ms measured with CUDA events: [pairs, collide, solve, finalize, total]
F2D_API void f2dWorld_GetLastStepTimes( b2WorldId worldId, float* out5 );
F2D_API void f2dWorld_EnablePhaseTiming( b2WorldId worldId, bool flag );
/// What the last step did: [islandPath, activeColours, awakeContacts, awakeBodies, maxIslandContacts, maxIslandBodies,
/// awakeIslands, mergedIslands, splitBodies, splitComponents, splitContacts, splitJoints]
F2D_API int f2dWorld_GetStepInfo( b2WorldId worldId, int* out, int cap );
/// In-kernel phase profile: ns per sub-phase (f2d::ProfSlot order) accumulated on the device since enabled
F2D_API void f2dWorld_EnableProfile( b2WorldId worldId, bool flag );
/* Measurement aid: bin the narrowphase work list by pair class (default on; only worlds with several shape types bin). */
F2D_API void f2dWorld_EnablePairClassBinning( b2WorldId worldId, bool flag );
F2D_API int f2dWorld_ReadProfile( b2WorldId worldId, unsigned long long* out, int cap );
/// Device-side step without the per-step host synchronisation / header readback (bench: inputs resident in HBM)
F2D_API void f2dWorld_StepAsync( b2WorldId worldId, float timeStep, int subStepCount );
F2D_API void f2dWorld_Synchronize( b2WorldId worldId );

// ---- introspection used by the parity tests (records in forge2d_b200_debug.h) ----------------------------------
struct f2dBodyRecord;
struct f2dContactRecord;
struct f2dIslandRecord;
struct f2dShapeRecord;
struct f2dTreeLeafRecord;
struct f2dJointRecord;
F2D_API int f2dDebug_AwakeOrder( b2WorldId id, int* bodyIds, int cap );
F2D_API int f2dDebug_MoveArray( b2WorldId id, int* keys, int cap );
F2D_API int f2dDebug_Bodies( b2WorldId id, struct f2dBodyRecord* out, int cap );
F2D_API int f2dDebug_Contacts( b2WorldId id, struct f2dContactRecord* out, int cap );
F2D_API int f2dDebug_Islands( b2WorldId id, struct f2dIslandRecord* out, int cap );
F2D_API int f2dDebug_Shapes( b2WorldId id, struct f2dShapeRecord* out, int cap );
F2D_API int f2dDebug_Tree( b2WorldId id, int treeType, struct f2dTreeLeafRecord* out, int cap );
F2D_API int f2dDebug_Joints( b2WorldId id, struct f2dJointRecord* out, int cap );
F2D_API void f2dDebug_ColorCounts( b2WorldId id, int* contactCounts, int* jointCounts );
F2D_API int f2dDebug_ColorContacts( b2WorldId id, int colorIndex, int* contactIds, int cap );
F2D_API int f2dDebug_AwakeContacts( b2WorldId id, int* contactIds, int cap );
F2D_API int f2dDebug_AwakeIslands( b2WorldId id, int* islandIds, int cap );

#ifdef __cplusplus
}
#endif
