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

// ---------------------------------------------------------------------------------------------------------------
// Extension (additive, `f2d` prefix): batches of independent worlds, sharded by world — config 5 of BASELINE.json.
// Stock b2WorldId cannot address 8192 worlds (B2_MAX_WORLDS = 128, constants.h:26-28), hence a separate handle.
typedef struct f2dBatch f2dBatch;
/// Replicates the current state of `templateWorld` into `count` device-resident worlds (one image each).
F2D_API f2dBatch* f2dBatch_Create( b2WorldId templateWorld, int count );
F2D_API void f2dBatch_Destroy( f2dBatch* batch );
/// One b2World_Step for every world of the batch; a single kernel sequence, one thread block per world.
F2D_API void f2dBatch_Step( f2dBatch* batch, float timeStep, int subStepCount );
/// Queues `steps` steps without host synchronisation in between (bench / headless simulation).
F2D_API void f2dBatch_StepN( f2dBatch* batch, float timeStep, int subStepCount, int steps );
F2D_API void f2dBatch_Synchronize( f2dBatch* batch );
/// Threads per world (block size) and resident blocks per SM the batch kernel is compiled for; returns 0 if unknown.
/// Available: 256x2 (default), 128x4, 64x8, 32x16, 128x8, 64x16.
F2D_API int f2dBatch_SetLaunchConfig( f2dBatch* batch, int threadsPerWorld, int blocksPerSM );
F2D_API int f2dBatch_GetWorldCount( f2dBatch* batch );
/// Body move events of every world -> host buffer: `out` receives count*maxBodies records, `counts[w]` valid ones.
F2D_API int f2dBatch_GetBodyEvents( f2dBatch* batch, b2BodyMoveEvent* out, int maxBodiesPerWorld, int* counts );
/// Same records through a pinned staging buffer owned by the batch; pointers valid until the next call.
F2D_API int f2dBatch_ReadBodyEvents( f2dBatch* batch, int maxBodiesPerWorld, const b2BodyMoveEvent** outEvents, const int** outCounts );
/// Per-world gravity: the batch counterpart of b2World_SetGravity (box2d.h:135); `gravity` holds `count` vectors.
F2D_API void f2dBatch_SetGravity( f2dBatch* batch, const b2Vec2* gravity, int count );
/// CUDA events on the batch's stream (slots 0..7) so callers can time device work without a torch stream.
F2D_API void f2dBatch_EventRecord( f2dBatch* batch, int slot );
F2D_API float f2dBatch_EventElapsedMs( f2dBatch* batch, int fromSlot, int toSlot );
F2D_API unsigned long long f2dBatch_GetWorldBytes( f2dBatch* batch ); ///< HBM bytes per world image
/// Copies world `index` of the batch back into an ordinary world (inspection, parity tests).
F2D_API void f2dBatch_DownloadWorld( f2dBatch* batch, int index, b2WorldId into );
/// Per-world x-translation of every body by `index * dx` (decorrelates otherwise identical worlds).
F2D_API uint32_t f2dBatch_GetErrorFlags( f2dBatch* batch );

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
