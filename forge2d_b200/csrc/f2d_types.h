// forge2d_b200 — world image layout.
//
// A world is ONE position-independent block of memory ("image"): a World header followed by every array the
// step touches, addressed by byte offsets from the header. The same bytes are valid on the host and in HBM, so
//   * host-side API calls (create body/shape/joint ...) mutate a host image with the same code the device runs,
//   * upload/download is a single memcpy, and
//   * a batch of N independent worlds is N images at a fixed stride in one HBM allocation (sharded by world).
//
// Unlike the reference (B2/src/world.h:44-175, solver_set.h:19-43) simulation records never move between
// "solver sets": BodySim / ContactSim / JointSim live in stable slots indexed by their id, and the set / colour
// membership that defines the reference's iteration ORDER (awake body order, colour array order, sleeping-set
// order) is kept in compact int32 id lists that are appended / swap-removed exactly where the
// reference appends / swap-removes whole structs. Order-dependent results are therefore identical while the
// bytes moved per structural edit drop from 100-196 B to 4 B.
#pragma once
#include "f2d_math.h"
#include <string.h>

namespace f2d
{

template <class T> struct Arr
{
	uint64_t off; // bytes from the World header
	int32_t count;
	int32_t cap;
};

// LIFO id pool, semantics of B2/src/id_pool.c:19-38 (contact-id reuse order feeds colouring order).
struct IdPool
{
	Arr<int32_t> free;
	int32_t next;
	int32_t pad;
};

enum : int
{
	kStaticSet = 0,
	kDisabledSet = 1,
	kAwakeSet = 2,
	kFirstSleepingSet = 3
};
enum : int
{
	kStaticBody = 0,
	kKinematicBody = 1,
	kDynamicBody = 2
};
enum : int
{
	kCircle = 0,
	kCapsule = 1,
	kSegment = 2,
	kPolygon = 3,
	kChainSegment = 4,
	kShapeTypeCount = 5
};
enum : int
{
	kDistanceJoint = 0,
	kFilterJoint,
	kMotorJoint,
	kMouseJoint,
	kPrismaticJoint,
	kRevoluteJoint,
	kWeldJoint,
	kWheelJoint
};

// --- public-ABI-compatible value types (layouts fixed by B2/include/box2d/collision.h, types.h, id.h) ----------
struct ShapeId
{
	int32_t index1;
	uint16_t world0, generation;
};
struct BodyId
{
	int32_t index1;
	uint16_t world0, generation;
};
struct ManifoldPoint // collision.h:498-529, 48 B
{
	V2 point, anchorA, anchorB;
	float separation, normalImpulse, tangentImpulse, totalNormalImpulse, normalVelocity;
	uint16_t id;
	bool persisted;
};
struct Manifold // collision.h:532-552, 112 B
{
	V2 normal;
	float rollingImpulse;
	ManifoldPoint points[2];
	int32_t pointCount;
};
struct Circle
{
	V2 center;
	float radius;
};
struct Capsule
{
	V2 c1, c2;
	float radius;
};
struct Poly // b2Polygon, 144 B
{
	V2 v[kMaxPolyVerts];
	V2 n[kMaxPolyVerts];
	V2 centroid;
	float radius;
	int32_t count;
};
struct Segment
{
	V2 p1, p2;
};
struct ChainSegment
{
	V2 ghost1;
	Segment segment;
	V2 ghost2;
	int32_t chainId;
};
struct Filter // types.h b2Filter
{
	uint64_t category, mask;
	int32_t group;
};
struct SimplexCache
{
	uint16_t count;
	uint8_t indexA[3], indexB[3];
};
struct BodyMoveEvent // types.h:1136-1142, 40 B
{
	Xf transform;
	BodyId bodyId;
	uint64_t userData;
	bool fellAsleep;
};
struct BeginTouchEvent // types.h:1050
{
	ShapeId a, b;
	Manifold manifold;
};
struct EndTouchEvent
{
	ShapeId a, b;
};
struct HitEvent // types.h:1082
{
	ShapeId a, b;
	V2 point, normal;
	float approachSpeed;
};
struct SensorEvent
{
	ShapeId sensor, visitor;
};

// --- internal records ---------------------------------------------------------------------------------------
// Record layout rule (all records below): a record is a whole number of 32-byte DRAM sectors, starts on a sector
// boundary (arrays start 256-byte aligned, f2d_image.h), and its fields are grouped so that one phase of the step finds
// what it reads in as few sectors as possible; `alignas` lets the compiler move 8- and 16-byte groups with one
// LDG.64 / STG.128 instead of one 4-byte access per field. Field meaning follows the reference; only the order is ours.

// Body data apart from the simulation record: B2/src/body.h:14-64
struct alignas( 32 ) Body
{
	// sector 0: narrowphase (setIndex, localIndex) and pair finding (contact / joint lists)
	int32_t setIndex, localIndex; // localIndex = position in the owning set's id list
	int32_t headContactKey, contactCount;
	int32_t headJointKey, jointCount;
	int32_t id, shapeCount;
	// sector 1: everything finalize reads and writes of this record (island vote, sleep, shape list, move event)
	int32_t islandId, headShapeId;
	float sleepThreshold, sleepTime;
	uint64_t userData;
	uint16_t generation;
	int8_t type;
	bool enableSleep : 1, isSpeedCapped : 1, fixedRotation : 1, isMarked : 1;
	int32_t bodyMoveIndex;
	// sector 2
	int32_t islandPrev, islandNext, headChainId;
	float mass, inertia;
	uint16_t colorMask; // per-colour membership bits (replaces the per-colour body bitsets, constraint_graph.h:24-42)
	uint16_t pad0;
	int32_t pad1[2];
	// sector 3
	char name[32];
};
static_assert( sizeof( Body ) == 128, "Body layout" );
// Body simulation data (stable slot = body id): B2/src/body.h:120-159
struct alignas( 32 ) BodySim
{
	// sector 0: everything the narrowphase reads of a body
	Xf transform;
	V2 localCenter;
	float invMass, invInertia;
	// sector 1: finalize
	V2 center;
	V2 center0;
	Rot rotation0;
	float minExtent, maxExtent;
	// sector 2: velocity integration constants, flags
	V2 force;
	float torque, linearDamping, angularDamping, gravityScale;
	bool isFast, isBullet, isSpeedCapped, allowFastRotation, enlargeAABB;
	bool pad0[3];
};
static_assert( sizeof( BodySim ) == 96, "BodySim layout" );
// Solver state, dense by awake index, 32 B = two float4: B2/src/body.h:66-83
struct alignas( 16 ) BodyState
{
	V2 v;
	float w;
	int32_t flags;
	V2 dp;
	Rot dq;
};
// B2/src/shape.h:13-52
struct alignas( 32 ) Shape
{
	// sector 0: boxes
	Box aabb, fatAABB;
	// sector 1: links, type, flags, friction
	int32_t type, bodyId, nextShapeId, proxyKey, sensorIndex;
	uint16_t generation;
	bool enableSensorEvents, enableContactEvents, enableHitEvents, enablePreSolveEvents, enlargedAABB;
	bool pad0;
	float friction;
	// sector 2: the rest of the material (narrowphase), cold ids
	float restitution, rollingResistance, tangentSpeed, density;
	int32_t userMaterialId, id, prevShapeId;
	uint32_t customColor;
	// sector 3 onwards: geometry (a box: vertices 96-127, normals 160-191, centroid / radius / count 224-239)
	union
	{
		Capsule capsule;
		Circle circle;
		Poly polygon;
		Segment segment;
		ChainSegment chainSegment;
	};
	// pair finding
	Filter filter;
	V2 localCentroid;
	uint64_t userData;
	int32_t pad1[2];
};
static_assert( sizeof( Shape ) == 288, "Shape layout" );

enum : uint32_t // contact.h:15-25
{
	kContactTouching = 0x1,
	kContactHitEvent = 0x2,
	kContactEnableContactEvents = 0x4
};
enum : uint32_t // contact.h:74-93
{
	kSimTouching = 0x00010000,
	kSimDisjoint = 0x00020000,
	kSimStartedTouching = 0x00040000,
	kSimStoppedTouching = 0x00080000,
	kSimEnableHitEvent = 0x00100000,
	kSimEnablePreSolve = 0x00200000,
	// ours, only ever set between the two halves of a callback-mediated narrowphase (f2d_step.h deferPreSolve): the
	// manifold holds the raw result of the manifold function and the host's pre-solve verdict is pending
	kSimPendingPreSolve = 0x00400000
};
struct Edge
{
	int32_t bodyId, prevKey, nextKey;
};
// Contact bookkeeping (slot = contact id): B2/src/contact.h:41-71. The reference's isMarked is the kContactMarked bit of flags.
constexpr uint32_t kContactMarked = 0x80000000u;
struct alignas( 32 ) Contact
{
	// sector 0: what walking a body's contact list reads (pair finding, island split rows)
	int32_t shapeIdA, shapeIdB;
	Edge edges[2];
	// sector 1
	int32_t setIndex, colorIndex, localIndex;
	int32_t islandPrev, islandNext, islandId;
	int32_t contactId;
	uint32_t flags;
};
static_assert( sizeof( Contact ) == 64, "Contact layout" );
// The manifold of a contact as it is STORED (the compute type, and the layout the public API hands out, is Manifold):
// the same 112 bytes regrouped into seven 16-byte chunks by who reads them together. The narrowphase needs chunks M0 and
// M1 of the previous step (feature ids and accumulated impulses, contact.c:552-586), prepare M0-M4, the impulse store
// writes M1 and M6, hit events read M2, M5 and M6.
struct alignas( 16 ) StoredManifold
{
	int32_t pointCount;																  // M0
	uint32_t ids; // points[0].id | points[1].id << 16
	float rollingImpulse;
	uint32_t persisted; // bit k: points[k].persisted
	float normalImpulse0, tangentImpulse0, normalImpulse1, tangentImpulse1;			  // M1
	V2 normal;																		  // M2
	float separation0, separation1;
	V2 anchorA0, anchorB0;															  // M3
	V2 anchorA1, anchorB1;															  // M4
	V2 point0, point1;																  // M5
	float totalNormalImpulse0, totalNormalImpulse1, normalVelocity0, normalVelocity1; // M6
};
static_assert( sizeof( StoredManifold ) == 112, "StoredManifold layout" );
F2D_HD void packManifold( StoredManifold& s, const Manifold& m )
{
	const ManifoldPoint& p0 = m.points[0];
	const ManifoldPoint& p1 = m.points[1];
	const uint32_t ids = (uint32_t)p0.id | ( (uint32_t)p1.id << 16 );
	const uint32_t persisted = ( p0.persisted ? 1u : 0u ) | ( p1.persisted ? 2u : 0u );
	store16( &s.pointCount, Q4{ floatFromBits( (uint32_t)m.pointCount ), floatFromBits( ids ), m.rollingImpulse, floatFromBits( persisted ) } );
	store16( &s.normalImpulse0, Q4{ p0.normalImpulse, p0.tangentImpulse, p1.normalImpulse, p1.tangentImpulse } );
	store16( &s.normal, Q4{ m.normal.x, m.normal.y, p0.separation, p1.separation } );
	store16( &s.anchorA0, Q4{ p0.anchorA.x, p0.anchorA.y, p0.anchorB.x, p0.anchorB.y } );
	store16( &s.anchorA1, Q4{ p1.anchorA.x, p1.anchorA.y, p1.anchorB.x, p1.anchorB.y } );
	store16( &s.point0, Q4{ p0.point.x, p0.point.y, p1.point.x, p1.point.y } );
	store16( &s.totalNormalImpulse0, Q4{ p0.totalNormalImpulse, p1.totalNormalImpulse, p0.normalVelocity, p1.normalVelocity } );
}
F2D_HD Manifold unpackManifold( const StoredManifold& s )
{
	const Q4 m0 = load16( &s.pointCount ), m1 = load16( &s.normalImpulse0 ), m2 = load16( &s.normal ), m3 = load16( &s.anchorA0 ),
			 m4 = load16( &s.anchorA1 ), m5 = load16( &s.point0 ), m6 = load16( &s.totalNormalImpulse0 );
	const uint32_t ids = floatBits( m0.y ), persisted = floatBits( m0.w );
	Manifold m;
	memset( &m, 0, sizeof( m ) ); // padding bytes too: manifolds are copied into event records
	m.normal = V2{ m2.x, m2.y };
	m.rollingImpulse = m0.z;
	m.pointCount = (int32_t)floatBits( m0.x );
	ManifoldPoint& p0 = m.points[0];
	ManifoldPoint& p1 = m.points[1];
	p0.point = V2{ m5.x, m5.y };
	p1.point = V2{ m5.z, m5.w };
	p0.anchorA = V2{ m3.x, m3.y };
	p0.anchorB = V2{ m3.z, m3.w };
	p1.anchorA = V2{ m4.x, m4.y };
	p1.anchorB = V2{ m4.z, m4.w };
	p0.separation = m2.z;
	p1.separation = m2.w;
	p0.normalImpulse = m1.x;
	p0.tangentImpulse = m1.y;
	p1.normalImpulse = m1.z;
	p1.tangentImpulse = m1.w;
	p0.totalNormalImpulse = m6.x;
	p1.totalNormalImpulse = m6.y;
	p0.normalVelocity = m6.z;
	p1.normalVelocity = m6.w;
	p0.id = (uint16_t)( ids & 0xffffu );
	p1.id = (uint16_t)( ids >> 16 );
	p0.persisted = ( persisted & 1u ) != 0;
	p1.persisted = ( persisted & 2u ) != 0;
	return m;
}
// Contact simulation data (stable slot = contact id): B2/src/contact.h:98-131. Sector 0 and 1 are all the narrowphase
// reads of a contact before it runs the manifold function.
struct alignas( 32 ) ContactSim
{
	int32_t shapeIdA, shapeIdB;
	int32_t bodyIdA, bodyIdB; // owners of the two shapes (ours: saves the shape -> body hop of the narrowphase gather)
	uint32_t simFlags;
	SimplexCache cache;
	int32_t pairClass; // typeA * kShapeTypeCount + typeB (fixed at creation): the narrowphase bins its work list by it
	StoredManifold manifold;
	int32_t bodySimIndexA, bodySimIndexB; // awake indices or kNull
	float friction, restitution;
	float invMassA, invIA, invMassB, invIB;
	float rollingResistance, tangentSpeed;
	int32_t pad1, pad2;
};
static_assert( sizeof( ContactSim ) == 192, "ContactSim layout" );
// B2/src/sensor.h:11-24. The two overlap lists of a sensor are blocks of World::sensorRefs
// (block = sensor index * 2 * World::sensorOverlapCap, second list sensorOverlapCap further); `flip` says which is
// "overlaps2". The per-sensor capacity follows the number of shapes of the world (f2d_image.h: an overlap list can never
// hold more), so a level-wide trigger region cannot overflow it.
constexpr int kMinSensorOverlapCap = 64;
struct ShapeRef
{
	int32_t shapeId;
	uint16_t generation, pad;
};
struct Sensor
{
	int32_t shapeId;
	int32_t count1, count2; // overlaps1 = previous step, overlaps2 = this step
	int32_t flip;			// 0: overlaps2 is the first half of the block, 1: the second half
};
// B2/src/island.h:25-57
struct Island
{
	int32_t setIndex, localIndex, islandId;
	int32_t headBody, tailBody, bodyCount;
	int32_t headContact, tailContact, contactCount;
	int32_t headJoint, tailJoint, jointCount;
	int32_t parentIsland, constraintRemoveCount;
};
// B2/src/joint.h:22-50
struct Joint
{
	uint64_t userData;
	int32_t setIndex, colorIndex, localIndex;
	Edge edges[2];
	int32_t jointId, islandId, islandPrev, islandNext;
	float drawSize;
	int32_t type;
	uint16_t generation;
	bool isMarked, collideConnected;
};
// Per-type joint data: B2/src/joint.h:52-245
struct DistanceJointData
{
	float length, hertz, dampingRatio, minLength, maxLength, maxMotorForce, motorSpeed;
	float impulse, lowerImpulse, upperImpulse, motorImpulse;
	int32_t indexA, indexB;
	V2 anchorA, anchorB, deltaCenter;
	Soft distanceSoftness;
	float axialMass;
	bool enableSpring, enableLimit, enableMotor;
};
struct MotorJointData
{
	V2 linearOffset;
	float angularOffset;
	V2 linearImpulse;
	float angularImpulse, maxForce, maxTorque, correctionFactor;
	int32_t indexA, indexB;
	V2 anchorA, anchorB, deltaCenter;
	float deltaAngle;
	M22 linearMass;
	float angularMass;
};
struct MouseJointData
{
	V2 targetA;
	float hertz, dampingRatio, maxForce;
	V2 linearImpulse;
	float angularImpulse;
	Soft linearSoftness, angularSoftness;
	int32_t indexB;
	V2 anchorB, deltaCenter;
	M22 linearMass;
};
struct PrismaticJointData
{
	V2 localAxisA, impulse;
	float springImpulse, motorImpulse, lowerImpulse, upperImpulse;
	float hertz, dampingRatio, targetTranslation, maxMotorForce, motorSpeed, referenceAngle, lowerTranslation, upperTranslation;
	int32_t indexA, indexB;
	V2 anchorA, anchorB, axisA, deltaCenter;
	float deltaAngle, axialMass;
	Soft springSoftness;
	bool enableSpring, enableLimit, enableMotor;
};
struct RevoluteJointData
{
	V2 linearImpulse;
	float springImpulse, motorImpulse, lowerImpulse, upperImpulse;
	float hertz, dampingRatio, targetAngle, maxMotorTorque, motorSpeed, referenceAngle, lowerAngle, upperAngle;
	int32_t indexA, indexB;
	V2 anchorA, anchorB, deltaCenter;
	float deltaAngle, axialMass;
	Soft springSoftness;
	bool enableSpring, enableMotor, enableLimit;
};
struct WeldJointData
{
	float referenceAngle, linearHertz, linearDampingRatio, angularHertz, angularDampingRatio;
	Soft linearSoftness, angularSoftness;
	V2 linearImpulse;
	float angularImpulse;
	int32_t indexA, indexB;
	V2 anchorA, anchorB, deltaCenter;
	float deltaAngle, axialMass;
};
struct WheelJointData
{
	V2 localAxisA;
	float perpImpulse, motorImpulse, springImpulse, lowerImpulse, upperImpulse;
	float maxMotorTorque, motorSpeed, lowerTranslation, upperTranslation, hertz, dampingRatio;
	int32_t indexA, indexB;
	V2 anchorA, anchorB, axisA, deltaCenter;
	float perpMass, motorMass, axialMass;
	Soft springSoftness;
	bool enableSpring, enableMotor, enableLimit;
};
// Joint simulation data (stable slot = joint id): B2/src/joint.h:247-278
struct JointSim
{
	int32_t jointId, bodyIdA, bodyIdB, type;
	V2 localOriginAnchorA, localOriginAnchorB;
	float invMassA, invMassB, invIA, invIB;
	float constraintHertz, constraintDampingRatio;
	Soft constraintSoftness;
	union
	{
		DistanceJointData distance;
		MotorJointData motor;
		MouseJointData mouse;
		RevoluteJointData revolute;
		PrismaticJointData prismatic;
		WeldJointData weld;
		WheelJointData wheel;
	};
};

// Broadphase BVH node, 48 B = three 16-byte sectors (box | links | category): B2/src/dynamic_tree.c:19-50
enum : uint16_t
{
	kNodeAllocated = 0x1,
	kNodeEnlarged = 0x2,
	kNodeLeaf = 0x4,
	kNodeMoved = 0x8 // replaces the reference's moveSet hash (broad_phase.h:74-82): set while the proxy is in moveArray
};
struct alignas( 16 ) TreeNode
{
	Box box;
	union
	{
		struct
		{
			int32_t child1, child2;
		};
		uint64_t userData; // leaf: shape id
	};
	int32_t parent; // or free-list next
	uint16_t height, flags;
	uint64_t category;
	uint64_t pad;
};
struct Tree
{
	Arr<TreeNode> nodes; // nodes.count = high-water mark of ever-used node slots
	Arr<int32_t> leafIndices;
	Arr<V2> leafCenters;
	Arr<int32_t> work; // rebuild stacks (kept in the image: a 1024-deep stack per GPU thread would cost GBs of local memory)
	int32_t root, nodeCount, freeList, proxyCount;
};
struct MovePair
{
	int32_t shapeA, shapeB, next;
};

// Sleeping solver set (id >= 3): four id lists carved from one block of World::sleepPool.
// Reference: B2/src/solver_set.h:19-43 with b2TrySleepIsland / b2WakeSolverSet (solver_set.c:37-422).
struct SolverSet
{
	int32_t setIndex; // kNull when the slot is free
	int32_t blockOff; // start in World::sleepPool; layout [bodies | contacts | joints | islands], capacities below
	int32_t bodyCap, contactCap, jointCap, islandCap;
	int32_t bodyCount, contactCount, jointCount, islandCount;
	int32_t prevBlock, nextBlock; // address-ordered block list (set ids) for compaction
};

// Step constants: B2/src/solver.h:75-138, built in world.c:742-771
constexpr int kPairClassCount = 32; // kShapeTypeCount^2 rounded up
struct StepCtx
{
	float dt, inv_dt, h, inv_h;
	int32_t classCount[kPairClassCount], classFill[kPairClassCount]; // narrowphase work list by pair class (stepCollide)
	int32_t subStepCount;
	Soft contactSoftness, staticSoftness;
	float restitutionThreshold, maxLinearVelocity;
	int32_t enableWarmStarting;
	int32_t awakeBodyCount, awakeContactCount, awakeJointCount;
	int32_t colorBase[kColorCount + 1]; // constraint slot base per colour
	int32_t activeColorCount;
	int32_t activeColors[kColorCount];
	int32_t bulletCount;
	int32_t hitCandidateCount;
	int32_t moveCount;
	int32_t pairCount;
	int32_t mergeCount; // awake islands with a parent at the start of the solve
	int32_t splitTarget; // island to split this step (copy of World::splitIslandId taken before the solve), or kNull
	int32_t splitBodies, splitContacts, splitJoints, splitComponents; // sizes of the split work arrays
	int32_t islandPath; // 1: constraints are solved island by island (one warp each), 0: colour by colour
	int32_t maxIslandContacts, maxIslandBodies;
	int32_t islandSolveCount; // awake islands when the island-parallel partition was built (a split may add islands later)
	int32_t orderedPairCount; // callback-mediated step: candidate pairs in creation order, waiting for the host's custom filter
	int32_t preSolveCount;	  // callback-mediated step: touching contacts waiting for the host's pre-solve verdict
	int32_t fastDeferredCount; // callback-mediated step: fast bodies whose continuous pass waits for the host's custom filter
	int32_t stateNeedsSerial;  // contact state pass: a begin-touch would wake a sleeping set, so the pass runs on one thread
	int32_t bodiesFinalized;   // stepSolve ran the body loop of finalize beside the island split; stepFinalize only casts the votes
	int32_t retryContacts;	   // != 0: the step stopped before its first structural edit because the contact arrays cannot take
							   // this many new contacts (kErrRetry); the host grows the image and runs the step again
	unsigned long long splitKey; // (sleepTime bits << 32) | ~simIndex  — arg-max over bodies that want an island split
};

// Contact constraint storage: structure-of-arrays, field-major, one slot per touching contact, slots of one colour
// contiguous (slot = colorBase[c] + index in colour). Field meaning follows b2ContactConstraintSIMD
// (B2/src/contact_solver.c:1034-1064), one lane per slot.
enum ConField : int
{
	cfInvMassA,
	cfInvMassB,
	cfInvIA,
	cfInvIB,
	cfNormalX,
	cfNormalY,
	cfFriction,
	cfTangentSpeed,
	cfRollingResistance,
	cfRollingMass,
	cfRollingImpulse,
	cfBiasRate,
	cfMassScale,
	cfImpulseScale,
	cfAnchorA1X,
	cfAnchorA1Y,
	cfAnchorB1X,
	cfAnchorB1Y,
	cfNormalMass1,
	cfTangentMass1,
	cfBaseSeparation1,
	cfNormalImpulse1,
	cfTotalNormalImpulse1,
	cfTangentImpulse1,
	cfAnchorA2X,
	cfAnchorA2Y,
	cfAnchorB2X,
	cfAnchorB2Y,
	cfBaseSeparation2,
	cfNormalImpulse2,
	cfTotalNormalImpulse2,
	cfTangentImpulse2,
	cfNormalMass2,
	cfTangentMass2,
	cfRestitution,
	cfRelativeVelocity1,
	cfRelativeVelocity2,
	cfIndexA, // int bits
	cfIndexB, // int bits
	cfPointCount, // int bits (overflow path only)
	cfFieldCount
};

enum ProfSlot : int
{
	pfBegin,
	pfPairQuery,
	pfPairCreate,
	pfTreeRebuild,
	pfNarrow,
	pfStatePass,
	pfSolveSetup,
	pfPrepare,
	pfIntegrateVel,
	pfWarmStart,
	pfSolve,
	pfIntegratePos,
	pfRelax,
	pfRestitution,
	pfStore,
	pfFinalizeBodies,
	pfHitEvents,
	pfEnlarge,
	pfBullets,
	pfSleep,
	pfEnd,
	pfSplitJoin,  // waiting for the island-split walk after the solver stages (accumulated separately from the phases above)
	pfSplitApply,
	pfTreeBeside, // the tree rebuild when it runs beside the contact-state pass (its time is inside pfStatePass)
	// inside the tree rebuild (charged by the leader of the team that rebuilds; their sum is pfTreeRebuild or pfTreeBeside)
	pfTreeCollect,
	pfTreePositions,
	pfTreeLevels,
	pfTreeTail,
	pfTreeRefit,
	kProfSlots = 32
};

enum : uint8_t // World::hostCallbacks
{
	kHostCustomFilter = 1, // b2CustomFilterFcn registered (broad_phase.c:267-278)
	kHostPreSolve = 2,	   // b2PreSolveFcn registered (contact.c:504-517)
	kHostMixing = 4		   // friction / restitution mixing callbacks in the world def: not supported
};
enum : uint32_t
{
	kErrCapacity = 1,		 // a fixed-capacity array overflowed inside the step
	kErrTreeStack = 2,		 // traversal stack overflow
	kErrUnsupported = 4,	 // feature not available on the device path
	kErrSleepPool = 8,
	kErrRetry = 16, // not an error for a single world (b2World_Step grows the image and repeats the step); a batch reports it
	// Per-step flag, cleared when the next step begins, never stops the stepping: an event or a sensor overlap did not fit
	// its array and was dropped (the simulation itself is unaffected)
	kErrTruncated = 32
};
constexpr uint32_t kErrFatal = kErrCapacity | kErrUnsupported; // a world with one of these set is not stepped any more

// World parameters + every array. Reference: B2/src/world.h:44-175.
struct World
{
	uint64_t magic;
	uint64_t imageBytes;
	uint32_t error;
	int32_t errorDetail;

	V2 gravity;
	float hitEventThreshold, restitutionThreshold, maxLinearSpeed, maxContactPushSpeed, contactSpeed, contactHertz,
		contactDampingRatio;
	float inv_h;
	uint64_t stepIndex;
	int32_t splitIslandId, endEventArrayIndex;
	uint16_t worldId, generation;
	bool enableSleep, locked, enableWarmStarting, enableContinuous, enableSpeculative, inUse;
	uint8_t hostCallbacks; // kHostCustomFilter | kHostPreSolve (| kHostMixing: unsupported): the step runs callback-mediated
	int32_t taskCount;
	int32_t sensorOverlapCap;	 // capacity of one overlap list of one sensor (World::sensorRefs)
	int32_t hitEventCapable;	 // shapes with enableHitEvents (hit-event scan is skipped when 0)
	int32_t contactEventCapable; // shapes with enableContactEvents (sizes the begin/end event arrays)

	StepCtx step;

	// in-kernel phase profile: rank 0 accumulates nanoseconds between marks (f2d_step.h F2D_MARK); off unless enabled
	uint64_t prof[kProfSlots];
	uint64_t profLast;
	int32_t profEnabled;
	uint32_t shapeTypeMask; // bit per shape type ever created in this world: one bit = one pair class, nothing to bin
	int32_t pairClassBinningOff, pad3[3]; // measurement aid (f2dWorld_EnablePairClassBinning): 1 = never bin

	// entities (slot = id)
	IdPool bodyIds, shapeIds, contactIds, jointIds, islandIds, setIds, chainIds;
	Arr<Body> bodies;
	Arr<BodySim> sims;
	Arr<Shape> shapes;
	Arr<Contact> contacts;
	Arr<ContactSim> contactSims;
	Arr<Joint> joints;
	Arr<JointSim> jointSims;
	Arr<Island> islands;
	Arr<SolverSet> sets;
	Arr<Sensor> sensors;	   // dense, swap-removed (shape.c:283-305); Shape::sensorIndex points here
	Arr<ShapeRef> sensorRefs;  // 2 * kSensorOverlapCap per sensor slot
	Arr<uint64_t> sensorBits;  // sensors whose overlap set changed this step (sensor.c:250-262)

	// set membership lists (ids)
	Arr<int32_t> staticBodies, disabledBodies, awakeBodies;
	Arr<BodyState> states; // aligned with awakeBodies
	Arr<int32_t> awakeContacts;	   // awake, non-touching contacts
	Arr<int32_t> disabledContacts; // non-touching contacts between sleeping bodies
	Arr<int32_t> disabledJoints;
	Arr<int32_t> staticJoints; // joints between two static bodies (joint.c:229-239)
	Arr<int32_t> awakeIslands;
	Arr<int32_t> sleepPool;
	int32_t sleepHead, sleepTail, sleepUsed, pad0;

	// constraint graph
	Arr<int32_t> colorContacts[kColorCount];
	Arr<int32_t> colorJoints[kColorCount];

	// broadphase
	Tree trees[3];
	Arr<int32_t> treeScratch; // team-parallel rebuild work arrays (f2d_tree_team.h)
	Arr<int32_t> moveArray;
	Arr<int32_t> moveHeads;
	Arr<MovePair> movePairs;
	Arr<int32_t> pairOffsets; // per moved proxy: start of its pairs in pairOrder
	Arr<int32_t> pairOrder;	  // pair indices in creation order

	// events
	Arr<BodyMoveEvent> moveEvents;
	Arr<BeginTouchEvent> beginEvents;
	Arr<EndTouchEvent> endEvents[2];
	Arr<HitEvent> hitEvents;
	Arr<SensorEvent> sensorBeginEvents;
	Arr<SensorEvent> sensorEndEvents[2];

	// per-step scratch
	Arr<uint64_t> contactBits;	// contact state changes by contact id (world.c:548-553)
	Arr<int32_t> stateOffsets;	// per word of contactBits: prefix sum of set bits
	Arr<int32_t> stateList;		// flagged contact ids, ascending
	Arr<uint64_t> enlargedBits; // by awake index (solver.c:1835-1840)
	Arr<uint64_t> islandBits;	// awake islands kept awake (solver.c:2024-2028)
	Arr<float> cons;			// cfFieldCount x consStride
	int32_t consStride, pad1;
	// island-parallel solve (f2d_step.h islandSolve): per awake island, its constraint slots grouped by colour and its
	// awake body indices
	Arr<int32_t> islSlotOff;  // [awake islands + 1]
	Arr<int32_t> islBodyOff;  // [awake islands + 1]
	Arr<int32_t> islColorOff; // [awake islands * 12 + 1] start of (island, colour) in islSlots
	Arr<int32_t> islColorFill;
	Arr<int32_t> islBodyFill;
	Arr<int32_t> islSlots;	  // [awake contacts]
	Arr<int32_t> islBodies;	  // [awake bodies]
	Arr<int32_t> bullets;
	Arr<float> integ;	  // 6 x bodies.cap, field-major by awake index: this step's velocity-integration constants
	Arr<int32_t> scan;	  // scan scratch (awake bodies + 1)
	Arr<int32_t> scratch; // island split stacks
	Arr<int32_t> splitScratch; // island split work arrays (f2d_step.h SplitView)

	// Device only: address of the image in HBM. The one-block-per-world kernels work on a copy of this header in
	// SHARED memory (every count, offset and step constant is then a shared-memory load instead of a global one), so
	// array addresses cannot be derived from the header's own address there; the kernels set this at entry. Host code
	// never reads it (ptr() below uses the header address itself on the host).
	uint64_t deviceBase;
};
static_assert( sizeof( World ) % 16 == 0, "the header is copied to shared memory in 16-byte words" );

constexpr uint64_t kWorldMagic = 0x4632444232303042ull; // "F2DB200B"

template <class T> F2D_HD T* ptr( World* w, const Arr<T>& a )
{
#if defined( __CUDA_ARCH__ )
	// every array of the image lives in HBM: telling the compiler so turns the generic LD / ST into LDG / STG
	T* p = reinterpret_cast<T*>( w->deviceBase + a.off );
	__builtin_assume( __isGlobal( p ) );
	return p;
#else
	return reinterpret_cast<T*>( reinterpret_cast<char*>( w ) + a.off );
#endif
}
template <class T> F2D_HD const T* ptr( const World* w, const Arr<T>& a )
{
#if defined( __CUDA_ARCH__ )
	const T* p = reinterpret_cast<const T*>( w->deviceBase + a.off );
	__builtin_assume( __isGlobal( p ) );
	return p;
#else
	return reinterpret_cast<const T*>( reinterpret_cast<const char*>( w ) + a.off );
#endif
}

F2D_HD void setError( World* w, uint32_t bits, int detail )
{
	if ( ( w->error & bits ) == 0 )
	{
		w->errorDetail = detail;
	}
	w->error |= bits;
}

// push with capacity guard: on overflow flags the world and returns slot 0 so the caller stays in bounds
template <class T> F2D_HD int push( World* w, Arr<T>& a, const T& v, int line )
{
	if ( a.count >= a.cap )
	{
		setError( w, kErrCapacity, line );
		return a.cap > 0 ? a.cap - 1 : 0;
	}
	ptr( w, a )[a.count] = v;
	return a.count++;
}
#define F2D_PUSH( w, arr, v ) ::f2d::push( w, arr, v, __LINE__ )
// an event record: when the array is full the event is dropped and the world flagged kErrTruncated (per-step, not fatal)
template <class T> F2D_HD void pushEvent( World* w, Arr<T>& a, const T& v, int line )
{
	if ( a.count >= a.cap )
	{
		setError( w, kErrTruncated, line );
		return;
	}
	ptr( w, a )[a.count] = v;
	a.count += 1;
}
#define F2D_PUSH_EVENT( w, arr, v ) ::f2d::pushEvent( w, arr, v, __LINE__ )

// swap-remove with the reference's semantics (B2/src/array.h:91-103): returns the index that moved or kNull
template <class T> F2D_HD int removeSwap( World* w, Arr<T>& a, int index )
{
	int moved = kNull;
	T* d = ptr( w, a );
	if ( index != a.count - 1 )
	{
		moved = a.count - 1;
		d[index] = d[moved];
	}
	a.count -= 1;
	return moved;
}

F2D_HD int allocId( World* w, IdPool& p )
{
	if ( p.free.count > 0 )
	{
		p.free.count -= 1;
		return ptr( w, p.free )[p.free.count];
	}
	return p.next++;
}
F2D_HD void freeId( World* w, IdPool& p, int id )
{
	F2D_PUSH( w, p.free, id );
}
F2D_HD int idCount( const IdPool& p ) { return p.next - p.free.count; }

F2D_HD int proxyType( int key ) { return key & 3; }
F2D_HD int proxyId( int key ) { return key >> 2; }
F2D_HD int proxyKey( int id, int type ) { return ( id << 2 ) | type; }

} // namespace f2d
