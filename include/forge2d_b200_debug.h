// forge2d_b200 — introspection records shared by the product's debug API (f2dDebug_*) and the
// reference tap (oracle/tap.c, which fills the same records from the compiled reference's internals).
// Parity tests compare these records field by field; integer fields must match bit-exactly.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// One record per allocated body id, ascending id.
// Reference source of each field: B2/src/body.h:14-113 (b2Body), :120-159 (b2BodySim), :66-83 (b2BodyState).
typedef struct f2dBodyRecord
{
	int32_t id, setIndex, localIndex, islandId, islandPrev, islandNext;
	int32_t type, headContactKey, contactCount, headShapeId, flags; // flags: 1=isFast 2=isBullet 4=isSpeedCapped(body) 8=enlargeAABB
	float px, py, qc, qs;     // origin transform
	float cx, cy;             // center of mass (world)
	float c0x, c0y, q0c, q0s; // previous center/rotation (TOI sweep start)
	float vx, vy, w;          // zero unless in the awake set
	float sleepTime;
	float invMass, invInertia, minExtent, maxExtent;
	float lcx, lcy;           // local center
} f2dBodyRecord;

// One record per allocated contact id, ascending id.
// Reference: B2/src/contact.h:41-71 (b2Contact), :98-131 (b2ContactSim), B2/include/box2d/collision.h:498-552.
typedef struct f2dContactRecord
{
	int32_t id, shapeIdA, shapeIdB, setIndex, colorIndex, localIndex;
	int32_t flags, simFlags, pointCount, id0, id1;
	int32_t islandId, islandPrev, islandNext;
	int32_t prevKeyA, nextKeyA, prevKeyB, nextKeyB;
	int32_t bodySimIndexA, bodySimIndexB;
	float nx, ny;
	float sep[2], ni[2], ti[2], tni[2], nv[2];
	float ax[2], ay[2], bx[2], by[2]; // anchors A and B (center-of-mass relative, world frame)
	float px[2], py[2];
	float friction, restitution, rollingImpulse;
} f2dContactRecord;

// One record per allocated island id, ascending id. Reference: B2/src/island.h:25-57.
typedef struct f2dIslandRecord
{
	int32_t id, setIndex, localIndex;
	int32_t headBody, tailBody, bodyCount;
	int32_t headContact, tailContact, contactCount;
	int32_t headJoint, tailJoint, jointCount;
	int32_t parentIsland, constraintRemoveCount;
} f2dIslandRecord;

// Shape AABB record, per allocated shape id. Reference: B2/src/shape.h:13-52.
typedef struct f2dShapeRecord
{
	int32_t id, bodyId, proxyKey, type, enlarged;
	float aabb[4], fat[4];
} f2dShapeRecord;

// One entry per leaf visited by a child1-first DFS of a broadphase tree: (proxyId, depth).
// Two binary trees with equal (leaf, depth) pre-order sequences have identical topology.
typedef struct f2dTreeLeafRecord
{
	int32_t proxyId, depth, enlargedAncestors;
	float box[4];
} f2dTreeLeafRecord;

// One record per allocated joint id. Reference: B2/src/joint.h:22-50 (b2Joint) + b2JointSim :247-278.
typedef struct f2dJointRecord
{
	int32_t id, type, setIndex, colorIndex, localIndex, bodyIdA, bodyIdB, islandId;
	float impulse[6]; // revolute: linearImpulse.x,.y, springImpulse, motorImpulse, lowerImpulse, upperImpulse
} f2dJointRecord;

#ifdef __cplusplus
}
#endif
