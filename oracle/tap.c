// TEST INFRASTRUCTURE — not part of the product. Compiled INTO oracle/_ref/libbox2d_ref.so together with the
// unmodified reference sources (see oracle/Makefile). It includes the reference's *internal* headers from where
// they lie under /root/reference (never copied) and dumps internal state into the records declared in
// include/forge2d_b200_debug.h so that tests can compare the CUDA engine against the real Box2D v3.1.1, bit for bit.
//
// Also provides a persistent pthread task system (b2WorldDef.enqueueTask/finishTask contract,
// B2/include/box2d/types.h:17-50,120-133) so the CPU baseline can use every host core.
#define _GNU_SOURCE
#include "body.h"
#include "broad_phase.h"
#include "constraint_graph.h"
#include "contact.h"
#include "island.h"
#include "joint.h"
#include "shape.h"
#include "solver_set.h"
#include "world.h"

#include "box2d/box2d.h"

#include "forge2d_b200_debug.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define TAP_API __attribute__( ( visibility( "default" ) ) )

// Layout of the tree node private to B2/src/dynamic_tree.c:19-50 (40 bytes).
typedef struct tapTreeNode
{
	b2AABB aabb;
	uint64_t categoryBits;
	union
	{
		struct
		{
			int32_t child1, child2;
		} children;
		uint64_t userData;
	};
	int32_t parent;
	uint16_t height;
	uint16_t flags;
} tapTreeNode;

static b2World* tapWorld( b2WorldId id )
{
	return b2GetWorldFromId( id );
}

TAP_API int tap_awake_order( b2WorldId id, int* bodyIds, int cap )
{
	b2World* w = tapWorld( id );
	b2SolverSet* awake = w->solverSets.data + b2_awakeSet;
	int n = awake->bodySims.count;
	for ( int i = 0; i < n && i < cap; ++i )
		bodyIds[i] = awake->bodySims.data[i].bodyId;
	return n;
}

TAP_API int tap_move_array( b2WorldId id, int* keys, int cap )
{
	b2World* w = tapWorld( id );
	int n = w->broadPhase.moveArray.count;
	for ( int i = 0; i < n && i < cap; ++i )
		keys[i] = w->broadPhase.moveArray.data[i];
	return n;
}

TAP_API int tap_bodies( b2WorldId id, f2dBodyRecord* out, int cap )
{
	b2World* w = tapWorld( id );
	int n = 0;
	for ( int i = 0; i < w->bodies.count; ++i )
	{
		b2Body* b = w->bodies.data + i;
		if ( b->setIndex == B2_NULL_INDEX )
			continue;
		if ( b->id != i )
			continue;
		if ( n < cap )
		{
			f2dBodyRecord* r = out + n;
			memset( r, 0, sizeof( *r ) );
			b2BodySim* sim = b2GetBodySim( w, b );
			r->id = i;
			r->setIndex = b->setIndex;
			r->localIndex = b->localIndex;
			r->islandId = b->islandId;
			r->islandPrev = b->islandPrev;
			r->islandNext = b->islandNext;
			r->type = b->type;
			r->headContactKey = b->headContactKey;
			r->contactCount = b->contactCount;
			r->headShapeId = b->headShapeId;
			r->flags = ( sim->isFast ? 1 : 0 ) | ( sim->isBullet ? 2 : 0 ) | ( b->isSpeedCapped ? 4 : 0 ) | ( sim->enlargeAABB ? 8 : 0 );
			r->px = sim->transform.p.x;
			r->py = sim->transform.p.y;
			r->qc = sim->transform.q.c;
			r->qs = sim->transform.q.s;
			r->cx = sim->center.x;
			r->cy = sim->center.y;
			r->c0x = sim->center0.x;
			r->c0y = sim->center0.y;
			r->q0c = sim->rotation0.c;
			r->q0s = sim->rotation0.s;
			if ( b->setIndex == b2_awakeSet )
			{
				b2BodyState* s = w->solverSets.data[b2_awakeSet].bodyStates.data + b->localIndex;
				r->vx = s->linearVelocity.x;
				r->vy = s->linearVelocity.y;
				r->w = s->angularVelocity;
			}
			r->sleepTime = b->sleepTime;
			r->invMass = sim->invMass;
			r->invInertia = sim->invInertia;
			r->minExtent = sim->minExtent;
			r->maxExtent = sim->maxExtent;
			r->lcx = sim->localCenter.x;
			r->lcy = sim->localCenter.y;
		}
		n += 1;
	}
	return n;
}

TAP_API int tap_contacts( b2WorldId id, f2dContactRecord* out, int cap )
{
	b2World* w = tapWorld( id );
	int n = 0;
	for ( int i = 0; i < w->contacts.count; ++i )
	{
		b2Contact* c = w->contacts.data + i;
		if ( c->contactId != i || c->setIndex == B2_NULL_INDEX )
			continue;
		if ( n < cap )
		{
			f2dContactRecord* r = out + n;
			memset( r, 0, sizeof( *r ) );
			b2ContactSim* s = b2GetContactSim( w, c );
			r->id = i;
			r->shapeIdA = c->shapeIdA;
			r->shapeIdB = c->shapeIdB;
			r->setIndex = c->setIndex;
			r->colorIndex = c->colorIndex;
			r->localIndex = c->localIndex;
			r->flags = (int)c->flags;
			r->simFlags = (int)s->simFlags;
			r->pointCount = s->manifold.pointCount;
			r->islandId = c->islandId;
			r->islandPrev = c->islandPrev;
			r->islandNext = c->islandNext;
			r->prevKeyA = c->edges[0].prevKey;
			r->nextKeyA = c->edges[0].nextKey;
			r->prevKeyB = c->edges[1].prevKey;
			r->nextKeyB = c->edges[1].nextKey;
			r->bodySimIndexA = s->bodySimIndexA;
			r->bodySimIndexB = s->bodySimIndexB;
			r->nx = s->manifold.normal.x;
			r->ny = s->manifold.normal.y;
			for ( int k = 0; k < s->manifold.pointCount && k < 2; ++k )
			{
				b2ManifoldPoint* mp = s->manifold.points + k;
				if ( k == 0 )
					r->id0 = mp->id;
				else
					r->id1 = mp->id;
				r->sep[k] = mp->separation;
				r->ni[k] = mp->normalImpulse;
				r->ti[k] = mp->tangentImpulse;
				r->tni[k] = mp->totalNormalImpulse;
				r->nv[k] = mp->normalVelocity;
				r->ax[k] = mp->anchorA.x;
				r->ay[k] = mp->anchorA.y;
				r->bx[k] = mp->anchorB.x;
				r->by[k] = mp->anchorB.y;
				r->px[k] = mp->point.x;
				r->py[k] = mp->point.y;
			}
			r->friction = s->friction;
			r->restitution = s->restitution;
			r->rollingImpulse = s->manifold.rollingImpulse;
		}
		n += 1;
	}
	return n;
}

TAP_API int tap_islands( b2WorldId id, f2dIslandRecord* out, int cap )
{
	b2World* w = tapWorld( id );
	int n = 0;
	for ( int i = 0; i < w->islands.count; ++i )
	{
		b2Island* s = w->islands.data + i;
		if ( s->islandId != i || s->setIndex == B2_NULL_INDEX )
			continue;
		if ( n < cap )
		{
			f2dIslandRecord* r = out + n;
			r->id = i;
			r->setIndex = s->setIndex;
			r->localIndex = s->localIndex;
			r->headBody = s->headBody;
			r->tailBody = s->tailBody;
			r->bodyCount = s->bodyCount;
			r->headContact = s->headContact;
			r->tailContact = s->tailContact;
			r->contactCount = s->contactCount;
			r->headJoint = s->headJoint;
			r->tailJoint = s->tailJoint;
			r->jointCount = s->jointCount;
			r->parentIsland = s->parentIsland;
			r->constraintRemoveCount = s->constraintRemoveCount;
		}
		n += 1;
	}
	return n;
}

TAP_API int tap_shapes( b2WorldId id, f2dShapeRecord* out, int cap )
{
	b2World* w = tapWorld( id );
	int n = 0;
	for ( int i = 0; i < w->shapes.count; ++i )
	{
		b2Shape* s = w->shapes.data + i;
		if ( s->id != i )
			continue;
		if ( n < cap )
		{
			f2dShapeRecord* r = out + n;
			r->id = i;
			r->bodyId = s->bodyId;
			r->proxyKey = s->proxyKey;
			r->type = s->type;
			r->enlarged = s->enlargedAABB;
			r->aabb[0] = s->aabb.lowerBound.x;
			r->aabb[1] = s->aabb.lowerBound.y;
			r->aabb[2] = s->aabb.upperBound.x;
			r->aabb[3] = s->aabb.upperBound.y;
			r->fat[0] = s->fatAABB.lowerBound.x;
			r->fat[1] = s->fatAABB.lowerBound.y;
			r->fat[2] = s->fatAABB.upperBound.x;
			r->fat[3] = s->fatAABB.upperBound.y;
		}
		n += 1;
	}
	return n;
}

TAP_API int tap_tree( b2WorldId id, int treeType, f2dTreeLeafRecord* out, int cap )
{
	b2World* w = tapWorld( id );
	b2DynamicTree* tree = w->broadPhase.trees + treeType;
	if ( tree->root == B2_NULL_INDEX || tree->nodeCount == 0 )
		return 0;
	const tapTreeNode* nodes = (const tapTreeNode*)tree->nodes;
	int n = 0;
	int stack[2048], depth[2048], enl[2048];
	int sp = 0;
	stack[sp] = tree->root;
	depth[sp] = 0;
	enl[sp] = 0;
	sp += 1;
	while ( sp > 0 )
	{
		sp -= 1;
		int ni = stack[sp], d = depth[sp], e = enl[sp];
		const tapTreeNode* node = nodes + ni;
		if ( node->flags & b2_leafNode )
		{
			if ( n < cap )
			{
				out[n].proxyId = ni;
				out[n].depth = d;
				out[n].enlargedAncestors = e;
				out[n].box[0] = node->aabb.lowerBound.x;
				out[n].box[1] = node->aabb.lowerBound.y;
				out[n].box[2] = node->aabb.upperBound.x;
				out[n].box[3] = node->aabb.upperBound.y;
			}
			n += 1;
		}
		else if ( sp < 2046 )
		{
			int e2 = e + ( ( node->flags & b2_enlargedNode ) ? 1 : 0 );
			// child1-first order
			stack[sp] = node->children.child2;
			depth[sp] = d + 1;
			enl[sp] = e2;
			sp += 1;
			stack[sp] = node->children.child1;
			depth[sp] = d + 1;
			enl[sp] = e2;
			sp += 1;
		}
	}
	return n;
}

TAP_API int tap_joints( b2WorldId id, f2dJointRecord* out, int cap )
{
	b2World* w = tapWorld( id );
	int n = 0;
	for ( int i = 0; i < w->joints.count; ++i )
	{
		b2Joint* j = w->joints.data + i;
		if ( j->jointId != i || j->setIndex == B2_NULL_INDEX )
			continue;
		if ( n < cap )
		{
			f2dJointRecord* r = out + n;
			memset( r, 0, sizeof( *r ) );
			b2JointSim* s = b2GetJointSim( w, j );
			r->id = i;
			r->type = j->type;
			r->setIndex = j->setIndex;
			r->colorIndex = j->colorIndex;
			r->localIndex = j->localIndex;
			r->bodyIdA = j->edges[0].bodyId;
			r->bodyIdB = j->edges[1].bodyId;
			r->islandId = j->islandId;
			switch ( j->type )
			{
				case b2_revoluteJoint:
					r->impulse[0] = s->revoluteJoint.linearImpulse.x;
					r->impulse[1] = s->revoluteJoint.linearImpulse.y;
					r->impulse[2] = s->revoluteJoint.springImpulse;
					r->impulse[3] = s->revoluteJoint.motorImpulse;
					r->impulse[4] = s->revoluteJoint.lowerImpulse;
					r->impulse[5] = s->revoluteJoint.upperImpulse;
					break;
				case b2_distanceJoint:
					r->impulse[0] = s->distanceJoint.impulse;
					r->impulse[1] = s->distanceJoint.lowerImpulse;
					r->impulse[2] = s->distanceJoint.upperImpulse;
					r->impulse[3] = s->distanceJoint.motorImpulse;
					break;
				case b2_motorJoint:
					r->impulse[0] = s->motorJoint.linearImpulse.x;
					r->impulse[1] = s->motorJoint.linearImpulse.y;
					r->impulse[2] = s->motorJoint.angularImpulse;
					break;
				case b2_mouseJoint:
					r->impulse[0] = s->mouseJoint.linearImpulse.x;
					r->impulse[1] = s->mouseJoint.linearImpulse.y;
					r->impulse[2] = s->mouseJoint.angularImpulse;
					break;
				case b2_prismaticJoint:
					r->impulse[0] = s->prismaticJoint.impulse.x;
					r->impulse[1] = s->prismaticJoint.impulse.y;
					r->impulse[2] = s->prismaticJoint.springImpulse;
					r->impulse[3] = s->prismaticJoint.motorImpulse;
					r->impulse[4] = s->prismaticJoint.lowerImpulse;
					r->impulse[5] = s->prismaticJoint.upperImpulse;
					break;
				case b2_weldJoint:
					r->impulse[0] = s->weldJoint.linearImpulse.x;
					r->impulse[1] = s->weldJoint.linearImpulse.y;
					r->impulse[2] = s->weldJoint.angularImpulse;
					break;
				case b2_wheelJoint:
					r->impulse[0] = s->wheelJoint.perpImpulse;
					r->impulse[1] = s->wheelJoint.motorImpulse;
					r->impulse[2] = s->wheelJoint.springImpulse;
					r->impulse[3] = s->wheelJoint.lowerImpulse;
					r->impulse[4] = s->wheelJoint.upperImpulse;
					break;
				default:
					break;
			}
		}
		n += 1;
	}
	return n;
}

TAP_API void tap_color_counts( b2WorldId id, int* contactCounts, int* jointCounts )
{
	b2World* w = tapWorld( id );
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i )
	{
		contactCounts[i] = w->constraintGraph.colors[i].contactSims.count;
		jointCounts[i] = w->constraintGraph.colors[i].jointSims.count;
	}
}

// Contact ids of one colour in array order (the Gauss-Seidel order inside the overflow colour).
TAP_API int tap_color_contacts( b2WorldId id, int colorIndex, int* contactIds, int cap )
{
	b2World* w = tapWorld( id );
	b2GraphColor* c = w->constraintGraph.colors + colorIndex;
	int n = c->contactSims.count;
	for ( int i = 0; i < n && i < cap; ++i )
		contactIds[i] = c->contactSims.data[i].contactId;
	return n;
}

TAP_API int tap_awake_contacts( b2WorldId id, int* contactIds, int cap )
{
	b2World* w = tapWorld( id );
	b2SolverSet* s = w->solverSets.data + b2_awakeSet;
	int n = s->contactSims.count;
	for ( int i = 0; i < n && i < cap; ++i )
		contactIds[i] = s->contactSims.data[i].contactId;
	return n;
}

TAP_API int tap_awake_islands( b2WorldId id, int* islandIds, int cap )
{
	b2World* w = tapWorld( id );
	b2SolverSet* s = w->solverSets.data + b2_awakeSet;
	int n = s->islandSims.count;
	for ( int i = 0; i < n && i < cap; ++i )
		islandIds[i] = s->islandSims.data[i].islandId;
	return n;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent pthread pool implementing Box2D's task interface for the all-cores CPU baseline.
// A task is a parallel-for over [0,itemCount) split into chunks of at least minRange; each chunk runs with a
// unique workerIndex < workerCount, as the contract requires.

#define TAP_MAX_THREADS 256
#define TAP_MAX_TASKS 256

typedef struct tapTask
{
	b2TaskCallback* fcn;
	void* ctx;
	int itemCount, chunk;
	atomic_int next;
	atomic_int done; // completed items
	atomic_int active;
} tapTask;

typedef struct tapPool
{
	pthread_t threads[TAP_MAX_THREADS];
	int threadCount; // worker threads (excluding caller)
	tapTask tasks[TAP_MAX_TASKS];
	atomic_int taskCount;
	atomic_int generation;
	atomic_int quit;
} tapPool;

typedef struct tapThreadArg
{
	tapPool* pool;
	int workerIndex;
} tapThreadArg;

static tapThreadArg tapArgs[TAP_MAX_THREADS];

static int tapRunChunks( tapTask* t, int workerIndex )
{
	int ran = 0;
	for ( ;; )
	{
		int start = atomic_fetch_add( &t->next, t->chunk );
		if ( start >= t->itemCount )
			break;
		int end = start + t->chunk;
		if ( end > t->itemCount )
			end = t->itemCount;
		t->fcn( start, end, (uint32_t)workerIndex, t->ctx );
		atomic_fetch_add( &t->done, end - start );
		ran = 1;
	}
	return ran;
}

static void* tapThreadMain( void* p )
{
	tapThreadArg* a = p;
	tapPool* pool = a->pool;
	int spins = 0;
	while ( atomic_load( &pool->quit ) == 0 )
	{
		int n = atomic_load( &pool->taskCount );
		int ran = 0;
		for ( int i = 0; i < n; ++i )
		{
			tapTask* t = pool->tasks + i;
			if ( atomic_load( &t->active ) && atomic_load( &t->next ) < t->itemCount )
				ran |= tapRunChunks( t, a->workerIndex );
		}
		if ( ran )
			spins = 0;
		else if ( ++spins > 2000 )
		{
			sched_yield();
			spins = 0;
		}
		else
			__builtin_ia32_pause();
	}
	return NULL;
}

TAP_API void* tap_pool_create( int workerCount )
{
	tapPool* pool = calloc( 1, sizeof( tapPool ) );
	pool->threadCount = workerCount - 1;
	if ( pool->threadCount > TAP_MAX_THREADS )
		pool->threadCount = TAP_MAX_THREADS;
	for ( int i = 0; i < pool->threadCount; ++i )
	{
		tapArgs[i].pool = pool;
		tapArgs[i].workerIndex = i + 1;
		pthread_create( pool->threads + i, NULL, tapThreadMain, tapArgs + i );
	}
	return pool;
}

TAP_API void tap_pool_destroy( void* p )
{
	tapPool* pool = p;
	atomic_store( &pool->quit, 1 );
	for ( int i = 0; i < pool->threadCount; ++i )
		pthread_join( pool->threads[i], NULL );
	free( pool );
}

static void* tapEnqueue( b2TaskCallback* task, int itemCount, int minRange, void* taskContext, void* userContext )
{
	tapPool* pool = userContext;
	int workers = pool->threadCount + 1;
	int slot = atomic_fetch_add( &pool->taskCount, 1 );
	if ( slot >= TAP_MAX_TASKS || itemCount <= minRange || workers == 1 )
	{
		if ( slot < TAP_MAX_TASKS )
			atomic_store( &pool->tasks[slot].active, 0 );
		// run inline; Box2D treats a NULL return as "already finished"
		// long single-item tasks (solver workers, tree rebuild) must still go to the pool
		if ( itemCount == 1 && workers > 1 && slot < TAP_MAX_TASKS )
		{
			tapTask* t = pool->tasks + slot;
			t->fcn = task;
			t->ctx = taskContext;
			t->itemCount = 1;
			t->chunk = 1;
			atomic_store( &t->next, 0 );
			atomic_store( &t->done, 0 );
			atomic_store( &t->active, 1 );
			return t;
		}
		task( 0, itemCount, 0, taskContext );
		return NULL;
	}
	tapTask* t = pool->tasks + slot;
	int chunk = ( itemCount + 4 * workers - 1 ) / ( 4 * workers );
	if ( chunk < minRange )
		chunk = minRange;
	t->fcn = task;
	t->ctx = taskContext;
	t->itemCount = itemCount;
	t->chunk = chunk;
	atomic_store( &t->next, 0 );
	atomic_store( &t->done, 0 );
	atomic_store( &t->active, 1 );
	return t;
}

static void tapFinish( void* userTask, void* userContext )
{
	(void)userContext;
	tapTask* t = userTask;
	if ( t == NULL )
		return;
	tapRunChunks( t, 0 );
	while ( atomic_load( &t->done ) < t->itemCount )
		__builtin_ia32_pause();
	atomic_store( &t->active, 0 );
}

// Box2D requires unique worker indices per *concurrently running* chunk of the same task. The solver stage
// (B2/src/solver.c:1691-1698) enqueues workerCount single-item tasks that spin on each other, so every one of
// them must run on its own thread: the pool threads pick them up and the caller runs what is left in finish.
TAP_API void tap_pool_begin_step( void* p )
{
	tapPool* pool = p;
	atomic_store( &pool->taskCount, 0 );
}

TAP_API b2WorldId tap_create_world_mt( const b2WorldDef* def, void* pool, int workerCount )
{
	b2WorldDef d = *def;
	d.workerCount = workerCount;
	d.enqueueTask = tapEnqueue;
	d.finishTask = tapFinish;
	d.userTaskContext = pool;
	return b2CreateWorld( &d );
}

TAP_API void tap_step_mt( b2WorldId id, float dt, int subSteps, void* pool )
{
	tap_pool_begin_step( pool );
	b2World_Step( id, dt, subSteps );
}

// ---------------------------------------------------------------------------------------------------------------
// Batch-of-worlds CPU baseline ("one single-worker world per thread, round-robin"): `threads`
// pthreads, thread t steps worlds t, t+threads, ... `steps` times each. Returns the wall time in seconds.
#include <time.h>
typedef struct tapBatchArg
{
	const b2WorldId* ids;
	int count, first, stride, steps, subSteps;
	float dt;
} tapBatchArg;

static void* tapBatchMain( void* p )
{
	tapBatchArg* a = p;
	for ( int s = 0; s < a->steps; ++s )
		for ( int i = a->first; i < a->count; i += a->stride )
			b2World_Step( a->ids[i], a->dt, a->subSteps );
	return NULL;
}

TAP_API double tap_step_worlds( const b2WorldId* ids, int count, float dt, int subSteps, int steps, int threads )
{
	if ( threads < 1 )
		threads = 1;
	if ( threads > TAP_MAX_THREADS )
		threads = TAP_MAX_THREADS;
	if ( threads > count )
		threads = count;
	pthread_t th[TAP_MAX_THREADS];
	tapBatchArg args[TAP_MAX_THREADS];
	struct timespec t0, t1;
	clock_gettime( CLOCK_MONOTONIC, &t0 );
	for ( int t = 0; t < threads; ++t )
	{
		args[t] = ( tapBatchArg ){ ids, count, t, threads, steps, subSteps, dt };
		if ( t > 0 )
			pthread_create( th + t, NULL, tapBatchMain, args + t );
	}
	tapBatchMain( args );
	for ( int t = 1; t < threads; ++t )
		pthread_join( th[t], NULL );
	clock_gettime( CLOCK_MONOTONIC, &t1 );
	return (double)( t1.tv_sec - t0.tv_sec ) + 1e-9 * (double)( t1.tv_nsec - t0.tv_nsec );
}
