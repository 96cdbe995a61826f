// forge2d_b200 — the world step as team-parallel phases.
//
// Phase order and every serial ordering decision follow b2World_Step (B2/src/world.c:695-812):
//   stepBegin      clear per-step event arrays / scratch bits                        world.c:709-712
//   stepPairs      broadphase pair finding + ordered contact creation                broad_phase.c:311-474
//   stepCollide    tree rebuild | narrowphase | ordered contact-state pass           world.c:488-693
//   stepSolve      island merge/split, prepare, substep loop, restitution, store     solver.c:1190-1720, :929-1106
//   stepFinalize   body finalize + continuous, hit events, proxy enlarge, sleep      solver.c:543-726, :1758-2051
// Parallel loops are `rank()/size()` strided; rank 0 executes the reference's serial sections.
#pragma once
#include "f2d_distance.h"
#include "f2d_joint.h"
#include "f2d_solver.h"
#include "f2d_tree_team.h"

namespace f2d
{

// ------------------------------------------------------------------------------------------------ begin
template <class Team> F2D_HDF inline void stepBegin( World* w, Team& t, float dt, int subStepCount )
{
	if ( t.rank() == 0 )
	{
		if ( w->profEnabled )
			w->profLast = profClock();
		w->moveEvents.count = 0;
		w->sensorBeginEvents.count = 0;
		w->beginEvents.count = 0;
		w->hitEvents.count = 0;
		w->taskCount = 0;
		w->locked = true;
		w->error &= ~kErrTruncated; // per-step flag

		// step context: world.c:742-771
		StepCtx& s = w->step;
		s.dt = dt;
		s.subStepCount = maxi( 1, subStepCount );
		if ( dt > 0.0f )
		{
			s.inv_dt = 1.0f / dt;
			s.h = dt / s.subStepCount;
			s.inv_h = s.subStepCount * s.inv_dt;
		}
		else
		{
			s.inv_dt = 0.0f;
			s.h = 0.0f;
			s.inv_h = 0.0f;
		}
		w->inv_h = s.inv_h;
		float contactHertz = minf( w->contactHertz, 0.125f * s.inv_h );
		s.contactSoftness = makeSoft( contactHertz, w->contactDampingRatio, s.h );
		s.staticSoftness = makeSoft( 2.0f * contactHertz, w->contactDampingRatio, s.h );
		w->contactSpeed = w->maxContactPushSpeed / s.staticSoftness.massScale;
		s.restitutionThreshold = w->restitutionThreshold;
		s.maxLinearVelocity = w->maxLinearSpeed;
		s.enableWarmStarting = w->enableWarmStarting ? 1 : 0;
		s.pairCount = 0;
		s.bulletCount = 0;
		s.splitKey = 0;
	}
	// contact-state bits are cleared here so the narrowphase can set them right after the tree rebuild starts
	uint64_t* bits = ptr( w, w->contactBits );
	for ( int i = t.rank(); i < w->contactBits.cap; i += t.size() )
		bits[i] = 0;
	t.sync();
	F2D_MARK( w, t, pfBegin );
}

// ------------------------------------------------------------------------------------------------ pairs
// body.c:1861-1895
F2D_HDF inline bool shouldBodiesCollide( World* w, const Body& bodyA, const Body& bodyB )
{
	if ( bodyA.type != kDynamicBody && bodyB.type != kDynamicBody )
		return false;
	int jointKey, otherBodyId;
	if ( bodyA.jointCount < bodyB.jointCount )
	{
		jointKey = bodyA.headJointKey;
		otherBodyId = bodyB.id;
	}
	else
	{
		jointKey = bodyB.headJointKey;
		otherBodyId = bodyA.id;
	}
	const Joint* joints = ptr( w, w->joints );
	while ( jointKey != kNull )
	{
		int jointId = jointKey >> 1;
		int edgeIndex = jointKey & 1;
		const Joint& joint = joints[jointId];
		if ( joint.collideConnected == false && joint.edges[edgeIndex ^ 1].bodyId == otherBodyId )
			return false;
		jointKey = joint.edges[edgeIndex].nextKey;
	}
	return true;
}

// Reads one word every `stride` bytes of [base, base + bytes): parks those sectors in the L1 of the calling SM. The sum
// is returned so that the loads stay alive.
F2D_HDF inline int warmSectors( const void* base, int bytes, int stride, int rank, int size )
{
	const char* p = static_cast<const char*>( base );
	int acc = 0;
	for ( int i = rank * stride; i < bytes; i += size * stride )
		acc += *reinterpret_cast<const int32_t*>( p + i );
	return acc;
}

// Is there already a contact between these two shapes? Replaces the reference's pairSet hash lookup
// (broad_phase.c:215-220): a contact for the pair exists iff it hangs off both bodies' contact lists.
F2D_HDF inline bool pairExists( World* w, const Body& bodyA, const Body& bodyB, int shapeIdA, int shapeIdB )
{
	const Contact* contacts = ptr( w, w->contacts );
	int key = bodyA.contactCount < bodyB.contactCount ? bodyA.headContactKey : bodyB.headContactKey;
	while ( key != kNull )
	{
		// sector 0 of the record in two 16-byte loads: shape ids, edge 0 | edge 0's next key, edge 1
		const Contact& c = contacts[key >> 1];
		const Q4 ids = load16( &c.shapeIdA ), links = load16( &c.edges[0].nextKey );
		const int a = (int)floatBits( ids.x ), b = (int)floatBits( ids.y );
		if ( ( ( a == shapeIdA ) & ( b == shapeIdB ) ) | ( ( a == shapeIdB ) & ( b == shapeIdA ) ) )
			return true;
		key = (int)floatBits( ( key & 1 ) ? links.w : links.x );
	}
	return false;
}

// One moved proxy: queries in the reference's tree order and filter order (broad_phase.c:160-302, 311-374).
// The reference's query callback does everything per hit. On a GPU that serialises a warp: the 32 lanes walk different
// paths, a lane that hits a leaf runs the callback - four levels of dependent loads (shapes, bodies, a walk over a body's
// contacts) - while the 31 others wait, and then the next lane hits (measured: 47 k cycles per query for ~20 dependent
// loads of its own). So the traversal only COLLECTS the hits that survive the tests it can make on the node it already
// holds (self, the both-moved rule), and the callback body runs afterwards over each lane's short list - the k-th
// candidate of every lane at the same time, so the loads of a level are in flight together. Order per proxy = hit order.
F2D_HDC inline void findPairsForProxy( World* w, int moveIndex )
{
	int32_t* heads = ptr( w, w->moveHeads );
	heads[moveIndex] = kNull;
	int queryKey = ptr( w, w->moveArray )[moveIndex];
	if ( queryKey == kNull )
		return;
	int queryType = proxyType( queryKey );
	int queryProxy = proxyId( queryKey );
	const TreeNode& queryNode = ptr( w, w->trees[queryType].nodes )[queryProxy];
	Box fat = queryNode.box;
	int queryShape = (int)queryNode.userData;
	const Shape* shapes = ptr( w, w->shapes );
	const Body* bodies = ptr( w, w->bodies );
	MovePair* pairs = ptr( w, w->movePairs );

	constexpr int kCandidates = 16; // (a full list is worked off on the spot)
	int32_t candidateKey[kCandidates], candidateShape[kCandidates];
	int candidateCount = 0;
	auto workOff = [&]() {
		for ( int k = 0; k < candidateCount; ++k )
		{
			const int key = candidateKey[k], shapeId = candidateShape[k];
			int shapeIdA, shapeIdB;
			if ( key < queryKey )
			{
				shapeIdA = shapeId;
				shapeIdB = queryShape;
			}
			else
			{
				shapeIdA = queryShape;
				shapeIdB = shapeId;
			}
			// Every test below only drops the pair, so their order is free: the cheap ones come first, evaluated without
			// branches, each round of loads issued together (a test per branch would be a round trip per test).
			const Shape& shapeA = shapes[shapeIdA];
			const Shape& shapeB = shapes[shapeIdB];
			const int bodyIdA = shapeA.bodyId, bodyIdB = shapeB.bodyId;
			const int sensorA = shapeA.sensorIndex, sensorB = shapeB.sensorIndex;
			const Filter filterA = shapeA.filter, filterB = shapeB.filter;
			const Body& bodyA = bodies[bodyIdA];
			const Body& bodyB = bodies[bodyIdB];
			const int typeA = bodyA.type, typeB = bodyB.type;
			const bool sameGroup = ( filterA.group == filterB.group ) & ( filterA.group != 0 );
			const bool masksMeet = ( ( filterA.mask & filterB.category ) != 0 ) & ( ( filterA.category & filterB.mask ) != 0 );
			const bool shapesCollide = sameGroup ? filterA.group > 0 : masksMeet; // shape.h:122-130
			const bool drop = ( bodyIdA == bodyIdB ) | ( sensorA != kNull ) | ( sensorB != kNull ) | ( shapesCollide == false ) |
							  ( ( typeA != kDynamicBody ) & ( typeB != kDynamicBody ) );
			if ( drop )
				continue;
			if ( pairExists( w, bodyA, bodyB, shapeIdA, shapeIdB ) )
				continue;
			if ( shouldBodiesCollide( w, bodyA, bodyB ) == false )
				continue;
			int pairIndex = atomAdd( &w->step.pairCount, 1 );
			if ( pairIndex >= w->movePairs.cap )
				continue; // counted, not stored: stepPairs sees pairCount > cap and asks the host for room (retryContacts)
			MovePair& p = pairs[pairIndex];
			p.shapeA = shapeIdA;
			p.shapeB = shapeIdB;
			p.next = heads[moveIndex]; // push-front: creation order is the reverse of hit order
			heads[moveIndex] = pairIndex;
		}
		candidateCount = 0;
	};

	for ( int pass = 0; pass < 3; ++pass )
	{
		int treeType;
		if ( queryType == kDynamicBody )
			treeType = pass == 0 ? kKinematicBody : ( pass == 1 ? kStaticBody : kDynamicBody );
		else
		{
			if ( pass != 2 )
				continue;
			treeType = kDynamicBody;
		}
		const Tree& tree = w->trees[treeType];
		treeQueryFlags( w, tree, fat, UINT64_MAX, [&]( int proxy, uint64_t userData, uint32_t flags ) -> bool {
			int key = proxyKey( proxy, treeType );
			if ( key == queryKey )
				return true;
			// both proxies moved: the pair belongs to exactly one of the two queries
			if ( queryType == kDynamicBody )
			{
				if ( treeType == kDynamicBody && key < queryKey && ( flags & kNodeMoved ) != 0 )
					return true;
			}
			else if ( flags & kNodeMoved )
			{
				return true;
			}
			if ( candidateCount == kCandidates )
				workOff();
			candidateKey[candidateCount] = key;
			candidateShape[candidateCount] = (int)userData;
			candidateCount += 1;
			return true;
		} );
	}
	workOff();
}

// The serial half of pair finding: contacts for the candidate pairs, in the reference's creation order (World::pairOrder
// over World::movePairs, built by stepPairs). Rank 0 creates; in a grid the other threads of its block warm its L1.
template <class Team> F2D_HDF inline void createOrderedPairs( World* w, Team& t, int total )
{
	const MovePair* pairs = ptr( w, w->movePairs );
	const int32_t* ordered = ptr( w, w->pairOrder );
	if constexpr ( Team::kHasSoloBlock )
	{
		// a grid: the other threads of the serial thread's block read what contact creation is about to chase, so
		// that it runs on L1 hits (see contactStatePass)
		if ( total > 0 && t.inSoloBlock() && t.rank() >= 32 )
		{
			const Shape* shapes = ptr( w, w->shapes );
			const Body* bodies = ptr( w, w->bodies );
			const Contact* contacts = ptr( w, w->contacts );
			int acc = 0;
			for ( int k = t.rank() - 32; k < total && k < 2048; k += t.soloSize() - 32 )
			{
				const MovePair& pair = pairs[ordered[k]];
				if ( pair.shapeA == kNull )
					continue;
				const Shape& a = shapes[pair.shapeA];
				const Shape& b = shapes[pair.shapeB];
				acc += a.bodyId + b.bodyId + a.type + b.type + (int)floatBits( a.restitution ) + (int)floatBits( b.restitution );
				const Body& ba = bodies[a.bodyId];
				const Body& bb = bodies[b.bodyId];
				acc += ba.setIndex + bb.setIndex;
				if ( ba.headContactKey != kNull )
					acc += contacts[ba.headContactKey >> 1].edges[ba.headContactKey & 1].prevKey;
				if ( bb.headContactKey != kNull )
					acc += contacts[bb.headContactKey >> 1].edges[bb.headContactKey & 1].prevKey;
			}
			if ( acc == 0x7fffffff ) // (never: keeps the loads alive)
				storeVolatile( &w->step.orderedPairCount, total );
		}
	}
	if ( total > 0 && t.rank() == 0 )
	{
		for ( int k = 0; k < total; ++k )
		{
			const MovePair& pair = pairs[ordered[k]];
			if ( pair.shapeA != kNull ) // kNull: rejected by the host's custom filter
				createContact( w, pair.shapeA, pair.shapeB );
		}
	}
}

// `part`: kPairsAll, or one half of a callback-mediated step: kPairsQuery stops once the candidate pairs stand in
// creation order (the host then runs the custom filter over them and marks the rejected ones, broad_phase.c:267-278),
// kPairsCreate resumes from there.
enum : int
{
	kPairsAll = 0,
	kPairsQuery = 1,
	kPairsCreate = 2,
	kPairsDefer = 3 // everything but the serial creation, which stepCollide runs beside the tree rebuild (createOrderedPairs)
};
template <class Team> F2D_HDF inline void stepPairs( World* w, Team& t, int part = kPairsAll )
{
	int moveCount = w->moveArray.count;
	if ( part != kPairsCreate )
	{
		if ( t.rank() == 0 )
		{
			w->step.orderedPairCount = 0;
			w->step.retryContacts = 0;
		}
		t.sync(); // every caller reads retryContacts right after this routine, whichever way it returns
	}
	if ( moveCount == 0 )
		return;
	if ( moveCount > w->moveHeads.cap )
	{
		if ( t.rank() == 0 )
			setError( w, kErrCapacity, __LINE__ );
		return;
	}
	if ( part != kPairsCreate )
	{
		if constexpr ( Team::kCanSplitTree && Team::kHasSoloBlock == false )
		{
			// One block with an SM to itself, and its L1 is cold at the start of the launch: every query below is a chain
			// of dependent loads (tree nodes, then per hit the shapes, the bodies and a walk over a body's contacts) that
			// would each go to L2. The whole team first reads one word of every 32-byte sector those chains can touch -
			// independent loads, one L2 latency in all - so the chains run on L1 hits.
			const Tree& tree = w->trees[kDynamicBody];
			int acc = warmSectors( ptr( w, tree.nodes ), tree.nodes.count * (int)sizeof( TreeNode ), 32, t.rank(), t.size() );
			acc += warmSectors( ptr( w, w->contacts ), w->contactIds.next * (int)sizeof( Contact ), (int)sizeof( Contact ), t.rank(), t.size() );
			acc += warmSectors( ptr( w, w->bodies ), w->bodyIds.next * (int)sizeof( Body ), (int)sizeof( Body ), t.rank(), t.size() );
			acc += warmSectors( reinterpret_cast<const char*>( ptr( w, w->shapes ) ) + offsetof( Shape, type ),
								w->shapeIds.next * (int)sizeof( Shape ), (int)sizeof( Shape ), t.rank(), t.size() );
			if ( acc == 0x7fffffff ) // (never: keeps the loads alive)
				storeVolatile( &w->step.orderedPairCount, 0 );
		}
		const int32_t* moves = ptr( w, w->moveArray );
		for ( int i = t.rank(); i < moveCount; i += t.size() )
		{
			// the leaf this thread queries next (its box and shape id start the next chain of dependent loads)
			if ( i + t.size() < moveCount )
			{
				int key = moves[i + t.size()];
				if ( key != kNull )
				{
					const char* leaf = reinterpret_cast<const char*>( ptr( w, w->trees[proxyType( key )].nodes ) + proxyId( key ) );
					prefetchL2( leaf );
					prefetchL2( leaf + sizeof( TreeNode ) - 1 );
				}
			}
#if defined( __CUDA_ARCH__ ) && defined( F2D_QUERY_CLOCK )
			{
				long long c0 = clock64();
				findPairsForProxy( w, i );
				long long c1 = clock64();
				atomicMax( reinterpret_cast<unsigned long long*>( &w->prof[29] ), (unsigned long long)( c1 - c0 ) );
				atomicAdd( reinterpret_cast<unsigned long long*>( &w->prof[30] ), (unsigned long long)( c1 - c0 ) );
				atomicAdd( reinterpret_cast<unsigned long long*>( &w->prof[31] ), 1ull );
			}
#else
			findPairsForProxy( w, i );
#endif
		}
	}
	t.sync();
	F2D_MARK( w, t, pfPairQuery );
	// Ordered creation (broad_phase.c:418-453): move-array order x per-proxy list order. Most moved proxies found
	// nothing new, so the lists are first compacted by the team (count, prefix sum, scatter) into one ordered array
	// and rank 0 only walks real pairs.
	{
		const int32_t* heads = ptr( w, w->moveHeads );
		const MovePair* pairs = ptr( w, w->movePairs );
		int32_t* offsets = ptr( w, w->pairOffsets );
		int32_t* ordered = ptr( w, w->pairOrder );
		int total = 0;
		if ( part != kPairsCreate )
		{
			for ( int i = t.rank(); i < moveCount; i += t.size() )
			{
				int n = 0;
				for ( int p = heads[i]; p != kNull; p = pairs[p].next )
					n += 1;
				offsets[i] = n;
			}
			t.sync();
			total = t.exclusiveScan( offsets, moveCount );
			if ( total > 0 && total <= w->pairOrder.cap )
			{
				for ( int i = t.rank(); i < moveCount; i += t.size() )
				{
					int out = offsets[i];
					for ( int p = heads[i]; p != kNull; p = pairs[p].next )
						ordered[out++] = p;
				}
				t.sync();
			}
			else
			{
				total = 0;
			}
			// Nothing structural has changed so far. The reference's arrays grow on demand; here every array that takes a
			// new contact was sized when the image was laid out, so if this step's new pairs do not fit (or the candidate
			// list itself overflowed) the step stops here and the host repeats it on a larger image.
			if ( t.rank() == 0 )
			{
				const int pairCount = w->step.pairCount;
				int need = 0;
				if ( pairCount > w->movePairs.cap || pairCount > w->pairOrder.cap )
					need = pairCount;
				else if ( idCount( w->contactIds ) + total > w->contacts.cap )
					need = total;
				if ( need > 0 )
				{
					w->step.retryContacts = need + 64;
					w->error |= kErrRetry;
					w->locked = false;
				}
			}
			t.sync();
			if ( w->step.retryContacts != 0 )
				return;
			if ( part == kPairsQuery || part == kPairsDefer )
			{
				if ( t.rank() == 0 )
					w->step.orderedPairCount = total;
				t.sync();
				if ( part == kPairsQuery )
					return;
			}
		}
		else
		{
			total = w->step.orderedPairCount;
		}
		if ( part != kPairsDefer )
			createOrderedPairs( w, t, total );
	}
	t.sync();
	// reset move buffer (broad_phase.c:460-462)
	{
		const int32_t* moves = ptr( w, w->moveArray );
		for ( int i = t.rank(); i < moveCount; i += t.size() )
		{
			int key = moves[i];
			if ( key != kNull )
				ptr( w, w->trees[proxyType( key )].nodes )[proxyId( key )].flags &= (uint16_t)~kNodeMoved;
		}
	}
	t.sync();
	if ( t.rank() == 0 )
		w->moveArray.count = 0;
	t.sync();
	F2D_MARK( w, t, pfPairCreate );
}

// ------------------------------------------------------------------------------------------------ collide
// world.c:357-445, one contact. The gather is written as two rounds of independent loads (ids and old impulses
// from the contact record; then shapes, bodies and body sims all at once) because the phase is bound by the
// latency of dependent loads, not by arithmetic.
// The touching transition of one contact after its update (world.c:417-437): flags for the ordered state pass
F2D_HD void markTouchTransition( uint64_t* bits, int contactId, uint32_t& simFlags, bool touching, bool wasTouching )
{
	if ( touching == true && wasTouching == false )
	{
		simFlags |= kSimStartedTouching;
		atomOr64( bits + ( contactId >> 6 ), 1ull << ( contactId & 63 ) );
	}
	else if ( touching == false && wasTouching == true )
	{
		simFlags |= kSimStoppedTouching;
		atomOr64( bits + ( contactId >> 6 ), 1ull << ( contactId & 63 ) );
	}
}

F2D_HDC inline void collideContact( World* w, int contactId, int workIndex )
{
	ContactSim& sim = ptr( w, w->contactSims )[contactId];
	const Shape* shapes = ptr( w, w->shapes );
	const Body* bodies = ptr( w, w->bodies );
	const BodySim* sims = ptr( w, w->sims );
	uint64_t* bits = ptr( w, w->contactBits );
	// ---- round 1: the contact record
	// (four 16-byte chunks = the first two sectors of the record: ids | flags | manifold M0 | M1)
	const Q4 k0 = load16( &sim.shapeIdA ), k1 = load16( &sim.simFlags ), m0 = load16( &sim.manifold.pointCount ),
			 m1 = load16( &sim.manifold.normalImpulse0 );
	const int shapeIdA = (int)floatBits( k0.x ), shapeIdB = (int)floatBits( k0.y );
	const int bodyIdA = (int)floatBits( k0.z ), bodyIdB = (int)floatBits( k0.w );
	uint32_t simFlags = floatBits( k1.x );
	OldImpulses old;
	old.pointCount = (int32_t)floatBits( m0.x );
	old.rollingImpulse = m0.z;
	old.id[0] = (uint16_t)( floatBits( m0.y ) & 0xffffu );
	old.id[1] = (uint16_t)( floatBits( m0.y ) >> 16 );
	old.normalImpulse[0] = m1.x;
	old.tangentImpulse[0] = m1.y;
	old.normalImpulse[1] = m1.z;
	old.tangentImpulse[1] = m1.w;
	// ---- round 2: everything the ids lead to
	const Shape& shapeA = shapes[shapeIdA];
	const Shape& shapeB = shapes[shapeIdB];
	const Body& bodyA = bodies[bodyIdA];
	const Body& bodyB = bodies[bodyIdB];
	const BodySim& simA = sims[bodyIdA];
	const BodySim& simB = sims[bodyIdB];
	const Box fatA = shapeA.fatAABB, fatB = shapeB.fatAABB;
	const int setA = bodyA.setIndex, localA = bodyA.localIndex;
	const int setB = bodyB.setIndex, localB = bodyB.localIndex;
	const Xf xfA = simA.transform, xfB = simB.transform;
	const V2 localCenterA = simA.localCenter, localCenterB = simB.localCenter;
	const float invMassA = simA.invMass, invIA = simA.invInertia;
	const float invMassB = simB.invMass, invIB = simB.invInertia;

	// (the whole round is in flight before the first early-out)
	F2D_ISSUE_F( fatA.lo.x );
	F2D_ISSUE_F( fatB.lo.x );
	F2D_ISSUE_I( setA );
	F2D_ISSUE_I( setB );
	F2D_ISSUE_F( xfA.p.x );
	F2D_ISSUE_F( xfB.p.x );
	F2D_ISSUE_F( localCenterA.x );
	F2D_ISSUE_F( localCenterB.x );
	bool overlap = boxOverlaps( fatA, fatB );
	if ( overlap == false )
	{
		sim.simFlags = ( simFlags | kSimDisjoint ) & ~kSimTouching;
		atomOr64( bits + ( contactId >> 6 ), 1ull << ( contactId & 63 ) );
		return;
	}
	bool wasTouching = ( simFlags & kSimTouching ) != 0;
	sim.bodySimIndexA = setA == kAwakeSet ? localA : kNull;
	sim.bodySimIndexB = setB == kAwakeSet ? localB : kNull;
	store16( &sim.invMassA, Q4{ invMassA, invIA, invMassB, invIB } );
	V2 centerOffsetA = rotate( xfA.q, localCenterA );
	V2 centerOffsetB = rotate( xfB.q, localCenterB );
	bool touching = updateContact( w, sim, simFlags, old, shapeA, xfA, centerOffsetA, shapeB, xfB, centerOffsetB );
	if ( simFlags & kSimPendingPreSolve )
	{
		// callback-mediated step: queue the contact for the host's pre-solve callback (position in the work list = the
		// reference's single-worker call order) and leave the rest to finishDeferredContact
		int slot = atomAdd( &w->step.preSolveCount, 1 );
		if ( slot < w->stateList.cap && slot < w->pairOrder.cap )
		{
			ptr( w, w->stateList )[slot] = contactId;
			ptr( w, w->pairOrder )[slot] = workIndex;
		}
		else
		{
			setError( w, kErrCapacity, __LINE__ );
		}
		sim.simFlags = simFlags;
		return;
	}
	markTouchTransition( bits, contactId, simFlags, touching, wasTouching );
	sim.simFlags = simFlags;
}

// Second half of a contact that waited for the host's pre-solve callback (contact.c:504-633, world.c:417-437):
// `approved` is the callback's return value.
F2D_HDF inline void finishDeferredContact( World* w, int contactId, bool approved )
{
	ContactSim& sim = ptr( w, w->contactSims )[contactId];
	const Shape* shapes = ptr( w, w->shapes );
	const BodySim* sims = ptr( w, w->sims );
	uint32_t simFlags = sim.simFlags & ~kSimPendingPreSolve;
	const bool wasTouching = ( simFlags & kSimTouching ) != 0;
	Manifold m = unpackManifold( sim.manifold );
	OldImpulses old = unparkOldImpulses( m );
	const Shape& shapeA = shapes[sim.shapeIdA];
	const Shape& shapeB = shapes[sim.shapeIdB];
	const BodySim& simA = sims[sim.bodyIdA];
	const BodySim& simB = sims[sim.bodyIdB];
	V2 centerOffsetA = rotate( simA.transform.q, simA.localCenter );
	V2 centerOffsetB = rotate( simB.transform.q, simB.localCenter );
	if ( approved == false )
		m.pointCount = 0; // "disable contact"
	bool touching = finishContactUpdate( w, sim, simFlags, m, old, approved, shapeA.enableHitEvents || shapeB.enableHitEvents,
										 centerOffsetA, centerOffsetB );
	markTouchTransition( ptr( w, w->contactBits ), contactId, simFlags, touching, wasTouching );
	sim.simFlags = simFlags;
}

// One contact of the ordered contact-state pass (world.c:595-684)
F2D_HDF inline void contactStateChange( World* w, int contactId )
{
	Contact* contacts = ptr( w, w->contacts );
	ContactSim* sims = ptr( w, w->contactSims );
	const Shape* shapes = ptr( w, w->shapes );
	Contact& c = contacts[contactId];
	ContactSim& sim = sims[contactId];
	int colorIndex = c.colorIndex;
	int localIndex = c.localIndex;
	const Shape& shapeA = shapes[c.shapeIdA];
	const Shape& shapeB = shapes[c.shapeIdB];
	uint32_t flags = c.flags;
	uint32_t simFlags = sim.simFlags;

	if ( simFlags & kSimDisjoint )
	{
		destroyContact( w, contactId, false );
	}
	else if ( simFlags & kSimStartedTouching )
	{
		if ( flags & kContactEnableContactEvents )
		{
			BeginTouchEvent ev;
			ev.a = makeShapeId( w, shapeA );
			ev.b = makeShapeId( w, shapeB );
			ev.manifold = unpackManifold( sim.manifold );
			F2D_PUSH_EVENT( w, w->beginEvents, ev );
		}
		c.flags |= kContactTouching;
		linkContact( w, c, contactId );
		sim.simFlags &= ~kSimStartedTouching;
		addContactToGraph( w, contactId );
		// remove from the awake non-touching list (world.c:472-485); c.localIndex now is the colour slot
		int moved = removeSwap( w, w->awakeContacts, localIndex );
		if ( moved != kNull )
			contacts[ptr( w, w->awakeContacts )[localIndex]].localIndex = localIndex;
	}
	else if ( simFlags & kSimStoppedTouching )
	{
		sim.simFlags &= ~kSimStoppedTouching;
		c.flags &= ~kContactTouching;
		if ( c.flags & kContactEnableContactEvents )
		{
			EndTouchEvent ev = { makeShapeId( w, shapeA ), makeShapeId( w, shapeB ) };
			F2D_PUSH_EVENT( w, w->endEvents[w->endEventArrayIndex], ev );
		}
		unlinkContact( w, c, contactId );
		int bodyIdA = c.edges[0].bodyId;
		int bodyIdB = c.edges[1].bodyId;
		// back to the awake non-touching list (world.c:461-470)
		c.colorIndex = kNull;
		c.localIndex = w->awakeContacts.count;
		F2D_PUSH( w, w->awakeContacts, contactId );
		removeContactFromGraph( w, bodyIdA, bodyIdB, colorIndex, localIndex );
	}
}

// The same state change as two halves that touch DISJOINT data and can therefore run on two threads at once, each
// visiting the flagged contacts in ascending id: the island half (touching flag, begin / end events, island link /
// unlink) and the graph half (colour graph, awake / colour id lists, body contact-edge lists, contact id pool, solver
// indices). `kind` is the contact's class, read from its flags by the compaction pass before either half changed them.
// Not usable when a begin-touch wakes a sleeping set (island.c:121-133): that moves bodies and contacts between sets and
// colours - contactStatePass then falls back to the one-thread version above.
enum : int
{
	kStateDisjoint = 0,
	kStateStarted = 1,
	kStateStopped = 2,
	kStateNone = 3
};
F2D_HDF inline void contactStateIslandHalf( World* w, int contactId, int kind )
{
	Contact* contacts = ptr( w, w->contacts );
	Contact& c = contacts[contactId];
	const Shape* shapes = ptr( w, w->shapes );
	const uint32_t flags = c.flags;
	if ( kind == kStateDisjoint )
	{
		// destroyContact( w, contactId, false ): end event of a touching contact, island unlink
		if ( ( flags & kContactTouching ) != 0 && ( flags & kContactEnableContactEvents ) != 0 )
		{
			EndTouchEvent ev = { makeShapeId( w, shapes[c.shapeIdA] ), makeShapeId( w, shapes[c.shapeIdB] ) };
			F2D_PUSH_EVENT( w, w->endEvents[w->endEventArrayIndex], ev );
		}
		if ( c.islandId != kNull )
			unlinkContact( w, c, contactId );
	}
	else if ( kind == kStateStarted )
	{
		if ( flags & kContactEnableContactEvents )
		{
			BeginTouchEvent ev;
			ev.a = makeShapeId( w, shapes[c.shapeIdA] );
			ev.b = makeShapeId( w, shapes[c.shapeIdB] );
			ev.manifold = unpackManifold( ptr( w, w->contactSims )[contactId].manifold );
			F2D_PUSH_EVENT( w, w->beginEvents, ev );
		}
		c.flags = flags | kContactTouching;
		linkContact( w, c, contactId );
	}
	else if ( kind == kStateStopped )
	{
		c.flags = flags & ~kContactTouching;
		if ( flags & kContactEnableContactEvents )
		{
			EndTouchEvent ev = { makeShapeId( w, shapes[c.shapeIdA] ), makeShapeId( w, shapes[c.shapeIdB] ) };
			F2D_PUSH_EVENT( w, w->endEvents[w->endEventArrayIndex], ev );
		}
		unlinkContact( w, c, contactId );
	}
}
F2D_HDF inline void contactStateGraphHalf( World* w, int contactId, int kind )
{
	Contact* contacts = ptr( w, w->contacts );
	Contact& c = contacts[contactId];
	ContactSim& sim = ptr( w, w->contactSims )[contactId];
	const int colorIndex = c.colorIndex;
	const int localIndex = c.localIndex;
	const int bodyIdA = c.edges[0].bodyId, bodyIdB = c.edges[1].bodyId;
	if ( kind == kStateDisjoint )
	{
		// destroyContact( w, contactId, false ): body contact-edge lists, graph / set list, record, id pool
		Body* bodies = ptr( w, w->bodies );
		const Edge edgeA = c.edges[0], edgeB = c.edges[1];
		if ( edgeA.prevKey != kNull )
			contacts[edgeA.prevKey >> 1].edges[edgeA.prevKey & 1].nextKey = edgeA.nextKey;
		if ( edgeA.nextKey != kNull )
			contacts[edgeA.nextKey >> 1].edges[edgeA.nextKey & 1].prevKey = edgeA.prevKey;
		if ( bodies[bodyIdA].headContactKey == ( ( contactId << 1 ) | 0 ) )
			bodies[bodyIdA].headContactKey = edgeA.nextKey;
		bodies[bodyIdA].contactCount -= 1;
		if ( edgeB.prevKey != kNull )
			contacts[edgeB.prevKey >> 1].edges[edgeB.prevKey & 1].nextKey = edgeB.nextKey;
		if ( edgeB.nextKey != kNull )
			contacts[edgeB.nextKey >> 1].edges[edgeB.nextKey & 1].prevKey = edgeB.prevKey;
		if ( bodies[bodyIdB].headContactKey == ( ( contactId << 1 ) | 1 ) )
			bodies[bodyIdB].headContactKey = edgeB.nextKey;
		bodies[bodyIdB].contactCount -= 1;
		if ( colorIndex != kNull )
			removeContactFromGraph( w, bodyIdA, bodyIdB, colorIndex, localIndex );
		else
			removeContactFromSetList( w, c );
		c.contactId = kNull;
		c.setIndex = kNull;
		c.colorIndex = kNull;
		c.localIndex = kNull;
		freeId( w, w->contactIds, contactId );
	}
	else if ( kind == kStateStarted )
	{
		sim.simFlags &= ~kSimStartedTouching;
		addContactToGraph( w, contactId );
		// remove from the awake non-touching list (world.c:472-485); c.localIndex now is the colour slot
		int moved = removeSwap( w, w->awakeContacts, localIndex );
		if ( moved != kNull )
			contacts[ptr( w, w->awakeContacts )[localIndex]].localIndex = localIndex;
	}
	else if ( kind == kStateStopped )
	{
		sim.simFlags &= ~kSimStoppedTouching;
		// back to the awake non-touching list (world.c:461-470)
		c.colorIndex = kNull;
		c.localIndex = w->awakeContacts.count;
		F2D_PUSH( w, w->awakeContacts, contactId );
		removeContactFromGraph( w, bodyIdA, bodyIdB, colorIndex, localIndex );
	}
}

// Ordered contact-state pass, ascending contact id (world.c:587-686). The flagged ids are compacted by the whole team
// (popcount per 64-bit word + prefix sum) together with their class; then the order-defining structural edits are
// applied one contact after the other - by two threads, one per half (see above), when the team has them.
constexpr int kStateKindShift = 28; // list entry = contact id | class << 28
// first half, the whole team: the compacted list (World::stateList); returns its length (0 also when a capacity ran out)
template <class Team> F2D_HDF inline int contactStateCollect( World* w, Team& t )
{
	const uint64_t* bits = ptr( w, w->contactBits );
	int wordCount = ( w->contactIds.next + 63 ) >> 6;
	if ( wordCount > w->contactBits.cap )
		wordCount = w->contactBits.cap;
	int32_t* offsets = ptr( w, w->stateOffsets );
	int32_t* list = ptr( w, w->stateList );
	if ( wordCount > w->stateOffsets.cap )
	{
		if ( t.rank() == 0 )
			setError( w, kErrCapacity, __LINE__ );
		return 0;
	}
	for ( int k = t.rank(); k < wordCount; k += t.size() )
	{
		uint64_t word = bits[k];
		int n = 0;
		while ( word != 0 )
		{
			word &= word - 1;
			n += 1;
		}
		offsets[k] = n;
	}
	if ( t.rank() == 0 )
		w->step.stateNeedsSerial = 0;
	t.sync();
	int total = t.exclusiveScan( offsets, wordCount );
	if ( total == 0 )
		return 0;
	if ( total > w->stateList.cap )
	{
		if ( t.rank() == 0 )
			setError( w, kErrCapacity, __LINE__ );
		return 0;
	}
	constexpr int kKindShift = kStateKindShift;
	{
		const ContactSim* sims = ptr( w, w->contactSims );
		const Contact* contacts = ptr( w, w->contacts );
		const Body* bodies = ptr( w, w->bodies );
		for ( int k = t.rank(); k < wordCount; k += t.size() )
		{
			uint64_t word = bits[k];
			int out = offsets[k];
			// lowest set bit first: the j-th flagged contact of every lane's word is handled in the same iteration (a
			// loop over all 64 bit positions runs the gathers below for one lane at a time)
			while ( word != 0 )
			{
				{
					const int id = 64 * k + lowestBit64( word );
					word &= word - 1;
					const uint32_t simFlags = sims[id].simFlags;
					int kind = kStateNone;
					if ( simFlags & kSimDisjoint )
						kind = kStateDisjoint;
					else if ( simFlags & kSimStartedTouching )
					{
						kind = kStateStarted;
						const Contact& c = contacts[id];
						if ( ( bodies[c.edges[0].bodyId].setIndex >= kFirstSleepingSet ) | ( bodies[c.edges[1].bodyId].setIndex >= kFirstSleepingSet ) )
							storeVolatile( &w->step.stateNeedsSerial, 1 );
					}
					else if ( simFlags & kSimStoppedTouching )
						kind = kStateStopped;
					list[out++] = id | ( kind << kKindShift );
				}
			}
		}
	}
	t.sync();
	return total;
}

// second half: the structural edits, in list order, on one or two threads of the team (the others return at once)
template <class Team> F2D_HDF inline void contactStateApply( World* w, Team& t, int total )
{
	constexpr int kKindShift = kStateKindShift;
	const int32_t* list = ptr( w, w->stateList );
	const bool twoThreads = t.size() >= 64 && w->step.stateNeedsSerial == 0 && w->contactIds.next < ( 1 << kKindShift );
	if constexpr ( Team::kHasSoloBlock )
	{
		// A grid: the records the two serial threads are about to chase were written by other SMs (narrowphase) and sit
		// in L2; the other threads of their block read them once, in parallel, which parks them in this SM's L1 - the
		// serial chains then run on L1 hits (~40 cycles) instead of L2 round trips (~700).
		if ( t.inSoloBlock() && t.rank() >= 64 )
		{
			const Contact* contacts = ptr( w, w->contacts );
			const ContactSim* sims = ptr( w, w->contactSims );
			const Body* bodies = ptr( w, w->bodies );
			const BodySim* bsims = ptr( w, w->sims );
			const Island* islands = ptr( w, w->islands );
			const Shape* shapes = ptr( w, w->shapes );
			int acc = 0;
			for ( int i = t.rank() - 64; i < total && i < 1024; i += t.soloSize() - 64 )
			{
				const int id = list[i] & ( ( 1 << kKindShift ) - 1 );
				const Contact& c = contacts[id];
				const int a = c.edges[0].bodyId, b = c.edges[1].bodyId;
				acc += c.islandId + c.colorIndex + c.edges[0].nextKey + c.edges[1].nextKey;
				acc += (int)sims[id].simFlags + sims[id].bodySimIndexA + sims[id].manifold.pointCount;
				const int ia = bodies[a].islandId, ib = bodies[b].islandId;
				acc += bodies[a].setIndex + bodies[b].setIndex + bodies[a].colorMask + bodies[b].colorMask;
				acc += floatBits( bsims[a].invMass ) + floatBits( bsims[b].invMass );
				if ( ia != kNull )
					acc += islands[ia].headContact + islands[ia].parentIsland;
				if ( ib != kNull )
					acc += islands[ib].headContact + islands[ib].parentIsland;
				acc += shapes[c.shapeIdA].generation + shapes[c.shapeIdB].generation;
			}
			if ( acc == 0x7fffffff ) // (never: keeps the loads alive)
				storeVolatile( &w->step.stateNeedsSerial, 0 );
		}
	}
	if ( twoThreads )
	{
		if ( t.rank() == 0 )
		{
			for ( int i = 0; i < total; ++i )
				contactStateIslandHalf( w, list[i] & ( ( 1 << kKindShift ) - 1 ), list[i] >> kKindShift );
		}
		else if ( t.rank() == 32 )
		{
			for ( int i = 0; i < total; ++i )
				contactStateGraphHalf( w, list[i] & ( ( 1 << kKindShift ) - 1 ), list[i] >> kKindShift );
		}
	}
	else if ( t.rank() == 0 )
	{
		for ( int i = 0; i < total; ++i )
			contactStateChange( w, list[i] & ( ( 1 << kKindShift ) - 1 ) );
	}
}

template <class Team> F2D_HDF inline void contactStatePass( World* w, Team& t )
{
	const int total = contactStateCollect( w, t );
	if ( total > 0 )
		contactStateApply( w, t, total );
}

// `part`: kCollideAll, or one half of a callback-mediated step: kCollideNarrow stops after the narrowphase (the host
// then answers the queued pre-solve callbacks: World::stateList = contact ids, rewritten by the host as verdicts in
// World::pairOrder), kCollideFinish completes those contacts and runs the ordered state pass.
enum : int
{
	kCollideAll = 0,
	kCollideNarrow = 1,
	kCollideFinish = 2,
	kCollideTreeOnly = 3,  // profiling aid (F2D_PROFILE_PHASE_LAUNCHES): the tree rebuild alone
	kCollideNarrowOnly = 4 // profiling aid: the narrowphase alone
};
template <class Team> F2D_HDF inline void stepCollide( World* w, Team& t, int part = kCollideAll, bool deferredPairs = false )
{
	if ( part == kCollideFinish )
	{
		const int pending = mini( w->step.preSolveCount, mini( w->stateList.cap, w->pairOrder.cap ) );
		const int32_t* ids = ptr( w, w->stateList );
		const int32_t* verdicts = ptr( w, w->pairOrder );
		for ( int k = t.rank(); k < pending; k += t.size() )
			finishDeferredContact( w, ids[k], verdicts[k] != 0 );
		t.sync();
		contactStatePass( w, t );
		t.sync();
		F2D_MARK( w, t, pfStatePass );
		return;
	}
	if ( t.rank() == 0 )
		w->step.preSolveCount = 0;

	// work list = colour 0..11 lists then the awake non-touching list (world.c:504-542)
	const int awakeBefore = w->awakeContacts.count; // (read by everybody before the barrier that ends the narrowphase)
	int total = awakeBefore;
	for ( int i = 0; i < kColorCount; ++i )
		total += w->colorContacts[i].count;

	// Rebuild of the dynamic and kinematic trees (world.c:499, broad_phase.c:488-492). The reference overlaps it with
	// the narrowphase on another worker; here the whole team does one after the other, each fully data-parallel
	// (measured: a single block rebuilding beside the narrowphase is slower than the grid doing both in turn).
	// a thread's indices only grow, so the list segment is tracked with a forward cursor (no per-thread table)
	struct WorkCursor
	{
		World* w;
		int total;
		int seg = -1, segStart = 0, segEnd = 0;
		const int32_t* list = nullptr;
		F2D_HDF int at( int i )
		{
			if ( i >= total )
				return kNull;
			while ( i >= segEnd )
			{
				seg += 1;
				segStart = segEnd;
				const Arr<int32_t>& a = seg < kColorCount ? w->colorContacts[seg] : w->awakeContacts;
				segEnd += a.count;
				list = ptr( w, a );
			}
			return list[i - segStart];
		}
	};
	// Narrowphase per pair class (contact.c:166-184 registers one manifold function per pair of shape types): a world
	// with more than one shape type first bins its work list by class - counts, bases, scatter into World::stateList,
	// which is free until the state pass - so that the lanes of a warp run the same manifold function instead of
	// serialising a switch over up to six of them. A world with one shape type (bench2d: boxes) has one class and skips
	// the binning. The work-list position of a contact only matters to the callback-mediated step (pre-solve call
	// order), which does not come through here.
	// (two shape types - say boxes on a ground segment - mean one dominant class: not worth two passes over the list)
	const bool binned = part == kCollideAll && w->pairClassBinningOff == 0 && popCount32( w->shapeTypeMask ) >= 3 &&
						total <= w->stateList.cap && total > 0;
	if ( binned )
	{
		const ContactSim* sims = ptr( w, w->contactSims );
		int32_t* sorted = ptr( w, w->stateList );
		typename Team::Lanes L;
		// one atomic per class and warp, not per contact (a grid has tens of thousands of contacts on a handful of counters)
		auto claim = [&]( bool active, int cls, int32_t* counters ) -> int {
			uint32_t pending = L.ballot( active );
			int slot = 0;
			while ( pending != 0 )
			{
				const int leader = lowestBit32( pending );
				const int leaderClass = L.from( cls, leader );
				const uint32_t same = L.ballot( active && cls == leaderClass );
				int base = 0;
				if ( L.lane() == leader )
					base = atomAdd( counters + leaderClass, popCount32( same ) );
				base = L.from( base, leader );
				if ( active && cls == leaderClass )
					slot = base + popCount32( same & ( ( 1u << L.lane() ) - 1u ) );
				pending &= ~same;
			}
			return slot;
		};
		for ( int k = t.rank(); k < kPairClassCount; k += t.size() )
			w->step.classCount[k] = 0;
		t.sync();
		{
			WorkCursor cursor{ w, total };
			for ( int base = t.rank() - L.lane(); base < total; base += t.size() )
			{
				const int i = base + L.lane();
				const bool active = i < total;
				const int cls = active ? ( sims[cursor.at( i )].pairClass & ( kPairClassCount - 1 ) ) : 0;
				claim( active, cls, w->step.classCount );
			}
		}
		t.sync();
		if ( t.rank() == 0 )
		{
			int base = 0;
			for ( int k = 0; k < kPairClassCount; ++k )
			{
				w->step.classFill[k] = base;
				base += w->step.classCount[k];
			}
		}
		t.sync();
		{
			WorkCursor cursor{ w, total };
			for ( int base = t.rank() - L.lane(); base < total; base += t.size() )
			{
				const int i = base + L.lane();
				const bool active = i < total;
				const int id = active ? cursor.at( i ) : 0;
				const int cls = active ? ( sims[id].pairClass & ( kPairClassCount - 1 ) ) : 0;
				const int slot = claim( active, cls, w->step.classFill );
				if ( active )
					sorted[slot] = id;
			}
		}
		t.sync();
	}
	auto narrowphase = [&]( int rank, int size ) {
		// A contact is a chain of dependent gathers (id -> contact record -> shapes / bodies / body sims). No prefetch of
		// the thread's next record: in a batch this phase keeps DRAM busy 96 % of the time (profiles/README.md, r02k), and a
		// prefetch moves whole 128-byte lines where the gathers touch single sectors.
		if ( binned )
		{
			const int32_t* sorted = ptr( w, w->stateList );
			for ( int i = rank; i < total; i += size )
				collideContact( w, sorted[i], i );
			return;
		}
		WorkCursor cursor{ w, total };
		for ( int i = rank; i < total; i += size )
			collideContact( w, cursor.at( i ), i );
	};
	// A large team rebuilds the trees later, beside the serial parts of the step (below).
	bool treeBeside = false;
	if constexpr ( Team::kCanSplitTree )
		treeBeside = part == kCollideAll && t.canSplitTree();
	if ( part != kCollideNarrowOnly && treeBeside == false )
	{
		treeRebuildTeam( w, t, w->trees[kDynamicBody] );
		treeRebuildTeam( w, t, w->trees[kKinematicBody] );
		t.sync();
		F2D_MARK( w, t, pfTreeRebuild );
	}
	if ( part == kCollideTreeOnly )
		return;
	narrowphase( t.rank(), t.size() );
	t.sync();
	F2D_MARK( w, t, pfNarrow );
	if ( part == kCollideNarrow || part == kCollideNarrowOnly )
		return;
	if constexpr ( Team::kCanSplitTree )
	{
		if ( treeBeside )
		{
			// Two parts of the step keep one or two threads busy and leave the rest of the team idle: the ordered creation
			// of this step's new contacts (stepPairs left it to us: `deferredPairs`) and the ordered contact-state pass.
			// Both touch contacts, bodies' contact lists, islands and the constraint graph; the rebuild touches the trees
			// and its own work arrays, and nothing reads the trees again before finalize enlarges proxies. So the front of
			// the team (two warps of a block / block 0 of a grid) runs the serial parts - with the narrowphase of the new
			// contacts, the last entries of the non-touching list, in between - while the rear rebuilds. The reference
			// runs the rebuild as a task beside the whole of b2Collide (world.c:499, joined at solver.c:1826-1831).
			if ( t.inFront() )
			{
				auto front = t.front();
				if ( deferredPairs )
				{
					createOrderedPairs( w, front, w->step.orderedPairCount );
					front.sync();
					F2D_MARK( w, front, pfPairCreate );
					const int created = w->awakeContacts.count - awakeBefore;
					const int32_t* list = ptr( w, w->awakeContacts );
					for ( int i = front.rank(); i < created; i += front.size() )
						collideContact( w, list[awakeBefore + i], total + i );
					front.sync();
					F2D_MARK( w, front, pfNarrow );
				}
				const int changed = contactStateCollect( w, front );
				if ( changed > 0 )
					contactStateApply( w, front, changed );
				front.sync();
				F2D_MARK( w, front, pfStatePass );
			}
			else
			{
				auto rear = t.rear();
				treeRebuildTeam( w, rear, w->trees[kDynamicBody] );
				treeRebuildTeam( w, rear, w->trees[kKinematicBody] );
				if ( rear.rank() == 0 && w->profEnabled )
					w->prof[pfTreeBeside] += profClock() - w->profLast;
			}
			t.sync();
			F2D_MARK( w, t, pfTreeRebuild ); // (what the front waited for the rear)
			return;
		}
	}
	contactStatePass( w, t );
	t.sync();
	F2D_MARK( w, t, pfStatePass );
}

// ------------------------------------------------------------------------------------------------ solve
// solver.c:65-129. Everything the velocity integration reads from the body record is constant during a step (forces
// and torques are only cleared by finalize), so the damping factors and velocity deltas - the reference recomputes
// them, from the same inputs, in every sub-step - are evaluated once per step into a dense field-major table by awake
// index: the sub-steps then read 24 coalesced bytes per body instead of gathering a 100-byte record.
F2D_HDF inline void prepareIntegrate( World* w, int awakeIndex, float h )
{
	const BodySim& sim = ptr( w, w->sims )[ptr( w, w->awakeBodies )[awakeIndex]];
	float* c = ptr( w, w->integ ) + awakeIndex;
	const int stride = w->integ.cap / 6;
	float linearDamping = 1.0f / ( 1.0f + h * sim.linearDamping );
	float angularDamping = 1.0f / ( 1.0f + h * sim.angularDamping );
	float gravityScale = sim.invMass > 0.0f ? sim.gravityScale : 0.0f;
	V2 linearVelocityDelta = add( mulSV( h * sim.invMass, sim.force ), mulSV( h * gravityScale, w->gravity ) );
	float angularVelocityDelta = h * sim.invInertia * sim.torque;
	c[0] = linearVelocityDelta.x;
	c[stride] = linearVelocityDelta.y;
	c[2 * stride] = angularVelocityDelta;
	c[3 * stride] = linearDamping;
	c[4 * stride] = angularDamping;
	c[5 * stride] = sim.allowFastRotation ? 1.0f : 0.0f;
}

F2D_HDF inline void integrateVelocity( World* w, int awakeIndex, float h, float maxLinearSpeed, float maxAngularSpeed )
{
	(void)h;
	BodyState& state = ptr( w, w->states )[awakeIndex];
	const float* c = ptr( w, w->integ ) + awakeIndex;
	const int stride = w->integ.cap / 6;
	const Q4 vw = load16( &state.v ); // v, w, flags
	V2 v = { vw.x, vw.y };
	float wv = vw.z;
	float maxLinearSpeedSquared = maxLinearSpeed * maxLinearSpeed;
	float maxAngularSpeedSquared = maxAngularSpeed * maxAngularSpeed;
	V2 linearVelocityDelta = { c[0], c[stride] };
	float angularVelocityDelta = c[2 * stride];
	float linearDamping = c[3 * stride];
	float angularDamping = c[4 * stride];
	v = mulAdd( linearVelocityDelta, linearDamping, v );
	wv = angularVelocityDelta + angularDamping * wv;
	bool capped = false;
	if ( dot( v, v ) > maxLinearSpeedSquared )
	{
		float ratio = maxLinearSpeed / length( v );
		v = mulSV( ratio, v );
		capped = true;
	}
	if ( wv * wv > maxAngularSpeedSquared && c[5 * stride] == 0.0f )
	{
		float ratio = maxAngularSpeed / absf( wv );
		wv *= ratio;
		capped = true;
	}
	if ( capped )
		ptr( w, w->sims )[ptr( w, w->awakeBodies )[awakeIndex]].isSpeedCapped = true;
	store16( &state.v, Q4{ v.x, v.y, wv, vw.w } );
}

// solver.c:182-199
F2D_HD void integratePosition( BodyState& state, float h )
{
	const Q4 vw = load16( &state.v ), pq = load16( &state.dp );
	const Rot dq = integrateRot( Rot{ pq.z, pq.w }, h * vw.z );
	const V2 dp = mulAdd( V2{ pq.x, pq.y }, h, V2{ vw.x, vw.y } );
	store16( &state.dp, Q4{ dp.x, dp.y, dq.c, dq.s } );
}

// Colour-parallel constraint stage: joints and contacts of colour `color` (independent inside a colour,
// solver.c:773-808). stage: 0 warm start, 1 solve(useBias), 2 relax, 3 restitution
template <class Team> F2D_HDF inline void colorStage( World* w, Team& t, int color, int stage )
{
	BodyState* states = ptr( w, w->states );
	const ConView c = conView( w );
	int jointCount = w->colorJoints[color].count;
	int contactCount = w->colorContacts[color].count;
	int base = w->step.colorBase[color];
	float inv_h = w->step.inv_h;
	float contactSpeed = w->contactSpeed;
	if ( stage != 3 && jointCount > 0 )
	{
		const int32_t* jl = ptr( w, w->colorJoints[color] );
		JointSim* jsims = ptr( w, w->jointSims );
		for ( int i = t.rank(); i < jointCount; i += t.size() )
		{
			JointSim& js = jsims[jl[i]];
			if ( stage == 0 )
				warmStartJoint( w, js, states );
			else
				solveJoint( w, js, states, stage == 1 );
		}
	}
	for ( int i = t.rank(); i < contactCount; i += t.size() )
	{
		int slot = base + i;
		if ( stage == 0 )
			warmStartSlot( c, slot, states );
		else if ( stage == 1 )
			solveSlot( c, slot, states, true, inv_h, contactSpeed );
		else if ( stage == 2 )
			solveSlot( c, slot, states, false, inv_h, contactSpeed );
		else
		{
			// the reference skips a 4-lane group when no lane has restitution (contact_solver.c:1982-1986)
			int g0 = ( i >> 2 ) << 2;
			bool any = false;
			for ( int k = g0; k < g0 + 4 && k < contactCount; ++k )
				any = any || ( c.f( cfRestitution, base + k ) != 0.0f );
			if ( any )
				restitutionSlot( c, slot, states, w->step.restitutionThreshold );
		}
	}
}

// Overflow colour, serial in array order, joints before contacts (solver.c:1008-1009, 1027-1028, 1055-1056, 1077)
F2D_HDF inline void overflowStage( World* w, int stage )
{
	BodyState* states = ptr( w, w->states );
	const ConView c = conView( w );
	int base = w->step.colorBase[kOverflow];
	int jointCount = w->colorJoints[kOverflow].count;
	int contactCount = w->colorContacts[kOverflow].count;
	if ( stage != 3 )
	{
		const int32_t* jl = ptr( w, w->colorJoints[kOverflow] );
		JointSim* jsims = ptr( w, w->jointSims );
		for ( int i = 0; i < jointCount; ++i )
		{
			if ( stage == 0 )
				warmStartJoint( w, jsims[jl[i]], states );
			else
				solveJoint( w, jsims[jl[i]], states, stage == 1 );
		}
	}
	for ( int i = 0; i < contactCount; ++i )
	{
		if ( stage == 0 )
			overflowWarmStart( c, base + i, states );
		else if ( stage == 1 )
			overflowSolve( c, base + i, states, true, w->step.inv_h, w->maxContactPushSpeed );
		else if ( stage == 2 )
			overflowSolve( c, base + i, states, false, w->step.inv_h, w->maxContactPushSpeed );
		else
			overflowRestitution( c, base + i, states, w->restitutionThreshold );
	}
}

template <class Team> F2D_HDF inline void constraintPass( World* w, Team& t, int stage )
{
	bool hasOverflow = w->colorContacts[kOverflow].count + w->colorJoints[kOverflow].count > 0;
	if ( hasOverflow )
	{
		if ( t.rank() == 0 )
			overflowStage( w, stage );
		t.sync();
	}
	int n = w->step.activeColorCount;
	for ( int k = 0; k < n; ++k )
	{
		colorStage( w, t, w->step.activeColors[k], stage );
		t.sync();
	}
}

// ------------------------------------------------------------------------------------------------ island-parallel solve
// When the awake world is many small islands (many_pyramids: 400 piles of 55 boxes) the colour-by-colour solve wastes
// its time in team-wide barriers: ~100 stages per step, each with a few thousand independent constraints. Islands
// share no awake body, so every island can run the complete sub-step loop on its own: one group (a warp) per island,
// colour after colour with warp-level synchronisation only. Inside a colour the constraints of an island touch
// disjoint bodies, exactly as in the global colour, so the floating-point results are identical to the colour-parallel
// path (and to the reference) whatever order the warps run in.
constexpr int kIslandPathMinIslands = 16;
constexpr int kIslandPathMaxContacts = 256; // ~32 constraints per colour: one pass of a warp
constexpr int kIslandPathMaxBodies = 128;

template <class Team> F2D_HDF inline void islandPartition( World* w, Team& t )
{
	const StepCtx& step = w->step;
	const int islandCount = w->awakeIslands.count;
	const int awakeBodyCount = step.awakeBodyCount;
	const int slotCount = step.awakeContactCount;
	const Island* islands = ptr( w, w->islands );
	const int32_t* awakeIslands = ptr( w, w->awakeIslands );
	const Contact* contacts = ptr( w, w->contacts );
	const Body* bodies = ptr( w, w->bodies );
	const int32_t* awakeBodies = ptr( w, w->awakeBodies );
	int32_t* slotOff = ptr( w, w->islSlotOff );
	int32_t* bodyOff = ptr( w, w->islBodyOff );
	int32_t* colorOff = ptr( w, w->islColorOff );
	int32_t* colorFill = ptr( w, w->islColorFill );
	int32_t* bodyFill = ptr( w, w->islBodyFill );
	int32_t* islSlots = ptr( w, w->islSlots );
	int32_t* islBodies = ptr( w, w->islBodies );

	// ---- group the constraint slots by (island, colour) and the awake bodies by island
	for ( int i = t.rank(); i < islandCount; i += t.size() )
	{
		const Island& is = islands[awakeIslands[i]];
		slotOff[i] = is.contactCount;
		bodyOff[i] = is.bodyCount;
		bodyFill[i] = 0;
	}
	for ( int i = t.rank(); i < islandCount * kColorCount; i += t.size() )
	{
		colorOff[i] = 0;
		colorFill[i] = 0;
	}
	t.sync();
	int slotTotal = t.exclusiveScan( slotOff, islandCount );
	int bodyTotal = t.exclusiveScan( bodyOff, islandCount );
	if ( slotTotal != slotCount || bodyTotal != awakeBodyCount )
	{
		// some constraint or body is not owned by an awake island: keep the colour-parallel path for this step
		t.sync();
		if ( t.rank() == 0 )
			w->step.islandPath = 0;
		t.sync();
		return;
	}
	for ( int slot = t.rank(); slot < slotCount; slot += t.size() )
	{
		int color = 0;
		while ( slot >= step.colorBase[color + 1] )
			color += 1;
		int contactId = ptr( w, w->colorContacts[color] )[slot - step.colorBase[color]];
		int li = islands[contacts[contactId].islandId].localIndex;
		atomAdd( colorOff + li * kColorCount + color, 1 );
	}
	t.sync();
	for ( int i = t.rank(); i < islandCount; i += t.size() )
	{
		int run = slotOff[i];
		for ( int c = 0; c < kColorCount; ++c )
		{
			int n = colorOff[i * kColorCount + c];
			colorOff[i * kColorCount + c] = run;
			run += n;
		}
	}
	t.sync();
	for ( int slot = t.rank(); slot < slotCount; slot += t.size() )
	{
		int color = 0;
		while ( slot >= step.colorBase[color + 1] )
			color += 1;
		int contactId = ptr( w, w->colorContacts[color] )[slot - step.colorBase[color]];
		int li = islands[contacts[contactId].islandId].localIndex;
		int k = li * kColorCount + color;
		islSlots[colorOff[k] + atomAdd( colorFill + k, 1 )] = slot;
	}
	for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
	{
		int li = islands[bodies[awakeBodies[i]].islandId].localIndex;
		islBodies[bodyOff[li] + atomAdd( bodyFill + li, 1 )] = i;
	}
	if ( t.rank() == 0 )
		w->step.islandSolveCount = islandCount;
	t.sync();
}

// One group per island: the whole sub-step loop with group-level synchronisation
template <class Team> F2D_HDF inline void islandSolve( World* w, Team& t )
{
	const StepCtx& step = w->step;
	// NOT w->awakeIslands.count: the island split (which runs beside or before these stages) destroys one island and
	// creates its pieces; the buckets below describe the islands as they were when islandPartition ran
	const int islandCount = step.islandSolveCount;
	const int awakeBodyCount = step.awakeBodyCount;
	const int slotCount = step.awakeContactCount;
	const int32_t* slotOff = ptr( w, w->islSlotOff );
	const int32_t* bodyOff = ptr( w, w->islBodyOff );
	const int32_t* colorOff = ptr( w, w->islColorOff );
	const int32_t* islSlots = ptr( w, w->islSlots );
	const int32_t* islBodies = ptr( w, w->islBodies );
	BodyState* states = ptr( w, w->states );
	const ConView c = conView( w );
	const float h = step.h;
	const float inv_h = step.inv_h;
	const float contactSpeed = w->contactSpeed;
	const float maxLinearSpeed = step.maxLinearVelocity;
	const float maxAngularSpeed = kMaxRotation * step.inv_dt;
	const int subStepCount = step.subStepCount;
	const int lane = t.lane(), lanes = t.groupSize();
	for ( int i = t.groupIndex(); i < islandCount; i += t.groupCount() )
	{
		const int b0 = bodyOff[i];
		const int b1 = i + 1 < islandCount ? bodyOff[i + 1] : awakeBodyCount;
		const int32_t* cOff = colorOff + i * kColorCount;
		const int sEnd = i + 1 < islandCount ? slotOff[i + 1] : slotCount;
		for ( int sub = 0; sub <= subStepCount; ++sub )
		{
			const bool restitutionPass = sub == subStepCount;
			if ( restitutionPass == false )
			{
				for ( int k = b0 + lane; k < b1; k += lanes )
					integrateVelocity( w, islBodies[k], h, maxLinearSpeed, maxAngularSpeed );
				t.groupSync();
			}
			// stages: warm start, solve | integrate positions | relax  -- or restitution alone after the last sub-step
			for ( int stage = restitutionPass ? 3 : 0; stage <= 3; ++stage )
			{
				if ( restitutionPass == false && stage == 3 )
					break;
				if ( stage == 2 )
				{
					for ( int k = b0 + lane; k < b1; k += lanes )
						integratePosition( states[islBodies[k]], h );
					t.groupSync();
				}
				for ( int color = 0; color < kOverflow; ++color )
				{
					const int s0 = cOff[color];
					const int s1 = color + 1 < kColorCount ? cOff[color + 1] : sEnd;
					if ( s0 == s1 )
						continue;
					for ( int k = s0 + lane; k < s1; k += lanes )
					{
						int slot = islSlots[k];
						if ( stage == 0 )
							warmStartSlot( c, slot, states );
						else if ( stage == 1 )
							solveSlot( c, slot, states, true, inv_h, contactSpeed );
						else if ( stage == 2 )
							solveSlot( c, slot, states, false, inv_h, contactSpeed );
						else
						{
							// 4-lane group rule of the reference (contact_solver.c:1982-1986), by index in the global colour
							int base = step.colorBase[color];
							int count = step.colorBase[color + 1] - base;
							int g0 = ( ( slot - base ) >> 2 ) << 2;
							bool any = false;
							for ( int q = g0; q < g0 + 4 && q < count; ++q )
								any = any || ( c.f( cfRestitution, base + q ) != 0.0f );
							if ( any )
								restitutionSlot( c, slot, states, step.restitutionThreshold );
						}
					}
					t.groupSync();
				}
			}
		}
	}
	t.sync();
}

// ------------------------------------------------------------------------------------------------ island split
// b2SplitIsland (island.c:602-840) is a depth-first walk whose visiting order defines the new islands' ids and the
// order of their body / contact / joint lists, so the walk itself stays serial. Everything around it is data-parallel:
//   prepare  list-rank the island's body list (pointer doubling) to get the seed order without walking the list, and
//            flatten every body's contact-edge and joint-edge lists into rows of (constraint id, other body position);
//   walk     one thread runs the reference's DFS over those dense rows (marks and stack are small dense arrays, so
//            every hop is an L1 hit instead of a pointer chase through 68..136-byte records) and only records the
//            visiting order; it also destroys the old island and creates the new ones, in DFS order;
//   apply    the whole team writes island ids, prev / next links, heads, tails and counts from the recorded order.
struct SplitView
{
	int32_t *next[2], *dist[2]; // list ranking, by body id
	int32_t* posOf;				// by body id: position in the island's body list
	int32_t* bodyAt;			// by position: body id
	int32_t *rowOff, *jrowOff;	// [n + 1] by position
	int32_t *edges, *jedges;	// (constraint id, other) pairs
	int32_t *contactMark, *jointMark, *bodyMark;
	int32_t *bodyOrder, *bodyComp, *contactOrder, *contactComp, *jointOrder, *jointComp;
	int32_t *compIsland, *stack;
};

// `n` = bodies of the island that is split. The arrays the serial walk reads and re-reads (edge rows, marks, stack) go
// into the team's shared-memory arena when the team has one and they fit (a single small world stepped by one block:
// the walk is then a chain of shared-memory hits instead of L1 / L2 round trips that the solver stages running beside
// it keep evicting); the visiting order it records, and everything else, stays in the image.
F2D_HD SplitView splitView( World* w, int32_t* arena, int arenaInts, int n )
{
	const int B = w->bodies.cap + 8, C = w->contacts.cap + 8, J = w->joints.cap + 8;
	int32_t* p = ptr( w, w->splitScratch );
	SplitView v;
	auto take = [&]( int n ) {
		int32_t* r = p;
		p += n;
		return r;
	};
	v.next[0] = take( B );
	v.next[1] = take( B );
	v.dist[0] = take( B );
	v.dist[1] = take( B );
	v.posOf = take( B );
	v.bodyAt = take( B );
	v.rowOff = take( B );
	v.jrowOff = take( B );
	v.bodyMark = take( B );
	v.bodyOrder = take( B );
	v.bodyComp = take( B );
	v.compIsland = take( B );
	v.stack = take( B );
	take( B );
	v.edges = take( 4 * C );
	v.contactMark = take( C );
	v.contactOrder = take( C );
	v.contactComp = take( C );
	v.jedges = take( 4 * J );
	v.jointMark = take( J );
	v.jointOrder = take( J );
	v.jointComp = take( J );
	if ( arena != nullptr )
	{
		// rows hold two ints per edge, a touching contact has at most two edges inside the island; marks are by id
		const int rows = n + 2, edgeInts = 4 * w->step.awakeContactCount + 8, contactIds = w->contactIds.next + 8;
		const int jedgeInts = 4 * w->jointIds.next + 8, jointIds = w->jointIds.next + 8;
		const int need = 4 * rows + edgeInts + contactIds + jedgeInts + jointIds;
		if ( need <= arenaInts )
		{
			int32_t* a = arena;
			auto carve = [&]( int count ) {
				int32_t* r = a;
				a += count;
				return r;
			};
			v.rowOff = carve( rows );
			v.jrowOff = carve( rows );
			v.bodyMark = carve( rows );
			v.stack = carve( rows );
			v.edges = carve( edgeInts );
			v.contactMark = carve( contactIds );
			v.jedges = carve( jedgeInts );
			v.jointMark = carve( jointIds );
		}
	}
	return v;
}
F2D_HD SplitView splitView( World* w ) { return splitView( w, nullptr, 0, 0 ); }

// Second word of a row entry: the other body's code (-2 disabled, -1 none / static, else its position) biased by 2, plus
// two flags that the walk would otherwise have to work out per pop with warp votes:
//   kRowFirstOther  no earlier edge of this row leads to the same body. Among the edges of a popped body that lead to an
//                   UNVISITED body X all are unmarked (a marked constraint between the two was processed when X was
//                   popped, and popped bodies are visited), so the one that pushes X - the first in row order - is known
//                   when the row is built;
//   kRowFirstId     first occurrence of the constraint in this row (a joint with both ends on one body is listed twice).
constexpr int kRowFirstOther = 1 << 28, kRowFirstId = 1 << 29, kRowCodeMask = ( 1 << 28 ) - 1;
F2D_HD int splitRowEntry( const int32_t* rows, int rowStart, int at, int id, int code )
{
	int flags = kRowFirstOther | kRowFirstId;
	for ( int k = rowStart; k < at; ++k )
	{
		if ( rows[2 * k] == id )
			flags &= ~kRowFirstId;
		if ( ( rows[2 * k + 1] & kRowCodeMask ) == code + 2 )
			flags &= ~kRowFirstOther;
	}
	return ( code + 2 ) | flags;
}

// Team-wide preparation. Leaves step.splitBodies == 0 when there is nothing to split (island.c:610-620).
template <class Team> F2D_HDF inline void splitPrepare( World* w, Team& t )
{
	const int baseId = w->step.splitTarget;
	if ( baseId == kNull )
		return;
	const Island& base = ptr( w, w->islands )[baseId];
	if ( base.setIndex != kAwakeSet || base.constraintRemoveCount == 0 )
		return;
	const int n = base.bodyCount;
	SplitView v = splitView( w, t.arenaPtr(), t.arenaSize(), n );
	Body* bodies = ptr( w, w->bodies );
	const Contact* contacts = ptr( w, w->contacts );
	const Joint* joints = ptr( w, w->joints );
	const int32_t* awakeBodies = ptr( w, w->awakeBodies );
	const int awakeCount = w->awakeBodies.count;
	constexpr int kNotMember = -2;

	// List ranking of the body list (dist = number of bodies after this one) by pointer doubling on dense arrays
	// indexed by AWAKE INDEX: the body records are chased once, here, and never again during the rounds.
	for ( int i = t.rank(); i < awakeCount; i += t.size() )
	{
		const Body& b = bodies[awakeBodies[i]];
		if ( b.islandId != baseId )
		{
			v.next[0][i] = kNotMember;
			v.next[1][i] = kNotMember;
			continue;
		}
		int nx = b.islandNext;
		v.next[0][i] = nx != kNull ? bodies[nx].localIndex : kNull; // island mates are awake: localIndex = awake index
		v.dist[0][i] = nx != kNull ? 1 : 0;
	}
	t.sync();
	int cur = 0;
	for ( int span = 1; span < n; span <<= 1 )
	{
		for ( int i = t.rank(); i < awakeCount; i += t.size() )
		{
			int nx = v.next[cur][i];
			if ( nx == kNotMember )
				continue;
			int d = v.dist[cur][i];
			if ( nx != kNull )
			{
				d += v.dist[cur][nx];
				nx = v.next[cur][nx];
			}
			v.next[cur ^ 1][i] = nx;
			v.dist[cur ^ 1][i] = d;
		}
		t.sync();
		cur ^= 1;
	}
	// rows: touching contact edges and joint edges per body, in edge-list order
	for ( int i = t.rank(); i < awakeCount; i += t.size() )
	{
		if ( v.next[cur][i] == kNotMember )
			continue;
		int id = awakeBodies[i];
		const Body& body = bodies[id];
		int pos = n - 1 - v.dist[cur][i];
		v.posOf[id] = pos;
		v.bodyAt[pos] = id;
		v.bodyMark[pos] = 0;
		int count = 0;
		for ( int key = body.headContactKey; key != kNull; )
		{
			const Contact& c = contacts[key >> 1];
			if ( c.flags & kContactTouching )
				count += 1;
			key = c.edges[key & 1].nextKey;
		}
		v.rowOff[pos] = count;
		int jcount = 0;
		for ( int key = body.headJointKey; key != kNull; key = joints[key >> 1].edges[key & 1].nextKey )
			jcount += 1;
		v.jrowOff[pos] = jcount;
	}
	t.sync();
	int edgeTotal = t.exclusiveScan( v.rowOff, n );
	int jedgeTotal = t.exclusiveScan( v.jrowOff, n );
	if ( 2 * edgeTotal > 4 * ( w->contacts.cap + 8 ) || 2 * jedgeTotal > 4 * ( w->joints.cap + 8 ) )
	{
		if ( t.rank() == 0 )
			setError( w, kErrCapacity, __LINE__ );
		return;
	}
	if ( t.rank() == 0 )
	{
		v.rowOff[n] = edgeTotal;
		v.jrowOff[n] = jedgeTotal;
	}
	for ( int i = t.rank(); i < awakeCount; i += t.size() )
	{
		if ( v.next[cur][i] == kNotMember )
			continue;
		int id = awakeBodies[i];
		const Body& body = bodies[id];
		int pos = v.posOf[id];
		int out = v.rowOff[pos];
		const int row0 = out;
		for ( int key = body.headContactKey; key != kNull; )
		{
			int contactId = key >> 1;
			int edgeIndex = key & 1;
			const Contact& c = contacts[contactId];
			key = c.edges[edgeIndex].nextKey;
			if ( ( c.flags & kContactTouching ) == 0 )
				continue;
			int otherId = c.edges[edgeIndex ^ 1].bodyId;
			const Body& other = bodies[otherId];
			// the other body as a POSITION in the island's body list (posOf of every member was written before the scans);
			// after the island merge both bodies of a touching contact are in this island unless one is static
			const int code = ( other.setIndex != kStaticSet && other.islandId == baseId ) ? v.posOf[otherId] : kNull;
			v.edges[2 * out] = contactId;
			v.edges[2 * out + 1] = splitRowEntry( v.edges, row0, out, contactId, code );
			v.contactMark[contactId] = 0;
			out += 1;
		}
		int jout = v.jrowOff[pos];
		const int jrow0 = jout;
		for ( int key = body.headJointKey; key != kNull; )
		{
			int jointId = key >> 1;
			int edgeIndex = key & 1;
			const Joint& j = joints[jointId];
			key = j.edges[edgeIndex].nextKey;
			int otherId = j.edges[edgeIndex ^ 1].bodyId;
			const Body& other = bodies[otherId];
			int code = other.setIndex == kDisabledSet ? -2 : ( ( other.setIndex == kAwakeSet && other.islandId == baseId ) ? v.posOf[otherId] : kNull );
			v.jedges[2 * jout] = jointId;
			v.jedges[2 * jout + 1] = splitRowEntry( v.jedges, jrow0, jout, jointId, code );
			v.jointMark[jointId] = 0;
			jout += 1;
		}
	}
	t.sync();
	if ( t.rank() == 0 )
		w->step.splitBodies = n;
	t.sync();
}

// The depth-first walk. Order of everything follows island.c:669-833: the visiting order is serial by nature (it
// defines the new islands' ids and list orders), but the edges of ONE popped body are independent tests against the
// marks, so the lanes of a warp take one edge each: an edge is new iff its constraint is unmarked (ballot -> ranks
// give the order slots), and among the new edges the first one (lowest lane) that leads to an unmarked body pushes it.
// Every lane keeps the same counters; lane 0 performs the island pool edits.
template <class Lanes> F2D_HDF inline void splitWalk( World* w, Lanes L, int32_t* arena, int arenaInts )
{
	const int n = w->step.splitBodies;
	if ( n == 0 )
		return;
	const int baseId = w->step.splitTarget;
	SplitView v = splitView( w, arena, arenaInts, n );
	const int lane = L.lane();
	const uint32_t below = lane == 0 ? 0u : ( 0xffffffffu >> ( 32 - lane ) );
	if ( lane == 0 )
		destroyIsland( w, baseId );
	int comp = 0, nb = 0, nc = 0, nj = 0;
	const int passes = v.jrowOff[n] > 0 ? 2 : 1; // an island without joints: no (empty) joint rows to look up per body
	for ( int seed = 0; seed < n; ++seed )
	{
		if ( v.bodyMark[seed] )
			continue;
		if ( lane == 0 )
		{
			v.stack[0] = seed;
			v.bodyMark[seed] = 1;
			v.compIsland[comp] = createIsland( w, kAwakeSet );
		}
		int sp = 1;
		// The entry on top of the stack is kept in a register (`top`, uniform across the lanes; kNull = not cached): the
		// body popped next is nearly always one pushed a moment ago, and reading it back from the stack array would add
		// one more dependent load to the four a pop already costs.
		int top = seed;
		L.sync();
		while ( sp > 0 )
		{
			sp -= 1;
			int pos = top != kNull ? top : v.stack[sp];
			top = kNull;
			if ( lane == 0 )
			{
				v.bodyOrder[nb] = pos;
				v.bodyComp[nb] = comp;
			}
			nb += 1;
			for ( int pass = 0; pass < passes; ++pass )
			{
				// pass 0: contact edges, pass 1: joint edges (island.c:700-767 then :769-812)
				const int32_t* rowOff = pass == 0 ? v.rowOff : v.jrowOff;
				const int32_t* rows = pass == 0 ? v.edges : v.jedges;
				int32_t* mark = pass == 0 ? v.contactMark : v.jointMark;
				int32_t* order = pass == 0 ? v.contactOrder : v.jointOrder;
				int32_t* orderComp = pass == 0 ? v.contactComp : v.jointComp;
				int& filled = pass == 0 ? nc : nj;
				const int end = rowOff[pos + 1];
				for ( int chunk = rowOff[pos]; chunk < end; chunk += L.count() )
				{
					const int e = chunk + lane;
					const bool live = e < end;
					const int id = live ? rows[2 * e] : 0;
					const int entry = live ? rows[2 * e + 1] : 1; // (1 = code -1, no flags)
					const int other = ( entry & kRowCodeMask ) - 2;
					// both marks are asked for together: each is one dependent hop behind the row entry
					const bool unmarked = live && mark[id] == 0;
					const bool unvisited = other >= 0 && v.bodyMark[other] == 0;
					// a constraint listed twice in one row counts once, at its first edge (see splitRowEntry)
					const bool fresh = unmarked && ( entry & kRowFirstId ) != 0;
					if ( fresh )
						mark[id] = 1;

					// joints to a disabled body are marked but neither recorded nor followed (island.c:785-789)
					const bool recorded = fresh && other != -2;
					const uint32_t recordedMask = L.ballot( recorded );
					if ( recorded )
					{
						int slot = filled + popCount32( recordedMask & below );
						order[slot] = id;
						orderComp[slot] = comp;
					}
					filled += popCount32( recordedMask );
					// several edges may lead to one body: the first of the row pushes it (see splitRowEntry)
					const bool pusher = recorded && unvisited && ( entry & kRowFirstOther ) != 0;
					const uint32_t pushMask = L.ballot( pusher );
					if ( pusher )
					{
						v.stack[sp + popCount32( pushMask & below )] = other;
						v.bodyMark[other] = 1;
						// its edge rows are read when it is popped: ask for them now (unless they are in shared memory)
						if ( arena == nullptr )
						{
							prefetchLine( v.rowOff + other );
							prefetchLine( v.edges + 2 * v.rowOff[other] );
						}
					}
					if ( pushMask != 0 )
					{
						sp += popCount32( pushMask );
						top = L.from( other, 31 - countLeadingZeros32( pushMask ) ); // the last one pushed
					}
					L.sync();
				}
			}
		}
		comp += 1;
	}
	if ( lane == 0 )
	{
		w->step.splitContacts = nc;
		w->step.splitJoints = nj;
		w->step.splitComponents = comp;
	}
}

// Team-wide: island ids, list links, heads, tails and counts from the recorded visiting order.
template <class Team> F2D_HDF inline void splitApply( World* w, Team& t )
{
	const int n = w->step.splitBodies;
	if ( n == 0 )
		return;
	SplitView v = splitView( w, t.arenaPtr(), t.arenaSize(), n );
	Island* islands = ptr( w, w->islands );
	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	for ( int j = t.rank(); j < n; j += t.size() )
	{
		int k = v.bodyComp[j];
		Island& is = islands[v.compIsland[k]];
		int id = v.bodyAt[v.bodyOrder[j]];
		bool first = j == 0 || v.bodyComp[j - 1] != k;
		bool last = j == n - 1 || v.bodyComp[j + 1] != k;
		Body& b = bodies[id];
		b.islandId = is.islandId;
		b.islandPrev = first ? kNull : v.bodyAt[v.bodyOrder[j - 1]];
		b.islandNext = last ? kNull : v.bodyAt[v.bodyOrder[j + 1]];
		if ( first )
			is.headBody = id;
		if ( last )
			is.tailBody = id;
		atomAdd( &is.bodyCount, 1 );
	}
	const int nc = w->step.splitContacts;
	for ( int j = t.rank(); j < nc; j += t.size() )
	{
		int k = v.contactComp[j];
		Island& is = islands[v.compIsland[k]];
		int id = v.contactOrder[j];
		bool first = j == 0 || v.contactComp[j - 1] != k;
		bool last = j == nc - 1 || v.contactComp[j + 1] != k;
		Contact& c = contacts[id];
		c.islandId = is.islandId;
		c.islandPrev = first ? kNull : v.contactOrder[j - 1];
		c.islandNext = last ? kNull : v.contactOrder[j + 1];
		if ( first )
			is.headContact = id;
		if ( last )
			is.tailContact = id;
		atomAdd( &is.contactCount, 1 );
	}
	const int nj = w->step.splitJoints;
	for ( int j = t.rank(); j < nj; j += t.size() )
	{
		int k = v.jointComp[j];
		Island& is = islands[v.compIsland[k]];
		int id = v.jointOrder[j];
		bool first = j == 0 || v.jointComp[j - 1] != k;
		bool last = j == nj - 1 || v.jointComp[j + 1] != k;
		Joint& jt = joints[id];
		jt.islandId = is.islandId;
		jt.islandPrev = first ? kNull : v.jointOrder[j - 1];
		jt.islandNext = last ? kNull : v.jointOrder[j + 1];
		if ( first )
			is.headJoint = id;
		if ( last )
			is.tailJoint = id;
		atomAdd( &is.jointCount, 1 );
	}
	t.sync();
}

// Prepare, the sub-step loop, restitution and impulse storage (solver.c:929-1106) on `t`: the whole team, or the crew
// left over when a side worker splits an island at the same time.
template <class Team> F2D_HDF inline void solveStages( World* w, Team& t )
{
	const int awakeBodyCount = w->step.awakeBodyCount;
	// prepare joints and contacts (all colours incl. overflow: prepare formulas are identical)
	{
		JointSim* jsims = ptr( w, w->jointSims );
		for ( int color = 0; color < kColorCount; ++color )
		{
			const int32_t* jl = ptr( w, w->colorJoints[color] );
			int n = w->colorJoints[color].count;
			for ( int i = t.rank(); i < n; i += t.size() )
				prepareJoint( w, jsims[jl[i]] );
		}
		{
			const float h = w->step.h;
			const int32_t* awakeBodies = ptr( w, w->awakeBodies );
			const BodySim* bsims = ptr( w, w->sims );
			for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
			{
				if ( i + t.size() < awakeBodyCount )
				{
					const BodySim& n = bsims[awakeBodies[i + t.size()]];
					prefetchL2( &n.force );
					prefetchL2( &n.gravityScale );
				}
				prepareIntegrate( w, i, h );
			}
		}
		const ConView c = conView( w );
		const BodyState* states = ptr( w, w->states );
		const ContactSim* csims = ptr( w, w->contactSims );
		float warmStartScale = w->enableWarmStarting ? 1.0f : 0.0f;
		int total = w->step.awakeContactCount;
		int color = -1, colorStart = 0, colorEnd = 0;
		const int32_t* list = nullptr;
		for ( int slot = t.rank(); slot < total; slot += t.size() )
		{
			while ( slot >= colorEnd )
			{
				color += 1;
				colorStart = colorEnd;
				colorEnd = w->step.colorBase[color + 1];
				list = ptr( w, w->colorContacts[color] );
			}
			// request the record of this thread's next slot while this one is prepared (same colour only: cheap to find)
			if ( slot + t.size() < colorEnd )
			{
				const ContactSim& n = csims[list[slot + t.size() - colorStart]];
				prefetchL2( &n );
				prefetchL2( &n.manifold.point0 );
			}
			prepareContactSlot( w, c, slot, list[slot - colorStart], states, warmStartScale );
		}
	}
	t.sync();
	F2D_MARK( w, t, pfPrepare );

	if ( w->step.islandPath )
	{
		islandSolve( w, t );
		F2D_MARK( w, t, pfSolve );
	}
	else
	{
		const float h = w->step.h;
		const float maxLinearSpeed = w->step.maxLinearVelocity;
		const float maxAngularSpeed = kMaxRotation * w->step.inv_dt;
		const int subStepCount = w->step.subStepCount;
		for ( int sub = 0; sub < subStepCount; ++sub )
		{
			for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
				integrateVelocity( w, i, h, maxLinearSpeed, maxAngularSpeed );
			t.sync();
			F2D_MARK( w, t, pfIntegrateVel );
			constraintPass( w, t, 0 );
			F2D_MARK( w, t, pfWarmStart );
			constraintPass( w, t, 1 );
			F2D_MARK( w, t, pfSolve );
			{
				BodyState* states = ptr( w, w->states );
				for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
					integratePosition( states[i], h );
			}
			t.sync();
			F2D_MARK( w, t, pfIntegratePos );
			constraintPass( w, t, 2 );
			F2D_MARK( w, t, pfRelax );
		}
		constraintPass( w, t, 3 );
		F2D_MARK( w, t, pfRestitution );
	}

	// store impulses (solver.c:1093-1097)
	{
		const ConView c = conView( w );
		int total = w->step.awakeContactCount;
		int overflowBase = w->step.colorBase[kOverflow];
		int color = -1, colorStart = 0, colorEnd = 0;
		const int32_t* list = nullptr;
		for ( int slot = t.rank(); slot < total; slot += t.size() )
		{
			while ( slot >= colorEnd )
			{
				color += 1;
				colorStart = colorEnd;
				colorEnd = w->step.colorBase[color + 1];
				list = ptr( w, w->colorContacts[color] );
			}
			storeSlot( w, c, slot, list[slot - colorStart], slot >= overflowBase );
		}
	}
	t.sync();
	F2D_MARK( w, t, pfStore );
}

// island.c:539-598. Pass 1 (every awake island ends up pointing straight at its root) is a read-only walk per island
// followed by one write; pass 2 (the order-defining list concatenations, last island first) stays serial but only runs
// when some island actually has a parent.
template <class Team> F2D_HDF inline void mergeAwakeIslandsTeam( World* w, Team& t )
{
	Island* islands = ptr( w, w->islands );
	const int32_t* awake = ptr( w, w->awakeIslands );
	int count = w->awakeIslands.count;
	int32_t* roots = ptr( w, w->scan );
	if ( count > w->scan.cap )
	{
		if ( t.rank() == 0 )
			mergeAwakeIslands( w );
		t.sync();
		return;
	}
	if ( t.rank() == 0 )
		w->step.mergeCount = 0;
	t.sync();
	for ( int i = t.rank(); i < count; i += t.size() )
	{
		int rootId = awake[i];
		while ( islands[rootId].parentIsland != kNull )
			rootId = islands[rootId].parentIsland;
		roots[i] = rootId;
		if ( rootId != awake[i] )
			atomAdd( &w->step.mergeCount, 1 );
	}
	t.sync();
	if ( w->step.mergeCount == 0 )
		return;
	for ( int i = t.rank(); i < count; i += t.size() )
	{
		if ( roots[i] != awake[i] )
			islands[awake[i]].parentIsland = roots[i];
	}
	t.sync();
	// Relabel (island.c:436-471) without walking the child islands' lists: every awake body, every touching contact
	// (= every entry of a colour array) and every awake joint looks up whether its island now has a parent.
	{
		Body* bodies = ptr( w, w->bodies );
		Contact* contacts = ptr( w, w->contacts );
		Joint* joints = ptr( w, w->joints );
		const int32_t* awakeBodies = ptr( w, w->awakeBodies );
		int bodyCount = w->awakeBodies.count;
		for ( int i = t.rank(); i < bodyCount; i += t.size() )
		{
			Body& b = bodies[awakeBodies[i]];
			if ( b.islandId != kNull && islands[b.islandId].parentIsland != kNull )
				b.islandId = islands[b.islandId].parentIsland;
		}
		for ( int color = 0; color < kColorCount; ++color )
		{
			const int32_t* cl = ptr( w, w->colorContacts[color] );
			int n = w->colorContacts[color].count;
			for ( int i = t.rank(); i < n; i += t.size() )
			{
				Contact& c = contacts[cl[i]];
				if ( c.islandId != kNull && islands[c.islandId].parentIsland != kNull )
					c.islandId = islands[c.islandId].parentIsland;
			}
			const int32_t* jl = ptr( w, w->colorJoints[color] );
			int nj = w->colorJoints[color].count;
			for ( int i = t.rank(); i < nj; i += t.size() )
			{
				Joint& j = joints[jl[i]];
				if ( j.islandId != kNull && islands[j.islandId].parentIsland != kNull )
					j.islandId = islands[j.islandId].parentIsland;
			}
		}
	}
	t.sync();
	// ... and rank 0 splices the lists onto the roots, last island first (order-defining). The merging islands are
	// compacted first so the serial loop only visits them. destroyIsland swap-removes from awakeIslands, which only
	// moves entries from the tail, i.e. islands this descending loop has already passed.
	int32_t* flags = ptr( w, w->islBodyFill );
	int32_t* merging = ptr( w, w->islSlotOff );
	if ( count > w->islBodyFill.cap || count > w->islSlotOff.cap )
	{
		if ( t.rank() == 0 )
			setError( w, kErrCapacity, __LINE__ );
		return;
	}
	for ( int i = t.rank(); i < count; i += t.size() )
		flags[i] = roots[i] != awake[i] ? 1 : 0;
	t.sync();
	int mergeTotal = t.exclusiveScan( flags, count );
	for ( int i = t.rank(); i < count; i += t.size() )
	{
		if ( roots[i] != awake[i] )
			merging[flags[i]] = awake[i];
	}
	t.sync();
	if ( t.rank() == 0 )
	{
		for ( int k = mergeTotal - 1; k >= 0; --k )
		{
			int islandId = merging[k];
			mergeIsland( w, islands[islandId], false );
			destroyIsland( w, islandId );
		}
	}
	t.sync();
}

template <class Team> F2D_HDF inline void stepSolve( World* w, Team& t )
{
	mergeAwakeIslandsTeam( w, t );
	{
		// size of the largest awake island (decides between the island-parallel and the colour-parallel solve)
		if ( t.rank() == 0 )
		{
			w->step.maxIslandContacts = 0;
			w->step.maxIslandBodies = 0;
		}
		t.sync();
		const Island* islands = ptr( w, w->islands );
		const int32_t* awake = ptr( w, w->awakeIslands );
		int count = w->awakeIslands.count;
		int maxC = 0, maxB = 0;
		for ( int i = t.rank(); i < count; i += t.size() )
		{
			const Island& is = islands[awake[i]];
			maxC = maxi( maxC, is.contactCount );
			maxB = maxi( maxB, is.bodyCount );
		}
		if ( maxC > 0 )
			atomMax32( &w->step.maxIslandContacts, maxC );
		if ( maxB > 0 )
			atomMax32( &w->step.maxIslandBodies, maxB );
		t.sync();
	}
	if ( t.rank() == 0 )
	{
		w->stepIndex += 1;
		StepCtx& s = w->step;
		s.splitTarget = w->splitIslandId;
		s.splitBodies = 0;
		s.awakeBodyCount = w->awakeBodies.count;
		// colour bases and the active colour list (solver.c:1237-1355); overflow keeps the last base
		int base = 0;
		s.activeColorCount = 0;
		s.awakeJointCount = 0;
		for ( int i = 0; i < kColorCount; ++i )
		{
			s.colorBase[i] = base;
			base += w->colorContacts[i].count;
			if ( i < kOverflow )
			{
				if ( w->colorContacts[i].count + w->colorJoints[i].count > 0 )
					s.activeColors[s.activeColorCount++] = i;
			}
			s.awakeJointCount += w->colorJoints[i].count;
		}
		s.colorBase[kColorCount] = base;
		s.awakeContactCount = base;
		// many small islands and nothing that needs a serial order: solve island by island (islandSolve)
		int overflowCount = w->colorContacts[kOverflow].count + w->colorJoints[kOverflow].count;
		s.islandPath = ( s.awakeJointCount == 0 && overflowCount == 0 && w->awakeIslands.count >= kIslandPathMinIslands &&
						 s.maxIslandContacts <= kIslandPathMaxContacts && s.maxIslandBodies <= kIslandPathMaxBodies &&
						 w->awakeIslands.count + 1 <= w->islSlotOff.cap )
						   ? 1
						   : 0;
		if ( base > w->consStride )
			setError( w, kErrCapacity, __LINE__ );
		if ( s.awakeBodyCount > w->moveEvents.cap )
			setError( w, kErrCapacity, __LINE__ );
		else if ( s.awakeBodyCount > 0 )
			w->moveEvents.count = s.awakeBodyCount;
	}
	t.sync();
	const int awakeBodyCount = w->step.awakeBodyCount;
	F2D_MARK( w, t, pfSolveSetup );
	if ( awakeBodyCount == 0 || ( w->error & kErrCapacity ) != 0 )
		return;

	if ( w->step.islandPath )
		islandPartition( w, t );

	// Island split (solver.c:1473-1485, 1700-1706). The serial depth-first walk touches island bookkeeping only, so, as
	// in the reference, it runs concurrently with the solver stages when the team can spare a side worker for it.
	splitPrepare( w, t );
	bool fork = false;
	if constexpr ( Team::kCanFork )
		fork = w->step.splitBodies > 0 && t.canFork();
	// (small teams - batches - keep the vote inside the body loop: the walk is hidden there anyway, and a second sweep
	// over the body records would cost them DRAM traffic)
	const bool overlapFinalize = fork && t.size() >= 512 && w->step.dt > 0.0f;
	if ( fork )
	{
		if constexpr ( Team::kCanFork )
		{
			if ( t.inSide() )
			{
				if ( t.inSideGroup() )
					splitWalk( w, typename Team::Lanes{}, t.arenaPtr(), t.arenaSize() );
			}
			else
			{
				auto crew = t.crew();
				solveStages( w, crew );
				// A large team is done with the stages well before the serial walk ends: it goes on with the body loop
				// of finalize, which needs nothing of the islands except the votes (cast after the join, stepFinalize).
				// The reference cannot do this ("cannot split islands in parallel with FinalizeBodies", solver.c:1478):
				// its finalize task reads the island of every body.
				if ( overlapFinalize )
					finalizeBodiesTeam( w, crew, false );
			}
		}
		t.sync();
		if ( overlapFinalize && t.rank() == 0 )
			w->step.bodiesFinalized = 1;
	}
	else
	{
		if ( t.inFirstGroup() )
			splitWalk( w, typename Team::Lanes{}, t.arenaPtr(), t.arenaSize() );
		solveStages( w, t );
		t.sync();
	}
	F2D_MARK( w, t, pfSplitJoin );
	splitApply( w, t );
	t.sync();
	F2D_MARK( w, t, pfSplitApply );
	if ( t.rank() == 0 )
		w->splitIslandId = kNull;
}

// ------------------------------------------------------------------------------------------------ finalize
struct ContinuousCtx
{
	World* w;
	BodySim* fastSim;
	const Shape* fastShape;
	V2 centroid1, centroid2;
	Sweep sweep;
	float fraction;
};

F2D_HD Sweep makeSweep( const BodySim& s )
{
	Sweep sw;
	sw.c1 = s.center0;
	sw.c2 = s.center;
	sw.q1 = s.rotation0;
	sw.q2 = s.transform.q;
	sw.localCenter = s.localCenter;
	return sw;
}

// solver.c:212-387 (custom filter / pre-solve callbacks are rejected at registration on the device path)
// b2CustomFilterFcn and b2PreSolveFcn inside the continuous pass: set by the callback-mediated step while the HOST runs
// solveContinuous for the deferred fast bodies (f2d_capi.inl stepWithHostCallbacks); never set on the device.
struct HostContinuousHooks
{
	bool ( *filter )( void* context, int shapeId, int fastShapeId ) = nullptr;					   // solver.c:271-282
	bool ( *preSolve )( void* context, int shapeId, int fastShapeId, Manifold* manifold ) = nullptr; // solver.c:366-379
	void* context = nullptr;
};
static HostContinuousHooks g_hostContinuous; // host code only

F2D_HDF inline bool continuousVisit( ContinuousCtx& ctx, int shapeId )
{
	World* w = ctx.w;
	const Shape& fastShape = *ctx.fastShape;
	if ( shapeId == fastShape.id )
		return true;
	const Shape& shape = ptr( w, w->shapes )[shapeId];
	if ( shape.bodyId == fastShape.bodyId )
		return true;
	if ( shape.sensorIndex != kNull )
		return true;
	if ( shouldShapesCollide( fastShape.filter, shape.filter ) == false )
		return true;
	const Body* bodies = ptr( w, w->bodies );
	const Body& body = bodies[shape.bodyId];
	const BodySim& bodySim = ptr( w, w->sims )[shape.bodyId];
	if ( bodySim.isBullet )
		return true;
	const Body& fastBody = bodies[fastShape.bodyId];
	if ( shouldBodiesCollide( w, fastBody, body ) == false )
		return true;
#if !defined( __CUDA_ARCH__ )
	if ( g_hostContinuous.filter != nullptr && g_hostContinuous.filter( g_hostContinuous.context, shapeId, fastShape.id ) == false )
		return true;
#endif

	if ( shape.type == kChainSegment )
	{
		Xf transform = bodySim.transform;
		V2 p1 = xfPoint( transform, shape.chainSegment.segment.p1 );
		V2 p2 = xfPoint( transform, shape.chainSegment.segment.p2 );
		V2 e = sub( p2, p1 );
		float len;
		e = lengthAndNormalize( &len, e );
		if ( len > kLinearSlop )
		{
			float offset1 = cross( sub( ctx.centroid1, p1 ), e );
			float offset2 = cross( sub( ctx.centroid2, p1 ), e );
			const float allowedFraction = 0.25f;
			if ( offset1 < 0.0f || offset1 - offset2 < allowedFraction * ctx.fastSim->minExtent )
				return true;
		}
	}

	ShapeProxy proxyA = makeShapeProxy( shape );
	ShapeProxy proxyB = makeShapeProxy( fastShape );
	Sweep sweepA = makeSweep( bodySim );
	float hitFraction = ctx.fraction;
	bool didHit = false;
	float toi = timeOfImpact( proxyA, proxyB, sweepA, ctx.sweep, ctx.fraction );
	if ( 0.0f < toi && toi < ctx.fraction )
	{
		hitFraction = toi;
		didHit = true;
	}
	else if ( 0.0f == toi )
	{
		// fallback to TOI of a small circle around the fast shape centroid
		V2 centroid = shapeCentroid( fastShape );
		float mn, mx;
		shapeExtent( fastShape, centroid, &mn, &mx );
		float radius = 0.25f * mn;
		ShapeProxy small = makeProxy( &centroid, 1, radius );
		toi = timeOfImpact( proxyA, small, sweepA, ctx.sweep, ctx.fraction );
		if ( 0.0f < toi && toi < ctx.fraction )
		{
			hitFraction = toi;
			didHit = true;
		}
	}
#if !defined( __CUDA_ARCH__ )
	if ( didHit && ( shape.enablePreSolveEvents || fastShape.enablePreSolveEvents ) && g_hostContinuous.preSolve != nullptr )
	{
		// the reference hands the callback a temporary manifold at the time of impact (b2ComputeManifold, contact.c:635-640)
		Xf transformA = sweepTransform( sweepA, hitFraction );
		Xf transformB = sweepTransform( ctx.sweep, hitFraction );
		SimplexCache cache;
		memset( &cache, 0, sizeof( cache ) );
		Manifold manifold = computeManifold( w, shape, transformA, fastShape, transformB, &cache );
		didHit = g_hostContinuous.preSolve( g_hostContinuous.context, shapeId, fastShape.id, &manifold );
	}
#endif
	if ( didHit )
		ctx.fraction = hitFraction;
	return true;
}

// solver.c:390-541
F2D_HDC inline void solveContinuous( World* w, int awakeIndex )
{
	int bodyId = ptr( w, w->awakeBodies )[awakeIndex];
	BodySim& fastSim = ptr( w, w->sims )[bodyId];
	const Body& fastBody = ptr( w, w->bodies )[bodyId];
	Shape* shapes = ptr( w, w->shapes );
	Sweep sweep = makeSweep( fastSim );
	Xf xf1;
	xf1.q = sweep.q1;
	xf1.p = sub( sweep.c1, rotate( sweep.q1, sweep.localCenter ) );
	Xf xf2;
	xf2.q = sweep.q2;
	xf2.p = sub( sweep.c2, rotate( sweep.q2, sweep.localCenter ) );

	ContinuousCtx ctx;
	ctx.w = w;
	ctx.sweep = sweep;
	ctx.fastSim = &fastSim;
	ctx.fraction = 1.0f;
	bool isBullet = fastSim.isBullet;

	int shapeId = fastBody.headShapeId;
	while ( shapeId != kNull )
	{
		Shape& fastShape = shapes[shapeId];
		shapeId = fastShape.nextShapeId;
		ctx.fastShape = &fastShape;
		ctx.centroid1 = xfPoint( xf1, fastShape.localCentroid );
		ctx.centroid2 = xfPoint( xf2, fastShape.localCentroid );
		Box box1 = fastShape.aabb;
		Box box2 = shapeAABB( fastShape, xf2 );
		Box box = boxUnion( box1, box2 );
		fastShape.aabb = box2; // no speculative margin here, as in the reference (solver.c:431-436)
		if ( fastShape.sensorIndex != kNull )
			continue;
		auto visit = [&]( int, uint64_t userData ) -> bool { return continuousVisit( ctx, (int)userData ); };
		treeQuery( w, w->trees[kStaticBody], box, UINT64_MAX, visit );
		if ( isBullet )
		{
			treeQuery( w, w->trees[kKinematicBody], box, UINT64_MAX, visit );
			treeQuery( w, w->trees[kDynamicBody], box, UINT64_MAX, visit );
		}
	}

	if ( ctx.fraction < 1.0f )
	{
		Rot q = nlerp( sweep.q1, sweep.q2, ctx.fraction );
		V2 c = lerp( sweep.c1, sweep.c2, ctx.fraction );
		V2 origin = sub( c, rotate( q, sweep.localCenter ) );
		Xf transform = { origin, q };
		fastSim.transform = transform;
		fastSim.center = c;
		fastSim.rotation0 = q;
		fastSim.center0 = c;
		ptr( w, w->moveEvents )[awakeIndex].transform = transform;

		shapeId = fastBody.headShapeId;
		while ( shapeId != kNull )
		{
			Shape& shape = shapes[shapeId];
			Box aabb = inflate( shapeAABB( shape, transform ), kSpeculative );
			shape.aabb = aabb;
			if ( boxContains( shape.fatAABB, aabb ) == false )
			{
				shape.fatAABB = inflate( aabb, kAabbMargin );
				shape.enlargedAABB = true;
				fastSim.enlargeAABB = true;
			}
			shapeId = shape.nextShapeId;
		}
	}
	else
	{
		fastSim.rotation0 = fastSim.transform.q;
		fastSim.center0 = fastSim.center;
		shapeId = fastBody.headShapeId;
		while ( shapeId != kNull )
		{
			Shape& shape = shapes[shapeId];
			if ( boxContains( shape.fatAABB, shape.aabb ) == false )
			{
				shape.fatAABB = inflate( shape.aabb, kAabbMargin );
				shape.enlargedAABB = true;
				fastSim.enlargeAABB = true;
			}
			shapeId = shape.nextShapeId;
		}
	}
}

constexpr int kEnlargeByList = -2; // finalizeBody -> stepFinalize: this body's enlarged proxies are found by walking its shapes

// solver.c:543-726, one awake body
F2D_HDF inline void finalizeBodyTail( World* w, int simIndex );
F2D_HDF inline void finalizeBodyVote( World* w, int simIndex );
// `vote`: cast the body's island vote (finalizeBodyVote) right here; false when the island split may still be running
// beside this loop (stepSolve), then the votes follow in a loop of their own once the islands are final
F2D_HDC inline void finalizeBody( World* w, int simIndex, bool vote )
{
	BodyState& state = ptr( w, w->states )[simIndex];
	int bodyId = ptr( w, w->awakeBodies )[simIndex];
	BodySim& sim = ptr( w, w->sims )[bodyId];
	Body& body = ptr( w, w->bodies )[bodyId];
	// The whole gather in 16-byte chunks, all issued before the first use: the state (2), the three sectors of the
	// simulation record (5, the sixth chunk holds only flags) and sector 1 of the body record (2).
	const Q4 stLo = load16( &state.v ), stHi = load16( &state.dp );
	const Q4 sXf = load16( &sim.transform ), sLc = load16( &sim.localCenter ), sCen = load16( &sim.center ),
			 sExt = load16( &sim.rotation0 );
	const Q4 bIsl = load16( &body.islandId );
	const int headShapeId = (int)floatBits( bIsl.y );
	// ... and the scalars the decisions below read (the flags of the simulation record sit in a sector of their own)
	const bool simIsBullet = sim.isBullet, simIsSpeedCapped = sim.isSpeedCapped;
	const int bodyType = body.type;
	const bool bodyEnableSleep = body.enableSleep;
	const uint16_t bodyGeneration = body.generation;
	const uint64_t bodyUserData = body.userData;
	const float bodySleepThreshold = bIsl.z;
	F2D_ISSUE_I( (int)simIsBullet | (int)simIsSpeedCapped << 1 );
	F2D_ISSUE_I( bodyType | (int)bodyEnableSleep << 8 | (int)bodyGeneration << 16 );
	{
		// the island and the first shape are needed at the end of this routine: start fetching them now, so that the
		// body -> island and body -> shape hops overlap with the arithmetic instead of following it
		const int islandId = (int)floatBits( bIsl.x );
		prefetchLine( ptr( w, w->islands ) + islandId );
		const char* shapeBytes = reinterpret_cast<const char*>( ptr( w, w->shapes ) + headShapeId );
		prefetchLine( shapeBytes );
		prefetchLine( shapeBytes + 96 ); // polygon vertices
		prefetchLine( shapeBytes + 224 ); // radius, count
	}
	const float timeStep = w->step.dt;
	const float invTimeStep = w->step.inv_dt;

	const V2 v = { stLo.x, stLo.y };
	const float wv = stLo.z;
	const V2 dp = { stHi.x, stHi.y };
	const Rot dq = { stHi.z, stHi.w };
	const V2 localCenter = { sLc.x, sLc.y };
	const float minExtent = sExt.z, maxExtent = sExt.w;
	const V2 center = add( V2{ sCen.x, sCen.y }, dp );
	Xf transform;
	transform.q = normalizeRot( mulRot( dq, Rot{ sXf.z, sXf.w } ) );
	float maxVelocity = length( v ) + absf( wv ) * maxExtent;
	float maxDeltaPosition = length( dp ) + absf( dq.s ) * maxExtent;
	float positionSleepFactor = 0.5f;
	float sleepVelocity = maxf( maxVelocity, positionSleepFactor * invTimeStep * maxDeltaPosition );
	store16( &state.dp, Q4{ 0.0f, 0.0f, 1.0f, 0.0f } );
	transform.p = sub( center, rotate( transform.q, localCenter ) );
	store16( &sim.transform, Q4{ transform.p.x, transform.p.y, transform.q.c, transform.q.s } );
	sim.center = center;

	body.bodyMoveIndex = simIndex;
	{
		// 40-byte record (types.h:1136-1142): two 16-byte chunks and the flag word (records are 8-byte aligned)
		BodyMoveEvent ev;
		ev.transform = transform;
		ev.bodyId = BodyId{ bodyId + 1, w->worldId, bodyGeneration };
		ev.userData = bodyUserData;
		ev.fellAsleep = false;
		ptr( w, w->moveEvents )[simIndex] = ev;
	}

	sim.force = V2{ 0.0f, 0.0f };
	sim.torque = 0.0f;
	body.isSpeedCapped = simIsSpeedCapped;
	sim.isSpeedCapped = false;
	sim.isFast = false;

	if ( w->enableSleep == false || bodyEnableSleep == false || sleepVelocity > bodySleepThreshold )
	{
		body.sleepTime = 0.0f;
		if ( bodyType == kDynamicBody && w->enableContinuous && maxVelocity * timeStep > 0.5f * minExtent )
		{
			sim.isFast = true;
			if ( simIsBullet )
			{
				int bulletIndex = atomAdd( &w->step.bulletCount, 1 );
				ptr( w, w->bullets )[bulletIndex] = simIndex;
			}
			else if ( w->hostCallbacks & ( kHostCustomFilter | kHostPreSolve ) )
			{
				// callback-mediated step: the continuous pass consults the host's custom filter per candidate shape and
				// its pre-solve callback per hit (solver.c:271-282, 366-379), so it - and everything of this routine
				// that depends on it - waits for the host.
				// Parked from the end of the bullet array downwards.
				int slot = atomAdd( &w->step.fastDeferredCount, 1 );
				ptr( w, w->bullets )[w->bullets.cap - 1 - slot] = simIndex;
				if ( vote )
					finalizeBodyVote( w, simIndex ); // does not depend on the continuous pass
				return;
			}
			else
			{
				solveContinuous( w, simIndex );
			}
		}
		else
		{
			sim.center0 = center;
			sim.rotation0 = transform.q;
		}
	}
	else
	{
		sim.center0 = center;
		sim.rotation0 = transform.q;
		body.sleepTime += timeStep;
	}

	if ( vote )
		finalizeBodyVote( w, simIndex );
	finalizeBodyTail( w, simIndex );
}

// The island vote of one awake body (solver.c:649-675): keeps its island awake, or proposes it for splitting. The only
// part of the body finalize that looks at islands; it depends on nothing the continuous pass changes.
F2D_HDF inline void finalizeBodyVote( World* w, int simIndex )
{
	int bodyId = ptr( w, w->awakeBodies )[simIndex];
	const Body& body = ptr( w, w->bodies )[bodyId];
	const Island& island = ptr( w, w->islands )[body.islandId];
	if ( body.sleepTime < kTimeToSleep )
	{
		int islandIndex = island.localIndex;
		atomOr64( ptr( w, w->islandBits ) + ( islandIndex >> 6 ), 1ull << ( islandIndex & 63 ) );
	}
	else if ( island.constraintRemoveCount > 0 )
	{
		// split candidate: largest sleepTime, first body in awake order on ties (solver.c:666-675 with one worker)
		unsigned long long key = ( (unsigned long long)floatBits( body.sleepTime ) << 32 ) |
								 (unsigned long long)( 0xFFFFFFFFu - (uint32_t)simIndex );
		atomMax64( &w->step.splitKey, key );
	}
}

// What finalizeBody does after the continuous pass of a fast body: the shape boxes
F2D_HDF inline void finalizeBodyTail( World* w, int simIndex )
{
	int bodyId = ptr( w, w->awakeBodies )[simIndex];
	BodySim& sim = ptr( w, w->sims )[bodyId];
	Body& body = ptr( w, w->bodies )[bodyId];

	Xf transform = sim.transform;
	bool isFast = sim.isFast;
	Shape* shapes = ptr( w, w->shapes );
	uint64_t* enlargedBits = ptr( w, w->enlargedBits );
	// Bodies with a single shape (nearly all) hand their enlarged proxy straight to the move-array pass: the new leaf
	// box is written here, where the shape record is at hand, and the key is parked by awake index, so that pass does
	// not have to walk body -> shape again. Other bodies (and fast bullets, whose boxes are not final yet) are marked
	// kEnlargeByList and go through the shape-list walk of stepFinalize.
	// (one shape: the head of the list has no successor - read from the shape record, which is needed anyway)
	const bool single = body.headShapeId != kNull && shapes[body.headShapeId].nextShapeId == kNull && ( sim.isBullet && isFast ) == false;
	int parkedKey = single ? kNull : kEnlargeByList;
	int shapeId = body.headShapeId;
	while ( shapeId != kNull )
	{
		Shape& shape = shapes[shapeId];
		// (what the rest of the iteration reads of the record, requested before the first store to it)
		const Box fat = shape.fatAABB;
		const int nextShapeId = shape.nextShapeId;
		const int proxyKey = shape.proxyKey;
		F2D_ISSUE_F( fat.lo.x );
		F2D_ISSUE_I( nextShapeId );
		F2D_ISSUE_I( proxyKey );
		if ( isFast )
		{
			atomOr64( enlargedBits + ( simIndex >> 6 ), 1ull << ( simIndex & 63 ) );
		}
		else
		{
			Box aabb = inflate( shapeAABB( shape, transform ), kSpeculative );
			shape.aabb = aabb;
			if ( boxContains( fat, aabb ) == false )
			{
				shape.fatAABB = inflate( aabb, kAabbMargin );
				shape.enlargedAABB = true;
				atomOr64( enlargedBits + ( simIndex >> 6 ), 1ull << ( simIndex & 63 ) );
			}
		}
		if ( single && shape.enlargedAABB )
		{
			parkedKey = proxyKey;
			TreeNode& leaf = ptr( w, w->trees[proxyType( parkedKey )].nodes )[proxyId( parkedKey )];
			leaf.box = shape.fatAABB;
			leaf.flags |= kNodeMoved;
			shape.enlargedAABB = false;
		}
		shapeId = nextShapeId;
	}
	ptr( w, w->islBodies )[simIndex] = parkedKey;
}

// The body loop of finalize (solver.c:1724-1747) on `t`: the whole team, or the crew while the island split still runs.
// `vote` as in finalizeBody.
template <class Team> F2D_HDF inline void finalizeBodiesTeam( World* w, Team& t, bool vote )
{
	const int awakeBodyCount = w->step.awakeBodyCount;
	if ( t.rank() == 0 )
		w->step.fastDeferredCount = 0;
	// clear bit sets (solver.c:1724-1733); the island bits belong to the votes
	int bodyWords = ( awakeBodyCount + 63 ) >> 6;
	uint64_t* eb = ptr( w, w->enlargedBits );
	for ( int i = t.rank(); i < bodyWords; i += t.size() )
		eb[i] = 0;
	if ( vote )
	{
		int islandWords = ( w->awakeIslands.count + 63 ) >> 6;
		uint64_t* ib = ptr( w, w->islandBits );
		for ( int i = t.rank(); i < islandWords; i += t.size() )
			ib[i] = 0;
	}
	t.sync();
	// the body and sim records of this thread's next body are requested into L2 while the current one is finalized
	const int32_t* awake = ptr( w, w->awakeBodies );
	const Body* bodyArr = ptr( w, w->bodies );
	const BodySim* simArr = ptr( w, w->sims );
	const int stride = t.size();
	for ( int i = t.rank(); i < awakeBodyCount; i += stride )
	{
		if ( i + stride < awakeBodyCount )
		{
			int next = awake[i + stride];
			prefetchL2( &simArr[next] );
			prefetchL2( &simArr[next].center );
			prefetchL2( &bodyArr[next].islandId );
		}
		finalizeBody( w, i, vote );
	}
	t.sync();
}

// The island votes on their own, once the islands are final (after the split has been applied)
template <class Team> F2D_HDF inline void finalizeVotesTeam( World* w, Team& t )
{
	const int awakeBodyCount = w->step.awakeBodyCount;
	int islandWords = ( w->awakeIslands.count + 63 ) >> 6;
	uint64_t* ib = ptr( w, w->islandBits );
	for ( int i = t.rank(); i < islandWords; i += t.size() )
		ib[i] = 0;
	t.sync();
	for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
		finalizeBodyVote( w, i );
	t.sync();
}

// Team-parallel enlarge of one leaf: the end state (leaf box replaced; every ancestor box = union with it; every
// ancestor flagged enlarged) equals what any serial order of b2DynamicTree_EnlargeProxy produces
// (dynamic_tree.c:827-866), because box growth is a commutative min/max and flags only accumulate.
F2D_HDF inline void enlargeLeafParallel( World* w, Tree& tree, int leaf, Box box )
{
	TreeNode* nodes = ptr( w, tree.nodes );
	nodes[leaf].box = box;
	int parent = nodes[leaf].parent;
	while ( parent != kNull )
	{
		TreeNode& n = nodes[parent];
		// Reads that bypass L1 (the atomics below are performed in L2, a line cached in L1 would never show them):
		// boxes only grow and flags only accumulate during this phase, so an old value can only cause a redundant
		// atomic, never a missed one - but a FRESH value is what makes the early exit below bite. Once an ancestor
		// already contains the box AND is flagged, whoever grew / flagged it is carrying both up to the root (or they
		// were that way before the phase, and the tree invariants - parent contains child, flags are upward closed -
		// cover the rest of the path).
#if defined( __CUDA_ARCH__ )
		const float4 nb = __ldcg( reinterpret_cast<const float4*>( &n.box ) );
		const int4 links = __ldcg( reinterpret_cast<const int4*>( &n.child1 ) ); // child1, child2, parent, height | flags << 16
		const Box nbox = { { nb.x, nb.y }, { nb.z, nb.w } };
		const int next = links.z;
		const bool flagged = ( ( (uint32_t)links.w >> 16 ) & kNodeEnlarged ) != 0;
#else
		const Box nbox = n.box;
		const int next = n.parent;
		const bool flagged = ( n.flags & kNodeEnlarged ) != 0;
#endif
		const bool contains = nbox.lo.x <= box.lo.x && nbox.lo.y <= box.lo.y && box.hi.x <= nbox.hi.x && box.hi.y <= nbox.hi.y;
		if ( contains && flagged )
			break;
		if ( box.lo.x < nbox.lo.x )
			atomMinF( &n.box.lo.x, box.lo.x );
		if ( box.lo.y < nbox.lo.y )
			atomMinF( &n.box.lo.y, box.lo.y );
		if ( nbox.hi.x < box.hi.x )
			atomMaxF( &n.box.hi.x, box.hi.x );
		if ( nbox.hi.y < box.hi.y )
			atomMaxF( &n.box.hi.y, box.hi.y );
		if ( flagged == false )
		{
			// height (low 16 bits) and flags (high 16 bits) share one 32-bit word
			atomOr32( reinterpret_cast<uint32_t*>( &n.height ), (uint32_t)kNodeEnlarged << 16 );
		}
		parent = next;
	}
}

// The same end state for up to kEnlargeWalks leaves of one thread at a time. A walk is a chain of dependent loads (box
// and links of the parent, then of its parent ...) as long as the tree is high (23-25 levels at 820 proxies), and a
// thread owns several walks, so it advances them in lockstep: the loads of one level of all its walks are issued
// together and cost one round trip instead of one each. `keys[first + k * stride]` are the proxy keys (kNull: skip);
// the leaf boxes have been written already.
constexpr int kEnlargeWalks = 4;
F2D_HDF inline void enlargeWalks( World* w, const int32_t* keys, int first, int stride, int total )
{
	TreeNode* nodes[kEnlargeWalks];
	int cur[kEnlargeWalks];
	Box box[kEnlargeWalks];
	bool any = false;
#if defined( __CUDA_ARCH__ )
#pragma unroll
#endif
	for ( int k = 0; k < kEnlargeWalks; ++k )
	{
		int m = first + k * stride;
		int key = m < total ? keys[m] : kNull;
		cur[k] = kNull;
		nodes[k] = nullptr;
		box[k] = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
		if ( key != kNull )
		{
			nodes[k] = ptr( w, w->trees[proxyType( key )].nodes );
			const TreeNode& leaf = nodes[k][proxyId( key )];
			box[k] = leaf.box;
			cur[k] = leaf.parent;
			any = any || cur[k] != kNull;
		}
	}
	while ( any )
	{
		any = false;
		Box nbox[kEnlargeWalks];
		int next[kEnlargeWalks];
		bool flagged[kEnlargeWalks];
#if defined( __CUDA_ARCH__ )
#pragma unroll
#endif
		for ( int k = 0; k < kEnlargeWalks; ++k )
		{
			nbox[k] = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
			next[k] = kNull;
			flagged[k] = true;
			if ( cur[k] == kNull )
				continue;
			TreeNode& n = nodes[k][cur[k]];
			// L1-bypassing reads for the same reason as in enlargeLeafParallel
#if defined( __CUDA_ARCH__ )
			const float4 nb = __ldcg( reinterpret_cast<const float4*>( &n.box ) );
			const int4 links = __ldcg( reinterpret_cast<const int4*>( &n.child1 ) );
			nbox[k] = Box{ { nb.x, nb.y }, { nb.z, nb.w } };
			next[k] = links.z;
			flagged[k] = ( ( (uint32_t)links.w >> 16 ) & kNodeEnlarged ) != 0;
#else
			nbox[k] = n.box;
			next[k] = n.parent;
			flagged[k] = ( n.flags & kNodeEnlarged ) != 0;
#endif
		}
#if defined( __CUDA_ARCH__ )
#pragma unroll
#endif
		for ( int k = 0; k < kEnlargeWalks; ++k )
		{
			if ( cur[k] == kNull )
				continue;
			TreeNode& n = nodes[k][cur[k]];
			const Box b = box[k], nb = nbox[k];
			const bool contains = nb.lo.x <= b.lo.x && nb.lo.y <= b.lo.y && b.hi.x <= nb.hi.x && b.hi.y <= nb.hi.y;
			if ( contains && flagged[k] )
			{
				cur[k] = kNull;
				continue;
			}
			if ( b.lo.x < nb.lo.x )
				atomMinF( &n.box.lo.x, b.lo.x );
			if ( b.lo.y < nb.lo.y )
				atomMinF( &n.box.lo.y, b.lo.y );
			if ( nb.hi.x < b.hi.x )
				atomMaxF( &n.box.hi.x, b.hi.x );
			if ( nb.hi.y < b.hi.y )
				atomMaxF( &n.box.hi.y, b.hi.y );
			if ( flagged[k] == false )
				atomOr32( reinterpret_cast<uint32_t*>( &n.height ), (uint32_t)kNodeEnlarged << 16 );
			cur[k] = next[k];
			any = any || cur[k] != kNull;
		}
	}
}

// Hit events, colour-major then array order: solver.c:1758-1818
F2D_HDF inline void reportHitEvents( World* w )
{
	float threshold = w->hitEventThreshold;
	const ContactSim* sims = ptr( w, w->contactSims );
	const Shape* shapes = ptr( w, w->shapes );
	for ( int i = 0; i < kColorCount; ++i )
	{
		const int32_t* list = ptr( w, w->colorContacts[i] );
		int n = w->colorContacts[i].count;
		for ( int j = 0; j < n; ++j )
		{
			const ContactSim& sim = sims[list[j]];
			if ( ( sim.simFlags & kSimEnableHitEvent ) == 0 )
				continue;
			HitEvent ev;
			memset( &ev, 0, sizeof( ev ) );
			ev.approachSpeed = threshold;
			bool hit = false;
			const Manifold manifold = unpackManifold( sim.manifold );
			for ( int k = 0; k < manifold.pointCount; ++k )
			{
				const ManifoldPoint& mp = manifold.points[k];
				float approachSpeed = -mp.normalVelocity;
				if ( approachSpeed > ev.approachSpeed && mp.totalNormalImpulse > 0.0f )
				{
					ev.approachSpeed = approachSpeed;
					ev.point = mp.point;
					hit = true;
				}
			}
			if ( hit )
			{
				ev.normal = manifold.normal;
				ev.a = makeShapeId( w, shapes[sim.shapeIdA] );
				ev.b = makeShapeId( w, shapes[sim.shapeIdB] );
				F2D_PUSH_EVENT( w, w->hitEvents, ev );
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ sensors
F2D_HD ShapeRef* sensorList( World* w, int sensorIndex, int which )
{
	return ptr( w, w->sensorRefs ) + (size_t)( 2 * sensorIndex + which ) * w->sensorOverlapCap;
}

// One sensor shape: sensor.c:101-212. Swaps its overlap lists, queries the three trees with the shape's AABB, keeps
// the visitors whose GJK distance is (numerically) zero, sorts them by (shape id, generation) and flags the sensor when
// the set differs from the previous step's.
F2D_HDC inline void sensorTask( World* w, int sensorIndex )
{
	Sensor& sensor = ptr( w, w->sensors )[sensorIndex];
	const Shape* shapes = ptr( w, w->shapes );
	const Shape& sensorShape = shapes[sensor.shapeId];
	uint64_t* bits = ptr( w, w->sensorBits );
	sensor.flip ^= 1;
	sensor.count1 = sensor.count2;
	sensor.count2 = 0;
	const ShapeRef* refs1 = sensorList( w, sensorIndex, sensor.flip ^ 1 );
	ShapeRef* refs2 = sensorList( w, sensorIndex, sensor.flip );
	const Body& body = ptr( w, w->bodies )[sensorShape.bodyId];
	if ( body.setIndex == kDisabledSet || sensorShape.enableSensorEvents == false )
	{
		if ( sensor.count1 != 0 )
			atomOr64( bits + ( sensorIndex >> 6 ), 1ull << ( sensorIndex & 63 ) );
		return;
	}
	const BodySim* sims = ptr( w, w->sims );
	const Xf transform = sims[sensorShape.bodyId].transform;
	const ShapeProxy sensorProxy = makeShapeProxy( sensorShape );
	int count2 = 0;
	auto visit = [&]( int, uint64_t userData ) -> bool {
		int shapeId = (int)userData;
		if ( shapeId == sensor.shapeId )
			return true;
		const Shape& other = shapes[shapeId];
		if ( other.enableSensorEvents == false )
			return true;
		if ( other.bodyId == sensorShape.bodyId )
			return true;
		if ( shouldShapesCollide( sensorShape.filter, other.filter ) == false )
			return true;
		Xf otherTransform = sims[other.bodyId].transform;
		SimplexCache cache;
		memset( &cache, 0, sizeof( cache ) );
		DistanceOutput out = shapeDistance( sensorProxy, makeShapeProxy( other ), transform, otherTransform, true, &cache );
		if ( ( out.distance < 10.0f * FLT_EPSILON ) == false )
			return true;
		if ( count2 >= w->sensorOverlapCap )
		{
			// cannot happen while the capacity follows the shape count (f2d_image.h); a capped world drops the overlap
			setError( w, kErrTruncated, __LINE__ );
			return true;
		}
		refs2[count2].shapeId = shapeId;
		refs2[count2].generation = other.generation;
		refs2[count2].pad = 0;
		count2 += 1;
		return true;
	};
	for ( int tree = 0; tree < 3; ++tree )
		treeQuery( w, w->trees[tree], sensorShape.aabb, sensorShape.filter.mask, visit );
	// (shape id, generation) ascending: keys are unique, so any sort reproduces qsort's result (sensor.c:185)
	for ( int i = 1; i < count2; ++i )
	{
		ShapeRef key = refs2[i];
		int j = i - 1;
		while ( j >= 0 && ( refs2[j].shapeId > key.shapeId || ( refs2[j].shapeId == key.shapeId && refs2[j].generation > key.generation ) ) )
		{
			refs2[j + 1] = refs2[j];
			j -= 1;
		}
		refs2[j + 1] = key;
	}
	sensor.count2 = count2;
	bool changed = sensor.count1 != count2;
	for ( int i = 0; i < count2 && changed == false; ++i )
		changed = refs1[i].shapeId != refs2[i].shapeId || refs1[i].generation != refs2[i].generation;
	if ( changed )
		atomOr64( bits + ( sensorIndex >> 6 ), 1ull << ( sensorIndex & 63 ) );
}

// Begin / end events of one flagged sensor: ordered merge of the two sorted lists (sensor.c:264-345)
F2D_HDF inline void sensorEvents( World* w, int sensorIndex )
{
	const Sensor& sensor = ptr( w, w->sensors )[sensorIndex];
	const Shape& sensorShape = ptr( w, w->shapes )[sensor.shapeId];
	const ShapeId sensorId = { sensor.shapeId + 1, w->worldId, sensorShape.generation };
	const ShapeRef* refs1 = sensorList( w, sensorIndex, sensor.flip ^ 1 );
	const ShapeRef* refs2 = sensorList( w, sensorIndex, sensor.flip );
	auto ended = [&]( const ShapeRef& r ) {
		SensorEvent ev = { sensorId, ShapeId{ r.shapeId + 1, w->worldId, r.generation } };
		F2D_PUSH_EVENT( w, w->sensorEndEvents[w->endEventArrayIndex], ev );
	};
	auto began = [&]( const ShapeRef& r ) {
		SensorEvent ev = { sensorId, ShapeId{ r.shapeId + 1, w->worldId, r.generation } };
		F2D_PUSH_EVENT( w, w->sensorBeginEvents, ev );
	};
	int i1 = 0, i2 = 0;
	while ( i1 < sensor.count1 && i2 < sensor.count2 )
	{
		const ShapeRef& r1 = refs1[i1];
		const ShapeRef& r2 = refs2[i2];
		if ( r1.shapeId == r2.shapeId )
		{
			if ( r1.generation < r2.generation )
			{
				ended( r1 );
				i1 += 1;
			}
			else if ( r1.generation > r2.generation )
			{
				began( r2 );
				i2 += 1;
			}
			else
			{
				i1 += 1;
				i2 += 1;
			}
		}
		else if ( r1.shapeId < r2.shapeId )
		{
			ended( r1 );
			i1 += 1;
		}
		else
		{
			began( r2 );
			i2 += 1;
		}
	}
	for ( ; i1 < sensor.count1; ++i1 )
		ended( refs1[i1] );
	for ( ; i2 < sensor.count2; ++i2 )
		began( refs2[i2] );
}

// sensor.c:215-351 b2OverlapSensors
template <class Team> F2D_HDF inline void overlapSensors( World* w, Team& t )
{
	const int sensorCount = w->sensors.count;
	if ( sensorCount == 0 )
		return;
	uint64_t* bits = ptr( w, w->sensorBits );
	const int words = ( sensorCount + 63 ) >> 6;
	for ( int k = t.rank(); k < words; k += t.size() )
		bits[k] = 0;
	t.sync();
	for ( int i = t.rank(); i < sensorCount; i += t.size() )
		sensorTask( w, i );
	t.sync();
	if ( t.rank() == 0 )
	{
		for ( int k = 0; k < words; ++k )
		{
			uint64_t word = bits[k];
			for ( int b = 0; word != 0; ++b, word >>= 1 )
			{
				if ( word & 1ull )
					sensorEvents( w, 64 * k + b );
			}
		}
	}
	t.sync();
}

// `part`: kFinalizeAll, or one third of a callback-mediated step. The continuous pass asks the host's filter about every
// candidate shape and its pre-solve callback about every hit (solver.c:271-282, 366-379), so there the HOST runs solveContinuous for the
// fast bodies - after kFinalizeBodies for ordinary bodies (their boxes feed the move array built by kFinalizeMoves), after
// kFinalizeMoves for bullets - and kFinalizeEnd resumes behind the bullets' continuous pass.
enum : int
{
	kFinalizeAll = 0,
	kFinalizeBodies = 1,
	kFinalizeMoves = 2,
	kFinalizeEnd = 3
};
template <class Team> F2D_HDF inline void stepFinalize( World* w, Team& t, int part = kFinalizeAll )
{
	const int awakeBodyCount = w->step.awakeBodyCount;
	const bool solved = awakeBodyCount > 0 && w->step.dt > 0.0f && ( w->error & kErrCapacity ) == 0;
	const bool doBodies = part == kFinalizeAll || part == kFinalizeBodies;
	const bool doMoves = part == kFinalizeAll || part == kFinalizeMoves;
	const bool doEnd = part == kFinalizeAll || part == kFinalizeEnd;
	if ( solved && doBodies )
	{
		// stepSolve has run the body loop already when it overlapped it with the island split (World::step.bodiesFinalized)
		if ( w->step.bodiesFinalized == 0 )
			finalizeBodiesTeam( w, t, true );
		else
			finalizeVotesTeam( w, t );
		t.sync();
		if ( t.rank() == 0 )
			w->step.bodiesFinalized = 0;
		F2D_MARK( w, t, pfFinalizeBodies );
	}
	if ( solved && doMoves )
	{
		uint64_t* eb = ptr( w, w->enlargedBits );
		if ( t.rank() == 0 && w->hitEventCapable > 0 )
			reportHitEvents( w );
		F2D_MARK( w, t, pfHitEvents );

		// Enlarged proxies -> broadphase, and next step's move array in ascending awake index x shape-list order
		// (solver.c:1835-1907). Count, scan, then scatter + enlarge in parallel.
		const int32_t* awakeBodies = ptr( w, w->awakeBodies );
		const Body* bodies = ptr( w, w->bodies );
		BodySim* sims = ptr( w, w->sims );
		Shape* shapes = ptr( w, w->shapes );
		int32_t* scan = ptr( w, w->scan );
		const int32_t* parked = ptr( w, w->islBodies ); // island-solve scratch, free here: see finalizeBody
		for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
		{
			int count = 0;
			const int key = parked[i];
			if ( key != kEnlargeByList )
			{
				count = key != kNull ? 1 : 0;
			}
			else if ( eb[i >> 6] & ( 1ull << ( i & 63 ) ) )
			{
				int bodyId = awakeBodies[i];
				const BodySim& sim = sims[bodyId];
				bool fastBullet = sim.isBullet && sim.isFast;
				for ( int s = bodies[bodyId].headShapeId; s != kNull; s = shapes[s].nextShapeId )
					count += ( fastBullet || shapes[s].enlargedAABB ) ? 1 : 0;
			}
			scan[i] = count;
		}
		t.sync();
		int moveTotal = t.exclusiveScan( scan, awakeBodyCount );
		if ( moveTotal > w->moveArray.cap )
		{
			if ( t.rank() == 0 )
				setError( w, kErrCapacity, __LINE__ );
			moveTotal = 0;
		}
		int32_t* moves = ptr( w, w->moveArray );
		int32_t* walks = ptr( w, w->moveHeads ); // pair-finding scratch, free here: proxy keys whose ancestors must grow
		if ( moveTotal > w->moveHeads.cap )
		{
			if ( t.rank() == 0 )
				setError( w, kErrCapacity, __LINE__ );
			moveTotal = 0;
		}
		for ( int i = t.rank(); i < awakeBodyCount; i += t.size() )
		{
			if ( moveTotal == 0 )
				continue;
			const int parkedKey = parked[i];
			if ( parkedKey != kEnlargeByList )
			{
				if ( parkedKey != kNull )
				{
					walks[scan[i]] = parkedKey;
					moves[scan[i]] = parkedKey;
				}
				continue;
			}
			if ( ( eb[i >> 6] & ( 1ull << ( i & 63 ) ) ) == 0 )
				continue;
			int bodyId = awakeBodies[i];
			const BodySim& sim = sims[bodyId];
			bool fastBullet = sim.isBullet && sim.isFast;
			int out = scan[i];
			for ( int s = bodies[bodyId].headShapeId; s != kNull; s = shapes[s].nextShapeId )
			{
				Shape& shape = shapes[s];
				if ( fastBullet )
				{
					walks[out] = kNull; // enlarged after the bullet's own continuous pass
					moves[out++] = shape.proxyKey;
					ptr( w, w->trees[proxyType( shape.proxyKey )].nodes )[proxyId( shape.proxyKey )].flags |= kNodeMoved;
				}
				else if ( shape.enlargedAABB )
				{
					int key = shape.proxyKey;
					TreeNode& leaf = ptr( w, w->trees[proxyType( key )].nodes )[proxyId( key )];
					leaf.box = shape.fatAABB;
					leaf.flags |= kNodeMoved;
					walks[out] = key;
					moves[out++] = key;
					shape.enlargedAABB = false;
				}
			}
		}
		t.sync();
		// the walks to the root, four per thread in lockstep (see enlargeWalks)
		for ( int first = t.rank(); first < moveTotal; first += kEnlargeWalks * t.size() )
			enlargeWalks( w, walks, first, t.size(), moveTotal );
		if ( t.rank() == 0 )
			w->moveArray.count = moveTotal;
		t.sync();
		F2D_MARK( w, t, pfEnlarge );
	}
	if ( solved && doEnd )
	{
		const int32_t* awakeBodies = ptr( w, w->awakeBodies );
		const Body* bodies = ptr( w, w->bodies );
		BodySim* sims = ptr( w, w->sims );
		Shape* shapes = ptr( w, w->shapes );
		uint64_t* ib = ptr( w, w->islandBits );
		// bullets: continuous against everything, then enlarge (solver.c:1915-1988)
		int bulletCount = w->step.bulletCount;
		if ( bulletCount > 0 )
		{
			const int32_t* bullets = ptr( w, w->bullets );
			if ( part == kFinalizeAll ) // kFinalizeEnd: the host has run the continuous pass of the bullets
			{
				for ( int i = t.rank(); i < bulletCount; i += t.size() )
					solveContinuous( w, bullets[i] );
			}
			t.sync();
			for ( int i = t.rank(); i < bulletCount; i += t.size() )
			{
				int bodyId = awakeBodies[bullets[i]];
				BodySim& sim = sims[bodyId];
				if ( sim.enlargeAABB == false )
					continue;
				sim.enlargeAABB = false;
				for ( int s = bodies[bodyId].headShapeId; s != kNull; s = shapes[s].nextShapeId )
				{
					Shape& shape = shapes[s];
					if ( shape.enlargedAABB == false )
						continue;
					shape.enlargedAABB = false;
					enlargeLeafParallel( w, w->trees[kDynamicBody], proxyId( shape.proxyKey ), shape.fatAABB );
				}
			}
			t.sync();
		}

		F2D_MARK( w, t, pfBullets );
		// island sleep: reverse scan of awake islands (solver.c:1995-2051). Which islands may fall asleep (no body voted
		// to stay awake, no constraint removed: solver_set.c:156-165) is found by the whole team; the serial, order-defining
		// part then only visits those, still last island first (a settled pile is hundreds of small islands, nearly
		// all of them kept awake: the serial scan over them used to cost as much as the whole finalize loop).
		const int islandCount = w->enableSleep ? w->awakeIslands.count : 0;
		int32_t* maySleep = ptr( w, w->scan ); // free since the move array was built
		const bool listed = islandCount <= w->scan.cap;
		int sleeperCount = -1;
		if ( listed )
		{
			const int32_t* islandList = ptr( w, w->awakeIslands );
			const Island* islands = ptr( w, w->islands );
			for ( int i = t.rank(); i < islandCount; i += t.size() )
			{
				const bool keptAwake = ( ib[i >> 6] & ( 1ull << ( i & 63 ) ) ) != 0;
				maySleep[i] = ( keptAwake == false && islands[islandList[i]].constraintRemoveCount == 0 ) ? 1 : 0;
			}
			t.sync();
			// flags -> prefix sums: the serial part below reads one number when (as nearly always) nobody falls asleep
			sleeperCount = t.exclusiveScan( maySleep, islandCount );
		}
		if ( t.rank() == 0 )
		{
			w->step.bulletCount = 0;
			if ( w->enableSleep )
			{
				unsigned long long key = w->step.splitKey;
				if ( key != 0 )
				{
					int simIndex = (int)( 0xFFFFFFFFu - (uint32_t)( key & 0xFFFFFFFFull ) );
					w->splitIslandId = bodies[awakeBodies[simIndex]].islandId;
				}
				const int32_t* islandList = ptr( w, w->awakeIslands );
				int above = sleeperCount; // prefix sum one past the current index
				for ( int islandIndex = sleeperCount == 0 ? -1 : islandCount - 1; islandIndex >= 0; islandIndex -= 1 )
				{
					bool skip;
					if ( listed )
					{
						const int here = maySleep[islandIndex];
						skip = here == above;
						above = here;
					}
					else
					{
						skip = ( ib[islandIndex >> 6] & ( 1ull << ( islandIndex & 63 ) ) ) != 0;
					}
					if ( skip )
						continue;
					// an island that falls asleep leaves the awake list by swap-remove: the entry that moves comes from
					// behind this index and has been visited already, so the flags by index stay valid
					trySleepIsland( w, islandList[islandIndex] );
				}
			}
		}
		t.sync();
		F2D_MARK( w, t, pfSleep );
	}

	if ( doEnd == false )
		return;
	// world.c:788-793
	overlapSensors( w, t );

	if ( t.rank() == 0 )
	{
		// swap end-event buffers (world.c:807-811)
		w->endEventArrayIndex = 1 - w->endEventArrayIndex;
		w->sensorEndEvents[w->endEventArrayIndex].count = 0;
		w->endEvents[w->endEventArrayIndex].count = 0;
		w->locked = false;
	}
	t.sync();
}

// Zero time step: only the event buffers advance (world.c:716-725)
template <class Team> F2D_HDF inline void stepZeroDt( World* w, Team& t )
{
	if ( t.rank() == 0 )
	{
		w->moveEvents.count = 0;
		w->sensorBeginEvents.count = 0;
		w->beginEvents.count = 0;
		w->hitEvents.count = 0;
		w->endEventArrayIndex = 1 - w->endEventArrayIndex;
		w->sensorEndEvents[w->endEventArrayIndex].count = 0;
		w->endEvents[w->endEventArrayIndex].count = 0;
	}
	t.sync();
}

// Launch granularity. kPhaseAll is the product path (one launch per step). The four coarse phases exist for per-phase
// device timing; the split ones for the callback-mediated step, where the host answers b2CustomFilterFcn between
// kPhasePairsQuery and kPhasePairsCreate and b2PreSolveFcn between kPhaseCollideNarrow and kPhaseCollideFinish.
enum Phase
{
	kPhaseAll = 0,
	kPhaseBeginPairs = 1,
	kPhaseCollide = 2,
	kPhaseSolve = 3,
	kPhaseFinalize = 4,
	kPhasePairsQuery = 5, // stepBegin + pair queries
	kPhasePairsCreate = 6,
	kPhaseCollideNarrow = 7,
	kPhaseCollideFinish = 8,
	kPhaseFinalizeBodies = 9, // see stepFinalize
	kPhaseFinalizeMoves = 10,
	kPhaseFinalizeEnd = 11,
	kPhaseCollideTreeOnly = 12, // profiling aids: see kCollideTreeOnly
	kPhaseCollideNarrowOnly = 13
};

// Whole step on one team (used by the CTA-per-world kernel and the host emulation)
template <class Team> F2D_HDF inline void stepWorld( World* w, Team& t, float dt, int subStepCount )
{
	if ( dt == 0.0f )
	{
		stepZeroDt( w, t );
		return;
	}
	stepBegin( w, t, dt, subStepCount );
	// a team that rebuilds the trees beside the serial parts of the step creates this step's contacts there too
	bool deferPairs = false;
	if constexpr ( Team::kCanSplitTree )
		deferPairs = t.canSplitTree() && w->hostCallbacks == 0;
	stepPairs( w, t, deferPairs ? kPairsDefer : kPairsAll );
	if ( w->step.retryContacts != 0 ) // see stepPairs: the host repeats the step on a larger image
		return;
	stepCollide( w, t, kCollideAll, deferPairs );
	stepSolve( w, t );
	stepFinalize( w, t );
}

// One launch-sized piece of the step (see Phase)
template <class Team> F2D_HDF inline void stepWorldPhase( World* w, Team& t, int phase, float dt, int subStepCount )
{
	if ( phase == kPhaseAll )
	{
		stepWorld( w, t, dt, subStepCount );
		return;
	}
	if ( dt == 0.0f )
	{
		if ( phase == kPhaseBeginPairs || phase == kPhasePairsQuery )
			stepZeroDt( w, t );
		return;
	}
	if ( phase != kPhaseBeginPairs && phase != kPhasePairsQuery && w->step.retryContacts != 0 )
		return; // the step stopped in stepPairs and will be repeated
	switch ( phase )
	{
		case kPhaseBeginPairs:
			stepBegin( w, t, dt, subStepCount );
			stepPairs( w, t );
			break;
		case kPhasePairsQuery:
			stepBegin( w, t, dt, subStepCount );
			stepPairs( w, t, kPairsQuery );
			break;
		case kPhasePairsCreate:
			stepPairs( w, t, kPairsCreate );
			break;
		case kPhaseCollide:
			stepCollide( w, t );
			break;
		case kPhaseCollideNarrow:
			stepCollide( w, t, kCollideNarrow );
			break;
		case kPhaseCollideFinish:
			stepCollide( w, t, kCollideFinish );
			break;
		case kPhaseCollideTreeOnly:
			stepCollide( w, t, kCollideTreeOnly );
			break;
		case kPhaseCollideNarrowOnly:
			stepCollide( w, t, kCollideNarrowOnly );
			break;
		case kPhaseSolve:
			stepSolve( w, t );
			break;
		case kPhaseFinalize:
			stepFinalize( w, t );
			break;
		case kPhaseFinalizeBodies:
			stepFinalize( w, t, kFinalizeBodies );
			break;
		case kPhaseFinalizeMoves:
			stepFinalize( w, t, kFinalizeMoves );
			break;
		case kPhaseFinalizeEnd:
			stepFinalize( w, t, kFinalizeEnd );
			break;
	}
}

} // namespace f2d
