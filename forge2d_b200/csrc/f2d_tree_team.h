// forge2d_b200 — team-parallel partial rebuild of the broadphase BVH.
//
// Produces exactly the tree of the reference's serial b2DynamicTree_Rebuild (B2/src/dynamic_tree.c:1872-1989 with
// b2BuildTree :1716-1869 and b2PartitionMid :1426-1532) — same items, same item order, same split at every node — but
// as data-parallel passes over the items:
//   collect   every node whose parent was dissolved (enlarged) is an item; its position in the reference's child1-first
//             DFS is the number of items that precede it = sum over its ancestors, when it hangs under child2, of the
//             item count under child1. Counts come from one walk to the root per item (integer atomics), ranks from
//             a second walk. No traversal stack, no serial DFS.
//   build     top-down median split without levels or team barriers: a long segment of the item array is split by the
//             lanes of one warp (centre bounds by shuffle reduction, then the reference's Hoare partition in closed form:
//             with L = #(centre < pivot), the k-th misplaced item of the left part (ascending) swaps with the k-th
//             misplaced item of the right part (descending); degenerate splits (L == 0 or L == n) leave the order
//             untouched and cut at n/2) and its two halves go back to a work queue that the warps drain together;
//             short segments are finished serially by one thread each.
//   refit     boxes, heights and categories bottom-up by last arrival: the second child subtree to complete refits
//             the node above it and carries on upwards.
// Internal nodes are recycled in place: the dissolved nodes are exactly as many as the build needs, so the free list
// and the node count are unchanged (the reference frees and re-allocates the same set; which internal node lands
// where is unobservable: proxies are leaves, queries follow child links).
#pragma once
#include "f2d_team.h"
#include "f2d_tree.h"

namespace f2d
{

struct TreeScratch
{
	int32_t* segOf;		   // [n] start index of the segment a position belongs to, kNull once the position is a child
	int32_t* segEnd[2];	   // [n] by segment start, double-buffered by level parity
	int32_t* segParent[2]; // [n] by segment start: (parent node << 1) | side, kNull for the root segment
	int32_t* segSplit;	   // [n] by segment start
	int32_t* scanLess;	   // [n+1] prefix sums of the "centre < pivot" flags
	int32_t* lessFlag;	   // [n] (unused: the flag is the difference of neighbouring prefix sums)
	int32_t* badLPos;	   // [n] positions of the misplaced items, left part ascending ...
	int32_t* badRPos;	   // [n] ... right part descending, stored from the segment start
	int32_t* freed;		   // [n] dissolved node recycled for boundary m at freed[m-1]
	int32_t* level;		   // [n] build depth of the node at boundary m, at level[m-1]
	int32_t* posParent;	   // [n] (node << 1) | side of the node a retired position hangs under
	float* lo[2][2];	   // [parity][axis][n] by segment start: bounds of the item centres
	float* hi[2][2];
	int32_t* under; // [nodeCap] items below a dissolved node, by node id
	int32_t* ctrl;	// [8] loop control
};

F2D_HD int treeScratchInts( int shapeCap ) { return 22 * ( shapeCap + 8 ) + 2 * shapeCap + 16 + 8; }

F2D_HD TreeScratch treeScratch( World* w )
{
	int n = w->shapes.cap + 8;
	int32_t* base = ptr( w, w->treeScratch );
	TreeScratch s;
	int k = 0;
	auto take = [&]() { return base + ( k++ ) * n; };
	s.segOf = take();
	s.segEnd[0] = take();
	s.segEnd[1] = take();
	s.segParent[0] = take();
	s.segParent[1] = take();
	s.segSplit = take();
	s.scanLess = take();
	k += 1; // scanLess has n + 1 entries
	s.lessFlag = take();
	s.badLPos = take();
	s.badRPos = take();
	s.freed = take();
	s.level = take();
	s.posParent = take();
	for ( int p = 0; p < 2; ++p )
		for ( int a = 0; a < 2; ++a )
		{
			s.lo[p][a] = reinterpret_cast<float*>( take() );
			s.hi[p][a] = reinterpret_cast<float*>( take() );
		}
	s.ctrl = take();
	s.under = s.ctrl + 8;
	return s;
}

// An item of the rebuild: a leaf, or an internal node that was not enlarged, directly below a dissolved node.
F2D_HD bool treeNodeIsDissolved( const TreeNode& n ) { return n.height > 0 && ( n.flags & kNodeEnlarged ) != 0; }

// Serial build of one small segment [a, e) of the item array: the reference's explicit-stack top-down build
// (dynamic_tree.c:1716-1869) with its in-place Hoare partition (treePartitionMid), recycling the dissolved node
// freed[m-1] for the split at m, and computing boxes / heights / categories on the way back up.
constexpr int kTreeSerialFinish = 12; // segments at most this long are finished by one thread each
constexpr int kTreeSerialFinishArena = 4;
constexpr int kTreeQueueBuildMaxTeam = 256; // teams up to this size use the work-queue build (treeSplitSegment)

// One short segment [a, e) of the item array, built and refitted by ONE thread. Many threads of a warp do this side by
// side on different segments, so the routine is two loops with as little data-dependent control flow as the algorithm
// allows - top-down: pop a range, partition it, create its node, push the halves (ranges of one item are hung right
// away); then bottom-up over the created nodes in reverse creation order (a node is created before everything below it).
// The recursive form of the reference (dynamic_tree.c:1404-1521: descend, come back, refit) is a three-state machine per
// lane, and a warp whose lanes sit in different states runs the states - each with its loads - one after the other
// (measured inside the batch kernel: 240-310 of a world's 650 us of rebuild).
F2D_HDC inline void treeFinishSegment( World* w, Tree& tree, TreeNode* nodes, int32_t* leafIndices, V2* leafCenters, const int32_t* freed,
									   int32_t* levelOf, int a, int e, int parentKey )
{
	constexpr int kDepth = kTreeSerialFinish + 2;
	int32_t rangeStart[kDepth], rangeEnd[kDepth], rangeParent[kDepth]; // parent = (node << 1) | side, kNull for the root
	int32_t created[kDepth];
	int createdCount = 0;
	int top = 0;
	rangeStart[0] = a;
	rangeEnd[0] = e;
	rangeParent[0] = parentKey;
	top = 1;
	while ( top > 0 )
	{
		top -= 1;
		const int start = rangeStart[top], end = rangeEnd[top], key = rangeParent[top];
		const int parentNode = key == kNull ? kNull : key >> 1;
		int index;
		if ( end - start == 1 )
		{
			index = leafIndices[start];
			nodes[index].parent = parentNode;
		}
		else
		{
			const int split = start + treePartitionMid( leafIndices + start, leafCenters + start, end - start );
			index = freed[split - 1];
			levelOf[split - 1] = 0x7fffffff; // refitted here, not by the level-synchronous pass
			TreeNode& node = nodes[index];
			node.box = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
			node.category = 1;
			node.height = 0;
			node.flags = kNodeAllocated;
			node.parent = parentNode;
			node.child1 = kNull;
			node.child2 = kNull;
			if ( createdCount >= kDepth || top + 2 > kDepth )
			{
				setError( w, kErrTreeStack, __LINE__ );
				return;
			}
			created[createdCount++] = index;
			rangeStart[top] = start;
			rangeEnd[top] = split;
			rangeParent[top] = ( index << 1 ) | 0;
			rangeStart[top + 1] = split;
			rangeEnd[top + 1] = end;
			rangeParent[top + 1] = ( index << 1 ) | 1;
			top += 2;
		}
		if ( key == kNull )
			tree.root = index;
		else if ( key & 1 )
			nodes[parentNode].child2 = index;
		else
			nodes[parentNode].child1 = index;
	}
	for ( int k = createdCount - 1; k >= 0; --k )
	{
		TreeNode& node = nodes[created[k]];
		const TreeNode& c1 = nodes[node.child1];
		const TreeNode& c2 = nodes[node.child2];
		node.box = boxUnion( c1.box, c2.box );
		node.height = (uint16_t)( 1 + maxU16( c1.height, c2.height ) );
		node.category = c1.category | c2.category;
	}
}

// A child subtree of the long-segment node `nodeIndex` is complete. The second arrival refits the node from its two
// children and carries on to the node above (bottom-up refit without levels or barriers).
F2D_HDF inline void treeArrive( TreeNode* nodes, int32_t* arrivals, int nodeIndex )
{
	while ( nodeIndex != kNull )
	{
#if defined( __CUDA_ARCH__ )
		__threadfence();
#endif
		if ( atomAdd( arrivals + nodeIndex, 1 ) == 0 )
			return;
#if defined( __CUDA_ARCH__ )
		__threadfence();
#endif
		TreeNode& node = nodes[nodeIndex];
		const int c1i = loadVolatile( &node.child1 ), c2i = loadVolatile( &node.child2 );
		const TreeNode& c1 = nodes[c1i];
		const TreeNode& c2 = nodes[c2i];
		node.box = boxUnion( c1.box, c2.box );
		node.height = (uint16_t)( 1 + maxU16( c1.height, c2.height ) );
		node.category = c1.category | c2.category;
		nodeIndex = node.parent;
	}
}

// One long segment [a, e) of the item array, split by the lanes of a warp: centre bounds, pivot, the reference's
// Hoare partition in closed form (with L = #(centre < pivot), the k-th misplaced item of the left part, ascending,
// swaps with the k-th misplaced item of the right part, descending; L == 0 or L == n leaves the order untouched and
// cuts at n/2), node creation, and the two child segments handed on: single items are hung at once, short segments
// go to the short list, long ones back to the queue.
template <class Lanes>
F2D_HDF inline void treeSplitSegment( World* w, Tree& tree, TreeNode* nodes, int32_t* leafIndices, V2* leafCenters, const TreeScratch& s,
									  int32_t* queue, int32_t* smallList, int a, Lanes L )
{
	const int lane = L.lane(), W = L.count();
	const uint32_t below = lane == 0 ? 0u : ( 0xffffffffu >> ( 32 - lane ) );
	const int e = s.segEnd[0][a];
	const int parentKey = s.segParent[0][a];
	const int count = e - a;
	int split = a + count / 2;
	if ( count > 2 )
	{
		float lx = FLT_MAX, ly = FLT_MAX, hx = -FLT_MAX, hy = -FLT_MAX;
		for ( int i = a + lane; i < e; i += W )
		{
			V2 c = leafCenters[i];
			lx = c.x < lx ? c.x : lx;
			ly = c.y < ly ? c.y : ly;
			hx = c.x > hx ? c.x : hx;
			hy = c.y > hy ? c.y : hy;
		}
		lx = L.reduceMin( lx );
		ly = L.reduceMin( ly );
		hx = L.reduceMax( hx );
		hy = L.reduceMax( hy );
		const bool useX = ( hx - lx ) > ( hy - ly );
		const float pivot = useX ? 0.5f * ( lx + hx ) : 0.5f * ( ly + hy );
		int lessCount = 0;
		for ( int i = a + lane; i < e; i += W )
		{
			V2 c = leafCenters[i];
			lessCount += ( useX ? c.x : c.y ) < pivot ? 1 : 0;
		}
		lessCount = L.reduceAdd( lessCount );
		if ( lessCount > 0 && lessCount < count )
		{
			split = a + lessCount;
			// misplaced positions, both sides ascending
			int32_t* badL = s.badLPos + a;
			int32_t* badR = s.badRPos + a;
			int nl = 0, nr = 0;
			for ( int chunk = a; chunk < split; chunk += W )
			{
				int i = chunk + lane;
				bool bad = false;
				if ( i < split )
				{
					V2 c = leafCenters[i];
					bad = ( ( useX ? c.x : c.y ) < pivot ) == false;
				}
				uint32_t mask = L.ballot( bad );
				if ( bad )
					badL[nl + popCount32( mask & below )] = i;
				nl += popCount32( mask );
			}
			for ( int chunk = split; chunk < e; chunk += W )
			{
				int i = chunk + lane;
				bool bad = false;
				if ( i < e )
				{
					V2 c = leafCenters[i];
					bad = ( useX ? c.x : c.y ) < pivot;
				}
				uint32_t mask = L.ballot( bad );
				if ( bad )
					badR[nr + popCount32( mask & below )] = i;
				nr += popCount32( mask );
			}
			L.sync();
			for ( int k = lane; k < nl; k += W )
			{
				int p = badL[k], q = badR[nr - 1 - k];
				int32_t ti = leafIndices[p];
				leafIndices[p] = leafIndices[q];
				leafIndices[q] = ti;
				V2 tc = leafCenters[p];
				leafCenters[p] = leafCenters[q];
				leafCenters[q] = tc;
			}
			L.sync();
		}
	}
	if ( lane == 0 )
	{
		const int nodeIndex = s.freed[split - 1];
		TreeNode& node = nodes[nodeIndex];
		node.box = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
		node.category = 1;
		node.height = 0;
		node.flags = kNodeAllocated;
		if ( parentKey == kNull )
		{
			node.parent = kNull;
			tree.root = nodeIndex;
		}
		else
		{
			node.parent = parentKey >> 1;
			if ( parentKey & 1 )
				nodes[parentKey >> 1].child2 = nodeIndex;
			else
				nodes[parentKey >> 1].child1 = nodeIndex;
		}
		int leafChildren = 0;
		int longChildren = 0;
		int longStart[2];
		for ( int side = 0; side < 2; ++side )
		{
			const int start = side == 0 ? a : split;
			const int end = side == 0 ? split : e;
			const int n = end - start;
			if ( n == 1 )
			{
				int item = leafIndices[start];
				if ( side == 0 )
					node.child1 = item;
				else
					node.child2 = item;
				nodes[item].parent = nodeIndex;
				leafChildren += 1;
			}
			else
			{
				s.segEnd[0][start] = end;
				s.segParent[0][start] = ( nodeIndex << 1 ) | side;
				if ( n > kTreeSerialFinish )
					longStart[longChildren++] = start;
				else
					smallList[atomAdd( s.ctrl + 3, 1 )] = start;
			}
		}
		s.under[nodeIndex] = leafChildren; // arrivals so far (treeArrive)
		if ( longChildren > 0 )
		{
			atomAdd( s.ctrl + 2, longChildren );
			int slot = atomAdd( s.ctrl + 1, longChildren );
#if defined( __CUDA_ARCH__ )
			__threadfence();
#endif
			for ( int k = 0; k < longChildren; ++k )
				storeVolatile( queue + slot + k, longStart[k] );
		}
#if defined( __CUDA_ARCH__ )
		__threadfence();
#endif
		atomAdd( s.ctrl + 2, -1 );
	}
	L.sync();
}

template <class Team> F2D_HDF inline void treeRebuildTeam( World* w, Team& t, Tree& tree )
{
	if ( tree.proxyCount == 0 || tree.root == kNull )
		return;
	TreeNode* nodes = ptr( w, tree.nodes );
	// nothing was enlarged since the last rebuild: the reference collects the root as the only item and returns it
	if ( treeNodeIsDissolved( nodes[tree.root] ) == false )
		return;
	const int nodeSlots = tree.nodes.count;
	if ( tree.proxyCount > tree.leafIndices.cap || tree.proxyCount + 8 > w->shapes.cap + 8 || nodeSlots > 2 * w->shapes.cap + 16 )
	{
		if ( t.rank() == 0 )
			setError( w, kErrCapacity, __LINE__ );
		return;
	}
	TreeScratch s = treeScratch( w );
	// A team with a shared-memory work area (one block that has an SM to itself) keeps the work arrays there, as far as
	// the area reaches: every pass below reads what the pass before it wrote, and from global memory those are L2 round
	// trips (stores do not fill L1, atomics bypass it) - about a microsecond per barrier interval, ~100 intervals.
	// (only the level-synchronous build of a large team: the work-queue build keeps under[] alive to the end)
	int32_t* const arena = t.size() > kTreeQueueBuildMaxTeam ? t.arenaPtr() : nullptr;
	const int arenaInts = arena != nullptr ? t.arenaSize() : 0;
	const bool inArena = arena != nullptr && nodeSlots + 8 <= arenaInts;
	if ( inArena )
	{
		s.ctrl = arena;
		s.under = arena + 8;
	}
	int32_t* leafIndices = ptr( w, tree.leafIndices );
	V2* leafCenters = ptr( w, tree.leafCenters );
	// profile marks inside the rebuild (the leader's clock; call after a sync of the team)
	const bool timed = t.rank() == 0 && w->profEnabled;
	uint64_t clock = timed ? profClock() : 0;
	auto mark = [&]( int slot ) {
		if ( timed )
		{
			uint64_t now = profClock();
			w->prof[slot] += now - clock;
			clock = now;
		}
	};

	// ---- collect: items under each dissolved node, bottom-up. A dissolved node has two children, each an item (counts 1)
	// or a dissolved node (counts what it gathered): the first child to report parks its count in under[], the second
	// adds its own and carries the sum one level up.
	for ( int i = t.rank(); i < nodeSlots; i += t.size() )
		s.under[i] = 0;
	t.sync();
	for ( int i = t.rank(); i < nodeSlots; i += t.size() )
	{
		const TreeNode& n = nodes[i];
		if ( ( n.flags & kNodeAllocated ) == 0 || treeNodeIsDissolved( n ) )
			continue;
		int parent = n.parent;
		if ( parent == kNull || treeNodeIsDissolved( nodes[parent] ) == false )
			continue; // inside a kept subtree (the root itself is dissolved here)
		int carry = 1;
		for ( int a = parent; a != kNull; a = nodes[a].parent )
		{
			int parkedCount = atomAdd( s.under + a, carry );
			if ( parkedCount == 0 )
				break;
			carry += parkedCount;
		}
	}
	t.sync();
	mark( pfTreeCollect );
	const int itemCount = s.under[tree.root];
	int32_t* doubling = reinterpret_cast<int32_t*>( s.lo[0][0] ); // 4 node-sized arrays of the pass below
	int doublingStride = 2 * ( w->shapes.cap + 8 );
	if ( inArena )
	{
		// arrays by item position, in order of how often the level loop goes through them. under[] is dead once the
		// positions are known; segOf and scanLess, first written after that, take its place.
		const int n = itemCount + 8;
		int arenaUsed = 8;
		if ( arenaUsed + 2 * n + 1 <= arenaInts )
		{
			s.segOf = arena + arenaUsed;
			s.scanLess = arena + arenaUsed + n;
		}
		arenaUsed += maxi( nodeSlots, 2 * n + 1 );
		auto take = [&]( int32_t*& array ) {
			if ( arenaUsed + n <= arenaInts )
			{
				array = arena + arenaUsed;
				arenaUsed += n;
			}
		};
		if ( arenaUsed + 8 * n <= arenaInts )
		{
			// the centre bounds stay one block (the pointer-doubling pass borrows it)
			for ( int p = 0; p < 2; ++p )
				for ( int a = 0; a < 2; ++a )
				{
					s.lo[p][a] = reinterpret_cast<float*>( arena + arenaUsed );
					s.hi[p][a] = reinterpret_cast<float*>( arena + arenaUsed + n );
					arenaUsed += 2 * n;
				}
			if ( 4 * nodeSlots <= 8 * n )
			{
				doubling = reinterpret_cast<int32_t*>( s.lo[0][0] );
				doublingStride = nodeSlots;
			}
		}
		take( s.segEnd[0] );
		take( s.segEnd[1] );
		take( s.segSplit );
		take( s.badLPos );
		take( s.badRPos );
		take( s.segParent[0] );
		take( s.segParent[1] );
	}
	// DFS position of every item and the boundary each dissolved node used to stand for: the number of items that
	// precede a node is the sum, over the ancestors it hangs under by child2, of the items under their child1. That is
	// a suffix sum along the path to the root, computed for all nodes at once by pointer doubling (log2(height) rounds
	// of "add what my pointer has gathered, then point where it points") instead of one walk to the root per node.
	{
		const int n2 = doublingStride; // node-sized arrays carved from the (not yet used) centre-bound arrays
		int32_t* base = doubling;
		int32_t* acc[2] = { base, base + n2 };
		int32_t* anc[2] = { base + 2 * n2, base + 3 * n2 };
		const int height = nodes[tree.root].height; // of the old tree: bounds the depth of every node
		for ( int i = t.rank(); i < nodeSlots; i += t.size() )
		{
			const TreeNode& n = nodes[i];
			int up = kNull, gathered = 0;
			if ( ( n.flags & kNodeAllocated ) != 0 )
			{
				int parent = n.parent;
				if ( parent != kNull && ( treeNodeIsDissolved( n ) || treeNodeIsDissolved( nodes[parent] ) ) )
				{
					const TreeNode& pn = nodes[parent];
					up = parent;
					if ( pn.child2 == i )
						gathered = treeNodeIsDissolved( nodes[pn.child1] ) ? s.under[pn.child1] : 1;
				}
			}
			acc[0][i] = gathered;
			anc[0][i] = up;
		}
		t.sync();
		int cur = 0;
		for ( int reach = 1; reach < height + 1; reach *= 2 )
		{
			const int nxt = cur ^ 1;
			for ( int i = t.rank(); i < nodeSlots; i += t.size() )
			{
				int up = anc[cur][i];
				int gathered = acc[cur][i];
				if ( up != kNull )
				{
					gathered += acc[cur][up];
					up = anc[cur][up];
				}
				acc[nxt][i] = gathered;
				anc[nxt][i] = up;
			}
			t.sync();
			cur = nxt;
		}
		const int32_t* beforeOf = acc[cur];
		for ( int i = t.rank(); i < nodeSlots; i += t.size() )
		{
			const TreeNode& n = nodes[i];
			if ( ( n.flags & kNodeAllocated ) == 0 )
				continue;
			bool dissolved = treeNodeIsDissolved( n );
			int parent = n.parent;
			if ( dissolved == false && ( parent == kNull || treeNodeIsDissolved( nodes[parent] ) == false ) )
				continue;
			int before = beforeOf[i];
			if ( dissolved )
			{
				const TreeNode& c1 = nodes[n.child1];
				int firstHalf = treeNodeIsDissolved( c1 ) ? s.under[n.child1] : 1;
				s.freed[before + firstHalf - 1] = i;
			}
			else
			{
				V2 c = boxCenter( n.box );
				leafIndices[before] = i;
				leafCenters[before] = c;
			}
		}
	}
	t.sync();
	mark( pfTreePositions );

	// Two builds of the same tree. Small teams (a few warps per world: batches) are bound by instructions per warp and
	// use the work-queue build; large teams (one 1024-thread block or a cooperative grid per world) have the threads to
	// take every item of a level at once and use the level-synchronous build.
	if ( t.size() <= kTreeQueueBuildMaxTeam )
	{
		// ---- build: long segments by warps from a work queue, short ones by single threads, refit by last arrival
		int32_t* queue = s.scanLess; // n + 1 entries
		int32_t* smallList = s.segOf;
		for ( int i = t.rank(); i < itemCount; i += t.size() )
			queue[i] = kNull;
		if ( t.rank() == 0 )
		{
			s.ctrl[0] = 0; // next ticket
			s.ctrl[1] = 0; // queue tail
			s.ctrl[2] = 0; // long segments published and not yet split
			s.ctrl[3] = 0; // short segments
			s.segEnd[0][0] = itemCount;
			s.segParent[0][0] = kNull;
		}
		t.sync();
		if ( t.rank() == 0 )
		{
			if ( itemCount > kTreeSerialFinish )
			{
				s.ctrl[1] = 1;
				s.ctrl[2] = 1;
				queue[0] = 0;
			}
			else
			{
				s.ctrl[3] = 1;
				smallList[0] = 0;
			}
		}
		t.sync();
		{
			typename Team::Lanes L;
			while ( true )
			{
				int ticket = 0;
				if ( L.lane() == 0 )
					ticket = atomAdd( s.ctrl + 0, 1 );
				ticket = L.broadcast( ticket );
				int a = kNull;
				if ( L.lane() == 0 )
				{
					while ( true )
					{
						a = ticket < itemCount ? loadVolatile( queue + ticket ) : kNull;
						if ( a != kNull || loadVolatile( s.ctrl + 2 ) == 0 )
							break;
					}
				}
				a = L.broadcast( a );
				if ( a == kNull )
					break;
				L.fence();
				treeSplitSegment( w, tree, nodes, leafIndices, leafCenters, s, queue, smallList, a, L );
			}
		}
		t.sync();
		mark( pfTreeLevels ); // (the work-queue build: long segments)
		// short segments: one thread each builds and refits its subtree serially, then reports to the node above
		{
			const int smallCount = s.ctrl[3];
			for ( int k = t.rank(); k < smallCount; k += t.size() )
			{
				int a0 = smallList[k];
				int parentKey = s.segParent[0][a0];
				treeFinishSegment( w, tree, nodes, leafIndices, leafCenters, s.freed, s.level, a0, s.segEnd[0][a0], parentKey );
				if ( parentKey != kNull )
					treeArrive( nodes, s.under, parentKey >> 1 );
			}
		}
		t.sync();
		mark( pfTreeTail ); // (short segments and the refit by last arrival)
		return;
	}

	// root bounds of the item centres for the first level
	if ( t.rank() == 0 )
	{
		for ( int a = 0; a < 2; ++a )
		{
			s.lo[0][a][0] = FLT_MAX;
			s.hi[0][a][0] = -FLT_MAX;
		}
	}
	t.sync();
	for ( int i = t.rank(); i < itemCount; i += t.size() )
	{
		V2 c = leafCenters[i];
		atomMinF( s.lo[0][0], c.x );
		atomMinF( s.lo[0][1], c.y );
		atomMaxF( s.hi[0][0], c.x );
		atomMaxF( s.hi[0][1], c.y );
	}
	t.sync();
	// ---- build, level by level (three passes and one prefix sum per level)
	for ( int i = t.rank(); i < itemCount; i += t.size() )
		s.segOf[i] = 0;
	if ( t.rank() == 0 )
	{
		s.segEnd[0][0] = itemCount;
		s.segParent[0][0] = kNull;
		s.ctrl[0] = 1; // a segment with >= 2 items exists at this level (itemCount >= 2 because the root is dissolved)
		s.ctrl[1] = 0;
		s.ctrl[2] = itemCount; // longest segment of this level
		s.ctrl[3] = 0;
	}
	t.sync();
	int level = 0;
	// level-synchronous while some segment is long; the short tail is finished per segment (treeFinishSegment)
	// (levels are cheap when the work arrays sit in shared memory: the serial tail starts later)
	const int serialFinish = inArena ? kTreeSerialFinishArena : kTreeSerialFinish;
	while ( s.ctrl[level & 1] != 0 && s.ctrl[2 + ( level & 1 )] > serialFinish )
	{
		const int cur = level & 1, nxt = cur ^ 1;
		const int32_t* segEnd = s.segEnd[cur];
		// pass A: side of the pivot (segments of <= 2 items split in the middle without looking at the centres). The
		// bounds and loop flags of the next level, accumulated in pass C, are cleared on the way: their last readers
		// were pass A and the loop test of the level before this one.
		if ( t.rank() == 0 )
		{
			s.ctrl[nxt] = 0;
			s.ctrl[2 + nxt] = 0;
		}
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			s.lo[nxt][0][i] = s.lo[nxt][1][i] = FLT_MAX;
			s.hi[nxt][0][i] = s.hi[nxt][1][i] = -FLT_MAX;
			int a = s.segOf[i];
			int less = 0;
			if ( a != kNull && segEnd[a] - a > 2 )
			{
				float lx = s.lo[cur][0][a], ly = s.lo[cur][1][a], hx = s.hi[cur][0][a], hy = s.hi[cur][1][a];
				bool useX = ( hx - lx ) > ( hy - ly );
				float pivot = useX ? 0.5f * ( lx + hx ) : 0.5f * ( ly + hy );
				V2 c = leafCenters[i];
				less = ( useX ? c.x : c.y ) < pivot ? 1 : 0;
			}
			s.scanLess[i] = less;
		}
		t.sync();
		const int totalLess = t.exclusiveScan( s.scanLess, itemCount ); // (ends with a barrier)
		auto lessBefore = [&]( int k ) { return k < itemCount ? s.scanLess[k] : totalLess; };
		// pass B: split point per segment; the k-th misplaced item of the left part (ascending) will swap with the k-th
		// misplaced item of the right part (descending) - both ranks follow from the one prefix sum
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			if ( a == kNull )
				continue;
			int e = segEnd[a];
			int count = e - a;
			int split = a + count / 2;
			if ( count > 2 )
			{
				int L = lessBefore( e ) - lessBefore( a );
				if ( L > 0 && L < count )
				{
					split = a + L;
					int lessRank = lessBefore( i ) - lessBefore( a );
					int lessInLeft = lessBefore( split ) - lessBefore( a );
					int less = lessBefore( i + 1 ) - lessBefore( i ); // the flag itself
					if ( i < split && less == 0 )
						s.badLPos[a + ( i - a ) - lessRank] = i;
					else if ( i >= split && less != 0 )
						s.badRPos[a + ( L - lessInLeft ) - 1 - ( lessRank - lessInLeft )] = i;
				}
			}
			if ( i == a )
				s.segSplit[a] = split;
		}
		t.sync();
		// pass C: swap, create the node of every segment, move every position to its child segment (or retire it) and
		// accumulate the centre bounds of the child segments
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			if ( a == kNull )
				continue;
			int e = segEnd[a];
			int m = s.segSplit[a];
			int count = e - a;
			int bad = 0;
			bool isBad = false;
			if ( count > 2 )
			{
				int L = lessBefore( e ) - lessBefore( a );
				if ( L > 0 && L < count )
				{
					bad = L - ( lessBefore( m ) - lessBefore( a ) );
					int less = lessBefore( i + 1 ) - lessBefore( i ); // the flag itself
					isBad = ( i < m && less == 0 ) || ( i >= m && less != 0 );
				}
			}
			const bool leftGrows = m - a > 2, rightGrows = e - m > 2;
			if ( i - a < bad )
			{
				int p = s.badLPos[i], q = s.badRPos[i];
				int32_t ti = leafIndices[p];
				leafIndices[p] = leafIndices[q];
				leafIndices[q] = ti;
				V2 cp = leafCenters[q], cq = leafCenters[p]; // centres after the swap
				leafCenters[p] = cp;
				leafCenters[q] = cq;
				if ( leftGrows )
				{
					atomMinF( s.lo[nxt][0] + a, cp.x );
					atomMinF( s.lo[nxt][1] + a, cp.y );
					atomMaxF( s.hi[nxt][0] + a, cp.x );
					atomMaxF( s.hi[nxt][1] + a, cp.y );
				}
				if ( rightGrows )
				{
					atomMinF( s.lo[nxt][0] + m, cq.x );
					atomMinF( s.lo[nxt][1] + m, cq.y );
					atomMaxF( s.hi[nxt][0] + m, cq.x );
					atomMaxF( s.hi[nxt][1] + m, cq.y );
				}
			}
			int nodeIndex = s.freed[m - 1];
			const bool left = i < m;
			if ( isBad == false && ( left ? leftGrows : rightGrows ) )
			{
				int b = left ? a : m;
				V2 c = leafCenters[i];
				atomMinF( s.lo[nxt][0] + b, c.x );
				atomMinF( s.lo[nxt][1] + b, c.y );
				atomMaxF( s.hi[nxt][0] + b, c.x );
				atomMaxF( s.hi[nxt][1] + b, c.y );
			}
			if ( left )
			{
				bool single = m - a == 1;
				s.segOf[i] = single ? kNull : a;
				if ( single )
					s.posParent[i] = ( nodeIndex << 1 ) | 0;
			}
			else
			{
				bool single = e - m == 1;
				s.segOf[i] = single ? kNull : m;
				if ( single )
					s.posParent[i] = ( nodeIndex << 1 ) | 1;
			}
			if ( i == a )
			{
				s.level[m - 1] = level;
				TreeNode& node = nodes[nodeIndex];
				node.box = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
				node.category = 1;
				node.height = 0;
				node.flags = kNodeAllocated;
				int parentKey = s.segParent[cur][a];
				if ( parentKey == kNull )
				{
					node.parent = kNull;
					tree.root = nodeIndex;
				}
				else
				{
					node.parent = parentKey >> 1;
					if ( parentKey & 1 )
						nodes[parentKey >> 1].child2 = nodeIndex;
					else
						nodes[parentKey >> 1].child1 = nodeIndex;
				}
				if ( m - a > 1 )
				{
					s.segEnd[nxt][a] = m;
					s.segParent[nxt][a] = ( nodeIndex << 1 ) | 0;
				}
				if ( e - m > 1 )
				{
					s.segEnd[nxt][m] = e;
					s.segParent[nxt][m] = ( nodeIndex << 1 ) | 1;
				}
				if ( m - a > 1 || e - m > 1 )
				{
					s.ctrl[nxt] = 1;
					atomMax32( s.ctrl + 2 + nxt, maxi( m - a, e - m ) );
				}
			}
		}
		t.sync();
		level += 1;
	}
	mark( pfTreeLevels );
	// short segments that are still open: one thread each builds the rest of its subtree serially
	if ( s.ctrl[level & 1] != 0 )
	{
		const int cur = level & 1;
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			if ( s.segOf[i] == i )
				treeFinishSegment( w, tree, nodes, leafIndices, leafCenters, s.freed, s.level, i, s.segEnd[cur][i], s.segParent[cur][i] );
		}
	}
	// every other position has retired under some node: hang the items
	for ( int i = t.rank(); i < itemCount; i += t.size() )
	{
		if ( s.segOf[i] != kNull )
			continue;
		int key = s.posParent[i];
		int item = leafIndices[i];
		nodes[item].parent = key >> 1;
		if ( key & 1 )
			nodes[key >> 1].child2 = item;
		else
			nodes[key >> 1].child1 = item;
	}
	t.sync();
	mark( pfTreeTail );

	// ---- refit bottom-up, one level at a time
	for ( int d = level - 1; d >= 0; --d )
	{
		for ( int j = t.rank(); j < itemCount - 1; j += t.size() )
		{
			if ( s.level[j] != d )
				continue;
			TreeNode& node = nodes[s.freed[j]];
			const TreeNode& c1 = nodes[node.child1];
			const TreeNode& c2 = nodes[node.child2];
			node.box = boxUnion( c1.box, c2.box );
			node.height = (uint16_t)( 1 + maxU16( c1.height, c2.height ) );
			node.category = c1.category | c2.category;
		}
		t.sync();
	}
	mark( pfTreeRefit );
}

} // namespace f2d
