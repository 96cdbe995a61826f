// forge2d_b200 — team-parallel partial rebuild of the broadphase BVH.
//
// Produces exactly the tree of the reference's serial b2DynamicTree_Rebuild (B2/src/dynamic_tree.c:1872-1989 with
// b2BuildTree :1716-1869 and b2PartitionMid :1426-1532) — same items, same item order, same split at every node — but
// as data-parallel passes over the items:
//   collect   every node whose parent was dissolved (enlarged) is an item; its position in the reference's child1-first
//             DFS is the number of items that precede it = sum over its ancestors, when it hangs under child2, of the
//             item count under child1. Counts come from one walk to the root per item (integer atomics), ranks from
//             a second walk. No traversal stack, no serial DFS.
//   build     level-synchronous top-down median split. The reference's Hoare partition has a closed form: with L =
//             #(centre < pivot), the k-th misplaced item of the left part (ascending) swaps with the k-th misplaced item
//             of the right part (descending); degenerate splits (L == 0 or L == n) leave the order untouched and cut
//             at n/2. Flags + prefix sums give every swap pair directly, all segments of a level at once.
//   refit     boxes, heights and categories bottom-up, one pass per level.
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
	int32_t* segOf;		// [n] start index of the segment a position belongs to
	int32_t* segEnd;	// [n] by segment start
	int32_t* segParent; // [n] by segment start: (parent node << 1) | side, kNull for the root segment
	int32_t* segSplit;	// [n] by segment start
	int32_t* scanLess;	// [n+1]
	int32_t* scanBadL;	// [n+1]
	int32_t* scanBadR;	// [n+1]
	int32_t* badLPos;	// [n]
	int32_t* badRPos;	// [n]
	int32_t* freed;		// [n] dissolved node recycled for boundary m at freed[m-1]
	int32_t* level;		// [n] build depth of the node at boundary m, at level[m-1]
	float* loX;			// [n] by segment start: bounds of the item centres
	float* loY;
	float* hiX;
	float* hiY;
	int32_t* under; // [nodeCap] items below a dissolved node, by node id
	int32_t* ctrl;	// [8] loop control
};

F2D_HD int treeScratchInts( int shapeCap ) { return 17 * ( shapeCap + 8 ) + 2 * shapeCap + 16 + 8; }

F2D_HD TreeScratch treeScratch( World* w, const Tree& t )
{
	int n = w->shapes.cap + 8;
	int32_t* base = ptr( w, w->treeScratch );
	TreeScratch s;
	s.segOf = base;
	s.segEnd = base + n;
	s.segParent = base + 2 * n;
	s.segSplit = base + 3 * n;
	s.scanLess = base + 4 * n;
	s.scanBadL = base + 5 * n;
	s.scanBadR = base + 6 * n;
	s.badLPos = base + 7 * n;
	s.badRPos = base + 8 * n;
	s.freed = base + 9 * n;
	s.level = base + 10 * n;
	s.loX = reinterpret_cast<float*>( base + 11 * n );
	s.loY = reinterpret_cast<float*>( base + 12 * n );
	s.hiX = reinterpret_cast<float*>( base + 13 * n );
	s.hiY = reinterpret_cast<float*>( base + 14 * n );
	s.ctrl = base + 15 * n;
	s.under = base + 15 * n + 8;
	(void)t;
	return s;
}

// An item of the rebuild: a leaf, or an internal node that was not enlarged, directly below a dissolved node.
F2D_HD bool treeNodeIsDissolved( const TreeNode& n ) { return n.height > 0 && ( n.flags & kNodeEnlarged ) != 0; }

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
	TreeScratch s = treeScratch( w, tree );
	int32_t* leafIndices = ptr( w, tree.leafIndices );
	V2* leafCenters = ptr( w, tree.leafCenters );

	// ---- collect: items under each dissolved node
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
		for ( int a = parent; a != kNull; a = nodes[a].parent )
			atomAdd( s.under + a, 1 );
	}
	t.sync();
	const int itemCount = s.under[tree.root];
	// DFS position of every item and the boundary each dissolved node used to stand for
	for ( int i = t.rank(); i < nodeSlots; i += t.size() )
	{
		const TreeNode& n = nodes[i];
		if ( ( n.flags & kNodeAllocated ) == 0 )
			continue;
		bool dissolved = treeNodeIsDissolved( n );
		int parent = n.parent;
		if ( dissolved == false && ( parent == kNull || treeNodeIsDissolved( nodes[parent] ) == false ) )
			continue;
		int before = 0;
		int child = i;
		for ( int a = parent; a != kNull; a = nodes[a].parent )
		{
			const TreeNode& an = nodes[a];
			if ( an.child2 == child )
			{
				const TreeNode& c1 = nodes[an.child1];
				before += treeNodeIsDissolved( c1 ) ? s.under[an.child1] : 1;
			}
			child = a;
		}
		if ( dissolved )
		{
			const TreeNode& c1 = nodes[n.child1];
			int firstHalf = treeNodeIsDissolved( c1 ) ? s.under[n.child1] : 1;
			s.freed[before + firstHalf - 1] = i;
		}
		else
		{
			leafIndices[before] = i;
			leafCenters[before] = boxCenter( n.box );
		}
	}
	t.sync();

	// ---- build, level by level
	for ( int i = t.rank(); i < itemCount; i += t.size() )
		s.segOf[i] = 0;
	if ( t.rank() == 0 )
	{
		s.segEnd[0] = itemCount;
		s.segParent[0] = kNull;
		s.loX[0] = s.loY[0] = FLT_MAX;
		s.hiX[0] = s.hiY[0] = -FLT_MAX;
		s.ctrl[0] = 1; // a segment with >= 2 items exists at this level (itemCount >= 2 because the root is dissolved)
		s.ctrl[1] = 0;
	}
	t.sync();
	int level = 0;
	while ( s.ctrl[level & 1] != 0 )
	{
		// P1: centre bounds per segment (segments of 2 split without looking at the centres)
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			if ( a == kNull || s.segEnd[a] - a <= 2 )
				continue;
			V2 c = leafCenters[i];
			atomMinF( s.loX + a, c.x );
			atomMinF( s.loY + a, c.y );
			atomMaxF( s.hiX + a, c.x );
			atomMaxF( s.hiY + a, c.y );
		}
		if ( t.rank() == 0 )
			s.ctrl[( level + 1 ) & 1] = 0;
		t.sync();
		// P2: which side of the pivot
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			int less = 0;
			if ( a != kNull && s.segEnd[a] - a > 2 )
			{
				float lx = s.loX[a], ly = s.loY[a], hx = s.hiX[a], hy = s.hiY[a];
				bool useX = ( hx - lx ) > ( hy - ly );
				float pivot = useX ? 0.5f * ( lx + hx ) : 0.5f * ( ly + hy );
				V2 c = leafCenters[i];
				less = ( useX ? c.x : c.y ) < pivot ? 1 : 0;
			}
			s.scanLess[i] = less;
			s.badLPos[i] = less; // kept for P3 (scanLess is overwritten by its prefix sums)
		}
		t.sync();
		int totalLess = t.exclusiveScan( s.scanLess, itemCount );
		if ( t.rank() == 0 )
			s.scanLess[itemCount] = totalLess;
		t.sync();
		// P3: split point per segment, misplaced items
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			int badL = 0, badR = 0;
			if ( a != kNull )
			{
				int e = s.segEnd[a];
				int count = e - a;
				int split = a + count / 2;
				if ( count > 2 )
				{
					int L = s.scanLess[e] - s.scanLess[a];
					if ( L > 0 && L < count )
					{
						split = a + L;
						int less = s.badLPos[i];
						badL = ( i < split && less == 0 ) ? 1 : 0;
						badR = ( i >= split && less != 0 ) ? 1 : 0;
					}
				}
				if ( i == a )
					s.segSplit[a] = split;
			}
			s.scanBadL[i] = badL;
			s.scanBadR[i] = badR;
		}
		t.sync();
		int totalBadL = t.exclusiveScan( s.scanBadL, itemCount );
		int totalBadR = t.exclusiveScan( s.scanBadR, itemCount );
		if ( t.rank() == 0 )
		{
			s.scanBadL[itemCount] = totalBadL;
			s.scanBadR[itemCount] = totalBadR;
		}
		t.sync();
		if ( totalBadL > 0 )
		{
			// P4: k-th misplaced-left (ascending) pairs with k-th misplaced-right (descending)
			for ( int i = t.rank(); i < itemCount; i += t.size() )
			{
				int a = s.segOf[i];
				if ( a == kNull )
					continue;
				int e = s.segEnd[a];
				bool isBadL = s.scanBadL[i + 1] != s.scanBadL[i];
				bool isBadR = s.scanBadR[i + 1] != s.scanBadR[i];
				if ( isBadL )
					s.badLPos[a + ( s.scanBadL[i] - s.scanBadL[a] )] = i;
				if ( isBadR )
				{
					int bad = s.scanBadR[e] - s.scanBadR[a];
					s.badRPos[a + ( bad - 1 - ( s.scanBadR[i] - s.scanBadR[a] ) )] = i;
				}
			}
			t.sync();
			// P5: swap
			for ( int i = t.rank(); i < itemCount; i += t.size() )
			{
				int a = s.segOf[i];
				if ( a == kNull )
					continue;
				int bad = s.scanBadL[s.segEnd[a]] - s.scanBadL[a];
				if ( i - a >= bad )
					continue;
				int p = s.badLPos[i], q = s.badRPos[i];
				int32_t ti = leafIndices[p];
				leafIndices[p] = leafIndices[q];
				leafIndices[q] = ti;
				V2 tc = leafCenters[p];
				leafCenters[p] = leafCenters[q];
				leafCenters[q] = tc;
			}
			t.sync();
		}
		// P6: one internal node per segment; children are items (segments of one) or next-level segments
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			if ( a != i )
				continue;
			int e = s.segEnd[a];
			int m = s.segSplit[a];
			int nodeIndex = s.freed[m - 1];
			s.level[m - 1] = level;
			TreeNode& node = nodes[nodeIndex];
			node.box = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
			node.category = 1;
			node.height = 0;
			node.flags = kNodeAllocated;
			int parentKey = s.segParent[a];
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
			bool more = false;
			if ( m - a == 1 )
			{
				int item = leafIndices[a];
				node.child1 = item;
				nodes[item].parent = nodeIndex;
			}
			else
			{
				s.segEnd[a] = m;
				s.segParent[a] = ( nodeIndex << 1 ) | 0;
				s.loX[a] = s.loY[a] = FLT_MAX;
				s.hiX[a] = s.hiY[a] = -FLT_MAX;
				more = true;
			}
			s.segEnd[m] = e;
			if ( e - m == 1 )
			{
				int item = leafIndices[m];
				node.child2 = item;
				nodes[item].parent = nodeIndex;
			}
			else
			{
				s.segParent[m] = ( nodeIndex << 1 ) | 1;
				s.loX[m] = s.loY[m] = FLT_MAX;
				s.hiX[m] = s.hiY[m] = -FLT_MAX;
				more = true;
			}
			if ( more )
				s.ctrl[( level + 1 ) & 1] = 1;
		}
		t.sync();
		// P7: positions move to their new segment (or retire when they became a child directly)
		for ( int i = t.rank(); i < itemCount; i += t.size() )
		{
			int a = s.segOf[i];
			if ( a == kNull )
				continue;
			int m = s.segSplit[a];
			if ( i < m )
				s.segOf[i] = ( m - a == 1 ) ? kNull : a;
			else
				s.segOf[i] = ( s.segEnd[m] - m == 1 ) ? kNull : m;
		}
		t.sync();
		level += 1;
	}

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
}

} // namespace f2d
