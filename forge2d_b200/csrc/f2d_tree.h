// forge2d_b200 — broadphase BVH with the reference's exact topology rules.
//
// The pair ORDER the reference emits depends on the history-dependent shape of its b2DynamicTree
// (SURVEY §9.1 O5, O16), so insertion (greedy SAH descent + rotations), removal, enlargement and the partial
// median-split rebuild are reproduced decision for decision; node storage, the free list and traversal are ours.
// Reference: B2/src/dynamic_tree.c (cited per function).
#pragma once
#include "f2d_types.h"

namespace f2d
{

constexpr int kTreeStack = 1024; // dynamic_tree.c:14 (rebuild stacks, stored in Tree::work)
constexpr int kQueryStack = 128; // per-thread traversal stack; deeper trees raise kErrTreeStack instead of asserting

// Free nodes form a LIFO list through `parent`; never-used slots are handed out in increasing order, which is
// exactly the order the reference's growing pool produces (dynamic_tree.c:122-157).
F2D_HDF inline int treeAllocNode( World* w, Tree& t )
{
	TreeNode* nodes = ptr( w, t.nodes );
	int id;
	if ( t.freeList != kNull )
	{
		id = t.freeList;
		t.freeList = nodes[id].parent;
	}
	else
	{
		if ( t.nodes.count >= t.nodes.cap )
		{
			setError( w, kErrCapacity, __LINE__ );
			return t.nodes.cap - 1;
		}
		id = t.nodes.count++;
	}
	TreeNode& n = nodes[id];
	n.box = Box{ { 0.0f, 0.0f }, { 0.0f, 0.0f } };
	n.category = 1; // B2_DEFAULT_CATEGORY_BITS
	n.child1 = kNull;
	n.child2 = kNull;
	n.parent = kNull;
	n.height = 0;
	n.flags = kNodeAllocated;
	t.nodeCount += 1;
	return id;
}

F2D_HD void treeFreeNode( World* w, Tree& t, int id ) // dynamic_tree.c:160-168
{
	TreeNode* nodes = ptr( w, t.nodes );
	nodes[id].parent = t.freeList;
	nodes[id].flags = 0;
	t.freeList = id;
	t.nodeCount -= 1;
}

F2D_HD uint16_t maxU16( uint16_t a, uint16_t b ) { return a > b ? a : b; }

// Greedy SAH sibling search: dynamic_tree.c:187-325
F2D_HDF inline int treeFindBestSibling( const TreeNode* nodes, int rootIndex, Box boxD )
{
	V2 centerD = boxCenter( boxD );
	float areaD = boxPerimeter( boxD );

	Box rootBox = nodes[rootIndex].box;
	float areaBase = boxPerimeter( rootBox );
	float directCost = boxPerimeter( boxUnion( rootBox, boxD ) );
	float inheritedCost = 0.0f;

	int bestSibling = rootIndex;
	float bestCost = directCost;

	int index = rootIndex;
	while ( nodes[index].height > 0 )
	{
		int child1 = nodes[index].child1;
		int child2 = nodes[index].child2;

		float cost = directCost + inheritedCost;
		if ( cost < bestCost )
		{
			bestSibling = index;
			bestCost = cost;
		}

		inheritedCost += directCost - areaBase;

		bool leaf1 = nodes[child1].height == 0;
		bool leaf2 = nodes[child2].height == 0;

		float lowerCost1 = FLT_MAX;
		Box box1 = nodes[child1].box;
		float directCost1 = boxPerimeter( boxUnion( box1, boxD ) );
		float area1 = 0.0f;
		if ( leaf1 )
		{
			float cost1 = directCost1 + inheritedCost;
			if ( cost1 < bestCost )
			{
				bestSibling = child1;
				bestCost = cost1;
			}
		}
		else
		{
			area1 = boxPerimeter( box1 );
			lowerCost1 = inheritedCost + directCost1 + minf( areaD - area1, 0.0f );
		}

		float lowerCost2 = FLT_MAX;
		Box box2 = nodes[child2].box;
		float directCost2 = boxPerimeter( boxUnion( box2, boxD ) );
		float area2 = 0.0f;
		if ( leaf2 )
		{
			float cost2 = directCost2 + inheritedCost;
			if ( cost2 < bestCost )
			{
				bestSibling = child2;
				bestCost = cost2;
			}
		}
		else
		{
			area2 = boxPerimeter( box2 );
			lowerCost2 = inheritedCost + directCost2 + minf( areaD - area2, 0.0f );
		}

		if ( leaf1 && leaf2 )
			break;
		if ( bestCost <= lowerCost1 && bestCost <= lowerCost2 )
			break;

		if ( lowerCost1 == lowerCost2 && leaf1 == false )
		{
			V2 d1 = sub( boxCenter( box1 ), centerD );
			V2 d2 = sub( boxCenter( box2 ), centerD );
			lowerCost1 = lengthSq( d1 );
			lowerCost2 = lengthSq( d2 );
		}

		if ( lowerCost1 < lowerCost2 && leaf1 == false )
		{
			index = child1;
			areaBase = area1;
			directCost = directCost1;
		}
		else
		{
			index = child2;
			areaBase = area2;
			directCost = directCost2;
		}
	}
	return bestSibling;
}

// One tree rotation: grandchild `gc` (child slot gcSlot of `lower`) swaps with `upperChild` (child slot upSlot of A).
// Shared tail of the four cases in dynamic_tree.c:338-623.
F2D_HD void treeSwap( TreeNode* nodes, int iA, int upSlot, int iLower, int lowSlot, int iUp, int iGc, int iKeep, Box newLowerBox )
{
	TreeNode& A = nodes[iA];
	TreeNode& L = nodes[iLower];
	TreeNode& U = nodes[iUp];
	TreeNode& G = nodes[iGc];
	TreeNode& K = nodes[iKeep];
	if ( upSlot == 1 )
		A.child1 = iGc;
	else
		A.child2 = iGc;
	if ( lowSlot == 1 )
		L.child1 = iUp;
	else
		L.child2 = iUp;
	U.parent = iLower;
	G.parent = iA;
	L.box = newLowerBox;
	L.height = (uint16_t)( 1 + maxU16( U.height, K.height ) );
	A.height = (uint16_t)( 1 + maxU16( L.height, G.height ) );
	L.category = U.category | K.category;
	A.category = L.category | G.category;
	L.flags |= ( U.flags | K.flags ) & kNodeEnlarged;
	A.flags |= ( L.flags | G.flags ) & kNodeEnlarged;
}

// dynamic_tree.c:338-623 b2RotateNodes
F2D_HDF inline void treeRotate( World* w, Tree& t, int iA )
{
	TreeNode* nodes = ptr( w, t.nodes );
	TreeNode& A = nodes[iA];
	if ( A.height < 2 )
		return;
	int iB = A.child1, iC = A.child2;
	TreeNode& B = nodes[iB];
	TreeNode& C = nodes[iC];

	if ( B.height == 0 )
	{
		// B leaf, C internal: candidates swap B<->F (C keeps G) or B<->G (C keeps F)
		int iF = C.child1, iG = C.child2;
		float costBase = boxPerimeter( C.box );
		Box boxBG = boxUnion( B.box, nodes[iG].box );
		float costBF = boxPerimeter( boxBG );
		Box boxBF = boxUnion( B.box, nodes[iF].box );
		float costBG = boxPerimeter( boxBF );
		if ( costBase < costBF && costBase < costBG )
			return;
		if ( costBF < costBG )
			treeSwap( nodes, iA, 1, iC, 1, iB, iF, iG, boxBG );
		else
			treeSwap( nodes, iA, 1, iC, 2, iB, iG, iF, boxBF );
	}
	else if ( C.height == 0 )
	{
		int iD = B.child1, iE = B.child2;
		float costBase = boxPerimeter( B.box );
		Box boxCE = boxUnion( C.box, nodes[iE].box );
		float costCD = boxPerimeter( boxCE );
		Box boxCD = boxUnion( C.box, nodes[iD].box );
		float costCE = boxPerimeter( boxCD );
		if ( costBase < costCD && costBase < costCE )
			return;
		if ( costCD < costCE )
			treeSwap( nodes, iA, 2, iB, 1, iC, iD, iE, boxCE );
		else
			treeSwap( nodes, iA, 2, iB, 2, iC, iE, iD, boxCD );
	}
	else
	{
		int iD = B.child1, iE = B.child2, iF = C.child1, iG = C.child2;
		float areaB = boxPerimeter( B.box );
		float areaC = boxPerimeter( C.box );
		float costBase = areaB + areaC;
		int best = 0;
		float bestCost = costBase;

		Box boxBG = boxUnion( B.box, nodes[iG].box );
		float costBF = areaB + boxPerimeter( boxBG );
		if ( costBF < bestCost )
		{
			best = 1;
			bestCost = costBF;
		}
		Box boxBF = boxUnion( B.box, nodes[iF].box );
		float costBG = areaB + boxPerimeter( boxBF );
		if ( costBG < bestCost )
		{
			best = 2;
			bestCost = costBG;
		}
		Box boxCE = boxUnion( C.box, nodes[iE].box );
		float costCD = areaC + boxPerimeter( boxCE );
		if ( costCD < bestCost )
		{
			best = 3;
			bestCost = costCD;
		}
		Box boxCD = boxUnion( C.box, nodes[iD].box );
		float costCE = areaC + boxPerimeter( boxCD );
		if ( costCE < bestCost )
		{
			best = 4;
		}
		switch ( best )
		{
			case 1:
				treeSwap( nodes, iA, 1, iC, 1, iB, iF, iG, boxBG );
				break;
			case 2:
				treeSwap( nodes, iA, 1, iC, 2, iB, iG, iF, boxBF );
				break;
			case 3:
				treeSwap( nodes, iA, 2, iB, 1, iC, iD, iE, boxCE );
				break;
			case 4:
				treeSwap( nodes, iA, 2, iB, 2, iC, iE, iD, boxCD );
				break;
			default:
				break;
		}
	}
}

// dynamic_tree.c:625-699
F2D_HDF inline void treeInsertLeaf( World* w, Tree& t, int leaf, bool shouldRotate )
{
	TreeNode* nodes = ptr( w, t.nodes );
	if ( t.root == kNull )
	{
		t.root = leaf;
		nodes[leaf].parent = kNull;
		return;
	}
	Box leafBox = nodes[leaf].box;
	int sibling = treeFindBestSibling( nodes, t.root, leafBox );

	int oldParent = nodes[sibling].parent;
	int newParent = treeAllocNode( w, t );
	nodes[newParent].parent = oldParent;
	nodes[newParent].userData = UINT64_MAX;
	nodes[newParent].box = boxUnion( leafBox, nodes[sibling].box );
	nodes[newParent].category = nodes[leaf].category | nodes[sibling].category;
	nodes[newParent].height = (uint16_t)( nodes[sibling].height + 1 );

	if ( oldParent != kNull )
	{
		if ( nodes[oldParent].child1 == sibling )
			nodes[oldParent].child1 = newParent;
		else
			nodes[oldParent].child2 = newParent;
	}
	else
	{
		t.root = newParent;
	}
	nodes[newParent].child1 = sibling;
	nodes[newParent].child2 = leaf;
	nodes[sibling].parent = newParent;
	nodes[leaf].parent = newParent;

	int index = nodes[leaf].parent;
	while ( index != kNull )
	{
		int c1 = nodes[index].child1, c2 = nodes[index].child2;
		nodes[index].box = boxUnion( nodes[c1].box, nodes[c2].box );
		nodes[index].category = nodes[c1].category | nodes[c2].category;
		nodes[index].height = (uint16_t)( 1 + maxU16( nodes[c1].height, nodes[c2].height ) );
		nodes[index].flags |= ( nodes[c1].flags | nodes[c2].flags ) & kNodeEnlarged;
		if ( shouldRotate )
			treeRotate( w, t, index );
		index = nodes[index].parent;
	}
}

// dynamic_tree.c:701-766
F2D_HDF inline void treeRemoveLeaf( World* w, Tree& t, int leaf )
{
	TreeNode* nodes = ptr( w, t.nodes );
	if ( leaf == t.root )
	{
		t.root = kNull;
		return;
	}
	int parent = nodes[leaf].parent;
	int grand = nodes[parent].parent;
	int sibling = nodes[parent].child1 == leaf ? nodes[parent].child2 : nodes[parent].child1;
	if ( grand != kNull )
	{
		if ( nodes[grand].child1 == parent )
			nodes[grand].child1 = sibling;
		else
			nodes[grand].child2 = sibling;
		nodes[sibling].parent = grand;
		treeFreeNode( w, t, parent );
		int index = grand;
		while ( index != kNull )
		{
			TreeNode& n = nodes[index];
			const TreeNode& a = nodes[n.child1];
			const TreeNode& b = nodes[n.child2];
			n.box = boxUnion( a.box, b.box );
			n.category = a.category | b.category;
			n.height = (uint16_t)( 1 + maxU16( a.height, b.height ) );
			index = n.parent;
		}
	}
	else
	{
		t.root = sibling;
		nodes[sibling].parent = kNull;
		treeFreeNode( w, t, parent );
	}
}

// dynamic_tree.c:770-792
F2D_HDF inline int treeCreateProxy( World* w, Tree& t, Box box, uint64_t category, uint64_t userData )
{
	int id = treeAllocNode( w, t );
	TreeNode& n = ptr( w, t.nodes )[id];
	n.box = box;
	n.userData = userData;
	n.category = category;
	n.height = 0;
	n.flags = kNodeAllocated | kNodeLeaf;
	treeInsertLeaf( w, t, id, true );
	t.proxyCount += 1;
	return id;
}

// dynamic_tree.c:794-804
F2D_HDF inline void treeDestroyProxy( World* w, Tree& t, int id )
{
	treeRemoveLeaf( w, t, id );
	treeFreeNode( w, t, id );
	t.proxyCount -= 1;
}

// dynamic_tree.c:811-825
F2D_HDF inline void treeMoveProxy( World* w, Tree& t, int id, Box box )
{
	treeRemoveLeaf( w, t, id );
	ptr( w, t.nodes )[id].box = box;
	treeInsertLeaf( w, t, id, false );
}

// dynamic_tree.c:827-866 (serial form)
F2D_HDF inline void treeEnlargeProxy( World* w, Tree& t, int id, Box box )
{
	TreeNode* nodes = ptr( w, t.nodes );
	nodes[id].box = box;
	int parent = nodes[id].parent;
	while ( parent != kNull )
	{
		bool changed = boxEnlarge( &nodes[parent].box, box );
		nodes[parent].flags |= kNodeEnlarged;
		parent = nodes[parent].parent;
		if ( changed == false )
			break;
	}
	while ( parent != kNull )
	{
		if ( nodes[parent].flags & kNodeEnlarged )
			break;
		nodes[parent].flags |= kNodeEnlarged;
		parent = nodes[parent].parent;
	}
}

// Median split about the centre of the centroid bounds, Hoare partition: dynamic_tree.c:1426-1532
F2D_HDF inline int treePartitionMid( int32_t* indices, V2* centers, int count )
{
	if ( count <= 2 )
		return count / 2;
	V2 lo = centers[0], hi = centers[0];
	for ( int i = 1; i < count; ++i )
	{
		lo = vmin( lo, centers[i] );
		hi = vmax( hi, centers[i] );
	}
	V2 d = sub( hi, lo );
	V2 c = { 0.5f * ( lo.x + hi.x ), 0.5f * ( lo.y + hi.y ) };
	int i1 = 0, i2 = count;
	bool useX = d.x > d.y;
	float pivot = useX ? c.x : c.y;
	while ( i1 < i2 )
	{
		while ( i1 < i2 && ( useX ? centers[i1].x : centers[i1].y ) < pivot )
			i1 += 1;
		while ( i1 < i2 && ( useX ? centers[i2 - 1].x : centers[i2 - 1].y ) >= pivot )
			i2 -= 1;
		if ( i1 < i2 )
		{
			int32_t ti = indices[i1];
			indices[i1] = indices[i2 - 1];
			indices[i2 - 1] = ti;
			V2 tc = centers[i1];
			centers[i1] = centers[i2 - 1];
			centers[i2 - 1] = tc;
			i1 += 1;
			i2 -= 1;
		}
	}
	if ( i1 > 0 && i1 < count )
		return i1;
	return count / 2;
}

struct RebuildItem
{
	int32_t nodeIndex, childCount, startIndex, splitIndex, endIndex;
};

// Top-down build over the collected items, explicit stack: dynamic_tree.c:1716-1869
F2D_HDF inline int treeBuild( World* w, Tree& t, int leafCount )
{
	TreeNode* nodes = ptr( w, t.nodes );
	int32_t* leafIndices = ptr( w, t.leafIndices );
	V2* leafCenters = ptr( w, t.leafCenters );
	if ( leafCount == 1 )
	{
		nodes[leafIndices[0]].parent = kNull;
		return leafIndices[0];
	}
	RebuildItem* stack = reinterpret_cast<RebuildItem*>( ptr( w, t.work ) + kTreeStack );
	int top = 0;
	stack[0].nodeIndex = treeAllocNode( w, t );
	stack[0].childCount = -1;
	stack[0].startIndex = 0;
	stack[0].endIndex = leafCount;
	stack[0].splitIndex = treePartitionMid( leafIndices, leafCenters, leafCount );

	while ( true )
	{
		RebuildItem* item = stack + top;
		item->childCount += 1;
		if ( item->childCount == 2 )
		{
			if ( top == 0 )
				break;
			RebuildItem* parentItem = stack + ( top - 1 );
			TreeNode& parentNode = nodes[parentItem->nodeIndex];
			if ( parentItem->childCount == 0 )
				parentNode.child1 = item->nodeIndex;
			else
				parentNode.child2 = item->nodeIndex;
			TreeNode& node = nodes[item->nodeIndex];
			node.parent = parentItem->nodeIndex;
			const TreeNode& c1 = nodes[node.child1];
			const TreeNode& c2 = nodes[node.child2];
			node.box = boxUnion( c1.box, c2.box );
			node.height = (uint16_t)( 1 + maxU16( c1.height, c2.height ) );
			node.category = c1.category | c2.category;
			top -= 1;
		}
		else
		{
			int startIndex, endIndex;
			if ( item->childCount == 0 )
			{
				startIndex = item->startIndex;
				endIndex = item->splitIndex;
			}
			else
			{
				startIndex = item->splitIndex;
				endIndex = item->endIndex;
			}
			int count = endIndex - startIndex;
			if ( count == 1 )
			{
				int childIndex = leafIndices[startIndex];
				TreeNode& node = nodes[item->nodeIndex];
				if ( item->childCount == 0 )
					node.child1 = childIndex;
				else
					node.child2 = childIndex;
				nodes[childIndex].parent = item->nodeIndex;
			}
			else
			{
				if ( top + 1 >= kTreeStack )
				{
					setError( w, kErrTreeStack, __LINE__ );
					break;
				}
				top += 1;
				RebuildItem* ni = stack + top;
				ni->nodeIndex = treeAllocNode( w, t );
				ni->childCount = -1;
				ni->startIndex = startIndex;
				ni->endIndex = endIndex;
				ni->splitIndex = treePartitionMid( leafIndices + startIndex, leafCenters + startIndex, count ) + startIndex;
			}
		}
	}
	TreeNode& root = nodes[stack[0].nodeIndex];
	const TreeNode& c1 = nodes[root.child1];
	const TreeNode& c2 = nodes[root.child2];
	root.box = boxUnion( c1.box, c2.box );
	root.height = (uint16_t)( 1 + maxU16( c1.height, c2.height ) );
	root.category = c1.category | c2.category;
	return stack[0].nodeIndex;
}

// Partial rebuild: enlarged internal nodes are dissolved, everything else is an item: dynamic_tree.c:1872-1989
F2D_HDF inline int treeRebuild( World* w, Tree& t, bool fullBuild )
{
	if ( t.proxyCount == 0 )
		return 0;
	if ( t.proxyCount > t.leafIndices.cap || t.proxyCount > t.leafCenters.cap )
	{
		setError( w, kErrCapacity, __LINE__ );
		return 0;
	}
	TreeNode* nodes = ptr( w, t.nodes );
	int32_t* leafIndices = ptr( w, t.leafIndices );
	V2* leafCenters = ptr( w, t.leafCenters );
	int leafCount = 0;
	int32_t* stack = ptr( w, t.work );
	int sp = 0;
	int nodeIndex = t.root;
	while ( true )
	{
		TreeNode& node = nodes[nodeIndex];
		if ( node.height == 0 || ( ( node.flags & kNodeEnlarged ) == 0 && fullBuild == false ) )
		{
			leafIndices[leafCount] = nodeIndex;
			leafCenters[leafCount] = boxCenter( node.box );
			leafCount += 1;
			node.parent = kNull;
		}
		else
		{
			int doomed = nodeIndex;
			nodeIndex = node.child1;
			if ( sp < kTreeStack )
				stack[sp++] = node.child2;
			else
				setError( w, kErrTreeStack, __LINE__ );
			treeFreeNode( w, t, doomed );
			continue;
		}
		if ( sp == 0 )
			break;
		nodeIndex = stack[--sp];
	}
	t.root = treeBuild( w, t, leafCount );
	return leafCount;
}

// dynamic_tree.c:1040-1075 b2DynamicTree_GetHeight
F2D_HD int treeHeight( const World* w, const Tree& t )
{
	if ( t.root == kNull )
		return 0;
	return ptr( w, t.nodes )[t.root].height;
}

// Overlap query with the reference's visit order (child2 subtree first; dynamic_tree.c:1114-1170).
// `visit(proxyId, userData)` returns false to stop.
// `visit( leaf id, userData, node flags )`
template <class F> F2D_HDF inline void treeQueryFlags( World* w, const Tree& t, Box box, uint64_t maskBits, F&& visit )
{
	if ( t.nodeCount == 0 || t.root == kNull )
		return;
	const TreeNode* nodes = ptr( w, t.nodes );
	// The reference pushes both children of an overlapping node and tests a node when it is popped: one dependent load
	// per visited node, and about half the visited nodes fail the test. Here the two children of the current node are
	// loaded together and tested at once; only nodes that passed are descended into or parked on the stack. Leaves are
	// reported in the same order (child2's subtree before child1's: the reference pops child2 first).
	struct Fields
	{
		int32_t child1, child2; // leaf: the low half of userData is the shape id
		uint32_t flags;
	};
	auto fieldsOf = []( const TreeNode& n ) {
		const Q4 q = load16( &n.child1 ); // child1, child2, parent, height | flags << 16
		return Fields{ (int32_t)floatBits( q.x ), (int32_t)floatBits( q.y ), floatBits( q.w ) >> 16 };
	};
	// (bitwise, not short-circuit: with || and && the compiler loads the box in pieces, one load - and one round trip -
	// per branch of the test; here the whole test is one round of loads and no branch)
	auto passes = [&]( const TreeNode& n ) {
		const Q4 b = load16( &n.box );
		const uint64_t category = n.category;
		const bool apart = ( box.lo.x > b.z ) | ( box.lo.y > b.w ) | ( b.x > box.hi.x ) | ( b.y > box.hi.y );
		return ( apart == false ) & ( ( category & maskBits ) != 0 );
	};
	int32_t stack[kQueryStack];
	int sp = 0;
	int id = t.root;
	if ( passes( nodes[id] ) == false )
		return;
	Fields cur = fieldsOf( nodes[id] );
	while ( true )
	{
		if ( cur.flags & kNodeLeaf )
		{
			if ( visit( id, (uint64_t)(uint32_t)cur.child1 | ( (uint64_t)(uint32_t)cur.child2 << 32 ), cur.flags ) == false )
				return;
		}
		else
		{
			const TreeNode& first = nodes[cur.child2];
			const TreeNode& second = nodes[cur.child1];
			const bool firstPasses = passes( first ), secondPasses = passes( second );
			const Fields firstFields = fieldsOf( first ), secondFields = fieldsOf( second );
			if ( firstPasses )
			{
				if ( secondPasses )
				{
					if ( sp < kQueryStack )
						stack[sp++] = cur.child1;
					else
						setError( w, kErrTreeStack, __LINE__ );
				}
				id = cur.child2;
				cur = firstFields;
				continue;
			}
			if ( secondPasses )
			{
				id = cur.child1;
				cur = secondFields;
				continue;
			}
		}
		if ( sp == 0 )
			return;
		id = stack[--sp];
		cur = fieldsOf( nodes[id] );
	}
}

// dynamic_tree.c:1114-1170 `visit( leaf id, userData )`
template <class F> F2D_HDF inline void treeQuery( World* w, const Tree& t, Box box, uint64_t maskBits, F&& visit )
{
	treeQueryFlags( w, t, box, maskBits, [&]( int id, uint64_t userData, uint32_t ) -> bool { return visit( id, userData ); } );
}

} // namespace f2d
