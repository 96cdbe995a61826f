// forge2d_b200 — entity bookkeeping on the world image: bodies, shapes, proxies, contacts, the constraint graph,
// islands and sleeping sets. Runs unchanged on the host (API-time mutations) and on the device (in-step serial
// sections executed by one thread of the world's team).
//
// These are the order-defining parts of the reference: every append / swap-remove / list insert
// happens at the same point, on the same list position, as in B2/src/{body,shape,contact,constraint_graph,
// island,solver_set,broad_phase}.c — but on int32 id lists over stable record slots (see f2d_types.h).
#pragma once
#include "f2d_manifold.h"
#include "f2d_tree.h"

namespace f2d
{

// ------------------------------------------------------------------------------------------------ geometry
// geometry.c:401-454
F2D_HDF inline Box shapeAABB( const Shape& s, Xf xf )
{
	switch ( s.type )
	{
		case kCapsule:
		{
			V2 v1 = xfPoint( xf, s.capsule.c1 );
			V2 v2 = xfPoint( xf, s.capsule.c2 );
			V2 r = { s.capsule.radius, s.capsule.radius };
			return Box{ sub( vmin( v1, v2 ), r ), add( vmax( v1, v2 ), r ) };
		}
		case kCircle:
		{
			V2 p = xfPoint( xf, s.circle.center );
			float r = s.circle.radius;
			return Box{ { p.x - r, p.y - r }, { p.x + r, p.y + r } };
		}
		case kPolygon:
		{
			V2 lo = xfPoint( xf, s.polygon.v[0] );
			V2 hi = lo;
			for ( int i = 1; i < s.polygon.count; ++i )
			{
				V2 v = xfPoint( xf, s.polygon.v[i] );
				lo = vmin( lo, v );
				hi = vmax( hi, v );
			}
			V2 r = { s.polygon.radius, s.polygon.radius };
			return Box{ sub( lo, r ), add( hi, r ) };
		}
		case kSegment:
		{
			V2 v1 = xfPoint( xf, s.segment.p1 );
			V2 v2 = xfPoint( xf, s.segment.p2 );
			return Box{ vmin( v1, v2 ), vmax( v1, v2 ) };
		}
		default:
		{
			V2 v1 = xfPoint( xf, s.chainSegment.segment.p1 );
			V2 v2 = xfPoint( xf, s.chainSegment.segment.p2 );
			return Box{ vmin( v1, v2 ), vmax( v1, v2 ) };
		}
	}
}

F2D_HD Box inflate( Box b, float m )
{
	Box o;
	o.lo.x = b.lo.x - m;
	o.lo.y = b.lo.y - m;
	o.hi.x = b.hi.x + m;
	o.hi.y = b.hi.y + m;
	return o;
}

// shape.c:~600 b2GetShapeCentroid
F2D_HD V2 shapeCentroid( const Shape& s )
{
	switch ( s.type )
	{
		case kCapsule:
			return lerp( s.capsule.c1, s.capsule.c2, 0.5f );
		case kCircle:
			return s.circle.center;
		case kPolygon:
			return s.polygon.centroid;
		case kSegment:
			return lerp( s.segment.p1, s.segment.p2, 0.5f );
		case kChainSegment:
			return lerp( s.chainSegment.segment.p1, s.chainSegment.segment.p2, 0.5f );
		default:
			return V2{ 0.0f, 0.0f };
	}
}

struct MassData
{
	float mass;
	V2 center;
	float inertia;
};
// geometry.c:220-240
F2D_HD MassData circleMass( const Circle& c, float density )
{
	float rr = c.radius * c.radius;
	MassData m;
	m.mass = density * kPi * rr;
	m.center = c.center;
	m.inertia = m.mass * ( 0.5f * rr + dot( c.center, c.center ) );
	return m;
}
// geometry.c:242-279
F2D_HD MassData capsuleMass( const Capsule& c, float density )
{
	float radius = c.radius;
	float rr = radius * radius;
	V2 p1 = c.c1, p2 = c.c2;
	float len = length( sub( p2, p1 ) );
	float ll = len * len;
	float circleMassV = density * ( kPi * radius * radius );
	float boxMass = density * ( 2.0f * radius * len );
	MassData m;
	m.mass = circleMassV + boxMass;
	m.center.x = 0.5f * ( p1.x + p2.x );
	m.center.y = 0.5f * ( p1.y + p2.y );
	float lc = 4.0f * radius / ( 3.0f * kPi );
	float h = 0.5f * len;
	float circleInertia = circleMassV * ( 0.5f * rr + h * h + 2.0f * h * lc );
	float boxInertia = boxMass * ( 4.0f * rr + ll ) / 12.0f;
	m.inertia = circleInertia + boxInertia;
	m.inertia += m.mass * dot( m.center, m.center );
	return m;
}
// geometry.c:281-399
F2D_HDF inline MassData polygonMass( const Poly& p, float density )
{
	if ( p.count == 1 )
	{
		Circle c = { p.v[0], p.radius };
		return circleMass( c, density );
	}
	if ( p.count == 2 )
	{
		Capsule c = { p.v[0], p.v[1], p.radius };
		return capsuleMass( c, density );
	}
	V2 vertices[kMaxPolyVerts];
	int count = p.count;
	float radius = p.radius;
	if ( radius > 0.0f )
	{
		float sqrt2 = 1.412f;
		for ( int i = 0; i < count; ++i )
		{
			int j = i == 0 ? count - 1 : i - 1;
			V2 mid = normalize( add( p.n[j], p.n[i] ) );
			vertices[i] = mulAdd( p.v[i], sqrt2 * radius, mid );
		}
	}
	else
	{
		for ( int i = 0; i < count; ++i )
			vertices[i] = p.v[i];
	}
	V2 center = { 0.0f, 0.0f };
	float area = 0.0f;
	float inertia = 0.0f;
	V2 r = vertices[0];
	const float inv3 = 1.0f / 3.0f;
	for ( int i = 1; i < count - 1; ++i )
	{
		V2 e1 = sub( vertices[i], r );
		V2 e2 = sub( vertices[i + 1], r );
		float D = cross( e1, e2 );
		float triangleArea = 0.5f * D;
		area += triangleArea;
		center = mulAdd( center, triangleArea * inv3, add( e1, e2 ) );
		float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
		float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
		float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
		inertia += ( 0.25f * inv3 * D ) * ( intx2 + inty2 );
	}
	MassData m;
	m.mass = density * area;
	float invArea = 1.0f / area;
	center.x *= invArea;
	center.y *= invArea;
	m.center = add( r, center );
	m.inertia = density * inertia;
	m.inertia += m.mass * ( dot( m.center, m.center ) - dot( center, center ) );
	return m;
}
// shape.c:707-720
F2D_HD MassData shapeMass( const Shape& s )
{
	switch ( s.type )
	{
		case kCapsule:
			return capsuleMass( s.capsule, s.density );
		case kCircle:
			return circleMass( s.circle, s.density );
		case kPolygon:
			return polygonMass( s.polygon, s.density );
		default:
			return MassData{ 0.0f, { 0.0f, 0.0f }, 0.0f };
	}
}
// shape.c:722-790
F2D_HDF inline void shapeExtent( const Shape& s, V2 localCenter, float* minExtent, float* maxExtent )
{
	*minExtent = 0.0f;
	*maxExtent = 0.0f;
	switch ( s.type )
	{
		case kCapsule:
		{
			float radius = s.capsule.radius;
			*minExtent = radius;
			V2 c1 = sub( s.capsule.c1, localCenter );
			V2 c2 = sub( s.capsule.c2, localCenter );
			*maxExtent = sqrtf( maxf( lengthSq( c1 ), lengthSq( c2 ) ) ) + radius;
		}
		break;
		case kCircle:
		{
			float radius = s.circle.radius;
			*minExtent = radius;
			*maxExtent = length( sub( s.circle.center, localCenter ) ) + radius;
		}
		break;
		case kPolygon:
		{
			const Poly& poly = s.polygon;
			float mn = kHuge;
			float maxSqr = 0.0f;
			for ( int i = 0; i < poly.count; ++i )
			{
				V2 v = poly.v[i];
				float planeOffset = dot( poly.n[i], sub( v, poly.centroid ) );
				mn = minf( mn, planeOffset );
				float dsq = lengthSq( sub( v, localCenter ) );
				maxSqr = maxf( maxSqr, dsq );
			}
			*minExtent = mn + poly.radius;
			*maxExtent = sqrtf( maxSqr ) + poly.radius;
		}
		break;
		case kSegment:
		{
			*minExtent = 0.0f;
			V2 c1 = sub( s.segment.p1, localCenter );
			V2 c2 = sub( s.segment.p2, localCenter );
			*maxExtent = sqrtf( maxf( lengthSq( c1 ), lengthSq( c2 ) ) );
		}
		break;
		case kChainSegment:
		{
			*minExtent = 0.0f;
			V2 c1 = sub( s.chainSegment.segment.p1, localCenter );
			V2 c2 = sub( s.chainSegment.segment.p2, localCenter );
			*maxExtent = sqrtf( maxf( lengthSq( c1 ), lengthSq( c2 ) ) );
		}
		break;
		default:
			break;
	}
}
F2D_HD float shapeRadius( const Shape& s ) // shape.h:107-120
{
	switch ( s.type )
	{
		case kCapsule:
			return s.capsule.radius;
		case kCircle:
			return s.circle.radius;
		case kPolygon:
			return s.polygon.radius;
		default:
			return 0.0f;
	}
}
F2D_HD bool shouldShapesCollide( const Filter& a, const Filter& b ) // shape.h:122-130
{
	if ( a.group == b.group && a.group != 0 )
		return a.group > 0;
	return ( a.mask & b.category ) != 0 && ( a.category & b.mask ) != 0;
}

// ------------------------------------------------------------------------------------------------ broadphase
// broad_phase.h:74-82 b2BufferMove — the moveSet hash is a flag on the proxy's tree node
F2D_HD void bufferMove( World* w, int key )
{
	TreeNode& n = ptr( w, w->trees[proxyType( key )].nodes )[proxyId( key )];
	if ( ( n.flags & kNodeMoved ) == 0 )
	{
		n.flags |= kNodeMoved;
		F2D_PUSH( w, w->moveArray, key );
	}
}
// broad_phase.c:71-89
F2D_HDF inline void unbufferMove( World* w, int key )
{
	TreeNode& n = ptr( w, w->trees[proxyType( key )].nodes )[proxyId( key )];
	if ( n.flags & kNodeMoved )
	{
		n.flags &= (uint16_t)~kNodeMoved;
		int32_t* mv = ptr( w, w->moveArray );
		for ( int i = 0; i < w->moveArray.count; ++i )
		{
			if ( mv[i] == key )
			{
				removeSwap( w, w->moveArray, i );
				break;
			}
		}
	}
}
// broad_phase.c:91-102
F2D_HD int bpCreateProxy( World* w, int type, Box box, uint64_t category, int shapeIndex, bool forcePairCreation )
{
	int id = treeCreateProxy( w, w->trees[type], box, category, (uint64_t)(uint32_t)shapeIndex );
	int key = proxyKey( id, type );
	if ( type != kStaticBody || forcePairCreation )
		bufferMove( w, key );
	return key;
}
// broad_phase.c:104-114
F2D_HD void bpDestroyProxy( World* w, int key )
{
	unbufferMove( w, key );
	treeDestroyProxy( w, w->trees[proxyType( key )], proxyId( key ) );
}
// broad_phase.c:116-123
F2D_HD void bpMoveProxy( World* w, int key, Box box )
{
	treeMoveProxy( w, w->trees[proxyType( key )], proxyId( key ), box );
	bufferMove( w, key );
}

// ------------------------------------------------------------------------------------------------ islands
// island.c:20-58
F2D_HDF inline int createIsland( World* w, int setIndex )
{
	int islandId = allocId( w, w->islandIds );
	if ( islandId == w->islands.count )
	{
		Island empty;
		memset( &empty, 0, sizeof( empty ) );
		F2D_PUSH( w, w->islands, empty );
	}
	Island& is = ptr( w, w->islands )[islandId];
	is.setIndex = setIndex;
	is.islandId = islandId;
	is.headBody = is.tailBody = kNull;
	is.bodyCount = 0;
	is.headContact = is.tailContact = kNull;
	is.contactCount = 0;
	is.headJoint = is.tailJoint = kNull;
	is.jointCount = 0;
	is.parentIsland = kNull;
	is.constraintRemoveCount = 0;
	if ( setIndex == kAwakeSet )
	{
		is.localIndex = w->awakeIslands.count;
		F2D_PUSH( w, w->awakeIslands, islandId );
	}
	else
	{
		// sleeping set created asleep (body created with isAwake=false): single island slot
		SolverSet& set = ptr( w, w->sets )[setIndex];
		is.localIndex = set.islandCount;
		int32_t* pool = ptr( w, w->sleepPool );
		if ( set.islandCount < set.islandCap )
			pool[set.blockOff + set.bodyCap + set.contactCap + set.jointCap + set.islandCount] = islandId;
		else
			setError( w, kErrSleepPool, __LINE__ );
		set.islandCount += 1;
	}
	return islandId;
}

F2D_HD int32_t* setIslandList( World* w, SolverSet& set )
{
	return ptr( w, w->sleepPool ) + set.blockOff + set.bodyCap + set.contactCap + set.jointCap;
}
F2D_HD int32_t* setBodyList( World* w, SolverSet& set ) { return ptr( w, w->sleepPool ) + set.blockOff; }
F2D_HD int32_t* setContactList( World* w, SolverSet& set ) { return ptr( w, w->sleepPool ) + set.blockOff + set.bodyCap; }
F2D_HD int32_t* setJointList( World* w, SolverSet& set )
{
	return ptr( w, w->sleepPool ) + set.blockOff + set.bodyCap + set.contactCap;
}

// island.c:60-90
F2D_HDF inline void destroyIsland( World* w, int islandId )
{
	if ( w->splitIslandId == islandId )
		w->splitIslandId = kNull;
	Island* islands = ptr( w, w->islands );
	Island& is = islands[islandId];
	if ( is.setIndex == kAwakeSet )
	{
		int moved = removeSwap( w, w->awakeIslands, is.localIndex );
		if ( moved != kNull )
		{
			int movedId = ptr( w, w->awakeIslands )[is.localIndex];
			islands[movedId].localIndex = is.localIndex;
		}
	}
	else
	{
		SolverSet& set = ptr( w, w->sets )[is.setIndex];
		int32_t* list = setIslandList( w, set );
		int last = set.islandCount - 1;
		if ( is.localIndex != last )
		{
			list[is.localIndex] = list[last];
			islands[list[is.localIndex]].localIndex = is.localIndex;
		}
		set.islandCount -= 1;
	}
	is.islandId = kNull;
	is.setIndex = kNull;
	is.localIndex = kNull;
	freeId( w, w->islandIds, islandId );
}

// island.c:92-113
// (`contactId` is passed explicitly: the slot index, not Contact::contactId, which the graph half of the contact state
// pass may be clearing at the same moment - f2d_step.h contactStateGraphHalf)
F2D_HD void addContactToIsland( World* w, int islandId, Contact& c, int contactId )
{
	Island& is = ptr( w, w->islands )[islandId];
	if ( is.headContact != kNull )
	{
		c.islandNext = is.headContact;
		ptr( w, w->contacts )[is.headContact].islandPrev = contactId;
	}
	is.headContact = contactId;
	if ( is.tailContact == kNull )
		is.tailContact = is.headContact;
	is.contactCount += 1;
	c.islandId = islandId;
}

F2D_HDF inline void wakeSolverSet( World* w, int setIndex );

// Union-find root with the reference's path-halving writes (island.c:141-176)
F2D_HD int islandRoot( Island* islands, int islandId )
{
	Island* is = islands + islandId;
	int parentId = is->parentIsland;
	while ( parentId != kNull )
	{
		Island* parent = islands + parentId;
		if ( parent->parentIsland != kNull )
			is->parentIsland = parent->parentIsland;
		is = parent;
		islandId = parentId;
		parentId = is->parentIsland;
	}
	return islandId;
}

// island.c:116-218
F2D_HDF inline void linkContact( World* w, Contact& c, int contactId )
{
	Body* bodies = ptr( w, w->bodies );
	int idA = c.edges[0].bodyId, idB = c.edges[1].bodyId;
	if ( bodies[idA].setIndex == kAwakeSet && bodies[idB].setIndex >= kFirstSleepingSet )
		wakeSolverSet( w, bodies[idB].setIndex );
	if ( bodies[idB].setIndex == kAwakeSet && bodies[idA].setIndex >= kFirstSleepingSet )
		wakeSolverSet( w, bodies[idA].setIndex );

	int islandIdA = bodies[idA].islandId;
	int islandIdB = bodies[idB].islandId;
	if ( islandIdA == islandIdB )
	{
		addContactToIsland( w, islandIdA, c, contactId );
		return;
	}
	Island* islands = ptr( w, w->islands );
	if ( islandIdA != kNull )
		islandIdA = islandRoot( islands, islandIdA );
	if ( islandIdB != kNull )
		islandIdB = islandRoot( islands, islandIdB );
	if ( islandIdA != islandIdB && islandIdA != kNull && islandIdB != kNull )
		islands[islandIdB].parentIsland = islandIdA;
	if ( islandIdA != kNull )
		addContactToIsland( w, islandIdA, c, contactId );
	else
		addContactToIsland( w, islandIdB, c, contactId );
}

// island.c:221-262
F2D_HDF inline void unlinkContact( World* w, Contact& c, int contactId )
{
	Island& is = ptr( w, w->islands )[c.islandId];
	Contact* contacts = ptr( w, w->contacts );
	if ( c.islandPrev != kNull )
		contacts[c.islandPrev].islandNext = c.islandNext;
	if ( c.islandNext != kNull )
		contacts[c.islandNext].islandPrev = c.islandPrev;
	if ( is.headContact == contactId )
		is.headContact = c.islandNext;
	if ( is.tailContact == contactId )
		is.tailContact = c.islandPrev;
	is.contactCount -= 1;
	is.constraintRemoveCount += 1;
	c.islandId = kNull;
	c.islandPrev = kNull;
	c.islandNext = kNull;
}

// island.c:264-289
F2D_HD void addJointToIsland( World* w, int islandId, Joint& j )
{
	Island& is = ptr( w, w->islands )[islandId];
	if ( is.headJoint != kNull )
	{
		j.islandNext = is.headJoint;
		ptr( w, w->joints )[is.headJoint].islandPrev = j.jointId;
	}
	is.headJoint = j.jointId;
	if ( is.tailJoint == kNull )
		is.tailJoint = is.headJoint;
	is.jointCount += 1;
	j.islandId = islandId;
}

F2D_HDF inline void mergeAwakeIslands( World* w );

// island.c:291-382
F2D_HDF inline void linkJoint( World* w, Joint& j, bool mergeIslands )
{
	Body* bodies = ptr( w, w->bodies );
	int idA = j.edges[0].bodyId, idB = j.edges[1].bodyId;
	if ( bodies[idA].setIndex == kAwakeSet && bodies[idB].setIndex >= kFirstSleepingSet )
		wakeSolverSet( w, bodies[idB].setIndex );
	else if ( bodies[idB].setIndex == kAwakeSet && bodies[idA].setIndex >= kFirstSleepingSet )
		wakeSolverSet( w, bodies[idA].setIndex );

	int islandIdA = bodies[idA].islandId;
	int islandIdB = bodies[idB].islandId;
	if ( islandIdA == islandIdB )
	{
		addJointToIsland( w, islandIdA, j );
		return;
	}
	Island* islands = ptr( w, w->islands );
	if ( islandIdA != kNull )
		islandIdA = islandRoot( islands, islandIdA );
	if ( islandIdB != kNull )
		islandIdB = islandRoot( islands, islandIdB );
	if ( islandIdA != islandIdB && islandIdA != kNull && islandIdB != kNull )
		islands[islandIdB].parentIsland = islandIdA;
	if ( islandIdA != kNull )
		addJointToIsland( w, islandIdA, j );
	else
		addJointToIsland( w, islandIdB, j );
	if ( mergeIslands )
		mergeAwakeIslands( w );
}

// island.c:384-425
F2D_HDF inline void unlinkJoint( World* w, Joint& j )
{
	Island& is = ptr( w, w->islands )[j.islandId];
	Joint* joints = ptr( w, w->joints );
	if ( j.islandPrev != kNull )
		joints[j.islandPrev].islandNext = j.islandNext;
	if ( j.islandNext != kNull )
		joints[j.islandNext].islandPrev = j.islandPrev;
	if ( is.headJoint == j.jointId )
		is.headJoint = j.islandNext;
	if ( is.tailJoint == j.jointId )
		is.tailJoint = j.islandPrev;
	is.jointCount -= 1;
	is.constraintRemoveCount += 1;
	j.islandId = kNull;
	j.islandPrev = kNull;
	j.islandNext = kNull;
}

// island.c:427-537 b2MergeIsland: child lists are appended to the root's (root-then-child order)
F2D_HDF inline void relabelIsland( World* w, const Island& island, int rootId )
{
	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	for ( int id = island.headBody; id != kNull; id = bodies[id].islandNext )
		bodies[id].islandId = rootId;
	for ( int id = island.headContact; id != kNull; id = contacts[id].islandNext )
		contacts[id].islandId = rootId;
	for ( int id = island.headJoint; id != kNull; id = joints[id].islandNext )
		joints[id].islandId = rootId;
}

F2D_HDF inline void mergeIsland( World* w, Island& island, bool relabel = true )
{
	int rootId = island.parentIsland;
	Island& root = ptr( w, w->islands )[rootId];
	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	if ( relabel )
		relabelIsland( w, island, rootId );

	bodies[root.tailBody].islandNext = island.headBody;
	bodies[island.headBody].islandPrev = root.tailBody;
	root.tailBody = island.tailBody;
	root.bodyCount += island.bodyCount;

	if ( root.headContact == kNull )
	{
		root.headContact = island.headContact;
		root.tailContact = island.tailContact;
		root.contactCount = island.contactCount;
	}
	else if ( island.headContact != kNull )
	{
		contacts[root.tailContact].islandNext = island.headContact;
		contacts[island.headContact].islandPrev = root.tailContact;
		root.tailContact = island.tailContact;
		root.contactCount += island.contactCount;
	}
	if ( root.headJoint == kNull )
	{
		root.headJoint = island.headJoint;
		root.tailJoint = island.tailJoint;
		root.jointCount = island.jointCount;
	}
	else if ( island.headJoint != kNull )
	{
		joints[root.tailJoint].islandNext = island.headJoint;
		joints[island.headJoint].islandPrev = root.tailJoint;
		root.tailJoint = island.tailJoint;
		root.jointCount += island.jointCount;
	}
	root.constraintRemoveCount += island.constraintRemoveCount;
}

// island.c:539-598
F2D_HDF inline void mergeAwakeIslands( World* w )
{
	Island* islands = ptr( w, w->islands );
	int32_t* awake = ptr( w, w->awakeIslands );
	int count = w->awakeIslands.count;
	for ( int i = 0; i < count; ++i )
	{
		int islandId = awake[i];
		Island* island = islands + islandId;
		int rootId = islandId;
		Island* root = island;
		while ( root->parentIsland != kNull )
		{
			Island* parent = islands + root->parentIsland;
			if ( parent->parentIsland != kNull )
				root->parentIsland = parent->parentIsland;
			rootId = root->parentIsland;
			root = parent;
		}
		if ( root != island )
			island->parentIsland = rootId;
	}
	for ( int i = count - 1; i >= 0; --i )
	{
		int islandId = awake[i];
		Island& island = islands[islandId];
		if ( island.parentIsland == kNull )
			continue;
		mergeIsland( w, island );
		destroyIsland( w, islandId );
	}
}

// Depth-first split of one island into connected components: island.c:602-840.
// Scratch: w->scratch holds the DFS stack and the seed list (2 x bodyCount ints).
F2D_HDF inline void splitIsland( World* w, int baseId )
{
	Island* islands = ptr( w, w->islands );
	Island& base = islands[baseId];
	int setIndex = base.setIndex;
	if ( setIndex != kAwakeSet )
		return;
	if ( base.constraintRemoveCount == 0 )
		return;
	int bodyCount = base.bodyCount;
	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	if ( 2 * bodyCount > w->scratch.cap )
	{
		setError( w, kErrCapacity, __LINE__ );
		return;
	}
	int32_t* stack = ptr( w, w->scratch );
	int32_t* bodyIds = stack + bodyCount;

	int index = 0;
	for ( int id = base.headBody; id != kNull; id = bodies[id].islandNext )
	{
		bodyIds[index++] = id;
		bodies[id].isMarked = false;
	}
	for ( int id = base.headContact; id != kNull; id = contacts[id].islandNext )
		contacts[id].flags &= ~kContactMarked;
	for ( int id = base.headJoint; id != kNull; id = joints[id].islandNext )
		joints[id].isMarked = false;

	destroyIsland( w, baseId );

	for ( int i = 0; i < bodyCount; ++i )
	{
		int seedIndex = bodyIds[i];
		Body& seed = bodies[seedIndex];
		if ( seed.isMarked )
			continue;
		int sp = 0;
		stack[sp++] = seedIndex;
		seed.isMarked = true;

		int islandId = createIsland( w, setIndex );
		islands = ptr( w, w->islands );
		Island& island = islands[islandId];

		while ( sp > 0 )
		{
			int bodyId = stack[--sp];
			Body& body = bodies[bodyId];
			body.islandId = islandId;
			if ( island.tailBody != kNull )
				bodies[island.tailBody].islandNext = bodyId;
			body.islandPrev = island.tailBody;
			body.islandNext = kNull;
			island.tailBody = bodyId;
			if ( island.headBody == kNull )
				island.headBody = bodyId;
			island.bodyCount += 1;

			int contactKey = body.headContactKey;
			while ( contactKey != kNull )
			{
				int contactId = contactKey >> 1;
				int edgeIndex = contactKey & 1;
				Contact& contact = contacts[contactId];
				contactKey = contact.edges[edgeIndex].nextKey;
				if ( contact.flags & kContactMarked )
					continue;
				if ( ( contact.flags & kContactTouching ) == 0 )
					continue;
				contact.flags |= kContactMarked;
				int otherBodyId = contact.edges[edgeIndex ^ 1].bodyId;
				Body& other = bodies[otherBodyId];
				if ( other.isMarked == false && other.setIndex != kStaticSet )
				{
					stack[sp++] = otherBodyId;
					other.isMarked = true;
				}
				contact.islandId = islandId;
				if ( island.tailContact != kNull )
					contacts[island.tailContact].islandNext = contactId;
				contact.islandPrev = island.tailContact;
				contact.islandNext = kNull;
				island.tailContact = contactId;
				if ( island.headContact == kNull )
					island.headContact = contactId;
				island.contactCount += 1;
			}

			int jointKey = body.headJointKey;
			while ( jointKey != kNull )
			{
				int jointId = jointKey >> 1;
				int edgeIndex = jointKey & 1;
				Joint& joint = joints[jointId];
				jointKey = joint.edges[edgeIndex].nextKey;
				if ( joint.isMarked )
					continue;
				joint.isMarked = true;
				int otherBodyId = joint.edges[edgeIndex ^ 1].bodyId;
				Body& other = bodies[otherBodyId];
				if ( other.setIndex == kDisabledSet )
					continue;
				if ( other.isMarked == false && other.setIndex == kAwakeSet )
				{
					stack[sp++] = otherBodyId;
					other.isMarked = true;
				}
				joint.islandId = islandId;
				if ( island.tailJoint != kNull )
					joints[island.tailJoint].islandNext = jointId;
				joint.islandPrev = island.tailJoint;
				joint.islandNext = kNull;
				island.tailJoint = jointId;
				if ( island.headJoint == kNull )
					island.headJoint = jointId;
				island.jointCount += 1;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ constraint graph
// First-fit colour over per-body colour masks: constraint_graph.c:68-134, 214-267.
// forContact: contacts with one static body never use colour 0.
F2D_HD int assignColor( Body& bodyA, Body& bodyB, bool staticA, bool staticB, bool forContact )
{
	if ( staticA == false && staticB == false )
	{
		uint32_t used = (uint32_t)bodyA.colorMask | (uint32_t)bodyB.colorMask;
		for ( int i = 0; i < kOverflow; ++i )
		{
			if ( used & ( 1u << i ) )
				continue;
			bodyA.colorMask |= (uint16_t)( 1u << i );
			bodyB.colorMask |= (uint16_t)( 1u << i );
			return i;
		}
	}
	else if ( staticA == false )
	{
		for ( int i = forContact ? 1 : 0; i < kOverflow; ++i )
		{
			if ( bodyA.colorMask & ( 1u << i ) )
				continue;
			bodyA.colorMask |= (uint16_t)( 1u << i );
			return i;
		}
	}
	else if ( staticB == false )
	{
		for ( int i = forContact ? 1 : 0; i < kOverflow; ++i )
		{
			if ( bodyB.colorMask & ( 1u << i ) )
				continue;
			bodyB.colorMask |= (uint16_t)( 1u << i );
			return i;
		}
	}
	return kOverflow;
}

// constraint_graph.c:68-182
F2D_HDF inline void addContactToGraph( World* w, int contactId )
{
	Contact& c = ptr( w, w->contacts )[contactId];
	ContactSim& sim = ptr( w, w->contactSims )[contactId];
	Body* bodies = ptr( w, w->bodies );
	Body& bodyA = bodies[c.edges[0].bodyId];
	Body& bodyB = bodies[c.edges[1].bodyId];
	bool staticA = bodyA.setIndex == kStaticSet;
	bool staticB = bodyB.setIndex == kStaticSet;
	int colorIndex = assignColor( bodyA, bodyB, staticA, staticB, true );
	c.colorIndex = colorIndex;
	c.localIndex = w->colorContacts[colorIndex].count;
	F2D_PUSH( w, w->colorContacts[colorIndex], contactId );

	const BodySim* sims = ptr( w, w->sims );
	if ( staticA )
	{
		sim.bodySimIndexA = kNull;
		sim.invMassA = 0.0f;
		sim.invIA = 0.0f;
	}
	else
	{
		sim.bodySimIndexA = bodyA.localIndex;
		sim.invMassA = sims[bodyA.id].invMass;
		sim.invIA = sims[bodyA.id].invInertia;
	}
	if ( staticB )
	{
		sim.bodySimIndexB = kNull;
		sim.invMassB = 0.0f;
		sim.invIB = 0.0f;
	}
	else
	{
		sim.bodySimIndexB = bodyB.localIndex;
		sim.invMassB = sims[bodyB.id].invMass;
		sim.invIB = sims[bodyB.id].invInertia;
	}
}

// constraint_graph.c:184-212
F2D_HDF inline void removeContactFromGraph( World* w, int bodyIdA, int bodyIdB, int colorIndex, int localIndex )
{
	if ( colorIndex != kOverflow )
	{
		Body* bodies = ptr( w, w->bodies );
		bodies[bodyIdA].colorMask &= (uint16_t)~( 1u << colorIndex );
		bodies[bodyIdB].colorMask &= (uint16_t)~( 1u << colorIndex );
	}
	int moved = removeSwap( w, w->colorContacts[colorIndex], localIndex );
	if ( moved != kNull )
	{
		int movedId = ptr( w, w->colorContacts[colorIndex] )[localIndex];
		ptr( w, w->contacts )[movedId].localIndex = localIndex;
	}
}

// constraint_graph.c:269-294
F2D_HDF inline void addJointToGraph( World* w, int jointId )
{
	Joint& j = ptr( w, w->joints )[jointId];
	Body* bodies = ptr( w, w->bodies );
	Body& bodyA = bodies[j.edges[0].bodyId];
	Body& bodyB = bodies[j.edges[1].bodyId];
	bool staticA = bodyA.setIndex == kStaticSet;
	bool staticB = bodyB.setIndex == kStaticSet;
	int colorIndex = assignColor( bodyA, bodyB, staticA, staticB, false );
	j.colorIndex = colorIndex;
	j.localIndex = w->colorJoints[colorIndex].count;
	F2D_PUSH( w, w->colorJoints[colorIndex], jointId );
}

// constraint_graph.c:296-322
F2D_HDF inline void removeJointFromGraph( World* w, int bodyIdA, int bodyIdB, int colorIndex, int localIndex )
{
	if ( colorIndex != kOverflow )
	{
		Body* bodies = ptr( w, w->bodies );
		bodies[bodyIdA].colorMask &= (uint16_t)~( 1u << colorIndex );
		bodies[bodyIdB].colorMask &= (uint16_t)~( 1u << colorIndex );
	}
	int moved = removeSwap( w, w->colorJoints[colorIndex], localIndex );
	if ( moved != kNull )
	{
		int movedId = ptr( w, w->colorJoints[colorIndex] )[localIndex];
		ptr( w, w->joints )[movedId].localIndex = localIndex;
	}
}

// ------------------------------------------------------------------------------------------------ id lists per set
// Removes contact `c` from the id list of the non-graph set that owns it (awake non-touching / disabled / sleeping).
F2D_HDF inline void removeContactFromSetList( World* w, Contact& c )
{
	Contact* contacts = ptr( w, w->contacts );
	if ( c.setIndex == kAwakeSet )
	{
		int moved = removeSwap( w, w->awakeContacts, c.localIndex );
		if ( moved != kNull )
			contacts[ptr( w, w->awakeContacts )[c.localIndex]].localIndex = c.localIndex;
	}
	else if ( c.setIndex == kDisabledSet )
	{
		int moved = removeSwap( w, w->disabledContacts, c.localIndex );
		if ( moved != kNull )
			contacts[ptr( w, w->disabledContacts )[c.localIndex]].localIndex = c.localIndex;
	}
	else
	{
		SolverSet& set = ptr( w, w->sets )[c.setIndex];
		int32_t* list = setContactList( w, set );
		int last = set.contactCount - 1;
		if ( c.localIndex != last )
		{
			list[c.localIndex] = list[last];
			contacts[list[c.localIndex]].localIndex = c.localIndex;
		}
		set.contactCount -= 1;
	}
}

// ------------------------------------------------------------------------------------------------ contacts
// contact.c:186-328. Returns the contact id or kNull when the type pair has no manifold function.
F2D_HDF inline int createContact( World* w, int shapeIdA, int shapeIdB )
{
	Shape* shapes = ptr( w, w->shapes );
	if ( pairHasManifold( shapes[shapeIdA].type, shapes[shapeIdB].type ) == false )
		return kNull;
	if ( pairIsPrimary( shapes[shapeIdA].type, shapes[shapeIdB].type ) == false )
	{
		int t = shapeIdA;
		shapeIdA = shapeIdB;
		shapeIdB = t;
	}
	Shape& shapeA = shapes[shapeIdA];
	Shape& shapeB = shapes[shapeIdB];
	Body* bodies = ptr( w, w->bodies );
	Body& bodyA = bodies[shapeA.bodyId];
	Body& bodyB = bodies[shapeB.bodyId];

	int setIndex = ( bodyA.setIndex == kAwakeSet || bodyB.setIndex == kAwakeSet ) ? kAwakeSet : kDisabledSet;

	int contactId = allocId( w, w->contactIds );
	if ( contactId == w->contacts.count )
	{
		if ( w->contacts.count >= w->contacts.cap || w->contactSims.count >= w->contactSims.cap )
		{
			setError( w, kErrCapacity, __LINE__ );
			w->contactIds.next -= 1;
			return kNull;
		}
		w->contacts.count += 1;
		w->contactSims.count += 1;
		memset( ptr( w, w->contacts ) + contactId, 0, sizeof( Contact ) );
	}
	Contact* contacts = ptr( w, w->contacts );
	Contact& c = contacts[contactId];
	c.contactId = contactId;
	c.setIndex = setIndex;
	c.colorIndex = kNull;
	Arr<int32_t>& list = setIndex == kAwakeSet ? w->awakeContacts : w->disabledContacts;
	c.localIndex = list.count;
	F2D_PUSH( w, list, contactId );
	c.islandId = kNull;
	c.islandPrev = kNull;
	c.islandNext = kNull;
	c.shapeIdA = shapeIdA;
	c.shapeIdB = shapeIdB;
	c.flags = 0;
	if ( shapeA.enableContactEvents || shapeB.enableContactEvents )
		c.flags |= kContactEnableContactEvents;

	{
		c.edges[0].bodyId = shapeA.bodyId;
		c.edges[0].prevKey = kNull;
		c.edges[0].nextKey = bodyA.headContactKey;
		int keyA = ( contactId << 1 ) | 0;
		if ( bodyA.headContactKey != kNull )
			contacts[bodyA.headContactKey >> 1].edges[bodyA.headContactKey & 1].prevKey = keyA;
		bodyA.headContactKey = keyA;
		bodyA.contactCount += 1;
	}
	{
		c.edges[1].bodyId = shapeB.bodyId;
		c.edges[1].prevKey = kNull;
		c.edges[1].nextKey = bodyB.headContactKey;
		int keyB = ( contactId << 1 ) | 1;
		if ( bodyB.headContactKey != kNull )
			contacts[bodyB.headContactKey >> 1].edges[bodyB.headContactKey & 1].prevKey = keyB;
		bodyB.headContactKey = keyB;
		bodyB.contactCount += 1;
	}

	ContactSim& sim = ptr( w, w->contactSims )[contactId];
	sim.bodySimIndexA = kNull;
	sim.bodySimIndexB = kNull;
	sim.invMassA = 0.0f;
	sim.invIA = 0.0f;
	sim.invMassB = 0.0f;
	sim.invIB = 0.0f;
	sim.shapeIdA = shapeIdA;
	sim.shapeIdB = shapeIdB;
	sim.bodyIdA = shapeA.bodyId;
	sim.bodyIdB = shapeB.bodyId;
	sim.pairClass = shapeA.type * kShapeTypeCount + shapeB.type;
	sim.pad1 = 0;
	sim.pad2 = 0;
	memset( &sim.cache, 0, sizeof( sim.cache ) );
	memset( &sim.manifold, 0, sizeof( sim.manifold ) );
	sim.friction = sqrtf( shapeA.friction * shapeB.friction );			  // world.c:88-92 default mixing
	sim.restitution = maxf( shapeA.restitution, shapeB.restitution );	  // world.c:94-98
	sim.rollingResistance = 0.0f;
	sim.tangentSpeed = 0.0f;
	sim.simFlags = 0;
	if ( shapeA.enablePreSolveEvents || shapeB.enablePreSolveEvents )
		sim.simFlags |= kSimEnablePreSolve;
	return contactId;
}

F2D_HD ShapeId makeShapeId( const World* w, const Shape& s )
{
	return ShapeId{ s.id + 1, w->worldId, s.generation };
}

F2D_HDF inline bool wakeBody( World* w, Body& body );

// contact.c:337-454
F2D_HDF inline void destroyContact( World* w, int contactId, bool wakeBodies )
{
	Contact* contacts = ptr( w, w->contacts );
	Contact& c = contacts[contactId];
	Body* bodies = ptr( w, w->bodies );
	Edge& edgeA = c.edges[0];
	Edge& edgeB = c.edges[1];
	int bodyIdA = edgeA.bodyId, bodyIdB = edgeB.bodyId;
	Body& bodyA = bodies[bodyIdA];
	Body& bodyB = bodies[bodyIdB];
	uint32_t flags = c.flags;
	bool touching = ( flags & kContactTouching ) != 0;

	if ( touching && ( flags & kContactEnableContactEvents ) != 0 )
	{
		const Shape* shapes = ptr( w, w->shapes );
		EndTouchEvent ev = { makeShapeId( w, shapes[c.shapeIdA] ), makeShapeId( w, shapes[c.shapeIdB] ) };
		F2D_PUSH_EVENT( w, w->endEvents[w->endEventArrayIndex], ev );
	}

	if ( edgeA.prevKey != kNull )
		contacts[edgeA.prevKey >> 1].edges[edgeA.prevKey & 1].nextKey = edgeA.nextKey;
	if ( edgeA.nextKey != kNull )
		contacts[edgeA.nextKey >> 1].edges[edgeA.nextKey & 1].prevKey = edgeA.prevKey;
	int edgeKeyA = ( contactId << 1 ) | 0;
	if ( bodyA.headContactKey == edgeKeyA )
		bodyA.headContactKey = edgeA.nextKey;
	bodyA.contactCount -= 1;

	if ( edgeB.prevKey != kNull )
		contacts[edgeB.prevKey >> 1].edges[edgeB.prevKey & 1].nextKey = edgeB.nextKey;
	if ( edgeB.nextKey != kNull )
		contacts[edgeB.nextKey >> 1].edges[edgeB.nextKey & 1].prevKey = edgeB.prevKey;
	int edgeKeyB = ( contactId << 1 ) | 1;
	if ( bodyB.headContactKey == edgeKeyB )
		bodyB.headContactKey = edgeB.nextKey;
	bodyB.contactCount -= 1;

	if ( c.islandId != kNull )
		unlinkContact( w, c, contactId );

	if ( c.colorIndex != kNull )
		removeContactFromGraph( w, bodyIdA, bodyIdB, c.colorIndex, c.localIndex );
	else
		removeContactFromSetList( w, c );

	c.contactId = kNull;
	c.setIndex = kNull;
	c.colorIndex = kNull;
	c.localIndex = kNull;
	freeId( w, w->contactIds, contactId );

	if ( wakeBodies && touching )
	{
		wakeBody( w, bodyA );
		wakeBody( w, bodyB );
	}
}

// What the narrowphase keeps of the previous manifold: feature ids and accumulated impulses (contact.c:552-586)
struct OldImpulses
{
	int32_t pointCount;
	float rollingImpulse;
	uint16_t id[2];
	float normalImpulse[2], tangentImpulse[2];
};

// Second half of b2UpdateContact, from the point where the pre-solve verdict is known (contact.c:519-633): `m` is the
// raw result of the manifold function, `touching` what the manifold (and the host's pre-solve callback, if any) decided.
F2D_HDF inline bool finishContactUpdate( World* w, ContactSim& sim, uint32_t& simFlags, Manifold& m, OldImpulses old, bool touching,
										 bool hitEvents, V2 centerOffsetA, V2 centerOffsetB )
{
	int pointCount = m.pointCount;
	if ( w->enableSpeculative == false && pointCount == 2 )
	{
		if ( m.points[0].separation > 1.5f * kLinearSlop )
		{
			m.points[0] = m.points[1];
			m.pointCount = 1;
		}
		else if ( m.points[0].separation > 1.5f * kLinearSlop )
		{
			m.pointCount = 1;
		}
		pointCount = m.pointCount;
	}

	if ( touching && hitEvents )
		simFlags |= kSimEnableHitEvent;
	else
		simFlags &= ~kSimEnableHitEvent;

	if ( pointCount > 0 )
		m.rollingImpulse = old.rollingImpulse;

#pragma unroll
	for ( int i = 0; i < 2; ++i )
	{
		if ( i >= pointCount )
			break;
		ManifoldPoint& mp2 = m.points[i];
		mp2.anchorA = sub( mp2.anchorA, centerOffsetA );
		mp2.anchorB = sub( mp2.anchorB, centerOffsetB );
		mp2.normalImpulse = 0.0f;
		mp2.tangentImpulse = 0.0f;
		mp2.totalNormalImpulse = 0.0f;
		mp2.normalVelocity = 0.0f;
		mp2.persisted = false;
		uint16_t id2 = mp2.id;
#pragma unroll
		for ( int j = 0; j < 2; ++j )
		{
			if ( j < old.pointCount && old.id[j] == id2 )
			{
				mp2.normalImpulse = old.normalImpulse[j];
				mp2.tangentImpulse = old.tangentImpulse[j];
				mp2.persisted = true;
				old.normalImpulse[j] = 0.0f;
				old.tangentImpulse[j] = 0.0f;
				break;
			}
		}
	}
	packManifold( sim.manifold, m );

	if ( touching )
		simFlags |= kSimTouching;
	else
		simFlags &= ~kSimTouching;
	return touching; // the caller stores simFlags once, after adding the transition bits
}

// Callback-mediated narrowphase: what finishContactUpdate needs of the OLD manifold is parked in fields of the raw NEW
// manifold that are still zero at this point (impulses, total impulse, normal velocity, rolling impulse), so a contact
// waiting for the host's pre-solve verdict needs no storage of its own. unparkOldImpulses restores the zeros.
F2D_HDF inline void parkOldImpulses( Manifold& m, const OldImpulses& old )
{
	for ( int j = 0; j < 2; ++j )
	{
		m.points[j].normalImpulse = old.normalImpulse[j];
		m.points[j].tangentImpulse = old.tangentImpulse[j];
		m.points[j].normalVelocity = floatFromBits( old.id[j] );
	}
	m.points[0].totalNormalImpulse = floatFromBits( (uint32_t)old.pointCount );
	m.rollingImpulse = old.rollingImpulse;
}
F2D_HDF inline OldImpulses unparkOldImpulses( Manifold& m )
{
	OldImpulses old;
	old.pointCount = (int32_t)floatBits( m.points[0].totalNormalImpulse );
	old.rollingImpulse = m.rollingImpulse;
	for ( int j = 0; j < 2; ++j )
	{
		old.id[j] = (uint16_t)floatBits( m.points[j].normalVelocity );
		old.normalImpulse[j] = m.points[j].normalImpulse;
		old.tangentImpulse[j] = m.points[j].tangentImpulse;
		m.points[j].normalImpulse = 0.0f;
		m.points[j].tangentImpulse = 0.0f;
		m.points[j].totalNormalImpulse = 0.0f;
		m.points[j].normalVelocity = 0.0f;
	}
	m.rollingImpulse = 0.0f;
	return old;
}

// Narrowphase update of one contact: contact.c:472-633. `old` was read together with the ids of the contact, so this
// routine issues no load that depends on another one. With a pre-solve callback registered (b2PreSolveFcn runs on the
// host), a touching contact with pre-solve events stops after the manifold function: kSimPendingPreSolve is set, the
// raw manifold is stored with the old impulses parked in it, and finishDeferredContact completes it once the host has
// answered (contact.c:504-517).
F2D_HDF inline bool updateContact( World* w, ContactSim& sim, uint32_t& simFlags, OldImpulses old, const Shape& shapeA, Xf xfA,
								   V2 centerOffsetA, const Shape& shapeB, Xf xfB, V2 centerOffsetB )
{
	// the material of both shapes: requested before the manifold function runs (the loads are in flight beside it), and
	// all of it before the first store below (a load cannot move above a store that might alias it)
	const float frictionA = shapeA.friction, frictionB = shapeB.friction;
	const float restitutionA = shapeA.restitution, restitutionB = shapeB.restitution;
	const float rollingA = shapeA.rollingResistance, rollingB = shapeB.rollingResistance;
	const float tangentSpeedA = shapeA.tangentSpeed, tangentSpeedB = shapeB.tangentSpeed;
	const bool hitEvents = shapeA.enableHitEvents | shapeB.enableHitEvents;
	F2D_ISSUE_F( frictionA );
	F2D_ISSUE_F( frictionB );
	F2D_ISSUE_F( restitutionA );
	F2D_ISSUE_F( restitutionB );
	F2D_ISSUE_F( rollingA );
	F2D_ISSUE_F( rollingB );
	F2D_ISSUE_F( tangentSpeedA );
	F2D_ISSUE_F( tangentSpeedB );
	F2D_ISSUE_I( (int)hitEvents );
	Manifold m = computeManifold( w, shapeA, xfA, shapeB, xfB, &sim.cache );

	float rollingResistance = 0.0f;
	if ( ( rollingA > 0.0f ) | ( rollingB > 0.0f ) )
	{
		float maxRadius = maxf( shapeRadius( shapeA ), shapeRadius( shapeB ) );
		rollingResistance = maxf( rollingA, rollingB ) * maxRadius;
	}
	sim.friction = sqrtf( frictionA * frictionB );
	sim.restitution = maxf( restitutionA, restitutionB );
	sim.rollingResistance = rollingResistance;
	sim.tangentSpeed = tangentSpeedA + tangentSpeedB;

	bool touching = m.pointCount > 0;
	if ( touching && ( w->hostCallbacks & kHostPreSolve ) != 0 && ( simFlags & kSimEnablePreSolve ) != 0 )
	{
		parkOldImpulses( m, old );
		packManifold( sim.manifold, m );
		simFlags |= kSimPendingPreSolve;
		return false; // not decided yet: the caller must not derive a touching transition
	}
	return finishContactUpdate( w, sim, simFlags, m, old, touching, hitEvents, centerOffsetA, centerOffsetB );
}

// ------------------------------------------------------------------------------------------------ sleeping sets
// Sleep-pool block allocator (ours): address-ordered block list, bump allocation, compaction on demand.
F2D_HDF inline void sleepPoolCompact( World* w )
{
	SolverSet* sets = ptr( w, w->sets );
	int32_t* pool = ptr( w, w->sleepPool );
	int cursor = 0;
	for ( int id = w->sleepHead; id != kNull; id = sets[id].nextBlock )
	{
		SolverSet& s = sets[id];
		int len = s.bodyCap + s.contactCap + s.jointCap + s.islandCap;
		if ( s.blockOff != cursor )
		{
			for ( int k = 0; k < len; ++k )
				pool[cursor + k] = pool[s.blockOff + k];
			s.blockOff = cursor;
		}
		cursor += len;
	}
	w->sleepUsed = cursor;
}

F2D_HDF inline bool sleepPoolAlloc( World* w, SolverSet& s, int setId, int bodyCap, int contactCap, int jointCap, int islandCap )
{
	int len = bodyCap + contactCap + jointCap + islandCap;
	if ( w->sleepUsed + len > w->sleepPool.cap )
		sleepPoolCompact( w );
	if ( w->sleepUsed + len > w->sleepPool.cap )
	{
		setError( w, kErrSleepPool, __LINE__ );
		return false;
	}
	s.blockOff = w->sleepUsed;
	s.bodyCap = bodyCap;
	s.contactCap = contactCap;
	s.jointCap = jointCap;
	s.islandCap = islandCap;
	s.bodyCount = s.contactCount = s.jointCount = s.islandCount = 0;
	w->sleepUsed += len;
	SolverSet* sets = ptr( w, w->sets );
	s.prevBlock = w->sleepTail;
	s.nextBlock = kNull;
	if ( w->sleepTail != kNull )
		sets[w->sleepTail].nextBlock = setId;
	else
		w->sleepHead = setId;
	w->sleepTail = setId;
	return true;
}

F2D_HDF inline void sleepPoolFree( World* w, SolverSet& s, int setId )
{
	SolverSet* sets = ptr( w, w->sets );
	if ( s.prevBlock != kNull )
		sets[s.prevBlock].nextBlock = s.nextBlock;
	else
		w->sleepHead = s.nextBlock;
	if ( s.nextBlock != kNull )
		sets[s.nextBlock].prevBlock = s.prevBlock;
	else
	{
		w->sleepTail = s.prevBlock;
		w->sleepUsed = s.prevBlock == kNull ? 0
											 : sets[s.prevBlock].blockOff + sets[s.prevBlock].bodyCap + sets[s.prevBlock].contactCap +
												   sets[s.prevBlock].jointCap + sets[s.prevBlock].islandCap;
	}
	(void)setId;
	s.prevBlock = s.nextBlock = kNull;
}

// Allocates a solver-set slot (solver_set.c:165-174 id reuse)
F2D_HDF inline int allocSolverSet( World* w )
{
	int setId = allocId( w, w->setIds );
	if ( setId == w->sets.count )
	{
		SolverSet empty;
		memset( &empty, 0, sizeof( empty ) );
		empty.setIndex = kNull;
		F2D_PUSH( w, w->sets, empty );
	}
	return setId;
}

// solver_set.c:22-35
F2D_HDF inline void destroySolverSet( World* w, int setIndex )
{
	SolverSet& s = ptr( w, w->sets )[setIndex];
	sleepPoolFree( w, s, setIndex );
	freeId( w, w->setIds, setIndex );
	memset( &s, 0, sizeof( s ) );
	s.setIndex = kNull;
	s.prevBlock = s.nextBlock = kNull;
}

F2D_HD BodyState identityState()
{
	BodyState s;
	s.v = V2{ 0.0f, 0.0f };
	s.w = 0.0f;
	s.flags = 0;
	s.dp = V2{ 0.0f, 0.0f };
	s.dq = Rot{ 1.0f, 0.0f };
	return s;
}

// solver_set.c:37-154
F2D_HDF inline void wakeSolverSet( World* w, int setIndex )
{
	SolverSet& set = ptr( w, w->sets )[setIndex];
	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	Island* islands = ptr( w, w->islands );

	const int32_t* bodyList = setBodyList( w, set );
	for ( int i = 0; i < set.bodyCount; ++i )
	{
		Body& body = bodies[bodyList[i]];
		body.setIndex = kAwakeSet;
		body.localIndex = w->awakeBodies.count;
		body.sleepTime = 0.0f;
		F2D_PUSH( w, w->awakeBodies, body.id );
		F2D_PUSH( w, w->states, identityState() );

		int contactKey = body.headContactKey;
		while ( contactKey != kNull )
		{
			int edgeIndex = contactKey & 1;
			int contactId = contactKey >> 1;
			Contact& c = contacts[contactId];
			contactKey = c.edges[edgeIndex].nextKey;
			if ( c.setIndex != kDisabledSet )
				continue;
			int localIndex = c.localIndex;
			c.setIndex = kAwakeSet;
			c.localIndex = w->awakeContacts.count;
			F2D_PUSH( w, w->awakeContacts, contactId );
			int moved = removeSwap( w, w->disabledContacts, localIndex );
			if ( moved != kNull )
				contacts[ptr( w, w->disabledContacts )[localIndex]].localIndex = localIndex;
		}
	}
	{
		const int32_t* list = setContactList( w, set );
		for ( int i = 0; i < set.contactCount; ++i )
		{
			addContactToGraph( w, list[i] );
			contacts[list[i]].setIndex = kAwakeSet;
		}
	}
	{
		const int32_t* list = setJointList( w, set );
		for ( int i = 0; i < set.jointCount; ++i )
		{
			addJointToGraph( w, list[i] );
			joints[list[i]].setIndex = kAwakeSet;
		}
	}
	{
		const int32_t* list = setIslandList( w, set );
		for ( int i = 0; i < set.islandCount; ++i )
		{
			Island& is = islands[list[i]];
			is.setIndex = kAwakeSet;
			is.localIndex = w->awakeIslands.count;
			F2D_PUSH( w, w->awakeIslands, list[i] );
		}
	}
	destroySolverSet( w, setIndex );
}

// solver_set.c:156-422
F2D_HDF inline void trySleepIsland( World* w, int islandId )
{
	Island* islands = ptr( w, w->islands );
	Island& island = islands[islandId];
	if ( island.constraintRemoveCount > 0 )
		return;

	int sleepSetId = allocSolverSet( w );
	SolverSet& sleepSet = ptr( w, w->sets )[sleepSetId];
	memset( &sleepSet, 0, sizeof( sleepSet ) );
	sleepSet.setIndex = sleepSetId;
	if ( sleepPoolAlloc( w, sleepSet, sleepSetId, island.bodyCount, island.contactCount, island.jointCount, 1 ) == false )
	{
		// cannot sleep this island: give the id back and leave it awake (flagged as an error)
		sleepSet.setIndex = kNull;
		sleepSet.prevBlock = sleepSet.nextBlock = kNull;
		freeId( w, w->setIds, sleepSetId );
		return;
	}

	Body* bodies = ptr( w, w->bodies );
	Contact* contacts = ptr( w, w->contacts );
	Joint* joints = ptr( w, w->joints );
	int32_t* awakeBodies = ptr( w, w->awakeBodies );
	BodyState* states = ptr( w, w->states );
	BodyMoveEvent* moveEvents = ptr( w, w->moveEvents );

	{
		int32_t* list = setBodyList( w, sleepSet );
		int bodyId = island.headBody;
		while ( bodyId != kNull )
		{
			Body& body = bodies[bodyId];
			if ( body.bodyMoveIndex != kNull )
			{
				moveEvents[body.bodyMoveIndex].fellAsleep = true;
				body.bodyMoveIndex = kNull;
			}
			int awakeIndex = body.localIndex;
			int sleepIndex = sleepSet.bodyCount;
			list[sleepSet.bodyCount++] = bodyId;

			int moved = removeSwap( w, w->awakeBodies, awakeIndex );
			if ( moved != kNull )
				bodies[awakeBodies[awakeIndex]].localIndex = awakeIndex;
			removeSwap( w, w->states, awakeIndex );
			(void)states;

			body.setIndex = sleepSetId;
			body.localIndex = sleepIndex;

			int contactKey = body.headContactKey;
			while ( contactKey != kNull )
			{
				int contactId = contactKey >> 1;
				int edgeIndex = contactKey & 1;
				Contact& c = contacts[contactId];
				contactKey = c.edges[edgeIndex].nextKey;
				if ( c.setIndex == kDisabledSet )
					continue;
				if ( c.colorIndex != kNull )
					continue;
				int otherBodyId = c.edges[edgeIndex ^ 1].bodyId;
				if ( bodies[otherBodyId].setIndex == kAwakeSet )
					continue;
				int localIndex = c.localIndex;
				c.setIndex = kDisabledSet;
				c.localIndex = w->disabledContacts.count;
				F2D_PUSH( w, w->disabledContacts, contactId );
				int movedC = removeSwap( w, w->awakeContacts, localIndex );
				if ( movedC != kNull )
					contacts[ptr( w, w->awakeContacts )[localIndex]].localIndex = localIndex;
			}
			bodyId = body.islandNext;
		}
	}
	{
		int32_t* list = setContactList( w, sleepSet );
		int contactId = island.headContact;
		while ( contactId != kNull )
		{
			Contact& c = contacts[contactId];
			int colorIndex = c.colorIndex;
			removeContactFromGraph( w, c.edges[0].bodyId, c.edges[1].bodyId, colorIndex, c.localIndex );
			int sleepIndex = sleepSet.contactCount;
			list[sleepSet.contactCount++] = contactId;
			c.setIndex = sleepSetId;
			c.colorIndex = kNull;
			c.localIndex = sleepIndex;
			contactId = c.islandNext;
		}
	}
	{
		int32_t* list = setJointList( w, sleepSet );
		int jointId = island.headJoint;
		while ( jointId != kNull )
		{
			Joint& j = joints[jointId];
			removeJointFromGraph( w, j.edges[0].bodyId, j.edges[1].bodyId, j.colorIndex, j.localIndex );
			int sleepIndex = sleepSet.jointCount;
			list[sleepSet.jointCount++] = jointId;
			j.setIndex = sleepSetId;
			j.colorIndex = kNull;
			j.localIndex = sleepIndex;
			jointId = j.islandNext;
		}
	}
	{
		int islandIndex = island.localIndex;
		setIslandList( w, sleepSet )[0] = islandId;
		sleepSet.islandCount = 1;
		int moved = removeSwap( w, w->awakeIslands, islandIndex );
		if ( moved != kNull )
			islands[ptr( w, w->awakeIslands )[islandIndex]].localIndex = islandIndex;
		island.setIndex = sleepSetId;
		island.localIndex = 0;
	}
}

// body.c:~480 b2WakeBody
F2D_HDF inline bool wakeBody( World* w, Body& body )
{
	if ( body.setIndex >= kFirstSleepingSet )
	{
		wakeSolverSet( w, body.setIndex );
		return true;
	}
	return false;
}

} // namespace f2d
