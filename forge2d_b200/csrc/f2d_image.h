// forge2d_b200 — host-side management of world images: capacities, layout, growth (re-layout).
// Host only. The layout is a pure function of the capacities, so host and device copies always agree.
#pragma once
#include "f2d_create.h"

#include <stdlib.h>
#include <vector>

namespace f2d
{

struct Caps
{
	int bodies, shapes, contacts, joints;
	int contactEvents, hitEvents; // event array capacities
	int sensors;				  // sensor shapes
	int sensorOverlap;			  // capacity of one overlap list of one sensor (derived: sensorOverlapCapFor)
};

// An overlap list holds shapes that overlap the sensor, each at most once: `shapes` entries always suffice. The total
// is bounded (64 MB of references); a world beyond that keeps the bounded capacity and drops overlaps (kErrTruncated).
inline int sensorOverlapCapFor( int shapes, int sensors )
{
	int cap = shapes < kMinSensorOverlapCap ? kMinSensorOverlapCap : shapes;
	const long long budget = ( 64ll << 20 ) / (long long)sizeof( ShapeRef ) / 2;
	const int n = sensors < 1 ? 1 : sensors;
	if ( (long long)cap * n > budget )
		cap = (int)( budget / n );
	return cap < kMinSensorOverlapCap ? kMinSensorOverlapCap : cap;
}

struct ArraySlot
{
	uint64_t* off;
	int32_t* count;
	int32_t* cap;
	int elemSize;
	int newCap;
	bool persistent; // contents survive a re-layout (count elements are copied)
};

// Enumerates every array of the image in a fixed order together with its capacity for `c`.
inline void collectArrays( World& w, const Caps& c, std::vector<ArraySlot>& out )
{
	auto add = [&]( auto& arr, int cap, bool persistent ) {
		using T = typename std::remove_reference<decltype( *ptr( &w, arr ) )>::type;
		out.push_back( ArraySlot{ &arr.off, &arr.count, &arr.cap, (int)sizeof( T ), cap < 4 ? 4 : cap, persistent } );
	};
	const int B = c.bodies, S = c.shapes, C = c.contacts, J = c.joints;
	add( w.bodyIds.free, B, true );
	add( w.shapeIds.free, S, true );
	add( w.contactIds.free, C, true );
	add( w.jointIds.free, J, true );
	add( w.islandIds.free, B, true );
	add( w.setIds.free, B + 8, true );
	add( w.chainIds.free, S, true ); // a chain owns at least one shape: never more chain ids than shape slots
	add( w.bodies, B, true );
	add( w.sims, B, true );
	add( w.shapes, S, true );
	add( w.contacts, C, true );
	add( w.contactSims, C, true );
	add( w.joints, J, true );
	add( w.jointSims, J, true );
	add( w.islands, B, true );
	add( w.sets, B + 8, true );
	const int N = c.sensors < 1 ? 1 : c.sensors;
	const int overlapCap = sensorOverlapCapFor( S, c.sensors );
	add( w.sensors, N, true );
	add( w.sensorRefs, 2 * overlapCap * N, true );
	add( w.sensorBits, N / 64 + 2, false );
	add( w.staticBodies, B, true );
	add( w.disabledBodies, B, true );
	add( w.awakeBodies, B, true );
	add( w.states, B, true );
	add( w.awakeContacts, C, true );
	add( w.disabledContacts, C, true );
	add( w.disabledJoints, J, true );
	add( w.staticJoints, J, true );
	add( w.awakeIslands, B, true );
	add( w.sleepPool, 2 * ( 2 * B + C + J ) + 64, true );
	for ( int i = 0; i < kColorCount; ++i )
	{
		// a colour holds each non-static body at most once; the overflow colour can hold everything
		add( w.colorContacts[i], i == kOverflow ? C : B, true );
		add( w.colorJoints[i], i == kOverflow ? J : B, true );
	}
	for ( int i = 0; i < 3; ++i )
	{
		add( w.trees[i].nodes, 2 * S + 16, true );
		add( w.trees[i].leafIndices, S + 4, false );
		add( w.trees[i].leafCenters, S + 4, false );
		add( w.trees[i].work, kTreeStack * 6, false );
	}
	add( w.treeScratch, 24 * ( S + 8 ) + 2 * S + 16 + 8, false );
	add( w.moveArray, S, true );
	add( w.moveHeads, S, false );
	add( w.movePairs, C, false );
	add( w.pairOffsets, S, false );
	add( w.pairOrder, C, false );
	add( w.moveEvents, B, true );
	add( w.beginEvents, c.contactEvents, true );
	add( w.endEvents[0], c.contactEvents, true );
	add( w.endEvents[1], c.contactEvents, true );
	add( w.hitEvents, c.hitEvents, true );
	// a step can begin or end at most one overlap per list slot of every sensor
	const int SE = c.sensors < 1 ? 4 : 2 * overlapCap * c.sensors;
	add( w.sensorBeginEvents, SE, true );
	add( w.sensorEndEvents[0], SE, true );
	add( w.sensorEndEvents[1], SE, true );
	add( w.contactBits, C / 64 + 2, false );
	add( w.stateOffsets, C / 64 + 2, false );
	add( w.stateList, C, false );
	add( w.enlargedBits, B / 64 + 2, false );
	add( w.islandBits, B / 64 + 2, false );
	add( w.cons, cfFieldCount * C, false );
	add( w.islSlotOff, B + 8, false );
	add( w.islBodyOff, B + 8, false );
	add( w.islColorOff, kColorCount * B + 8, false );
	add( w.islColorFill, kColorCount * B + 8, false );
	add( w.islBodyFill, B + 8, false );
	add( w.islSlots, C, false );
	add( w.islBodies, B, false );
	add( w.bullets, B, false );
	add( w.integ, 6 * B, false );
	add( w.scan, B + 8, false );
	add( w.scratch, 2 * B + 64, false );
	add( w.splitScratch, 14 * ( B + 8 ) + 7 * ( C + 8 ) + 7 * ( J + 8 ), false );
}

inline uint64_t alignUp( uint64_t v, uint64_t a ) { return ( v + a - 1 ) / a * a; }

// Assigns offsets/capacities for `c` in `w` (counts untouched) and returns the image size in bytes.
inline uint64_t layoutImage( World& w, const Caps& c )
{
	std::vector<ArraySlot> slots;
	collectArrays( w, c, slots );
	uint64_t cursor = alignUp( sizeof( World ), 256 );
	for ( ArraySlot& s : slots )
	{
		*s.off = cursor;
		*s.cap = s.newCap;
		cursor = alignUp( cursor + (uint64_t)s.elemSize * (uint64_t)s.newCap, 256 );
	}
	w.consStride = c.contacts < 4 ? 4 : c.contacts;
	w.sensorOverlapCap = sensorOverlapCapFor( c.shapes, c.sensors );
	w.imageBytes = cursor;
	return cursor;
}

typedef void* ( *ImageAlloc )( size_t );
typedef void ( *ImageFree )( void* );

// New empty world image with the defaults of b2CreateWorld (B2/src/world.c:100-261)
inline World* imageCreate( const Caps& c, ImageAlloc alloc )
{
	World tmp;
	memset( &tmp, 0, sizeof( tmp ) );
	uint64_t bytes = layoutImage( tmp, c );
	World* w = static_cast<World*>( alloc( bytes ) );
	memset( w, 0, bytes );
	*w = tmp;
	w->magic = kWorldMagic;
	w->splitIslandId = kNull;
	w->sleepHead = w->sleepTail = kNull;
	w->enableWarmStarting = true;
	w->enableSpeculative = true;
	for ( int i = 0; i < 3; ++i )
	{
		w->trees[i].root = kNull;
		w->trees[i].freeList = kNull;
	}
	// the three fixed solver sets: static 0, disabled 1, awake 2 (world.c:140-157)
	for ( int i = 0; i < 3; ++i )
	{
		int id = allocSolverSet( w );
		SolverSet& s = ptr( w, w->sets )[id];
		s.setIndex = id;
		s.prevBlock = s.nextBlock = kNull;
	}
	return w;
}

// Re-layout into larger capacities, preserving every persistent array and the header.
inline World* imageRelayout( World* old, const Caps& c, ImageAlloc alloc, ImageFree release )
{
	World header = *old;
	uint64_t bytes = layoutImage( header, c );
	World* w = static_cast<World*>( alloc( bytes ) );
	memset( w, 0, bytes );
	*w = header;

	Caps dummy = c;
	std::vector<ArraySlot> oldSlots, newSlots;
	World oldHeaderCopy = *old;
	oldHeaderCopy.sleepPool.count = old->sleepUsed; // the pool is bump-allocated; its Arr::count is not maintained
	collectArrays( oldHeaderCopy, dummy, oldSlots ); // offsets/counts as stored in the old header
	collectArrays( *w, c, newSlots );
	for ( size_t i = 0; i < newSlots.size(); ++i )
	{
		if ( newSlots[i].persistent == false )
			continue;
		int count = *oldSlots[i].count;
		// arrays addressed by id rather than by count keep their whole old capacity
		int n = count;
		if ( n > *oldSlots[i].cap )
			n = *oldSlots[i].cap;
		memcpy( reinterpret_cast<char*>( w ) + *newSlots[i].off, reinterpret_cast<const char*>( old ) + *oldSlots[i].off,
				(size_t)n * (size_t)newSlots[i].elemSize );
	}
	// the overlap lists are blocks of the per-sensor capacity: when that changed, every block moves
	if ( old->sensorOverlapCap != w->sensorOverlapCap && old->sensors.count > 0 )
	{
		const ShapeRef* from = reinterpret_cast<const ShapeRef*>( reinterpret_cast<const char*>( old ) + old->sensorRefs.off );
		ShapeRef* to = reinterpret_cast<ShapeRef*>( reinterpret_cast<char*>( w ) + w->sensorRefs.off );
		const int oldCap = old->sensorOverlapCap, newCap = w->sensorOverlapCap;
		const int keep = oldCap < newCap ? oldCap : newCap;
		memset( to, 0, sizeof( ShapeRef ) * (size_t)w->sensorRefs.cap );
		for ( int list = 0; list < 2 * old->sensors.count; ++list )
			memcpy( to + (size_t)list * newCap, from + (size_t)list * oldCap, sizeof( ShapeRef ) * (size_t)keep );
		w->sensorRefs.count = 2 * newCap * w->sensors.count;
	}
	release( old );
	return w;
}

inline int roundCap( int v, int minimum )
{
	int c = minimum;
	while ( c < v )
		c += c >> 1;
	return ( c + 63 ) / 64 * 64;
}

} // namespace f2d
