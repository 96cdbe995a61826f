// forge2d_b200 — implementation of the C ABI declared in include/forge2d_b200.h.
// Included by exactly one translation unit per library:
//   * forge2d_b200/csrc/f2d_cuda.cu   (the product: steps run as CUDA kernels, no CPU path)
//   * tests/emu/f2d_emu.cpp           (test-only host emulation of the same step code, never shipped)
// The including TU provides the backend hooks declared below.
#include "f2d_image.h"
#include "f2d_mutate.h"
#include "f2d_query.h"
#include "f2d_step.h"

#include "../../include/forge2d_b200.h"
#include "../../include/forge2d_b200_debug.h"

#include <algorithm>
#include <float.h>
#include <stdio.h>
#include <string>
#include <vector>

namespace f2d
{

constexpr int kMaxWorlds = 128; // B2_MAX_WORLDS, constants.h:26-28
constexpr int kSecretCookie = 1152023; // core.h:116

enum SyncState
{
	kInSync,
	kHostNewer,
	kDeviceNewer
};

struct HostWorld
{
	World* img = nullptr; // host image (pinned when the CUDA backend is active)
	Caps caps{};
	SyncState state = kInSync;
	bool inUse = false;
	uint16_t generation = 0;
	int launchMode = -1;
	void* backend = nullptr; // device mirror, streams, events
	bool eventsFresh = true; // event arrays of `img` reflect the last step
	// Chain records (shape.h b2ChainShape): host-only bookkeeping — the step only ever sees the chain-SEGMENT shapes
	struct Chain
	{
		int id = -1, bodyId = -1, nextChainId = -1;
		uint16_t generation = 0;
		std::vector<int> shapeIndices;
		std::vector<b2SurfaceMaterial> materials;
	};
	std::vector<Chain> chains;
	bool headerFresh = true; // false between f2dWorld_StepAsync and the arrival of that step's header (headerCurrent)
	float asyncTimeStep = 0.0f;
	int asyncSubSteps = 0;
	// Partial host <-> device synchronisation while the device copy is the newer one (SyncState::kDeviceNewer):
	//   bodyMirrorFresh  the host copies of the body records, simulation records and solver states are current (one
	//                    download of those three arrays after a step serves every position / velocity getter);
	//   dirty            byte ranges of the host image that per-frame mutators (forces, impulses, velocities of awake
	//                    bodies) have changed; uploaded right before the next step instead of the whole image.
	bool bodyMirrorFresh = false;
	std::vector<std::pair<uint64_t, uint32_t>> dirty;
	bool wholeImageSync = false; // the next download / upload moves the whole image, work arrays included (host work in the middle of a step)
	unsigned layoutSerial = 0; // bumped by every re-layout of the image: the device copy of another layout is replaced whole
	// b2World_GetProfile: the in-kernel phase marks as they stood at the previous call
	uint64_t profSeen[kProfSlots] = {};
	uint64_t profStepSeen = 0;
	// host callbacks of the callback-mediated step (world.c:1710-1740); the image only carries World::hostCallbacks
	b2CustomFilterFcn* customFilterFcn = nullptr;
	void* customFilterContext = nullptr;
	b2PreSolveFcn* preSolveFcn = nullptr;
	void* preSolveContext = nullptr;
};

static HostWorld g_worlds[kMaxWorlds];
static std::string g_lastError;
static long long g_launchCount = 0;

static void reportError( const char* fmt, int a = 0, int b = 0 )
{
	char buf[512];
	snprintf( buf, sizeof( buf ), fmt, a, b );
	g_lastError = buf;
	fprintf( stderr, "forge2d_b200: %s\n", buf );
}

// ---- backend hooks (defined by the including translation unit) ------------------------------------------------
static void* backendHostAlloc( size_t bytes );
static void backendHostFree( void* p );
static bool backendAvailable();
/// Runs one step on the authoritative copy; after return `hw.img`'s HEADER (sizeof(World)) is current.
static void backendStep( HostWorld& hw, float dt, int subSteps, bool synchronous );
static void backendSynchronize( HostWorld& hw );
/// Makes the full host image current (device -> host) when the device copy is newer.
static void backendDownload( HostWorld& hw );
/// Makes one byte range of the host image current.
static void backendDownloadRange( HostWorld& hw, uint64_t off, uint64_t bytes );
struct ByteRange
{
	uint64_t off, bytes;
};
static void backendDownloadRanges( HostWorld& hw, std::initializer_list<ByteRange> ranges ); // one synchronisation for all
/// Callback-mediated step: the step as a sequence of launches (f2d_step.h Phase) with the host in between.
/// Begin makes the authoritative copy current, Phase runs one piece, UploadRange pushes host edits of one byte range of
/// the image, End leaves `hw.img`'s header current like backendStep. DownloadRange synchronises with the phases queued.
static void backendPhaseBegin( HostWorld& hw );
static void backendPhase( HostWorld& hw, float dt, int subSteps, int phase );
static void backendUploadRange( HostWorld& hw, uint64_t off, uint64_t bytes );
static void backendPhaseEnd( HostWorld& hw );
static void backendRelease( HostWorld& hw );
static void backendStepTimes( HostWorld& hw, float* out5 );
static void backendEnableTiming( HostWorld& hw, bool flag );

static HostWorld* worldFromIndex0( int index0 )
{
	if ( index0 < 0 || index0 >= kMaxWorlds || g_worlds[index0].inUse == false )
		return nullptr;
	return g_worlds + index0;
}
static HostWorld* worldFromId( b2WorldId id )
{
	HostWorld* hw = worldFromIndex0( (int)id.index1 - 1 );
	if ( hw == nullptr || hw->generation != id.generation )
		return nullptr;
	return hw;
}

// After f2dWorld_StepAsync the header copy of that step may still be in flight: everything that reads the header of
// `hw.img` (counts, capacities in use, error flags) waits for it first.
static void headerCurrent( HostWorld& hw )
{
	if ( hw.headerFresh == false )
	{
		backendSynchronize( hw );
		hw.headerFresh = true;
	}
}

// Pending host-side edits of a device-newer image go to the device (before a step, or before the image is downloaded)
static void flushDirty( HostWorld& hw )
{
	if ( hw.dirty.empty() )
		return;
	std::sort( hw.dirty.begin(), hw.dirty.end() );
	uint64_t off = hw.dirty[0].first, end = off + hw.dirty[0].second;
	for ( size_t i = 1; i <= hw.dirty.size(); ++i )
	{
		if ( i < hw.dirty.size() && hw.dirty[i].first <= end + 256 ) // close ranges travel as one copy
		{
			end = std::max( end, hw.dirty[i].first + hw.dirty[i].second );
			continue;
		}
		backendUploadRange( hw, off, end - off );
		if ( i < hw.dirty.size() )
		{
			off = hw.dirty[i].first;
			end = off + hw.dirty[i].second;
		}
	}
	hw.dirty.clear();
}

// Host image current and writable
static World* hostImage( HostWorld& hw )
{
	headerCurrent( hw );
	if ( hw.state == kDeviceNewer )
	{
		flushDirty( hw );
		backendDownload( hw );
		hw.state = kInSync;
		hw.eventsFresh = true;
		hw.headerFresh = true;
	}
	return hw.img;
}
static World* mutableImage( HostWorld& hw )
{
	World* w = hostImage( hw );
	hw.state = kHostNewer;
	return w;
}

// The body-level part of the host image current (body records, simulation records, solver states; the header always
// is): what position / velocity getters and the per-frame mutators need. After a step this costs one download of those
// three arrays (a few hundred KB for bench2d) instead of the whole image (megabytes).
static World* bodyImage( HostWorld& hw )
{
	headerCurrent( hw );
	if ( hw.state == kDeviceNewer && hw.bodyMirrorFresh == false )
	{
		World* w = hw.img;
		flushDirty( hw );
		backendDownloadRanges( hw, { { w->bodies.off, (uint64_t)w->bodies.count * sizeof( Body ) },
									 { w->sims.off, (uint64_t)w->sims.count * sizeof( BodySim ) },
									 { w->states.off, (uint64_t)w->awakeBodies.count * sizeof( BodyState ) } } );
		hw.bodyMirrorFresh = true;
	}
	return hw.img;
}
// A host-side change of `bytes` at `p` inside the image: recorded for upload when the device copy is the newer one
// (otherwise the host image is the master already and the caller has marked it so)
static void touchRange( HostWorld& hw, const void* p, size_t bytes )
{
	if ( hw.state == kDeviceNewer )
		hw.dirty.push_back( { (uint64_t)( reinterpret_cast<const char*>( p ) - reinterpret_cast<const char*>( hw.img ) ), (uint32_t)bytes } );
}

static Caps capsWith( const World* w, int B, int S, int C, int J )
{
	Caps c;
	c.bodies = B;
	c.shapes = S;
	c.contacts = C;
	c.joints = J;
	c.contactEvents = w != nullptr && w->contactEventCapable > 0 ? C : 16;
	c.hitEvents = w != nullptr && w->hitEventCapable > 0 ? C : 16;
	c.sensors = 0;
	c.sensorOverlap = sensorOverlapCapFor( S, 0 );
	return c;
}

// Grows the image when any entity array is short of `need*` free slots.
static void reserve( HostWorld& hw, int needBodies, int needShapes, int needContacts, int needJoints, int needSensors = 0 )
{
	World* w = hw.img;
	Caps c = hw.caps;
	bool grow = false;
	int wantB = w->bodyIds.next + needBodies;
	int wantS = w->shapeIds.next + needShapes;
	int wantC = w->contactIds.next + needContacts;
	int wantJ = w->jointIds.next + needJoints;
	if ( wantB > c.bodies )
	{
		c.bodies = roundCap( wantB, 64 );
		grow = true;
	}
	if ( wantS > c.shapes )
	{
		c.shapes = roundCap( wantS, 64 );
		grow = true;
	}
	if ( wantC > c.contacts )
	{
		c.contacts = roundCap( wantC, 256 );
		grow = true;
	}
	if ( wantJ > c.joints )
	{
		c.joints = roundCap( wantJ, 64 );
		grow = true;
	}
	if ( w->sensors.count + needSensors > c.sensors )
	{
		c.sensors = roundCap( w->sensors.count + needSensors, 8 );
		grow = true;
	}
	if ( sensorOverlapCapFor( c.shapes, c.sensors ) != c.sensorOverlap )
	{
		c.sensorOverlap = sensorOverlapCapFor( c.shapes, c.sensors );
		grow = true;
	}
	Caps ev = capsWith( w, c.bodies, c.shapes, c.contacts, c.joints );
	if ( ev.contactEvents != c.contactEvents || ev.hitEvents != c.hitEvents )
	{
		c.contactEvents = ev.contactEvents;
		c.hitEvents = ev.hitEvents;
		grow = true;
	}
	if ( grow )
	{
		hw.img = imageRelayout( w, c, backendHostAlloc, backendHostFree );
		hw.layoutSerial += 1;
		hw.caps = c;
	}
}

static BodyId toBodyId( b2BodyId id ) { return BodyId{ id.index1, id.world0, id.generation }; }

static bool checkWorldError( HostWorld& hw, const char* where )
{
	World* w = hw.img;
	if ( w->error != 0 )
	{
		char buf[256];
		snprintf( buf, sizeof( buf ), "%s: world error flags 0x%x (detail line %d)%s", where, w->error, w->errorDetail,
				  ( w->error & kErrCapacity ) ? " [capacity exceeded inside the step]" : "" );
		g_lastError = buf;
		fprintf( stderr, "forge2d_b200: %s\n", buf );
		return true;
	}
	return false;
}

} // namespace f2d

using namespace f2d;

extern "C" {

// ---- defaults (B2/src/types.c:9-88) ---------------------------------------------------------------------------
b2WorldDef b2DefaultWorldDef( void )
{
	b2WorldDef def;
	memset( &def, 0, sizeof( def ) );
	def.gravity.x = 0.0f;
	def.gravity.y = -10.0f;
	def.hitEventThreshold = 1.0f;
	def.restitutionThreshold = 1.0f;
	def.maxContactPushSpeed = 3.0f;
	def.contactHertz = 30.0;
	def.contactDampingRatio = 10.0f;
	def.maximumLinearSpeed = 400.0f;
	def.enableSleep = true;
	def.enableContinuous = true;
	def.internalValue = kSecretCookie;
	return def;
}
b2BodyDef b2DefaultBodyDef( void )
{
	b2BodyDef def;
	memset( &def, 0, sizeof( def ) );
	def.type = b2_staticBody;
	def.rotation.c = 1.0f;
	def.rotation.s = 0.0f;
	def.sleepThreshold = 0.05f;
	def.gravityScale = 1.0f;
	def.enableSleep = true;
	def.isAwake = true;
	def.isEnabled = true;
	def.internalValue = kSecretCookie;
	return def;
}
b2Filter b2DefaultFilter( void )
{
	b2Filter f = { 1, UINT64_MAX, 0 };
	return f;
}
b2SurfaceMaterial b2DefaultSurfaceMaterial( void )
{
	b2SurfaceMaterial m;
	memset( &m, 0, sizeof( m ) );
	m.friction = 0.6f;
	return m;
}
b2ShapeDef b2DefaultShapeDef( void )
{
	b2ShapeDef def;
	memset( &def, 0, sizeof( def ) );
	def.material.friction = 0.6f;
	def.density = 1.0f;
	def.filter = b2DefaultFilter();
	def.updateBodyMass = true;
	def.invokeContactCreation = true;
	def.internalValue = kSecretCookie;
	return def;
}
b2RevoluteJointDef b2DefaultRevoluteJointDef( void )
{
	b2RevoluteJointDef def;
	memset( &def, 0, sizeof( def ) );
	def.drawSize = 0.25f;
	def.internalValue = kSecretCookie;
	return def;
}

// ---- geometry helpers (B2/src/geometry.c:133-203) --------------------------------------------------------------
b2Polygon b2MakeBox( float hx, float hy )
{
	b2Polygon s;
	memset( &s, 0, sizeof( s ) );
	s.count = 4;
	s.vertices[0] = b2Vec2{ -hx, -hy };
	s.vertices[1] = b2Vec2{ hx, -hy };
	s.vertices[2] = b2Vec2{ hx, hy };
	s.vertices[3] = b2Vec2{ -hx, hy };
	s.normals[0] = b2Vec2{ 0.0f, -1.0f };
	s.normals[1] = b2Vec2{ 1.0f, 0.0f };
	s.normals[2] = b2Vec2{ 0.0f, 1.0f };
	s.normals[3] = b2Vec2{ -1.0f, 0.0f };
	s.radius = 0.0f;
	s.centroid = b2Vec2{ 0.0f, 0.0f };
	return s;
}
b2Polygon b2MakeSquare( float h )
{
	return b2MakeBox( h, h );
}
b2Polygon b2MakeOffsetRoundedBox( float hx, float hy, b2Vec2 center, b2Rot rotation, float radius )
{
	Xf xf = { { center.x, center.y }, { rotation.c, rotation.s } };
	b2Polygon s;
	memset( &s, 0, sizeof( s ) );
	s.count = 4;
	const V2 lv[4] = { { -hx, -hy }, { hx, -hy }, { hx, hy }, { -hx, hy } };
	const V2 ln[4] = { { 0.0f, -1.0f }, { 1.0f, 0.0f }, { 0.0f, 1.0f }, { -1.0f, 0.0f } };
	for ( int i = 0; i < 4; ++i )
	{
		V2 v = xfPoint( xf, lv[i] );
		V2 n = rotate( xf.q, ln[i] );
		s.vertices[i] = b2Vec2{ v.x, v.y };
		s.normals[i] = b2Vec2{ n.x, n.y };
	}
	s.radius = radius;
	s.centroid = center;
	return s;
}
// geometry.c:41-75 (+ centroid :17-39)
b2Polygon b2MakePolygon( const b2Hull* hull, float radius )
{
	if ( hull->count < 3 )
		return b2MakeSquare( 0.5f );
	b2Polygon s;
	memset( &s, 0, sizeof( s ) );
	s.count = hull->count;
	s.radius = radius;
	for ( int i = 0; i < s.count; ++i )
		s.vertices[i] = hull->points[i];
	for ( int i = 0; i < s.count; ++i )
	{
		int i2 = i + 1 < s.count ? i + 1 : 0;
		V2 edge = sub( V2{ s.vertices[i2].x, s.vertices[i2].y }, V2{ s.vertices[i].x, s.vertices[i].y } );
		V2 n = normalize( crossVS( edge, 1.0f ) );
		s.normals[i] = b2Vec2{ n.x, n.y };
	}
	V2 center = { 0.0f, 0.0f };
	float area = 0.0f;
	V2 origin = { s.vertices[0].x, s.vertices[0].y };
	const float inv3 = 1.0f / 3.0f;
	for ( int i = 1; i < s.count - 1; ++i )
	{
		V2 e1 = sub( V2{ s.vertices[i].x, s.vertices[i].y }, origin );
		V2 e2 = sub( V2{ s.vertices[i + 1].x, s.vertices[i + 1].y }, origin );
		float a = 0.5f * cross( e1, e2 );
		center = mulAdd( center, a * inv3, add( e1, e2 ) );
		area += a;
	}
	float invArea = 1.0f / area;
	center.x *= invArea;
	center.y *= invArea;
	center = add( origin, center );
	s.centroid = b2Vec2{ center.x, center.y };
	return s;
}
} // extern "C"
namespace f2d
{
// Quickhull over <= 8 welded points with the decision rules of B2/src/hull.c:13-82 (signed distance to the right of
// the directed edge a->b, first-maximum tie rule, 2*slop rejection). Emits the chain strictly between a and b, in
// order, through `emit`; the candidate list is passed by value so both sub-chains filter the same right-hand set.
struct HullPoints
{
	V2 p[B2_MAX_POLYGON_VERTICES];
	int n;
};
static void hullChain( V2 a, V2 b, HullPoints cand, b2Hull& out )
{
	if ( cand.n == 0 )
		return;
	V2 e = normalize( sub( b, a ) );
	HullPoints right;
	right.n = 0;
	int best = 0;
	float bestDistance = 0.0f;
	for ( int i = 0; i < cand.n; ++i )
	{
		float d = cross( sub( cand.p[i], a ), e );
		if ( i == 0 || d > bestDistance )
		{
			best = i;
			bestDistance = d;
		}
		if ( d > 0.0f )
			right.p[right.n++] = cand.p[i];
	}
	if ( bestDistance < 2.0f * kLinearSlop )
		return;
	V2 apex = cand.p[best];
	hullChain( a, apex, right, out );
	out.points[out.count++] = b2Vec2{ apex.x, apex.y };
	hullChain( apex, b, right, out );
}
} // namespace f2d
extern "C" {
// collision.h:228, B2/src/hull.c:84-265: weld, pick two extreme points, quickhull both sides, drop collinear points.
b2Hull b2ComputeHull( const b2Vec2* points, int count )
{
	b2Hull hull;
	memset( &hull, 0, sizeof( hull ) );
	if ( count < 3 || count > B2_MAX_POLYGON_VERTICES )
		return hull;
	const float tolSqr = 16.0f * kLinearSlop * kLinearSlop;
	V2 lo = { FLT_MAX, FLT_MAX }, hi = { -FLT_MAX, -FLT_MAX };
	HullPoints ps;
	ps.n = 0;
	for ( int i = 0; i < count; ++i )
	{
		V2 vi = { points[i].x, points[i].y };
		lo = V2{ minf( lo.x, vi.x ), minf( lo.y, vi.y ) };
		hi = V2{ maxf( hi.x, vi.x ), maxf( hi.y, vi.y ) };
		bool unique = true;
		for ( int j = 0; j < i && unique; ++j )
			unique = !( distanceSq( vi, V2{ points[j].x, points[j].y } ) < tolSqr );
		if ( unique )
			ps.p[ps.n++] = vi;
	}
	if ( ps.n < 3 )
		return hull;
	// farthest point from the box centre, then the farthest point from it; each leaves the set by swap-with-last
	auto takeFarthest = [&]( V2 from ) {
		int f = 0;
		float best = distanceSq( from, ps.p[0] );
		for ( int i = 1; i < ps.n; ++i )
		{
			float d = distanceSq( from, ps.p[i] );
			if ( d > best )
			{
				f = i;
				best = d;
			}
		}
		V2 r = ps.p[f];
		ps.p[f] = ps.p[ps.n - 1];
		ps.n -= 1;
		return r;
	};
	V2 center = { 0.5f * ( lo.x + hi.x ), 0.5f * ( lo.y + hi.y ) };
	V2 p1 = takeFarthest( center );
	V2 p2 = takeFarthest( p1 );
	HullPoints right, left;
	right.n = left.n = 0;
	V2 e = normalize( sub( p2, p1 ) );
	for ( int i = 0; i < ps.n; ++i )
	{
		float d = cross( sub( ps.p[i], p1 ), e );
		if ( d >= 2.0f * kLinearSlop )
			right.p[right.n++] = ps.p[i];
		else if ( d <= -2.0f * kLinearSlop )
			left.p[left.n++] = ps.p[i];
	}
	b2Hull side1, side2;
	side1.count = side2.count = 0;
	hullChain( p1, p2, right, side1 );
	hullChain( p2, p1, left, side2 );
	if ( side1.count == 0 && side2.count == 0 )
		return hull;
	hull.points[hull.count++] = b2Vec2{ p1.x, p1.y };
	for ( int i = 0; i < side1.count; ++i )
		hull.points[hull.count++] = side1.points[i];
	hull.points[hull.count++] = b2Vec2{ p2.x, p2.y };
	for ( int i = 0; i < side2.count; ++i )
		hull.points[hull.count++] = side2.points[i];
	// remove nearly collinear middle points, restarting the scan after every removal
	for ( int i = 0; hull.count > 2 && i < hull.count; )
	{
		int i2 = ( i + 1 ) % hull.count, i3 = ( i + 2 ) % hull.count;
		V2 s1 = { hull.points[i].x, hull.points[i].y }, s2 = { hull.points[i2].x, hull.points[i2].y };
		V2 s3 = { hull.points[i3].x, hull.points[i3].y };
		V2 r = normalize( sub( s3, s1 ) );
		if ( cross( sub( s2, s1 ), r ) <= 2.0f * kLinearSlop )
		{
			for ( int j = i2; j < hull.count - 1; ++j )
				hull.points[j] = hull.points[j + 1];
			hull.count -= 1;
			i = 0;
		}
		else
			i += 1;
	}
	if ( hull.count < 3 )
		hull.count = 0;
	return hull;
}

// ---- world -----------------------------------------------------------------------------------------------------
b2WorldId b2CreateWorld( const b2WorldDef* def )
{
	int index = -1;
	for ( int i = 0; i < kMaxWorlds; ++i )
	{
		if ( g_worlds[i].inUse == false )
		{
			index = i;
			break;
		}
	}
	if ( index < 0 )
		return b2WorldId{ 0, 0 };
	HostWorld& hw = g_worlds[index];
	uint16_t generation = hw.generation;
	hw = HostWorld();
	hw.generation = generation;
	hw.inUse = true;
	hw.caps = capsWith( nullptr, 64, 64, 256, 64 );
	hw.img = imageCreate( hw.caps, backendHostAlloc );
	World* w = hw.img;
	w->worldId = (uint16_t)index;
	w->generation = generation;
	w->inUse = true;
	w->gravity = V2{ def->gravity.x, def->gravity.y };
	w->hitEventThreshold = def->hitEventThreshold;
	w->restitutionThreshold = def->restitutionThreshold;
	w->maxLinearSpeed = def->maximumLinearSpeed;
	w->maxContactPushSpeed = def->maxContactPushSpeed;
	w->contactHertz = def->contactHertz;
	w->contactDampingRatio = def->contactDampingRatio;
	w->enableSleep = def->enableSleep;
	w->enableContinuous = def->enableContinuous;
	if ( def->frictionCallback != nullptr || def->restitutionCallback != nullptr )
	{
		reportError( "b2CreateWorld: host friction/restitution callbacks cannot run inside the device step; default mixing is used" );
		w->hostCallbacks |= kHostMixing;
	}
	hw.state = kHostNewer;
	return b2WorldId{ (uint16_t)( index + 1 ), hw.generation };
}

void b2DestroyWorld( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	backendRelease( *hw );
	backendHostFree( hw->img );
	uint16_t generation = hw->generation;
	*hw = HostWorld();
	hw->generation = (uint16_t)( generation + 1 );
}

bool b2World_IsValid( b2WorldId id )
{
	return worldFromId( id ) != nullptr;
}

} // extern "C"
namespace f2d
{
static void prepareStep( HostWorld& hw )
{
	// capacity headroom for what the step itself may create (contacts from buffered moves)
	World* w = hw.img; // header is current in every sync state
	int needContacts = 4 * w->moveArray.count + 64;
	if ( w->contactIds.next + needContacts > hw.caps.contacts )
	{
		hostImage( hw );
		reserve( hw, 0, 0, needContacts, 0 );
		hw.state = kHostNewer;
	}
}

// b2World_Step for a world with b2CustomFilterFcn / b2PreSolveFcn registered. Both are host functions that must run on
// the calling thread during the call (SURVEY §8b threading), so the step is cut where the reference calls them:
//   pairs query | custom filter over the candidate pairs (broad_phase.c:267-278) | contact creation,
//   narrowphase | pre-solve over the touching contacts that ask for it (contact.c:504-517) | rest of the update,
// with only the few arrays the callbacks need crossing PCIe. Call order = the reference's with one worker: moved proxies
// in move-array order x tree-hit order; contacts in work-list order (colours 0..11, then the non-touching list).
template <class T> static void downloadArray( HostWorld& hw, const Arr<T>& a, int first, int count )
{
	if ( count > 0 )
		backendDownloadRange( hw, a.off + (uint64_t)first * sizeof( T ), (uint64_t)count * sizeof( T ) );
}
template <class T> static void uploadArray( HostWorld& hw, const Arr<T>& a, int first, int count )
{
	if ( count > 0 )
		backendUploadRange( hw, a.off + (uint64_t)first * sizeof( T ), (uint64_t)count * sizeof( T ) );
}
static void stepWithHostCallbacks( HostWorld& hw, float dt, int subSteps )
{
	World* w = hw.img;
	const int moveCountBefore = w->moveArray.count; // header is current in every sync state
	backendPhaseBegin( hw );
	backendPhase( hw, dt, subSteps, kPhasePairsQuery );
	if ( dt != 0.0f && moveCountBefore > 0 )
	{
		backendDownloadRange( hw, 0, sizeof( World ) );
		if ( w->step.retryContacts != 0 )
		{
			backendPhaseEnd( hw ); // stopped for lack of room, no callback has run: b2World_Step grows the image and repeats
			return;
		}
	}
	if ( hw.customFilterFcn != nullptr && dt != 0.0f && moveCountBefore > 0 )
	{
		const int moveCount = w->moveArray.count;
		const int total = w->step.orderedPairCount;
		if ( total > 0 )
		{
			const int pairCount = w->step.pairCount < w->movePairs.cap ? w->step.pairCount : w->movePairs.cap;
			downloadArray( hw, w->pairOffsets, 0, moveCount );
			downloadArray( hw, w->pairOrder, 0, total );
			downloadArray( hw, w->movePairs, 0, pairCount );
			const int32_t* offsets = ptr( w, w->pairOffsets );
			const int32_t* ordered = ptr( w, w->pairOrder );
			MovePair* pairs = ptr( w, w->movePairs );
			const Shape* shapes = ptr( w, w->shapes ); // generations never change on the device
			bool rejected = false;
			for ( int i = 0; i < moveCount; ++i )
			{
				const int begin = offsets[i], end = i + 1 < moveCount ? offsets[i + 1] : total;
				for ( int k = end - 1; k >= begin; --k ) // creation order is the reverse of hit order
				{
					MovePair& pair = pairs[ordered[k]];
					b2ShapeId idA = { pair.shapeA + 1, w->worldId, shapes[pair.shapeA].generation };
					b2ShapeId idB = { pair.shapeB + 1, w->worldId, shapes[pair.shapeB].generation };
					if ( hw.customFilterFcn( idA, idB, hw.customFilterContext ) == false )
					{
						pair.shapeA = kNull;
						rejected = true;
					}
				}
			}
			if ( rejected )
				uploadArray( hw, w->movePairs, 0, pairCount );
		}
	}
	backendPhase( hw, dt, subSteps, kPhasePairsCreate );
	backendPhase( hw, dt, subSteps, kPhaseCollideNarrow );
	if ( hw.preSolveFcn != nullptr && dt != 0.0f )
	{
		backendDownloadRange( hw, 0, sizeof( World ) );
		int pending = w->step.preSolveCount;
		if ( pending > w->stateList.cap )
			pending = w->stateList.cap;
		if ( pending > w->pairOrder.cap )
			pending = w->pairOrder.cap;
		if ( pending > 0 )
		{
			downloadArray( hw, w->stateList, 0, pending );
			downloadArray( hw, w->pairOrder, 0, pending );
			const int32_t* ids = ptr( w, w->stateList );
			int32_t* workIndexThenVerdict = ptr( w, w->pairOrder );
			std::vector<int> order( pending );
			int lo = ids[0], hi = ids[0];
			for ( int k = 0; k < pending; ++k )
			{
				order[k] = k;
				lo = ids[k] < lo ? ids[k] : lo;
				hi = ids[k] > hi ? ids[k] : hi;
			}
			std::sort( order.begin(), order.end(), [&]( int a, int b ) { return workIndexThenVerdict[a] < workIndexThenVerdict[b]; } );
			downloadArray( hw, w->contactSims, lo, hi - lo + 1 );
			const ContactSim* sims = ptr( w, w->contactSims );
			const Shape* shapes = ptr( w, w->shapes );
			std::vector<int32_t> verdicts( pending );
			for ( int k : order )
			{
				const ContactSim& sim = sims[ids[k]];
				Manifold manifold = unpackManifold( sim.manifold ); // what the reference hands out: the raw result of the manifold function
				unparkOldImpulses( manifold );
				b2ShapeId idA = { sim.shapeIdA + 1, w->worldId, shapes[sim.shapeIdA].generation };
				b2ShapeId idB = { sim.shapeIdB + 1, w->worldId, shapes[sim.shapeIdB].generation };
				verdicts[k] = hw.preSolveFcn( idA, idB, reinterpret_cast<b2Manifold*>( &manifold ), hw.preSolveContext ) ? 1 : 0;
			}
			for ( int k = 0; k < pending; ++k )
				workIndexThenVerdict[k] = verdicts[k];
			uploadArray( hw, w->pairOrder, 0, pending );
		}
	}
	backendPhase( hw, dt, subSteps, kPhaseCollideFinish );
	backendPhase( hw, dt, subSteps, kPhaseSolve );
	if ( dt == 0.0f )
	{
		backendPhase( hw, dt, subSteps, kPhaseFinalize );
		backendPhaseEnd( hw );
		return;
	}
	// The continuous pass consults the custom filter per candidate shape while it walks the trees and the pre-solve
	// callback per hit (solver.c:271-282, 366-379): like the world queries, that traversal runs on the host image, with the
	// same solveContinuous the device runs, for the (rare) fast bodies of this step only; the image crosses PCIe only
	// in steps that have any.
	struct Hooks
	{
		static b2ShapeId id( World* w, int shapeId ) { return b2ShapeId{ shapeId + 1, w->worldId, ptr( w, w->shapes )[shapeId].generation }; }
		static bool filter( void* context, int shapeId, int fastShapeId )
		{
			HostWorld* hw = static_cast<HostWorld*>( context );
			return hw->customFilterFcn( id( hw->img, shapeId ), id( hw->img, fastShapeId ), hw->customFilterContext );
		}
		static bool preSolve( void* context, int shapeId, int fastShapeId, Manifold* manifold )
		{
			HostWorld* hw = static_cast<HostWorld*>( context );
			return hw->preSolveFcn( id( hw->img, shapeId ), id( hw->img, fastShapeId ), reinterpret_cast<b2Manifold*>( manifold ),
									hw->preSolveContext );
		}
	};
	auto onHostImage = [&]( auto&& work ) {
		hw.wholeImageSync = true; // the work arrays of the step in flight too (parked bodies, enlarged-proxy bits ...), both ways
		backendDownload( hw );	  // synchronises with the phases queued
		g_hostContinuous.filter = hw.customFilterFcn != nullptr ? Hooks::filter : nullptr;
		g_hostContinuous.preSolve = hw.preSolveFcn != nullptr ? Hooks::preSolve : nullptr;
		g_hostContinuous.context = &hw;
		work();
		g_hostContinuous = HostContinuousHooks{};
		hw.state = kHostNewer;
		hw.wholeImageSync = true;
		backendPhaseBegin( hw ); // uploads
	};
	backendPhase( hw, dt, subSteps, kPhaseFinalizeBodies );
	backendDownloadRange( hw, 0, sizeof( World ) );
	if ( w->step.fastDeferredCount > 0 )
	{
		onHostImage( [&]() {
			const int n = w->step.fastDeferredCount;
			const int32_t* parkedFromEnd = ptr( w, w->bullets ) + w->bullets.cap - n;
			std::vector<int> order( parkedFromEnd, parkedFromEnd + n );
			std::sort( order.begin(), order.end() ); // awake order = the reference's order with one worker
			for ( int simIndex : order )
			{
				solveContinuous( w, simIndex );
				finalizeBodyTail( w, simIndex ); // (the island vote of a deferred body was cast by the device loop)
			}
		} );
	}
	backendPhase( hw, dt, subSteps, kPhaseFinalizeMoves );
	backendDownloadRange( hw, 0, sizeof( World ) );
	if ( w->step.bulletCount > 0 )
	{
		onHostImage( [&]() {
			const int32_t* bullets = ptr( w, w->bullets );
			std::vector<int> order( bullets, bullets + w->step.bulletCount );
			std::sort( order.begin(), order.end() );
			for ( int simIndex : order )
				solveContinuous( w, simIndex );
		} );
	}
	backendPhase( hw, dt, subSteps, kPhaseFinalizeEnd );
	backendPhaseEnd( hw );
}

// The step stops before its first structural edit when this step's new contacts do not fit the arrays of the image
// (f2d_step.h stepPairs; the reference's arrays grow on demand): grow and repeat. `stepped`: the first attempt has run
// already (f2dWorld_StepAsync) and only its verdict is looked at.
static void stepUntilItFits( HostWorld& hw, float timeStep, int subStepCount, bool stepped )
{
	for ( int attempt = 0;; ++attempt )
	{
		if ( stepped == false )
		{
			if ( hw.img->hostCallbacks & ( kHostCustomFilter | kHostPreSolve ) )
				stepWithHostCallbacks( hw, timeStep, subStepCount );
			else
				backendStep( hw, timeStep, subStepCount, true );
		}
		stepped = false;
		const int need = hw.img->step.retryContacts;
		if ( need == 0 || attempt == 4 )
			break;
		World* w = hostImage( hw );
		w->step.retryContacts = 0;
		w->error &= ~kErrRetry;
		reserve( hw, 0, 0, need + ( need >> 1 ), 0 );
		hw.img->step.retryContacts = 0;
		hw.img->error &= ~kErrRetry;
		hw.state = kHostNewer;
	}
}

} // namespace f2d
extern "C" {
void b2World_Step( b2WorldId worldId, float timeStep, int subStepCount )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	if ( backendAvailable() == false )
	{
		reportError( "b2World_Step: no CUDA device available - this library has no CPU fallback" );
		return;
	}
	if ( hw->img->error & kErrUnsupported )
	{
		checkWorldError( *hw, "b2World_Step refused" );
		return;
	}
	headerCurrent( *hw );
	prepareStep( *hw );
	stepUntilItFits( *hw, timeStep, subStepCount, false );
	hw->eventsFresh = ( hw->state != kDeviceNewer );
	checkWorldError( *hw, "b2World_Step" );
}

void f2dWorld_StepAsync( b2WorldId worldId, float timeStep, int subStepCount )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr || backendAvailable() == false )
		return;
	if ( hw->img->hostCallbacks & ( kHostCustomFilter | kHostPreSolve ) )
	{
		reportError( "f2dWorld_StepAsync: a world with host callbacks must be stepped with b2World_Step (the callbacks run on the calling thread)" );
		return;
	}
	headerCurrent( *hw ); // (of the step before this one: prepareStep reads it)
	prepareStep( *hw );
	backendStep( *hw, timeStep, subStepCount, false );
	hw->eventsFresh = false;
	hw->headerFresh = false;
	hw->asyncTimeStep = timeStep;
	hw->asyncSubSteps = subStepCount;
}
void f2dWorld_Synchronize( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	const bool pending = hw->headerFresh == false;
	backendSynchronize( *hw );
	hw->headerFresh = true;
	// an asynchronous step that stopped for lack of contact room is repeated here, on a larger image
	if ( pending && hw->img->step.retryContacts != 0 )
		stepUntilItFits( *hw, hw->asyncTimeStep, hw->asyncSubSteps, true );
	checkWorldError( *hw, "f2dWorld_Synchronize" );
}

} // extern "C"
// Event arrays: pointers into the host image, refreshed by range downloads (B2/src/world.c:1491-1555)
template <class T> static void refreshArray( HostWorld& hw, const Arr<T>& a )
{
	headerCurrent( hw );
	if ( hw.state == kDeviceNewer && a.count > 0 )
		backendDownloadRange( hw, a.off, (uint64_t)a.count * sizeof( T ) );
}
extern "C" {

b2BodyEvents b2World_GetBodyEvents( b2WorldId worldId )
{
	b2BodyEvents ev = { nullptr, 0 };
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return ev;
	World* w = hw->img;
	refreshArray( *hw, w->moveEvents );
	ev.moveEvents = reinterpret_cast<b2BodyMoveEvent*>( ptr( w, w->moveEvents ) );
	ev.moveCount = w->moveEvents.count;
	return ev;
}
b2SensorEvents b2World_GetSensorEvents( b2WorldId worldId )
{
	b2SensorEvents ev = { nullptr, nullptr, 0, 0 };
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return ev;
	World* w = hw->img;
	int endIndex = 1 - w->endEventArrayIndex;
	refreshArray( *hw, w->sensorBeginEvents );
	refreshArray( *hw, w->sensorEndEvents[endIndex] );
	ev.beginEvents = reinterpret_cast<b2SensorBeginTouchEvent*>( ptr( w, w->sensorBeginEvents ) );
	ev.endEvents = reinterpret_cast<b2SensorEndTouchEvent*>( ptr( w, w->sensorEndEvents[endIndex] ) );
	ev.beginCount = w->sensorBeginEvents.count;
	ev.endCount = w->sensorEndEvents[endIndex].count;
	return ev;
}
b2ContactEvents b2World_GetContactEvents( b2WorldId worldId )
{
	b2ContactEvents ev = { nullptr, nullptr, nullptr, 0, 0, 0 };
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return ev;
	World* w = hw->img;
	int endIndex = 1 - w->endEventArrayIndex; // previous buffer is exposed (world.c:1538)
	refreshArray( *hw, w->beginEvents );
	refreshArray( *hw, w->endEvents[endIndex] );
	refreshArray( *hw, w->hitEvents );
	ev.beginEvents = reinterpret_cast<b2ContactBeginTouchEvent*>( ptr( w, w->beginEvents ) );
	ev.endEvents = reinterpret_cast<b2ContactEndTouchEvent*>( ptr( w, w->endEvents[endIndex] ) );
	ev.hitEvents = reinterpret_cast<b2ContactHitEvent*>( ptr( w, w->hitEvents ) );
	ev.beginCount = w->beginEvents.count;
	ev.endCount = w->endEvents[endIndex].count;
	ev.hitCount = w->hitEvents.count;
	return ev;
}

void b2World_EnableSleeping( b2WorldId worldId, bool flag )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	World* w = mutableImage( *hw );
	if ( flag == w->enableSleep )
		return;
	w->enableSleep = flag;
	if ( flag == false )
	{
		// world.c:1710-1732: wake every sleeping set
		int setCount = w->sets.count;
		for ( int i = kFirstSleepingSet; i < setCount; ++i )
		{
			SolverSet& s = ptr( w, w->sets )[i];
			if ( s.setIndex != kNull && s.bodyCount > 0 )
			{
				reserve( *hw, 0, 0, 0, 0 );
				w = hw->img;
				wakeSolverSet( w, i );
			}
		}
	}
}
bool b2World_IsSleepingEnabled( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	return hw ? hw->img->enableSleep : false;
}
void b2World_EnableContinuous( b2WorldId worldId, bool flag )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw )
		mutableImage( *hw )->enableContinuous = flag;
}
bool b2World_IsContinuousEnabled( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	return hw ? hw->img->enableContinuous : false;
}
void b2World_EnableWarmStarting( b2WorldId worldId, bool flag )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw )
		mutableImage( *hw )->enableWarmStarting = flag;
}
void b2World_SetGravity( b2WorldId worldId, b2Vec2 gravity )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw )
		mutableImage( *hw )->gravity = V2{ gravity.x, gravity.y };
}
b2Vec2 b2World_GetGravity( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return b2Vec2{ 0, 0 };
	return b2Vec2{ hw->img->gravity.x, hw->img->gravity.y };
}
b2Counters b2World_GetCounters( b2WorldId worldId )
{
	b2Counters s;
	memset( &s, 0, sizeof( s ) );
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return s;
	World* w = hostImage( *hw );
	s.bodyCount = idCount( w->bodyIds );
	s.shapeCount = idCount( w->shapeIds );
	s.contactCount = idCount( w->contactIds );
	s.jointCount = idCount( w->jointIds );
	s.islandCount = idCount( w->islandIds );
	s.staticTreeHeight = treeHeight( w, w->trees[kStaticBody] );
	s.treeHeight = maxi( treeHeight( w, w->trees[kDynamicBody] ), treeHeight( w, w->trees[kKinematicBody] ) );
	s.byteCount = (int)w->imageBytes;
	s.taskCount = w->taskCount;
	for ( int i = 0; i < kColorCount; ++i )
		s.colorCounts[i] = w->colorContacts[i].count + w->colorJoints[i].count;
	return s;
}
int b2World_GetAwakeBodyCount( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return 0;
	headerCurrent( *hw );
	return hw->img->awakeBodies.count;
}

// ---- bodies ----------------------------------------------------------------------------------------------------
b2BodyId b2CreateBody( b2WorldId worldId, const b2BodyDef* def )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr || def->internalValue != kSecretCookie )
		return b2BodyId{ 0, 0, 0 };
	mutableImage( *hw );
	reserve( *hw, 2, 0, 0, 0 );
	World* w = hw->img;
	if ( w->locked )
		return b2BodyId{ 0, 0, 0 };
	BodyParams p;
	p.type = (int)def->type;
	p.position = V2{ def->position.x, def->position.y };
	p.rotation = Rot{ def->rotation.c, def->rotation.s };
	p.linearVelocity = V2{ def->linearVelocity.x, def->linearVelocity.y };
	p.angularVelocity = def->angularVelocity;
	p.linearDamping = def->linearDamping;
	p.angularDamping = def->angularDamping;
	p.gravityScale = def->gravityScale;
	p.sleepThreshold = def->sleepThreshold;
	p.name = def->name;
	p.userData = (uint64_t)(uintptr_t)def->userData;
	p.enableSleep = def->enableSleep;
	p.isAwake = def->isAwake;
	p.fixedRotation = def->fixedRotation;
	p.isBullet = def->isBullet;
	p.isEnabled = def->isEnabled;
	p.allowFastRotation = def->allowFastRotation;
	int bodyId = createBody( w, p );
	return b2BodyId{ bodyId + 1, w->worldId, ptr( w, w->bodies )[bodyId].generation };
}

static Body* bodyFromId( b2BodyId id, HostWorld** outWorld, bool forWrite )
{
	HostWorld* hw = worldFromIndex0( id.world0 );
	if ( hw == nullptr )
		return nullptr;
	World* w = forWrite ? mutableImage( *hw ) : hostImage( *hw );
	int index = id.index1 - 1;
	if ( index < 0 || index >= w->bodies.count )
		return nullptr;
	Body* b = ptr( w, w->bodies ) + index;
	if ( b->setIndex == kNull || b->id != index || b->generation != id.generation )
		return nullptr;
	if ( outWorld )
		*outWorld = hw;
	return b;
}

// Read access to what bodyImage keeps current (body record, simulation record, solver state)
static Body* bodyFromIdLight( b2BodyId id, HostWorld** outWorld )
{
	HostWorld* hw = worldFromIndex0( id.world0 );
	if ( hw == nullptr )
		return nullptr;
	World* w = bodyImage( *hw );
	int index = id.index1 - 1;
	if ( index < 0 || index >= w->bodies.count )
		return nullptr;
	Body* b = ptr( w, w->bodies ) + index;
	if ( b->setIndex == kNull || b->id != index || b->generation != id.generation )
		return nullptr;
	if ( outWorld )
		*outWorld = hw;
	return b;
}
// A per-frame mutator (force, impulse, velocity) of a body: when the body is awake - or need not be woken - the change
// is a field of its simulation record / solver state, recorded as a dirty range; waking a sleeping body is a
// structural edit and takes the full path (`wakes` = the call would wake the body if it slept).
static Body* bodyForLightEdit( b2BodyId id, HostWorld** outWorld, bool wakes )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromIdLight( id, &hw );
	if ( b == nullptr )
		return nullptr;
	if ( hw->state != kDeviceNewer || ( wakes && b->setIndex >= kFirstSleepingSet ) )
		return bodyFromId( id, outWorld, true );
	if ( outWorld )
		*outWorld = hw;
	return b;
}

bool b2Body_IsValid( b2BodyId id )
{
	return bodyFromId( id, nullptr, false ) != nullptr;
}
b2BodyType b2Body_GetType( b2BodyId id )
{
	Body* b = bodyFromId( id, nullptr, false );
	return b ? (b2BodyType)b->type : b2_staticBody;
}
b2Transform b2Body_GetTransform( b2BodyId id )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromIdLight( id, &hw );
	b2Transform t = { { 0, 0 }, { 1, 0 } };
	if ( b == nullptr )
		return t;
	const BodySim& s = ptr( hw->img, hw->img->sims )[b->id];
	t.p = b2Vec2{ s.transform.p.x, s.transform.p.y };
	t.q = b2Rot{ s.transform.q.c, s.transform.q.s };
	return t;
}
b2Vec2 b2Body_GetPosition( b2BodyId id )
{
	return b2Body_GetTransform( id ).p;
}
b2Rot b2Body_GetRotation( b2BodyId id )
{
	return b2Body_GetTransform( id ).q;
}
b2Vec2 b2Body_GetLinearVelocity( b2BodyId id )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromIdLight( id, &hw );
	if ( b == nullptr || b->setIndex != kAwakeSet )
		return b2Vec2{ 0, 0 };
	const BodyState& s = ptr( hw->img, hw->img->states )[b->localIndex];
	return b2Vec2{ s.v.x, s.v.y };
}
float b2Body_GetAngularVelocity( b2BodyId id )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromIdLight( id, &hw );
	if ( b == nullptr || b->setIndex != kAwakeSet )
		return 0.0f;
	return ptr( hw->img, hw->img->states )[b->localIndex].w;
}
// body.c:~1000 b2Body_SetLinearVelocity: wakes the body when the velocity is non-zero
void b2Body_SetLinearVelocity( b2BodyId id, b2Vec2 v )
{
	HostWorld* hw = nullptr;
	const bool wakes = v.x * v.x + v.y * v.y > 0.0f;
	Body* b = bodyForLightEdit( id, &hw, wakes );
	if ( b == nullptr || b->type == kStaticBody )
		return;
	World* w = hw->img;
	if ( wakes && b->setIndex >= kFirstSleepingSet )
		wakeBody( w, *b );
	if ( b->setIndex != kAwakeSet )
		return;
	ptr( w, w->states )[b->localIndex].v = V2{ v.x, v.y };
	touchRange( *hw, &ptr( w, w->states )[b->localIndex], sizeof( BodyState ) );
}
void b2Body_SetAngularVelocity( b2BodyId id, float wv )
{
	HostWorld* hw = nullptr;
	Body* b = bodyForLightEdit( id, &hw, wv != 0.0f );
	if ( b == nullptr || b->type == kStaticBody || b->fixedRotation )
		return;
	World* w = hw->img;
	if ( wv != 0.0f && b->setIndex >= kFirstSleepingSet )
		wakeBody( w, *b );
	if ( b->setIndex != kAwakeSet )
		return;
	ptr( w, w->states )[b->localIndex].w = wv;
	touchRange( *hw, &ptr( w, w->states )[b->localIndex], sizeof( BodyState ) );
}
float b2Body_GetMass( b2BodyId id )
{
	Body* b = bodyFromId( id, nullptr, false );
	return b ? b->mass : 0.0f;
}
float b2Body_GetRotationalInertia( b2BodyId id )
{
	Body* b = bodyFromId( id, nullptr, false );
	return b ? b->inertia : 0.0f;
}
b2Vec2 b2Body_GetLocalCenterOfMass( b2BodyId id )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( id, &hw, false );
	if ( b == nullptr )
		return b2Vec2{ 0, 0 };
	const BodySim& s = ptr( hw->img, hw->img->sims )[b->id];
	return b2Vec2{ s.localCenter.x, s.localCenter.y };
}
b2Vec2 b2Body_GetWorldCenterOfMass( b2BodyId id )
{
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( id, &hw, false );
	if ( b == nullptr )
		return b2Vec2{ 0, 0 };
	const BodySim& s = ptr( hw->img, hw->img->sims )[b->id];
	return b2Vec2{ s.center.x, s.center.y };
}
bool b2Body_IsAwake( b2BodyId id )
{
	Body* b = bodyFromId( id, nullptr, false );
	return b ? b->setIndex == kAwakeSet : false;
}
int b2Body_GetShapeCount( b2BodyId id )
{
	Body* b = bodyFromId( id, nullptr, false );
	return b ? b->shapeCount : 0;
}
int b2Body_GetContactCapacity( b2BodyId id )
{
	Body* b = bodyFromId( id, nullptr, false );
	return b ? b->contactCount : 0;
}

// ---- shapes ----------------------------------------------------------------------------------------------------
static b2ShapeId createShapeCommon( b2BodyId bodyId, const b2ShapeDef* def, const void* geometry, int type )
{
	if ( def->internalValue != kSecretCookie )
		return b2ShapeId{ 0, 0, 0 };
	HostWorld* hw = nullptr;
	Body* b = bodyFromId( bodyId, &hw, true );
	if ( b == nullptr )
		return b2ShapeId{ 0, 0, 0 };
	int bodyIndex = b->id;
	reserve( *hw, 0, 2, 0, 0, def->isSensor ? 1 : 0 );
	World* w = hw->img;
	if ( w->locked )
		return b2ShapeId{ 0, 0, 0 };
	ShapeParams p;
	p.userData = (uint64_t)(uintptr_t)def->userData;
	p.friction = def->material.friction;
	p.restitution = def->material.restitution;
	p.rollingResistance = def->material.rollingResistance;
	p.tangentSpeed = def->material.tangentSpeed;
	p.userMaterialId = def->material.userMaterialId;
	p.customColor = def->material.customColor;
	p.density = def->density;
	p.filter = Filter{ def->filter.categoryBits, def->filter.maskBits, def->filter.groupIndex };
	p.isSensor = def->isSensor;
	p.enableSensorEvents = def->enableSensorEvents;
	p.enableContactEvents = def->enableContactEvents;
	p.enableHitEvents = def->enableHitEvents;
	p.enablePreSolveEvents = def->enablePreSolveEvents;
	p.invokeContactCreation = def->invokeContactCreation;
	p.updateBodyMass = def->updateBodyMass;
	int before = w->contactEventCapable + w->hitEventCapable;
	int shapeId = createShape( w, bodyIndex, p, geometry, type );
	if ( w->contactEventCapable + w->hitEventCapable != before )
	{
		reserve( *hw, 0, 0, 0, 0 ); // event arrays grow with the first event-enabled shape
		w = hw->img;
	}
	return b2ShapeId{ shapeId + 1, w->worldId, ptr( w, w->shapes )[shapeId].generation };
}
b2ShapeId b2CreateCircleShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Circle* circle )
{
	Circle c = { { circle->center.x, circle->center.y }, circle->radius };
	return createShapeCommon( bodyId, def, &c, kCircle );
}
b2ShapeId b2CreateCapsuleShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Capsule* capsule )
{
	V2 c1 = { capsule->center1.x, capsule->center1.y }, c2 = { capsule->center2.x, capsule->center2.y };
	float lengthSqr = distanceSq( c1, c2 );
	if ( lengthSqr <= kLinearSlop * kLinearSlop ) // shape.c:198-208
	{
		Circle c = { lerp( c1, c2, 0.5f ), capsule->radius };
		return createShapeCommon( bodyId, def, &c, kCircle );
	}
	Capsule c = { c1, c2, capsule->radius };
	return createShapeCommon( bodyId, def, &c, kCapsule );
}
b2ShapeId b2CreatePolygonShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Polygon* polygon )
{
	static_assert( sizeof( Poly ) == sizeof( b2Polygon ), "polygon layout" );
	Poly p;
	memcpy( &p, polygon, sizeof( p ) );
	return createShapeCommon( bodyId, def, &p, kPolygon );
}
b2ShapeId b2CreateSegmentShape( b2BodyId bodyId, const b2ShapeDef* def, const b2Segment* segment )
{
	V2 p1 = { segment->point1.x, segment->point1.y }, p2 = { segment->point2.x, segment->point2.y };
	float lengthSqr = distanceSq( p1, p2 );
	if ( lengthSqr <= kLinearSlop * kLinearSlop ) // shape.c:215-223
		return b2ShapeId{ 0, 0, 0 };
	Segment s = { p1, p2 };
	return createShapeCommon( bodyId, def, &s, kSegment );
}
static Shape* shapeFromId( b2ShapeId id, HostWorld** outWorld )
{
	HostWorld* hw = worldFromIndex0( id.world0 );
	if ( hw == nullptr )
		return nullptr;
	World* w = hostImage( *hw );
	int index = id.index1 - 1;
	if ( index < 0 || index >= w->shapes.count )
		return nullptr;
	Shape* s = ptr( w, w->shapes ) + index;
	if ( s->id != index || s->generation != id.generation )
		return nullptr;
	if ( outWorld )
		*outWorld = hw;
	return s;
}
bool b2Shape_IsValid( b2ShapeId id )
{
	return shapeFromId( id, nullptr ) != nullptr;
}
b2BodyId b2Shape_GetBody( b2ShapeId id )
{
	HostWorld* hw = nullptr;
	Shape* s = shapeFromId( id, &hw );
	if ( s == nullptr )
		return b2BodyId{ 0, 0, 0 };
	World* w = hw->img;
	return b2BodyId{ s->bodyId + 1, w->worldId, ptr( w, w->bodies )[s->bodyId].generation };
}
b2AABB b2Shape_GetAABB( b2ShapeId id )
{
	Shape* s = shapeFromId( id, nullptr );
	b2AABB a = { { 0, 0 }, { 0, 0 } };
	if ( s )
	{
		a.lowerBound = b2Vec2{ s->aabb.lo.x, s->aabb.lo.y };
		a.upperBound = b2Vec2{ s->aabb.hi.x, s->aabb.hi.y };
	}
	return a;
}

// ---- joints ----------------------------------------------------------------------------------------------------
b2JointId b2CreateRevoluteJoint( b2WorldId worldId, const b2RevoluteJointDef* def )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr || def->internalValue != kSecretCookie )
		return b2JointId{ 0, 0, 0 };
	mutableImage( *hw );
	reserve( *hw, 0, 0, 0, 2 );
	World* w = hw->img;
	if ( w->locked )
		return b2JointId{ 0, 0, 0 };
	RevoluteParams p;
	p.bodyIdA = def->bodyIdA.index1 - 1;
	p.bodyIdB = def->bodyIdB.index1 - 1;
	p.localAnchorA = V2{ def->localAnchorA.x, def->localAnchorA.y };
	p.localAnchorB = V2{ def->localAnchorB.x, def->localAnchorB.y };
	p.referenceAngle = def->referenceAngle;
	p.targetAngle = def->targetAngle;
	p.enableSpring = def->enableSpring;
	p.hertz = def->hertz;
	p.dampingRatio = def->dampingRatio;
	p.enableLimit = def->enableLimit;
	p.lowerAngle = def->lowerAngle;
	p.upperAngle = def->upperAngle;
	p.enableMotor = def->enableMotor;
	p.maxMotorTorque = def->maxMotorTorque;
	p.motorSpeed = def->motorSpeed;
	p.drawSize = def->drawSize;
	p.collideConnected = def->collideConnected;
	p.userData = (uint64_t)(uintptr_t)def->userData;
	int jointId = createRevoluteJoint( w, p );
	return b2JointId{ jointId + 1, w->worldId, ptr( w, w->joints )[jointId].generation };
}
bool b2Joint_IsValid( b2JointId id )
{
	HostWorld* hw = worldFromIndex0( id.world0 );
	if ( hw == nullptr )
		return false;
	World* w = hostImage( *hw );
	int index = id.index1 - 1;
	if ( index < 0 || index >= w->joints.count )
		return false;
	const Joint& j = ptr( w, w->joints )[index];
	return j.jointId == index && j.generation == id.generation;
}

#include "f2d_capi_ext.inl"

// ---- diagnostics -----------------------------------------------------------------------------------------------
int f2dHasDevice( void )
{
	return backendAvailable() ? 1 : 0;
}
const char* f2dGetLastError( void )
{
	return g_lastError.c_str();
}
void f2dClearLastError( void )
{
	g_lastError.clear();
}
void f2dWorld_SetLaunchMode( b2WorldId worldId, int mode )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw )
		hw->launchMode = mode;
}
uint32_t f2dWorld_GetErrorFlags( b2WorldId worldId )
{
	HostWorld* hw = worldFromId( worldId );
	return hw ? hw->img->error : 0;
}
long long f2dWorld_GetKernelLaunchCount( void )
{
	return g_launchCount;
}
void f2dWorld_GetLastStepTimes( b2WorldId worldId, float* out5 )
{
	HostWorld* hw = worldFromId( worldId );
	for ( int i = 0; i < 5; ++i )
		out5[i] = 0.0f;
	if ( hw )
		backendStepTimes( *hw, out5 );
}
// What the last step did: [islandPath, activeColorCount, awakeContactCount, awakeBodyCount, maxIslandContacts,
// maxIslandBodies, awakeIslandCount, mergeCount, splitBodies, splitComponents, splitContacts, splitJoints]
int f2dWorld_GetStepInfo( b2WorldId worldId, int* out, int cap )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return 0;
	const World* w = hw->img; // the header is current after every step
	int v[12] = { w->step.islandPath,		 w->step.activeColorCount, w->step.awakeContactCount, w->step.awakeBodyCount,
				  w->step.maxIslandContacts, w->step.maxIslandBodies,  w->awakeIslands.count,	  w->step.mergeCount,
				  w->step.splitBodies,		 w->step.splitComponents,  w->step.splitContacts,	  w->step.splitJoints };
	for ( int i = 0; i < 12 && i < cap; ++i )
		out[i] = v[i];
	return 12;
}
// Measurement aid: the narrowphase of a world with several shape types with / without its work list binned by pair class
void f2dWorld_EnablePairClassBinning( b2WorldId worldId, bool flag )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	mutableImage( *hw )->pairClassBinningOff = flag ? 0 : 1;
}
// In-kernel profile (nanoseconds per f2d::ProfSlot accumulated by rank 0 since it was enabled / last read)
void f2dWorld_EnableProfile( b2WorldId worldId, bool flag )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return;
	World* w = mutableImage( *hw );
	w->profEnabled = flag ? 1 : 0;
	memset( w->prof, 0, sizeof( w->prof ) );
}
int f2dWorld_ReadProfile( b2WorldId worldId, unsigned long long* out, int cap )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return 0;
	for ( int i = 0; i < kProfSlots && i < cap; ++i )
		out[i] = hw->img->prof[i];
	return kProfSlots;
}
// b2World_GetProfile (box2d.h:169, types.h:466-490): milliseconds per phase, averaged over the steps since the previous
// call. The numbers come from the in-kernel phase marks (rank 0 reads %globaltimer between the phases of the step), which
// the first call switches on: like the reference's profile they are always available afterwards, for a few clock reads
// per step. Phases the device step does not have as such (prepareStages, sensors) report 0.
b2Profile b2World_GetProfile( b2WorldId worldId )
{
	b2Profile p;
	memset( &p, 0, sizeof( p ) );
	HostWorld* hw = worldFromId( worldId );
	if ( hw == nullptr )
		return p;
	World* w = hw->img; // the header is current in every sync state
	if ( w->profEnabled == 0 )
	{
		w = mutableImage( *hw );
		w->profEnabled = 1;
		memset( w->prof, 0, sizeof( w->prof ) );
		memset( hw->profSeen, 0, sizeof( hw->profSeen ) );
		hw->profStepSeen = w->stepIndex;
		return p;
	}
	const uint64_t steps = w->stepIndex - hw->profStepSeen;
	if ( steps == 0 )
		return p;
	auto ms = [&]( int slot ) {
		const uint64_t ns = w->prof[slot] - hw->profSeen[slot];
		return (float)( (double)ns * 1e-6 / (double)steps );
	};
	p.pairs = ms( pfBegin ) + ms( pfPairQuery ) + ms( pfPairCreate );
	p.collide = ms( pfTreeRebuild ) + ms( pfNarrow ) + ms( pfStatePass );
	p.mergeIslands = ms( pfSolveSetup );
	p.prepareConstraints = ms( pfPrepare );
	p.integrateVelocities = ms( pfIntegrateVel );
	p.warmStart = ms( pfWarmStart );
	p.solveImpulses = ms( pfSolve );
	p.integratePositions = ms( pfIntegratePos );
	p.relaxImpulses = ms( pfRelax );
	p.applyRestitution = ms( pfRestitution );
	p.storeImpulses = ms( pfStore );
	p.splitIslands = ms( pfSplitJoin ) + ms( pfSplitApply );
	p.solveConstraints = p.prepareConstraints + p.integrateVelocities + p.warmStart + p.solveImpulses + p.integratePositions +
						 p.relaxImpulses + p.applyRestitution + p.storeImpulses + p.splitIslands;
	p.transforms = ms( pfFinalizeBodies );
	p.hitEvents = ms( pfHitEvents );
	p.refit = ms( pfEnlarge );
	p.bullets = ms( pfBullets );
	p.sleepIslands = ms( pfSleep );
	p.solve = p.mergeIslands + p.solveConstraints + p.transforms + p.hitEvents + p.refit + p.bullets + p.sleepIslands;
	p.step = p.pairs + p.collide + p.solve + ms( pfEnd );
	for ( int i = 0; i < kProfSlots; ++i )
		hw->profSeen[i] = w->prof[i];
	hw->profStepSeen = w->stepIndex;
	return p;
}

void f2dWorld_EnablePhaseTiming( b2WorldId worldId, bool flag )
{
	HostWorld* hw = worldFromId( worldId );
	if ( hw )
		backendEnableTiming( *hw, flag );
}

// ---- introspection (same records as oracle/tap.c) ----------------------------------------------------------------
int f2dDebug_AwakeOrder( b2WorldId id, int* bodyIds, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = w->awakeBodies.count;
	for ( int i = 0; i < n && i < cap; ++i )
		bodyIds[i] = ptr( w, w->awakeBodies )[i];
	return n;
}
int f2dDebug_MoveArray( b2WorldId id, int* keys, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = w->moveArray.count;
	for ( int i = 0; i < n && i < cap; ++i )
		keys[i] = ptr( w, w->moveArray )[i];
	return n;
}
int f2dDebug_Bodies( b2WorldId id, f2dBodyRecord* out, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = 0;
	for ( int i = 0; i < w->bodies.count; ++i )
	{
		const Body& b = ptr( w, w->bodies )[i];
		if ( b.setIndex == kNull || b.id != i )
			continue;
		if ( n < cap )
		{
			f2dBodyRecord* r = out + n;
			memset( r, 0, sizeof( *r ) );
			const BodySim& sim = ptr( w, w->sims )[i];
			r->id = i;
			r->setIndex = b.setIndex;
			r->localIndex = b.localIndex;
			r->islandId = b.islandId;
			r->islandPrev = b.islandPrev;
			r->islandNext = b.islandNext;
			r->type = b.type;
			r->headContactKey = b.headContactKey;
			r->contactCount = b.contactCount;
			r->headShapeId = b.headShapeId;
			r->flags = ( sim.isFast ? 1 : 0 ) | ( sim.isBullet ? 2 : 0 ) | ( b.isSpeedCapped ? 4 : 0 ) | ( sim.enlargeAABB ? 8 : 0 );
			r->px = sim.transform.p.x;
			r->py = sim.transform.p.y;
			r->qc = sim.transform.q.c;
			r->qs = sim.transform.q.s;
			r->cx = sim.center.x;
			r->cy = sim.center.y;
			r->c0x = sim.center0.x;
			r->c0y = sim.center0.y;
			r->q0c = sim.rotation0.c;
			r->q0s = sim.rotation0.s;
			if ( b.setIndex == kAwakeSet )
			{
				const BodyState& s = ptr( w, w->states )[b.localIndex];
				r->vx = s.v.x;
				r->vy = s.v.y;
				r->w = s.w;
			}
			r->sleepTime = b.sleepTime;
			r->invMass = sim.invMass;
			r->invInertia = sim.invInertia;
			r->minExtent = sim.minExtent;
			r->maxExtent = sim.maxExtent;
			r->lcx = sim.localCenter.x;
			r->lcy = sim.localCenter.y;
		}
		n += 1;
	}
	return n;
}
int f2dDebug_Contacts( b2WorldId id, f2dContactRecord* out, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = 0;
	for ( int i = 0; i < w->contacts.count; ++i )
	{
		const Contact& c = ptr( w, w->contacts )[i];
		if ( c.contactId != i || c.setIndex == kNull )
			continue;
		if ( n < cap )
		{
			f2dContactRecord* r = out + n;
			memset( r, 0, sizeof( *r ) );
			const ContactSim& s = ptr( w, w->contactSims )[i];
			const Manifold manifold = unpackManifold( s.manifold );
			r->id = i;
			r->shapeIdA = c.shapeIdA;
			r->shapeIdB = c.shapeIdB;
			r->setIndex = c.setIndex;
			r->colorIndex = c.colorIndex;
			r->localIndex = c.localIndex;
			r->flags = (int)( c.flags & ~kContactMarked );
			r->simFlags = (int)s.simFlags;
			r->pointCount = manifold.pointCount;
			r->islandId = c.islandId;
			r->islandPrev = c.islandPrev;
			r->islandNext = c.islandNext;
			r->prevKeyA = c.edges[0].prevKey;
			r->nextKeyA = c.edges[0].nextKey;
			r->prevKeyB = c.edges[1].prevKey;
			r->nextKeyB = c.edges[1].nextKey;
			r->bodySimIndexA = s.bodySimIndexA;
			r->bodySimIndexB = s.bodySimIndexB;
			r->nx = manifold.normal.x;
			r->ny = manifold.normal.y;
			for ( int k = 0; k < manifold.pointCount && k < 2; ++k )
			{
				const ManifoldPoint& mp = manifold.points[k];
				if ( k == 0 )
					r->id0 = mp.id;
				else
					r->id1 = mp.id;
				r->sep[k] = mp.separation;
				r->ni[k] = mp.normalImpulse;
				r->ti[k] = mp.tangentImpulse;
				r->tni[k] = mp.totalNormalImpulse;
				r->nv[k] = mp.normalVelocity;
				r->ax[k] = mp.anchorA.x;
				r->ay[k] = mp.anchorA.y;
				r->bx[k] = mp.anchorB.x;
				r->by[k] = mp.anchorB.y;
				r->px[k] = mp.point.x;
				r->py[k] = mp.point.y;
			}
			r->friction = s.friction;
			r->restitution = s.restitution;
			r->rollingImpulse = manifold.rollingImpulse;
		}
		n += 1;
	}
	return n;
}
int f2dDebug_Islands( b2WorldId id, f2dIslandRecord* out, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = 0;
	for ( int i = 0; i < w->islands.count; ++i )
	{
		const Island& s = ptr( w, w->islands )[i];
		if ( s.islandId != i || s.setIndex == kNull )
			continue;
		if ( n < cap )
		{
			f2dIslandRecord* r = out + n;
			r->id = i;
			r->setIndex = s.setIndex;
			r->localIndex = s.localIndex;
			r->headBody = s.headBody;
			r->tailBody = s.tailBody;
			r->bodyCount = s.bodyCount;
			r->headContact = s.headContact;
			r->tailContact = s.tailContact;
			r->contactCount = s.contactCount;
			r->headJoint = s.headJoint;
			r->tailJoint = s.tailJoint;
			r->jointCount = s.jointCount;
			r->parentIsland = s.parentIsland;
			r->constraintRemoveCount = s.constraintRemoveCount;
		}
		n += 1;
	}
	return n;
}
int f2dDebug_Shapes( b2WorldId id, f2dShapeRecord* out, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = 0;
	for ( int i = 0; i < w->shapes.count; ++i )
	{
		const Shape& s = ptr( w, w->shapes )[i];
		if ( s.id != i )
			continue;
		if ( n < cap )
		{
			f2dShapeRecord* r = out + n;
			r->id = i;
			r->bodyId = s.bodyId;
			r->proxyKey = s.proxyKey;
			r->type = s.type;
			r->enlarged = s.enlargedAABB;
			r->aabb[0] = s.aabb.lo.x;
			r->aabb[1] = s.aabb.lo.y;
			r->aabb[2] = s.aabb.hi.x;
			r->aabb[3] = s.aabb.hi.y;
			r->fat[0] = s.fatAABB.lo.x;
			r->fat[1] = s.fatAABB.lo.y;
			r->fat[2] = s.fatAABB.hi.x;
			r->fat[3] = s.fatAABB.hi.y;
		}
		n += 1;
	}
	return n;
}
int f2dDebug_Tree( b2WorldId id, int treeType, f2dTreeLeafRecord* out, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	const Tree& tree = w->trees[treeType];
	if ( tree.root == kNull || tree.nodeCount == 0 )
		return 0;
	const TreeNode* nodes = ptr( w, tree.nodes );
	int n = 0;
	std::vector<int> stack, depth, enl;
	stack.push_back( tree.root );
	depth.push_back( 0 );
	enl.push_back( 0 );
	while ( stack.empty() == false )
	{
		int ni = stack.back(), d = depth.back(), e = enl.back();
		stack.pop_back();
		depth.pop_back();
		enl.pop_back();
		const TreeNode& node = nodes[ni];
		if ( node.flags & kNodeLeaf )
		{
			if ( n < cap )
			{
				out[n].proxyId = ni;
				out[n].depth = d;
				out[n].enlargedAncestors = e;
				out[n].box[0] = node.box.lo.x;
				out[n].box[1] = node.box.lo.y;
				out[n].box[2] = node.box.hi.x;
				out[n].box[3] = node.box.hi.y;
			}
			n += 1;
		}
		else
		{
			int e2 = e + ( ( node.flags & kNodeEnlarged ) ? 1 : 0 );
			stack.push_back( node.child2 );
			depth.push_back( d + 1 );
			enl.push_back( e2 );
			stack.push_back( node.child1 );
			depth.push_back( d + 1 );
			enl.push_back( e2 );
		}
	}
	return n;
}
int f2dDebug_Joints( b2WorldId id, f2dJointRecord* out, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	int n = 0;
	for ( int i = 0; i < w->joints.count; ++i )
	{
		const Joint& j = ptr( w, w->joints )[i];
		if ( j.jointId != i || j.setIndex == kNull )
			continue;
		if ( n < cap )
		{
			f2dJointRecord* r = out + n;
			memset( r, 0, sizeof( *r ) );
			const JointSim& s = ptr( w, w->jointSims )[i];
			r->id = i;
			r->type = j.type;
			r->setIndex = j.setIndex;
			r->colorIndex = j.colorIndex;
			r->localIndex = j.localIndex;
			r->bodyIdA = j.edges[0].bodyId;
			r->bodyIdB = j.edges[1].bodyId;
			r->islandId = j.islandId;
			switch ( j.type )
			{
				case kRevoluteJoint:
					r->impulse[0] = s.revolute.linearImpulse.x;
					r->impulse[1] = s.revolute.linearImpulse.y;
					r->impulse[2] = s.revolute.springImpulse;
					r->impulse[3] = s.revolute.motorImpulse;
					r->impulse[4] = s.revolute.lowerImpulse;
					r->impulse[5] = s.revolute.upperImpulse;
					break;
				case kDistanceJoint:
					r->impulse[0] = s.distance.impulse;
					r->impulse[1] = s.distance.lowerImpulse;
					r->impulse[2] = s.distance.upperImpulse;
					r->impulse[3] = s.distance.motorImpulse;
					break;
				case kMotorJoint:
					r->impulse[0] = s.motor.linearImpulse.x;
					r->impulse[1] = s.motor.linearImpulse.y;
					r->impulse[2] = s.motor.angularImpulse;
					break;
				case kMouseJoint:
					r->impulse[0] = s.mouse.linearImpulse.x;
					r->impulse[1] = s.mouse.linearImpulse.y;
					r->impulse[2] = s.mouse.angularImpulse;
					break;
				case kPrismaticJoint:
					r->impulse[0] = s.prismatic.impulse.x;
					r->impulse[1] = s.prismatic.impulse.y;
					r->impulse[2] = s.prismatic.springImpulse;
					r->impulse[3] = s.prismatic.motorImpulse;
					r->impulse[4] = s.prismatic.lowerImpulse;
					r->impulse[5] = s.prismatic.upperImpulse;
					break;
				case kWeldJoint:
					r->impulse[0] = s.weld.linearImpulse.x;
					r->impulse[1] = s.weld.linearImpulse.y;
					r->impulse[2] = s.weld.angularImpulse;
					break;
				case kWheelJoint:
					r->impulse[0] = s.wheel.perpImpulse;
					r->impulse[1] = s.wheel.motorImpulse;
					r->impulse[2] = s.wheel.springImpulse;
					r->impulse[3] = s.wheel.lowerImpulse;
					r->impulse[4] = s.wheel.upperImpulse;
					break;
				default:
					break;
			}
		}
		n += 1;
	}
	return n;
}
void f2dDebug_ColorCounts( b2WorldId id, int* contactCounts, int* jointCounts )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return;
	World* w = hostImage( *hw );
	for ( int i = 0; i < kColorCount; ++i )
	{
		contactCounts[i] = w->colorContacts[i].count;
		jointCounts[i] = w->colorJoints[i].count;
	}
}
static int copyIdList( HostWorld* hw, const Arr<int32_t>& a, int* out, int cap )
{
	World* w = hw->img;
	for ( int i = 0; i < a.count && i < cap; ++i )
		out[i] = ptr( w, a )[i];
	return a.count;
}
int f2dDebug_ColorContacts( b2WorldId id, int colorIndex, int* contactIds, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	return copyIdList( hw, w->colorContacts[colorIndex], contactIds, cap );
}
int f2dDebug_AwakeContacts( b2WorldId id, int* contactIds, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	return copyIdList( hw, w->awakeContacts, contactIds, cap );
}
int f2dDebug_AwakeIslands( b2WorldId id, int* islandIds, int cap )
{
	HostWorld* hw = worldFromId( id );
	if ( hw == nullptr )
		return 0;
	World* w = hostImage( *hw );
	return copyIdList( hw, w->awakeIslands, islandIds, cap );
}

} // extern "C"
