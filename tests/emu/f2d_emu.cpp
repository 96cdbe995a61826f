// TEST INFRASTRUCTURE — host emulation of the device step (never shipped, never loaded by the product).
//
// Compiles the exact team-parallel step code of forge2d_b200/csrc (f2d_step.h) for ONE host thread (SerialTeam)
// behind the same C ABI, so the order-defining logic can be debugged and parity-checked against the compiled
// reference (oracle/_ref) in containers without a GPU. The product library (f2d_cuda.cu) contains no such path:
// its b2World_Step only launches CUDA kernels.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../forge2d_b200/csrc/f2d_capi.inl"

namespace f2d
{
static void* backendHostAlloc( size_t bytes ) { return malloc( bytes ); }
static void backendHostFree( void* p ) { free( p ); }
static bool backendAvailable() { return true; }
static void backendStep( HostWorld& hw, float dt, int subSteps, bool )
{
	SerialTeam team;
	stepWorld( hw.img, team, dt, subSteps );
	hw.state = kInSync;
}
static void backendPhaseBegin( HostWorld& ) {}
static void backendPhase( HostWorld& hw, float dt, int subSteps, int phase )
{
	SerialTeam team;
	stepWorldPhase( hw.img, team, phase, dt, subSteps );
}
static void backendUploadRange( HostWorld&, uint64_t, uint64_t ) {}
static void backendPhaseEnd( HostWorld& hw ) { hw.state = kInSync; }
static void backendSynchronize( HostWorld& ) {}
static void backendDownload( HostWorld& ) {}
static void backendDownloadRange( HostWorld&, uint64_t, uint64_t ) {}
static void backendDownloadRanges( HostWorld&, std::initializer_list<ByteRange> ) {}
static void backendRelease( HostWorld& ) {}
static void backendStepTimes( HostWorld&, float* ) {}
static void backendEnableTiming( HostWorld&, bool ) {}
} // namespace f2d

// batch emulation: independent copies of the template image stepped one after another
struct f2dBatch
{
	std::vector<f2d::World*> worlds;
	uint64_t bytes;
};
extern "C" {
f2dBatch* f2dBatch_Create( b2WorldId templateWorld, int count )
{
	f2d::HostWorld* hw = f2d::worldFromId( templateWorld );
	if ( hw == nullptr )
		return nullptr;
	f2d::prepareStep( *hw );
	f2dBatch* b = new f2dBatch();
	b->bytes = hw->img->imageBytes;
	for ( int i = 0; i < count; ++i )
	{
		f2d::World* w = (f2d::World*)malloc( b->bytes );
		memcpy( w, hw->img, b->bytes );
		b->worlds.push_back( w );
	}
	return b;
}
void f2dBatch_Destroy( f2dBatch* b )
{
	for ( f2d::World* w : b->worlds )
		free( w );
	delete b;
}
void f2dBatch_Step( f2dBatch* b, float dt, int sub )
{
	f2d::SerialTeam team;
	for ( f2d::World* w : b->worlds )
		f2d::stepWorld( w, team, dt, sub );
}
void f2dBatch_StepN( f2dBatch* b, float dt, int sub, int steps )
{
	for ( int i = 0; i < steps; ++i )
		f2dBatch_Step( b, dt, sub );
}
void f2dBatch_Synchronize( f2dBatch* ) {}
int f2dBatch_GetWorldCount( f2dBatch* b ) { return (int)b->worlds.size(); }
int f2dBatch_GetBodyEvents( f2dBatch* b, b2BodyMoveEvent* out, int maxBodies, int* counts )
{
	int total = 0;
	for ( size_t i = 0; i < b->worlds.size(); ++i )
	{
		f2d::World* w = b->worlds[i];
		int n = w->moveEvents.count < maxBodies ? w->moveEvents.count : maxBodies;
		memcpy( out + i * maxBodies, f2d::ptr( w, w->moveEvents ), n * sizeof( b2BodyMoveEvent ) );
		counts[i] = n;
		total += n;
	}
	return total;
}
void f2dBatch_DownloadWorld( f2dBatch* b, int index, b2WorldId into )
{
	f2d::HostWorld* hw = f2d::worldFromId( into );
	if ( hw == nullptr || index < 0 || index >= (int)b->worlds.size() )
		return;
	f2d::World* src = b->worlds[index];
	uint16_t worldId = hw->img->worldId, generation = hw->img->generation;
	free( hw->img );
	hw->img = (f2d::World*)malloc( b->bytes );
	memcpy( hw->img, src, b->bytes );
	hw->img->worldId = worldId;
	hw->img->generation = generation;
	hw->state = f2d::kInSync;
}
uint32_t f2dBatch_GetErrorFlags( f2dBatch* b )
{
	uint32_t e = 0;
	for ( f2d::World* w : b->worlds )
		e |= w->error;
	return e;
}
b2Hull b2ComputeHull( const b2Vec2* points, int count );
}
