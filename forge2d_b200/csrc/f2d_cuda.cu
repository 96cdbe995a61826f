// forge2d_b200 — CUDA backend (sm_100a): kernels that run the team-parallel step phases of f2d_step.h on the GPU and
// the host glue that keeps the device-resident world image in sync with the C ABI.
//
// Execution mapping (B200: 148 SMs, 227 KB smem/SM, 126 MB L2):
//   * batch of worlds  -> one thread block per world (CtaTeam, `__syncthreads` between phases), grid = #worlds;
//     worlds are independent, so there is no inter-block traffic at all and shards map 1:1 onto GPUs.
//   * one small world  -> a single 1024-thread block (barrier cost ~tens of ns instead of a multi-us grid barrier;
//     the whole working set of a bench2d-sized world lives in that SM's L1/L2).
//   * one large world  -> a cooperative grid of one block per SM (GridTeam, grid-wide barrier between phases).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (no FMA contraction: bit parity with the
// reference's SSE2 arithmetic).
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stddef.h>

#include "f2d_capi.inl"
#include "f2d_launch.h"

namespace f2d
{

// ------------------------------------------------------------------------------------------------ host glue
static int g_deviceState = -1; // -1 unknown, 0 none, 1 ok
static int g_smCount = 0;

static bool backendAvailable()
{
	if ( g_deviceState < 0 )
	{
		int n = 0;
		cudaError_t e = cudaGetDeviceCount( &n );
		if ( e != cudaSuccess || n == 0 )
		{
			cudaGetLastError();
			g_deviceState = 0;
		}
		else
		{
			int dev = 0;
			cudaGetDevice( &dev );
			cudaDeviceGetAttribute( &g_smCount, cudaDevAttrMultiProcessorCount, dev );
			g_deviceState = 1;
		}
	}
	return g_deviceState == 1;
}

static bool cudaOk( cudaError_t e, const char* what )
{
	if ( e == cudaSuccess )
		return true;
	char buf[400];
	snprintf( buf, sizeof( buf ), "CUDA error in %s: %s", what, cudaGetErrorString( e ) );
	g_lastError = buf;
	fprintf( stderr, "forge2d_b200: %s\n", buf );
	return false;
}

// Host images are pinned when a device exists so that uploads/downloads run at full PCIe rate.
static void* backendHostAlloc( size_t bytes )
{
	size_t total = bytes + 256;
	char* raw = nullptr;
	bool pinned = false;
	if ( backendAvailable() )
	{
		void* p = nullptr;
		if ( cudaMallocHost( &p, total ) == cudaSuccess )
		{
			raw = static_cast<char*>( p );
			pinned = true;
		}
		else
			cudaGetLastError();
	}
	if ( raw == nullptr )
		raw = static_cast<char*>( aligned_alloc( 256, ( total + 255 ) / 256 * 256 ) );
	raw[0] = pinned ? 1 : 0;
	return raw + 256;
}
static void backendHostFree( void* p )
{
	if ( p == nullptr )
		return;
	char* raw = static_cast<char*>( p ) - 256;
	if ( raw[0] == 1 )
		cudaFreeHost( raw );
	else
		free( raw );
}

struct DeviceMirror
{
	World* dev = nullptr;
	size_t devBytes = 0;
	cudaStream_t stream = nullptr;
	int32_t* blockTotals = nullptr;
	unsigned layoutSerial = 0; // HostWorld::layoutSerial of the image `dev` holds
	cudaEvent_t ev[6] = {};
	bool timing = false;
	float times[5] = { 0, 0, 0, 0, 0 };
};

static DeviceMirror* mirror( HostWorld& hw )
{
	if ( hw.backend == nullptr )
	{
		DeviceMirror* m = new DeviceMirror();
		cudaStreamCreateWithFlags( &m->stream, cudaStreamNonBlocking );
		cudaMalloc( &m->blockTotals, 4096 * sizeof( int32_t ) );
		cudaMemset( m->blockTotals, 0, 4096 * sizeof( int32_t ) ); // [0,192): three GridBarriers (team, crew, rear), then per-block scan totals
		for ( int i = 0; i < 6; ++i )
			cudaEventCreate( &m->ev[i] );
		hw.backend = m;
	}
	return static_cast<DeviceMirror*>( hw.backend );
}


// bytes moved between host and device by the single-world paths (f2dGetTransferBytes: tests assert what a frame costs)
static unsigned long long g_bytesH2D = 0, g_bytesD2H = 0;

static bool hostImagePinned( const World* img ) { return ( reinterpret_cast<const char*>( img ) - 256 )[0] == 1; }

// `hostHeader`: the kernel mirrors the world header into the (pinned, device-accessible) host image itself
static bool launchWorld( HostWorld& hw, DeviceMirror* m, float dt, int sub, int phase, bool mirrorHeader = false )
{
	void* hostHeader = mirrorHeader && hostImagePinned( hw.img ) ? hw.img : nullptr;
	int mode = hw.launchMode;
	if ( mode < 0 )
		mode = hw.img->awakeBodies.count + hw.img->shapeIds.next > 4096 ? 1 : 0;
	g_launchCount += 1;
	if ( mode == 0 )
		return cudaOk( launchSingleCta( m->dev, dt, sub, phase, hostHeader, m->stream ), "stepWorldsCta launch" );
	int blocks = g_smCount;
	if ( blocks > 4000 )
		blocks = 4000;
	return cudaOk( launchSingleGrid( m->dev, m->blockTotals, blocks, dt, sub, phase, hostHeader, m->stream ), "stepWorldGrid cooperative launch" );
}

// The byte ranges of an image that hold live state, by its header: the header itself and the first `count` elements of
// every persistent array - what a re-layout preserves (imageRelayout). The step's work arrays (constraint fields, tree
// and island-split scratch, pair lists ...) and the unused tails of the entity arrays are most of an image and never
// need to cross PCIe. Ranges closer than a copy call is worth are merged.
static void liveRanges( const World& header, const Caps& caps, std::vector<ByteRange>& out )
{
	World copy = header;
	copy.sleepPool.count = copy.sleepUsed; // bump-allocated: its Arr::count is not maintained (see imageRelayout)
	std::vector<ArraySlot> slots;
	collectArrays( copy, caps, slots );
	out.clear();
	out.push_back( ByteRange{ 0, sizeof( World ) } );
	const uint64_t kMergeGap = 16 * 1024;
	for ( const ArraySlot& slot : slots )
	{
		if ( slot.persistent == false )
			continue;
		const int n = *slot.count < *slot.cap ? *slot.count : *slot.cap;
		if ( n <= 0 )
			continue;
		const uint64_t off = *slot.off, bytes = (uint64_t)n * (uint64_t)slot.elemSize;
		ByteRange& last = out.back();
		if ( off <= last.off + last.bytes + kMergeGap )
			last.bytes = off + bytes - last.off;
		else
			out.push_back( ByteRange{ off, bytes } );
	}
}

// Device copy allocated and current (uploads the host image when it is newer or the image was re-laid out)
static bool deviceImageCurrent( HostWorld& hw, DeviceMirror* m )
{
	World* img = hw.img;
	if ( hw.state == kHostNewer || m->dev == nullptr || m->devBytes != img->imageBytes || m->layoutSerial != hw.layoutSerial )
	{
		const bool newLayout = m->devBytes != img->imageBytes || m->layoutSerial != hw.layoutSerial;
		if ( newLayout || hw.wholeImageSync )
		{
			// a new allocation takes the whole image once (zeroes in everything that is not live state)
			if ( newLayout )
			{
				if ( m->dev )
					cudaFree( m->dev );
				m->dev = nullptr;
				if ( cudaOk( cudaMalloc( &m->dev, img->imageBytes ), "cudaMalloc(world image)" ) == false )
					return false;
				m->devBytes = img->imageBytes;
				m->layoutSerial = hw.layoutSerial;
			}
			hw.wholeImageSync = false;
			if ( cudaOk( cudaMemcpyAsync( m->dev, img, img->imageBytes, cudaMemcpyHostToDevice, m->stream ), "upload world image" ) == false )
				return false;
			g_bytesH2D += img->imageBytes;
		}
		else
		{
			std::vector<ByteRange> ranges;
			liveRanges( *img, hw.caps, ranges );
			for ( const ByteRange& r : ranges )
			{
				cudaMemcpyAsync( reinterpret_cast<char*>( m->dev ) + r.off, reinterpret_cast<const char*>( img ) + r.off, r.bytes,
								 cudaMemcpyHostToDevice, m->stream );
				g_bytesH2D += r.bytes;
			}
			if ( cudaOk( cudaGetLastError(), "upload world image" ) == false )
				return false;
		}
		hw.dirty.clear(); // all live state just went up
	}
	else if ( hw.state == kDeviceNewer )
	{
		flushDirty( hw ); // per-frame edits of a device-newer image (forces, impulses, velocities)
	}
	return true;
}

static void backendPhaseBegin( HostWorld& hw )
{
	deviceImageCurrent( hw, mirror( hw ) );
}
static void backendPhase( HostWorld& hw, float dt, int subSteps, int phase )
{
	DeviceMirror* m = mirror( hw );
	if ( m->dev != nullptr )
		launchWorld( hw, m, dt, subSteps, phase );
}
static void backendUploadRange( HostWorld& hw, uint64_t off, uint64_t bytes )
{
	DeviceMirror* m = mirror( hw );
	if ( m->dev == nullptr || bytes == 0 )
		return;
	cudaMemcpyAsync( reinterpret_cast<char*>( m->dev ) + off, reinterpret_cast<const char*>( hw.img ) + off, bytes, cudaMemcpyHostToDevice,
					 m->stream );
	g_bytesH2D += bytes;
	cudaOk( cudaStreamSynchronize( m->stream ), "upload range" ); // the host buffer is reused right away
}
static void backendPhaseEnd( HostWorld& hw )
{
	DeviceMirror* m = mirror( hw );
	if ( m->dev == nullptr )
		return;
	cudaMemcpyAsync( hw.img, m->dev, sizeof( World ), cudaMemcpyDeviceToHost, m->stream );
	g_bytesD2H += sizeof( World );
	hw.state = kDeviceNewer;
	hw.bodyMirrorFresh = false;
	cudaOk( cudaStreamSynchronize( m->stream ), "world step (callback-mediated)" );
}

static void backendStep( HostWorld& hw, float dt, int subSteps, bool synchronous )
{
	DeviceMirror* m = mirror( hw );
	World* img = hw.img;
	if ( deviceImageCurrent( hw, m ) == false )
		return;
	if ( m->timing )
	{
		cudaEventRecord( m->ev[0], m->stream );
		for ( int p = kPhaseBeginPairs; p <= kPhaseFinalize; ++p )
		{
			launchWorld( hw, m, dt, subSteps, p );
			cudaEventRecord( m->ev[p], m->stream );
		}
	}
	else
	{
		// the header travels back with every step (counters, error flags, event counts, capacities in use): written by
		// the kernel into the pinned host image, or copied when the image is not pinned
		launchWorld( hw, m, dt, subSteps, kPhaseAll, true );
	}
	if ( m->timing || hostImagePinned( img ) == false )
		cudaMemcpyAsync( img, m->dev, sizeof( World ), cudaMemcpyDeviceToHost, m->stream );
	g_bytesD2H += sizeof( World );
	hw.state = kDeviceNewer;
	hw.bodyMirrorFresh = false;
	if ( synchronous )
	{
		cudaOk( cudaStreamSynchronize( m->stream ), "world step" );
		if ( m->timing )
		{
			float total = 0.0f;
			for ( int p = 1; p <= 4; ++p )
			{
				cudaEventElapsedTime( &m->times[p - 1], m->ev[p - 1], m->ev[p] );
				total += m->times[p - 1];
			}
			m->times[4] = total;
		}
	}
}

static void backendSynchronize( HostWorld& hw )
{
	DeviceMirror* m = mirror( hw );
	cudaOk( cudaStreamSynchronize( m->stream ), "synchronize" );
}

static void backendDownload( HostWorld& hw )
{
	DeviceMirror* m = mirror( hw );
	if ( m->dev == nullptr )
		return;
	if ( hw.wholeImageSync )
	{
		hw.wholeImageSync = false;
		cudaStreamSynchronize( m->stream );
		cudaOk( cudaMemcpy( hw.img, m->dev, hw.img->imageBytes, cudaMemcpyDeviceToHost ), "download world image" );
		g_bytesD2H += hw.img->imageBytes;
		return;
	}
	// the device's header says what is live (the host copy of the header may be one asynchronous step behind)
	cudaMemcpyAsync( hw.img, m->dev, sizeof( World ), cudaMemcpyDeviceToHost, m->stream );
	cudaStreamSynchronize( m->stream );
	std::vector<ByteRange> ranges;
	liveRanges( *hw.img, hw.caps, ranges );
	for ( const ByteRange& r : ranges )
	{
		cudaMemcpyAsync( reinterpret_cast<char*>( hw.img ) + r.off, reinterpret_cast<const char*>( m->dev ) + r.off, r.bytes,
						 cudaMemcpyDeviceToHost, m->stream );
		g_bytesD2H += r.bytes;
	}
	cudaOk( cudaStreamSynchronize( m->stream ), "download world image" );
}

static void backendDownloadRange( HostWorld& hw, uint64_t off, uint64_t bytes )
{
	DeviceMirror* m = mirror( hw );
	if ( m->dev == nullptr || bytes == 0 )
		return;
	cudaMemcpyAsync( reinterpret_cast<char*>( hw.img ) + off, reinterpret_cast<const char*>( m->dev ) + off, bytes, cudaMemcpyDeviceToHost,
					 m->stream );
	g_bytesD2H += bytes;
	cudaOk( cudaStreamSynchronize( m->stream ), "download range" );
}

static void backendDownloadRanges( HostWorld& hw, std::initializer_list<ByteRange> ranges )
{
	DeviceMirror* m = mirror( hw );
	if ( m->dev == nullptr )
		return;
	for ( const ByteRange& r : ranges )
	{
		if ( r.bytes == 0 )
			continue;
		cudaMemcpyAsync( reinterpret_cast<char*>( hw.img ) + r.off, reinterpret_cast<const char*>( m->dev ) + r.off, r.bytes,
						 cudaMemcpyDeviceToHost, m->stream );
		g_bytesD2H += r.bytes;
	}
	cudaOk( cudaStreamSynchronize( m->stream ), "download ranges" );
}

static void backendRelease( HostWorld& hw )
{
	if ( hw.backend == nullptr )
		return;
	DeviceMirror* m = static_cast<DeviceMirror*>( hw.backend );
	cudaStreamSynchronize( m->stream );
	if ( m->dev )
		cudaFree( m->dev );
	cudaFree( m->blockTotals );
	for ( int i = 0; i < 6; ++i )
		cudaEventDestroy( m->ev[i] );
	cudaStreamDestroy( m->stream );
	delete m;
	hw.backend = nullptr;
}

static void backendStepTimes( HostWorld& hw, float* out5 )
{
	DeviceMirror* m = mirror( hw );
	for ( int i = 0; i < 5; ++i )
		out5[i] = m->times[i];
}
static void backendEnableTiming( HostWorld& hw, bool flag )
{
	mirror( hw )->timing = flag;
}

} // namespace f2d

// ------------------------------------------------------------------------------------------------ batch extension
struct f2dBatch
{
	char* dev = nullptr;
	unsigned long long stride = 0;
	int count = 0;
	cudaStream_t stream = nullptr;
	f2d::BodyMoveEvent* devEvents = nullptr;
	int* devCounts = nullptr;
	int eventCap = 0;
	unsigned int* devError = nullptr;
	f2d::Caps caps{};
	int threads = 128, blocksPerSM = 8; // launch configuration of the batch kernel (f2dBatch_SetLaunchConfig)
	cudaEvent_t events[8] = {};
	b2BodyMoveEvent* hostEvents = nullptr; // pinned staging for f2dBatch_ReadBodyEvents
	int* hostCounts = nullptr;
	int hostEventCap = 0;
	// f2dBatch_StepAndReadBodyEvents: the batch is stepped in slices of worlds, one stream each, so that the read-back
	// of a slice crosses PCIe while the next slices are still being stepped
	static constexpr int kSlices = 16;
	cudaStream_t sliceStreams[kSlices] = {};
	cudaEvent_t sliceDone[kSlices] = {};
	cudaEvent_t inputsReady = nullptr;
	// growth path: a world whose new contacts do not fit stops before its first structural edit (kErrRetry, f2d_step.h
	// stepPairs); the batch then moves every image into a larger layout on the device and lets the stopped worlds repeat
	// the step
	f2d::World layout{};			   // host copy of the header layout shared by all images (offsets, capacities)
	unsigned int* hostStatus = nullptr; // pinned: [0] OR of the error flags, [1] largest retryContacts
	unsigned int* devStatus = nullptr;
	unsigned int* devWorldFlags = nullptr;
	int growths = 0;
	// Launch mode: several worlds per block with phase-aligned teams (stepWorldsGang, the default), or one world per
	// block in the configuration `threads` x `blocksPerSM` (f2dBatch_SetLaunchConfig)
	bool gang = true;
	int* devQueue = nullptr; // world queues of the gang kernel: [0] the batch stream, [1 + i] slice i
	// pipelined step + read-back (f2dBatch_StepPipelined): two device / pinned host staging buffers, a copy stream
	struct Pipe
	{
		void* dev[2] = { nullptr, nullptr };
		void* host[2] = { nullptr, nullptr };
		int* devCounts[2] = { nullptr, nullptr };
		int* hostCounts[2] = { nullptr, nullptr }; // [count] counts, then one word: OR of the error flags
		unsigned int* devStatus[2] = { nullptr, nullptr };
		cudaEvent_t gathered[2] = { nullptr, nullptr }, copied[2] = { nullptr, nullptr };
		cudaStream_t copyStream = nullptr;
		size_t bytes = 0;
		int maxBodies = 0, format = 0;
		long long calls = 0;
		float lastDt = 0.0f;
		int lastSub = 0;
	} pipe;
	bool checkEveryStep = false; // (the last step of a call is checked by f2dBatch_Synchronize: StepN itself never blocks on it)
	bool pendingStatus = false;
	float pendingDt = 0.0f;
	int pendingSub = 0;
};

namespace f2d
{
// One step of `n` worlds starting at `base` on `stream` (`onlyRetry`: only the worlds that stopped for contact room)
static bool batchLaunchStep( f2dBatch* b, char* base, int n, float dt, int sub, bool onlyRetry, int queueIndex, cudaStream_t stream )
{
	g_launchCount += 1;
	if ( b->gang )
		return launchBatchStepGang( base, b->stride, n, dt, sub, onlyRetry ? 1 : 0, g_smCount, b->devQueue + queueIndex, stream );
	return launchBatchStep( b->threads, b->blocksPerSM, base, b->stride, n, dt, sub, onlyRetry ? -1 : 1, stream );
}

static void batchInitStatus( f2dBatch* b )
{
	cudaMalloc( &b->devQueue, ( f2dBatch::kSlices + 1 ) * sizeof( int ) );
	cudaMalloc( &b->devError, sizeof( unsigned int ) );
	cudaMemsetAsync( b->devError, 0, sizeof( unsigned int ), b->stream );
	cudaMalloc( &b->devStatus, 2 * sizeof( unsigned int ) );
	cudaMalloc( &b->devWorldFlags, (size_t)b->count * sizeof( unsigned int ) );
	cudaMallocHost( &b->hostStatus, 2 * sizeof( unsigned int ) );
}

// OR of the error flags of all worlds and the largest contact room a stopped world asked for; synchronises the stream
static unsigned int batchStatus( f2dBatch* b, int* retryContacts )
{
	cudaMemsetAsync( b->devStatus, 0, 2 * sizeof( unsigned int ), b->stream );
	launchGatherErrors( b->dev, b->stride, b->count, b->devStatus, b->stream );
	launchGatherWorldStatus( b->dev, b->stride, b->count, nullptr, reinterpret_cast<int*>( b->devStatus + 1 ), b->stream );
	g_launchCount += 2;
	cudaMemcpyAsync( b->hostStatus, b->devStatus, 2 * sizeof( unsigned int ), cudaMemcpyDeviceToHost, b->stream );
	cudaStreamSynchronize( b->stream );
	if ( retryContacts != nullptr )
		*retryContacts = (int)b->hostStatus[1];
	return b->hostStatus[0];
}

// Moves every image of the batch into a layout with room for `needContacts` more contacts (all on the device)
static bool batchGrow( f2dBatch* b, int needContacts )
{
	Caps caps = b->caps;
	caps.contacts = roundCap( caps.contacts + needContacts + ( needContacts >> 1 ), 256 );
	if ( b->layout.contactEventCapable > 0 )
		caps.contactEvents = caps.contacts;
	if ( b->layout.hitEventCapable > 0 )
		caps.hitEvents = caps.contacts;
	World newHeader = b->layout;
	const uint64_t bytes = layoutImage( newHeader, caps );
	const unsigned long long newStride = ( bytes + 255ull ) / 256ull * 256ull;
	std::vector<ArraySlot> oldSlots, newSlots;
	World oldHeader = b->layout;
	collectArrays( oldHeader, b->caps, oldSlots );
	collectArrays( newHeader, caps, newSlots );
	std::vector<RelayoutSlot> table( newSlots.size() );
	for ( size_t i = 0; i < newSlots.size(); ++i )
	{
		// ArraySlot::off points at Arr::off, the first member of the Arr record
		table[i].headerOffset = (int32_t)( reinterpret_cast<const char*>( newSlots[i].off ) - reinterpret_cast<const char*>( &newHeader ) );
		table[i].elemSize = newSlots[i].elemSize;
		table[i].persistent = newSlots[i].persistent ? 1 : 0;
	}
	char* newDev = nullptr;
	RelayoutSlot* devTable = nullptr;
	World* devHeader = nullptr;
	if ( cudaOk( cudaMalloc( &newDev, newStride * (unsigned long long)b->count ), "cudaMalloc(batch growth)" ) == false )
		return false;
	cudaMalloc( &devTable, table.size() * sizeof( RelayoutSlot ) );
	cudaMalloc( &devHeader, sizeof( World ) );
	cudaMemsetAsync( newDev, 0, newStride * (unsigned long long)b->count, b->stream );
	cudaMemcpyAsync( devTable, table.data(), table.size() * sizeof( RelayoutSlot ), cudaMemcpyHostToDevice, b->stream );
	cudaMemcpyAsync( devHeader, &newHeader, sizeof( World ), cudaMemcpyHostToDevice, b->stream );
	launchRelayoutWorlds( b->dev, b->stride, newDev, newStride, b->count, devTable, (int)table.size(), devHeader, b->stream );
	g_launchCount += 1;
	const bool ok = cudaOk( cudaStreamSynchronize( b->stream ), "batch growth" );
	cudaFree( devTable );
	cudaFree( devHeader );
	if ( ok == false )
	{
		cudaFree( newDev );
		return false;
	}
	cudaFree( b->dev );
	b->dev = newDev;
	b->stride = newStride;
	b->caps = caps;
	b->layout = newHeader;
	b->growths += 1;
	return true;
}

// After a step: worlds that stopped for contact room get it and repeat the step. Returns the remaining error flags.
static unsigned int batchResolveRetries( f2dBatch* b, float dt, int sub )
{
	int need = 0;
	unsigned int flags = batchStatus( b, &need );
	for ( int attempt = 0; ( flags & kErrRetry ) != 0 && attempt < 4; ++attempt )
	{
		if ( batchGrow( b, need ) == false )
			break;
		batchLaunchStep( b, b->dev, b->count, dt, sub, true, 0, b->stream );
		flags = batchStatus( b, &need );
	}
	if ( flags & ( kErrFatal | kErrRetry ) )
		reportError( "f2dBatch: world error flags 0x%x after the step (f2dBatch_GetWorldErrors names the worlds)", flags );
	return flags;
}
} // namespace f2d

extern "C" {

f2dBatch* f2dBatch_Create( b2WorldId templateWorld, int count )
{
	using namespace f2d;
	HostWorld* hw = worldFromId( templateWorld );
	if ( hw == nullptr || count <= 0 )
		return nullptr;
	if ( backendAvailable() == false )
	{
		reportError( "f2dBatch_Create: no CUDA device available - this library has no CPU fallback" );
		return nullptr;
	}
	if ( hw->img->hostCallbacks & ( kHostCustomFilter | kHostPreSolve ) )
	{
		reportError( "f2dBatch_Create: the template world has host callbacks registered; a batch steps without the host in the loop" );
		return nullptr;
	}
	hostImage( *hw );
	// every world of the batch keeps the template's capacities: leave room for the contacts the first steps create
	reserve( *hw, 0, 0, 4 * hw->img->moveArray.count + hw->img->shapeIds.next + 256, 0 );
	hw->state = kHostNewer;
	World* img = hw->img;
	f2dBatch* b = new f2dBatch();
	b->count = count;
	b->caps = hw->caps;
	b->stride = ( img->imageBytes + 255ull ) / 256ull * 256ull;
	if ( cudaOk( cudaMalloc( &b->dev, b->stride * (unsigned long long)count ), "cudaMalloc(batch)" ) == false )
	{
		delete b;
		return nullptr;
	}
	cudaStreamCreateWithFlags( &b->stream, cudaStreamNonBlocking );
	for ( int i = 0; i < 8; ++i )
		cudaEventCreate( &b->events[i] );
	cudaMemcpyAsync( b->dev, img, img->imageBytes, cudaMemcpyHostToDevice, b->stream );
	// replicate by doubling: log2(count) device-to-device copies
	int have = 1;
	while ( have < count )
	{
		int n = have < count - have ? have : count - have;
		cudaMemcpyAsync( b->dev + b->stride * (unsigned long long)have, b->dev, b->stride * (unsigned long long)n, cudaMemcpyDeviceToDevice,
						 b->stream );
		have += n;
	}
	b->layout = *img;
	batchInitStatus( b );
	cudaOk( cudaStreamSynchronize( b->stream ), "batch upload" );
	return b;
}

// A batch of DIFFERENT worlds: every world is brought to one common image layout (the largest capacities of any of
// them, plus contact head-room) and uploaded; `worlds` may name a world more than once. The host worlds stay usable.
f2dBatch* f2dBatch_CreateFromWorlds( const b2WorldId* worlds, int count )
{
	using namespace f2d;
	if ( worlds == nullptr || count <= 0 )
		return nullptr;
	if ( backendAvailable() == false )
	{
		reportError( "f2dBatch_CreateFromWorlds: no CUDA device available - this library has no CPU fallback" );
		return nullptr;
	}
	Caps common{};
	int contactEventCapable = 0, hitEventCapable = 0;
	for ( int i = 0; i < count; ++i )
	{
		HostWorld* hw = worldFromId( worlds[i] );
		if ( hw == nullptr )
		{
			reportError( "f2dBatch_CreateFromWorlds: world %d is not valid", i );
			return nullptr;
		}
		if ( hw->img->hostCallbacks & ( kHostCustomFilter | kHostPreSolve ) )
		{
			reportError( "f2dBatch_CreateFromWorlds: world %d has host callbacks registered; a batch steps without the host in the loop", i );
			return nullptr;
		}
		World* w = hostImage( *hw );
		const Caps& c = hw->caps;
		const int wantContacts = w->contactIds.next + 4 * w->moveArray.count + w->shapeIds.next + 256;
		common.bodies = std::max( common.bodies, c.bodies );
		common.shapes = std::max( common.shapes, c.shapes );
		common.contacts = std::max( common.contacts, std::max( c.contacts, roundCap( wantContacts, 256 ) ) );
		common.joints = std::max( common.joints, c.joints );
		common.sensors = std::max( common.sensors, c.sensors );
		contactEventCapable = std::max( contactEventCapable, w->contactEventCapable );
		hitEventCapable = std::max( hitEventCapable, w->hitEventCapable );
	}
	// event arrays: sized for the contacts of a world as soon as ANY world of the batch can emit such events
	common.contactEvents = contactEventCapable > 0 ? common.contacts : 16;
	common.hitEvents = hitEventCapable > 0 ? common.contacts : 16;
	common.sensorOverlap = sensorOverlapCapFor( common.shapes, common.sensors );
	f2dBatch* b = new f2dBatch();
	b->count = count;
	b->caps = common;
	for ( int i = 0; i < count; ++i )
	{
		HostWorld* hw = worldFromId( worlds[i] );
		const Caps& c = hw->caps;
		if ( c.bodies != common.bodies || c.shapes != common.shapes || c.contacts != common.contacts || c.joints != common.joints ||
			 c.sensors != common.sensors || c.contactEvents != common.contactEvents || c.hitEvents != common.hitEvents ||
			 c.sensorOverlap != common.sensorOverlap )
		{
			hw->img = imageRelayout( hw->img, common, backendHostAlloc, backendHostFree );
			hw->layoutSerial += 1;
			hw->caps = common;
			hw->state = kHostNewer;
		}
		if ( i == 0 )
		{
			b->stride = ( hw->img->imageBytes + 255ull ) / 256ull * 256ull;
			if ( cudaOk( cudaMalloc( &b->dev, b->stride * (unsigned long long)count ), "cudaMalloc(batch)" ) == false )
			{
				delete b;
				return nullptr;
			}
			cudaStreamCreateWithFlags( &b->stream, cudaStreamNonBlocking );
			for ( int k = 0; k < 8; ++k )
				cudaEventCreate( &b->events[k] );
			b->layout = *hw->img;
		}
		if ( hw->img->imageBytes > b->stride )
		{
			reportError( "f2dBatch_CreateFromWorlds: world %d does not fit the common layout", i );
			f2dBatch_Destroy( b );
			return nullptr;
		}
		cudaMemcpyAsync( b->dev + b->stride * (unsigned long long)i, hw->img, hw->img->imageBytes, cudaMemcpyHostToDevice, b->stream );
	}
	batchInitStatus( b );
	cudaOk( cudaStreamSynchronize( b->stream ), "batch upload" );
	return b;
}

static void pipeRelease( f2dBatch* b );

void f2dBatch_Destroy( f2dBatch* b )
{
	if ( b == nullptr )
		return;
	cudaStreamSynchronize( b->stream );
	cudaFree( b->dev );
	cudaFree( b->devEvents );
	cudaFree( b->devCounts );
	cudaFree( b->devError );
	if ( b->pipe.copyStream )
		cudaStreamSynchronize( b->pipe.copyStream );
	pipeRelease( b );
	cudaFree( b->devStatus );
	cudaFree( b->devWorldFlags );
	cudaFree( b->devQueue );
	if ( b->hostStatus )
		cudaFreeHost( b->hostStatus );
	if ( b->hostEvents )
		cudaFreeHost( b->hostEvents );
	if ( b->hostCounts )
		cudaFreeHost( b->hostCounts );
	for ( int i = 0; i < 8; ++i )
		cudaEventDestroy( b->events[i] );
	for ( int i = 0; i < f2dBatch::kSlices; ++i )
	{
		if ( b->sliceStreams[i] )
			cudaStreamDestroy( b->sliceStreams[i] );
		if ( b->sliceDone[i] )
			cudaEventDestroy( b->sliceDone[i] );
	}
	if ( b->inputsReady )
		cudaEventDestroy( b->inputsReady );
	cudaStreamDestroy( b->stream );
	delete b;
}

void f2dBatch_StepN( f2dBatch* b, float dt, int sub, int steps )
{
	using namespace f2d;
	if ( b == nullptr )
		return;
	// one launch per step keeps every world of the batch in lock-step (and gives ncu one launch per step)
	for ( int s = 0; s < steps; ++s )
	{
		if ( batchLaunchStep( b, b->dev, b->count, dt, sub, false, 0, b->stream ) == false )
		{
			reportError( "f2dBatch_Step: no batch kernel for %d threads x %d blocks/SM", b->threads, b->blocksPerSM );
			return;
		}
		// one cheap flag gather per step (queued behind it): a world that needs more contact room must get it before the
		// NEXT step, or it would fall behind the others. The host only waits when a flag is up.
		cudaMemsetAsync( b->devStatus, 0, 2 * sizeof( unsigned int ), b->stream );
		launchGatherErrors( b->dev, b->stride, b->count, b->devStatus, b->stream );
		g_launchCount += 1;
		if ( s + 1 < steps || b->checkEveryStep )
		{
			cudaMemcpyAsync( b->hostStatus, b->devStatus, sizeof( unsigned int ), cudaMemcpyDeviceToHost, b->stream );
			cudaStreamSynchronize( b->stream );
			if ( b->hostStatus[0] & kErrRetry )
				batchResolveRetries( b, dt, sub );
		}
		else
		{
			b->pendingStatus = true;
			b->pendingDt = dt;
			b->pendingSub = sub;
			cudaMemcpyAsync( b->hostStatus, b->devStatus, sizeof( unsigned int ), cudaMemcpyDeviceToHost, b->stream );
		}
	}
	cudaOk( cudaGetLastError(), "stepWorldsCta(batch) launch" );
}

int f2dBatch_SetLaunchConfig( f2dBatch* b, int threads, int blocksPerSM )
{
	if ( b == nullptr || f2d::batchConfigExists( threads, blocksPerSM ) == false )
		return 0;
	b->threads = threads;
	b->blocksPerSM = blocksPerSM;
	b->gang = false;
	return 1;
}

// 1 (default): several worlds per thread block, phase-aligned (stepWorldsGang); 0: one world per block
void f2dBatch_SetGangMode( f2dBatch* b, int on )
{
	if ( b )
		b->gang = on != 0;
}

void f2dBatch_Step( f2dBatch* b, float dt, int sub )
{
	f2dBatch_StepN( b, dt, sub, 1 );
	f2dBatch_Synchronize( b );
}

void f2dBatch_Synchronize( f2dBatch* b )
{
	if ( b == nullptr )
		return;
	f2d::cudaOk( cudaStreamSynchronize( b->stream ), "batch step" );
	if ( b->pendingStatus )
	{
		b->pendingStatus = false;
		if ( b->hostStatus[0] & f2d::kErrRetry )
			f2d::batchResolveRetries( b, b->pendingDt, b->pendingSub );
		else if ( b->hostStatus[0] & f2d::kErrFatal )
			f2d::reportError( "f2dBatch: world error flags 0x%x after the step (f2dBatch_GetWorldErrors names the worlds)", b->hostStatus[0] );
	}
}

// Error flags of every world (f2d kErr* bits; 0 = fine): `out` takes f2dBatch_GetWorldCount entries
int f2dBatch_GetWorldErrors( f2dBatch* b, uint32_t* out, int cap )
{
	using namespace f2d;
	if ( b == nullptr || out == nullptr )
		return 0;
	f2dBatch_Synchronize( b );
	launchGatherWorldStatus( b->dev, b->stride, b->count, b->devWorldFlags, nullptr, b->stream );
	g_launchCount += 1;
	const int n = cap < b->count ? cap : b->count;
	cudaMemcpyAsync( out, b->devWorldFlags, (size_t)n * sizeof( unsigned int ), cudaMemcpyDeviceToHost, b->stream );
	cudaStreamSynchronize( b->stream );
	int bad = 0;
	for ( int i = 0; i < n; ++i )
		bad += out[i] != 0 ? 1 : 0;
	return bad;
}
int f2dBatch_GetGrowthCount( f2dBatch* b ) { return b ? b->growths : 0; }

int f2dBatch_GetWorldCount( f2dBatch* b )
{
	return b ? b->count : 0;
}

int f2dBatch_GetBodyEvents( f2dBatch* b, b2BodyMoveEvent* out, int maxBodies, int* counts )
{
	using namespace f2d;
	if ( b == nullptr )
		return 0;
	int need = b->count * maxBodies;
	if ( need > b->eventCap )
	{
		cudaFree( b->devEvents );
		cudaFree( b->devCounts );
		cudaMalloc( &b->devEvents, (size_t)need * sizeof( BodyMoveEvent ) );
		cudaMalloc( &b->devCounts, (size_t)b->count * sizeof( int ) );
		b->eventCap = need;
	}
	launchGatherMoveEvents( b->dev, b->stride, b->count, b->devEvents, maxBodies, b->devCounts, b->stream );
	g_launchCount += 1;
	cudaMemcpyAsync( out, b->devEvents, (size_t)need * sizeof( BodyMoveEvent ), cudaMemcpyDeviceToHost, b->stream );
	cudaMemcpyAsync( counts, b->devCounts, (size_t)b->count * sizeof( int ), cudaMemcpyDeviceToHost, b->stream );
	cudaOk( cudaStreamSynchronize( b->stream ), "batch events" );
	int total = 0;
	for ( int i = 0; i < b->count; ++i )
		total += counts[i];
	return total;
}

// Same gather, but into a pinned staging buffer owned by the batch (full PCIe rate); like the reference's event arrays
// the returned pointers stay valid until the next call (B2/src/world.c:1491-1555 ownership rule).
int f2dBatch_ReadBodyEvents( f2dBatch* b, int maxBodies, const b2BodyMoveEvent** outEvents, const int** outCounts )
{
	using namespace f2d;
	if ( b == nullptr )
		return 0;
	int need = b->count * maxBodies;
	if ( need > b->eventCap )
	{
		cudaFree( b->devEvents );
		cudaFree( b->devCounts );
		cudaMalloc( &b->devEvents, (size_t)need * sizeof( BodyMoveEvent ) );
		cudaMalloc( &b->devCounts, (size_t)b->count * sizeof( int ) );
		b->eventCap = need;
	}
	if ( need > b->hostEventCap )
	{
		if ( b->hostEvents )
			cudaFreeHost( b->hostEvents );
		if ( b->hostCounts )
			cudaFreeHost( b->hostCounts );
		cudaMallocHost( &b->hostEvents, (size_t)need * sizeof( BodyMoveEvent ) );
		cudaMallocHost( &b->hostCounts, (size_t)b->count * sizeof( int ) );
		b->hostEventCap = need;
	}
	launchGatherMoveEvents( b->dev, b->stride, b->count, b->devEvents, maxBodies, b->devCounts, b->stream );
	g_launchCount += 1;
	cudaMemcpyAsync( b->hostEvents, b->devEvents, (size_t)need * sizeof( BodyMoveEvent ), cudaMemcpyDeviceToHost, b->stream );
	cudaMemcpyAsync( b->hostCounts, b->devCounts, (size_t)b->count * sizeof( int ), cudaMemcpyDeviceToHost, b->stream );
	cudaOk( cudaStreamSynchronize( b->stream ), "batch events" );
	*outEvents = b->hostEvents;
	*outCounts = b->hostCounts;
	int total = 0;
	for ( int i = 0; i < b->count; ++i )
		total += b->hostCounts[i];
	return total;
}

static bool ensureEventBuffers( f2dBatch* b, int maxBodies )
{
	using namespace f2d;
	int need = b->count * maxBodies;
	if ( need > b->eventCap )
	{
		cudaFree( b->devEvents );
		cudaFree( b->devCounts );
		cudaMalloc( &b->devEvents, (size_t)need * sizeof( BodyMoveEvent ) );
		cudaMalloc( &b->devCounts, (size_t)b->count * sizeof( int ) );
		b->eventCap = need;
	}
	if ( need > b->hostEventCap )
	{
		if ( b->hostEvents )
			cudaFreeHost( b->hostEvents );
		if ( b->hostCounts )
			cudaFreeHost( b->hostCounts );
		cudaMallocHost( &b->hostEvents, (size_t)need * sizeof( BodyMoveEvent ) );
		cudaMallocHost( &b->hostCounts, (size_t)b->count * sizeof( int ) );
		b->hostEventCap = need;
	}
	return cudaOk( cudaGetLastError(), "batch event buffers" );
}

// One step of every world AND the body move events of that step in the pinned staging buffer, as one call. Same
// results as f2dBatch_Step followed by f2dBatch_ReadBodyEvents; the difference is the schedule: the worlds are stepped
// in slices on separate streams (whatever was queued on the batch stream before - e.g. f2dBatch_SetGravity - is waited
// for first), each slice's events are gathered and copied to the host as soon as that slice is done, so all but the
// last slice's copy overlaps with stepping. Worlds are independent (B2/src/world.c:33-35), slices share nothing.
int f2dBatch_StepAndReadBodyEvents( f2dBatch* b, float dt, int sub, int maxBodies, const b2BodyMoveEvent** outEvents, const int** outCounts )
{
	using namespace f2d;
	if ( b == nullptr || maxBodies <= 0 )
		return 0;
	if ( ensureEventBuffers( b, maxBodies ) == false )
		return 0;
	if ( b->inputsReady == nullptr )
	{
		cudaEventCreateWithFlags( &b->inputsReady, cudaEventDisableTiming );
		// earlier slices get the higher stream priority: their blocks are dispatched first, so the slices finish one
		// after the other (and their copies start one after the other) instead of all together at the end
		int least = 0, greatest = 0;
		cudaDeviceGetStreamPriorityRange( &least, &greatest );
		for ( int i = 0; i < f2dBatch::kSlices; ++i )
		{
			int priority = greatest + i < least ? greatest + i : least;
			cudaStreamCreateWithPriority( &b->sliceStreams[i], cudaStreamNonBlocking, priority );
			cudaEventCreateWithFlags( &b->sliceDone[i], cudaEventDisableTiming );
		}
	}
	// slices of whole waves (resident blocks of the whole GPU) when the batch is large enough for that to matter
	int slices = f2dBatch::kSlices;
	if ( const char* e = getenv( "F2D_BATCH_SLICES" ) ) // tuning aid
		slices = atoi( e ) < 1 ? 1 : ( atoi( e ) > f2dBatch::kSlices ? f2dBatch::kSlices : atoi( e ) );
	while ( slices > 1 && b->count / slices < 148 * b->blocksPerSM / 2 )
		slices -= 1;
	cudaEventRecord( b->inputsReady, b->stream );
	int start = 0;
	for ( int i = 0; i < slices; ++i )
	{
		int end = (int)( (long long)b->count * ( i + 1 ) / slices );
		int n = end - start;
		cudaStream_t st = b->sliceStreams[i];
		cudaStreamWaitEvent( st, b->inputsReady, 0 );
		if ( n > 0 )
		{
			char* base = b->dev + b->stride * (unsigned long long)start;
			if ( batchLaunchStep( b, base, n, dt, sub, false, 1 + i, st ) == false )
			{
				reportError( "f2dBatch_StepAndReadBodyEvents: no batch kernel for %d threads x %d blocks/SM", b->threads, b->blocksPerSM );
				return 0;
			}
			BodyMoveEvent* devOut = b->devEvents + (size_t)start * maxBodies;
			launchGatherMoveEvents( base, b->stride, n, devOut, maxBodies, b->devCounts + start, st );
			g_launchCount += 1;
			cudaMemcpyAsync( b->hostEvents + (size_t)start * maxBodies, devOut, (size_t)n * maxBodies * sizeof( BodyMoveEvent ),
							 cudaMemcpyDeviceToHost, st );
			cudaMemcpyAsync( b->hostCounts + start, b->devCounts + start, (size_t)n * sizeof( int ), cudaMemcpyDeviceToHost, st );
		}
		cudaEventRecord( b->sliceDone[i], st );
		cudaStreamWaitEvent( b->stream, b->sliceDone[i], 0 );
		start = end;
	}
	cudaOk( cudaStreamSynchronize( b->stream ), "batch step + events" );
	{
		// a world that stopped for contact room repeats the step on the grown images; the events are then read again
		int need = 0;
		if ( batchStatus( b, &need ) & kErrRetry )
		{
			batchResolveRetries( b, dt, sub );
			launchGatherMoveEvents( b->dev, b->stride, b->count, b->devEvents, maxBodies, b->devCounts, b->stream );
			g_launchCount += 1;
			cudaMemcpyAsync( b->hostEvents, b->devEvents, (size_t)b->count * maxBodies * sizeof( BodyMoveEvent ), cudaMemcpyDeviceToHost,
							 b->stream );
			cudaMemcpyAsync( b->hostCounts, b->devCounts, (size_t)b->count * sizeof( int ), cudaMemcpyDeviceToHost, b->stream );
			cudaOk( cudaStreamSynchronize( b->stream ), "batch events after growth" );
		}
	}
	*outEvents = b->hostEvents;
	*outCounts = b->hostCounts;
	int total = 0;
	for ( int i = 0; i < b->count; ++i )
		total += b->hostCounts[i];
	return total;
}

// ---- pipelined step + read-back ------------------------------------------------------------------------------
// Call k queues step k, the packing of its results into a device staging buffer and their copy to pinned host memory
// (on a copy stream), then returns the results of step k-1, whose copy ran while step k was being computed: the GPU
// never waits for PCIe and the host never waits for more than max(step, copy). The first call returns nothing;
// f2dBatch_FlushPipelined returns the results of the last step. format 0: b2BodyMoveEvent records (40 bytes per body);
// format 1: transforms only (b2Transform, 16 bytes per body, awake order) - 2.5x less PCIe traffic.
static void pipeRelease( f2dBatch* b )
{
	f2dBatch::Pipe& p = b->pipe;
	for ( int i = 0; i < 2; ++i )
	{
		cudaFree( p.dev[i] );
		cudaFree( p.devCounts[i] ); // (devStatus is the last word of this allocation)
		if ( p.host[i] )
			cudaFreeHost( p.host[i] );
		if ( p.hostCounts[i] )
			cudaFreeHost( p.hostCounts[i] );
		if ( p.gathered[i] )
			cudaEventDestroy( p.gathered[i] );
		if ( p.copied[i] )
			cudaEventDestroy( p.copied[i] );
	}
	if ( p.copyStream )
		cudaStreamDestroy( p.copyStream );
	p = f2dBatch::Pipe();
}

static bool pipeEnsure( f2dBatch* b, int maxBodies, int format )
{
	f2dBatch::Pipe& p = b->pipe;
	if ( p.copyStream != nullptr && p.maxBodies == maxBodies && p.format == format )
		return true;
	if ( p.copyStream != nullptr )
	{
		cudaStreamSynchronize( b->stream );
		cudaStreamSynchronize( p.copyStream );
		pipeRelease( b );
	}
	const size_t record = format == 1 ? sizeof( f2d::Xf ) : sizeof( f2d::BodyMoveEvent );
	p.bytes = (size_t)b->count * (size_t)maxBodies * record;
	p.maxBodies = maxBodies;
	p.format = format;
	cudaStreamCreateWithFlags( &p.copyStream, cudaStreamNonBlocking );
	for ( int i = 0; i < 2; ++i )
	{
		cudaMalloc( &p.dev[i], p.bytes );
		cudaMalloc( &p.devCounts[i], ( (size_t)b->count + 1 ) * sizeof( int ) );
		p.devStatus[i] = reinterpret_cast<unsigned int*>( p.devCounts[i] + b->count );
		cudaMallocHost( &p.host[i], p.bytes );
		cudaMallocHost( &p.hostCounts[i], ( (size_t)b->count + 1 ) * sizeof( int ) );
		cudaEventCreateWithFlags( &p.gathered[i], cudaEventDisableTiming );
		cudaEventCreateWithFlags( &p.copied[i], cudaEventDisableTiming );
	}
	p.calls = 0;
	return f2d::cudaOk( cudaGetLastError(), "pipelined batch buffers" );
}

// Results of the step queued by call `call` (its copy is waited for); handles a world that stopped for contact room
static int pipeCollect( f2dBatch* b, long long call, const void** out, const int** counts )
{
	using namespace f2d;
	f2dBatch::Pipe& p = b->pipe;
	const int slot = (int)( call & 1 );
	cudaEventSynchronize( p.copied[slot] );
	const unsigned int flags = (unsigned int)p.hostCounts[slot][b->count];
	if ( flags & kErrRetry )
	{
		// rare: drain the pipeline, grow the images, let the stopped worlds catch up (they missed this step and, if one
		// was queued behind it, the next). Such a world reports no events for the step it had to wait in.
		cudaStreamSynchronize( b->stream );
		const int behind = (int)( p.calls - call );
		int need = 0;
		batchStatus( b, &need );
		for ( int attempt = 0; attempt < 4; ++attempt )
		{
			if ( batchGrow( b, need ) == false )
				break;
			for ( int k = 0; k < behind; ++k )
				batchLaunchStep( b, b->dev, b->count, p.lastDt, p.lastSub, true, 0, b->stream );
			if ( ( batchStatus( b, &need ) & kErrRetry ) == 0 )
				break;
		}
	}
	else if ( flags & kErrFatal )
	{
		reportError( "f2dBatch: world error flags 0x%x after the step (f2dBatch_GetWorldErrors names the worlds)", (int)flags );
	}
	*out = p.host[slot];
	*counts = p.hostCounts[slot];
	int total = 0;
	for ( int i = 0; i < b->count; ++i )
		total += p.hostCounts[slot][i];
	return total;
}

int f2dBatch_StepPipelined( f2dBatch* b, float dt, int sub, int maxBodies, int format, const void** out, const int** counts )
{
	using namespace f2d;
	if ( b == nullptr || maxBodies <= 0 || out == nullptr || counts == nullptr || ( format != 0 && format != 1 ) )
		return 0;
	if ( pipeEnsure( b, maxBodies, format ) == false )
		return 0;
	f2dBatch::Pipe& p = b->pipe;
	const long long call = p.calls;
	const int slot = (int)( call & 1 );
	// the staging buffers of this slot were last used by call - 2: its copy must be over before they are overwritten
	cudaStreamWaitEvent( b->stream, p.copied[slot], 0 );
	if ( batchLaunchStep( b, b->dev, b->count, dt, sub, false, 0, b->stream ) == false )
	{
		reportError( "f2dBatch_StepPipelined: no batch kernel for %d threads x %d blocks/SM", b->threads, b->blocksPerSM );
		return 0;
	}
	cudaMemsetAsync( p.devStatus[slot], 0, sizeof( unsigned int ), b->stream );
	if ( format == 1 )
	{
		launchGatherTransforms( b->dev, b->stride, b->count, p.dev[slot], maxBodies, p.devCounts[slot], p.devStatus[slot], b->stream );
	}
	else
	{
		launchGatherMoveEvents( b->dev, b->stride, b->count, static_cast<BodyMoveEvent*>( p.dev[slot] ), maxBodies, p.devCounts[slot], b->stream );
		launchGatherErrors( b->dev, b->stride, b->count, p.devStatus[slot], b->stream );
		g_launchCount += 1;
	}
	g_launchCount += 1;
	cudaEventRecord( p.gathered[slot], b->stream );
	cudaStreamWaitEvent( p.copyStream, p.gathered[slot], 0 );
	cudaMemcpyAsync( p.host[slot], p.dev[slot], p.bytes, cudaMemcpyDeviceToHost, p.copyStream );
	cudaMemcpyAsync( p.hostCounts[slot], p.devCounts[slot], ( (size_t)b->count + 1 ) * sizeof( int ), cudaMemcpyDeviceToHost, p.copyStream );
	cudaEventRecord( p.copied[slot], p.copyStream );
	p.calls = call + 1;
	p.lastDt = dt;
	p.lastSub = sub;
	if ( call == 0 )
	{
		*out = nullptr;
		*counts = nullptr;
		return 0;
	}
	return pipeCollect( b, call - 1, out, counts );
}

int f2dBatch_FlushPipelined( f2dBatch* b, const void** out, const int** counts )
{
	if ( b == nullptr || out == nullptr || counts == nullptr || b->pipe.calls == 0 )
		return 0;
	return pipeCollect( b, b->pipe.calls - 1, out, counts );
}

// Per-world gravity (the batch counterpart of b2World_SetGravity, box2d.h:135): one strided host->device copy that
// patches World::gravity of every image; queued on the batch stream ahead of the next step.
void f2dBatch_SetGravity( f2dBatch* b, const b2Vec2* gravity, int count )
{
	using namespace f2d;
	if ( b == nullptr || count <= 0 )
		return;
	if ( count > b->count )
		count = b->count;
	cudaOk( cudaMemcpy2DAsync( b->dev + offsetof( World, gravity ), b->stride, gravity, sizeof( b2Vec2 ), sizeof( b2Vec2 ), (size_t)count,
							   cudaMemcpyHostToDevice, b->stream ),
			"batch gravity upload" );
}

// Moves every world of the batch by its own offset (offsets[k] for world k): replicas of one template placed side by
// side, or decorrelated for a benchmark (same physics, different floating-point rounding per world).
void f2dBatch_TranslateWorlds( f2dBatch* b, const b2Vec2* offsets, int count )
{
	using namespace f2d;
	if ( b == nullptr || offsets == nullptr || count <= 0 )
		return;
	if ( count > b->count )
		count = b->count;
	void* dev = nullptr;
	cudaMalloc( &dev, (size_t)count * sizeof( b2Vec2 ) );
	cudaMemcpyAsync( dev, offsets, (size_t)count * sizeof( b2Vec2 ), cudaMemcpyHostToDevice, b->stream );
	launchTranslateWorlds( b->dev, b->stride, count, dev, b->stream );
	g_launchCount += 1;
	cudaOk( cudaStreamSynchronize( b->stream ), "batch translate" );
	cudaFree( dev );
}

void f2dBatch_EventRecord( f2dBatch* b, int slot )
{
	if ( b && slot >= 0 && slot < 8 )
		cudaEventRecord( b->events[slot], b->stream );
}
float f2dBatch_EventElapsedMs( f2dBatch* b, int from, int to )
{
	float ms = -1.0f;
	if ( b && from >= 0 && from < 8 && to >= 0 && to < 8 )
	{
		cudaEventSynchronize( b->events[to] );
		cudaEventElapsedTime( &ms, b->events[from], b->events[to] );
	}
	return ms;
}
unsigned long long f2dBatch_GetWorldBytes( f2dBatch* b )
{
	return b ? b->stride : 0ull;
}

int f2dSetDevice( int device )
{
	f2d::g_deviceState = -1;
	return cudaSetDevice( device ) == cudaSuccess && f2d::backendAvailable() ? 1 : 0;
}
// Bytes moved between host and device by single-world calls since the library was loaded
void f2dGetTransferBytes( unsigned long long* hostToDevice, unsigned long long* deviceToHost )
{
	if ( hostToDevice )
		*hostToDevice = f2d::g_bytesH2D;
	if ( deviceToHost )
		*deviceToHost = f2d::g_bytesD2H;
}
void* f2dHostAlloc( unsigned long long bytes )
{
	void* p = nullptr;
	if ( cudaMallocHost( &p, (size_t)bytes ) != cudaSuccess )
	{
		cudaGetLastError();
		return nullptr;
	}
	return p;
}
void f2dHostFree( void* p )
{
	if ( p )
		cudaFreeHost( p );
}

void f2dBatch_DownloadWorld( f2dBatch* b, int index, b2WorldId into )
{
	using namespace f2d;
	HostWorld* hw = worldFromId( into );
	if ( b == nullptr || hw == nullptr || index < 0 || index >= b->count )
		return;
	cudaStreamSynchronize( b->stream );
	World header;
	cudaMemcpy( &header, b->dev + b->stride * (unsigned long long)index, sizeof( World ), cudaMemcpyDeviceToHost );
	uint16_t worldId = hw->img->worldId, generation = hw->img->generation;
	backendHostFree( hw->img );
	hw->img = static_cast<World*>( backendHostAlloc( header.imageBytes ) );
	cudaOk( cudaMemcpy( hw->img, b->dev + b->stride * (unsigned long long)index, header.imageBytes, cudaMemcpyDeviceToHost ),
			"download batch world" );
	hw->img->worldId = worldId;
	hw->img->generation = generation;
	hw->caps = b->caps;
	hw->state = kHostNewer;
}

uint32_t f2dBatch_GetErrorFlags( f2dBatch* b )
{
	using namespace f2d;
	if ( b == nullptr )
		return 0;
	cudaMemsetAsync( b->devError, 0, sizeof( unsigned int ), b->stream );
	launchGatherErrors( b->dev, b->stride, b->count, b->devError, b->stream );
	g_launchCount += 1;
	unsigned int e = 0;
	cudaMemcpyAsync( &e, b->devError, sizeof( e ), cudaMemcpyDeviceToHost, b->stream );
	cudaStreamSynchronize( b->stream );
	return e;
}

} // extern "C"
