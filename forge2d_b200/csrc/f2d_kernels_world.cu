// forge2d_b200 — kernel for ONE small world: a single 1024-thread block (barrier cost ~tens of ns, working set in that
// SM's L1/L2), plus the batch gather kernels. The cooperative-grid kernel for large worlds is in f2d_kernels_grid.cu.
#include "f2d_kernels.cuh"

#include <stdlib.h>

namespace f2d
{

// Gathers the body move events of every world of a batch into one dense buffer (one block per world).
__global__ void gatherMoveEvents( const char* base, unsigned long long stride, int worldCount, BodyMoveEvent* out, int maxBodies,
								  int* counts )
{
	int wi = (int)blockIdx.x;
	if ( wi >= worldCount )
		return;
	const World* w = reinterpret_cast<const World*>( base + (unsigned long long)wi * stride );
	int n = min( w->moveEvents.count, maxBodies );
	// not ptr(): a batch that was uploaded but never stepped has no deviceBase yet
	const BodyMoveEvent* src = reinterpret_cast<const BodyMoveEvent*>( reinterpret_cast<const char*>( w ) + w->moveEvents.off );
	// 40-byte records copied as 8-byte words: coalesced, no struct padding games
	const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>( src );
	unsigned long long* d8 = reinterpret_cast<unsigned long long*>( out + (size_t)wi * maxBodies );
	int words = n * (int)( sizeof( BodyMoveEvent ) / 8 );
	for ( int i = (int)threadIdx.x; i < words; i += (int)blockDim.x )
		d8[i] = s8[i];
	if ( threadIdx.x == 0 )
		counts[wi] = n;
}

// The body transforms of every world of a batch, packed: 16 bytes per body (position, rotation) in awake order, the
// per-world counts, and the OR of all error flags in status[0] (one block per world).
__global__ void gatherTransforms( const char* base, unsigned long long stride, int worldCount, Xf* out, int maxBodies, int* counts,
								  unsigned int* status )
{
	int wi = (int)blockIdx.x;
	if ( wi >= worldCount )
		return;
	const World* w = reinterpret_cast<const World*>( base + (unsigned long long)wi * stride );
	int n = min( w->moveEvents.count, maxBodies );
	const BodyMoveEvent* src = reinterpret_cast<const BodyMoveEvent*>( reinterpret_cast<const char*>( w ) + w->moveEvents.off );
	Q4* dst = reinterpret_cast<Q4*>( out + (size_t)wi * maxBodies );
	for ( int i = (int)threadIdx.x; i < n; i += (int)blockDim.x )
	{
		const Xf xf = src[i].transform; // 40-byte records: 8-byte aligned
		dst[i] = Q4{ xf.p.x, xf.p.y, xf.q.c, xf.q.s };
	}
	if ( threadIdx.x == 0 )
	{
		counts[wi] = n;
		if ( w->error != 0 )
			atomicOr( status, w->error );
	}
}

__global__ void gatherErrors( const char* base, unsigned long long stride, int worldCount, unsigned int* out )
{
	int wi = (int)( blockIdx.x * blockDim.x + threadIdx.x );
	if ( wi >= worldCount )
		return;
	const World* w = reinterpret_cast<const World*>( base + (unsigned long long)wi * stride );
	if ( w->error )
		atomicOr( out, w->error );
}



// Per-world error flags and the contact room a stopped world asked for (World::step.retryContacts), one thread per world
__global__ void gatherWorldStatus( const char* base, unsigned long long stride, int worldCount, unsigned int* flags, int* retryMax )
{
	int wi = (int)( blockIdx.x * blockDim.x + threadIdx.x );
	if ( wi >= worldCount )
		return;
	const World* w = reinterpret_cast<const World*>( base + (unsigned long long)wi * stride );
	if ( flags != nullptr )
		flags[wi] = w->error;
	if ( retryMax != nullptr && ( w->error & kErrRetry ) != 0 )
		atomicMax( retryMax, w->step.retryContacts );
}

// Moves every world of a batch into a larger image layout (more contact room): one block per world. `slots` lists, for
// every array of the image, where its Arr record sits in the header, its element size and whether its contents
// persist; `newHeader` is a header laid out for the new capacities (offsets, capacities, consStride, imageBytes).
__global__ void relayoutWorlds( const char* oldBase, unsigned long long oldStride, char* newBase, unsigned long long newStride, int worldCount,
								const RelayoutSlot* slots, int slotCount, const World* newHeader )
{
	const int wi = (int)blockIdx.x;
	if ( wi >= worldCount )
		return;
	const char* from = oldBase + (unsigned long long)wi * oldStride;
	char* to = newBase + (unsigned long long)wi * newStride;
	const World* oldW = reinterpret_cast<const World*>( from );
	// header: the old one, with the new layout patched in below
	for ( int i = (int)threadIdx.x; i < (int)( sizeof( World ) / 16 ); i += (int)blockDim.x )
		reinterpret_cast<uint4*>( to )[i] = reinterpret_cast<const uint4*>( from )[i];
	__syncthreads();
	for ( int k = 0; k < slotCount; ++k )
	{
		const RelayoutSlot slot = slots[k];
		const Arr<char>& oldArr = *reinterpret_cast<const Arr<char>*>( from + slot.headerOffset );
		const Arr<char>& newArr = *reinterpret_cast<const Arr<char>*>( reinterpret_cast<const char*>( newHeader ) + slot.headerOffset );
		if ( slot.persistent )
		{
			const int keep = oldArr.cap < newArr.cap ? oldArr.cap : newArr.cap;
			const unsigned long long bytes = (unsigned long long)keep * (unsigned long long)slot.elemSize; // multiple of 4
			const uint32_t* src = reinterpret_cast<const uint32_t*>( from + oldArr.off );
			uint32_t* dst = reinterpret_cast<uint32_t*>( to + newArr.off );
			for ( unsigned long long i = threadIdx.x; i < bytes / 4; i += blockDim.x )
				dst[i] = src[i];
		}
		if ( threadIdx.x == 0 )
		{
			Arr<char>& out = *reinterpret_cast<Arr<char>*>( to + slot.headerOffset );
			out.off = newArr.off;
			out.cap = newArr.cap;
			out.count = oldArr.count;
		}
	}
	if ( threadIdx.x == 0 )
	{
		World* w = reinterpret_cast<World*>( to );
		w->consStride = newHeader->consStride;
		w->imageBytes = newHeader->imageBytes;
		w->sensorOverlapCap = oldW->sensorOverlapCap;
	}
}

// Translates every world of a batch by its own offset: bodies (origin, centre of mass, sweep start), shape boxes and
// the boxes of all broadphase tree nodes. One block per world. Nothing else in the image holds world coordinates that
// the step reads (manifold anchors are relative to the bodies; joint frames are local).
__global__ void translateWorlds( char* base, unsigned long long stride, int worldCount, const V2* offsets )
{
	const int wi = (int)blockIdx.x;
	if ( wi >= worldCount )
		return;
	World* w = reinterpret_cast<World*>( base + (unsigned long long)wi * stride );
	char* image = reinterpret_cast<char*>( w );
	const V2 d = offsets[wi];
	BodySim* sims = reinterpret_cast<BodySim*>( image + w->sims.off );
	const Body* bodies = reinterpret_cast<const Body*>( image + w->bodies.off );
	for ( int i = (int)threadIdx.x; i < w->bodies.count; i += (int)blockDim.x )
	{
		if ( bodies[i].id != i )
			continue;
		BodySim& s = sims[i];
		s.transform.p = add( s.transform.p, d );
		s.center = add( s.center, d );
		s.center0 = add( s.center0, d );
	}
	Shape* shapes = reinterpret_cast<Shape*>( image + w->shapes.off );
	for ( int i = (int)threadIdx.x; i < w->shapes.count; i += (int)blockDim.x )
	{
		Shape& s = shapes[i];
		if ( s.id != i )
			continue;
		s.aabb = Box{ add( s.aabb.lo, d ), add( s.aabb.hi, d ) };
		s.fatAABB = Box{ add( s.fatAABB.lo, d ), add( s.fatAABB.hi, d ) };
	}
	for ( int t = 0; t < 3; ++t )
	{
		TreeNode* nodes = reinterpret_cast<TreeNode*>( image + w->trees[t].nodes.off );
		for ( int i = (int)threadIdx.x; i < w->trees[t].nodes.count; i += (int)blockDim.x )
		{
			TreeNode& n = nodes[i];
			if ( n.flags & kNodeAllocated )
				n.box = Box{ add( n.box.lo, d ), add( n.box.hi, d ) };
		}
	}
}

constexpr int kSingleCtaThreads = 512;

// The one-block kernel gets a shared-memory work area (CtaTeam::arena): the block has the SM to itself.
static int singleCtaArenaBytes()
{
	static int bytes = -1;
	if ( bytes < 0 )
	{
		const char* kb = getenv( "F2D_SINGLE_ARENA_KB" ); // tuning aid
		bytes = ( kb != nullptr ? atoi( kb ) : 56 ) * 1024; // with the static 4 KB: the 64 KB carve-out (measured on bench2d: 96 KB costs 60 us per frame, less L1)
		if ( bytes > 200 * 1024 )
			bytes = 200 * 1024;
		if ( bytes > 0 &&
			 cudaFuncSetAttribute( stepWorldsCta<kSingleCtaThreads, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes ) != cudaSuccess )
		{
			cudaGetLastError();
			bytes = 0;
		}
	}
	return bytes;
}

bool launchBatchStepSolo( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
						  cudaStream_t stream )
{
	if ( threads != kSingleCtaThreads || blocksPerSM != 1 )
		return false;
	int smCount = 0, device = 0;
	cudaGetDevice( &device );
	cudaDeviceGetAttribute( &smCount, cudaDevAttrMultiProcessorCount, device );
	const int arenaBytes = singleCtaArenaBytes();
	const int blocks = worldCount < smCount ? worldCount : smCount;
	stepWorldsCta<kSingleCtaThreads, 1><<<blocks, kSingleCtaThreads, arenaBytes, stream>>>( base, stride, worldCount, dt, sub, kPhaseAll, steps,
																							arenaBytes, nullptr );
	return cudaGetLastError() == cudaSuccess;
}

cudaError_t launchSingleCta( World* dev, float dt, int sub, int phase, void* hostHeader, cudaStream_t stream )
{
	const int arenaBytes = singleCtaArenaBytes();
	stepWorldsCta<kSingleCtaThreads, 1><<<1, kSingleCtaThreads, arenaBytes, stream>>>( reinterpret_cast<char*>( dev ), 0ull, 1, dt, sub, phase, 1,
																					   arenaBytes, static_cast<uint4*>( hostHeader ) );
	return cudaGetLastError();
}

void launchGatherMoveEvents( const char* base, unsigned long long stride, int worldCount, BodyMoveEvent* out, int maxBodies, int* counts,
							 cudaStream_t stream )
{
	gatherMoveEvents<<<worldCount, 256, 0, stream>>>( base, stride, worldCount, out, maxBodies, counts );
}
void launchGatherTransforms( const char* base, unsigned long long stride, int worldCount, void* out, int maxBodies, int* counts,
							 unsigned int* status, cudaStream_t stream )
{
	gatherTransforms<<<worldCount, 256, 0, stream>>>( base, stride, worldCount, static_cast<Xf*>( out ), maxBodies, counts, status );
}
void launchGatherWorldStatus( const char* base, unsigned long long stride, int worldCount, unsigned int* flags, int* retryMax,
							 cudaStream_t stream )
{
	gatherWorldStatus<<<( worldCount + 255 ) / 256, 256, 0, stream>>>( base, stride, worldCount, flags, retryMax );
}
void launchRelayoutWorlds( const char* oldBase, unsigned long long oldStride, char* newBase, unsigned long long newStride, int worldCount,
						   const RelayoutSlot* slots, int slotCount, const World* newHeader, cudaStream_t stream )
{
	relayoutWorlds<<<worldCount, 256, 0, stream>>>( oldBase, oldStride, newBase, newStride, worldCount, slots, slotCount, newHeader );
}
void launchTranslateWorlds( char* base, unsigned long long stride, int worldCount, const void* offsets, cudaStream_t stream )
{
	translateWorlds<<<worldCount, 256, 0, stream>>>( base, stride, worldCount, static_cast<const V2*>( offsets ) );
}
void launchGatherErrors( const char* base, unsigned long long stride, int worldCount, unsigned int* out, cudaStream_t stream )
{
	gatherErrors<<<( worldCount + 255 ) / 256, 256, 0, stream>>>( base, stride, worldCount, out );
}

} // namespace f2d
