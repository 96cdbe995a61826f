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

__global__ void gatherErrors( const char* base, unsigned long long stride, int worldCount, unsigned int* out )
{
	int wi = (int)( blockIdx.x * blockDim.x + threadIdx.x );
	if ( wi >= worldCount )
		return;
	const World* w = reinterpret_cast<const World*>( base + (unsigned long long)wi * stride );
	if ( w->error )
		atomicOr( out, w->error );
}



constexpr int kSingleCtaThreads = 1024;

// The one-block kernel gets a shared-memory work area (CtaTeam::arena): the block has the SM to itself.
static int singleCtaArenaBytes()
{
	static int bytes = -1;
	if ( bytes < 0 )
	{
		const char* kb = getenv( "F2D_SINGLE_ARENA_KB" ); // tuning aid
		bytes = ( kb != nullptr ? atoi( kb ) : 96 ) * 1024;
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
void launchGatherErrors( const char* base, unsigned long long stride, int worldCount, unsigned int* out, cudaStream_t stream )
{
	gatherErrors<<<( worldCount + 255 ) / 256, 256, 0, stream>>>( base, stride, worldCount, out );
}

} // namespace f2d
