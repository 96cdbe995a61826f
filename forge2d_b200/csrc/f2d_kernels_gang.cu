// forge2d_b200 — batch kernel with several worlds per thread block (stepWorldsGang): the default for batches.
#define F2D_SMALL_TEAM_KERNELS 1 // (f2d_math.h F2D_HDC)
#include "f2d_kernels.cuh"

namespace f2d
{

constexpr int kGangTeamThreads = 128;
constexpr int kGangTeams = 7; // 896 threads, 72 registers each: one block per SM (measured on 8192 decorrelated bench2d worlds:
							  // 128x7 23.8 ms per step, 192x5 24.9, 256x4 25.2, 96x10 25.1, 64x14 25.2; one world per block 33.5)

bool launchBatchStepGang( char* base, unsigned long long stride, int worldCount, float dt, int sub, int onlyRetry, int smCount, int* queue,
						  cudaStream_t stream )
{
	if ( worldCount <= 0 )
		return true;
	cudaMemsetAsync( queue, 0, sizeof( int ), stream );
	const int gangs = ( worldCount + kGangTeams - 1 ) / kGangTeams;
	const int blocks = gangs < smCount ? gangs : smCount;
	stepWorldsGang<kGangTeamThreads, kGangTeams><<<blocks, kGangTeamThreads * kGangTeams, 0, stream>>>( base, stride, worldCount, dt, sub, queue,
																									  onlyRetry );
	return cudaGetLastError() == cudaSuccess;
}

} // namespace f2d
