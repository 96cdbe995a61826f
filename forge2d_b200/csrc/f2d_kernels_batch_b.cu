// forge2d_b200 — batch kernel (one thread block per world), configuration 64x16.
// One step kernel per translation unit: the variants compile in parallel (and ptxas 12.9 crashes on a module that holds
// two instantiations of the step).
#define F2D_SMALL_TEAM_KERNELS 1 // (f2d_math.h F2D_HDC)
#include "f2d_kernels.cuh"

namespace f2d
{
bool launchBatchStepB( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
						cudaStream_t stream )
{
	if ( threads == 64 && blocksPerSM == 16 )
	{
		stepWorldsCta<64, 16><<<worldCount, 64, 0, stream>>>( base, stride, worldCount, dt, sub, kPhaseAll, steps, 0, nullptr );
		return true;
	}
	return false;
}
} // namespace f2d
