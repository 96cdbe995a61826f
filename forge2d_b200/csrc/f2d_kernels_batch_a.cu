// forge2d_b200 — batch kernels (one thread block per world), configurations 256x2, 256x4.
// Separate translation unit so the variants compile in parallel.
#include "f2d_kernels.cuh"

namespace f2d
{
bool launchBatchStepA( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
						cudaStream_t stream )
{
	if ( threads == 256 && blocksPerSM == 2 )
	{
		stepWorldsCta<256, 2><<<worldCount, 256, 0, stream>>>( base, stride, worldCount, dt, sub, kPhaseAll, steps );
		return true;
	}
	if ( threads == 256 && blocksPerSM == 4 )
	{
		stepWorldsCta<256, 4><<<worldCount, 256, 0, stream>>>( base, stride, worldCount, dt, sub, kPhaseAll, steps );
		return true;
	}
	return false;
}
} // namespace f2d
