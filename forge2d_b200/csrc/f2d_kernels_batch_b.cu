// forge2d_b200 — batch kernels (one thread block per world), configurations 32x32.
// Separate translation unit so the variants compile in parallel.
#include "f2d_kernels.cuh"

namespace f2d
{
bool launchBatchStepB( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
						cudaStream_t stream )
{
	if ( threads == 32 && blocksPerSM == 32 )
	{
		stepWorldsCta<32, 32><<<worldCount, 32, 0, stream>>>( base, stride, worldCount, dt, sub, kPhaseAll, steps );
		return true;
	}
	return false;
}
} // namespace f2d
