// forge2d_b200 — batch kernel (one thread block per world), configuration 128x8 (the default).
// One step kernel per translation unit: the variants compile in parallel (and ptxas 12.9 crashes on a module that holds
// two instantiations of the step).
#define F2D_SMALL_TEAM_KERNELS 1 // (f2d_math.h F2D_HDC)
#include "f2d_kernels.cuh"

#include <stdlib.h>

namespace f2d
{
bool launchBatchStepC( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
						cudaStream_t stream )
{
	if ( threads == 128 && blocksPerSM == 8 )
	{
		// profiling aid: F2D_PROFILE_PHASE_LAUNCHES=1 launches the phases of the step one after the other, so that ncu
		// attributes duration and DRAM traffic to pairs / tree rebuild / narrowphase / state pass / solve / finalize
		// (the results are the same)
		static const bool phaseLaunches = getenv( "F2D_PROFILE_PHASE_LAUNCHES" ) != nullptr;
		if ( phaseLaunches && steps == 1 )
		{
			const int phases[] = { kPhaseBeginPairs, kPhaseCollideTreeOnly, kPhaseCollideNarrowOnly, kPhaseCollideFinish, kPhaseSolve, kPhaseFinalize };
			for ( int phase : phases )
				stepWorldsCta<128, 8><<<worldCount, 128, 0, stream>>>( base, stride, worldCount, dt, sub, phase, 1, 0, nullptr );
			return true;
		}
		stepWorldsCta<128, 8><<<worldCount, 128, 0, stream>>>( base, stride, worldCount, dt, sub, kPhaseAll, steps, 0, nullptr );
		return true;
	}
	return false;
}
} // namespace f2d
