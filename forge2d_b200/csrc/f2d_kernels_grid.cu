// forge2d_b200 — kernel for ONE large world: a cooperative grid of one block per SM.
#include "f2d_kernels.cuh"

namespace f2d
{

constexpr int kGridThreads = 512;

cudaError_t launchSingleGrid( World* dev, int32_t* blockTotals, int blocks, float dt, int sub, int phase, void* hostHeader,
							  cudaStream_t stream )
{
	uint4* host = static_cast<uint4*>( hostHeader );
	void* args[] = { &dev, &blockTotals, &dt, &sub, &phase, &host };
	return cudaLaunchCooperativeKernel( (void*)stepWorldGrid<kGridThreads>, dim3( blocks ), dim3( kGridThreads ), args, 0, stream );
}

} // namespace f2d
