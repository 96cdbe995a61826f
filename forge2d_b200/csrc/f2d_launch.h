// forge2d_b200 — host-visible launcher declarations (no device code): implemented by the f2d_kernels_*.cu units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f2d_step.h" // World, Phase

namespace f2d
{
bool launchBatchStepA( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
					   cudaStream_t stream );
bool launchBatchStepB( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
					   cudaStream_t stream );
bool launchBatchStepC( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
					   cudaStream_t stream );
// several worlds per block, phase-aligned (f2d_kernels.cuh stepWorldsGang); `queue`: one device int per concurrent launch
bool launchBatchStepGang( char* base, unsigned long long stride, int worldCount, float dt, int sub, int onlyRetry, int smCount, int* queue,
						  cudaStream_t stream );
// one world per SM at a time in the single-world kernel (512 threads, shared-memory work area, front / rear teams): the
// worlds are taken from the batch in a block-stride loop
bool launchBatchStepSolo( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub, int steps,
						  cudaStream_t stream );
inline bool batchConfigExists( int threads, int blocksPerSM )
{
	const int known[][2] = { { 256, 4 }, { 128, 8 }, { 64, 16 }, { 512, 1 } };
	for ( auto& k : known )
		if ( k[0] == threads && k[1] == blocksPerSM )
			return true;
	return false;
}
inline bool launchBatchStep( int threads, int blocksPerSM, char* base, unsigned long long stride, int worldCount, float dt, int sub,
							 int steps, cudaStream_t stream )
{
	return launchBatchStepA( threads, blocksPerSM, base, stride, worldCount, dt, sub, steps, stream ) ||
		   launchBatchStepB( threads, blocksPerSM, base, stride, worldCount, dt, sub, steps, stream ) ||
		   launchBatchStepC( threads, blocksPerSM, base, stride, worldCount, dt, sub, steps, stream ) ||
		   launchBatchStepSolo( threads, blocksPerSM, base, stride, worldCount, dt, sub, steps, stream );
}
// hostHeader: device-accessible address of the pinned host image (the kernel mirrors the header there), or nullptr
cudaError_t launchSingleCta( World* dev, float dt, int sub, int phase, void* hostHeader, cudaStream_t stream );
cudaError_t launchSingleGrid( World* dev, int32_t* blockTotals, int blocks, float dt, int sub, int phase, void* hostHeader,
							  cudaStream_t stream );
// batch growth (f2d_kernels_world.cu relayoutWorlds): one entry per array of the image
struct RelayoutSlot
{
	int32_t headerOffset; // byte offset of the array's Arr record inside World
	int32_t elemSize;
	int32_t persistent;
};
void launchGatherWorldStatus( const char* base, unsigned long long stride, int worldCount, unsigned int* flags, int* retryMax,
							 cudaStream_t stream );
void launchRelayoutWorlds( const char* oldBase, unsigned long long oldStride, char* newBase, unsigned long long newStride, int worldCount,
						   const RelayoutSlot* slots, int slotCount, const World* newHeader, cudaStream_t stream );
void launchTranslateWorlds( char* base, unsigned long long stride, int worldCount, const void* offsets, cudaStream_t stream );
// gather kernels of the batch extension
struct BodyMoveEvent;
void launchGatherMoveEvents( const char* base, unsigned long long stride, int worldCount, BodyMoveEvent* out, int maxBodies, int* counts,
							 cudaStream_t stream );
void launchGatherTransforms( const char* base, unsigned long long stride, int worldCount, void* out, int maxBodies, int* counts,
							 unsigned int* status, cudaStream_t stream );
void launchGatherErrors( const char* base, unsigned long long stride, int worldCount, unsigned int* out, cudaStream_t stream );
} // namespace f2d
