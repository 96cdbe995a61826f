// forge2d_b200 — CUDA kernels of the world step (sm_100a): execution teams + kernel templates, shared by the kernel
// translation units (f2d_kernels_*.cu). The step phases themselves live in f2d_step.h and are instantiated here for
//   * CtaTeam  — one thread block per world (`__syncthreads` between phases): batches of worlds, small worlds;
//   * GridTeam — one cooperative grid per world (grid-wide barrier between phases): one large world on all 148 SMs.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "f2d_launch.h"
#include "f2d_step.h"

namespace cg = cooperative_groups;

namespace f2d
{

// ------------------------------------------------------------------------------------------------ device teams
// A team of threads inside one thread block: the whole block (one world per block), or one of several equal slices
// of a block that steps several worlds side by side (stepWorldsGang). Slices are whole warps; every team has its own
// hardware barrier (bar.sync id, count), its own shared scratch, and a second barrier for its crew.
//
// The crew = the team minus its last warp: lets that warp run a serial side task (island split) concurrently with the
// barrier-separated solver stages.
// In-place exclusive scan of data[0..n) by a slice of a thread block (whole warps, `scratch` = 64 ints of shared memory
// owned by the slice); returns the total. Collective over the slice, ONE barrier per chunk of size() items: every warp
// publishes its sum, and after the barrier every warp scans the published sums for itself (the sums of two successive
// chunks live in different halves of the scratch, so a warp that runs ahead cannot overwrite what a slower one still reads).
template <class Slice> __device__ __forceinline__ int sliceExclusiveScan( const Slice& t, int32_t* scratch, int32_t* data, int n )
{
	const int nt = t.size(), tid = t.rank();
	const int lane = tid & 31, warp = tid >> 5, warps = nt >> 5;
	int carry = 0;
	int parity = 0;
	for ( int base = 0; base < n; base += nt, parity ^= 32 )
	{
		int i = base + tid;
		int v = i < n ? data[i] : 0;
		// warp inclusive scan
		int x = v;
		for ( int off = 1; off < 32; off <<= 1 )
		{
			int y = __shfl_up_sync( 0xffffffffu, x, off );
			if ( lane >= off )
				x += y;
		}
		int32_t* warpSums = scratch + parity;
		if ( lane == 31 )
			warpSums[warp] = x;
		t.sync();
		// sums of the warps before this one, and of all of them
		int ws = lane < warps ? warpSums[lane] : 0;
		int before = lane < warp ? ws : 0;
		for ( int off = 16; off > 0; off >>= 1 )
		{
			ws += __shfl_xor_sync( 0xffffffffu, ws, off );
			before += __shfl_xor_sync( 0xffffffffu, before, off );
		}
		if ( i < n )
			data[i] = carry + before + ( x - v );
		carry += ws;
	}
	// (the next collective of the slice starts with a barrier of its own before anybody reads data[])
	t.sync();
	return carry;
}

// The first four warps of a thread block as a team of their own (hardware barrier 3): the serial parts of the step and the
// little parallel work between them, while the rear rebuilds the trees (stepCollide).
struct CtaFront
{
	// (as in the solo block of a grid, the warps beside the two serial threads read ahead of them: their L1 is the
	// block's, and the narrowphase has just swept it)
	static constexpr bool kHasSoloBlock = true;
	static constexpr int kThreads = 128;
	int32_t* scratch;
	typedef WarpLanes Lanes;
	__device__ bool inSoloBlock() const { return true; }
	__device__ int soloSize() const { return kThreads; }
	__device__ int rank() const { return (int)threadIdx.x; }
	__device__ int size() const { return kThreads; }
	__device__ void sync() const { asm volatile( "bar.sync 3, %0;" ::"n"( kThreads ) : "memory" ); }
	__device__ int exclusiveScan( int32_t* data, int n ) const { return sliceExclusiveScan( *this, scratch, data, n ); }
};

// Block 0 of a cooperative grid as a team of its own: the same role as CtaFront.
struct GridFront
{
	static constexpr bool kHasSoloBlock = true;
	int32_t* scratch;
	typedef WarpLanes Lanes;
	__device__ bool inSoloBlock() const { return true; }
	__device__ int soloSize() const { return (int)blockDim.x; }
	__device__ int rank() const { return (int)threadIdx.x; }
	__device__ int size() const { return (int)blockDim.x; }
	__device__ void sync() const { __syncthreads(); }
	__device__ int exclusiveScan( int32_t* data, int n ) const { return sliceExclusiveScan( *this, scratch, data, n ); }
};

// All warps of a thread block but the front's, as a team of their own (hardware barrier 2): see CtaTeamT::rear.
struct CtaRear
{
	int32_t* scratch;
	int32_t* arena; // the block's shared-memory work area (CtaTeamT::arena), free while the rear is on its own
	int arenaInts;
	typedef WarpLanes Lanes;
	__device__ int32_t* arenaPtr() const { return arena; }
	__device__ int arenaSize() const { return arenaInts; }
	__device__ int rank() const { return (int)threadIdx.x - CtaFront::kThreads; }
	__device__ int size() const { return (int)blockDim.x - CtaFront::kThreads; }
	__device__ void sync() const { asm volatile( "bar.sync 2, %0;" ::"r"( (int)blockDim.x - CtaFront::kThreads ) : "memory" ); }
	__device__ int exclusiveScan( int32_t* data, int n ) const { return sliceExclusiveScan( *this, scratch, data, n ); }
};

template <bool kWholeBlock> struct CtaCrewT
{
	int n, tid, barId;
	int32_t* arena;
	int arenaInts;
	__device__ int32_t* arenaPtr() const { return arena; }
	__device__ int arenaSize() const { return arenaInts; }
	// (a whole-block team reads its rank and size from the special registers, as the round-1 kernels did: held in the
	// struct they cost registers the solver stages do not have at 64 per thread)
	__device__ int groupCount() const { return size() >> 5; }
	__device__ int groupIndex() const { return rank() >> 5; }
	__device__ int lane() const { return (int)( threadIdx.x & 31 ); }
	__device__ int groupSize() const { return 32; }
	__device__ void groupSync() const { __syncwarp(); }
	__device__ int rank() const
	{
		if constexpr ( kWholeBlock )
			return (int)threadIdx.x;
		else
			return tid;
	}
	__device__ int size() const
	{
		if constexpr ( kWholeBlock )
			return (int)blockDim.x - 32;
		else
			return n;
	}
	__device__ void sync() const
	{
		if constexpr ( kWholeBlock )
			asm volatile( "bar.sync 1, %0;" ::"r"( (int)blockDim.x - 32 ) : "memory" ); // the crew of a whole-block team: barrier 1
		else
			asm volatile( "bar.sync %0, %1;" ::"r"( barId ), "r"( n ) : "memory" );
	}
};

// kWholeBlock: the team is the whole thread block (barrier 0 = __syncthreads, which the compiler knows); otherwise a
// slice of the block with barrier ids held in registers. kLarge: a whole-block team of 512 threads or more (one world
// that has an SM to itself), which can put its rear on a task of its own.
template <bool kWholeBlock, bool kLarge = false> struct CtaTeamT
{
	static constexpr bool kHasSoloBlock = false;
	static constexpr bool kCanFork = true;
	int32_t* smem; // 64 ints of shared scratch owned by the team
	int tid, nthreads; // rank inside the team, threads of the team
	int barId, crewBarId; // hardware barriers of the team and of its crew (0 = the block's own barrier)
	// Dynamic shared memory the block may use as a work area (0 in batches, where the resident worlds of an SM live on
	// its L1; a single small world has the SM - and its 228 KB - to itself)
	int32_t* arena = nullptr;
	int arenaInts = 0;
	// the whole block as one team
	static __device__ CtaTeamT block( int32_t* smem ) { return CtaTeamT{ smem, (int)threadIdx.x, (int)blockDim.x, 0, 1 }; }
	// Forking pays when the crew keeps most of the team: with two warps the solver would run on one while the other
	// walks (measured: the walking warp then waits ~2x its walk for the crew), so two-warp teams walk first, then solve;
	// from four warps up the crew of three or more hides the walk (measured at 128 threads: 8 % of warp time idle otherwise).
	__device__ bool canFork() const { return size() >= 128; }
	__device__ bool inSide() const { return rank() >= size() - 32; }
	__device__ bool isSideLeader() const { return rank() == size() - 32; }
	typedef WarpLanes Lanes;
	__device__ bool inSideGroup() const { return rank() >= size() - 32; }
	__device__ bool inFirstGroup() const { return rank() < 32; }
	__device__ CtaCrewT<kWholeBlock> crew() const { return CtaCrewT<kWholeBlock>{ size() - 32, rank(), crewBarId, arena, arenaInts }; }
	__device__ int groupCount() const { return size() >> 5; }
	__device__ int groupIndex() const { return rank() >> 5; }
	__device__ int lane() const { return (int)( threadIdx.x & 31 ); }
	__device__ int groupSize() const { return 32; }
	__device__ void groupSync() const { __syncwarp(); }
	__device__ int32_t* arenaPtr() const { return arena; }
	__device__ int arenaSize() const { return arenaInts; }
	__device__ int rank() const
	{
		if constexpr ( kWholeBlock )
			return (int)threadIdx.x;
		else
			return tid;
	}
	__device__ int size() const
	{
		if constexpr ( kWholeBlock )
			return (int)blockDim.x;
		else
			return nthreads;
	}
	__device__ void sync() const
	{
		if constexpr ( kWholeBlock )
			__syncthreads();
		else
			asm volatile( "bar.sync %0, %1;" ::"r"( barId ), "r"( nthreads ) : "memory" );
	}
	// in-place exclusive scan of data[0..n) in global memory; returns the total. Team-wide collective.
	__device__ int exclusiveScan( int32_t* data, int n ) const { return sliceExclusiveScan( *this, smem, data, n ); }
	// The rear of a large whole-block team: everybody but the first two warps, with a barrier and scan scratch of its own
	// (stepCollide: the rear rebuilds the trees while two threads of the front apply the ordered contact-state changes)
	static constexpr bool kCanSplitTree = kWholeBlock && kLarge;
	__device__ bool canSplitTree() const { return true; }
	__device__ bool inFront() const { return threadIdx.x < CtaFront::kThreads; }
	__device__ CtaRear rear() const { return CtaRear{ smem + 64, arena, arenaInts }; }
	__device__ CtaFront front() const { return CtaFront{ smem }; }
};
typedef CtaTeamT<true> CtaTeam;	 // one world per block (and the solo block of a grid)
typedef CtaTeamT<true, true> LargeCtaTeam; // one world per block of >= 512 threads
typedef CtaTeamT<false> GangTeam; // one of several worlds of a block (stepWorldsGang)

// Grid-wide barrier state in global memory: a monotonically increasing arrival counter (wraps harmlessly) and the
// count it had when the previous kernel on the stream finished.
struct GridBarrier
{
	unsigned int arrivals;
	unsigned int pad0[31];
	unsigned int base;
	unsigned int pad1[31];
};

__device__ __forceinline__ unsigned int loadAcquire( const unsigned int* p )
{
	unsigned int v;
	asm volatile( "ld.acquire.gpu.global.u32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
	return v;
}

// One cooperative grid per world, one block per SM. The barrier is hand-rolled (one atomic arrival per block, thread 0
// spins on an acquire load, gpu-scope fences on both sides so the other SMs' writes are visible and this SM's L1 is
// invalidated): ~0.5 us instead of the ~2.2 us measured for cooperative_groups' grid.sync() on 148 blocks.
__device__ __forceinline__ void gridBarrierWait( GridBarrier* barrier, unsigned int& gen, unsigned int parts )
{
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		gen += parts;
		// release: the block's writes (ordered before this by the bar.sync above) become visible gpu-wide with the
		// arrival; acquire: the spin load orders the other blocks' writes before everything after the bar.sync below
		asm volatile( "red.release.gpu.global.add.u32 [%0], 1;" ::"l"( &barrier->arrivals ) : "memory" );
		while ( (int)( loadAcquire( &barrier->arrivals ) - gen ) < 0 )
		{
		}
	}
	__syncthreads();
}

// All blocks of the grid but the last: see CtaCrew.
struct GridCrew
{
	GridBarrier* barrier;
	unsigned int* gen;
	__device__ int groupCount() const { return (int)( ( ( gridDim.x - 1 ) * blockDim.x ) >> 5 ); }
	__device__ int groupIndex() const { return (int)( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 ); }
	__device__ int lane() const { return (int)( threadIdx.x & 31 ); }
	__device__ int groupSize() const { return 32; }
	__device__ void groupSync() const { __syncwarp(); }
	__device__ int rank() const { return (int)( blockIdx.x * blockDim.x + threadIdx.x ); }
	__device__ int size() const { return (int)( ( gridDim.x - 1 ) * blockDim.x ); }
	__device__ void sync() const { gridBarrierWait( barrier, *gen, gridDim.x - 1 ); }
};

// In-place exclusive scan by the blocks [firstBlock, firstBlock + nb) of a grid: each block scans one contiguous tile,
// the tile offsets are added after a barrier of those blocks (`sync`). Returns the total.
template <class Sync>
__device__ __forceinline__ int gridSliceScan( int firstBlock, int nb, int32_t* smem, int32_t* blockTotals, Sync&& sync, int32_t* data, int n )
{
	const int b0 = (int)blockIdx.x - firstBlock;
	const int tile = ( n + nb - 1 ) / nb;
	const int begin = min( n, b0 * tile );
	const int end = min( n, begin + tile );
	CtaTeam cta = CtaTeam::block( smem );
	int total = cta.exclusiveScan( data + begin, end - begin );
	if ( threadIdx.x == 0 )
		blockTotals[b0] = total;
	sync();
	int offset = 0, sum = 0;
	for ( int b = 0; b < nb; ++b )
	{
		int v = blockTotals[b];
		if ( b < b0 )
			offset += v;
		sum += v;
	}
	for ( int i = begin + (int)threadIdx.x; i < end; i += (int)blockDim.x )
		data[i] += offset;
	sync();
	return sum;
}

// All blocks of the grid but the first, as a team of their own (grid barrier 2): rebuilds the trees while block 0 applies
// the ordered contact-state changes (stepCollide).
struct GridRear
{
	GridBarrier* barrier;
	unsigned int* gen;
	int32_t* smem;
	int32_t* blockTotals;
	typedef WarpLanes Lanes;
	__device__ int32_t* arenaPtr() const { return nullptr; }
	__device__ int arenaSize() const { return 0; }
	__device__ int rank() const { return (int)( ( blockIdx.x - 1 ) * blockDim.x + threadIdx.x ); }
	__device__ int size() const { return (int)( ( gridDim.x - 1 ) * blockDim.x ); }
	__device__ void sync() const { gridBarrierWait( barrier, *gen, gridDim.x - 1 ); }
	__device__ int exclusiveScan( int32_t* data, int n ) const
	{
		return gridSliceScan( 1, (int)gridDim.x - 1, smem, blockTotals, [&]() { sync(); }, data, n );
	}
};

struct GridTeam
{
	static constexpr bool kHasSoloBlock = true;
	static constexpr bool kCanFork = true;
	static constexpr bool kCanSplitTree = true;
	unsigned int crewGen, rearGen;
	__device__ bool canSplitTree() const { return gridDim.x >= 2; }
	__device__ bool inFront() const { return blockIdx.x == 0; }
	__device__ GridRear rear() { return GridRear{ barrier + 2, &rearGen, smem, blockTotals }; }
	__device__ GridFront front() const { return GridFront{ smem }; }
	__device__ bool canFork() const { return gridDim.x >= 2; }
	__device__ bool inSide() const { return blockIdx.x == gridDim.x - 1; }
	__device__ bool isSideLeader() const { return blockIdx.x == gridDim.x - 1 && threadIdx.x == 0; }
	typedef WarpLanes Lanes;
	__device__ bool inSideGroup() const { return blockIdx.x == gridDim.x - 1 && threadIdx.x < 32; }
	__device__ bool inFirstGroup() const { return blockIdx.x == 0 && threadIdx.x < 32; }
	__device__ GridCrew crew() { return GridCrew{ barrier + 1, &crewGen }; }
	__device__ int groupCount() const { return (int)( ( gridDim.x * blockDim.x ) >> 5 ); }
	__device__ int groupIndex() const { return (int)( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 ); }
	__device__ int lane() const { return (int)( threadIdx.x & 31 ); }
	__device__ int groupSize() const { return 32; }
	__device__ void groupSync() const { __syncwarp(); }
	int32_t* smem;
	int32_t* blockTotals; // gridDim.x ints in global memory
	GridBarrier* barrier;
	unsigned int gen; // arrivals expected once the next barrier completes (meaningful in thread 0)
	__device__ int32_t* arenaPtr() const { return nullptr; }
	__device__ int arenaSize() const { return 0; }
	__device__ int rank() const { return (int)( blockIdx.x * blockDim.x + threadIdx.x ); }
	__device__ int size() const { return (int)( gridDim.x * blockDim.x ); }
	__device__ void begin()
	{
		gen = 0;
		crewGen = 0;
		rearGen = 0;
		if ( threadIdx.x == 0 )
		{
			gen = *reinterpret_cast<volatile unsigned int*>( &barrier[0].base );
			crewGen = *reinterpret_cast<volatile unsigned int*>( &barrier[1].base );
			rearGen = *reinterpret_cast<volatile unsigned int*>( &barrier[2].base );
		}
	}
	__device__ void end()
	{
		// every block executed the same number of barriers; block 0 publishes the counts for the next launch
		if ( blockIdx.x == 0 && threadIdx.x == 0 )
		{
			*reinterpret_cast<volatile unsigned int*>( &barrier[0].base ) = gen;
			*reinterpret_cast<volatile unsigned int*>( &barrier[1].base ) = crewGen;
		}
		// (block 0 is not part of the rear)
		if ( blockIdx.x == 1 && threadIdx.x == 0 )
			*reinterpret_cast<volatile unsigned int*>( &barrier[2].base ) = rearGen;
	}
	__device__ void sync() { gridBarrierWait( barrier, gen, gridDim.x ); }
	// the tree rebuild runs on block 0 alone (block-level barriers) while the other blocks do the narrowphase
	__device__ bool inSoloBlock() const { return blockIdx.x == 0; }
	__device__ int soloSize() const { return (int)blockDim.x; }
	__device__ bool hasOutsideSolo() const { return gridDim.x > 1; }
	__device__ CtaTeam soloTeam() const { return CtaTeam::block( smem ); }
	__device__ int rankOutsideSolo() const { return (int)( ( blockIdx.x - 1 ) * blockDim.x + threadIdx.x ); }
	__device__ int sizeOutsideSolo() const { return (int)( ( gridDim.x - 1 ) * blockDim.x ); }
	__device__ int exclusiveScan( int32_t* data, int n )
	{
		return gridSliceScan( 0, (int)gridDim.x, smem, blockTotals, [&]() { sync(); }, data, n );
	}
};

// ------------------------------------------------------------------------------------------------ kernels

template <class Team> __device__ __forceinline__ void runPhase( World* w, Team& t, int phase, float dt, int sub )
{
	stepWorldPhase( w, t, phase, dt, sub );
}

// One thread block per world; worlds are `stride` bytes apart. Grid-stride over worlds. The block steps a copy of the
// World header held in shared memory (see World::deviceBase) and writes it back when the world is done.
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__( kThreads, kMinBlocks )
	stepWorldsCta( char* base, unsigned long long stride, int worldCount, float dt, int sub, int phase, int steps, int arenaBytes,
				   uint4* hostHeader )
{
	typedef CtaTeamT<true, ( kThreads >= 512 )> Team;
	__shared__ int32_t smem[128]; // scan scratch of the team, and of its rear (CtaRear)
	__shared__ uint4 header[sizeof( World ) / 16];
	extern __shared__ int4 dynamicShared[];
	Team team = Team::block( smem );
	if ( arenaBytes > 0 )
	{
		team.arena = reinterpret_cast<int32_t*>( dynamicShared );
		team.arenaInts = arenaBytes / 4;
	}
	World* w = reinterpret_cast<World*>( header );
	for ( int wi = (int)blockIdx.x; wi < worldCount; wi += (int)gridDim.x )
	{
		uint4* image = reinterpret_cast<uint4*>( base + (unsigned long long)wi * stride );
		for ( int i = (int)threadIdx.x; i < (int)( sizeof( World ) / 16 ); i += kThreads )
			header[i] = image[i];
		__syncthreads();
		if ( threadIdx.x == 0 )
			w->deviceBase = reinterpret_cast<uint64_t>( image );
		__syncthreads();
		// steps < 0: only the worlds that stopped for more contact room (kErrRetry) take (one) step - the batch has grown
		// their images in between (f2dBatch growth path)
		const bool onlyRetry = steps < 0;
		if ( onlyRetry == false || ( w->error & kErrRetry ) != 0 )
		{
			if ( onlyRetry && threadIdx.x == 0 )
			{
				w->error &= ~kErrRetry;
				w->step.retryContacts = 0;
			}
			__syncthreads();
			for ( int s = 0; s < ( onlyRetry ? 1 : steps ); ++s )
			{
				const bool fatal = ( w->error & kErrFatal ) != 0;
				__syncthreads(); // every thread has read the flags before rank 0 touches them again (stepBegin)
				if ( fatal )
					break;
				runPhase( w, team, phase, dt, sub );
			}
		}
		__syncthreads();
		for ( int i = (int)threadIdx.x; i < (int)( sizeof( World ) / 16 ); i += kThreads )
			image[i] = header[i];
		// A single world stepped synchronously: the header (counters, error flags, event counts) also goes straight into
		// the pinned host image over PCIe - 2 KB of posted writes instead of a D2H copy operation after the kernel.
		if ( hostHeader != nullptr )
		{
			for ( int i = (int)threadIdx.x; i < (int)( sizeof( World ) / 16 ); i += kThreads )
				hostHeader[i] = header[i];
		}
		__syncthreads();
	}
}

// Several worlds per thread block, side by side: kTeams teams of kTeamThreads threads, one world each, kept within one
// coarse phase of the step of each other. The worlds of a block therefore run the SAME phase (or two adjacent ones),
// and the SM's instruction caches hold one phase's code instead of the whole step's: with independent one-world blocks
// the eight resident worlds of an SM drift into different phases (real batches are not in lock-step), and the
// step - ~3 MB of SASS - no longer fits: ncu shows as many stall cycles waiting for instructions as waiting for memory
// (profiles/README.md, r03m). Blocks are persistent and take gangs of worlds from a queue.
template <int kTeamThreads, int kTeams>
__global__ void __launch_bounds__( kTeamThreads* kTeams, 1 )
	stepWorldsGang( char* base, unsigned long long stride, int worldCount, float dt, int sub, int* queue, int onlyRetry )
{
	__shared__ int32_t scratch[kTeams][64];
	__shared__ uint4 headers[kTeams][sizeof( World ) / 16];
	__shared__ int gangFirst;
	__shared__ int32_t phaseDone[kTeams]; // phases each team has finished (of the current gang)
	const int teamIndex = (int)threadIdx.x / kTeamThreads, tid = (int)threadIdx.x % kTeamThreads;
	// barrier 0: the block; 1 .. kTeams: the teams; kTeams + 1 .. 2 kTeams: their crews (16 hardware barriers per block)
	// (teams under 128 threads never fork: CtaTeam::canFork)
	static_assert( ( kTeamThreads >= 128 ? 2 : 1 ) * kTeams + 1 <= 16, "one hardware barrier per team and per crew, plus the block's" );
	GangTeam team{ scratch[teamIndex], tid, kTeamThreads, 1 + teamIndex, 1 + kTeams + teamIndex };
	World* w = reinterpret_cast<World*>( headers[teamIndex] );
	while ( true )
	{
		__syncthreads();
		if ( threadIdx.x == 0 )
			gangFirst = atomicAdd( queue, kTeams );
		__syncthreads();
		const int wi = gangFirst + teamIndex;
		if ( gangFirst >= worldCount )
			break;
		const bool have = wi < worldCount;
		uint4* image = reinterpret_cast<uint4*>( base + (unsigned long long)( have ? wi : 0 ) * stride );
		bool active = false;
		if ( have )
		{
			for ( int i = tid; i < (int)( sizeof( World ) / 16 ); i += kTeamThreads )
				headers[teamIndex][i] = image[i];
			team.sync();
			if ( tid == 0 )
				w->deviceBase = reinterpret_cast<uint64_t>( image );
			team.sync();
			// onlyRetry: only the worlds that stopped for more contact room take the step (the batch has grown the images)
			const uint32_t error = w->error;
			active = ( error & kErrFatal ) == 0 && ( onlyRetry == 0 || ( error & kErrRetry ) != 0 );
			team.sync();
			if ( active && onlyRetry && tid == 0 )
			{
				w->error &= ~kErrRetry;
				w->step.retryContacts = 0;
			}
			team.sync();
		}
		// Phase alignment with one phase of slack: a team may enter phase q as soon as EVERY team has finished phase
		// q - 2, so at most two adjacent phases run on the SM at a time (their code fits the instruction caches) and a
		// world that is slow in one phase - an island split, a continuous-collision pass - only holds the others up when it
		// falls a whole phase behind. (Measured on 8192 decorrelated bench2d worlds: strict barrier per phase 23.9 ms per
		// step, one phase of slack 23.5, two phases 23.9.)
		const int phases[] = { kPhaseBeginPairs, kPhaseCollideTreeOnly, kPhaseCollideNarrowOnly, kPhaseCollideFinish, kPhaseSolve,
							   kPhaseFinalize };
		constexpr int kPhaseCount = (int)( sizeof( phases ) / sizeof( phases[0] ) );
		if ( tid == 0 )
			phaseDone[teamIndex] = active ? 0 : kPhaseCount;
		__syncthreads();
		for ( int q = 0; q < kPhaseCount; ++q )
		{
			if ( active == false )
				break;
			if ( q >= 2 )
			{
				if ( tid == 0 )
				{
					bool ready = false;
					while ( ready == false )
					{
						ready = true;
						for ( int k = 0; k < kTeams; ++k )
							ready = ready && atomicAdd( &phaseDone[k], 0 ) >= q - 1; // (atomic read of the shared flag)
					}
				}
				team.sync();
			}
			stepWorldPhase( w, team, phases[q], dt, sub );
			team.sync();
			if ( tid == 0 )
			{
				__threadfence_block();
				atomicExch( &phaseDone[teamIndex], q + 1 );
			}
		}
		__syncthreads();
		if ( have )
		{
			for ( int i = tid; i < (int)( sizeof( World ) / 16 ); i += kTeamThreads )
				image[i] = headers[teamIndex][i];
		}
	}
}

// One cooperative grid per world.
template <int kThreads>
__global__ void __launch_bounds__( kThreads, 1 )
	stepWorldGrid( World* w, int32_t* blockTotals, float dt, int sub, int phase, uint4* hostHeader )
{
	__shared__ int32_t smem[64];
	GridTeam team{ 0u, 0u, smem, blockTotals + 192, reinterpret_cast<GridBarrier*>( blockTotals ), 0u }; // three barriers, then the scan totals
	// the grid shares the header in HBM; every block stores the same address, so each may rely on it after its own barrier
	if ( threadIdx.x == 0 )
		w->deviceBase = reinterpret_cast<uint64_t>( w );
	__syncthreads();
	if ( w->error & kErrFatal )
		return;
	team.begin();
	runPhase( w, team, phase, dt, sub );
	if ( hostHeader != nullptr )
	{
		// header -> pinned host image (see stepWorldsCta), by block 0 once every block is done with the step
		team.sync();
		if ( blockIdx.x == 0 )
		{
			const uint4* header = reinterpret_cast<const uint4*>( w );
			for ( int i = (int)threadIdx.x; i < (int)( sizeof( World ) / 16 ); i += kThreads )
				hostHeader[i] = __ldcg( header + i );
		}
	}
	team.end();
}

} // namespace f2d
