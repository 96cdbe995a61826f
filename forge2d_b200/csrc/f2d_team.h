// forge2d_b200 — execution-team abstraction.
//
// Every phase of the world step is written once as a team-parallel routine: `for (i = team.rank(); i < n; i +=
// team.size())` loops separated by `team.sync()`, with the reference's order-defining serial sections run by rank 0.
// Three teams execute the same code:
//   * CtaTeam   — one CUDA thread block per world (`__syncthreads`): batches of independent worlds, small worlds;
//   * GridTeam  — one cooperative grid per world (grid-wide barrier): a single large world across all 148 SMs;
//   * SerialTeam — one host thread; used by the host-side API paths and by the CPU emulation build in tests/.
#pragma once
#include "f2d_types.h"

namespace f2d
{

#if defined( __CUDA_ARCH__ )
F2D_HD void atomOr64( uint64_t* p, uint64_t v ) { atomicOr( reinterpret_cast<unsigned long long*>( p ), (unsigned long long)v ); }
F2D_HD int atomAdd( int32_t* p, int v ) { return atomicAdd( p, v ); }
F2D_HD void atomMax64( unsigned long long* p, unsigned long long v ) { atomicMax( p, v ); }
F2D_HD void atomOr32( uint32_t* p, uint32_t v ) { atomicOr( p, v ); }
F2D_HD void atomMax32( int32_t* p, int32_t v ) { atomicMax( p, v ); }
// float min/max through the ordered-int trick (valid for any mix of signs, NaN-free inputs)
F2D_HD void atomMinF( float* p, float v )
{
	v += 0.0f; // -0.0 -> +0.0: as an int, -0.0 would beat every negative value
	if ( v >= 0.0f )
		atomicMin( reinterpret_cast<int*>( p ), __float_as_int( v ) );
	else
		atomicMax( reinterpret_cast<unsigned int*>( p ), __float_as_uint( v ) );
}
F2D_HD void atomMaxF( float* p, float v )
{
	if ( v >= 0.0f )
		atomicMax( reinterpret_cast<int*>( p ), __float_as_int( v ) );
	else
		atomicMin( reinterpret_cast<unsigned int*>( p ), __float_as_uint( v ) );
}
#else
F2D_HD void atomOr64( uint64_t* p, uint64_t v ) { *p |= v; }
F2D_HD int atomAdd( int32_t* p, int v )
{
	int o = *p;
	*p += v;
	return o;
}
F2D_HD void atomMax64( unsigned long long* p, unsigned long long v )
{
	if ( v > *p )
		*p = v;
}
F2D_HD void atomOr32( uint32_t* p, uint32_t v ) { *p |= v; }
F2D_HD void atomMax32( int32_t* p, int32_t v )
{
	if ( v > *p )
		*p = v;
}
F2D_HD void atomMinF( float* p, float v )
{
	if ( v < *p )
		*p = v;
}
F2D_HD void atomMaxF( float* p, float v )
{
	if ( *p < v )
		*p = v;
}
#endif

F2D_HD int loadVolatile( const int32_t* p ) { return *reinterpret_cast<const volatile int32_t*>( p ); }
F2D_HD void storeVolatile( int32_t* p, int v ) { *reinterpret_cast<volatile int32_t*>( p ) = v; }

#if defined( __CUDA_ARCH__ )
F2D_HD uint64_t profClock()
{
	unsigned long long t;
	asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t ) );
	return t;
}
#else
F2D_HD uint64_t profClock() { return 0; }
#endif

// Phase mark: rank 0 charges the time since the previous mark to `slot`. Call right after a team.sync().
#define F2D_MARK( w, t, slot )                                                                                                 \
	do                                                                                                                         \
	{                                                                                                                          \
		if ( ( t ).rank() == 0 && ( w )->profEnabled )                                                                         \
		{                                                                                                                      \
			uint64_t now_ = ::f2d::profClock();                                                                                \
			( w )->prof[slot] += now_ - ( w )->profLast;                                                                       \
			( w )->profLast = now_;                                                                                            \
		}                                                                                                                      \
	} while ( 0 )

F2D_HD int popCount32( uint32_t x )
{
#if defined( __CUDA_ARCH__ )
	return __popc( x );
#else
	int n = 0;
	while ( x != 0 )
	{
		x &= x - 1;
		n += 1;
	}
	return n;
#endif
}
// leading zero bits (x != 0)
F2D_HD int countLeadingZeros32( uint32_t x )
{
#if defined( __CUDA_ARCH__ )
	return __clz( (int)x );
#else
	return __builtin_clz( x );
#endif
}
// index of the lowest set bit (x != 0)
F2D_HD int lowestBit32( uint32_t x )
{
#if defined( __CUDA_ARCH__ )
	return __ffs( (int)x ) - 1;
#else
	int n = 0;
	while ( ( x & 1u ) == 0 )
	{
		x >>= 1;
		n += 1;
	}
	return n;
#endif
}

// index of the lowest set bit (x != 0)
F2D_HD int lowestBit64( uint64_t x )
{
#if defined( __CUDA_ARCH__ )
	return __ffsll( (long long)x ) - 1;
#else
	return __builtin_ctzll( x );
#endif
}

// The lanes that walk a graph together (island split): one host thread, or the 32 lanes of a warp.
struct SoloLanes
{
	F2D_HD int lane() const { return 0; }
	F2D_HD int count() const { return 1; }
	F2D_HD uint32_t ballot( bool p ) const { return p ? 1u : 0u; }
	F2D_HD uint32_t matchAny( int ) const { return 1u; }
	F2D_HD bool firstOfEqual( bool active, int ) const { return active; }
	F2D_HD void sync() const {}
	F2D_HD int broadcast( int v ) const { return v; }
	F2D_HD int from( int v, int ) const { return v; }
	F2D_HD int reduceAdd( int v ) const { return v; }
	F2D_HD float reduceMin( float v ) const { return v; }
	F2D_HD float reduceMax( float v ) const { return v; }
	F2D_HD void fence() const {}
};
#if defined( __CUDA_ARCH__ )
struct WarpLanes
{
	F2D_HD int lane() const { return (int)( threadIdx.x & 31 ); }
	F2D_HD int count() const { return 32; }
	F2D_HD uint32_t ballot( bool p ) const { return __ballot_sync( 0xffffffffu, p ); }
	F2D_HD uint32_t matchAny( int key ) const { return __match_any_sync( 0xffffffffu, key ); }
	// Is this lane the lowest of the active lanes that hold `key`? One shuffle and one ballot per DISTINCT key among the
	// active lanes (usually one or two) - match.any costs a pass per distinct value over all 32 lanes, active or not.
	F2D_HD bool firstOfEqual( bool active, int key ) const
	{
		uint32_t pending = __ballot_sync( 0xffffffffu, active );
		bool first = false;
		while ( pending != 0 )
		{
			const int leader = __ffs( (int)pending ) - 1;
			const int leaderKey = __shfl_sync( 0xffffffffu, key, leader );
			const uint32_t same = __ballot_sync( 0xffffffffu, active && key == leaderKey );
			first = first || (int)( threadIdx.x & 31 ) == leader;
			pending &= ~same;
		}
		return first;
	}
	F2D_HD void sync() const { __syncwarp(); }
	F2D_HD int broadcast( int v ) const { return __shfl_sync( 0xffffffffu, v, 0 ); }
	F2D_HD int from( int v, int sourceLane ) const { return __shfl_sync( 0xffffffffu, v, sourceLane ); } // every lane gets sourceLane's v
	F2D_HD int reduceAdd( int v ) const
	{
		for ( int off = 16; off > 0; off >>= 1 )
			v += __shfl_xor_sync( 0xffffffffu, v, off );
		return v;
	}
	// min / max in the reference's `a < b ? a : b` / `a > b ? a : b` forms (every lane ends with the same value)
	F2D_HD float reduceMin( float v ) const
	{
		for ( int off = 16; off > 0; off >>= 1 )
		{
			float o = __shfl_xor_sync( 0xffffffffu, v, off );
			v = o < v ? o : v;
		}
		return v;
	}
	F2D_HD float reduceMax( float v ) const
	{
		for ( int off = 16; off > 0; off >>= 1 )
		{
			float o = __shfl_xor_sync( 0xffffffffu, v, off );
			v = o > v ? o : v;
		}
		return v;
	}
	F2D_HD void fence() const { __threadfence(); }
};
#else
typedef SoloLanes WarpLanes;
#endif

struct SerialTeam
{
	typedef SoloLanes Lanes;
	F2D_HD bool inFirstGroup() const { return true; }
	F2D_HD int32_t* arenaPtr() const { return nullptr; } // shared-memory work area of the team (device block teams only)
	F2D_HD int arenaSize() const { return 0; }
	static constexpr bool kHasSoloBlock = false;
	static constexpr bool kCanFork = false;
	static constexpr bool kCanSplitTree = false;
	F2D_HD int rank() const { return 0; }
	F2D_HD int size() const { return 1; }
	F2D_HD void sync() const {}
	// groups: independent sub-teams that only need to synchronise among themselves (warps on the device)
	F2D_HD int groupCount() const { return 1; }
	F2D_HD int groupIndex() const { return 0; }
	F2D_HD int lane() const { return 0; }
	F2D_HD int groupSize() const { return 1; }
	F2D_HD void groupSync() const {}
	// in-place exclusive scan of data[0..n), returns the total
	F2D_HD int exclusiveScan( int32_t* data, int n ) const
	{
		int sum = 0;
		for ( int i = 0; i < n; ++i )
		{
			int v = data[i];
			data[i] = sum;
			sum += v;
		}
		return sum;
	}
};

} // namespace f2d
