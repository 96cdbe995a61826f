// Micro-benchmark (development aid): what does a global store do to a line that sits in the L1 of the storing SM?
// One thread: load (fill L1), store, load again, timed with clock64. Also: a load after another thread's store + bar.sync.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_store_probe l1_store_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe( int* data, long long* out )
{
	__shared__ int sink;
	int* p = data + 4096;
	if ( threadIdx.x == 0 )
	{
		long long t0 = clock64();
		int a = *p; // miss: L2 or DRAM
		sink = a;
		if ( sink == 123456789 )
			out[7] = 1; // the clock read below waits for the load
		long long t1 = clock64();
		int b = p[1] + sink; // same sector: L1 hit
		sink = b;
		if ( sink == 123456789 )
			out[7] = 1; // the clock read below waits for the load
		long long t2 = clock64();
		p[2] = b + 1; // store to the cached line
		__threadfence_block();
		long long t3 = clock64();
		int c = p[2] + sink; // after own store
		sink = c;
		if ( sink == 123456789 )
			out[7] = 1; // the clock read below waits for the load
		long long t4 = clock64();
		out[0] = t1 - t0;
		out[1] = t2 - t1;
		out[2] = t4 - t3;
	}
	__syncthreads();
	// other thread stores, bar.sync, thread 0 loads
	int* q = data + 8192;
	if ( threadIdx.x == 0 )
		sink = q[0]; // fill
	__syncthreads();
	if ( threadIdx.x == 64 )
		q[1] = 7;
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		long long t0 = clock64();
		int c = q[1] + sink;
		sink = c;
		if ( sink == 123456789 )
			out[7] = 1; // the clock read below waits for the load
		long long t1 = clock64();
		out[3] = t1 - t0;
		out[5] = c;
	}
	__syncthreads();
	// store to a line that was never loaded, then load it (write-allocate?)
	int* r = data + 16384;
	if ( threadIdx.x == 64 )
		r[0] = 9;
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		long long t0 = clock64();
		int c = r[0] + sink;
		sink = c;
		if ( sink == 123456789 )
			out[7] = 1; // the clock read below waits for the load
		long long t1 = clock64();
		out[4] = t1 - t0;
		out[6] = c;
	}
}

int main()
{
	int* data;
	long long* out;
	cudaMalloc( &data, 1 << 20 );
	cudaMemset( data, 0, 1 << 20 );
	cudaMallocManaged( &out, 64 );
	for ( int rep = 0; rep < 3; ++rep )
	{
		probe<<<1, 128>>>( data, out );
		cudaDeviceSynchronize();
		printf( "first load %lld, L1 hit %lld, load after own store %lld, load after another thread's store + bar %lld, load of a stored (never loaded) line %lld cycles\n",
				out[0], out[1], out[2], out[3], out[4] );
	}
	return 0;
}
