// fetch_probe.cu -- how many DRAM sectors does ONE random 32-byte index-block load cost on this device?
// ncu on k_sa showed ~4 L2 sectors looked up (and filled from DRAM) per 32-byte request (profiles/r2_k_sa.md).  This
// program runs the dependent random-sector walk of k_probe_gather with different load instructions and different
// cudaLimitMaxL2FetchGranularity settings; run it plainly for the times and under
//   ncu --metrics dram__sectors_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
// for the sectors per request.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fetch_probe fetch_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 4
struct Blk { uint32_t r[8]; };

template <int MODE> __device__ __forceinline__ Blk load32(const uint4 *p)
{
	Blk b;
	if (MODE == 0) asm volatile("ld.global.nc.L1::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]), "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p));
	else if (MODE == 1) { uint4 x = __ldg(p), y = __ldg(p + 1); b.r[0] = x.x; b.r[1] = x.y; b.r[2] = x.z; b.r[3] = x.w; b.r[4] = y.x; b.r[5] = y.y; b.r[6] = y.z; b.r[7] = y.w; }
	else if (MODE == 2) asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]), "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p));
	else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]), "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p));
	else if (MODE == 4) {
		asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]) : "l"(p));
		asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p + 1));
	} else if (MODE == 5) {
		asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]) : "l"(p));
		asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p + 1));
	} else if (MODE == 6) {
		asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]) : "l"(p));
		asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p + 1));
	} else if (MODE == 7) {
		unsigned long long pol;
		asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]), "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]) : "l"(p), "l"(pol));
	} else if (MODE == 8) { // one 16-byte load only (half a block): does the request size matter?
		asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]) : "l"(p));
		b.r[4] = b.r[5] = b.r[6] = b.r[7] = 0;
	} else { // 9: a 4-byte load
		asm volatile("ld.global.u32 %0, [%1];" : "=r"(b.r[0]) : "l"(p));
		b.r[1] = b.r[2] = b.r[3] = b.r[4] = b.r[5] = b.r[6] = b.r[7] = 0;
	}
	return b;
}

template <int MODE> __global__ void __launch_bounds__(256) k_walk(const uint4 *__restrict__ buf, uint32_t n_blk, int steps, uint32_t seed, uint32_t *__restrict__ sink)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t x[CHAINS], acc = 0;
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) x[c] = (t * CHAINS + c + seed) * 2654435761u + 12345u;
	for (int s = 0; s < steps; ++s) {
		Blk o[CHAINS];
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) o[c] = load32<MODE>(buf + 2 * (size_t)(((uint64_t)x[c] * n_blk) >> 32));
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) { acc += (o[c].r[1] ^ o[c].r[2]) + (o[c].r[3] ^ o[c].r[4]) + (o[c].r[5] ^ o[c].r[6]) + o[c].r[7]; x[c] = (x[c] ^ o[c].r[0]) * 1664525u + 1013904223u; }
	}
	sink[t] = acc;
}

__global__ void k_fill(uint4 *buf, size_t n)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t h = (uint32_t)i * 2246822519u + 374761393u;
		buf[i] = make_uint4(h, h ^ 0x9e3779b9u, h * 3u, h * 7u);
	}
}

template <int MODE> static void run(const char *name, const uint4 *buf, uint32_t n_blk, uint32_t *sink, int reps)
{
	const int grid = 148 * 8, steps = 48;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_walk<MODE><<<grid, 256>>>(buf, n_blk, steps, 1u, sink);
	cudaEventRecord(e0);
	for (int r = 0; r < reps; ++r) k_walk<MODE><<<grid, 256>>>(buf, n_blk, steps, 77u + r, sink);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const cudaError_t err = cudaGetLastError();
	const double loads = (double)grid * 256 * CHAINS * steps * reps;
	printf("mode %d %-34s %8.3f ms  %7.2f G loads/s  %8.1f GB/s of 32-byte blocks  %s\n", MODE, name, ms / reps, loads / ms * 1e-6, loads * 32 / ms * 1e-6,
	       err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char **argv)
{
	const char *lim = argc > 1 ? argv[1] : "none";
	const double gb = argc > 2 ? atof(argv[2]) : 3.0;
	const int reps = argc > 3 ? atoi(argv[3]) : 3;
	cudaSetDevice(0);
	cudaFree(0);
	size_t before = 0, after = 0;
	cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
	cudaError_t rc = cudaSuccess;
	if (lim[0] != 'n') rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(lim));
	cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
	printf("cudaLimitMaxL2FetchGranularity: default %zu, asked %s -> rc %d (%s), now %zu; buffer %.1f GB\n", before, lim, (int)rc, cudaGetErrorString(rc), after, gb);
	cudaGetLastError();
	const size_t n16 = (size_t)(gb * 1e9 / 16) & ~(size_t)1;
	const uint32_t n_blk = (uint32_t)(n16 / 2);
	uint4 *buf; uint32_t *sink;
	if (cudaMalloc(&buf, n16 * 16) != cudaSuccess || cudaMalloc(&sink, 148 * 8 * 256 * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
	k_fill<<<148 * 8, 256>>>(buf, n16);
	cudaDeviceSynchronize();
	run<0>("nc.L1::evict_last.v8 (library)", buf, n_blk, sink, reps);
	run<1>("2 x __ldg(uint4)", buf, n_blk, sink, reps);
	run<2>("ld.global.v8", buf, n_blk, sink, reps);
	run<3>("nc.L1::no_allocate.v8", buf, n_blk, sink, reps);
	run<4>("2 x ld.global.cg.v4", buf, n_blk, sink, reps);
	run<5>("2 x ld.global.L2::64B.v4", buf, n_blk, sink, reps);
	run<6>("2 x ld.global.L2::128B.v4", buf, n_blk, sink, reps);
	run<7>("nc.no_allocate.L2 evict_first.v8", buf, n_blk, sink, reps);
	run<8>("1 x ld.global.nc.v4 (16 B)", buf, n_blk, sink, reps);
	run<9>("1 x ld.global.u32 (4 B)", buf, n_blk, sink, reps);
	cudaDeviceSynchronize();
	return 0;
}
