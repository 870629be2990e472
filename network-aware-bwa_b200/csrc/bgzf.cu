// bgzf.cu -- host side of the device BGZF codec (bgzf.cuh): the C-ABI calls, staging, packing of the members.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>

#include "../../include/bwa_gpu.h"
#include "bgzf.cuh"

namespace bwagpu {
int hostprep_fail(const char *fmt, ...); // bwagpu.cu: records the message for bwa_gpu_last_error, returns 1
int primary_device();                     // bwagpu.cu: the first device of bwa_gpu_init, -1 before it
void count_bgzf(int launches, double ms, int64_t bytes_in, int64_t bytes_out, int inflate); // bwagpu.cu: running totals
}
using bwagpu::hostprep_fail;

#define BCK(call)                                                                                                 \
	do {                                                                                                          \
		cudaError_t e_ = (call);                                                                                  \
		if (e_ != cudaSuccess) return hostprep_fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

namespace {

// wait for a stream without spinning on a host core (the host code of a bam2bam run needs them all)
static cudaError_t stream_sync_blocking(cudaStream_t st, cudaEvent_t &ev)
{
	cudaError_t e = cudaSuccess;
	if (!ev) e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventRecord(ev, st);
	return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}

__global__ void __launch_bounds__(bgzf::T, 2) k_bgzf_deflate(bgzf::Params P)
{
	extern __shared__ __align__(16) unsigned char bgzf_smem_raw[];
	bgzf::Smem &s = *reinterpret_cast<bgzf::Smem *>(bgzf_smem_raw);
	for (int k = blockIdx.x; k < P.n_blocks; k += gridDim.x) {
		const long long off = (long long)k * bgzf::IN_MAX;
		const int len = (int)(P.n_bytes - off < bgzf::IN_MAX ? P.n_bytes - off : bgzf::IN_MAX);
		bgzf::deflate_block(s, P.in + off, len, P.level, P.out + (size_t)k * bgzf::OUT_STRIDE, P.clen + k,
		                    P.tok + (size_t)blockIdx.x * 65536, P.x2n);
	}
}

// member offsets in the packed stream: exclusive prefix of clen (one CTA; a call has at most a few thousand blocks)
__global__ void k_bgzf_offsets(const int32_t *__restrict__ clen, int n, long long *__restrict__ off)
{
	__shared__ long long part[1024];
	__shared__ long long carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024) {
		const int i = base + (int)threadIdx.x;
		const long long v = i < n ? clen[i] : 0;
		part[threadIdx.x] = v;
		__syncthreads();
		for (int d = 1; d < 1024; d <<= 1) {
			const long long o = (int)threadIdx.x >= d ? part[threadIdx.x - d] : 0;
			__syncthreads();
			part[threadIdx.x] += o;
			__syncthreads();
		}
		if (i < n) off[i] = carry + part[threadIdx.x] - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry += part[1023];
		__syncthreads();
	}
	if (threadIdx.x == 0) off[n] = carry;
}

__global__ void k_bgzf_pack(const uint8_t *__restrict__ members, const int32_t *__restrict__ clen, const long long *__restrict__ off,
                            uint8_t *__restrict__ packed)
{
	const uint8_t *src = members + (size_t)blockIdx.x * bgzf::OUT_STRIDE;
	uint8_t *dst = packed + off[blockIdx.x];
	const int n = clen[blockIdx.x];
	// the destination starts at any byte: leading bytes up to a 4-byte boundary, then whole words assembled from two source words
	const int lead = (int)((4 - ((uintptr_t)dst & 3)) & 3);
	for (int i = threadIdx.x; i < lead && i < n; i += blockDim.x) dst[i] = src[i];
	const int words = n > lead ? (n - lead) >> 2 : 0;
	const uint32_t *src4 = (const uint32_t *)src;
	uint32_t *dst4 = (uint32_t *)(dst + lead);
	for (int w = threadIdx.x; w < words; w += blockDim.x) {
		const int b = lead + 4 * w; // source byte offset
		dst4[w] = __funnelshift_r(src4[b >> 2], src4[(b >> 2) + 1], (b & 3) << 3);
	}
	for (int i = lead + 4 * words + (int)threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(bgzf::INF_T) k_bgzf_inflate(const uint8_t *__restrict__ in, const bgzf::InfJob *__restrict__ jobs, int n,
                                                               uint8_t *out, int *__restrict__ status)
{
	__shared__ bgzf::InfTables tables[bgzf::INF_T];
	const int j = blockIdx.x * bgzf::INF_T + threadIdx.x;
	if (j >= n) return;
	const bgzf::InfJob J = jobs[j];
	status[j] = bgzf::inflate_member(in + J.in_off, J.in_len, out + J.out_off, J.out_len, tables[threadIdx.x]);
}

__global__ void __launch_bounds__(32 * bgzf::INFW_WARPS) k_bgzf_inflate_warp(const uint8_t *__restrict__ in, const bgzf::InfJob *__restrict__ jobs, int n,
                                                                             uint8_t *out, int *__restrict__ status)
{
	__shared__ bgzf::InfWarp W[bgzf::INFW_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int j = blockIdx.x * bgzf::INFW_WARPS + warp; j < n; j += gridDim.x * bgzf::INFW_WARPS) {
		const bgzf::InfJob J = jobs[j];
		const int st = bgzf::inflate_member_warp(in + J.in_off, J.in_len, out + J.out_off, J.out_len, W[warp]);
		if (lane == 0) status[j] = st;
		__syncwarp();
	}
}

struct Inflater {
	std::mutex mu;
	int dev = -1;
	cudaStream_t st = nullptr;
	cudaEvent_t ev_sync = nullptr;
	uint8_t *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
	size_t cap_in = 0, cap_out = 0, cap_h_in = 0, cap_h_out = 0;
	bgzf::InfJob *d_jobs = nullptr, *h_jobs = nullptr;
	int *d_status = nullptr, *h_status = nullptr;
	size_t cap_jobs = 0;
	void release()
	{
		if (dev < 0) return;
		cudaSetDevice(dev);
		cudaFree(d_in); cudaFree(d_out); cudaFree(d_jobs); cudaFree(d_status);
		cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFreeHost(h_jobs); cudaFreeHost(h_status);
		if (ev_sync) cudaEventDestroy(ev_sync);
		ev_sync = nullptr;
		if (st) cudaStreamDestroy(st);
		d_in = d_out = h_in = h_out = nullptr; d_jobs = h_jobs = nullptr; d_status = h_status = nullptr;
		cap_in = cap_out = cap_h_in = cap_h_out = cap_jobs = 0;
		st = nullptr; dev = -1;
	}
};
Inflater g_inflater;

struct Codec {
	std::mutex mu;
	int dev = -1;
	cudaStream_t st = nullptr;
	cudaEvent_t ev_sync = nullptr;
	uint8_t *d_in = nullptr, *d_out = nullptr, *d_packed = nullptr;
	size_t cap_in = 0, cap_blocks = 0;
	int32_t *d_clen = nullptr;
	long long *d_off = nullptr;
	uint32_t *d_tok = nullptr;
	int grid = 0;
	uint8_t *h_in = nullptr, *h_packed = nullptr; // pinned staging
	size_t cap_h_in = 0, cap_h_packed = 0;
	int32_t *h_clen = nullptr;
	long long *h_off = nullptr;
	size_t cap_h_blocks = 0;
	uint32_t x2n[32];
	void release()
	{
		if (dev < 0) return;
		cudaSetDevice(dev);
		cudaFree(d_in); cudaFree(d_out); cudaFree(d_packed); cudaFree(d_clen); cudaFree(d_off); cudaFree(d_tok);
		cudaFreeHost(h_in); cudaFreeHost(h_packed); cudaFreeHost(h_clen); cudaFreeHost(h_off);
		if (ev_sync) cudaEventDestroy(ev_sync);
		ev_sync = nullptr;
		if (st) cudaStreamDestroy(st);
		d_in = d_out = d_packed = nullptr; d_clen = nullptr; d_off = nullptr; d_tok = nullptr;
		h_in = h_packed = nullptr; h_clen = nullptr; h_off = nullptr;
		cap_in = cap_blocks = cap_h_in = cap_h_packed = cap_h_blocks = 0;
		st = nullptr; dev = -1; grid = 0;
	}
};
Codec g_codec;

int codec_prepare(Codec &c, size_t n_bytes, size_t n_blocks, bool in_is_pinned)
{
	const int dev = bwagpu::primary_device();
	if (dev < 0) return hostprep_fail("bwa_gpu_bgzf_deflate: bwa_gpu_init has not been called (no CPU fallback)");
	if (c.dev != dev) {
		c.release();
		c.dev = dev;
		BCK(cudaSetDevice(dev));
		BCK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
		int n_sm = 0;
		BCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
		c.grid = 2 * n_sm;
		BCK(cudaFuncSetAttribute(k_bgzf_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(bgzf::Smem)));
		BCK(cudaMalloc((void **)&c.d_tok, (size_t)c.grid * 65536 * sizeof(uint32_t)));
		bgzf::crc_x2n_table(c.x2n);
	}
	BCK(cudaSetDevice(dev));
	if (n_bytes > c.cap_in) {
		cudaFree(c.d_in); c.d_in = nullptr; c.cap_in = 0;
		const size_t want = n_bytes + n_bytes / 4 + 65536;
		BCK(cudaMalloc((void **)&c.d_in, want));
		c.cap_in = want;
	}
	if (n_blocks > c.cap_blocks) {
		cudaFree(c.d_out); cudaFree(c.d_packed); cudaFree(c.d_clen); cudaFree(c.d_off);
		c.d_out = c.d_packed = nullptr; c.d_clen = nullptr; c.d_off = nullptr; c.cap_blocks = 0;
		const size_t want = n_blocks + n_blocks / 4 + 16;
		BCK(cudaMalloc((void **)&c.d_out, want * bgzf::OUT_STRIDE));
		BCK(cudaMalloc((void **)&c.d_packed, want * bgzf::OUT_STRIDE + 16));
		BCK(cudaMalloc((void **)&c.d_clen, want * sizeof(int32_t)));
		BCK(cudaMalloc((void **)&c.d_off, (want + 1) * sizeof(long long)));
		c.cap_blocks = want;
	}
	if (!in_is_pinned && n_bytes > c.cap_h_in) {
		cudaFreeHost(c.h_in); c.h_in = nullptr; c.cap_h_in = 0;
		const size_t want = n_bytes + n_bytes / 4 + 65536;
		BCK(cudaMallocHost((void **)&c.h_in, want));
		c.cap_h_in = want;
	}
	if (n_blocks * (size_t)bgzf::OUT_STRIDE > c.cap_h_packed) {
		cudaFreeHost(c.h_packed); c.h_packed = nullptr; c.cap_h_packed = 0;
		const size_t want = (n_blocks + n_blocks / 4 + 16) * (size_t)bgzf::OUT_STRIDE;
		BCK(cudaMallocHost((void **)&c.h_packed, want));
		c.cap_h_packed = want;
	}
	if (n_blocks > c.cap_h_blocks) {
		cudaFreeHost(c.h_clen); cudaFreeHost(c.h_off); c.h_clen = nullptr; c.h_off = nullptr; c.cap_h_blocks = 0;
		const size_t want = n_blocks + n_blocks / 4 + 16;
		BCK(cudaMallocHost((void **)&c.h_clen, want * sizeof(int32_t)));
		BCK(cudaMallocHost((void **)&c.h_off, (want + 1) * sizeof(long long)));
		c.cap_h_blocks = want;
	}
	return 0;
}

} // namespace

namespace bwagpu {
void bgzf_release()
{
	{ std::lock_guard<std::mutex> g(g_codec.mu); g_codec.release(); }
	{ std::lock_guard<std::mutex> g(g_inflater.mu); g_inflater.release(); }
}
}

extern "C" void *bwa_gpu_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (bwagpu::primary_device() < 0) { hostprep_fail("bwa_gpu_host_alloc: bwa_gpu_init has not been called"); return nullptr; }
	cudaSetDevice(bwagpu::primary_device());
	const cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
	if (e != cudaSuccess) { hostprep_fail("bwa_gpu_host_alloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
	return p;
}

extern "C" void bwa_gpu_host_free(void *p)
{
	if (p) cudaFreeHost(p);
}

extern "C" int bwa_gpu_bgzf_deflate(const uint8_t *in, int64_t n_bytes, int level, const uint8_t **out, int64_t *out_bytes,
                                    const int32_t **member_len, int32_t *n_members, double *kernel_ms)
{
	if (n_bytes < 0 || !out || !out_bytes || (n_bytes > 0 && !in)) return hostprep_fail("bwa_gpu_bgzf_deflate: bad arguments");
	Codec &c = g_codec;
	std::lock_guard<std::mutex> g(c.mu);
	const size_t n_blocks = ((size_t)n_bytes + bgzf::IN_MAX - 1) / bgzf::IN_MAX;
	*out = nullptr; *out_bytes = 0;
	if (member_len) *member_len = nullptr;
	if (n_members) *n_members = (int32_t)n_blocks;
	if (kernel_ms) *kernel_ms = 0;
	if (n_blocks == 0) return 0;
	if (n_blocks > 0x7fffffffu / bgzf::OUT_STRIDE * 64) return hostprep_fail("bwa_gpu_bgzf_deflate: %lld bytes in one call", (long long)n_bytes);
	cudaPointerAttributes attr;
	bool pinned = cudaPointerGetAttributes(&attr, in) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (codec_prepare(c, (size_t)n_bytes, n_blocks, pinned)) return 1;
	const uint8_t *src = in;
	if (!pinned) { memcpy(c.h_in, in, (size_t)n_bytes); src = c.h_in; }
	cudaEvent_t e0, e1;
	BCK(cudaEventCreate(&e0)); BCK(cudaEventCreate(&e1));
	BCK(cudaMemcpyAsync(c.d_in, src, (size_t)n_bytes, cudaMemcpyHostToDevice, c.st));
	bgzf::Params P;
	P.in = c.d_in; P.n_bytes = n_bytes; P.n_blocks = (int)n_blocks; P.level = level;
	P.out = c.d_out; P.clen = c.d_clen; P.tok = c.d_tok;
	memcpy(P.x2n, c.x2n, sizeof(P.x2n));
	const int grid = (int)(n_blocks < (size_t)c.grid ? n_blocks : (size_t)c.grid);
	BCK(cudaEventRecord(e0, c.st));
	k_bgzf_deflate<<<grid, bgzf::T, sizeof(bgzf::Smem), c.st>>>(P);
	k_bgzf_offsets<<<1, 1024, 0, c.st>>>(c.d_clen, (int)n_blocks, c.d_off);
	k_bgzf_pack<<<(unsigned)n_blocks, 256, 0, c.st>>>(c.d_out, c.d_clen, c.d_off, c.d_packed);
	BCK(cudaEventRecord(e1, c.st));
	BCK(cudaGetLastError());
	BCK(cudaMemcpyAsync(c.h_clen, c.d_clen, n_blocks * sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));
	BCK(cudaMemcpyAsync(c.h_off, c.d_off, (n_blocks + 1) * sizeof(long long), cudaMemcpyDeviceToHost, c.st));
	BCK(stream_sync_blocking(c.st, c.ev_sync));
	const long long total = c.h_off[n_blocks];
	BCK(cudaMemcpyAsync(c.h_packed, c.d_packed, (size_t)total, cudaMemcpyDeviceToHost, c.st));
	BCK(stream_sync_blocking(c.st, c.ev_sync));
	{
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (kernel_ms) *kernel_ms = ms;
		bwagpu::count_bgzf(3, ms, n_bytes, total, 0);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	*out = c.h_packed; *out_bytes = total;
	if (member_len) *member_len = c.h_clen;
	return 0;
}

static bool is_pinned(const void *p)
{
	cudaPointerAttributes attr;
	const bool yes = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	return yes;
}

extern "C" int bwa_gpu_bgzf_inflate(const uint8_t *in, int64_t n_bytes, int32_t n_members, const int64_t *member_off, uint8_t *out,
                                    int64_t out_cap, int64_t *out_off, double *kernel_ms)
{
	if (n_members < 0 || n_bytes < 0 || (n_members && (!in || !member_off || !out || !out_off))) return hostprep_fail("bwa_gpu_bgzf_inflate: bad arguments");
	if (kernel_ms) *kernel_ms = 0;
	if (out_off) out_off[0] = 0;
	if (n_members == 0) return 0;
	const int dev = bwagpu::primary_device();
	if (dev < 0) return hostprep_fail("bwa_gpu_bgzf_inflate: bwa_gpu_init has not been called (no CPU fallback)");
	struct CallTrace { // BWAGPU_TRACE=1: the call's wall time on stderr
		int n; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
		~CallTrace()
		{
			static const bool on = getenv("BWAGPU_TRACE") && atoi(getenv("BWAGPU_TRACE"));
			if (on) fprintf(stderr, "[trace] bgzf_inflate of %d members: %.1f ms\n", n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
		}
	} call_trace{n_members};
	Inflater &c = g_inflater;
	std::lock_guard<std::mutex> g(c.mu);
	if (c.dev != dev) {
		c.release();
		c.dev = dev;
		BCK(cudaSetDevice(dev));
		BCK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
	}
	BCK(cudaSetDevice(dev));
	const size_t n = (size_t)n_members;
	if (n > c.cap_jobs) {
		cudaFree(c.d_jobs); cudaFree(c.d_status); cudaFreeHost(c.h_jobs); cudaFreeHost(c.h_status);
		c.d_jobs = c.h_jobs = nullptr; c.d_status = c.h_status = nullptr; c.cap_jobs = 0;
		const size_t want = n + n / 4 + 64;
		BCK(cudaMalloc((void **)&c.d_jobs, want * sizeof(bgzf::InfJob)));
		BCK(cudaMalloc((void **)&c.d_status, want * sizeof(int)));
		BCK(cudaMallocHost((void **)&c.h_jobs, want * sizeof(bgzf::InfJob)));
		BCK(cudaMallocHost((void **)&c.h_status, want * sizeof(int)));
		c.cap_jobs = want;
	}
	// the members' extents and, from their trailers (ISIZE: bgzf.c:325), where each one's bytes go in the output stream
	int64_t total = 0;
	for (size_t k = 0; k < n; ++k) {
		const int64_t a = member_off[k], b = member_off[k + 1];
		if (a < 0 || b > n_bytes || b - a < 26 || b - a > 65536 + 64) return hostprep_fail("bwa_gpu_bgzf_inflate: member %zu has no valid extent", k);
		uint32_t isize;
		memcpy(&isize, in + b - 4, 4);
		if (isize > 65536) return hostprep_fail("bwa_gpu_bgzf_inflate: member %zu claims %u bytes (a BGZF block holds at most 65536)", k, isize);
		c.h_jobs[k].in_off = a; c.h_jobs[k].in_len = (int)(b - a);
		c.h_jobs[k].out_off = total; c.h_jobs[k].out_len = (int)isize;
		out_off[k] = total;
		total += isize;
	}
	out_off[n] = total;
	if (total > out_cap) return hostprep_fail("bwa_gpu_bgzf_inflate: %lld bytes of output, room for %lld", (long long)total, (long long)out_cap);
	const size_t span = (size_t)(member_off[n] - member_off[0]);
	const uint8_t *src = in + member_off[0];
	for (size_t k = 0; k < n; ++k) c.h_jobs[k].in_off -= member_off[0];
	if (span > c.cap_in) {
		cudaFree(c.d_in); c.d_in = nullptr; c.cap_in = 0;
		const size_t want = span + span / 4 + 65536;
		BCK(cudaMalloc((void **)&c.d_in, want));
		c.cap_in = want;
	}
	if ((size_t)total + 1 > c.cap_out) {
		cudaFree(c.d_out); c.d_out = nullptr; c.cap_out = 0;
		const size_t want = (size_t)total + (size_t)total / 4 + 65536;
		BCK(cudaMalloc((void **)&c.d_out, want));
		c.cap_out = want;
	}
	const bool in_pinned = is_pinned(src), out_pinned = is_pinned(out);
	if (!in_pinned) {
		if (span > c.cap_h_in) {
			cudaFreeHost(c.h_in); c.h_in = nullptr; c.cap_h_in = 0;
			const size_t want = span + span / 4 + 65536;
			BCK(cudaMallocHost((void **)&c.h_in, want));
			c.cap_h_in = want;
		}
		memcpy(c.h_in, src, span);
		src = c.h_in;
	}
	if (!out_pinned && (size_t)total + 1 > c.cap_h_out) {
		cudaFreeHost(c.h_out); c.h_out = nullptr; c.cap_h_out = 0;
		const size_t want = (size_t)total + (size_t)total / 4 + 65536;
		BCK(cudaMallocHost((void **)&c.h_out, want));
		c.cap_h_out = want;
	}
	cudaEvent_t e0, e1;
	BCK(cudaEventCreate(&e0)); BCK(cudaEventCreate(&e1));
	BCK(cudaMemcpyAsync(c.d_in, src, span, cudaMemcpyHostToDevice, c.st));
	BCK(cudaMemcpyAsync(c.d_jobs, c.h_jobs, n * sizeof(bgzf::InfJob), cudaMemcpyHostToDevice, c.st));
	BCK(cudaEventRecord(e0, c.st));
	{
		static const int form = getenv("BWAGPU_INFLATE_THREAD_FORM") ? atoi(getenv("BWAGPU_INFLATE_THREAD_FORM")) : 0; // 1: the thread-per-member kernel (A/B)
		if (form) k_bgzf_inflate<<<(unsigned)((n + bgzf::INF_T - 1) / bgzf::INF_T), bgzf::INF_T, 0, c.st>>>(c.d_in, c.d_jobs, (int)n, c.d_out, c.d_status);
		else {
			int n_sm = 148;
			cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
			const unsigned blocks = (unsigned)std::min<size_t>((n + bgzf::INFW_WARPS - 1) / bgzf::INFW_WARPS, (size_t)n_sm * 6);
			k_bgzf_inflate_warp<<<blocks, 32 * bgzf::INFW_WARPS, 0, c.st>>>(c.d_in, c.d_jobs, (int)n, c.d_out, c.d_status);
		}
	}
	BCK(cudaEventRecord(e1, c.st));
	BCK(cudaGetLastError());
	BCK(cudaMemcpyAsync(c.h_status, c.d_status, n * sizeof(int), cudaMemcpyDeviceToHost, c.st));
	if (total) BCK(cudaMemcpyAsync(out_pinned ? out : c.h_out, c.d_out, (size_t)total, cudaMemcpyDeviceToHost, c.st));
	BCK(stream_sync_blocking(c.st, c.ev_sync));
	{
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (kernel_ms) *kernel_ms = ms;
		bwagpu::count_bgzf(1, ms, (int64_t)span, total, 1);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	for (size_t k = 0; k < n; ++k)
		if (c.h_status[k]) return hostprep_fail("bwa_gpu_bgzf_inflate: member %zu (at byte %lld of the input) is not a valid BGZF member (code %d)", k,
		                                        (long long)member_off[k], c.h_status[k]);
	if (!out_pinned && total) memcpy(out, c.h_out, (size_t)total);
	return 0;
}
