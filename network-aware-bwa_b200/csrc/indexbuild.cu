// indexbuild.cu -- FM-index construction on the device.
//
// Replaces the compute of `bwa index` (bwtindex.c:102-190): pac -> BWT (is.c / bwt_gen/: bwt_pac2bwt, bwtmisc.c:56-101)
// -> occurrence counts interleaved every 128 bases (bwt_bwtupdate_core, bwtmisc.c:125-152) -> suffix array sampled
// every 32 rows (bwt_cal_sa, bwt.c:48-70), for the forward strand and for the reversed (NOT complemented) strand
// (bwa_pac_rev_core, bwtmisc.c:168-193).  The output is the reference's own bwt_t image, word for word what
// bwt_restore_bwt + bwt_restore_sa load from `bwa index` files (the BWT of a string is unique), so the files this
// builder writes are interchangeable with the reference's.
//
// The reference builds the BWT by induced sorting (< 2 Gb) or the incremental BWT-SW builder (hours for 3.1 Gb) and the
// sampled SA by walking LF over the whole BWT.  Here the suffix array is sorted outright in HBM -- 180 GB hold it for
// any genome the 32-bit bwtint_t admits -- by prefix doubling on 64-bit radix sorts (cub):
//   round 0   sort all suffixes by their first 21 symbols (3 bits each, 0 = past the end);
//   round r   re-sort only the suffixes that still share their group with another one (Larsson-Sadakane refinement):
//             key = (group start << 32 | rank of the suffix h symbols further on), h doubling every round.
// On a genome-like string nearly every suffix is alone after round 0, so later rounds touch only repeats.
// Everything fits u32 (n < 2^32 - 64).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bwa_gpu.h"

namespace bwagpu {
int hostprep_fail(const char *fmt, ...); // bwagpu.cu: records the message for bwa_gpu_last_error, returns 1
}
using bwagpu::hostprep_fail;

#define ICK(call)                                                                                                 \
	do {                                                                                                          \
		cudaError_t e_ = (call);                                                                                  \
		if (e_ != cudaSuccess) return hostprep_fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

namespace {

typedef unsigned long long u64;
const int K0 = 21; // symbols in the round-0 key

struct MaxOp {
	__device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

template <typename T> struct Dev {
	T *p = nullptr;
	size_t n = 0;
	int alloc(size_t count)
	{
		release();
		if (count == 0) count = 1;
		cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
		if (e != cudaSuccess) return hostprep_fail("index build: cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
		n = count;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
	~Dev() { release(); }
};

const int GRID = 148 * 16, BLK = 256;
#define GSTRIDE(i, n) for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < (u64)(n); i += (u64)gridDim.x * blockDim.x)

// pac (4 bases per byte, first base in the top bits: bntseq.c:207-210) -> one byte per base; rev: T[i] = base n-1-i
__global__ void k_unpack(const uint8_t *__restrict__ pac, u64 n, int rev, uint8_t *__restrict__ T)
{
	GSTRIDE(i, n) {
		const u64 s = rev ? n - 1 - i : i;
		T[i] = (pac[s >> 2] >> ((~s & 3) << 1)) & 3;
	}
}

__global__ void k_keys(const uint8_t *__restrict__ T, u64 n, u64 *__restrict__ key, uint32_t *__restrict__ idx)
{
	GSTRIDE(i, n) {
		u64 k = 0;
#pragma unroll
		for (int j = 0; j < K0; ++j) {
			const u64 p = i + j;
			k = k << 3 | (p < n ? (u64)T[p] + 1 : 0);
		}
		key[i] = k;
		idx[i] = (uint32_t)i;
	}
}

// head[i] = i when sorted position i starts a group (its key differs from its predecessor's), else 0
__global__ void k_heads(const u64 *__restrict__ ks, u64 n, uint32_t *__restrict__ head)
{
	GSTRIDE(i, n) head[i] = (i == 0 || ks[i] != ks[i - 1]) ? (uint32_t)i : 0u;
}

// grp[i] = sorted position of the head of i's group.  rank[sa[i]] = grp[i]; unsorted[i] = 1 when the group has > 1 member
__global__ void k_rank_flag(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ grp, u64 n, uint32_t *__restrict__ rank,
                            uint32_t *__restrict__ unsorted)
{
	GSTRIDE(i, n) {
		const uint32_t g = grp[i];
		rank[sa[i]] = g;
		const bool single = g == (uint32_t)i && (i + 1 == n || grp[i + 1] == (uint32_t)(i + 1));
		unsorted[i] = single ? 0u : 1u;
	}
}

// compaction: elements with flag[i] set go to position pos[i] (exclusive scan of flag)
__global__ void k_compact(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ a,
                          const uint32_t *__restrict__ b, u64 n, uint32_t *__restrict__ oa, uint32_t *__restrict__ ob)
{
	GSTRIDE(i, n) if (flag[i]) { oa[pos[i]] = a[i]; ob[pos[i]] = b[i]; }
}

__global__ void k_keys2(const uint32_t *__restrict__ pidx, const uint32_t *__restrict__ pgrp, const uint32_t *__restrict__ rank, u64 m,
                        u64 n, u64 h, u64 *__restrict__ key)
{
	GSTRIDE(j, m) {
		const u64 nx = (u64)pidx[j] + h;
		key[j] = (u64)pgrp[j] << 32 | (nx < n ? (u64)rank[nx] + 1 : 0ull);
	}
}

// list heads after the re-sort: gh = first list index of the old group (high word), sh = of the new sub-group (whole key)
__global__ void k_heads2(const u64 *__restrict__ ks, u64 m, uint32_t *__restrict__ gh, uint32_t *__restrict__ sh)
{
	GSTRIDE(j, m) {
		const bool g = j == 0 || (ks[j] >> 32) != (ks[j - 1] >> 32);
		gh[j] = g ? (uint32_t)j : 0u;
		sh[j] = (g || ks[j] != ks[j - 1]) ? (uint32_t)j : 0u;
	}
}

__global__ void k_refine(const u64 *__restrict__ ks, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ gh,
                         const uint32_t *__restrict__ sh, u64 m, uint32_t *__restrict__ sa, uint32_t *__restrict__ rank,
                         uint32_t *__restrict__ newgrp, uint32_t *__restrict__ keep)
{
	GSTRIDE(j, m) {
		const uint32_t g = (uint32_t)(ks[j] >> 32);
		const uint32_t ng = g + (sh[j] - gh[j]);
		sa[g + ((uint32_t)j - gh[j])] = idx[j];
		rank[idx[j]] = ng;
		newgrp[j] = ng;
		const bool single = sh[j] == (uint32_t)j && (j + 1 == m || sh[j + 1] == (uint32_t)(j + 1));
		keep[j] = single ? 0u : 1u;
	}
}

// BWT0 (the BWT without the '$' row, is.c:212-213), 16 bases per word, first base in the top bits (bwtmisc.c:97-98)
__global__ void k_bwt_words(const uint8_t *__restrict__ T, const uint32_t *__restrict__ sa, u64 n, u64 primary, u64 nw,
                            uint32_t *__restrict__ words)
{
	GSTRIDE(w, nw) {
		uint32_t x = 0;
		for (int t = 0; t < 16; ++t) {
			const u64 j = w * 16 + t;
			uint32_t c = 0;
			if (j < n) {
				const u64 r = j < primary ? j : j + 1;          // row of the full matrix (row `primary` is the '$' row)
				const u64 s = r == 0 ? n : (u64)sa[r - 1];      // SA_full[0] = n (the empty suffix sorts first)
				c = T[s - 1];
			}
			x |= c << ((15 - t) << 1);
		}
		words[w] = x;
	}
}

// per 128-base block: how many of each symbol it holds
__global__ void k_block_counts(const uint32_t *__restrict__ words, u64 n, u64 nw, u64 nb, uint32_t *__restrict__ c0,
                               uint32_t *__restrict__ c1, uint32_t *__restrict__ c2, uint32_t *__restrict__ c3)
{
	GSTRIDE(b, nb) {
		uint32_t n1 = 0, n2 = 0, n3 = 0, valid = 0;
		for (int w = 0; w < 8; ++w) {
			const u64 gw = b * 8 + w;
			if (gw >= nw) break;
			const uint32_t x = words[gw];
			const u64 left = n - gw * 16;
			const uint32_t v = left >= 16 ? 16u : (uint32_t)left;
			const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
			n1 += __popc(lo & ~hi); n2 += __popc(hi & ~lo); n3 += __popc(hi & lo); // padding bases are 0: never counted here
			valid += v;
		}
		c0[b] = valid - n1 - n2 - n3; c1[b] = n1; c2[b] = n2; c3[b] = n3;
	}
}

// the reference's array: per block 4 running counts then up to 8 words; the totals follow the last word
__global__ void k_assemble(const uint32_t *__restrict__ words, const uint32_t *__restrict__ e0, const uint32_t *__restrict__ e1,
                           const uint32_t *__restrict__ e2, const uint32_t *__restrict__ e3, u64 nw, u64 nb, uint32_t *__restrict__ out)
{
	GSTRIDE(b, nb + 1) {
		if (b < nb) {
			uint32_t *o = out + b * 12;
			o[0] = e0[b]; o[1] = e1[b]; o[2] = e2[b]; o[3] = e3[b];
			for (int w = 0; w < 8; ++w) {
				const u64 gw = b * 8 + w;
				if (gw < nw) o[4 + w] = words[gw];
			}
		} else {
			uint32_t *o = out + nw + 4 * nb;
			o[0] = e0[nb]; o[1] = e1[nb]; o[2] = e2[nb]; o[3] = e3[nb];
		}
	}
}

__global__ void k_sa_sample(const uint32_t *__restrict__ sa, u64 n_sa, uint32_t intv, uint32_t *__restrict__ out)
{
	GSTRIDE(k, n_sa) out[k] = k == 0 ? 0xffffffffu : sa[k * intv - 1]; // bwt.c:62-69: sa[0] = -1
}

struct Scratch {
	Dev<uint8_t> tmp;
	int reserve(size_t bytes)
	{
		if (bytes <= tmp.n) return 0;
		return tmp.alloc(bytes + bytes / 8 + 256);
	}
};

int max_scan(Scratch &S, uint32_t *d, u64 n, cudaStream_t st)
{
	size_t tb = 0;
	ICK(cub::DeviceScan::InclusiveScan((void *)nullptr, tb, d, d, MaxOp(), (long long)n, st));
	if (S.reserve(tb)) return 1;
	ICK(cub::DeviceScan::InclusiveScan((void *)S.tmp.p, tb, d, d, MaxOp(), (long long)n, st));
	return 0;
}

int excl_sum(Scratch &S, const uint32_t *in, uint32_t *out, u64 n, cudaStream_t st)
{
	size_t tb = 0;
	ICK(cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, in, out, (long long)n, st));
	if (S.reserve(tb)) return 1;
	ICK(cub::DeviceScan::ExclusiveSum((void *)S.tmp.p, tb, in, out, (long long)n, st));
	return 0;
}

int sort_pairs(Scratch &S, const u64 *kin, u64 *kout, const uint32_t *vin, uint32_t *vout, u64 n, int end_bit, cudaStream_t st)
{
	size_t tb = 0;
	ICK(cub::DeviceRadixSort::SortPairs((void *)nullptr, tb, kin, kout, vin, vout, (long long)n, 0, end_bit, st));
	if (S.reserve(tb)) return 1;
	ICK(cub::DeviceRadixSort::SortPairs((void *)S.tmp.p, tb, kin, kout, vin, vout, (long long)n, 0, end_bit, st));
	return 0;
}

int bits_for(u64 v)
{
	int b = 1;
	while (b < 64 && (v >> b)) ++b;
	return b;
}

// One strand.  pac on the device; out: the reference's bwt_t fields, arrays malloc()'d like bwt_restore_bwt / bwt_restore_sa do.
int build_strand(const uint8_t *d_pac, u64 n, int rev, bwt_t *out, cudaStream_t st, double *ms_sort)
{
	Scratch S;
	Dev<uint8_t> T;
	Dev<uint32_t> sa, rank;
	if (T.alloc(n)) return 1;
	k_unpack<<<GRID, BLK, 0, st>>>(d_pac, n, rev, T.p);
	ICK(cudaGetLastError());
	cudaEvent_t e0, e1;
	ICK(cudaEventCreate(&e0)); ICK(cudaEventCreate(&e1));
	ICK(cudaEventRecord(e0, st));

	Dev<uint32_t> pidx, pgrp; // the suffixes still sharing a group: text position, sorted position of the group's head
	u64 m = 0;
	{
		// ---- round 0: every suffix by its first K0 symbols
		Dev<u64> ks;
		{
			Dev<u64> kin;
			Dev<uint32_t> iin;
			if (kin.alloc(n) || iin.alloc(n) || ks.alloc(n) || sa.alloc(n)) return 1;
			k_keys<<<GRID, BLK, 0, st>>>(T.p, n, kin.p, iin.p);
			ICK(cudaGetLastError());
			if (sort_pairs(S, kin.p, ks.p, iin.p, sa.p, n, 3 * K0, st)) return 1;
			ICK(cudaStreamSynchronize(st));
		}
		Dev<uint32_t> grp, flag;
		if (grp.alloc(n)) return 1;
		k_heads<<<GRID, BLK, 0, st>>>(ks.p, n, grp.p);
		ICK(cudaGetLastError());
		ICK(cudaStreamSynchronize(st));
		ks.release();
		if (max_scan(S, grp.p, n, st)) return 1;
		if (rank.alloc(n) || flag.alloc(n + 1)) return 1;
		k_rank_flag<<<GRID, BLK, 0, st>>>(sa.p, grp.p, n, rank.p, flag.p);
		ICK(cudaGetLastError());
		ICK(cudaMemsetAsync(flag.p + n, 0, 4, st));
		Dev<uint32_t> pos;
		if (pos.alloc(n + 1)) return 1;
		if (excl_sum(S, flag.p, pos.p, n + 1, st)) return 1;
		uint32_t m32 = 0;
		ICK(cudaMemcpyAsync(&m32, pos.p + n, 4, cudaMemcpyDeviceToHost, st));
		ICK(cudaStreamSynchronize(st));
		m = m32;
		if (m) {
			if (pidx.alloc(m) || pgrp.alloc(m)) return 1;
			k_compact<<<GRID, BLK, 0, st>>>(flag.p, pos.p, sa.p, grp.p, n, pidx.p, pgrp.p);
			ICK(cudaGetLastError());
			ICK(cudaStreamSynchronize(st));
		}
	}
	// ---- refinement rounds: only the suffixes still in the list
	const int kbits = 32 + bits_for(n); // high word = group start (< n), low word = rank + 1 (<= n)
	int rounds = 0;
	for (u64 h = K0; m > 0; h <<= 1, ++rounds) {
		if (rounds > 40) return hostprep_fail("index build: prefix doubling did not converge");
		Dev<u64> kin, ks;
		Dev<uint32_t> idx, gh, sh, ng, keep, pos;
		if (kin.alloc(m) || ks.alloc(m) || idx.alloc(m) || gh.alloc(m) || sh.alloc(m) || ng.alloc(m) || keep.alloc(m + 1) || pos.alloc(m + 1)) return 1;
		k_keys2<<<GRID, BLK, 0, st>>>(pidx.p, pgrp.p, rank.p, m, n, h, kin.p);
		ICK(cudaGetLastError());
		if (sort_pairs(S, kin.p, ks.p, pidx.p, idx.p, m, kbits > 64 ? 64 : kbits, st)) return 1;
		k_heads2<<<GRID, BLK, 0, st>>>(ks.p, m, gh.p, sh.p);
		ICK(cudaGetLastError());
		if (max_scan(S, gh.p, m, st) || max_scan(S, sh.p, m, st)) return 1;
		k_refine<<<GRID, BLK, 0, st>>>(ks.p, idx.p, gh.p, sh.p, m, sa.p, rank.p, ng.p, keep.p);
		ICK(cudaGetLastError());
		ICK(cudaMemsetAsync(keep.p + m, 0, 4, st));
		if (excl_sum(S, keep.p, pos.p, m + 1, st)) return 1;
		uint32_t m2 = 0;
		ICK(cudaMemcpyAsync(&m2, pos.p + m, 4, cudaMemcpyDeviceToHost, st));
		ICK(cudaStreamSynchronize(st));
		if (m2) {
			Dev<uint32_t> nidx, ngrp;
			if (nidx.alloc(m2) || ngrp.alloc(m2)) return 1;
			k_compact<<<GRID, BLK, 0, st>>>(keep.p, pos.p, idx.p, ng.p, m, nidx.p, ngrp.p);
			ICK(cudaGetLastError());
			ICK(cudaStreamSynchronize(st));
			std::swap(pidx.p, nidx.p); std::swap(pidx.n, nidx.n);
			std::swap(pgrp.p, ngrp.p); std::swap(pgrp.n, ngrp.n);
		}
		m = m2;
	}
	pidx.release(); pgrp.release();
	ICK(cudaEventRecord(e1, st));

	// ---- primary: the row of the full matrix whose suffix is the whole text (is.c:209-211)
	uint32_t r0 = 0;
	ICK(cudaMemcpyAsync(&r0, rank.p, 4, cudaMemcpyDeviceToHost, st));
	ICK(cudaStreamSynchronize(st));
	rank.release();
	const u64 primary = (u64)r0 + 1;

	// ---- BWT words, occurrence counts, the reference's interleaved array
	const u64 nw = (n + 15) >> 4, nb = (n + 127) >> 7;
	const u64 bwt_size = nw + 4 * (nb + 1); // bwtmisc.c:133-135
	Dev<uint32_t> words, c[4], e[4], outw;
	if (words.alloc(nw) || outw.alloc(bwt_size)) return 1;
	for (int q = 0; q < 4; ++q) if (c[q].alloc(nb + 1) || e[q].alloc(nb + 1)) return 1;
	k_bwt_words<<<GRID, BLK, 0, st>>>(T.p, sa.p, n, primary, nw, words.p);
	ICK(cudaGetLastError());
	for (int q = 0; q < 4; ++q) ICK(cudaMemsetAsync(c[q].p + nb, 0, 4, st));
	k_block_counts<<<GRID, BLK, 0, st>>>(words.p, n, nw, nb, c[0].p, c[1].p, c[2].p, c[3].p);
	ICK(cudaGetLastError());
	for (int q = 0; q < 4; ++q) if (excl_sum(S, c[q].p, e[q].p, nb + 1, st)) return 1;
	k_assemble<<<GRID, BLK, 0, st>>>(words.p, e[0].p, e[1].p, e[2].p, e[3].p, nw, nb, outw.p);
	ICK(cudaGetLastError());
	uint32_t tot[4];
	for (int q = 0; q < 4; ++q) ICK(cudaMemcpyAsync(&tot[q], e[q].p + nb, 4, cudaMemcpyDeviceToHost, st));
	ICK(cudaStreamSynchronize(st));
	T.release(); words.release();

	// ---- sampled suffix array (bwt_cal_sa(bwt, 32), bwtindex.c:173,185)
	const uint32_t intv = 32;
	const u64 n_sa = (n + intv) / intv;
	Dev<uint32_t> sas;
	if (sas.alloc(n_sa)) return 1;
	k_sa_sample<<<GRID, BLK, 0, st>>>(sa.p, n_sa, intv, sas.p);
	ICK(cudaGetLastError());

	// ---- the reference's bwt_t image
	memset(out, 0, sizeof(*out));
	out->primary = (bwtint_t)primary;
	out->seq_len = (bwtint_t)n;
	out->L2[0] = 0;
	for (int q = 0; q < 4; ++q) out->L2[q + 1] = out->L2[q] + tot[q];
	out->bwt_size = (bwtint_t)bwt_size;
	out->bwt = (uint32_t *)malloc(bwt_size * 4);
	out->sa_intv = (int)intv;
	out->n_sa = (bwtint_t)n_sa;
	out->sa = (bwtint_t *)malloc(n_sa * sizeof(bwtint_t));
	if (!out->bwt || !out->sa) return hostprep_fail("index build: out of host memory");
	ICK(cudaMemcpyAsync(out->bwt, outw.p, bwt_size * 4, cudaMemcpyDeviceToHost, st));
	ICK(cudaMemcpyAsync(out->sa, sas.p, n_sa * 4, cudaMemcpyDeviceToHost, st));
	ICK(cudaStreamSynchronize(st));
	for (int i = 0; i != 256; ++i) { // bwt_gen_cnt_table (bwt.c:36-45)
		uint32_t x = 0;
		for (int j = 0; j != 4; ++j) x |= (((i & 3) == j) + ((i >> 2 & 3) == j) + ((i >> 4 & 3) == j) + (i >> 6 == j)) << (j << 3);
		out->cnt_table[i] = x;
	}
	float ms = 0;
	ICK(cudaEventElapsedTime(&ms, e0, e1));
	if (ms_sort) *ms_sort += ms;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if (getenv("BWAGPU_TRACE")) fprintf(stderr, "[index build] strand %d: %llu bases, suffix sort %.0f ms, %d refinement rounds\n", rev, n, ms, rounds);
	return 0;
}

} // namespace

// ------------------------------------------------------------------ C-ABI
extern "C" int bwa_gpu_index_build(const ubyte_t *pac, int64_t l_pac, int device, bwt_t *fwd, bwt_t *rev)
{
	if (!pac || l_pac <= 0 || !fwd || !rev) return hostprep_fail("bwa_gpu_index_build: bad argument");
	if ((uint64_t)l_pac >= 0xffffffc0ull) return hostprep_fail("bwa_gpu_index_build: %lld bases do not fit the 32-bit bwtint_t (bwtindex.c:103-106)", (long long)l_pac);
	int have = 0;
	if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return hostprep_fail("bwa_gpu_index_build: no CUDA device (this library has no CPU fallback)");
	if (device < 0 || device >= have) return hostprep_fail("bwa_gpu_index_build: device %d out of range", device);
	ICK(cudaSetDevice(device));
	cudaStream_t st;
	ICK(cudaStreamCreate(&st));
	const size_t nbytes = (size_t)(l_pac / 4 + 1);
	Dev<uint8_t> d_pac;
	if (d_pac.alloc(nbytes)) return 1;
	ICK(cudaMemcpyAsync(d_pac.p, pac, nbytes, cudaMemcpyHostToDevice, st));
	double ms = 0;
	int rc = build_strand(d_pac.p, (u64)l_pac, 0, fwd, st, &ms);
	if (!rc) rc = build_strand(d_pac.p, (u64)l_pac, 1, rev, st, &ms);
	cudaStreamDestroy(st);
	return rc;
}

extern "C" void bwa_gpu_index_free(bwt_t *b)
{
	if (!b) return;
	free(b->bwt); free(b->sa);
	b->bwt = 0; b->sa = 0;
}

static int dump_words(FILE *f, const void *p, size_t n_words)
{
	const char *q = (const char *)p;
	size_t left = n_words * 4;
	while (left) { // fwrite in pieces: some libcs cap a single call
		const size_t m = left < ((size_t)1 << 30) ? left : (size_t)1 << 30;
		if (fwrite(q, 1, m, f) != m) return 1;
		q += m; left -= m;
	}
	return 0;
}

// bwt_dump_bwt (bwtio.c:17-25) and bwt_dump_sa (bwtio.c:27-38) for both strands: <prefix>.bwt/.rbwt/.sa/.rsa
extern "C" int bwa_gpu_index_write(const char *prefix, const bwt_t *fwd, const bwt_t *rev)
{
	if (!prefix || !fwd || !rev) return hostprep_fail("bwa_gpu_index_write: bad argument");
	const bwt_t *b[2] = {fwd, rev};
	static const char *ext_bwt[2] = {".bwt", ".rbwt"}, *ext_sa[2] = {".sa", ".rsa"};
	for (int s = 0; s < 2; ++s) {
		std::string fn = std::string(prefix) + ext_bwt[s];
		FILE *f = fopen(fn.c_str(), "wb");
		if (!f) return hostprep_fail("bwa_gpu_index_write: cannot open %s", fn.c_str());
		int bad = fwrite(&b[s]->primary, sizeof(bwtint_t), 1, f) != 1 || fwrite(b[s]->L2 + 1, sizeof(bwtint_t), 4, f) != 4 ||
		          dump_words(f, b[s]->bwt, b[s]->bwt_size);
		bad |= fclose(f) != 0;
		if (bad) return hostprep_fail("bwa_gpu_index_write: error writing %s", fn.c_str());
		fn = std::string(prefix) + ext_sa[s];
		f = fopen(fn.c_str(), "wb");
		if (!f) return hostprep_fail("bwa_gpu_index_write: cannot open %s", fn.c_str());
		const bwtint_t intv = (bwtint_t)b[s]->sa_intv;
		bad = fwrite(&b[s]->primary, sizeof(bwtint_t), 1, f) != 1 || fwrite(b[s]->L2 + 1, sizeof(bwtint_t), 4, f) != 4 ||
		      fwrite(&intv, sizeof(bwtint_t), 1, f) != 1 || fwrite(&b[s]->seq_len, sizeof(bwtint_t), 1, f) != 1 ||
		      dump_words(f, b[s]->sa + 1, b[s]->n_sa - 1);
		bad |= fclose(f) != 0;
		if (bad) return hostprep_fail("bwa_gpu_index_write: error writing %s", fn.c_str());
	}
	return 0;
}
