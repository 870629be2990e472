// fmindex.cuh -- K1: FM-index occurrence primitives on the HBM layout (device side).
//
// Replaces bwt_occ / bwt_2occ / bwt_occ4 / bwt_2occ4 / bwt_invPsi / bwt_sa
// (reference bwt.c:72-216, bwt.h:61-75).
//
// Layout.  The reference stores, per 128 bases, 4 u32 running counts + 8
// u32 words of 2-bit bases (48 B, stride 48 B: an occ query straddles two or three 32 B
// DRAM sectors).  Here one block is exactly ONE 32-byte sector, 32-byte aligned:
//
//     struct { u32 cnt[4];   // # of A,C,G,T in BWT0[0 .. 64*b)
//              u64 lo, hi; } // bit j of lo/hi = low/high bit of base BWT0[64*b + j]
//
// so every occ / occ4 is a single 256-bit load and three 64-bit popcounts, with no
// byte-LUT (bwt.c:155-157), no "c == 0" correction for masked-out pairs (bwt.c:112,
// 143,150,174) and no per-word loop.  BWT0 is the BWT with the '$' row removed, as in the
// reference (is.c:212-213), hence the `k >= primary -> k-1` shift (bwt.c:99,167).
//
// Query convention: callers turn the reference's argument k (inclusive upper index, may
// be (u32)-1) into j = "number of leading BWT0 symbols to count":
//     j = (k == -1) ? 0 : (k >= primary ? k : k + 1)        0 <= j <= seq_len
// Block j>>6 always exists because the array has (seq_len>>6)+1 blocks.
#pragma once
#ifndef BWAGPU_L2HINT
#define BWAGPU_L2HINT 0 // 1: L2 evict_last policy on index loads (A/B switch)
#endif
#ifndef BWAGPU_IDX_STREAM
#define BWAGPU_IDX_STREAM 0 // 1: index loads L1::evict_first + L2 evict_first; 2: L1::no_allocate + L2 evict_first (A/B switch)
#endif
#ifndef BWAGPU_CTX_KEEP
#define BWAGPU_CTX_KEEP 0 // 1: context-entry loads L1::evict_last + L2 evict_last (A/B switch)
#endif
#ifndef BWAGPU_IDX_L2_64B
#define BWAGPU_IDX_L2_64B 1 // 1: index loads carry .L2::64B -- the L2 fills 2 sectors per miss instead of the whole 128-byte line (0: plain loads, for A/B runs)
#endif
#ifndef BWAGPU_LDG256
#define BWAGPU_LDG256 1 // sm_100a has LDG.E.256; set to 0 for two LDG.128
#endif
#include <stdint.h>
#ifdef BWAGPU_HOST_EMU
#include "host_emu_shim.h" // tests/host_emu: runs the kernel bodies on the CPU for logic tests only
#else
#include <cuda_runtime.h>
#endif

namespace bwagpu {

struct DevIndex {
	const uint4 *blk;    // 2 uint4 per block: {cnt0..3}, {lo.lo32, lo.hi32, hi.lo32, hi.hi32}
	const uint32_t *sa;  // sampled suffix array, sa[0] = 0xffffffff (bwt.c:69)
	uint32_t primary, seq_len, n_sa, sa_intv;
	uint32_t L2[5];
	uint32_t n_blk;
};

struct OccBlock {
	uint32_t c0, c1, c2, c3;
	uint64_t lo, hi;
};

// One 32-byte sector.  ld.global.nc.v8 would be a single LDG.256; two LDG.128 to the same
// sector cost one extra L1 wavefront but are accepted by every CUDA 12.x ptxas.
__device__ __forceinline__ OccBlock load_block(const DevIndex &ix, uint32_t b)
{
	const uint4 *p = ix.blk + 2 * (size_t)b;
	OccBlock o;
#if BWAGPU_LDG256
	uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#if BWAGPU_IDX_STREAM
	// an index far larger than L2 has no reuse beyond a trip's own two lookups: keep it out of the way of the searches'
	// state (context entries, stack records), which is re-read for as long as a read lasts
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#if BWAGPU_IDX_STREAM == 2
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
#else
	asm volatile("ld.global.nc.L1::evict_first.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
#endif
	             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
	             : "l"(p), "l"(pol));
#elif BWAGPU_L2HINT
	// keep index sectors in L2 ahead of the streaming stack / width traffic
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
	             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
	             : "l"(p), "l"(pol));
#elif BWAGPU_IDX_L2_64B
	// Measured (scripts/probe/fetch_probe.cu, profiles/r2_fetch_probe.md): a plain load that misses L2 makes it fill the whole
	// 128-byte line from DRAM -- 3.9 sectors per random 32-byte block -- whatever cudaLimitMaxL2FetchGranularity says;
	// .L2::64B is the smallest fill the ISA offers (1.9-2.0 sectors per block)
	asm volatile("ld.global.nc.L1::evict_last.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
	             : "l"(p));
#else
	asm volatile("ld.global.nc.L1::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
	             : "l"(p));
#endif
	o.c0 = r0; o.c1 = r1; o.c2 = r2; o.c3 = r3;
	o.lo = (uint64_t)r5 << 32 | r4;
	o.hi = (uint64_t)r7 << 32 | r6;
#else
	uint4 a = __ldg(p), b2 = __ldg(p + 1);
	o.c0 = a.x; o.c1 = a.y; o.c2 = a.z; o.c3 = a.w;
	o.lo = (uint64_t)b2.y << 32 | b2.x;
	o.hi = (uint64_t)b2.w << 32 | b2.z;
#endif
	return o;
}

// reference argument k  ->  count j  (see header comment)
__device__ __forceinline__ uint32_t occ_arg(const DevIndex &ix, uint32_t k)
{
	return k == 0xffffffffu ? 0u : (k >= ix.primary ? k : k + 1);
}

// all four occurrence counts of the first j symbols, given j's block
__device__ __forceinline__ void occ4_in_block(const OccBlock &o, uint32_t j, uint32_t cnt[4])
{
	const uint32_t r = j & 63u;
	const uint64_t m = (1ull << r) - 1ull;
	const uint32_t n3 = __popcll(o.hi & o.lo & m);
	const uint32_t n2 = __popcll(o.hi & ~o.lo & m);
	const uint32_t n1 = __popcll(~o.hi & o.lo & m);
	cnt[0] = o.c0 + (r - n1 - n2 - n3);
	cnt[1] = o.c1 + n1;
	cnt[2] = o.c2 + n2;
	cnt[3] = o.c3 + n3;
}

// occurrence count of one symbol c (0..3) among the first j symbols
__device__ __forceinline__ uint32_t occ1_in_block(const OccBlock &o, uint32_t j, uint32_t c)
{
	const uint32_t r = j & 63u;
	const uint64_t m = (1ull << r) - 1ull;
	const uint64_t h = (c & 2u) ? o.hi : ~o.hi;
	const uint64_t l = (c & 1u) ? o.lo : ~o.lo;
	const uint32_t base = c == 0 ? o.c0 : c == 1 ? o.c1 : c == 2 ? o.c2 : o.c3;
	return base + __popcll(h & l & m);
}

// bwt_2occ4(bwt, km1, l, cntk, cntl) -- bwt.c:179-216.  Requires km1 < l or km1 == -1.
// `fetch_ref` (optional) accumulates the number of occ blocks the REFERENCE layout reads
// for this query (1 if both ends share a 128-base block, else 2; a -1 end costs nothing),
// `fetch_own` the number of 32-byte blocks actually loaded here.
template <bool STATS>
__device__ __forceinline__ void occ4_pair(const DevIndex &ix, uint32_t km1, uint32_t l, uint32_t cntk[4],
                                          uint32_t cntl[4], uint32_t &fetch_ref, uint32_t &fetch_own)
{
	const uint32_t jk = occ_arg(ix, km1), jl = occ_arg(ix, l);
	const uint32_t bk = jk >> 6, bl = jl >> 6;
	OccBlock ol = load_block(ix, bl);
	OccBlock ok = ol;
	if (bk != bl) ok = load_block(ix, bk);
	occ4_in_block(ok, jk, cntk);
	occ4_in_block(ol, jl, cntl);
	if (STATS) {
		fetch_own += (bk != bl) ? 2u : 1u;
		if (km1 == 0xffffffffu) fetch_ref += 1u;
		else {
			// the reference compares the inclusive positions (k>=primary ? k-1 : k) >> 7
			const uint32_t pk = km1 >= ix.primary ? km1 - 1 : km1, pl = l >= ix.primary ? l - 1 : l;
			fetch_ref += (pk >> 7) == (pl >> 7) ? 1u : 2u;
		}
	}
}

// bwt_2occ(bwt, km1, l, c, &ok, &ol) -- bwt.c:118-153
template <bool STATS>
__device__ __forceinline__ void occ1_pair(const DevIndex &ix, uint32_t km1, uint32_t l, uint32_t c, uint32_t &ck,
                                          uint32_t &cl, uint32_t &fetch_ref, uint32_t &fetch_own)
{
	const uint32_t jk = occ_arg(ix, km1), jl = occ_arg(ix, l);
	const uint32_t bk = jk >> 6, bl = jl >> 6;
	OccBlock ol = load_block(ix, bl);
	OccBlock ok = ol;
	if (bk != bl) ok = load_block(ix, bk);
	ck = occ1_in_block(ok, jk, c);
	cl = occ1_in_block(ol, jl, c);
	if (STATS) {
		fetch_own += (bk != bl) ? 2u : 1u;
		// bwt_occ short-circuits k == seq_len without touching the index (bwt.c:97)
		const uint32_t lcost = l == ix.seq_len ? 0u : 1u;
		if (km1 == 0xffffffffu) fetch_ref += lcost;
		else {
			const uint32_t pk = km1 >= ix.primary ? km1 - 1 : km1, pl = l >= ix.primary ? l - 1 : l;
			fetch_ref += (pk >> 7) == (pl >> 7) ? 1u : 1u + lcost;
		}
	}
}

// bwt_invPsi (bwt.h:71-75): one LF step.  The symbol BWT0[p] and its rank come from the
// SAME 32-byte block (p = k or k-1), so a step is one sector, against B0 + occ in the
// reference (two touches of a 48-byte block).
__device__ __forceinline__ uint32_t inv_psi(const DevIndex &ix, uint32_t k)
{
	if (k == ix.primary) return 0;
	const uint32_t p = k < ix.primary ? k : k - 1;
	const OccBlock o = load_block(ix, p >> 6);
	const uint32_t r = p & 63u;
	const uint32_t c = (uint32_t)((o.hi >> r) & 1ull) << 1 | (uint32_t)((o.lo >> r) & 1ull);
	const uint64_t m = r == 63u ? ~0ull : ((2ull << r) - 1ull); // symbols 0..r inclusive
	const uint64_t h = (c & 2u) ? o.hi : ~o.hi;
	const uint64_t l = (c & 1u) ? o.lo : ~o.lo;
	const uint32_t base = c == 0 ? o.c0 : c == 1 ? o.c1 : c == 2 ? o.c2 : o.c3;
	const uint32_t c_base = c == 0 ? ix.L2[0] : c == 1 ? ix.L2[1] : c == 2 ? ix.L2[2] : ix.L2[3];
	return c_base + base + __popcll(h & l & m);
}

} // namespace bwagpu
