// kernels.cuh -- K0 (index re-layout), K2 (width), K3 (gapped search), K4 (SA gather).
// Hand-written for sm_100a; integer-only (every floating-point decision of the reference
// -- bwa_cal_maxdiff, bwtaln.c:37-49 -- is taken on the host and shipped as an integer).
#pragma once
#include "fmindex.cuh"

namespace bwagpu {

#define STATE_M 0u
#define STATE_I 1u
#define STATE_D 2u
#define NIL 0xffffffffu
#define WB_EQ 0x8000u
#define WB_BID 0x7fffu
// width sub-arrays start on 8-entry boundaries so that K2 can flush them with 16/32-byte stores
#define WSTRIDE(n) ((((size_t)(n) + 1) + 7) & ~(size_t)7)
#define ARENA_CHUNK_LOG 10
#define ARENA_CHUNK (1u << ARENA_CHUNK_LOG)

// ------------------------------------------------------------------ K0: re-layout
// raw = the reference's bwt_t::bwt (12-word blocks, bwtmisc.c:125-152) on the device.
__global__ void k_relayout(const uint32_t *__restrict__ raw, uint32_t seq_len, uint32_t n_blk, uint4 *__restrict__ blk,
                           uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_blk) return;
	const uint64_t p0 = (uint64_t)b << 6;
	uint32_t c[4] = {t0, t1, t2, t3}; // totals: used by the terminal block when seq_len % 64 == 0
	uint64_t lo = 0, hi = 0;
	if (p0 < seq_len) {
		const uint32_t rb = b >> 1, half = b & 1u;
		const uint32_t *q = raw + (size_t)rb * 12;
		c[0] = q[0]; c[1] = q[1]; c[2] = q[2]; c[3] = q[3];
		const uint32_t n_words = (seq_len + 15) >> 4; // words the reference emitted in total
		for (uint32_t w = 0; w < 8; ++w) {
			const uint32_t gw = rb * 8 + w; // global word index
			const uint32_t x = gw < n_words ? q[4 + w] : 0u;
			if (half == 1 && w < 4) { // second half: fold the first 64 bases into the counts
				for (int t = 0; t < 16; ++t) {
					const uint64_t pos = (uint64_t)gw * 16 + t;
					if (pos < seq_len) ++c[(x >> ((15 - t) << 1)) & 3u];
				}
			} else if ((w >> 2) == half) {
				for (int t = 0; t < 16; ++t) {
					const uint32_t base = (x >> ((15 - t) << 1)) & 3u;
					const int j = (int)(w & 3u) * 16 + t;
					lo |= (uint64_t)(base & 1u) << j;
					hi |= (uint64_t)(base >> 1) << j;
				}
			}
		}
	}
	blk[2 * (size_t)b] = make_uint4(c[0], c[1], c[2], c[3]);
	blk[2 * (size_t)b + 1] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// ------------------------------------------------------------------ batch description
struct GapOpt { // integer subset of gap_opt_t (bwtaln.h:143-153)
	int s_mm, s_gapo, s_gape, mode;
	int indel_end_skip, max_del_occ, max_entries;
	int max_gape, max_seed_diff, seed_len, max_top2;
};

struct ReadMeta {        // 16 B per read, filled by the host
	uint32_t seq_off;    // into the packed base array (chunk-local)
	uint32_t w_off;      // into the width arena (entries)
	uint16_t len;
	uint8_t max_diff;    // local_opt.max_diff for this read (bwtaln.c:102,126)
	uint8_t max_gapo;    // after the clamp of bwtaln.c:103
	uint32_t n_amb;      // bases > 3 in the read (the "too many N" test of bwtgap.c:118-123)
};

struct Batch {
	DevIndex ix[2];
	GapOpt opt;
	int n_reads;
	const uint8_t *__restrict__ seq; // byte j of a read: seq[0][j] | seq[1][j] << 4
	const ReadMeta *__restrict__ meta;
	uint32_t *w;       // width arena: w values (only gap_shadow reads them back)
	uint16_t *bid;     // width arena: bid | (w[i-1] == w[i]) << 15 -- all the pruning tests need
	uint2 *ctx;        // context arena (k_ctx, used by k_search_warp): everything a popped node at (strand, i) needs, in one 8-byte word
	uint16_t *ctx16;   // compact context (k_ctx16, used by k_search): one 16-bit entry per (strand, position)
	// results, per read
	int32_t *n_aln;       // -1 = not done (ran out of arena / pool: retried in the next pass)
	int32_t *max_entries;
	uint32_t *pool_off;   // offset of the read's alns in the unordered pool
	uint4 *pool;
	uint32_t pool_cap;
	unsigned int *pool_count;
	// work list
	const int32_t *jobs; // NULL = identity
	int n_jobs;
	int *work_counter;
	int32_t *overflow_ids;
	int *overflow_count;
	// per-slot scratch
	uint4 *ent;
	uint32_t *nxt;
	uint32_t *heads; // global bucket heads (only when BWAGPU_SMEM_HEADS == 0)
	uint32_t cap, n_stacks;
	uint32_t pop_cap;  // pass 0 only: a read that pops more nodes than this is handed to the next pass (0 = no limit)
	// arena overflow: chunks of ARENA_CHUNK records from a pool shared by all threads of the launch.  A
	// thread takes chunks as its search deepens and hands all but one back when the read is finished
	// (lock-free stack with a version tag), so the pool only has to cover the reads in flight
	uint4 *xent;
	uint32_t *xnxt;
	uint32_t *ctab;            // per thread: ids of the chunks it owns
	uint32_t ctab_stride;      // max chunks per thread
	uint32_t x_chunks;         // chunks in the pool
	unsigned int *x_next;      // bump counter (chunks never handed out yet)
	unsigned long long *x_free_top; // recycled chunks: lock-free stack, {tag:32 | head:32} against ABA
	uint32_t *x_free_next;     // link of the recycled-chunk stack
	// stats (STATS builds only)
	// [32 + d] occurrence lookups at depth d = min(31, len - i); [64 + b] reads whose search popped [2^b, 2^(b+1)) nodes
	unsigned long long *stats; // [10] pops served from the arena (memory)  [0] ref fetches [1] own fetches [2] pops [3] pushes [4] records stored [5] pruned pops [6] expansions [7] exact-tail steps [8] derive trips
};

__device__ __forceinline__ unsigned long long gtime()
{
#ifdef BWAGPU_HOST_EMU
	return 0;
#else
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
#endif
}

__device__ __forceinline__ uint32_t sel4(uint32_t c, const uint32_t v[4])
{
	return c == 0 ? v[0] : c == 1 ? v[1] : c == 2 ? v[2] : v[3];
}

// ------------------------------------------------------------------ K2: width
// bwt_cal_width (bwtaln.c:52-76) for the 4 strings of each read: thread = (read, a, seed).
// w[a] is computed on index a with seq[a] (bwtaln.c:123-124), the seed widths on the last
// seed_len bases (128-129).  4 independent dependent chains per read.
template <bool STATS>
__global__ void __launch_bounds__(128) k_width(const Batch B)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int job = t >> 2;
	uint32_t f_ref = 0, f_own = 0;
	if (job < B.n_jobs) {
		// later passes recompute the widths of their reads: gap_shadow (bwtgap.c:81-91) edited
		// them in place during the attempt that overflowed
		const int r = B.jobs ? B.jobs[job] : job;
		const ReadMeta m = B.meta[r];
		const int a = t & 1, seed = (t >> 1) & 1;
		const int len = m.len;
		const bool has_seed = len > B.opt.seed_len;
		if (len > 0 && (!seed || has_seed)) {
			const int n = seed ? B.opt.seed_len : len;
			const uint8_t *s = B.seq + m.seq_off + (seed ? len - n : 0);
			const size_t wo = (size_t)m.w_off + (seed ? 2 * WSTRIDE(len) + (size_t)a * WSTRIDE(n) : (size_t)a * WSTRIDE(len));
			uint32_t *w = B.w + wo;
			uint16_t *bd = B.bid + wo;
			const DevIndex &ix = B.ix[a];
			uint32_t k = 0, l = ix.seq_len;
			uint32_t bid = 0, prev_w = 0;
			// results are collected 8 at a time and flushed as one 32-byte (w) and one 16-byte (wb)
			// store: a 4- or 2-byte store per step costs a whole sector of L2 write bandwidth each
			uint32_t wq[8], bq[4];
			for (int i0 = 0; i0 <= n; i0 += 8) {
#pragma unroll
				for (int t = 0; t < 8; ++t) {
					const int i = i0 + t;
					uint32_t wi = 0, bv = 0;
					if (i < n) {
						const uint32_t c = (s[i] >> (a << 2)) & 15u;
						if (c < 4) {
							uint32_t ok, ol;
							occ1_pair<STATS>(ix, k - 1, l, c, ok, ol, f_ref, f_own);
							k = ix.L2[c] + ok + 1;
							l = ix.L2[c] + ol;
						}
						if (k > l || c > 3) { // restart
							k = 0; l = ix.seq_len; ++bid;
						}
						wi = l - k + 1;
						bv = bid | ((i > 0 && wi == prev_w) ? WB_EQ : 0u); // bid | (w[i-1] == w[i]) << 15
						prev_w = wi;
					} else if (i == n) { // the terminator entry (bwtaln.c:73-74)
						wi = 0;
						bv = (bid + 1) | (prev_w == 0 ? WB_EQ : 0u);
					}
					wq[t] = wi;
					if (t & 1) bq[t >> 1] |= bv << 16; else bq[t >> 1] = bv;
				}
				*reinterpret_cast<uint4 *>(w + i0) = make_uint4(wq[0], wq[1], wq[2], wq[3]);
				*reinterpret_cast<uint4 *>(w + i0 + 4) = make_uint4(wq[4], wq[5], wq[6], wq[7]);
				*reinterpret_cast<uint4 *>(bd + i0) = make_uint4(bq[0], bq[1], bq[2], bq[3]);
			}
		}
	}
	if (STATS) {
		atomicAdd(B.stats + 0, (unsigned long long)f_ref);
		atomicAdd(B.stats + 1, (unsigned long long)f_own);
	}
}

// ------------------------------------------------------------------ K2b: per-position context words
// A node popped at position i of strand a needs width[i-1], width[i-2].bid, the two seed-width entries
// and the read bases i-1 and i-2 (bwtgap.c:156-157, 206-215, 160-165) before it can be pruned or
// expanded: six scattered 1- and 2-byte loads.  k_ctx packs them into ONE 8-byte word per (strand, i),
// indexed like the width arena, so that a trip of K3 issues a single context load next to its two
// occurrence blocks:
//   .x = CW(wb[i-1]) | str[i-1] << 12 | (CW(wb[i-2]) & 0xfff | str[i-2] << 12) << 16
//   .y = CW(swb[si]) | (CW(swb[si-1]) & 0xfff) << 16        si = (i-1) - (len - seed_len) >= 1, else 0
// CW(v) = min(bid, 4095) | EQ << 15: bid is only ever compared with values <= max_diff <= 254.
#define CW_BID 0x0fffu
#define CW_KEEP 0x70007000u // the two base fields of .x
__device__ __forceinline__ uint32_t cw_of(uint32_t v)
{
	const uint32_t b = v & WB_BID;
	return (b > CW_BID ? CW_BID : b) | (v & WB_EQ);
}

__global__ void __launch_bounds__(256) k_ctx(const Batch B)
{
#ifdef BWAGPU_HOST_EMU
	const int job = (int)blockIdx.x, lane = 0, nl = 1;
#else
	const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int job = (int)(gt >> 5), lane = (int)(gt & 31), nl = 32;
#endif
	if (job >= B.n_jobs) return;
	const int r = B.jobs ? B.jobs[job] : job;
	const ReadMeta m = B.meta[r];
	const int len = m.len, seed_len = B.opt.seed_len;
	if (len == 0) return;
	const bool has_seed = len > seed_len;
	const uint8_t *s = B.seq + m.seq_off;
	for (int a = 0; a < 2; ++a) {
		const uint16_t *wb = B.bid + m.w_off + (size_t)a * WSTRIDE(len);
		const uint16_t *swb = B.bid + m.w_off + 2 * WSTRIDE(len) + (size_t)a * WSTRIDE(seed_len);
		uint2 *cx = B.ctx + m.w_off + (size_t)a * WSTRIDE(len);
		for (int i = lane; i <= len; i += nl) {
			uint32_t lo = 0, hi = 0;
			if (i >= 1) lo = cw_of(wb[i - 1]) | ((uint32_t)(s[i - 1] >> (a << 2)) & 7u) << 12;
			if (i >= 2) lo |= ((cw_of(wb[i - 2]) & CW_BID) | ((uint32_t)(s[i - 2] >> (a << 2)) & 7u) << 12) << 16;
			if (has_seed) {
				const int si = (i - 1) - (len - seed_len);
				if (si >= 1) hi = cw_of(swb[si]) | (cw_of(swb[si - 1]) & CW_BID) << 16;
			}
			cx[i] = make_uint2(lo, hi);
		}
	}
}

// ------------------------------------------------------------------ K2c: compact per-position context for k_search
// The 8-byte words above cost a 32-byte sector per popped node and, for the reads in flight, more bytes than L2 holds
// (132 k threads x 2 strands x ~100 positions x 8 B = 210 MB): every pop was a DRAM sector of which a quarter was used.
// k_search reads ONE 16-bit entry per (strand, position t) instead,
//     E[t] = min(bid(w[t]), 127) | (w[t-1] == w[t]) << 7 | min(bid(seed_w[t - off]), 15) << 8 | (seed EQ) << 12 | str[t] << 13
// (off = len - seed_len; the seed fields are 0 before the seed window), indexed like the width arena.  A node popped at
// position i looks at E[i-1] and E[i-2] -- two 2-byte loads from the same sector -- and a read's two strands are ~400 bytes
// = 13 sectors in all, which stay in L1/L2 for the whole search.  bid saturates at 127 / 15: it is only ever compared with
// values <= max_diff <= 126 and <= max_seed_diff <= 14 (checked on the host).
#define C16_BID 0x7fu
#define C16_EQ 0x80u
#define C16_SBID_SHIFT 8
#define C16_SEQ 0x1000u
#define C16_BASE_SHIFT 13
#define C16_KEEP 0xff00u // seed fields + base: what gap_shadow never changes
__device__ __forceinline__ uint32_t c16_width(uint32_t v) // bid | EQ << 15  ->  the width fields of an entry
{
	const uint32_t b = v & WB_BID;
	return (b > C16_BID ? C16_BID : b) | ((v & WB_EQ) ? C16_EQ : 0u);
}

__global__ void __launch_bounds__(256) k_ctx16(const Batch B)
{
#ifdef BWAGPU_HOST_EMU
	const int job = (int)blockIdx.x, lane = 0, nl = 1;
#else
	const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int job = (int)(gt >> 5), lane = (int)(gt & 31), nl = 32;
#endif
	if (job >= B.n_jobs) return;
	const int r = B.jobs ? B.jobs[job] : job;
	const ReadMeta m = B.meta[r];
	const int len = m.len, seed_len = B.opt.seed_len;
	if (len == 0) return;
	const bool has_seed = len > seed_len;
	const int off = len - seed_len;
	const uint8_t *s = B.seq + m.seq_off;
	for (int a = 0; a < 2; ++a) {
		const uint16_t *wb = B.bid + m.w_off + (size_t)a * WSTRIDE(len);
		const uint16_t *swb = B.bid + m.w_off + 2 * WSTRIDE(len) + (size_t)a * WSTRIDE(seed_len);
		uint16_t *cx = B.ctx16 + m.w_off + (size_t)a * WSTRIDE(len);
		for (int t = lane; t < len; t += nl) {
			uint32_t e = c16_width(wb[t]) | ((uint32_t)(s[t] >> (a << 2)) & 7u) << C16_BASE_SHIFT;
			if (has_seed && t >= off) {
				const uint32_t sv = swb[t - off], sb = sv & WB_BID;
				e |= (sb > 15u ? 15u : sb) << C16_SBID_SHIFT | ((sv & WB_EQ) ? C16_SEQ : 0u);
			}
			cx[t] = (uint16_t)e;
		}
	}
}

// ------------------------------------------------------------------ K3: gapped search
// bwt_match_gap (bwtgap.c:104-266) with its gap_stack (bwtgap.c:13-79), one read per
// thread, persistent threads pulling reads from a global counter.
//
// Exactness: the traversal order is the reference's -- lowest non-empty score bucket
// first, LIFO inside a bucket, children in the order insertion / deletions c=0..3 /
// mismatches c=(str[i]+j)&3 j=1..4.  Buckets are singly linked lists threaded through a
// per-thread arena (the reference uses realloc-doubling arrays; only the order is
// observable).  What is stored differs from the reference, what is popped does not:
//   * the match continuation (pushed last, same score as the entry just popped, hence
//     always the very next pop) is carried in registers (`held`) and never stored;
//   * GROUP RECORDS: the children one expansion sends to the same bucket (the gap children:
//     insertion + deletions; the mismatch children) are consecutive in that bucket's LIFO,
//     so they are stored as ONE 16-byte record = the parent node + a bit mask of the
//     children still to come.  Popping takes the highest set bit (the child pushed last),
//     rewrites the mask in place (the record stays on top) and re-derives the child's
//     SA interval from the parent's with one extra occurrence lookup -- and only if the
//     child survives the cheap pruning tests.  An expansion therefore stores at most two
//     records instead of up to nine entries;
//   * PHANTOMS: once the first hit fixed best_score, children scoring above best_score +
//     s_mm can only ever end the search when popped (bwtgap.c:144); they are counted but
//     not stored;
//   * the arena slot freed by the latest pop is kept in a register for the next push;
//   * the read's HITS are records of the same arena, chained in discovery order (the duplicate
//     test of bwtgap.c:179-183 walks the chain), so a read may have any number of them.
// n_entries counts every child individually, stored or not, so max_entries and the
// `> opt->max_entries` stop (bwtgap.c:139-140) are the reference's.
//
// Memory: a small private arena per thread; when a search outgrows it, the arena continues in
// 1024-record chunks of a pool shared by the launch (POOLED instantiation; chunks are recycled
// through a lock-free stack when the read ends).  Reads that run out of arena (pass 0) or find
// the pool dry (pass 1) are flagged and retried from scratch by the next pass (bwagpu.cu).
//
// The per-thread control flow is a small state machine so that the lanes of a warp meet
// at ONE occurrence lookup per trip whatever each lane is doing (expanding a node,
// deriving a group child, or walking the exact-match tail of bwt.c:237-252).
struct Entry {
	uint32_t k, l;
	uint32_t pos;  // plain: i | last_diff_pos << 16      group: parent i (after --i) | child mask << 16
	uint32_t tag;  // n_mm | n_gapo << 8 | n_gape << 16 | state << 24 | a << 26 | kind << 27 | (KIND_MM: read base at i) << 29
};

#define E_I(e) ((int)((e).pos & 0xffffu))
#define E_LDP(e) ((int)((e).pos >> 16))
#define E_MM(e) ((int)((e).tag & 0xffu))
#define E_GO(e) ((int)(((e).tag >> 8) & 0xffu))
#define E_GE(e) ((int)(((e).tag >> 16) & 0xffu))
#define E_ST(e) (((e).tag >> 24) & 3u)
#define E_A(e) (((e).tag >> 26) & 1u)
#define KIND_PLAIN 0u
#define KIND_GAP 1u // bit 0 = insertion (child i = parent i), bit 1+c = deletion of c (child i = parent i + 1)
#define KIND_MM 2u  // bit j-1 = mismatch child c = (str[i] + j) & 3

enum { MODE_NEW = 0, MODE_POP = 1, MODE_POPWAIT = 2, MODE_DERIVE = 3, MODE_EXACT = 4, MODE_EXPAND = 5, MODE_DONE = 6 };

#ifndef BWAGPU_MINBLOCKS
#define BWAGPU_MINBLOCKS 7 // __launch_bounds__ second argument: 7 blocks per SM = 72 registers, no spills (6 at the 79 it takes unbounded: 148.3 ms, 7: 143.0, 8 = 64 registers with spills: 160.2; C4, 2 M reads)
#endif
// Measured on B200, 10M x 76bp (profiles/r1_ab_experiments.md).  Bucket heads in shared memory were a wash
// while six context loads per node competed for L1; with the single context word (k_ctx) they win 5 %
// and are the default.  An L1 prefetch of the next record to pop is slower; kept as a switch.
#ifndef BWAGPU_SMEM_HEADS
#define BWAGPU_SMEM_HEADS 1
#endif
#ifndef BWAGPU_PREFETCH_TOP
#define BWAGPU_PREFETCH_TOP 0
#endif
#ifndef BWAGPU_STREAM_STACK
#define BWAGPU_STREAM_STACK 0 // 1: stack records use streaming (evict-first) loads/stores (A/B switch)
#endif
#ifndef BWAGPU_BATCH_POP
#define BWAGPU_BATCH_POP 0 // N > 0 (needs BWAGPU_CONVERGE=1): memory pops wait until N lanes of the warp want one
#endif
#ifndef BWAGPU_TOP_REG
#define BWAGPU_TOP_REG 0 // 1: the record pushed last is kept in registers until the next push displaces it
#endif
#ifndef BWAGPU_ARENA_G
#define BWAGPU_ARENA_G 7 // log2 of the arena interleaving granule (records)
#endif
// elements to allocate per thread for a private arena of `cap` records (whole granules)
#define ARENA_ALLOC(cap) ((((size_t)(cap)) + ((1u << BWAGPU_ARENA_G) - 1u)) & ~(size_t)((1u << BWAGPU_ARENA_G) - 1u))
#ifndef BWAGPU_NO_FREELIST
#define BWAGPU_NO_FREELIST 1 // 1: pass 0 recycles only the slot of the latest pop (no free list; deeper reads go to pass 1)
#endif
#ifndef BWAGPU_CTX_SMEM
#define BWAGPU_CTX_SMEM 1 // 1: each thread keeps the 32-byte sector of context entries it read last in shared memory
#endif
// bytes of dynamic shared memory per block of 128 threads: the bucket heads (+ the hit list's two ends), then the context sectors
#define SEARCH_SMEM_HEADS(n_stacks, head_bytes) ((((size_t)128 * ((n_stacks) + 2) * (head_bytes)) + 15) & ~(size_t)15)
#define SEARCH_SMEM(n_stacks, head_bytes) ((BWAGPU_SMEM_HEADS ? SEARCH_SMEM_HEADS(n_stacks, head_bytes) : 0) + (size_t)128 * 32 + (size_t)128 * 16) // + context sectors + pop cache
#ifndef BWAGPU_PREFETCH_HELD
#define BWAGPU_PREFETCH_HELD 0 // 1: an expansion prefetches (L1) the two index blocks of its match child before it pushes the other children; measured 182.5 against 142.0 ms (a prefetch is one more request, and requests are what the kernel pays for): off
#endif
#ifndef BWAGPU_POP_CACHE
#define BWAGPU_POP_CACHE 1 // 1 (needs BWAGPU_EMBED_NXT): the group record a child was just popped from stays in shared memory while it has children left
#endif
#if BWAGPU_POP_CACHE && BWAGPU_SPLIT_POP
#error "BWAGPU_POP_CACHE and BWAGPU_SPLIT_POP are alternatives"
#endif
#ifndef BWAGPU_EMBED_NXT
#define BWAGPU_EMBED_NXT 1 // 1: pass 0 keeps a record's list link inside the record (no second array, one request per push / pop)
#endif
#if BWAGPU_EMBED_NXT && !BWAGPU_NO_FREELIST
#error "BWAGPU_EMBED_NXT needs BWAGPU_NO_FREELIST (pass 0 keeps no free list)"
#endif
#ifndef BWAGPU_SPLIT_POP
#define BWAGPU_SPLIT_POP 0 // 1: a pop from the arena is split over two trips (load issued in one, record used in the next); measured 184 ms against 180 ms on C4 (profiles/r2_ab_experiments.md): off
#endif
#if BWAGPU_SPLIT_POP && BWAGPU_TOP_REG
#error "BWAGPU_SPLIT_POP and BWAGPU_TOP_REG are alternatives"
#endif
#ifndef BWAGPU_CONVERGE
#define BWAGPU_CONVERGE 0 // 1: lanes stay in the loop until the whole warp is done and re-converge every trip
#endif

// POOLED = false: the private arena only (no chunk indirection anywhere in the loop) -- pass 0, where
// nearly every read of an ordinary workload finishes.  POOLED = true: the arena continues in chunks of
// the shared pool -- the passes that pick up the reads whose search went deeper.
// STDMODE = true: gap_opt_t::mode has BWA_MODE_GAPE set and LOGGAP / NONSTOP clear (the default of
// gap_init_opt, bwtaln.c:19-35, and of every BASELINE.json config): the three tests fold at compile time.
template <bool POOLED> struct HeadT { typedef uint32_t type; };
#if !defined(BWAGPU_HOST_EMU)
template <> struct HeadT<false> { typedef uint16_t type; }; // pass 0: private arena of <= 65535 records
#endif

// Non-empty score buckets.  The general form covers 256 buckets (the ABI's limit); SMALL covers scores
// 0..63 in one word -- enough for every default-option workload (76 bp, -n 0.04 -o 1: scores <= 27) -- and
// is what pass 0 uses: a push to a bucket >= 64 sends the read to the next pass like an arena overflow.
template <bool SMALL> struct BucketMask {
	uint64_t m0, m1, m2, m3;
	__device__ __forceinline__ void reset() { m0 = m1 = m2 = m3 = 0; }
	__device__ __forceinline__ bool fits(int) const { return true; }
	__device__ __forceinline__ void set(int s)
	{
		const uint64_t bit = 1ull << (s & 63);
		const int w = s >> 6;
		m0 |= w == 0 ? bit : 0ull; m1 |= w == 1 ? bit : 0ull; m2 |= w == 2 ? bit : 0ull; m3 |= w == 3 ? bit : 0ull;
	}
	__device__ __forceinline__ void clear(int s)
	{
		const uint64_t bit = 1ull << (s & 63);
		const int w = s >> 6;
		m0 &= ~(w == 0 ? bit : 0ull); m1 &= ~(w == 1 ? bit : 0ull); m2 &= ~(w == 2 ? bit : 0ull); m3 &= ~(w == 3 ? bit : 0ull);
	}
	__device__ __forceinline__ bool test(int s) const
	{
		const uint64_t w = s < 64 ? m0 : s < 128 ? m1 : s < 192 ? m2 : m3;
		return (w >> (s & 63)) & 1ull;
	}
	__device__ __forceinline__ bool any() const { return (m0 | m1 | m2 | m3) != 0; }
	__device__ __forceinline__ int lowest() const
	{
		if (m0) return __ffsll((long long)m0) - 1;
		if (m1) return 64 + __ffsll((long long)m1) - 1;
		if (m2) return 128 + __ffsll((long long)m2) - 1;
		return 192 + __ffsll((long long)m3) - 1;
	}
};
template <> struct BucketMask<true> {
	uint64_t m0;
	__device__ __forceinline__ void reset() { m0 = 0; }
	__device__ __forceinline__ bool fits(int s) const { return s < 64; }
	__device__ __forceinline__ void set(int s) { m0 |= 1ull << s; }
	__device__ __forceinline__ void clear(int s) { m0 &= ~(1ull << s); }
	__device__ __forceinline__ bool test(int s) const { return (m0 >> s) & 1ull; }
	__device__ __forceinline__ bool any() const { return m0 != 0; }
	__device__ __forceinline__ int lowest() const { return __ffsll((long long)m0) - 1; }
};

// Per-thread state is kept deliberately small (registers decide how many reads an SM has in flight, and the
// kernel is latency-bound): the node being processed is (k, l, e_pos, e_tag) with no second copy; the match
// continuation overwrites k/l in place and is marked by one flag; the context word stays packed; per-read
// constants share one register; the hit list's ends live beside the bucket heads.
template <bool STATS, bool POOLED, bool STDMODE>
__global__ void __launch_bounds__(128, BWAGPU_MINBLOCKS) k_search(const Batch B)
{
	typedef typename HeadT<POOLED>::type head_t;
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	// private arenas are interleaved across the grid in granules of 2^BWAGPU_ARENA_G records: records
	// [g * G, (g + 1) * G) of thread `slot` sit at ((g * NT) + slot) * G.  A thread's neighbouring records still
	// share sectors, and the low indices every search keeps re-using form ONE dense region for the whole grid
	// instead of one hot spot per 32 KB arena (fewer pages in the TLB, even use of L2 sets)
	const size_t NT = (size_t)gridDim.x * blockDim.x;
	constexpr uint32_t AG = BWAGPU_ARENA_G, AM = (1u << BWAGPU_ARENA_G) - 1u;
	uint4 *const ent = B.ent + ((size_t)slot << AG);
	uint32_t *const nxt = B.nxt + ((size_t)slot << AG);
	// bucket list heads: shared memory, bucket-major (heads[s * blockDim + tid]); two more slots per thread
	// hold the first and last record of the read's hit list
#ifdef BWAGPU_HOST_EMU
	static head_t s_heads[264];
	head_t *const heads = s_heads + threadIdx.x;
	const uint32_t HS = blockDim.x; // stride between buckets
	static uint32_t s_cxw[8], s_pcw[4];
	uint32_t *const cxw = s_cxw; // context sector: word w of this thread at cxw[w * CS]
	uint32_t *const pcw = s_pcw; // pop cache: word w at pcw[w * CS]
#elif BWAGPU_SMEM_HEADS
	extern __shared__ __align__(16) unsigned char s_heads_raw[];
	head_t *const heads = reinterpret_cast<head_t *>(s_heads_raw) + threadIdx.x;
	const uint32_t HS = blockDim.x;
	uint32_t *const cxw = reinterpret_cast<uint32_t *>(s_heads_raw + SEARCH_SMEM_HEADS(B.n_stacks, sizeof(head_t))) + threadIdx.x;
	uint32_t *const pcw = cxw + 8 * blockDim.x;
#else
	extern __shared__ __align__(16) unsigned char s_heads_raw[];
	uint32_t *const heads = B.heads + (size_t)slot * (B.n_stacks + 2);
	const uint32_t HS = 1;
	uint32_t *const cxw = reinterpret_cast<uint32_t *>(s_heads_raw) + threadIdx.x;
	uint32_t *const pcw = cxw + 8 * blockDim.x;
#endif
	const uint32_t CS = blockDim.x; // stride between the words of a thread's context sector
	// CONTEXT SECTOR.  Every popped node reads two context entries (k_ctx16), and the match chain walks them downwards one
	// position per trip -- but with ~900 searches per SM a sector does not survive in L1 from one trip to the next: nearly
	// every pop was a request of its own to L2 (a third of the kernel's read requests, and the kernel is bound by their
	// number).  Each thread therefore keeps the last sector it fetched (16 entries) in shared memory; cx_tag = which one.
	uint32_t cx_tag = 0xffffffffu;
	const uint32_t HIT_HEAD = B.n_stacks * HS, HIT_TAIL = (B.n_stacks + 1) * HS; // valid while n_aln > 0
	const GapOpt &O = B.opt;
	const bool gape_mode = STDMODE || (O.mode & 0x01), loggap = !STDMODE && (O.mode & 0x04), nonstop = !STDMODE && (O.mode & 0x10);
	// The reversed genome has the forward genome's base composition, so C() (bwt_t::L2) and
	// seq_len are the same for both indexes; only the block array and `primary` differ.
	const uint32_t C1 = B.ix[0].L2[1], C2 = B.ix[0].L2[2], C3 = B.ix[0].L2[3];

	// record idx -> address: the private arena first, then this thread's pool chunks
	const uint32_t CAP0 = B.cap;
	uint32_t *const ctab = B.ctab + (size_t)slot * B.ctab_stride;
	uint32_t n_chunks = 0; // pool chunks owned by this thread right now
	auto ent_at = [&](uint32_t idx) -> uint4 * {
		if (!POOLED || idx < CAP0) return ent + (((size_t)(idx >> AG) * NT) << AG) + (idx & AM);
		const uint32_t o = idx - CAP0;
		return B.xent + ((size_t)ctab[o >> ARENA_CHUNK_LOG] << ARENA_CHUNK_LOG) + (o & (ARENA_CHUNK - 1));
	};
	auto nxt_at = [&](uint32_t idx) -> uint32_t * {
		if (!POOLED || idx < CAP0) return nxt + (((size_t)(idx >> AG) * NT) << AG) + (idx & AM);
		const uint32_t o = idx - CAP0;
		return B.xnxt + ((size_t)ctab[o >> ARENA_CHUNK_LOG] << ARENA_CHUNK_LOG) + (o & (ARENA_CHUNK - 1));
	};

	// EMBEDDED LINKS (pass 0).  k_search is bound by the number of memory requests it makes (profiles/r2_fetch_probe.md), and
	// a push used to make two (the 16-byte record and its 4-byte link in a second array), a pop from the arena two more.
	// Pass 0's arena has at most 65535 records and its reads at most 255 bases (longer ones go to the next pass), so the
	// link fits into the record: word z = i | (last_diff_pos or child mask) << 8 | link << 16; a hit record keeps its link
	// in the upper half of its score word.  0xffff = end of list.
	constexpr bool EMB = BWAGPU_EMBED_NXT && !POOLED;
	auto link16 = [](uint32_t idx) -> uint32_t { return idx == NIL ? 0xffffu : idx; };
	auto unlink16 = [](uint32_t v) -> uint32_t { return v == 0xffffu ? NIL : v; };
	// POP CACHE (pass 0).  A group record stays on top of its bucket while its children are popped one by one, and between
	// two of them the thread pops nothing else from the arena (a child's own pushes go to buckets of higher score when all
	// penalties are positive): every child used to cost a load of the record and a store of its shrunken mask.  The record
	// now moves to shared memory with the first child and is served from there (pend_idx = which record; its mask in the
	// arena goes stale and is written back -- one byte -- only if another group takes the slot while it has children left).
	constexpr bool PC = BWAGPU_POP_CACHE && EMB;
	uint32_t pend_idx = NIL;

	int mode = MODE_NEW;
	// per-read state
	int rid = -1, max_diff = 0;
	uint32_t rd = 0; // len | max_gapo << 16 | local_opt.max_diff << 24   (ReadMeta)
#define RD_LEN ((int)(rd & 0xffffu))
#define RD_GAPO ((int)((rd >> 16) & 0xffu))
#define RD_MAXDIFF ((int)(rd >> 24))
	uint32_t w_off = 0; // the read's offset into the width / context arenas
	int best_score = 0, best_cnt = 0, n_aln = 0, n_entries = 0, max_entries = 0;
	uint32_t pops_left = 0; // pass 0: nodes this read may still pop before it is handed to the next pass
	uint32_t read_pops = 0; // STATS
	bool overflow = false;
	// Bucket lists.  heads[s] in memory is only meaningful while bit s of the mask is set, and
	// the head of the bucket popped last lives in a register (cur_s / cur_head), so neither a
	// reset pass nor a head load per pop is needed.
	BucketMask<!POOLED> mask;
	mask.reset();
	int cur_s = -1;
	uint32_t cur_head = NIL;
	uint32_t bump = 0, free_head = NIL, spare = NIL;
	// the node being processed: interval (k, l), e_pos = i | last_diff_pos << 16, e_tag = n_mm | n_gapo << 8 |
	// n_gape << 16 | state << 24 | a << 26.  held: the match child of the node just expanded is the next pop
	// (pushed last, same score); its interval is already in k/l, its position in i, its counts are e_tag's.
	uint32_t k = 0, l = 0, e_pos = 0, e_tag = 0;
	bool held = false;
	uint32_t cc = 0; // MODE_DERIVE: which child interval of (k, l) to take; MODE_EXACT: the base at i-1
#if BWAGPU_TOP_REG
	// One record lives in registers instead of the arena (lr_s = its bucket, < 0: none): the most recent push into the
	// lowest bucket pushed to so far.  A push into the same or a lower bucket displaces it (the old one is written to the
	// arena first, so it stays below the new one); a push into a higher bucket goes to the arena directly.  The register
	// record is therefore always the top of its bucket, and a pop takes it whenever no lower bucket has records in
	// memory -- the common case at the end of a path (the mismatch group of the last expansion).  The remaining
	// children of a group are then popped from registers too: their own pushes all go to higher buckets.
	uint32_t lr_k = 0, lr_l = 0, lr_pos = 0, lr_tag = 0;
	int lr_s = -1;
#endif
#if BWAGPU_SPLIT_POP
	// SPLIT POP (A/B switch, off).  A record popped from the arena costs its trip two dependent memory round trips (the
	// record, then the occurrence blocks of its interval), and with a few lanes of every warp popping from memory in every
	// trip the whole warp pays both (ncu: 15 % of all stall samples at the record's first use, 4.8 lanes active).  With the
	// switch on such a lane only ISSUES the record's loads (MODE_POPWAIT, not active in this trip's lookup) and takes the
	// record up at the top of the next trip: its latency passes under the wait for the other lanes' occurrence blocks.  The
	// lane's stack cannot change in between (only the lane itself pushes to it), so the pop is the reference's.  Measured
	// slower (the lane's extra trip and 7 more registers cost more than the shorter wait saves).
	uint4 pq = make_uint4(0u, 0u, 0u, 0u);
	uint32_t pnx = NIL;
#endif
	int m = 0, i = 0; // i doubles as the exact tail's cursor
	// context of the current node, loaded together with its occurrence blocks (k_ctx16): cw = E[i-1] | E[i-2] << 16, i.e.
	// width[i-1] (bid, EQ), width[i-2].bid, the two seed-width entries, str[i-1], str[i-2]
	uint32_t cw = 0u;
	const uint16_t *const ctx16 = B.ctx16;
#define CW_BID1 ((int)(cw & C16_BID))                       /* width[i-1].bid */
#define CW_EQ1 (cw & C16_EQ)                                 /* width[i-2].w == width[i-1].w */
#define CW_C1 ((cw >> C16_BASE_SHIFT) & 7u)                  /* str[i-1] */
#define CW_B2 ((int)((cw >> 16) & C16_BID))                  /* width[i-2].bid */
#define CW_CN ((cw >> (16 + C16_BASE_SHIFT)) & 7u)           /* str[i-2] */
#define CW_SBID1 ((int)((cw >> C16_SBID_SHIFT) & 15u))       /* seed_width[si].bid */
#define CW_SEQ1 (cw & C16_SEQ)
#define CW_S2 ((int)((cw >> (16 + C16_SBID_SHIFT)) & 15u))   /* seed_width[si-1].bid */
	uint32_t f_ref = 0, f_own = 0, n_pops = 0, n_pushes = 0, n_stored = 0, n_pruned = 0, n_expand = 0, n_exact = 0, n_derive = 0, n_trips = 0, n_mempop = 0;
	if (STATS) atomicMin(B.stats + 12, gtime());

	auto score_of = [&](int mm, int go, int ge) { return mm * O.s_mm + go * O.s_gapo + ge * O.s_gape; };

	auto chunk_alloc = [&]() -> uint32_t {
		unsigned long long old = *(volatile unsigned long long *)B.x_free_top;
		while ((uint32_t)old != NIL) { // pop a recycled chunk
			const uint32_t head = (uint32_t)old;
			const uint32_t nx = ((volatile uint32_t *)B.x_free_next)[head];
			const unsigned long long nw = (((old >> 32) + 1ull) << 32) | nx;
			const unsigned long long prev = atomicCAS(B.x_free_top, old, nw);
			if (prev == old) return head;
			old = prev;
		}
		const uint32_t c = atomicAdd(B.x_next, 1u);
		return c < B.x_chunks ? c : NIL;
	};
	auto chunk_free = [&](uint32_t c) {
		unsigned long long old = *(volatile unsigned long long *)B.x_free_top;
		for (;;) {
			((volatile uint32_t *)B.x_free_next)[c] = (uint32_t)old;
			__threadfence();
			const unsigned long long nw = (((old >> 32) + 1ull) << 32) | c;
			const unsigned long long prev = atomicCAS(B.x_free_top, old, nw);
			if (prev == old) return;
			old = prev;
		}
	};

	// one arena record: the slot freed by the latest pop, else (pooled passes) the free list, else fresh space
	// (private arena, then chunks of the shared pool).  NIL + overflow when there is none.  Pass 0 keeps no free
	// list: a slot freed while `spare` is occupied is not reused, and a read that runs through the private arena
	// that way goes to the next pass.
	auto alloc_rec = [&]() -> uint32_t {
		uint32_t idx;
		if (spare != NIL) { idx = spare; spare = NIL; }
		else if ((POOLED || !BWAGPU_NO_FREELIST) && free_head != NIL) { idx = free_head; free_head = *nxt_at(idx); }
		else {
			if (!POOLED) {
				if (bump == CAP0) { overflow = true; return NIL; } // deeper than the private arena: next pass
			} else if (bump == CAP0 + (n_chunks << ARENA_CHUNK_LOG)) { // arena full: take one more chunk from the pool
				const uint32_t c = n_chunks < B.ctab_stride ? chunk_alloc() : NIL;
				if (c == NIL) { overflow = true; return NIL; } // pool dry: retried in the guaranteed pass
				ctab[n_chunks++] = c;
			}
			idx = bump++;
		}
		return idx;
	};

	// links one record into bucket s's list in the arena
	auto store_rec = [&](uint32_t rk, uint32_t rl, uint32_t pos, uint32_t tag, int s) {
		const uint32_t idx = alloc_rec();
		if (idx == NIL) return;
		if (EMB) {
			uint32_t below;
			if (s == cur_s) { below = cur_head; cur_head = idx; }
			else { below = mask.test(s) ? (uint32_t)heads[s * HS] : NIL; heads[s * HS] = (head_t)idx; }
			*ent_at(idx) = make_uint4(rk, rl, (pos & 0xffu) | ((pos >> 16) & 0xffu) << 8 | link16(below) << 16, tag);
		} else {
			*ent_at(idx) = make_uint4(rk, rl, pos, tag);
			if (s == cur_s) { *nxt_at(idx) = cur_head; cur_head = idx; }
			else { *nxt_at(idx) = mask.test(s) ? (uint32_t)heads[s * HS] : NIL; heads[s * HS] = (head_t)idx; }
		}
		mask.set(s);
	};

	// gap_push (bwtgap.c:45-64): one record holding n_children nodes of score s
	auto push_rec = [&](uint32_t rk, uint32_t rl, uint32_t pos, uint32_t tag, int s, int n_children) {
		n_entries += n_children;
		if (STATS) n_pushes += n_children;
		if (n_aln > 0 && !nonstop && s > best_score + O.s_mm) return; // phantoms: counted, never stored
		if (STATS) ++n_stored;
		if (!mask.fits(s)) { overflow = true; return; }
#if BWAGPU_TOP_REG
		if (lr_s >= 0 && s > lr_s) { store_rec(rk, rl, pos, tag, s); return; } // the register keeps the lower bucket's top
		if (lr_s >= 0) store_rec(lr_k, lr_l, lr_pos, lr_tag, lr_s); // displaced (same or higher bucket): now an ordinary arena record
		lr_k = rk; lr_l = rl; lr_pos = pos; lr_tag = tag; lr_s = s;
#else
		store_rec(rk, rl, pos, tag, s);
#endif
	};

	// copies the finished read's results out and resets the per-slot stack
	auto finish_read = [&]() {
		uint32_t off = 0;
		if (STATS && !overflow) atomicAdd(B.stats + 64 + (read_pops ? 31 - __clz((int)read_pops) : 0), 1ull);
		if (!overflow && n_aln > 0) {
			off = atomicAdd(B.pool_count, (unsigned int)n_aln);
			if (off + (uint32_t)n_aln > B.pool_cap) overflow = true;
		}
		if (overflow) {
			B.n_aln[rid] = -1;
			const int o = atomicAdd(B.overflow_count, 1);
			B.overflow_ids[o] = rid;
		} else {
			uint32_t h = n_aln > 0 ? (uint32_t)heads[HIT_HEAD] : NIL;
			for (int j = 0; j < n_aln; ++j) {
				uint4 q = *ent_at(h);
				if (EMB) { h = unlink16(q.w >> 16); q.w &= 0xffffu; } else h = *nxt_at(h);
				B.pool[off + j] = q;
			}
			B.n_aln[rid] = n_aln;
			B.pool_off[rid] = off;
			B.max_entries[rid] = max_entries;
		}
#if BWAGPU_TOP_REG
		lr_s = -1;
#endif
		mask.reset(); cur_s = -1; cur_head = NIL;
		pend_idx = NIL;
		bump = 0; free_head = NIL; spare = NIL; held = false; n_entries = 0; n_aln = 0;
		if (POOLED) while (n_chunks > 1) chunk_free(ctab[--n_chunks]); // keep one chunk, recycle the rest
	};

	// action for found hits (bwtgap.c:167-200).  Returns false when the search must stop.
	auto process_hit = [&](uint32_t hk, uint32_t hl) -> bool {
		const int mm = (int)(e_tag & 0xffu), go = (int)((e_tag >> 8) & 0xffu), ge = (int)((e_tag >> 16) & 0xffu);
		const int score = score_of(mm, go, ge);
		const int len = RD_LEN;
		bool do_add = true;
		if (n_aln == 0) {
			best_score = score;
			int best_diff = mm + go;
			if (gape_mode) best_diff += ge;
			if (!nonstop) max_diff = (best_diff + 1 > RD_MAXDIFF) ? RD_MAXDIFF : best_diff + 1; // top2 behaviour
		}
		if (score == best_score) best_cnt += (int)(hl - hk + 1);
		else if (best_cnt > O.max_top2) return false; // top2b behaviour
		if (go) { // the hit may have been found already (gap in a tandem repeat)
			uint32_t h = n_aln > 0 ? (uint32_t)heads[HIT_HEAD] : NIL;
			for (int j = 0; j < n_aln; ++j) {
				const uint4 q = *ent_at(h);
				if (q.y == hk && q.z == hl) { do_add = false; break; }
				h = EMB ? unlink16(q.w >> 16) : *nxt_at(h);
			}
		}
		if (do_add) {
			// gap_shadow (bwtgap.c:81-91) on the searched strand's width array, then refresh the packed
			// (bid, w[t-1]==w[t]) view and the context words of the positions it may have changed
			const uint32_t a = (e_tag >> 26) & 1u;
			const size_t wo = (size_t)w_off + (size_t)a * WSTRIDE(len);
			uint32_t *w = B.w + wo;
			uint16_t *wb = B.bid + wo;
			uint16_t *cx = B.ctx16 + wo;
			const uint32_t x = hl - hk + 1, mx = B.ix[1 - a].seq_len;
			const int ldp = (int)(e_pos >> 16);
			uint32_t j = 0, prev = 0;
			for (int t = 0; t <= ldp && t <= len; ++t) {
				uint32_t wv = w[t], bid = wb[t] & WB_BID;
				if (t < ldp) {
					if (wv > x) { wv -= x; w[t] = wv; }
					else if (wv == x) { bid = 1; wv = mx - (++j); w[t] = wv; }
				}
				const uint32_t nb = bid | ((t > 0 && wv == prev) ? WB_EQ : 0u);
				wb[t] = (uint16_t)nb;
				if (t < len) cx[t] = (uint16_t)((cx[t] & C16_KEEP) | c16_width(nb)); // the entry of position t: its width fields
				prev = wv;
			}
			cx_tag = 0xffffffffu; // the sector kept in shared memory may hold entries just rewritten
			const uint32_t idx = alloc_rec();
			if (idx != NIL) {
				if (EMB) {
					*ent_at(idx) = make_uint4((uint32_t)mm | (uint32_t)go << 8 | (uint32_t)ge << 16 | a << 24, hk, hl, (uint32_t)score | 0xffffu << 16);
					if (n_aln > 0) reinterpret_cast<uint16_t *>(&ent_at((uint32_t)heads[HIT_TAIL])->w)[1] = (uint16_t)idx; // the tail's link (upper half of its score word)
					else heads[HIT_HEAD] = (head_t)idx;
				} else {
					*ent_at(idx) = make_uint4((uint32_t)mm | (uint32_t)go << 8 | (uint32_t)ge << 16 | a << 24, hk, hl, (uint32_t)score);
					*nxt_at(idx) = NIL;
					if (n_aln > 0) *nxt_at((uint32_t)heads[HIT_TAIL]) = idx; else heads[HIT_HEAD] = (head_t)idx;
				}
				heads[HIT_TAIL] = (head_t)idx;
				++n_aln;
			}
		}
		return true;
	};

	// hit / exact tail / expansion for the node (k, l, e_pos, e_tag) at position i (bwtgap.c:160-165, 201).
	// Returns true when the trip goes on to an occurrence lookup.
	auto decide = [&]() -> bool {
		if (i == 0) {
			mode = MODE_POP;
			if (!process_hit(k, l)) { finish_read(); mode = MODE_NEW; }
			return false;
		}
		if (m == 0 && (((e_tag >> 24) & 3u) == STATE_M || gape_mode || (int)((e_tag >> 16) & 0xffu) == O.max_gape)) { // no diff allowed
			cc = CW_C1;
			if (cc > 3u) { mode = MODE_POP; return false; } // N in the tail: no match
			mode = MODE_EXACT;
		} else {
			--i;
			mode = MODE_EXPAND;
		}
		return true;
	};

	// The loop body is written without early `continue`s: every lane walks the same sequence
	// of guarded blocks (NEW / POP / loads / pruning / consume), so the warp re-converges
	// after each block and ALL lanes that need memory this trip issue their loads together.
	for (;;) {
		bool fresh = false, need_derive = false;
		bool active = true; // takes part in this trip's occurrence lookup
		if (STATS) ++n_trips;
		if (mode == MODE_NEW) {
			const int job = atomicAdd(B.work_counter, 1);
			if (job >= B.n_jobs) {
				if (STATS) atomicMin(B.stats + 11, gtime());
				break;
			}
			rid = B.jobs ? B.jobs[job] : job;
			const ReadMeta md = B.meta[rid];
			const int len = md.len;
			rd = (uint32_t)md.len | (uint32_t)md.max_gapo << 16 | (uint32_t)md.max_diff << 24;
			overflow = false;
			n_aln = 0; max_entries = 0; best_cnt = 0; n_entries = 0;
			pops_left = (!POOLED && B.pop_cap) ? B.pop_cap : 0xffffffffu;
			if (STATS) read_pops = 0;
			max_diff = md.max_diff;
			// len == 0: bwtaln.c:134 (aln = 0, n_aln = 0); too many N: bwtgap.c:118-123
			// (*pmax_entries is left untouched there)
			if (len == 0 || (int)md.n_amb > max_diff) {
				B.n_aln[rid] = 0; B.pool_off[rid] = 0; B.max_entries[rid] = 0;
				active = false; // stays in MODE_NEW
			} else if (EMB && len > 255) { // positions do not fit the packed record: the next pass takes the read
				overflow = true;
				mode = MODE_POP;
			} else {
				w_off = md.w_off;
				cx_tag = 0xffffffffu;
				best_score = score_of(max_diff + 1, (int)md.max_gapo + 1, O.max_gape + 1);
				// the two root nodes (bwtgap.c:127-128): strand 0 stored, strand 1 (popped first) held
				push_rec(0u, B.ix[0].seq_len, (uint32_t)len, 0u, 0, 1);
				k = 0u; l = B.ix[0].seq_len; i = len; e_tag = 1u << 26;
				held = true; ++n_entries;
				if (STATS) ++n_pushes;
				mode = MODE_POP;
			}
		}

		if (active && (mode == MODE_POP || mode == MODE_POPWAIT)) {
			const bool second = BWAGPU_SPLIT_POP && mode == MODE_POPWAIT; // the record asked for in the last trip has arrived
			bool stop = false;
			if (!second) {
				if (!POOLED && pops_left-- == 0) overflow = true; // a straggler: the warp-per-read pass does it faster than one lane can
				stop = overflow || n_entries == 0;
				if (!stop) {
					if (max_entries < n_entries) max_entries = n_entries;
					// > max_entries (bwtgap.c:140); only phantoms left: the reference pops one and stops
#if BWAGPU_TOP_REG
					stop = n_entries > O.max_entries || (!held && !mask.any() && lr_s < 0);
#else
					stop = n_entries > O.max_entries || (!held && !mask.any());
#endif
				}
			}
			if (stop) { finish_read(); mode = MODE_NEW; active = false; }
			else {
				bool took = true; // a node was popped in this trip
				// gap_pop (bwtgap.c:66-79)
				if (!second && held) { // the match child: counts of its parent, state M, no difference at its position
					held = false;
					e_pos = (uint32_t)i;
					e_tag &= ~(3u << 24);
				} else {
#if BWAGPU_TOP_REG
					const bool from_lr = lr_s >= 0 && (!mask.any() || lr_s <= mask.lowest());
#else
					const bool from_lr = false;
#endif
					uint4 q;
					uint4 *qp = nullptr;
					uint32_t idx = NIL, nx = NIL;
					int s = 0;
					if (from_lr) {
#if BWAGPU_TOP_REG
						q = make_uint4(lr_k, lr_l, lr_pos, lr_tag);
#endif
					} else {
						s = mask.lowest();
						if (s != cur_s) {
							if (cur_s >= 0 && mask.test(cur_s)) heads[cur_s * HS] = (head_t)cur_head;
							cur_s = s; cur_head = heads[s * HS];
						}
						idx = cur_head;
						qp = ent_at(idx);
#if BWAGPU_SPLIT_POP
						if (!second) { // first half: issue the loads and sit this trip's lookup out
							pq = *qp;
							if (!EMB) pnx = *nxt_at(idx);
							mode = MODE_POPWAIT;
							took = false; active = false;
						}
						q = pq; nx = pnx;
#else
						if (PC && idx == pend_idx) q = make_uint4(pcw[0], pcw[CS], pcw[2 * CS], pcw[3 * CS]); // the group popped from last
						else q = *qp;
						if (!EMB) nx = *nxt_at(idx); // issued with the entry load, used only when the record is unlinked
#endif
						if (EMB) { // z = i | (last_diff_pos or mask) << 8 | link << 16  ->  the plain form the code below works on
							nx = unlink16(q.z >> 16);
							q.z = (q.z & 0xffu) | ((q.z >> 8) & 0xffu) << 16;
						}
					}
					if (took) {
						if (STATS && !from_lr) ++n_mempop;
						mode = MODE_POP;
						const uint32_t kind = (q.w >> 27) & 3u;
						uint32_t gm = 0, b = 0;
						if (kind != KIND_PLAIN) {
							gm = (q.z >> 16) & 31u;
							b = 31u - (uint32_t)__clz((int)gm); // child pushed last = highest bit
							gm &= ~(1u << b);
						}
						if (from_lr) {
#if BWAGPU_TOP_REG
							if (gm) lr_pos = (q.z & 0xffffu) | gm << 16; // stays on top with one child fewer
							else lr_s = -1;
#endif
						} else {
							bool unlink = gm == 0;
#if BWAGPU_TOP_REG
							if (gm && lr_s < 0) { // the rest of the group moves to registers: it stays the top of the lowest bucket until it is used up
								lr_k = q.x; lr_l = q.y; lr_pos = (q.z & 0xffffu) | gm << 16; lr_tag = q.w; lr_s = s;
								unlink = true;
							}
#endif
							if (!unlink) { // record stays on top with one child fewer
								if (PC) {
									if (idx != pend_idx) {
										if (pend_idx != NIL) reinterpret_cast<uint8_t *>(&ent_at(pend_idx)->z)[1] = (uint8_t)(pcw[2 * CS] >> 8); // write the displaced group's mask back
										pcw[0] = q.x; pcw[CS] = q.y; pcw[3 * CS] = q.w;
										pend_idx = idx;
									}
									pcw[2 * CS] = (q.z & 0xffu) | gm << 8 | link16(nx) << 16;
								} else qp->z = EMB ? ((q.z & 0xffu) | gm << 8 | link16(nx) << 16) : ((q.z & 0xffffu) | gm << 16);
							} else {
								if (PC && idx == pend_idx) pend_idx = NIL;
								cur_head = nx;
								if (cur_head == NIL) mask.clear(s);
								if ((POOLED || !BWAGPU_NO_FREELIST) && spare != NIL) { *nxt_at(spare) = free_head; free_head = spare; }
								spare = idx;
							}
						}
						k = q.x; l = q.y;
						if (kind == KIND_PLAIN) { e_pos = q.z; e_tag = q.w; }
						else {
							const uint32_t pi = q.z & 0xffffu, pst = (q.w >> 24) & 3u, a = (q.w >> 26) & 1u;
							uint32_t mm = q.w & 0xffu, go = (q.w >> 8) & 0xffu, ge = (q.w >> 16) & 0xffu, ci, st;
							if (kind == KIND_MM) {
								const uint32_t cb = q.w >> 29; // the parent's read base at pi, kept in the record
								cc = (cb + b + 1u) & 3u;
								++mm; ci = pi; st = STATE_M; need_derive = true;
							} else {
								if (pst == STATE_M) ++go; else ++ge;
								if (b == 0) { ci = pi; st = STATE_I; }
								else { ci = pi + 1u; st = STATE_D; cc = b - 1u; need_derive = true; }
							}
							e_pos = ci | ci << 16; // every group child is a difference: last_diff_pos = its own i
							e_tag = mm | go << 8 | ge << 16 | st << 24 | a << 26;
						}
						i = (int)(e_pos & 0xffffu);
					}
				}
				if (took) {
					--n_entries;
					if (STATS) { ++n_pops; ++read_pops; }
					fresh = true;
				}
			}
		}

		// ---- issue every load this trip depends on before using any of them: the two
		// occurrence blocks (bwt_2occ4 at bwtgap.c:202 / bwt_2occ at bwt.c:245) and the context word --
		// the node's own for a node just popped, the next read base for an exact tail.  One memory round trip per trip.
		const uint32_t a = (e_tag >> 26) & 1u;
		const DevIndex &ix = B.ix[1 - a];
		uint32_t jk = 0, jl = 0;
		OccBlock ob_l = {0, 0, 0, 0, 0, 0}, ob_k = {0, 0, 0, 0, 0, 0};
		if (active) {
			jk = occ_arg(ix, k - 1); jl = occ_arg(ix, l);
			ob_l = load_block(ix, jl >> 6);
			ob_k = load_block(ix, jk >> 6); // same sector as ob_l for narrow intervals (L1 hit)
			if (fresh | (mode == MODE_EXACT)) { // fresh: the node's own entries E[i-1], E[i-2]; exact tail: E[i-2] (its base = str[i-2]) in the low half
				const int j = fresh ? i : i - 1;
#if BWAGPU_CTX_SMEM
				uint32_t e1 = 0u, e2 = 0u;
				if (j >= 1) {
					const uint32_t g1 = w_off + a * (uint32_t)WSTRIDE(RD_LEN) + (uint32_t)j - 1u; // entry index of position j - 1
					if ((g1 >> 4) != cx_tag) { // fetch its sector (the arena is 32-byte aligned) and keep it
						cx_tag = g1 >> 4;
						const uint4 *sp = reinterpret_cast<const uint4 *>(ctx16 + ((size_t)cx_tag << 4));
						const uint4 s0 = sp[0], s1 = sp[1];
						cxw[0] = s0.x; cxw[CS] = s0.y; cxw[2 * CS] = s0.z; cxw[3 * CS] = s0.w;
						cxw[4 * CS] = s1.x; cxw[5 * CS] = s1.y; cxw[6 * CS] = s1.z; cxw[7 * CS] = s1.w;
					}
					e1 = (cxw[((g1 & 15u) >> 1) * CS] >> ((g1 & 1u) << 4)) & 0xffffu;
					if (fresh && j >= 2) {
						const uint32_t g2 = g1 - 1u;
						e2 = (g2 >> 4) == cx_tag ? (cxw[((g2 & 15u) >> 1) * CS] >> ((g2 & 1u) << 4)) & 0xffffu : (uint32_t)ctx16[g2];
					}
				}
#else
				const uint16_t *cp = ctx16 + ((size_t)w_off + (size_t)a * WSTRIDE(RD_LEN) + (size_t)j);
#if BWAGPU_CTX_KEEP && !defined(BWAGPU_HOST_EMU)
				uint32_t e1 = 0u, e2 = 0u;
				{
					unsigned long long pol;
					unsigned short v;
					asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
					if (j >= 1) { asm volatile("ld.global.L1::evict_last.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(cp - 1), "l"(pol)); e1 = v; }
					if (fresh && j >= 2) { asm volatile("ld.global.L1::evict_last.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(cp - 2), "l"(pol)); e2 = v; }
				}
#else
				const uint32_t e1 = j >= 1 ? (uint32_t)cp[-1] : 0u;
				const uint32_t e2 = (fresh && j >= 2) ? (uint32_t)cp[-2] : 0u;
#endif
#endif
				cw = e1 | e2 << 16;
			}
			if (STATS) {
				{ const int dpt = RD_LEN - i; atomicAdd(B.stats + 32 + (dpt < 0 ? 0 : dpt > 31 ? 31 : dpt), 1ull); }
				f_own += (jk >> 6) != (jl >> 6) ? 2u : 1u;
				if (k == 0) f_ref += 1u;
				else {
					const uint32_t km1 = k - 1;
					const uint32_t pk = km1 >= ix.primary ? km1 - 1 : km1, pl = l >= ix.primary ? l - 1 : l;
					f_ref += (pk >> 7) == (pl >> 7) ? 1u : 2u;
				}
			}
		}

		if (active && fresh) { // pruning tests of bwtgap.c:144-157
			const int mm = (int)(e_tag & 0xffu), go = (int)((e_tag >> 8) & 0xffu), ge = (int)((e_tag >> 16) & 0xffu);
			if (!nonstop && score_of(mm, go, ge) > best_score + O.s_mm) { finish_read(); mode = MODE_NEW; active = false; }
			else {
				m = max_diff - (mm + go);
				if (gape_mode) m -= ge;
				if (m < 0 || (i > 0 && m < CW_BID1)) { if (STATS) ++n_pruned; active = false; } // stays in MODE_POP
				else if (need_derive) mode = MODE_DERIVE;
				else active = decide();
			}
		}

		if (active) {
			uint32_t nk[4], nl[4]; // the four one-symbol extensions of (k, l): [nk[c], nl[c]]
			{
				uint32_t cnt_k[4], cnt_l[4];
				occ4_in_block(ob_k, jk, cnt_k);
				occ4_in_block(ob_l, jl, cnt_l);
				nk[0] = cnt_k[0] + 1; nk[1] = C1 + cnt_k[1] + 1; nk[2] = C2 + cnt_k[2] + 1; nk[3] = C3 + cnt_k[3] + 1;
				nl[0] = cnt_l[0]; nl[1] = C1 + cnt_l[1]; nl[2] = C2 + cnt_l[2]; nl[3] = C3 + cnt_l[3];
			}

			if (mode == MODE_DERIVE) { // k,l were the parent's: take child cc's interval
				if (STATS) ++n_derive;
				k = sel4(cc, nk);
				l = sel4(cc, nl);
				decide(); // next trip looks the child's own interval up
			} else if (mode == MODE_EXACT) { // bwt_match_exact_alt (bwt.c:237-252), one base per trip
				if (STATS) ++n_exact;
				k = sel4(cc, nk);
				l = sel4(cc, nl);
				--i;
				if (k > l) mode = MODE_POP;
				else if (i == 0) {
					mode = MODE_POP;
					if (!process_hit(k, l)) { finish_read(); mode = MODE_NEW; }
				} else {
					// the base the next step matches, str[i-1]: a tail's first step still holds the node's own word
					// (str[i_old-2] is its second base field), later steps loaded word i_old-1
					cc = fresh ? CW_CN : CW_C1;
					if (cc > 3u) mode = MODE_POP;
				}
			} else { // ---- MODE_EXPAND (bwtgap.c:201-259); i was already decremented
				if (STATS) ++n_expand;
				const int mm = (int)(e_tag & 0xffu), go = (int)((e_tag >> 8) & 0xffu), ge = (int)((e_tag >> 16) & 0xffu);
				const uint32_t st = (e_tag >> 24) & 3u;
				const uint32_t occ = l - k + 1;
				const int len = RD_LEN, max_gapo = RD_GAPO;
				const bool has_seed = len > O.seed_len;
				bool allow_diff = true, allow_M = true;
				if (i > 0) {
					const int b1 = CW_B2; // width[i-1].bid (i was decremented: the node's own i-2)
					if (b1 > m - 1) allow_diff = false;
					else if (b1 == m - 1 && CW_BID1 == m - 1 && CW_EQ1) allow_M = false;
					if (has_seed) {
						const int si = i - (len - O.seed_len);
						if (si > 0) {
							const int m_seed = m - max_diff + O.max_seed_diff; // max_seed_diff - (mm + go [+ ge]) (bwtgap.c:153-155)
							const int s1 = CW_S2; // seed_width[si-1].bid
							if (s1 > m_seed - 1) allow_diff = false;
							else if (s1 == m_seed - 1 && CW_SBID1 == m_seed - 1 && CW_SEQ1) allow_M = false;
						}
					}
				}
				const uint32_t ci = CW_C1;
				// which of the four one-symbol extensions are non-empty (k' <= l')
				const uint32_t V = (nk[0] <= nl[0] ? 1u : 0u) | (nk[1] <= nl[1] ? 2u : 0u) | (nk[2] <= nl[2] ? 4u : 0u) | (nk[3] <= nl[3] ? 8u : 0u);
#if BWAGPU_PREFETCH_HELD && !defined(BWAGPU_HOST_EMU)
				// The match child is the next pop of this lane and its interval is known here, a few hundred issue slots (the
				// pushes below, at 3-5 lanes each) before the next trip asks for its occurrence blocks: start them now.
				if (ci < 4 && ((V >> ci) & 1u)) {
					const uint32_t pjk = occ_arg(ix, sel4(ci, nk) - 1u) >> 6, pjl = occ_arg(ix, sel4(ci, nl)) >> 6;
					asm volatile("prefetch.global.L1 [%0];" :: "l"(ix.blk + 2 * (size_t)pjl));
					if (pjk != pjl) asm volatile("prefetch.global.L1 [%0];" :: "l"(ix.blk + 2 * (size_t)pjk));
				}
#endif
				const int score = score_of(mm, go, ge);
				const uint32_t ptag = e_tag; // plain node: no kind / base bits
				if (allow_diff) {
					// indels (bwtgap.c:218-247): one KIND_GAP record
					int tmp;
					if (loggap) {
						const uint32_t v = (uint32_t)(ge + go);
						tmp = (v ? 31 - __clz((int)v) : 0) / 2 + 1; // int_log2 (bwtgap.c:93-102)
					} else tmp = go + ge;
					if (i >= O.indel_end_skip + tmp && len - i >= O.indel_end_skip + tmp) {
						uint32_t gm = 0;
						int gs = score + O.s_gape;
						if (st == STATE_M) { if (go < max_gapo) { gm = 1u | V << 1; gs = score + O.s_gapo; } }
						else if (st == STATE_I) { if (ge < O.max_gape) gm = 1u; }
						else if (ge < O.max_gape && (ge + go < max_diff || occ < (uint32_t)O.max_del_occ)) gm = V << 1;
						if (gm) {
							uint32_t rk = k, rl = l, rpos = (uint32_t)i | gm << 16, rtag = ptag | KIND_GAP << 27;
							if (!(gm & (gm - 1))) { // a single child is stored as the plain node it is (no derive trip later)
								const uint32_t ngo = st == STATE_M ? go + 1 : go, nge = st == STATE_M ? ge : ge + 1;
								const uint32_t tg = (uint32_t)mm | ngo << 8 | nge << 16 | a << 26;
								if (gm == 1u) { rpos = (uint32_t)i | (uint32_t)i << 16; rtag = tg | STATE_I << 24; }
								else {
									const uint32_t c = 30u - (uint32_t)__clz((int)gm); // bit 1 + c
									rk = sel4(c, nk); rl = sel4(c, nl);
									rpos = (uint32_t)(i + 1) | (uint32_t)(i + 1) << 16; rtag = tg | STATE_D << 24;
								}
							}
							push_rec(rk, rl, rpos, rtag, gs, __popc(gm));
						}
					}
					if (allow_M) { // mismatches (bwtgap.c:248-257): one KIND_MM record; the match is held
						uint32_t mmask = 0;
#pragma unroll
						for (uint32_t j = 1; j <= 3; ++j) mmask |= ((V >> ((ci + j) & 3u)) & 1u) << (j - 1);
						if (ci > 3) mmask |= (V & 1u) << 3; // N: j = 4 is a mismatch too, c = (4 + 4) & 3 = 0
						if (mmask) {
							uint32_t rk = k, rl = l, rpos = (uint32_t)i | mmask << 16, rtag = ptag | KIND_MM << 27 | ci << 29;
							if (!(mmask & (mmask - 1))) {
								const uint32_t c = (ci + (32u - (uint32_t)__clz((int)mmask))) & 3u; // bit j-1 -> c = (ci + j) & 3
								rk = sel4(c, nk); rl = sel4(c, nl);
								rpos = (uint32_t)i | (uint32_t)i << 16;
								rtag = (uint32_t)(mm + 1) | (uint32_t)go << 8 | (uint32_t)ge << 16 | STATE_M << 24 | a << 26;
							}
							push_rec(rk, rl, rpos, rtag, score + O.s_mm, __popc(mmask));
						}
					}
				}
				if (ci < 4 && ((V >> ci) & 1u)) { // the match: last push, next pop -> stays in registers
					k = sel4(ci, nk);
					l = sel4(ci, nl);
					held = true; ++n_entries;
					if (STATS) ++n_pushes;
				}
				mode = MODE_POP;
			}
		}
	}
#undef RD_LEN
#undef RD_GAPO
#undef RD_MAXDIFF
#undef CW_BID1
#undef CW_EQ1
#undef CW_C1
#undef CW_B2
#undef CW_CN
#undef CW_SBID1
#undef CW_SEQ1
#undef CW_S2

	if (STATS) {
		atomicAdd(B.stats + 0, (unsigned long long)f_ref);
		atomicAdd(B.stats + 1, (unsigned long long)f_own);
		atomicAdd(B.stats + 2, (unsigned long long)n_pops);
		atomicAdd(B.stats + 3, (unsigned long long)n_pushes);
		atomicAdd(B.stats + 4, (unsigned long long)n_stored);
		atomicAdd(B.stats + 5, (unsigned long long)n_pruned);
		atomicAdd(B.stats + 6, (unsigned long long)n_expand);
		atomicAdd(B.stats + 7, (unsigned long long)n_exact);
		atomicAdd(B.stats + 8, (unsigned long long)n_derive);
		atomicAdd(B.stats + 9, (unsigned long long)n_trips);
		atomicAdd(B.stats + 10, (unsigned long long)n_mempop);
		atomicMax(B.stats + 13, gtime());
	}
}

// ------------------------------------------------------------------ measurement: random-sector gather ceiling
// What a kernel made of nothing but dependent random 32-byte sector loads sustains on this device (SURVEY.md §8d:
// "ceiling = measured random 32 B-sector read throughput").  Each thread walks CHAINS independent chains; the next
// index of a chain depends on the sector just loaded, as an FM-index step does.
template <int CHAINS>
__global__ void __launch_bounds__(256) k_probe_gather(const uint4 *__restrict__ buf, uint32_t n_blk, int steps, uint32_t seed,
                                                      uint32_t *__restrict__ sink)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t x[CHAINS], acc = 0;
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) x[c] = (t * CHAINS + c + seed) * 2654435761u + 12345u;
	DevIndex ix = {};
	ix.blk = buf;
	for (int s = 0; s < steps; ++s) {
		OccBlock o[CHAINS];
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) o[c] = load_block(ix, (uint32_t)(((uint64_t)x[c] * n_blk) >> 32));
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) { acc += o[c].c3; x[c] = (x[c] ^ o[c].c0) * 1664525u + 1013904223u; }
	}
	sink[t] = acc;
}

// ------------------------------------------------------------------ job ordering for K3
// Sort key per read from the widths K2 just computed: D = min over strands of the lower
// bound on differences for the whole read (width[len-1].bid).  Reads with 1 <= D <= max_diff
// explore the most (the larger D the longer), D = 0 (an exact match exists) and D > max_diff
// (pruned at the root, bwtgap.c:157) the least.  K3 hands jobs out in key order: the longest
// searches start first (shorter tail) and the lanes of a warp work on similar reads (less
// divergence).  Order of processing is not observable in the results.
__global__ void k_job_keys(const Batch B, uint8_t *__restrict__ keys, int32_t *__restrict__ ids)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= B.n_reads) return;
	const ReadMeta m = B.meta[r];
	uint32_t key = 255;
	if (m.len > 0) {
		const uint16_t *wb = B.bid + m.w_off;
		const uint32_t d0 = wb[m.len - 1] & WB_BID, d1 = wb[WSTRIDE(m.len) + m.len - 1] & WB_BID;
		const uint32_t d = d0 < d1 ? d0 : d1, md = m.max_diff;
		key = (d >= 1 && d <= md) ? md - d : md + 1 + (d == 0 ? 0u : 1u);
	}
	keys[r] = (uint8_t)key;
	ids[r] = r;
}

// ------------------------------------------------------------------ ordered compaction of the aln pool
// gather: pool (completion order) -> out (read order, offsets = exclusive scan of n_aln)
__global__ void k_gather_aln(int n_reads, const int32_t *__restrict__ n_aln, const uint32_t *__restrict__ pool_off,
                             const uint32_t *__restrict__ out_off, const uint4 *__restrict__ pool, uint4 *__restrict__ out)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int n = n_aln[r];
	const uint32_t src = pool_off[r], dst = out_off[r];
	for (int j = 0; j < n; ++j) out[dst + j] = pool[src + j];
}

// ------------------------------------------------------------------ K4: SA -> coordinate
// bwt_sa (bwt.c:72-81): walk LF to the next sampled row; one 32-byte block per step.
struct IndexPair { DevIndex ix[2]; }; // [0] forward, [1] reverse

__global__ void __launch_bounds__(256) k_sa(const IndexPair P, long long n,
                                            const uint32_t *__restrict__ q, const uint8_t *__restrict__ which,
                                            uint32_t *__restrict__ out)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	const bool fw = which[t] != 0;
	DevIndex ix; // field-wise select keeps the struct in registers (no local copy of the param)
	ix.blk = fw ? P.ix[0].blk : P.ix[1].blk;
	ix.sa = fw ? P.ix[0].sa : P.ix[1].sa;
	ix.primary = fw ? P.ix[0].primary : P.ix[1].primary;
	ix.seq_len = fw ? P.ix[0].seq_len : P.ix[1].seq_len;
	ix.sa_intv = fw ? P.ix[0].sa_intv : P.ix[1].sa_intv;
#pragma unroll
	for (int j = 0; j < 5; ++j) ix.L2[j] = fw ? P.ix[0].L2[j] : P.ix[1].L2[j];
	uint32_t k = q[t], steps = 0;
	const uint32_t intv = ix.sa_intv;
	while (k % intv != 0) {
		++steps;
		k = inv_psi(ix, k);
	}
	k /= intv;
	uint32_t sv = 0xffffffffu;
	if (k) {
#if BWAGPU_IDX_L2_64B && !defined(BWAGPU_HOST_EMU)
		asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(sv) : "l"(ix.sa + k)); // one random 4-byte read: no 128-byte fill
#else
		sv = ix.sa[k];
#endif
	}
	out[t] = steps + sv;
}

} // namespace bwagpu
