// search_warp.cuh -- K3w: bwt_match_gap (bwtgap.c:104-266) with ONE WARP per read, for the searches that
// outgrow pass 0 of k_search (kernels.cuh).
//
// Why: k_search gives a read one thread.  That is the right shape while hundreds of thousands of reads are in
// flight, but a deep search (ancient-DNA options: no seed, -o 2; max_entries up to 10^5..10^6 per read) is then
// one lane walking a dependent chain of memory round trips, ~2 us per popped entry, and the launch lasts as long
// as its deepest read.  Here the 32 lanes of a warp drain the SAME read's lowest score bucket together.
//
// What makes that exact.  The reference pops the lowest non-empty bucket LIFO.  An entry popped from bucket s
// pushes its children into buckets s + s_mm / s + s_gapo / s + s_gape -- all above s -- except the match child,
// which has score s, is pushed last and is therefore the very next pop.  So the pops the reference makes before
// it touches the entry BELOW the current top of bucket s are exactly the current top and its chain of match
// descendants.  The top T entries of bucket s therefore start T independent CHAINS, and the reference's order is
// "chain 0 completely, then chain 1, ...".  Lane j runs chain j.  What must be reconstructed is the order of
// the pushes in the three target buckets (chain 0's first, in its own order, then chain 1's, ...) and the value
// of stack->n_entries at every pop (max_entries, bwtgap.c:139-140).  Both are prefix sums over the lanes:
//
//   pass A  every lane runs its chain WITHOUT storing: it counts its pushes per target bucket, the net change of
//           n_entries and the highest n_entries it sees relative to its start;
//   scan    exclusive prefix sums give each lane the position of its first push in each target bucket and the
//           absolute n_entries it starts from;
//   pass B  the lanes run their chains again (occurrence blocks and context words are now L1 hits) and store
//           each pushed entry at its final position: buckets are dense arrays (chunks of 1024 entries from the
//           launch's shared pool, linked downwards), so the next take of a bucket's top 32 entries is coalesced.
//
// Only a HIT changes what later pops see (max_diff, best_score, gap_shadow's edits of w[], bwtgap.c:167-200), and
// a chain ends at its hit.  If lane h is the first lane whose chain hits, lanes 0..h are committed, lane h then
// processes its hit alone, and lanes h+1.. are simply not committed -- they wrote nothing in pass A, and their
// entries are still on the bucket; they are taken again next round under the new state.  The same rule handles
// the `n_entries > opt->max_entries` stop (the first lane that would see it is found from the prefix sums and
// stops at the exact pop in pass B) and the round is shrunk to one lane when a penalty is zero (a push could
// then land in the bucket being drained).  Everything else -- pruning by the width bounds, indel_end_skip,
// max_del_occ, phantoms (counted, not stored), first-in-order duplicate test -- is the reference's, entry by entry.
//
// TEAM (template parameter): the warps that share a read -- 1 (several reads per block) or all warps of the block (one read per
// block, rounds of up to 32 * TEAM chains; scans, first-event search and broadcasts then go through a few words of shared memory
// and __syncthreads).  Same lanes per SM; a deep read finishes sooner, the bulk rate is a little lower.  The host picks by the
// size of the pass (bwagpu.cu).
//
// Memory: all of a read's entries and hits live in chunks of the launch's shared pool (Batch::xent; the word
// xnxt[chunk << 10] links a chunk to the one below it).  A read that finds the pool dry is flagged and retried by
// the guaranteed pass of k_search, like before.
#pragma once
#include "kernels.cuh"

namespace bwagpu {

#ifndef WK_WARPS_PER_BLOCK
#define WK_WARPS_PER_BLOCK 4 // 4 warps x 5 blocks per SM measured 6 % faster than 8 x 2 (profiles/r1_ab_experiments.md)
#endif
#ifndef BWAGPU_WARP_TEAM_DEFAULT
#define BWAGPU_WARP_TEAM_DEFAULT 0 // 1: the warps of a block share one read by default (BWAGPU_WARP_TEAM overrides at run time)
#endif
#ifndef WK_MINBLOCKS
#define WK_MINBLOCKS 5 // __launch_bounds__ second argument: 96 registers
#endif
#define WK_WARPS WK_WARPS_PER_BLOCK // warps (= reads in flight) per block
#define WK_NB 257     // 256 score buckets (the ABI's limit) + the read's hit list
#define WK_HITS 256
#define WK_SEG 16     // chunks one round may touch per target bucket
#define WK_WORDS_PER_WARP (2 * WK_NB + 3 * (WK_SEG + 1) + 64 + 1 + 16) // + the chunk cache (WK_CACHE ids + fill) + team scratch

// chunks of the shared pool: bump counter first, then a lock-free stack of recycled chunks ({tag:32 | head:32} against ABA)
__device__ __forceinline__ uint32_t pool_chunk_alloc(const Batch &B)
{
	unsigned long long old = *(volatile unsigned long long *)B.x_free_top;
	while ((uint32_t)old != NIL) {
		const uint32_t head = (uint32_t)old;
		const uint32_t nx = ((volatile uint32_t *)B.x_free_next)[head];
		const unsigned long long nw = (((old >> 32) + 1ull) << 32) | nx;
		const unsigned long long prev = atomicCAS(B.x_free_top, old, nw);
		if (prev == old) return head;
		old = prev;
	}
	const uint32_t c = atomicAdd(B.x_next, 1u);
	return c < B.x_chunks ? c : NIL;
}

__device__ __forceinline__ void pool_chunk_free(const Batch &B, uint32_t c)
{
	unsigned long long old = *(volatile unsigned long long *)B.x_free_top;
	for (;;) {
		((volatile uint32_t *)B.x_free_next)[c] = (uint32_t)old;
		__threadfence();
		const unsigned long long nw = (((old >> 32) + 1ull) << 32) | c;
		const unsigned long long prev = atomicCAS(B.x_free_top, old, nw);
		if (prev == old) return;
		old = prev;
	}
}

// Recycled chunks of the warp pass: a lock-free stack of BATCHES.  A batch is a head chunk plus up to WK_BATCH_MAX
// more chunk ids written into the head chunk's own link words (xnxt[head << 10 | 1] = next batch, [.. | 2] = count,
// [.. | 3 ..] = ids), so taking or returning 16 chunks is ONE compare-and-swap on the pool's head word
// ({tag:32 | head:32} against ABA).  k_search's one-chunk-at-a-time list (kernels.cuh) is a different launch.
#define WK_BATCH_MAX 15
__device__ __forceinline__ void pool_batch_push(const Batch &B, uint32_t head)
{
	volatile uint32_t *const xw = (volatile uint32_t *)B.xnxt + ((size_t)head << ARENA_CHUNK_LOG);
	unsigned long long old = *(volatile unsigned long long *)B.x_free_top;
	for (;;) {
		xw[1] = (uint32_t)old;
		__threadfence();
		const unsigned long long nw = (((old >> 32) + 1ull) << 32) | head;
		const unsigned long long prev = atomicCAS(B.x_free_top, old, nw);
		if (prev == old) return;
		old = prev;
	}
}

// Per-warp chunk cache (shared memory; one lane at a time uses it).  2368 warps allocating and freeing a chunk every few
// rounds through the pool's single head word serialise on it (measured: 300 us per round); with the cache the pool sees one
// bump-counter add per WK_BUMP allocations, or one compare-and-swap per batch.
#define WK_CACHE 64
#define WK_BUMP 8
__device__ __forceinline__ uint32_t wk_alloc(const Batch &B, uint32_t *cache, uint32_t *cache_n)
{
	const uint32_t n = *cache_n;
	if (n) { *cache_n = n - 1; return cache[n - 1]; }
	if (*(volatile unsigned int *)B.x_next < B.x_chunks) { // chunks never handed out yet
		const uint32_t c0 = atomicAdd(B.x_next, (unsigned int)WK_BUMP);
		if (c0 < B.x_chunks) {
			const uint32_t m = min((uint32_t)WK_BUMP, B.x_chunks - c0);
			for (uint32_t q = 1; q < m; ++q) cache[q - 1] = c0 + q;
			*cache_n = m - 1;
			return c0;
		}
	}
	unsigned long long old = *(volatile unsigned long long *)B.x_free_top; // a batch of recycled chunks
	while ((uint32_t)old != NIL) {
		const uint32_t head = (uint32_t)old;
		volatile uint32_t *const xw = (volatile uint32_t *)B.xnxt + ((size_t)head << ARENA_CHUNK_LOG);
		const uint32_t nx = xw[1];
		const unsigned long long nw = (((old >> 32) + 1ull) << 32) | nx;
		const unsigned long long prev = atomicCAS(B.x_free_top, old, nw);
		if (prev == old) {
			const uint32_t m = xw[2];
			for (uint32_t q = 0; q < m; ++q) cache[q] = xw[3 + q];
			*cache_n = m;
			return head;
		}
		old = prev;
	}
	return NIL;
}

__device__ __forceinline__ void wk_free(const Batch &B, uint32_t *cache, uint32_t *cache_n, uint32_t c)
{
	uint32_t n = *cache_n;
	if (n == WK_CACHE) { // full: hand the upper half back as one batch
		const uint32_t head = cache[WK_CACHE - 1];
		volatile uint32_t *const xw = (volatile uint32_t *)B.xnxt + ((size_t)head << ARENA_CHUNK_LOG);
		xw[2] = WK_BATCH_MAX;
		for (uint32_t q = 0; q < WK_BATCH_MAX; ++q) xw[3 + q] = cache[WK_CACHE - 2 - q];
		pool_batch_push(B, head);
		n = WK_CACHE - 1 - WK_BATCH_MAX;
	}
	cache[n] = c;
	*cache_n = n + 1;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, v, d);
		if (lane >= d) v += t;
	}
	return v;
}

__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
	return v;
}


// ---- a read's TEAM: one warp (TEAM = 1, several reads per block) or all TEAM warps of a block (one read per block, rounds of up to
// 32 * TEAM chains: the same lanes per SM, but a deep read finishes TEAM times sooner).  `lane` is the index within the team; tsm is
// a few words of shared memory per team for the cross-warp steps.
template <int TEAM> __device__ __forceinline__ void team_sync()
{
	if (TEAM == 1) __syncwarp(); else __syncthreads();
}

template <int TEAM> __device__ __forceinline__ int team_incl_scan(int v, int lane, int *tsm, int &total)
{
	const int w = warp_incl_scan(v, lane & 31);
	if (TEAM == 1) { total = __shfl_sync(0xffffffffu, w, 31); return w; }
	if ((lane & 31) == 31) tsm[lane >> 5] = w;
	__syncthreads();
	int pre = 0, tot = 0;
#pragma unroll
	for (int q = 0; q < TEAM; ++q) { const int t = tsm[q]; if (q < (lane >> 5)) pre += t; tot += t; }
	__syncthreads();
	total = tot;
	return w + pre;
}

template <int TEAM> __device__ __forceinline__ int team_max(int v, int lane, int *tsm)
{
	v = warp_max(v);
	if (TEAM == 1) return v;
	if ((lane & 31) == 0) tsm[lane >> 5] = v;
	__syncthreads();
	int r = tsm[0];
#pragma unroll
	for (int q = 1; q < TEAM; ++q) r = max(r, tsm[q]);
	__syncthreads();
	return r;
}

template <int TEAM> __device__ __forceinline__ int team_bcast(int v, int src, int lane, int *tsm)
{
	if (TEAM == 1) return __shfl_sync(0xffffffffu, v, src);
	if (lane == src) tsm[0] = v;
	__syncthreads();
	const int r = tsm[0];
	__syncthreads();
	return r;
}

template <bool STDMODE, bool WSTATS, int TEAM>
__global__ void __launch_bounds__(WK_WARPS * 32, WK_MINBLOCKS) k_search_warp(const Batch B)
{
#ifndef BWAGPU_WARP_EMU // tests/host_emu/warp_emu.cpp supplies the array
	extern __shared__ __align__(16) uint32_t wk_smem[];
#endif
	constexpr int TW = 32 * TEAM; // lanes of a team
	const int lane = (int)threadIdx.x % TW, wib = (int)threadIdx.x / TW; // lane within the team, team within the block
	uint32_t *const cnt = wk_smem + (size_t)wib * WK_WORDS_PER_WARP; // entries per bucket
	uint32_t *const top = cnt + WK_NB;                                // chunk holding a bucket's top entry (NIL: empty)
	uint32_t *const seg = top + WK_NB; // per target slot: [0] = number of the first chunk this round writes, [1 + r] = chunk ids
	uint32_t *const cache = seg + 3 * (WK_SEG + 1), *const cache_n = cache + 64; // chunk cache, kept from read to read
	int *const tsm = reinterpret_cast<int *>(cache_n + 1);
	if (lane == 0) *cache_n = 0;
	team_sync<TEAM>();
	const GapOpt &O = B.opt;
	const bool gape_mode = STDMODE || (O.mode & 0x01), loggap = !STDMODE && (O.mode & 0x04), nonstop = !STDMODE && (O.mode & 0x10);
	const uint32_t C1 = B.ix[0].L2[1], C2 = B.ix[0].L2[2], C3 = B.ix[0].L2[3];
	const int limit = O.max_entries;
	// target buckets of one expansion: s + s_mm (slot 0), s + s_gapo, s + s_gape; equal penalties share a slot so that
	// their pushes keep the order of the expansion
	const int d_mm = O.s_mm, d_go = O.s_gapo, d_ge = O.s_gape;
	const int slot_go = d_go == d_mm ? 0 : 1;
	const int slot_ge = d_ge == d_mm ? 0 : (d_ge == d_go ? slot_go : 2);
	const bool serial_only = d_mm <= 0 || d_go <= 0 || d_ge <= 0; // a push could land in the bucket being drained
	uint4 *const xent = B.xent;
	volatile uint32_t *const xlink = (volatile uint32_t *)B.xnxt; // xlink[chunk << 10] = the chunk below

	auto score_of = [&](int mm, int go, int ge) { return mm * O.s_mm + go * O.s_gapo + ge * O.s_gape; };

	for (;;) { // reads
		int job = 0;
		if (lane == 0) job = atomicAdd(B.work_counter, 1);
		job = team_bcast<TEAM>(job, 0, lane, tsm);
		if (job >= B.n_jobs) break;
		const int rid = B.jobs ? B.jobs[job] : job;
		const ReadMeta md = B.meta[rid];
		const int len = md.len, rd_maxdiff = md.max_diff, max_gapo = md.max_gapo;
		int max_diff = rd_maxdiff;
		// len == 0: bwtaln.c:134; too many N: bwtgap.c:118-123
		if (len == 0 || (int)md.n_amb > max_diff) {
			if (lane == 0) { B.n_aln[rid] = 0; B.pool_off[rid] = 0; B.max_entries[rid] = 0; }
			continue;
		}
		const uint32_t w_off = md.w_off;
		const bool has_seed = len > O.seed_len;
		const uint8_t *const seqp = B.seq + md.seq_off;
		int best_score = score_of(max_diff + 1, max_gapo + 1, O.max_gape + 1), best_cnt = 0, n_aln = 0, n_entries = 0, max_entries = 0;
		bool overflow = false;
		BucketMask<false> mask;
		mask.reset();
		for (int b = lane; b < WK_NB; b += TW) { cnt[b] = 0; top[b] = NIL; }
		team_sync<TEAM>();
		{ // the two root entries (bwtgap.c:127-128): strand 0 below strand 1
			uint32_t c = NIL;
			if (lane == 0) {
				c = wk_alloc(B, cache, cache_n);
				if (c != NIL) {
					xlink[(size_t)c << ARENA_CHUNK_LOG] = NIL;
					xent[((size_t)c << ARENA_CHUNK_LOG) + 0] = make_uint4(0u, B.ix[0].seq_len, (uint32_t)len, 0u);
					xent[((size_t)c << ARENA_CHUNK_LOG) + 1] = make_uint4(0u, B.ix[0].seq_len, (uint32_t)len, 1u << 26);
					cnt[0] = 2; top[0] = c;
				}
			}
			c = (uint32_t)team_bcast<TEAM>((int)c, 0, lane, tsm);
			if (c == NIL) overflow = true;
			else { mask.set(0); n_entries = 2; }
			team_sync<TEAM>();
		}

		int tcap = TW; // lanes per round: halved after a round that had to drop lanes, doubled otherwise
		// diagnostics (BWAGPU_WARP_STATS=1 prints them): rounds, lanes taken / committed, chain steps, clocks per phase
		unsigned long long st_rounds = 0, st_taken = 0, st_com = 0, st_steps = 0, st_maxsteps = 0, st_ck[6] = {0, 0, 0, 0, 0, 0};
		int steps = 0;
		while (!overflow) { // rounds
			long long ck0 = clock64();
			if (n_entries == 0) break;
			if (max_entries < n_entries) max_entries = n_entries;
			if (n_entries > limit) break;  // bwtgap.c:140
			if (!mask.any()) break;        // only phantoms left: the reference pops one and stops (bwtgap.c:144)
			const int s = mask.lowest();
			if (!nonstop && s > best_score + O.s_mm) break; // bwtgap.c:144
			const uint32_t cs = cnt[s];
			const int T = serial_only ? 1 : (int)min((uint32_t)tcap, cs);
			const bool mine = lane < T;

			// ---- take: lane j gets the j-th entry from the top of bucket s
			uint4 e = make_uint4(0u, 0u, 0u, 0u);
			{
				const uint32_t tq = (cs - 1) >> ARENA_CHUNK_LOG;
				const uint32_t ctop = top[s];
				uint32_t cbelow = NIL;
				if (((cs - (uint32_t)T) >> ARENA_CHUNK_LOG) != tq) cbelow = xlink[(size_t)ctop << ARENA_CHUNK_LOG];
				if (mine) {
					const uint32_t p = cs - 1 - (uint32_t)lane;
					const uint32_t c = (p >> ARENA_CHUNK_LOG) == tq ? ctop : cbelow;
					e = __ldcg(xent + ((size_t)c << ARENA_CHUNK_LOG) + (p & (ARENA_CHUNK - 1)));
				}
			}

			// ---- one chain: the entry and its match descendants (results in the variables below)
			uint32_t n0 = 0, n1 = 0, n2 = 0;   // pushes stored per target slot
			int net = 0, maxpre = 0;           // net change of n_entries; its highest value at a pop, relative to the chain's start
			bool hit = false, brk = false, ovf = false;
			uint32_t hk = 0, hl = 0, h_tag = 0, h_ldp = 0;
			auto chain = [&](const bool WR, const int base, const uint32_t wb0, const uint32_t wb1, const uint32_t wb2, const bool hitA, const uint32_t hkA,
			                 const uint32_t hlA) {
				uint32_t k = e.x, l = e.y, ldp = e.z >> 16;
				int i = (int)(e.z & 0xffffu);
				uint32_t mm = e.w & 0xffu, go = (e.w >> 8) & 0xffu, ge = (e.w >> 16) & 0xffu, st = (e.w >> 24) & 3u;
				const uint32_t a = (e.w >> 26) & 1u;
				const DevIndex &ix = B.ix[1 - a];
				const uint2 *const cx = B.ctx + (size_t)w_off + (size_t)a * WSTRIDE(len);
				int pre = 0;
				n0 = n1 = n2 = 0; maxpre = 0; hit = brk = ovf = false;
				steps = 0;
				auto push = [&](uint32_t pk, uint32_t pl, uint32_t ppos, uint32_t ptag, int delta, int slot) {
					++pre; // gap_push: ++stack->n_entries (bwtgap.c:62)
					const int sb = s + delta;
					if (n_aln > 0 && !nonstop && sb > best_score + O.s_mm) return; // phantom: can only ever end the search when popped
					if (sb >= WK_HITS) { ovf = true; return; }
					const uint32_t idx = slot == 0 ? n0++ : slot == 1 ? n1++ : n2++;
					if (WR) {
						const uint32_t p = (slot == 0 ? wb0 : slot == 1 ? wb1 : wb2) + idx;
						const uint32_t *sg = seg + slot * (WK_SEG + 1);
						const uint32_t c = sg[1 + (p >> ARENA_CHUNK_LOG) - sg[0]];
						xent[((size_t)c << ARENA_CHUNK_LOG) + (p & (ARENA_CHUNK - 1))] = make_uint4(pk, pl, ppos, ptag);
					}
				};
				bool exact = false; // walking the exact-match tail (bwt_match_exact_alt, bwt.c:237-252), one base per iteration
				uint32_t xc = 0;    // ... and the base it matches next
				for (;;) {
					// one memory round trip per iteration whatever the lane is doing: both occurrence blocks, and the
					// node's context word (expansion) or the next read base (exact tail), are issued before any is used
					const uint32_t jk = occ_arg(ix, k - 1), jl = occ_arg(ix, l);
					const OccBlock ob_l = load_block(ix, jl >> 6);
					const OccBlock ob_k = load_block(ix, jk >> 6);
					uint2 cw = make_uint2(0u, 0u);
					if (exact) xc = (uint32_t)(seqp[i - 1] >> (a << 2)) & 15u;
					else cw = cx[i]; // k_ctx: width[i-1], width[i-2].bid, the seed widths, str[i-1], str[i-2]
					int m = 0;
					if (!exact) {
						// loop top of bwtgap.c:139-141 for this pop
						++steps;
						if (pre > maxpre) maxpre = pre;
						if (WR && base + pre > limit) { brk = true; break; }
						--pre; // gap_pop
						m = max_diff - (int)(mm + go);
						if (gape_mode) m -= (int)ge;
						if (m < 0) break;
						if (i > 0 && m < (int)(cw.x & CW_BID)) break; // bwtgap.c:157
						if (i == 0) {
							hit = true; hk = k; hl = l; h_ldp = ldp; h_tag = mm | go << 8 | ge << 16 | a << 24;
							break;
						}
						if (m == 0 && (st == STATE_M || gape_mode || (int)ge == O.max_gape)) { // no diff allowed: exact tail
							if (WR) { // pass B: the tail pushes nothing and its outcome is known from pass A
								if (hitA) { hit = true; hk = hkA; hl = hlA; h_ldp = ldp; h_tag = mm | go << 8 | ge << 16 | a << 24; }
								break;
							}
							exact = true;
							xc = (cw.x >> 12) & 7u; // str[i-1]
						}
					}
					uint32_t nk[4], nl[4];
					{
						uint32_t ck[4], cl[4];
						occ4_in_block(ob_k, jk, ck);
						occ4_in_block(ob_l, jl, cl);
						nk[0] = ck[0] + 1; nk[1] = C1 + ck[1] + 1; nk[2] = C2 + ck[2] + 1; nk[3] = C3 + ck[3] + 1;
						nl[0] = cl[0]; nl[1] = C1 + cl[1]; nl[2] = C2 + cl[2]; nl[3] = C3 + cl[3];
					}
					if (exact) {
						if (xc > 3u) break; // an N: no match
						k = sel4(xc, nk); l = sel4(xc, nl);
						--i;
						if (k > l) break;
						if (i == 0) {
							hit = true; hk = k; hl = l; h_ldp = ldp; h_tag = mm | go << 8 | ge << 16 | a << 24;
							break;
						}
						continue;
					}
					--i;
					const uint32_t occ = l - k + 1;
					bool allow_diff = true, allow_M = true;
					if (i > 0) { // bwtgap.c:206-215; the node's context word was packed for position i + 1
						const uint32_t wb1 = cw.x & (CW_BID | WB_EQ);      // width[i]: bid, w[i-1] == w[i]
						const int b1 = (int)((cw.x >> 16) & CW_BID);       // width[i-1].bid
						if (b1 > m - 1) allow_diff = false;
						else if (b1 == m - 1 && (int)(wb1 & WB_BID) == m - 1 && (wb1 & WB_EQ)) allow_M = false;
						if (has_seed) {
							const int si = i - (len - O.seed_len);
							if (si > 0) {
								const int m_seed = m - max_diff + O.max_seed_diff;
								const uint32_t sw1 = cw.y & (CW_BID | WB_EQ);
								const int s1 = (int)((cw.y >> 16) & CW_BID);
								if (s1 > m_seed - 1) allow_diff = false;
								else if (s1 == m_seed - 1 && (int)(sw1 & WB_BID) == m_seed - 1 && (sw1 & WB_EQ)) allow_M = false;
							}
						}
					}
					const uint32_t ci = (cw.x >> 12) & 7u; // str[i]
					const uint32_t V = (nk[0] <= nl[0] ? 1u : 0u) | (nk[1] <= nl[1] ? 2u : 0u) | (nk[2] <= nl[2] ? 4u : 0u) | (nk[3] <= nl[3] ? 8u : 0u);
					const uint32_t ui = (uint32_t)i;
					if (allow_diff) { // indels (bwtgap.c:217-247)
						int tmp;
						if (loggap) {
							const uint32_t v = ge + go;
							tmp = (v ? 31 - __clz((int)v) : 0) / 2 + 1;
						} else tmp = (int)(go + ge);
						if (i >= O.indel_end_skip + tmp && len - i >= O.indel_end_skip + tmp) {
							if (st == STATE_M) {
								if ((int)go < max_gapo) {
									const uint32_t tg = mm | (go + 1) << 8 | ge << 16 | a << 26;
									push(k, l, ui | ui << 16, tg | STATE_I << 24, d_go, slot_go);
#pragma unroll
									for (uint32_t c = 0; c < 4; ++c)
										if ((V >> c) & 1u) push(nk[c], nl[c], (ui + 1) | (ui + 1) << 16, tg | STATE_D << 24, d_go, slot_go);
								}
							} else if (st == STATE_I) {
								if ((int)ge < O.max_gape) push(k, l, ui | ui << 16, mm | go << 8 | (ge + 1) << 16 | STATE_I << 24 | a << 26, d_ge, slot_ge);
							} else if ((int)ge < O.max_gape && ((int)(ge + go) < max_diff || occ < (uint32_t)O.max_del_occ)) {
								const uint32_t tg = mm | go << 8 | (ge + 1) << 16 | STATE_D << 24 | a << 26;
#pragma unroll
								for (uint32_t c = 0; c < 4; ++c)
									if ((V >> c) & 1u) push(nk[c], nl[c], (ui + 1) | (ui + 1) << 16, tg, d_ge, slot_ge);
							}
						}
					}
					if (allow_diff && allow_M) { // mismatches (bwtgap.c:248-257); the match itself is not stored (see below)
						const uint32_t tg = (mm + 1) | go << 8 | ge << 16 | STATE_M << 24 | a << 26;
#pragma unroll
						for (uint32_t j = 1; j <= 3; ++j) {
							const uint32_t c = (ci + j) & 3u;
							if ((V >> c) & 1u) push(nk[c], nl[c], ui | ui << 16, tg, d_mm, 0);
						}
						if (ci > 3u && (V & 1u)) push(nk[0], nl[0], ui | ui << 16, tg, d_mm, 0); // N: j = 4 is a mismatch too, c = (4 + 4) & 3
					}
					// the match child: pushed last with the score of the entry just popped, hence the next pop
					if (ci > 3u || !((V >> ci) & 1u)) break;
					k = sel4(ci, nk); l = sel4(ci, nl);
					++pre; st = STATE_M; ldp = 0;
				}
				net = pre;
			};

			// ---- pass A: count
			long long ck1 = clock64();
			if (mine) chain(false, 0, 0u, 0u, 0u, false, 0u, 0u);
			team_sync<TEAM>();
			long long ck2 = clock64();
			if (WSTATS && TEAM == 1) {
				st_steps += (unsigned long long)__reduce_add_sync(0xffffffffu, mine ? steps : 0);
				st_maxsteps += (unsigned long long)__reduce_max_sync(0xffffffffu, mine ? steps : 0);
			}

			// ---- who commits: everything up to the first lane whose chain changes what later pops see
			const int netA = mine ? net : 0;
			int tot_unused;
			const int base_j = n_entries + team_incl_scan<TEAM>(netA, lane, tsm, tot_unused) - netA; // n_entries when this lane's chain starts
			const bool brkA = mine && base_j + maxpre > limit;
			int h;
			{
				const int first = -team_max<TEAM>((mine && (hit || brkA || ovf)) ? -lane : -(1 << 20), lane, tsm); // the first such lane
				h = first < T ? first : T - 1;
			}
			uint32_t N0, N1, N2, wb0, wb1, wb2;
			for (;;) {
				const bool com = lane <= h;
				int t0, t1, t2;
				const int i0 = team_incl_scan<TEAM>(com ? (int)n0 : 0, lane, tsm, t0), i1 = team_incl_scan<TEAM>(com ? (int)n1 : 0, lane, tsm, t1),
				          i2 = team_incl_scan<TEAM>(com ? (int)n2 : 0, lane, tsm, t2);
				N0 = (uint32_t)t0; N1 = (uint32_t)t1; N2 = (uint32_t)t2;
				wb0 = (uint32_t)i0 - n0; wb1 = (uint32_t)i1 - n1; wb2 = (uint32_t)i2 - n2; // offsets within this round's pushes
				const uint32_t room = (WK_SEG - 1) << ARENA_CHUNK_LOG;
				if (N0 <= room && N1 <= room && N2 <= room) break;
				if (h == 0) { overflow = true; break; } // one chain alone outgrows a round's chunk table: the thread kernel takes the read
				h = 0;
			}
			if (overflow) break;
			const bool com = lane <= h;

			// ---- commit the pops: bucket s loses its top h + 1 entries
			{
				const uint32_t newc = cs - (uint32_t)(h + 1);
				if (lane == 0) {
					uint32_t c = top[s];
					int q = (int)((cs - 1) >> ARENA_CHUNK_LOG);
					const int keep = newc ? (int)((newc - 1) >> ARENA_CHUNK_LOG) : -1;
					while (q > keep) {
						const uint32_t below = xlink[(size_t)c << ARENA_CHUNK_LOG];
						wk_free(B, cache, cache_n, c);
						c = below; --q;
					}
					top[s] = newc ? c : NIL;
					cnt[s] = newc;
				}
				if (newc == 0) mask.clear(s);
				team_sync<TEAM>();
			}

			// ---- room for the pushes: chunk ids of every chunk this round writes, per target slot
			{
				int fail = 0;
				if (lane == 0) {
					for (int t = 0; t < 3 && !fail; ++t) {
						const uint32_t N = t == 0 ? N0 : t == 1 ? N1 : N2;
						if (N == 0) continue;
						const int b = s + (t == 0 ? d_mm : t == slot_go ? d_go : d_ge);
						const uint32_t p0 = cnt[b];
						const uint32_t q0 = p0 >> ARENA_CHUNK_LOG, q1 = (p0 + N - 1) >> ARENA_CHUNK_LOG;
						const int have = p0 ? (int)((p0 - 1) >> ARENA_CHUNK_LOG) : -1; // number of the bucket's current top chunk
						uint32_t *sg = seg + t * (WK_SEG + 1);
						uint32_t cur = top[b];
						sg[0] = q0;
						for (uint32_t q = q0; q <= q1; ++q) {
							if ((int)q != have) {
								const uint32_t c = wk_alloc(B, cache, cache_n);
								if (c == NIL) { fail = 1; break; }
								xlink[(size_t)c << ARENA_CHUNK_LOG] = cur;
								cur = c;
							}
							sg[1 + q - q0] = cur;
						}
						top[b] = cur; // on failure the chunks linked so far are still reachable from top[b] and are freed with the read
						if (!fail) cnt[b] = p0 + N;
					}
				}
				fail = team_bcast<TEAM>(fail, 0, lane, tsm);
				if (fail) { overflow = true; break; } // pool dry: retried by the guaranteed pass
				team_sync<TEAM>();
				if (N0) { mask.set(s + d_mm); wb0 += cnt[s + d_mm] - N0; }
				if (N1) { const int b = s + (slot_go == 1 ? d_go : d_ge); mask.set(b); wb1 += cnt[b] - N1; }
				if (N2) { mask.set(s + d_ge); wb2 += cnt[s + d_ge] - N2; }
			}

			// ---- pass B: the committed lanes store their pushes at their final positions
			long long ck3 = clock64();
			if (com) { const bool hA = hit; const uint32_t kA = hk, lA = hl; chain(true, base_j, wb0, wb1, wb2, hA, kA, lA); }
			team_sync<TEAM>();
			long long ck4 = clock64();
			if (WSTATS) {
				++st_rounds; st_taken += (unsigned long long)T; st_com += (unsigned long long)(h + 1);
				st_ck[0] += (unsigned long long)(ck1 - ck0); st_ck[1] += (unsigned long long)(ck2 - ck1); st_ck[2] += (unsigned long long)(ck3 - ck2);
				st_ck[3] += (unsigned long long)(ck4 - ck3);
			}
			{
				int tot;
				(void)team_incl_scan<TEAM>(com ? net : 0, lane, tsm, tot);
				n_entries += tot;
				const int mx = team_max<TEAM>(com ? base_j + maxpre : 0, lane, tsm);
				if (max_entries < mx) max_entries = mx;
			}
			{
				const int flags = team_max<TEAM>(com ? ((ovf ? 2 : 0) | (brk ? 1 : 0)) : 0, lane, tsm); // only lane h can have either
				if (flags & 2) { overflow = true; break; }
				if (flags & 1) break; // bwtgap.c:140
			}

			// ---- the hit of lane h, if its chain ended in one (bwtgap.c:167-200), processed by that lane alone
			if (team_bcast<TEAM>((int)hit, h, lane, tsm)) {
				int stop = 0;
				if (lane == h) {
					const int mmh = (int)(h_tag & 0xffu), goh = (int)((h_tag >> 8) & 0xffu), geh = (int)((h_tag >> 16) & 0xffu);
					const uint32_t a = (h_tag >> 24) & 1u;
					const int score = s;
					bool do_add = true;
					if (n_aln == 0) {
						best_score = score;
						int best_diff = mmh + goh;
						if (gape_mode) best_diff += geh;
						if (!nonstop) max_diff = (best_diff + 1 > rd_maxdiff) ? rd_maxdiff : best_diff + 1; // top2 behaviour
					}
					if (score == best_score) best_cnt += (int)(hl - hk + 1);
					else if (best_cnt > O.max_top2) { stop = 1; do_add = false; } // top2b behaviour
					if (do_add && goh) { // the hit may have been found already (gap in a tandem repeat)
						uint32_t c = top[WK_HITS];
						for (int q = (n_aln - 1) >> ARENA_CHUNK_LOG; q >= 0 && do_add && n_aln > 0; --q) {
							const int lo = q << ARENA_CHUNK_LOG, hi = min(n_aln, lo + (int)ARENA_CHUNK);
							for (int p = lo; p < hi; ++p) {
								const uint4 qh = __ldcg(xent + ((size_t)c << ARENA_CHUNK_LOG) + (p & (ARENA_CHUNK - 1)));
								if (qh.y == hk && qh.z == hl) { do_add = false; break; }
							}
							c = xlink[(size_t)c << ARENA_CHUNK_LOG];
						}
					}
					if (do_add) {
						// gap_shadow (bwtgap.c:81-91) on the searched strand's width array, then the packed view and the context words
						const size_t wo = (size_t)w_off + (size_t)a * WSTRIDE(len);
						uint32_t *w = B.w + wo;
						uint16_t *wb = B.bid + wo;
						uint2 *cx = B.ctx + wo;
						const uint32_t x = hl - hk + 1, mx = B.ix[1 - a].seq_len;
						const int ldp = (int)h_ldp;
						uint32_t j = 0, prev = 0, prev_nb = 0;
						for (int t = 0; t <= ldp && t <= len; ++t) {
							uint32_t wv = w[t], bid = wb[t] & WB_BID;
							if (t < ldp) {
								if (wv > x) { wv -= x; w[t] = wv; }
								else if (wv == x) { bid = 1; wv = mx - (++j); w[t] = wv; }
							}
							const uint32_t nb = bid | ((t > 0 && wv == prev) ? WB_EQ : 0u);
							wb[t] = (uint16_t)nb;
							if (t + 1 <= len) {
								uint32_t *p = &cx[t + 1].x;
								*p = (*p & CW_KEEP) | cw_of(nb) | (cw_of(prev_nb) & CW_BID) << 16;
							}
							prev = wv; prev_nb = nb;
						}
						// append to the read's hit list
						uint32_t c = top[WK_HITS];
						if ((n_aln & (int)(ARENA_CHUNK - 1)) == 0) {
							const uint32_t nc = wk_alloc(B, cache, cache_n);
							if (nc == NIL) stop = 2;
							else { xlink[(size_t)nc << ARENA_CHUNK_LOG] = c; top[WK_HITS] = c = nc; }
						}
						if (stop != 2) {
							xent[((size_t)c << ARENA_CHUNK_LOG) + ((uint32_t)n_aln & (ARENA_CHUNK - 1))] =
								make_uint4((uint32_t)mmh | (uint32_t)goh << 8 | (uint32_t)geh << 16 | a << 24, hk, hl, (uint32_t)score);
							++n_aln;
							cnt[WK_HITS] = (uint32_t)n_aln;
						}
					}
					__threadfence_block();
				}
				best_score = team_bcast<TEAM>(best_score, h, lane, tsm);
				max_diff = team_bcast<TEAM>(max_diff, h, lane, tsm);
				best_cnt = team_bcast<TEAM>(best_cnt, h, lane, tsm);
				n_aln = team_bcast<TEAM>(n_aln, h, lane, tsm);
				stop = team_bcast<TEAM>(stop, h, lane, tsm);
				team_sync<TEAM>();
				if (stop == 2) { overflow = true; break; }
				if (stop) break;
			}
			tcap = h < T - 1 ? max(1, min(tcap >> 1, h + 1)) : min(TW, tcap << 1);
			if (WSTATS) st_ck[4] += (unsigned long long)(clock64() - ck4);
		}
		if (WSTATS && lane == 0) {
			atomicAdd(B.stats + 16, st_rounds); atomicAdd(B.stats + 17, st_taken); atomicAdd(B.stats + 18, st_com);
			atomicAdd(B.stats + 19, st_steps); atomicAdd(B.stats + 20, st_maxsteps);
			for (int q = 0; q < 5; ++q) atomicAdd(B.stats + 21 + q, st_ck[q]);
			atomicAdd(B.stats + 26, 1ull);
		}

		// ---- results, then every chunk of the read goes back to the pool
		uint32_t off = 0;
		if (!overflow && n_aln > 0) {
			if (lane == 0) off = atomicAdd(B.pool_count, (unsigned int)n_aln);
			off = (uint32_t)team_bcast<TEAM>((int)off, 0, lane, tsm);
			if (off + (uint32_t)n_aln > B.pool_cap) overflow = true;
		}
		if (overflow) {
			if (lane == 0) {
				B.n_aln[rid] = -1;
				const int o = atomicAdd(B.overflow_count, 1);
				B.overflow_ids[o] = rid;
			}
		} else {
			uint32_t c = top[WK_HITS];
			for (int q = n_aln > 0 ? (n_aln - 1) >> ARENA_CHUNK_LOG : -1; q >= 0; --q) {
				const int lo = q << ARENA_CHUNK_LOG, hi = min(n_aln, lo + (int)ARENA_CHUNK);
				for (int p = lo + lane; p < hi; p += TW)
					B.pool[off + (uint32_t)p] = __ldcg(xent + ((size_t)c << ARENA_CHUNK_LOG) + (p & (ARENA_CHUNK - 1)));
				c = xlink[(size_t)c << ARENA_CHUNK_LOG];
			}
			if (lane == 0) { B.n_aln[rid] = n_aln; B.pool_off[rid] = off; B.max_entries[rid] = max_entries; }
		}
		team_sync<TEAM>();
		{ // The chunks of the read go to the warp's cache first (the next read needs a chunk per non-empty bucket at once);
		  // what does not fit is returned in batches: the first such chunk a lane meets heads a batch, the next ones are listed in it
			uint32_t head = NIL, m = 0;
			volatile uint32_t *xw = nullptr;
			for (int b = lane; b < WK_NB; b += TW) {
				uint32_t c = top[b];
				while (c != NIL) {
					const uint32_t below = xlink[(size_t)c << ARENA_CHUNK_LOG];
					const uint32_t pos = atomicAdd(cache_n, 1u);
					if (pos < WK_CACHE) cache[pos] = c;
					else {
						atomicSub(cache_n, 1u);
						if (head == NIL) { head = c; m = 0; xw = (volatile uint32_t *)B.xnxt + ((size_t)head << ARENA_CHUNK_LOG); }
						else xw[3 + m++] = c;
						if (m == WK_BATCH_MAX) { xw[2] = m; pool_batch_push(B, head); head = NIL; }
					}
					c = below;
				}
			}
			if (head != NIL) { xw[2] = m; pool_batch_push(B, head); }
		}
		team_sync<TEAM>();
	}
}

} // namespace bwagpu
