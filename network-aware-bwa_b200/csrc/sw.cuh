// sw.cuh -- K5: batched mate-rescue Smith-Waterman (device + launcher).
//
// Replaces passes 1 and 2 of aln_local_core (stdaln.c:529-696) as bwa_sw_core calls it
// (bwape.c:456: aln_param_bwa = gap open 26, extend 9, matrix aln_sm_maq: +11 / -19, N -13;
// stdaln.c:206-212,227).  One warp per job, inter-sequence parallel across warps.
//
// Pass 1 (score + end cell, the ~len1*len2 cell bulk) is an anti-diagonal wavefront: lane t
// owns a strip of C consecutive reference columns whose H/E state lives in registers, and
// handles read row j at time step j + t; the strip's right edge (H and the running F)
// moves to lane t+1 by warp shuffle.  Windows wider than 32*C_MAX columns are swept in
// column super-blocks, the edge between two super-blocks going through shared memory.
// Each cell applies the reference's update literally, including its two quirks (E is only
// carried while H above is >= q+r+1, stdaln.c:615; F is only refreshed while the cell to
// the left is positive, stdaln.c:611), so every H is the reference's H.  The end cell is
// the FIRST maximum in (read row, ref column) order (stdaln.c:623-625): lanes keep their
// first strict maximum and the warp reduces on (score desc, row asc, column asc).
//
// Pass 2 (start cell) is the reference's adaptive band walked backwards from the end cell
// (stdaln.c:638-696).  Its band limits for row j depend on the finished row j+1, so rows
// cannot be pipelined; instead each ROW is spread over the lanes.  A row is computed in two
// phases on shared memory: (A) every band cell reads the previous row's H/E and forms its
// F-less value; (B) the horizontal chain F(t+1) = max(F(t) - r, h0(t) - q - r) is a max-plus
// prefix scan (5 shuffles per 32 cells).  This reproduces the reference's H exactly: its
// `if (last_h > 0)` guard only ever skips non-positive F values, which cannot change an H >= 0.
// The running maximum, the `score_r - qr == score_f` early stop and the band update are then
// applied with the sequential first-in-order semantics (ballot + find-first).
//
// Pass 3 (banded global alignment for the CIGAR, stdaln.c:723-735) is not part of this
// call; it and the accept/reject arithmetic in doubles (bwape.c:592-600) stay on the host.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "../../include/bwa_gpu.h"
#include "sw_cell.h"

namespace bwagpu {

#define SW_Q 26
#define SW_R 9
#define SW_QR 35
#define SW_MAXSC 11
#define SW_CMAX 16 // columns per lane per super-block -> 512 reference columns per sweep

struct SwJob {
	long long beg;   // pac coordinate of the first window base
	int len1;        // window length after clipping to l_pac (bwape.c:447-448)
	int len2;        // read length
	long long q_off; // offset of the read in the packed read array
};

__device__ __forceinline__ int sw_sc(int r, int q) { return q > 3 ? -13 : (r == q ? 11 : -19); }
__device__ __forceinline__ int pac_base(const uint8_t *pac, long long k) { return (pac[k >> 2] >> ((~k & 3) << 1)) & 3; }

// One sweep of pass 1 over <= 32 * C reference columns (col0+1 .. col_end): lane t owns C consecutive columns -- the
// strip's G = H - qr and E of the previous row and its bases live in registers -- and handles read row j at step j + t - 1;
// the strip's right edge (G and the running F) moves to lane t+1 by shuffle, the edge between two sweeps through edge_h /
// edge_f.  Columns past col_end are padding: base 5 equals no read base, and a cell that only mismatches can only pass on
// values strictly below a real cell's (F < the H to its left, E < the H above, diagonal + mismatch < the diagonal), so
// padding never holds the first maximum.  The first maximum of a row's strip comes from one key per cell,
// H * 16 + (15 - column in the strip): a maximum over keys is the highest H at its lowest column.
template <int C>
__device__ __forceinline__ void sw_sweep(const uint8_t *q, const uint8_t *refb, int *edge_h, int *edge_f, const bool from_edge,
                                         const bool last_sb, const int col0, const int col_end, const int len2, const int lane,
                                         int &best, int &best_i, int &best_j)
{
	static_assert(C >= 2 && C <= 16 && (C & 1) == 0, "columns per lane");
	const unsigned full = 0xffffffffu;
	const int my0 = col0 + lane * C; // my columns: my0+1 .. my0+C
	int G[C], E[C], R[C];
#pragma unroll
	for (int c = 0; c < C; ++c) {
		G[c] = -SWC_QR; E[c] = 0;
		R[c] = my0 + c < col_end ? refb[my0 + c] : 5;
	}
	int g_out = -SWC_QR, f_out = 0, diag_in = -SWC_QR;
	for (int s = 0; s < len2 + 31; ++s) {
		const int j = s - lane + 1;
		int g_in = __shfl_up_sync(full, g_out, 1);
		int f_in = __shfl_up_sync(full, f_out, 1);
		const bool row_ok = j >= 1 && j <= len2;
		if (lane == 0) {
			g_in = (from_edge && row_ok) ? edge_h[j] - SWC_QR : -SWC_QR;
			f_in = (from_edge && row_ok) ? edge_f[j] : 0;
		}
		if (row_ok) {
			const SwRow row = swc_row(q[j - 1]);
			int left = g_in, f = f_in, diag = diag_in, key = 0;
			diag_in = g_in;
#pragma unroll
			for (int c = 0; c < C; c += 2) {
				const int up0 = G[c], up1 = G[c + 1];
				const int h0 = swc_cell(row, R[c], up0, E[c], diag, left, f);
				G[c] = h0 - SWC_QR;
				const int h1 = swc_cell(row, R[c + 1], up1, E[c + 1], up0, G[c], f);
				G[c + 1] = h1 - SWC_QR;
				diag = up1; left = G[c + 1];
				key = swc_max3(key, h0 * 16 + (15 - c), h1 * 16 + (14 - c));
			}
			g_out = left; f_out = f;
			if (!last_sb && lane == 31) { edge_h[j] = left + SWC_QR; edge_f[j] = f; }
			if ((key >> 4) > best) { best = key >> 4; best_j = j; best_i = my0 + 16 - (key & 15); }
		}
	}
}

// smem per warp: read bases (len2_max bytes, padded to 4), edge_h/edge_f (len2_max+1 ints each),
// rev_h/rev_e (len1_max+2 ints each)
__global__ void __launch_bounds__(128) k_sw(const uint8_t *__restrict__ pac, const SwJob *__restrict__ jobs, int n_jobs,
                                            const uint8_t *__restrict__ reads, bwa_gpu_sw_res_t *__restrict__ res,
                                            int len1_max, int len2_max, int *work_counter, int *__restrict__ score_r_out)
{
	extern __shared__ int smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int q_words = (len2_max + 4) >> 2, r_words = (len1_max + 4) >> 2;
	const int per_warp = q_words + r_words + 2 * (len2_max + 1) + 4 * (len1_max + 2);
	int *base = smem + warp * per_warp;
	uint8_t *q = (uint8_t *)base;
	uint8_t *refb = (uint8_t *)(base + q_words); // the window's bases, unpacked once per job
	int *edge_h = base + q_words + r_words, *edge_f = edge_h + len2_max + 1;
	int *rh = edge_f + len2_max + 1, *re = rh + len1_max + 2;
	int *th0 = re + len1_max + 2, *tee = th0 + len1_max + 2; // pass 2, phase A -> phase B
	const unsigned full = 0xffffffffu;

	for (;;) {
		int job = 0;
		if (lane == 0) job = atomicAdd(work_counter, 1);
		job = __shfl_sync(full, job, 0);
		if (job >= n_jobs) break;
		const SwJob J = jobs[job];
		const int len1 = J.len1, len2 = J.len2;
		if (len1 <= 0 || len2 <= 0) { // stdaln.c:559
			if (lane == 0) { bwa_gpu_sw_res_t r = {-1, 0, 0, 0, 0}; res[job] = r; if (score_r_out) score_r_out[job] = 0; }
			continue;
		}
		for (int t = lane; t < len2; t += 32) q[t] = reads[J.q_off + t];
		for (int t = lane; t < len1; t += 32) refb[t] = (uint8_t)pac_base(pac, J.beg + t);
		__syncwarp();

		// ---------------- pass 1: forward score, wavefront over column super-blocks
		int best = 0, best_i = 0, best_j = 0;
		const int n_sb = (len1 + 32 * SW_CMAX - 1) / (32 * SW_CMAX);
		for (int sb = 0; sb < n_sb; ++sb) {
			const int col0 = sb * 32 * SW_CMAX;                 // columns col0+1 .. col0+cols
			const int cols = min(len1 - col0, 32 * SW_CMAX);
			const bool last_sb = sb == n_sb - 1;
			switch ((((cols + 31) >> 5) + 1) >> 1) {            // columns per lane, rounded up to even (warp-uniform)
			case 1: sw_sweep<2>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			case 2: sw_sweep<4>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			case 3: sw_sweep<6>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			case 4: sw_sweep<8>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			case 5: sw_sweep<10>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			case 6: sw_sweep<12>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			case 7: sw_sweep<14>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			default: sw_sweep<16>(q, refb, edge_h, edge_f, sb != 0, last_sb, col0, col0 + cols, len2, lane, best, best_i, best_j); break;
			}
			__syncwarp();
		}
		// first maximum in (row, column) order across lanes
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			const int ob = __shfl_xor_sync(full, best, d), oi = __shfl_xor_sync(full, best_i, d), oj = __shfl_xor_sync(full, best_j, d);
			const bool take = ob > best || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)));
			if (take) { best = ob; best_i = oi; best_j = oj; }
		}
		const int score_f = best, end_i = best_i, end_j = best_j;
		int start_i = 0, start_j = 0, flag = 0, score_r_final = 0;
		if (score_f > 32000) flag = 1; // the reference would have rescaled (stdaln.c:587-606); not restated

		// ---------------- pass 2: reverse band from the end cell (stdaln.c:638-696), one row at a time
		if (!flag && score_f >= 1 && end_i > 0 && end_j > 0) {
			const int NEG = -(1 << 28);
			for (int i = lane; i <= end_i + 1; i += 32) { rh[i] = 0; re[i] = 0; }
			__syncwarp();
			int score_r = sw_sc(refb[end_i - 1], q[end_j - 1]);
			score_r_final = score_r - SW_QR; // kept up to date below: what aln_local_core compares with score_f (stdaln.c:714-715)
			start_i = end_i; start_j = end_j;
			if (lane == 0) rh[end_i] = SW_QR + score_r;
			__syncwarp();
			int start = end_i - 1, end = end_i - 3;
			if (end <= 0) end = 0;
			const int target = score_f + SW_QR;
			bool stop = false;
			for (int j = end_j - 1; j != 0 && !stop; --j) {
				if (start < end) { flag = 2; break; } // never observed; the reference would run off its array
				const int qj = q[j - 1];
				const int B = start - end; // band cells i = start, start-1, ..., end+1  <->  t = 0 .. B-1
				// phase A: F-less cell values from the previous row's state
				for (int t = lane; t < B; t += 32) {
					const int i = start - t, x = i + 1;
					int h0 = rh[x] + sw_sc(refb[i - 1], qj);
					if (h0 < 0) h0 = 0;
					const int left = rh[x - 1];
					int ee = re[x] > left - SW_Q ? re[x] - SW_R : left - SW_QR;
					if (ee < 0) ee = 0;
					if (h0 < ee) h0 = ee;
					th0[t] = h0; tee[t] = ee;
				}
				__syncwarp();
				// phase B: horizontal chain by prefix scan, writes, running maximum
				int F0 = 0, run = score_r;
				for (int tb = 0; tb < B && !stop; tb += 32) {
					const int t = tb + lane;
					const bool valid = t < B;
					const int i = start - t;
					const int h0 = valid ? th0[t] : 0;
					// exclusive prefix max of h0(s) - qr + r*s over the lanes before me
					int v = valid ? h0 - SW_QR + SW_R * lane : NEG, pm = v;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) pm = max(pm, __shfl_up_sync(full, pm, d)); // a lane below d gets its own value back
					int pex = __shfl_up_sync(full, pm, 1);
					if (lane == 0) pex = NEG;
					int F = F0 - SW_R * lane;
					if (lane > 0 && pex - SW_R * (lane - 1) > F) F = pex - SW_R * (lane - 1);
					const int cur = h0 > F ? h0 : F;
					if (valid) { rh[i] = cur; re[i + 1] = tee[t]; }
					// running strict maximum in cell order, and the early stop of stdaln.c:684-686.  A cell stops the pass when it is a
					// new strict maximum AND equals the target: that is the first cell >= target, provided it equals the target
					// and nothing before the chunk reached it (a later cell == target has an earlier cell >= target before it).
					const int cm = valid ? cur : NEG;
					const unsigned ge = __ballot_sync(full, cm >= target), eq = __ballot_sync(full, cm == target);
					const bool hit = run < target && ge != 0u && ((eq >> (__ffs((int)ge) - 1)) & 1u);
					if (hit) {
						const int ln = __ffs((int)ge) - 1;
						score_r = target; start_i = start - (tb + ln); start_j = j;
						score_r_final = score_r - SW_QR;
						stop = true;
					} else {
						const int cmax = __reduce_max_sync(full, cm);
						if (cmax > run) {
							const unsigned at = __ballot_sync(full, valid && cur == cmax);
							run = cmax; start_i = start - (tb + __ffs((int)at) - 1); start_j = j;
						}
						// carry the chain into the next 32 cells
						const int Fn = F - SW_R > h0 - SW_QR ? F - SW_R : h0 - SW_QR;
						F0 = __shfl_sync(full, Fn, 31);
					}
				}
				if (stop) break;
				score_r = run;
				score_r_final = score_r - SW_QR;
				__syncwarp();
				if (lane == 0) { rh[start + 1] = 0; re[end + 1] = 0; }
				__syncwarp();
				// recalculate the boundaries of the band (stdaln.c:691-695)
				if (rh[start] <= SW_QR) --start;
				if (start <= 0) start = 0;
				end = start_i - (start_j - j) - (score_r + (start_j - j) * SW_MAXSC) / SW_R - 1;
				if (end <= 0) end = 0;
				__syncwarp();
			}
		}
		if (lane == 0) {
			bwa_gpu_sw_res_t r;
			if (score_r_out) score_r_out[job] = score_r_final;
			r.score = flag ? -2 - flag : score_f;
			r.start_i = start_i; r.start_j = start_j; r.end_i = end_i; r.end_j = end_j;
			res[job] = r;
		}
		__syncwarp();
	}
}

// ------------------------------------------------------------------ K6: banded global alignment + traceback
// aln_global_core (stdaln.c:345-525) with aln_param_bwa's scores, restated literally, one job per
// thread: the three row regimes of the band (left edge / inside / right edge), end-gap costs on the
// first and last row and column, the M/I/D traceback matrix (2 bits per state, one byte per cell) and
// the backtrace.  The CIGAR is produced straight from the backtrace (aln_path2cigar32 +
// bwa_aln_path2cigar, stdaln.c:1009-1039, bwtaln.c:396-406).  Sequential per job -- a job is only
// ~len2 x band cells -- with per-thread scratch in global memory.
#define G_INF (-1073741823)

struct GScore { int M, I, D; };

// Per-thread scratch is interleaved across the 32 lanes of a warp (element x of lane t lives at
// base[x * 32 + t]): lanes walk their jobs in near lock-step, so a row access by the warp is one
// contiguous segment instead of 32 scattered ones.
struct GRow {
	GScore *p;
	__device__ __forceinline__ GScore &operator[](int i) const { return p[(size_t)i * 32]; }
};

struct GlobalOut {
	int score, n_cigar, start_i, start_j, end_i, end_j;
};

__device__ __forceinline__ int g_setM(uint8_t &c, const GScore &p, int sc)
{
	if (p.M >= p.I) {
		if (p.M >= p.D) { c = (c & ~3u) | 0u; return p.M + sc; }
		c = (c & ~3u) | 2u; return p.D + sc;
	}
	if (p.I > p.D) { c = (c & ~3u) | 1u; return p.I + sc; }
	c = (c & ~3u) | 2u; return p.D + sc;
}
__device__ __forceinline__ int g_setI(uint8_t &c, const GScore &p, int ext)
{
	if (p.M - SW_Q > p.I) { c = (c & ~12u) | 0u; return p.M - SW_Q - ext; }
	c = (c & ~12u) | (1u << 2); return p.I - ext;
}
__device__ __forceinline__ int g_setD(uint8_t &c, const GScore &p, int ext)
{
	if (p.M - SW_Q > p.D) { c = (c & ~48u) | 0u; return p.M - SW_Q - ext; }
	c = (c & ~48u) | (2u << 4); return p.D - ext;
}

// ref base i (1-based, window-relative) = pac_base(pac, rbeg + i - 1); query base j = q[j - 1].
// cells: (len2 + 1) * width bytes; sc0/sc1: len1 + 2 GScore each; cig: len1 + len2 + 2 u16 (filled from its END).
// pac == nullptr: the reference bases are given one per byte at refb[0 .. len1) instead (bwa_gpu_global_align_seqs).
__device__ GlobalOut global_align_dev(const uint8_t *__restrict__ pac, const uint8_t *__restrict__ refb, long long rbeg, int len1, const uint8_t *__restrict__ q,
                                      int len2, int gap_end, int band, uint8_t *cells, GRow sc0, GRow sc1,
                                      uint16_t *cig, int cig_cap)
{
	GlobalOut out = {0, 0, 0, 0, 0, 0};
	if (len1 == 0 || len2 == 0) return out;
	const int eend = gap_end >= 0 ? gap_end : SW_R;
	int b1, b2;
	if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; }
	else { b1 = band; b2 = len2 - len1 + band; }
	if (b1 > len1) b1 = len1;
	if (b2 > len2) b2 = len2;
	const int width = (b1 + b2 <= len1) ? b1 + b2 + 1 : len1 + 1;
	GRow curr = sc0, last = sc1, sw;
#define GCELL(j, i) cells[((size_t)(j) * width + ((j) > b2 ? (i) - ((j) - b2) : (i))) * 32]
#define GSC(i) sw_sc(pac ? pac_base(pac, rbeg + (i) - 1) : (int)refb[(i) - 1], qj)
	int i, j, end;
	curr[0].M = 0; curr[0].I = curr[0].D = G_INF;
	for (i = 1; i < b1; ++i) {
		curr[i].M = curr[i].I = G_INF;
		curr[i].D = g_setD(GCELL(0, i), curr[i - 1], eend);
	}
	sw = curr; curr = last; last = sw;
	const int tmp_end = b2 < len2 ? b2 : len2 - 1;
	for (j = 1; j <= tmp_end + 1; ++j) { // part 1 (+ its last-row variant)
		const bool last_row = j == tmp_end + 1;
		if (last_row && !(j == len2 && b2 != len2 - 1)) break;
		const int qj = q[j - 1];
		curr[0].M = curr[0].D = G_INF;
		curr[0].I = g_setI(GCELL(j, 0), last[0], eend);
		end = (j + b1 <= len1 + 1) ? j + b1 - 1 : len1;
		for (i = 1; i != end; ++i) {
			uint8_t &c = GCELL(j, i);
			curr[i].M = g_setM(c, last[i - 1], GSC(i));
			curr[i].I = g_setI(c, last[i], SW_R);
			curr[i].D = g_setD(c, curr[i - 1], last_row ? eend : SW_R);
		}
		uint8_t &c = GCELL(j, i);
		curr[i].M = g_setM(c, last[i - 1], GSC(i));
		curr[i].D = g_setD(c, curr[i - 1], last_row ? eend : SW_R);
		if (j + b1 - 1 > len1) curr[i].I = g_setI(c, last[i], eend);
		else curr[i].I = G_INF;
		sw = curr; curr = last; last = sw;
	}
	for (; j <= len2 - b2 + 1; ++j) { // part 2
		const int qj = q[j - 1];
		curr[j - b2].M = curr[j - b2].I = curr[j - b2].D = G_INF;
		end = j + b1 - 1;
		for (i = j - b2 + 1; i != end; ++i) {
			uint8_t &c = GCELL(j, i);
			curr[i].M = g_setM(c, last[i - 1], GSC(i));
			curr[i].I = g_setI(c, last[i], SW_R);
			curr[i].D = g_setD(c, curr[i - 1], SW_R);
		}
		uint8_t &c = GCELL(j, i);
		curr[i].M = g_setM(c, last[i - 1], GSC(i));
		curr[i].D = g_setD(c, curr[i - 1], SW_R);
		curr[i].I = G_INF;
		sw = curr; curr = last; last = sw;
	}
	for (; j <= len2; ++j) { // part 3 and the last row
		const bool last_row = j == len2;
		const int qj = q[j - 1];
		curr[j - b2].M = curr[j - b2].I = curr[j - b2].D = G_INF;
		for (i = j - b2 + 1; i < len1; ++i) {
			uint8_t &c = GCELL(j, i);
			curr[i].M = g_setM(c, last[i - 1], GSC(i));
			curr[i].I = g_setI(c, last[i], SW_R);
			curr[i].D = g_setD(c, curr[i - 1], last_row ? eend : SW_R);
		}
		uint8_t &c = GCELL(j, i);
		curr[i].M = g_setM(c, last[len1 - 1], GSC(i));
		curr[i].I = g_setI(c, last[i], eend);
		curr[i].D = g_setD(c, curr[i - 1], last_row ? eend : SW_R);
		sw = curr; curr = last; last = sw;
	}
	// backtrace, emitting CIGAR runs from the end of the alignment backwards
	i = len1; j = len2;
	uint8_t cc = GCELL(j, i);
	int mx = last[len1].M, type = cc & 3, ctype = 0;
	if (last[len1].I > mx) { mx = last[len1].I; type = (cc >> 2) & 3; ctype = 1; }
	if (last[len1].D > mx) { mx = last[len1].D; type = (cc >> 4) & 3; ctype = 2; }
	out.score = mx; out.end_i = i; out.end_j = j;
	int pos = cig_cap, run_type = ctype, run_len = 1; // path[0] = (len1, len2, ctype)
	int pi = i, pj = j; // coordinates of the latest path element
	int n_path = 1;
	for (;;) {
		if (ctype == 0) { --i; --j; } else if (ctype == 1) --j; else --i;
		cc = GCELL(j, i);
		ctype = type;
		type = type == 0 ? (cc & 3) : type == 1 ? ((cc >> 2) & 3) : ((cc >> 4) & 3);
		if (!(i || j)) break; // this element is path[path_len]: not part of the path
		++n_path;
		pi = i; pj = j;
		if (ctype == run_type) ++run_len;
		else { cig[--pos] = (uint16_t)(run_type << 14 | run_len); run_type = ctype; run_len = 1; }
	}
	cig[--pos] = (uint16_t)(run_type << 14 | run_len);
	out.n_cigar = cig_cap - pos;
	// move the CIGAR to the front of the buffer in start -> end order (it already is, just shifted)
	for (int t = 0; t < out.n_cigar; ++t) cig[t] = cig[pos + t];
	out.start_i = pi; out.start_j = pj;
	(void)n_path;
#undef GCELL
#undef GSC
	return out;
}

struct PathJob { // one K6 job; mode 0: plain global alignment, mode 1: third pass of aln_local_core inside a K5 box
	long long beg;   // pac coordinate of the window -- or, without a pac (explicit reference bases), their offset in the byte array
	int len1, len2;
	long long q_off;
	long long cig_off;
};

__global__ void __launch_bounds__(128) k_global(const uint8_t *__restrict__ pac, const PathJob *__restrict__ jobs, int n_jobs,
                                                const uint8_t *__restrict__ reads, int gap_end, int band,
                                                const bwa_gpu_sw_res_t *__restrict__ boxes, const int *__restrict__ score_r,
                                                bwa_gpu_path_res_t *__restrict__ res, uint16_t *__restrict__ cigars,
                                                uint8_t *cells_all, size_t cells_stride, GScore *sc_all, size_t sc_stride,
                                                int *work_counter, const int *__restrict__ ids, const int *__restrict__ n_ids)
{
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, warp = tid >> 5, lane = tid & 31;
	uint8_t *cells = cells_all + warp * cells_stride * 32 + lane;
	GRow sc0, sc1;
	sc0.p = sc_all + warp * 2 * sc_stride * 32 + lane;
	sc1.p = sc0.p + sc_stride * 32;
	if (ids) n_jobs = *n_ids; // the jobs k_global_warp could not hold in shared memory
	for (;;) {
		int job = atomicAdd(work_counter, 1);
		if (job >= n_jobs) break;
		if (ids) job = ids[job];
		const PathJob J = jobs[job];
		uint16_t *cig = cigars + J.cig_off;
		const int cig_cap = J.len1 + J.len2 + 2;
		bwa_gpu_path_res_t r;
		r.cigar_off = J.cig_off;
		if (!boxes) { // refine_gapped_core's call (bwase.c:212)
			const GlobalOut o = pac ? global_align_dev(pac, nullptr, J.beg, J.len1, reads + J.q_off, J.len2, gap_end, band, cells, sc0, sc1, cig, cig_cap)
			                        : global_align_dev(nullptr, reads + J.beg, 0, J.len1, reads + J.q_off, J.len2, gap_end, band, cells, sc0, sc1, cig, cig_cap);
			r.score = o.score; r.n_cigar = o.n_cigar;
			r.start_i = o.start_i; r.start_j = o.start_j; r.end_i = o.end_i; r.end_j = o.end_j;
		} else { // aln_local_core's third pass (stdaln.c:723-745)
			const bwa_gpu_sw_res_t b = boxes[job];
			r.score = b.score; r.n_cigar = 0;
			r.start_i = b.start_i; r.start_j = b.start_j; r.end_i = b.end_i; r.end_j = b.end_j;
			if (b.score >= 1 && b.end_i > 0 && b.end_j > 0) { // score_f >= thres (= 1) and a local match exists
				const int score_f = b.score, sr = score_r[job];
				const int l1 = b.end_i - b.start_i + 1, l2 = b.end_j - b.start_j + 1;
				int span = (b.end_i - b.start_i > b.end_j - b.start_j ? b.end_i - b.start_i : b.end_j - b.start_j) + 1;
				GlobalOut o;
				for (int bw = 50;; bw <<= 1) {
					o = global_align_dev(pac, nullptr, J.beg + b.start_i - 1, l1, reads + J.q_off + b.start_j - 1, l2, -1, bw, cells, sc0, sc1, cig, cig_cap);
					if (o.score == sr || score_f == o.score) break;
					if (bw > span) break;
				}
				if (sr > o.score && score_f > o.score) r.score = -1; // the reference's "Potential bug" branch
				else r.score = o.score;
				r.n_cigar = o.n_cigar;
				r.start_i = o.start_i + b.start_i - 1; r.start_j = o.start_j + b.start_j - 1;
				r.end_i = o.end_i + b.start_i - 1; r.end_j = o.end_j + b.start_j - 1;
			}
		}
		res[job] = r;
	}
}

// ------------------------------------------------------------------ K6, one WARP per job, everything in shared memory
// The same recurrence as global_align_dev, reorganised so that nothing of a job touches global memory but its bases and
// its CIGAR: the two score rows (M, I, D per reference column), the window's bases and the traceback matrix (one byte per
// band cell) live in shared memory.  A row is computed by the lanes together: M and I of a cell depend on the previous row
// only (one column per lane, 32 at a time); D(i) = max(M(i-1) - q, D(i-1)) - ext along the row is a max-plus prefix scan --
// with E(i) = D(i) + i ext it is the running maximum of A(i) = M(i-1) - q + (i-1) ext, and the traceback bit of a cell
// ("came from M" iff M(i-1) - q > D(i-1), strictly: stdaln.c's set_D) is A(i) > E(i-1).  Integer arithmetic throughout, so
// every score and every bit is the sequential code's.  Lane 0 walks the traceback.  A job whose window is wider than
// GW_ROWCAP columns or whose band needs more than GW_CELLCAP cells is handed to the thread-per-job kernel above.
#define GW_ROWCAP 256
#define GW_CELLCAP 16384
struct GwSmem {
	int M[2][GW_ROWCAP + 2], I[2][GW_ROWCAP + 2], D[2][GW_ROWCAP + 2];
	uint8_t ref[GW_ROWCAP + 8];
	uint8_t cells[GW_CELLCAP];
};

__device__ __forceinline__ int gw_setM(int pM, int pI, int pD, int sc, int &bits)
{
	if (pM >= pI) {
		if (pM >= pD) { bits = 0; return pM + sc; }
		bits = 2; return pD + sc;
	}
	if (pI > pD) { bits = 1; return pI + sc; }
	bits = 2; return pD + sc;
}
__device__ __forceinline__ int gw_setI(int pM, int pI, int ext, int &bit)
{
	if (pM - SW_Q > pI) { bit = 0; return pM - SW_Q - ext; }
	bit = 1; return pI - ext;
}

// false: the job does not fit this kernel's shared memory (nothing was written)
__device__ bool global_align_warp(GwSmem &S, const uint8_t *__restrict__ pac, const uint8_t *__restrict__ refb, long long rbeg, int len1,
                                  const uint8_t *__restrict__ q, int len2, int gap_end, int band, uint16_t *cig, int cig_cap, GlobalOut &out)
{
	const int lane = threadIdx.x & 31;
	out.score = out.n_cigar = out.start_i = out.start_j = out.end_i = out.end_j = 0;
	if (len1 == 0 || len2 == 0) return true;
	const int eend = gap_end >= 0 ? gap_end : SW_R;
	int b1, b2;
	if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; }
	else { b1 = band; b2 = len2 - len1 + band; }
	if (b1 > len1) b1 = len1;
	if (b2 > len2) b2 = len2;
	const int width = (b1 + b2 <= len1) ? b1 + b2 + 1 : len1 + 1;
	if (len1 > GW_ROWCAP || (long long)(len2 + 1) * width > GW_CELLCAP) return false;
	__syncwarp();
	for (int i = lane; i < len1; i += 32) S.ref[i] = (uint8_t)(pac ? pac_base(pac, rbeg + i) : (int)refb[i]);
	const int tmp_end = b2 < len2 ? b2 : len2 - 1;
	int cur = 0;
	for (int j = 0; j <= len2; ++j) {
		// the row's regime (the three loops of aln_global_core and its first row)
		int lo, hi, hi_I, extD, qj = 0;
		bool lo_inf;
		if (j == 0) { lo = 0; hi = b1 - 1; lo_inf = false; hi_I = 0; extD = eend; }
		else {
			qj = q[j - 1];
			if (j <= tmp_end || (j == len2 && b2 == len2)) { // part 1 (and its last-row form)
				lo = 0; lo_inf = false;
				hi = (j + b1 <= len1 + 1) ? j + b1 - 1 : len1;
				hi_I = (j + b1 - 1 > len1) ? 1 : 0;
				extD = (j == tmp_end + 1) ? eend : SW_R;
			} else if (j <= len2 - b2 + 1) { // part 2
				lo = j - b2; lo_inf = true; hi = j + b1 - 1; hi_I = 0; extD = SW_R;
			} else { // part 3
				lo = j - b2; lo_inf = true; hi = len1; hi_I = 1; extD = (j == len2) ? eend : SW_R;
			}
		}
		int *cM = S.M[cur], *cI = S.I[cur], *cD = S.D[cur];
		const int *lM = S.M[cur ^ 1], *lI = S.I[cur ^ 1], *lD = S.D[cur ^ 1];
		uint8_t *crow = S.cells + (size_t)j * width - (j > b2 ? j - b2 : 0);
		// phase A: M and I of every column, the D of the row's first column
		for (int i = lo + lane; i <= hi; i += 32) {
			int m, ii, bm = 0, bi = 0;
			if (j == 0) { m = i == 0 ? 0 : G_INF; ii = G_INF; }
			else if (i == lo && lo_inf) { m = G_INF; ii = G_INF; }
			else if (i == 0) { m = G_INF; ii = gw_setI(lM[0], lI[0], eend, bi); }
			else {
				m = gw_setM(lM[i - 1], lI[i - 1], lD[i - 1], sw_sc(S.ref[i - 1], qj), bm);
				if (i < hi) ii = gw_setI(lM[i], lI[i], SW_R, bi);
				else if (hi_I) ii = gw_setI(lM[i], lI[i], eend, bi);
				else ii = G_INF;
			}
			cM[i] = m; cI[i] = ii;
			if (i == lo) cD[i] = G_INF;
			crow[i] = (uint8_t)(bm | bi << 2);
		}
		__syncwarp();
		// phase B: D along the row
		int carry = G_INF + lo * extD; // E(lo)
		for (int base = lo + 1; base <= hi; base += 32) {
			const int i = base + lane;
			const bool live = i <= hi;
			const int a = live ? cM[i - 1] - SW_Q + (i - 1) * extD : -2000000000;
			int inc = a;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(0xffffffffu, inc, d);
				if (lane >= d && o > inc) inc = o;
			}
			int before = __shfl_up_sync(0xffffffffu, inc, 1);
			before = lane == 0 ? carry : (before > carry ? before : carry); // E(i-1)
			const int e = inc > carry ? inc : carry;                        // E(i)
			if (live) {
				cD[i] = e - i * extD;
				if (!(a > before)) crow[i] |= (uint8_t)(2u << 4);
			}
			carry = __shfl_sync(0xffffffffu, e, 31);
		}
		__syncwarp();
		cur ^= 1;
	}
	// backtrace on lane 0 (the same walk as global_align_dev's)
	const int fin = cur ^ 1;
	int res[6] = {0, 0, 0, 0, 0, 0};
	if (lane == 0) {
#define WCELL(j, i) S.cells[(size_t)(j) * width + ((j) > b2 ? (i) - ((j) - b2) : (i))]
		int i = len1, j = len2;
		uint8_t cc = WCELL(j, i);
		int mx = S.M[fin][len1], type = cc & 3, ctype = 0;
		if (S.I[fin][len1] > mx) { mx = S.I[fin][len1]; type = (cc >> 2) & 3; ctype = 1; }
		if (S.D[fin][len1] > mx) { mx = S.D[fin][len1]; type = (cc >> 4) & 3; ctype = 2; }
		int pos = cig_cap, run_type = ctype, run_len = 1;
		int pi = i, pj = j;
		for (;;) {
			if (ctype == 0) { --i; --j; } else if (ctype == 1) --j; else --i;
			cc = WCELL(j, i);
			ctype = type;
			type = type == 0 ? (cc & 3) : type == 1 ? ((cc >> 2) & 3) : ((cc >> 4) & 3);
			if (!(i || j)) break;
			pi = i; pj = j;
			if (ctype == run_type) ++run_len;
			else { cig[--pos] = (uint16_t)(run_type << 14 | run_len); run_type = ctype; run_len = 1; }
		}
		cig[--pos] = (uint16_t)(run_type << 14 | run_len);
		const int n_cigar = cig_cap - pos;
		for (int t = 0; t < n_cigar; ++t) cig[t] = cig[pos + t];
		res[0] = mx; res[1] = n_cigar; res[2] = pi; res[3] = pj; res[4] = len1; res[5] = len2;
#undef WCELL
	}
	out.score = __shfl_sync(0xffffffffu, res[0], 0); out.n_cigar = __shfl_sync(0xffffffffu, res[1], 0);
	out.start_i = __shfl_sync(0xffffffffu, res[2], 0); out.start_j = __shfl_sync(0xffffffffu, res[3], 0);
	out.end_i = __shfl_sync(0xffffffffu, res[4], 0); out.end_j = __shfl_sync(0xffffffffu, res[5], 0);
	return true;
}

#define GW_WARPS 4
__global__ void __launch_bounds__(32 * GW_WARPS) k_global_warp(const uint8_t *__restrict__ pac, const PathJob *__restrict__ jobs, int n_jobs,
                                                               const uint8_t *__restrict__ reads, int gap_end, int band,
                                                               const bwa_gpu_sw_res_t *__restrict__ boxes, const int *__restrict__ score_r,
                                                               bwa_gpu_path_res_t *__restrict__ res, uint16_t *__restrict__ cigars,
                                                               int *work_counter, int *__restrict__ big_ids, int *n_big)
{
	extern __shared__ __align__(16) unsigned char gw_raw[];
	GwSmem &S = reinterpret_cast<GwSmem *>(gw_raw)[threadIdx.x >> 5];
	const int lane = threadIdx.x & 31;
	for (;;) {
		int job = 0;
		if (lane == 0) job = atomicAdd(work_counter, 1);
		job = __shfl_sync(0xffffffffu, job, 0);
		if (job >= n_jobs) break;
		const PathJob J = jobs[job];
		uint16_t *cig = cigars + J.cig_off;
		const int cig_cap = J.len1 + J.len2 + 2;
		bwa_gpu_path_res_t r;
		bool fits = true;
		r.cigar_off = J.cig_off;
		if (!boxes) { // refine_gapped_core's call (bwase.c:212)
			GlobalOut o;
			fits = pac ? global_align_warp(S, pac, nullptr, J.beg, J.len1, reads + J.q_off, J.len2, gap_end, band, cig, cig_cap, o)
			           : global_align_warp(S, nullptr, reads + J.beg, 0, J.len1, reads + J.q_off, J.len2, gap_end, band, cig, cig_cap, o);
			r.score = o.score; r.n_cigar = o.n_cigar;
			r.start_i = o.start_i; r.start_j = o.start_j; r.end_i = o.end_i; r.end_j = o.end_j;
		} else { // aln_local_core's third pass (stdaln.c:723-745)
			const bwa_gpu_sw_res_t b = boxes[job];
			r.score = b.score; r.n_cigar = 0;
			r.start_i = b.start_i; r.start_j = b.start_j; r.end_i = b.end_i; r.end_j = b.end_j;
			if (b.score >= 1 && b.end_i > 0 && b.end_j > 0) {
				const int score_f = b.score, sr = score_r[job];
				const int l1 = b.end_i - b.start_i + 1, l2 = b.end_j - b.start_j + 1;
				const int span = (b.end_i - b.start_i > b.end_j - b.start_j ? b.end_i - b.start_i : b.end_j - b.start_j) + 1;
				GlobalOut o;
				for (int bw = 50;; bw <<= 1) {
					fits = global_align_warp(S, pac, nullptr, J.beg + b.start_i - 1, l1, reads + J.q_off + b.start_j - 1, l2, -1, bw, cig, cig_cap, o);
					if (!fits) break;
					if (o.score == sr || score_f == o.score) break;
					if (bw > span) break;
				}
				if (sr > o.score && score_f > o.score) r.score = -1;
				else r.score = o.score;
				r.n_cigar = o.n_cigar;
				r.start_i = o.start_i + b.start_i - 1; r.start_j = o.start_j + b.start_j - 1;
				r.end_i = o.end_i + b.start_i - 1; r.end_j = o.end_j + b.start_j - 1;
			}
		}
		if (lane == 0) {
			if (fits) res[job] = r;
			else big_ids[atomicAdd(n_big, 1)] = job;
		}
	}
}

// device buffers that survive between calls (cudaMalloc of the K6 scratch costs more than the kernels)
struct SwScratch {
	void *p[12] = {};
	size_t cap[12] = {};
	cudaEvent_t ev_sync = nullptr; // blocking-sync event: the waiting host thread sleeps
	cudaError_t sync(cudaStream_t st)
	{
		cudaError_t e = cudaSuccess;
		if (!ev_sync) e = cudaEventCreateWithFlags(&ev_sync, cudaEventBlockingSync | cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaEventRecord(ev_sync, st);
		return e != cudaSuccess ? e : cudaEventSynchronize(ev_sync);
	}
	cudaError_t reserve(int k, size_t bytes)
	{
		if (bytes <= cap[k]) return cudaSuccess;
		if (p[k]) cudaFree(p[k]);
		p[k] = nullptr; cap[k] = 0;
		const size_t want = bytes + bytes / 4 + 256;
		const cudaError_t e = cudaMalloc(&p[k], want);
		if (e == cudaSuccess) cap[k] = want;
		return e;
	}
	void release()
	{
		for (int k = 0; k < 12; ++k) { if (p[k]) cudaFree(p[k]); p[k] = nullptr; cap[k] = 0; }
		if (ev_sync) cudaEventDestroy(ev_sync);
		ev_sync = nullptr;
	}
};

struct SwCounts { long long cells_fwd = 0, h2d = 0, d2h = 0; int launches = 0; };

// host launcher.  mode 0: K5 only (res); mode 1: K5 + third pass (pres, cigars); mode 2: plain banded global
// alignment with (gap_end, band) (pres, cigars).  kernel_ms[0] = k_sw, kernel_ms[1] = k_global (CUDA events on `st`).
static int sw_batch(SwScratch &S, cudaStream_t st, const uint8_t *d_pac, int64_t l_pac, int n, const bwa_gpu_sw_job_t *jobs, int mode,
                    int gap_end, int band, bwa_gpu_sw_res_t *res, bwa_gpu_path_res_t *pres, std::vector<uint16_t> *cigars,
                    int (*fail)(const char *, ...), double *kernel_ms, SwCounts *counts = nullptr)
{
	if (kernel_ms) kernel_ms[0] = kernel_ms[1] = 0;
	if (cigars) cigars->clear();
	if (n == 0) return 0;
	std::vector<SwJob> hj(n);
	std::vector<PathJob> pj(mode ? n : 0);
	int len1_max = 1, len2_max = 1;
	long long q_total = 0, cig_total = 0;
	for (int i = 0; i < n; ++i) {
		const bwa_gpu_sw_job_t &j = jobs[i];
		if (j.beg < 0 || j.len < 0 || j.reglen < 0 || (j.len > 0 && !j.seq)) return fail("job %d is malformed", i);
		long long l = j.reglen;
		if (j.beg + l > l_pac) l = l_pac - j.beg; // the reference copies at most up to l_pac (bwape.c:447-448, bwase.c:201-208)
		if (l < 0) l = 0;
		hj[i].beg = j.beg; hj[i].len1 = (int)l; hj[i].len2 = j.len; hj[i].q_off = q_total;
		if (mode) {
			pj[i].beg = j.beg; pj[i].len1 = (int)l; pj[i].len2 = j.len; pj[i].q_off = q_total; pj[i].cig_off = cig_total;
			cig_total += l + j.len + 2;
		}
		q_total += j.len;
		if (hj[i].len1 > len1_max) len1_max = hj[i].len1;
		if (j.len > len2_max) len2_max = j.len;
	}
	std::vector<uint8_t> hq((size_t)q_total + 1);
	for (int i = 0; i < n; ++i)
		for (int t = 0; t < jobs[i].len; ++t) hq[(size_t)hj[i].q_off + t] = jobs[i].seq[t] > 3 ? 4 : jobs[i].seq[t];
	SwJob *d_jobs = nullptr; uint8_t *d_q = nullptr; bwa_gpu_sw_res_t *d_res = nullptr; int *d_cnt = nullptr, *d_sr = nullptr;
	PathJob *d_pj = nullptr; bwa_gpu_path_res_t *d_pres = nullptr; uint16_t *d_cig = nullptr; uint8_t *d_cells = nullptr; GScore *d_sc = nullptr;
	int *d_big = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
	cudaError_t e;
	auto cleanup = [&]() {
		if (e0) cudaEventDestroy(e0);
		if (e1) cudaEventDestroy(e1);
		if (e2) cudaEventDestroy(e2);
	};
#define SWALLOC(ptr, k, bytes) do { e = S.reserve(k, bytes); if (e != cudaSuccess) { cleanup(); return fail("cudaMalloc(%zu): %s", (size_t)(bytes), cudaGetErrorString(e)); } ptr = (decltype(ptr))S.p[k]; } while (0)
#define SWCK(x) do { e = (x); if (e != cudaSuccess) { cleanup(); return fail("%s: %s", #x, cudaGetErrorString(e)); } } while (0)
	SWALLOC(d_q, 0, hq.size());
	SWALLOC(d_cnt, 1, 4 * sizeof(int)); // work counters: [0] k_sw, [1] k_global_warp, [3] k_global; [2] = jobs passed on to k_global
	SWCK(cudaMemcpyAsync(d_q, hq.data(), hq.size(), cudaMemcpyHostToDevice, st));
	SWCK(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(int), st));
	SWCK(cudaEventCreate(&e0)); SWCK(cudaEventCreate(&e1)); SWCK(cudaEventCreate(&e2));
	int dev = 0, n_sm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
	if (counts) {
		counts->h2d += (long long)hq.size() + (long long)n * (long long)(sizeof(SwJob) + (mode ? sizeof(PathJob) : 0));
		counts->d2h += (long long)n * (long long)(mode ? sizeof(bwa_gpu_path_res_t) : sizeof(bwa_gpu_sw_res_t)) + (mode ? 2 * cig_total : 0);
		if (mode != 2) for (int i = 0; i < n; ++i) counts->cells_fwd += (long long)hj[i].len1 * hj[i].len2;
		counts->launches += (mode != 2) + 2 * (mode != 0);
	}
	// ---- every allocation and staging copy first, so that the event window holds kernels only
	size_t smem = 0, cells_stride = 0, sc_stride = 0, threads = 0;
	int blocks = 0;
	if (mode != 2) { // K5
		const int q_words = (len2_max + 4) >> 2, r_words = (len1_max + 4) >> 2;
		smem = (size_t)4 * (q_words + r_words + 2 * (len2_max + 1) + 4 * (len1_max + 2)) * sizeof(int);
		if (smem > 200 * 1024) { cleanup(); return fail("window %d x read %d needs %zu B of shared memory per block", len1_max, len2_max, smem); }
		SWALLOC(d_jobs, 2, (size_t)n * sizeof(SwJob));
		SWALLOC(d_res, 3, (size_t)n * sizeof(bwa_gpu_sw_res_t));
		SWALLOC(d_sr, 4, (size_t)n * sizeof(int));
		SWCK(cudaMemcpyAsync(d_jobs, hj.data(), (size_t)n * sizeof(SwJob), cudaMemcpyHostToDevice, st));
		SWCK(cudaFuncSetAttribute(k_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		int bps = 1;
		SWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_sw, 128, smem));
		if (bps < 1) bps = 1;
		blocks = n_sm * bps;
		if (blocks > (n + 3) / 4) blocks = (n + 3) / 4;
	}
	if (mode) { // K6
		cells_stride = ((size_t)(len2_max + 1) * (len1_max + 1) + 15) & ~(size_t)15;
		sc_stride = (size_t)len1_max + 2;
		const size_t per_thread = cells_stride + 2 * sc_stride * sizeof(GScore);
		threads = std::min<size_t>((size_t)n_sm * 256, ((size_t)n + 127) / 128 * 128); // only what k_global_warp passes on comes here
		const size_t budget = (size_t)2 << 30;
		if (threads * per_thread > budget) threads = std::max<size_t>(128, budget / per_thread / 128 * 128);
		SWALLOC(d_big, 10, (size_t)n * sizeof(int));
		SWCK(cudaFuncSetAttribute(k_global_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GW_WARPS * sizeof(GwSmem))));
		SWALLOC(d_pj, 5, (size_t)n * sizeof(PathJob));
		SWALLOC(d_pres, 6, (size_t)n * sizeof(bwa_gpu_path_res_t));
		SWALLOC(d_cig, 7, (size_t)cig_total * sizeof(uint16_t) + 16);
		SWALLOC(d_cells, 8, threads * cells_stride);
		SWALLOC(d_sc, 9, threads * 2 * sc_stride * sizeof(GScore));
		SWCK(cudaMemcpyAsync(d_pj, pj.data(), (size_t)n * sizeof(PathJob), cudaMemcpyHostToDevice, st));
		cigars->resize((size_t)cig_total);
	}
	SWCK(cudaEventRecord(e0, st));
	if (mode != 2) {
		k_sw<<<blocks, 128, smem, st>>>(d_pac, d_jobs, n, d_q, d_res, len1_max, len2_max, d_cnt, d_sr);
		SWCK(cudaGetLastError());
	}
	SWCK(cudaEventRecord(e2, st));
	if (mode) {
		const int wblocks = std::min(2 * n_sm, (n + GW_WARPS - 1) / GW_WARPS);
		k_global_warp<<<wblocks, 32 * GW_WARPS, GW_WARPS * sizeof(GwSmem), st>>>(d_pac, d_pj, n, d_q, gap_end, band, mode == 1 ? d_res : nullptr,
		                                                                       d_sr, d_pres, d_cig, d_cnt + 1, d_big, d_cnt + 2);
		SWCK(cudaGetLastError());
		k_global<<<(unsigned)(threads / 128), 128, 0, st>>>(d_pac, d_pj, n, d_q, gap_end, band, mode == 1 ? d_res : nullptr, d_sr, d_pres,
		                                                   d_cig, d_cells, cells_stride, d_sc, sc_stride, d_cnt + 3, d_big, d_cnt + 2);
		SWCK(cudaGetLastError());
	}
	SWCK(cudaEventRecord(e1, st));
	if (mode != 2 && res) SWCK(cudaMemcpyAsync(res, d_res, (size_t)n * sizeof(bwa_gpu_sw_res_t), cudaMemcpyDeviceToHost, st));
	if (mode) {
		SWCK(cudaMemcpyAsync(pres, d_pres, (size_t)n * sizeof(bwa_gpu_path_res_t), cudaMemcpyDeviceToHost, st));
		if (cig_total) SWCK(cudaMemcpyAsync(cigars->data(), d_cig, (size_t)cig_total * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
	}
	SWCK(S.sync(st));
	if (kernel_ms) {
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e2); kernel_ms[0] = ms;
		cudaEventElapsedTime(&ms, e2, e1); kernel_ms[1] = ms;
	}
#undef SWCK
#undef SWALLOC
	cleanup();
	if (mode != 2 && res)
		for (int i = 0; i < n; ++i)
			if (res[i].score <= -3) return fail("job %d hit an unsupported case (code %d)", i, res[i].score);
	return 0;
}

// bwa_gpu_global_align_seqs: aln_global_core on explicit sequence pairs (both given one base per byte).  The same K6 kernel,
// its reference bases read from the byte array the reads are in instead of the packed genome.
static int ga_seqs_batch(SwScratch &S, cudaStream_t st, int n, const bwa_gpu_ga_job_t *jobs, int gap_end, int band, bwa_gpu_path_res_t *pres,
                         std::vector<uint16_t> *cigars, int (*fail)(const char *, ...), double *kernel_ms, SwCounts *counts = nullptr)
{
	if (kernel_ms) *kernel_ms = 0;
	cigars->clear();
	if (n == 0) return 0;
	std::vector<PathJob> pj(n);
	int len1_max = 1, len2_max = 1;
	long long b_total = 0, cig_total = 0;
	for (int i = 0; i < n; ++i) {
		const bwa_gpu_ga_job_t &j = jobs[i];
		if (j.reflen < 0 || j.len < 0 || (j.reflen > 0 && !j.ref) || (j.len > 0 && !j.seq)) return fail("job %d is malformed", i);
		pj[i].beg = b_total; pj[i].len1 = j.reflen; pj[i].len2 = j.len; pj[i].q_off = b_total + j.reflen; pj[i].cig_off = cig_total;
		b_total += (long long)j.reflen + j.len;
		cig_total += (long long)j.reflen + j.len + 2;
		if (j.reflen > len1_max) len1_max = j.reflen;
		if (j.len > len2_max) len2_max = j.len;
	}
	std::vector<uint8_t> hb((size_t)b_total + 1);
	for (int i = 0; i < n; ++i) {
		uint8_t *d = hb.data() + pj[i].beg;
		for (int t = 0; t < jobs[i].reflen; ++t) d[t] = jobs[i].ref[t] > 3 ? 4 : jobs[i].ref[t];
		d += jobs[i].reflen;
		for (int t = 0; t < jobs[i].len; ++t) d[t] = jobs[i].seq[t] > 3 ? 4 : jobs[i].seq[t];
	}
	uint8_t *d_b = nullptr; int *d_cnt = nullptr; PathJob *d_pj = nullptr; bwa_gpu_path_res_t *d_pres = nullptr; uint16_t *d_cig = nullptr;
	uint8_t *d_cells = nullptr; GScore *d_sc = nullptr;
	int *d_big = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	cudaError_t e;
	auto cleanup = [&]() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
#define SWALLOC(ptr, k, bytes) do { e = S.reserve(k, bytes); if (e != cudaSuccess) { cleanup(); return fail("cudaMalloc(%zu): %s", (size_t)(bytes), cudaGetErrorString(e)); } ptr = (decltype(ptr))S.p[k]; } while (0)
#define SWCK(x) do { e = (x); if (e != cudaSuccess) { cleanup(); return fail("%s: %s", #x, cudaGetErrorString(e)); } } while (0)
	int dev = 0, n_sm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
	const size_t cells_stride = ((size_t)(len2_max + 1) * (len1_max + 1) + 15) & ~(size_t)15, sc_stride = (size_t)len1_max + 2;
	const size_t per_thread = cells_stride + 2 * sc_stride * sizeof(GScore);
	size_t threads = std::min<size_t>((size_t)n_sm * 256, ((size_t)n + 127) / 128 * 128);
	const size_t budget = (size_t)2 << 30;
	if (threads * per_thread > budget) threads = std::max<size_t>(128, budget / per_thread / 128 * 128);
	SWALLOC(d_b, 0, hb.size());
	SWALLOC(d_cnt, 1, 4 * sizeof(int));
	SWALLOC(d_big, 10, (size_t)n * sizeof(int));
	SWCK(cudaFuncSetAttribute(k_global_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GW_WARPS * sizeof(GwSmem))));
	SWALLOC(d_pj, 5, (size_t)n * sizeof(PathJob));
	SWALLOC(d_pres, 6, (size_t)n * sizeof(bwa_gpu_path_res_t));
	SWALLOC(d_cig, 7, (size_t)cig_total * sizeof(uint16_t) + 16);
	SWALLOC(d_cells, 8, threads * cells_stride);
	SWALLOC(d_sc, 9, threads * 2 * sc_stride * sizeof(GScore));
	SWCK(cudaMemcpyAsync(d_b, hb.data(), hb.size(), cudaMemcpyHostToDevice, st));
	SWCK(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(int), st));
	SWCK(cudaMemcpyAsync(d_pj, pj.data(), (size_t)n * sizeof(PathJob), cudaMemcpyHostToDevice, st));
	cigars->resize((size_t)cig_total);
	if (counts) {
		counts->h2d += (long long)hb.size() + (long long)n * (long long)sizeof(PathJob);
		counts->d2h += (long long)n * (long long)sizeof(bwa_gpu_path_res_t) + 2 * cig_total;
		counts->launches += 2;
	}
	SWCK(cudaEventCreate(&e0)); SWCK(cudaEventCreate(&e1));
	SWCK(cudaEventRecord(e0, st));
	k_global_warp<<<std::min(2 * n_sm, (n + GW_WARPS - 1) / GW_WARPS), 32 * GW_WARPS, GW_WARPS * sizeof(GwSmem), st>>>(
	    nullptr, d_pj, n, d_b, gap_end, band, nullptr, nullptr, d_pres, d_cig, d_cnt + 1, d_big, d_cnt + 2);
	SWCK(cudaGetLastError());
	k_global<<<(unsigned)(threads / 128), 128, 0, st>>>(nullptr, d_pj, n, d_b, gap_end, band, nullptr, nullptr, d_pres, d_cig, d_cells, cells_stride,
	                                                   d_sc, sc_stride, d_cnt + 3, d_big, d_cnt + 2);
	SWCK(cudaGetLastError());
	SWCK(cudaEventRecord(e1, st));
	SWCK(cudaMemcpyAsync(pres, d_pres, (size_t)n * sizeof(bwa_gpu_path_res_t), cudaMemcpyDeviceToHost, st));
	if (cig_total) SWCK(cudaMemcpyAsync(cigars->data(), d_cig, (size_t)cig_total * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
	SWCK(S.sync(st));
	{
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (kernel_ms) *kernel_ms = ms;
	}
#undef SWCK
#undef SWALLOC
	cleanup();
	return 0;
}

} // namespace bwagpu
