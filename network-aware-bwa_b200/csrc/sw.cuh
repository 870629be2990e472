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

// smem per warp: read bases (len2_max bytes, padded to 4), edge_h/edge_f (len2_max+1 ints each),
// rev_h/rev_e (len1_max+2 ints each)
__global__ void __launch_bounds__(128) k_sw(const uint8_t *__restrict__ pac, const SwJob *__restrict__ jobs, int n_jobs,
                                            const uint8_t *__restrict__ reads, bwa_gpu_sw_res_t *__restrict__ res,
                                            int len1_max, int len2_max, int *work_counter)
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
			if (lane == 0) { bwa_gpu_sw_res_t r = {-1, 0, 0, 0, 0}; res[job] = r; }
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
			const int C = (cols + 31) >> 5;                     // columns per lane in this sweep
			const int my0 = col0 + lane * C;                    // my columns: my0+1 .. my0+C (clipped to len1)
			const int my_n = max(0, min(C, col0 + cols - my0));
			int H[SW_CMAX], E[SW_CMAX], R[SW_CMAX];
#pragma unroll
			for (int c = 0; c < SW_CMAX; ++c) {
				H[c] = 0; E[c] = 0;
				R[c] = c < my_n ? refb[my0 + c] : 0;
			}
			int h_out = 0, f_out = 0, diag_in = 0;
			const bool last_sb = sb == n_sb - 1;
			for (int s = 0; s < len2 + 31; ++s) {
				const int j = s - lane + 1;
				int h_in = __shfl_up_sync(full, h_out, 1);
				int f_in = __shfl_up_sync(full, f_out, 1);
				const bool row_ok = j >= 1 && j <= len2;
				if (lane == 0) {
					h_in = (sb && row_ok) ? edge_h[j] : 0;
					f_in = (sb && row_ok) ? edge_f[j] : 0;
				}
				if (row_ok && my_n > 0) {
					const int qj = q[j - 1];
					int last_h = h_in, f = f_in, diag = diag_in;
					diag_in = h_in;
#pragma unroll
					for (int c = 0; c < SW_CMAX; ++c) {
						if (c < my_n) {
							const int up = H[c];
							int h = diag + sw_sc(R[c], qj), e = 0;
							if (h < 0) h = 0;
							if (last_h > 0) { // stdaln.c:611-614
								f = f > last_h - SW_Q ? f - SW_R : last_h - SW_QR;
								if (h < f) h = f;
							}
							if (up >= SW_QR + 1) { // stdaln.c:615-619
								e = E[c] > up - SW_Q ? E[c] - SW_R : up - SW_QR;
								if (h < e) h = e;
							}
							E[c] = e; H[c] = h;
							diag = up; last_h = h;
							if (best < h) { best = h; best_i = my0 + c + 1; best_j = j; }
						}
					}
					h_out = last_h; f_out = f;
					if (!last_sb && lane == 31) { edge_h[j] = last_h; edge_f[j] = f; }
				}
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
		int start_i = 0, start_j = 0, flag = 0;
		if (score_f > 32000) flag = 1; // the reference would have rescaled (stdaln.c:587-606); not restated

		// ---------------- pass 2: reverse band from the end cell (stdaln.c:638-696), one row at a time
		if (!flag && score_f >= 1 && end_i > 0 && end_j > 0) {
			const int NEG = -(1 << 28);
			for (int i = lane; i <= end_i + 1; i += 32) { rh[i] = 0; re[i] = 0; }
			__syncwarp();
			int score_r = sw_sc(refb[end_i - 1], q[end_j - 1]);
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
					for (int d = 1; d < 32; d <<= 1) {
						const int o = __shfl_up_sync(full, pm, d);
						if (lane >= d && o > pm) pm = o;
					}
					int pex = __shfl_up_sync(full, pm, 1);
					if (lane == 0) pex = NEG;
					int F = F0 - SW_R * lane;
					if (lane > 0 && pex - SW_R * (lane - 1) > F) F = pex - SW_R * (lane - 1);
					const int cur = h0 > F ? h0 : F;
					if (valid) { rh[i] = cur; re[i + 1] = tee[t]; }
					// running strict maximum in cell order, and the early stop of stdaln.c:684-686
					int cm = valid ? cur : NEG, im = cm;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const int o = __shfl_up_sync(full, im, d);
						if (lane >= d && o > im) im = o;
					}
					int ex = __shfl_up_sync(full, im, 1); // maximum over everything before this cell, earlier rows included
					if (lane == 0 || ex < run) ex = run;
					const bool newmax = valid && cur > ex;
					const unsigned hit = __ballot_sync(full, newmax && cur == target);
					if (hit) {
						const int ln = __ffs((int)hit) - 1;
						score_r = target; start_i = start - (tb + ln); start_j = j;
						stop = true;
					} else {
						const int cmax = __shfl_sync(full, im, 31);
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
			r.score = flag ? -2 - flag : score_f;
			r.start_i = start_i; r.start_j = start_j; r.end_i = end_i; r.end_j = end_j;
			res[job] = r;
		}
		__syncwarp();
	}
}

// host launcher: stages jobs + reads, runs k_sw, copies results back
static int sw_batch(cudaStream_t st, const uint8_t *d_pac, int64_t l_pac, int n, const bwa_gpu_sw_job_t *jobs,
                    bwa_gpu_sw_res_t *res, int (*fail)(const char *, ...), double *kernel_ms)
{
	if (kernel_ms) *kernel_ms = 0;
	if (n == 0) return 0;
	std::vector<SwJob> hj(n);
	int len1_max = 1, len2_max = 1;
	long long q_total = 0;
	for (int i = 0; i < n; ++i) {
		const bwa_gpu_sw_job_t &j = jobs[i];
		if (j.beg < 0 || j.len < 0 || j.reglen < 0 || (j.len > 0 && !j.seq)) return fail("bwa_gpu_mate_sw: job %d is malformed", i);
		long long l = j.reglen;
		if (j.beg + l > l_pac) l = l_pac - j.beg; // bwa_sw_core copies at most up to l_pac (bwape.c:447-448)
		if (l < 0) l = 0;
		hj[i].beg = j.beg; hj[i].len1 = (int)l; hj[i].len2 = j.len; hj[i].q_off = q_total;
		q_total += j.len;
		if (hj[i].len1 > len1_max) len1_max = hj[i].len1;
		if (j.len > len2_max) len2_max = j.len;
	}
	const int q_words = (len2_max + 4) >> 2, r_words = (len1_max + 4) >> 2;
	const size_t smem = (size_t)4 * (q_words + r_words + 2 * (len2_max + 1) + 4 * (len1_max + 2)) * sizeof(int);
	if (smem > 200 * 1024) return fail("bwa_gpu_mate_sw: window %d x read %d needs %zu B of shared memory per block", len1_max, len2_max, smem);
	std::vector<uint8_t> hq((size_t)q_total + 1);
	for (int i = 0; i < n; ++i)
		for (int t = 0; t < jobs[i].len; ++t) hq[(size_t)hj[i].q_off + t] = jobs[i].seq[t] > 3 ? 4 : jobs[i].seq[t];
	SwJob *d_jobs = nullptr; uint8_t *d_q = nullptr; bwa_gpu_sw_res_t *d_res = nullptr; int *d_cnt = nullptr;
	cudaError_t e;
#define SWCK(x) do { e = (x); if (e != cudaSuccess) { cudaFree(d_jobs); cudaFree(d_q); cudaFree(d_res); cudaFree(d_cnt); \
		return fail("bwa_gpu_mate_sw: %s: %s", #x, cudaGetErrorString(e)); } } while (0)
	SWCK(cudaMalloc((void **)&d_jobs, (size_t)n * sizeof(SwJob)));
	SWCK(cudaMalloc((void **)&d_q, hq.size()));
	SWCK(cudaMalloc((void **)&d_res, (size_t)n * sizeof(bwa_gpu_sw_res_t)));
	SWCK(cudaMalloc((void **)&d_cnt, sizeof(int)));
	SWCK(cudaMemcpyAsync(d_jobs, hj.data(), (size_t)n * sizeof(SwJob), cudaMemcpyHostToDevice, st));
	SWCK(cudaMemcpyAsync(d_q, hq.data(), hq.size(), cudaMemcpyHostToDevice, st));
	SWCK(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
	SWCK(cudaFuncSetAttribute(k_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int dev = 0, n_sm = 148, bps = 1;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
	SWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_sw, 128, smem));
	if (bps < 1) bps = 1;
	int blocks = n_sm * bps;
	if (blocks > (n + 3) / 4) blocks = (n + 3) / 4;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0, st);
	k_sw<<<blocks, 128, smem, st>>>(d_pac, d_jobs, n, d_q, d_res, len1_max, len2_max, d_cnt);
	cudaEventRecord(e1, st);
	SWCK(cudaGetLastError());
	SWCK(cudaMemcpyAsync(res, d_res, (size_t)n * sizeof(bwa_gpu_sw_res_t), cudaMemcpyDeviceToHost, st));
	SWCK(cudaStreamSynchronize(st));
	{
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (kernel_ms) *kernel_ms = ms;
		cudaEventDestroy(e0); cudaEventDestroy(e1);
	}
#undef SWCK
	cudaFree(d_jobs); cudaFree(d_q); cudaFree(d_res); cudaFree(d_cnt);
	for (int i = 0; i < n; ++i)
		if (res[i].score <= -3) return fail("bwa_gpu_mate_sw: job %d hit an unsupported case (code %d)", i, res[i].score);
	return 0;
}

} // namespace bwagpu
