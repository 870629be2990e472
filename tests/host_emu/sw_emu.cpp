// sw_emu.cpp -- TEST INFRASTRUCTURE: pass 1 of K5 (k_sw / sw_sweep in csrc/sw.cuh) on the CPU, built around the SAME cell
// function the kernel uses (csrc/sw_cell.h): strips of C columns per lane with padding columns, G = H - qr state, the
// unconditional F update, the first-maximum keys and the reduction over lanes on (score desc, row asc, column asc).
// The wavefront's timing is not emulated -- lane t needs lane t-1's right edge of the same row, nothing else -- so a lane
// runs all its rows before the next lane starts.  tests/test_kernel_logic.py compares (score, end_i, end_j) with the
// reference's aln_local_core.
#include <cstdint>
#include <vector>
#include "../../network-aware-bwa_b200/csrc/sw_cell.h"

namespace {
const int CMAX = 16;

template <int C>
void sweep(const uint8_t *q, const uint8_t *refb, std::vector<int> &edge_h, std::vector<int> &edge_f, bool from_edge, int col0, int col_end,
           int len2, int *best, int *best_i, int *best_j)
{
	std::vector<int> in_g(len2 + 1), in_f(len2 + 1), out_g(len2 + 1), out_f(len2 + 1);
	for (int j = 1; j <= len2; ++j) {
		in_g[j] = from_edge ? edge_h[j] - SWC_QR : -SWC_QR;
		in_f[j] = from_edge ? edge_f[j] : 0;
	}
	for (int lane = 0; lane < 32; ++lane) {
		const int my0 = col0 + lane * C;
		int G[C], E[C], R[C];
		for (int c = 0; c < C; ++c) { G[c] = -SWC_QR; E[c] = 0; R[c] = my0 + c < col_end ? refb[my0 + c] : 5; }
		int diag_in = -SWC_QR;
		for (int j = 1; j <= len2; ++j) {
			const SwRow row = swc_row(q[j - 1]);
			int left = in_g[j], f = in_f[j], diag = diag_in, key = 0;
			diag_in = in_g[j];
			for (int c = 0; c < C; c += 2) {
				const int up0 = G[c], up1 = G[c + 1];
				const int h0 = swc_cell(row, R[c], up0, E[c], diag, left, f);
				G[c] = h0 - SWC_QR;
				const int h1 = swc_cell(row, R[c + 1], up1, E[c + 1], up0, G[c], f);
				G[c + 1] = h1 - SWC_QR;
				diag = up1; left = G[c + 1];
				key = swc_max3(key, h0 * 16 + (15 - c), h1 * 16 + (14 - c));
			}
			out_g[j] = left; out_f[j] = f;
			if ((key >> 4) > best[lane]) { best[lane] = key >> 4; best_j[lane] = j; best_i[lane] = my0 + 16 - (key & 15); }
		}
		in_g.swap(out_g); in_f.swap(out_f);
	}
	for (int j = 1; j <= len2; ++j) { edge_h[j] = in_g[j] + SWC_QR; edge_f[j] = in_f[j]; } // lane 31's right edge
}
} // namespace

extern "C" int sw_emu_pass1(const uint8_t *refb, int len1, const uint8_t *q, int len2, int *out /* score, end_i, end_j */)
{
	out[0] = out[1] = out[2] = 0;
	if (len1 <= 0 || len2 <= 0) { out[0] = -1; return 0; }
	int best[32] = {0}, best_i[32] = {0}, best_j[32] = {0};
	std::vector<int> edge_h(len2 + 1, 0), edge_f(len2 + 1, 0);
	const int n_sb = (len1 + 32 * CMAX - 1) / (32 * CMAX);
	for (int sb = 0; sb < n_sb; ++sb) {
		const int col0 = sb * 32 * CMAX;
		const int cols = len1 - col0 < 32 * CMAX ? len1 - col0 : 32 * CMAX;
#define SWEEP(C) sweep<C>(q, refb, edge_h, edge_f, sb != 0, col0, col0 + cols, len2, best, best_i, best_j)
		switch ((((cols + 31) >> 5) + 1) >> 1) {
		case 1: SWEEP(2); break;
		case 2: SWEEP(4); break;
		case 3: SWEEP(6); break;
		case 4: SWEEP(8); break;
		case 5: SWEEP(10); break;
		case 6: SWEEP(12); break;
		case 7: SWEEP(14); break;
		default: SWEEP(16); break;
		}
#undef SWEEP
	}
	int b = best[0], bi = best_i[0], bj = best_j[0];
	for (int l = 1; l < 32; ++l)
		if (best[l] > b || (best[l] == b && (best_j[l] < bj || (best_j[l] == bj && best_i[l] < bi)))) { b = best[l]; bi = best_i[l]; bj = best_j[l]; }
	out[0] = b; out[1] = bi; out[2] = bj;
	return 0;
}
