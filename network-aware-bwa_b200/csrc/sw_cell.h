// sw_cell.h -- the forward cell of K5 (pass 1 of aln_local_core, stdaln.c:608-627), written once for the device
// (DPX instructions: VIADDMNMX / VIMNMX3 on sm_100a) and for the host (tests/host_emu/sw_emu.cpp checks the
// algebra below against the reference on a box without a GPU).
//
// The reference's cell, with q = gap open 26, r = gap extend 9, qr = 35 (aln_param_bwa):
//     h = max(0, H(j-1,i-1) + score);
//     if (H(j,i-1) > 0)   { f = f > H(j,i-1) - q ? f - r : H(j,i-1) - qr;  h = max(h, f); }   // stdaln.c:611-614
//     if (H(j-1,i) > qr)  { e = E > H(j-1,i) - q ? E - r : H(j-1,i) - qr;  h = max(h, e); }   // stdaln.c:615-619
//     else e = 0;
// Restated:
//   * `x > y - q ? x - r : y - qr` is max(x - r, y - qr) (the two arms agree on a tie).
//   * F: the guard only ever skips an f <= 0.  (After every cell f <= max(H, 0): an updated f was folded into that cell's
//     H, and a skipped update means H(j,i-1) = 0 and, by the same invariant one cell earlier, f <= 0.)  A non-positive f
//     never changes an H >= 0 and never becomes positive again (it only loses r per column, and the other arm
//     H - qr > 0 replaces it in both forms), so the update runs unconditionally here.
//   * E: the guard is observable (E <= 26 can survive below an H(j-1,i) <= 35; the reference drops it to 0), so it stays.
//   * state is kept as G = H - qr: the same number feeds F of the next column and E of the next row without a
//     subtraction of its own, and the guard of E is G > 0.  The diagonal term adds the score + qr instead.
// One cell = 2 (score select) + 1 (F) + 3 (E) + 2 (H) + 1 (G) + 1.5 (first-maximum key) lane instructions.
#pragma once

#define SWC_Q 26
#define SWC_R 9
#define SWC_QR 35

#if defined(__CUDA_ARCH__)
#define SWC_FN __device__ __forceinline__
SWC_FN int swc_addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }   // max(a + b, c)
SWC_FN int swc_max_relu(int a, int b) { return __vimax_s32_relu(a, b); }         // max(a, b, 0)
SWC_FN int swc_max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
#else
#define SWC_FN static inline
SWC_FN int swc_addmax(int a, int b, int c) { return a + b > c ? a + b : c; }
SWC_FN int swc_max_relu(int a, int b) { int m = a > b ? a : b; return m > 0 ? m : 0; }
SWC_FN int swc_max3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
#endif

// One row's constants: the read base as the reference base it can equal (7 for an N: equals nothing) and the score of a
// mismatch, both + qr (aln_sm_maq: +11 / -19, N -13; stdaln.c:206-212).
struct SwRow {
	int qm, miss_qr;
};
SWC_FN SwRow swc_row(int qj)
{
	SwRow r;
	r.qm = qj > 3 ? 7 : qj;
	r.miss_qr = (qj > 3 ? -13 : -19) + SWC_QR;
	return r;
}

// In: up_g = H(j-1,i) - qr (this column, previous row), e_up = E(j-1,i), diag_g = H(j-1,i-1) - qr, left_g = H(j,i-1) - qr,
// f = the running F of the row, rbase = reference base of the column (5 = padding: equals no read base).
// Out: H(j,i) (returned), f and e_up updated in place.
SWC_FN int swc_cell(const SwRow row, int rbase, int up_g, int &e_up, int diag_g, int left_g, int &f)
{
	const int sc = rbase == row.qm ? 11 + SWC_QR : row.miss_qr;
	f = swc_addmax(f, -SWC_R, left_g);
	int e = swc_addmax(e_up, -SWC_R, up_g);
	e = up_g > 0 ? e : 0;
	e_up = e;
	return swc_max_relu(swc_addmax(diag_g, sc, f), e);
}
