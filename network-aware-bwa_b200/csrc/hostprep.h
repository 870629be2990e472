// hostprep.h -- host-side preparation shared by the library (bwagpu.cu) and the CPU
// kernel-logic emulation used by the tests (tests/host_emu/): the integer decisions the
// reference takes in floating point, and the packing of a read into the device format.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../include/bwa_gpu.h"
#include "kernels.cuh"

namespace bwagpu {

int hostprep_fail(const char *fmt, ...); // provided by the including translation unit

// ------------------------------------------------------------------ host-side integer decisions
// bwa_cal_maxdiff (bwtaln.c:37-49): smallest k whose Poisson(l*err) tail drops below
// thres.  Double arithmetic + libm exp on the HOST, as in the reference; the device only
// ever sees the resulting integer.
static inline int cal_maxdiff(int l, double err, double thres)
{
	double elambda = exp(-l * err);
	double sum = elambda, y = 1.0;
	int x = 1;
	for (int k = 1; k < 1000; ++k) {
		y *= l * err;
		x = (int)((uint32_t)x * (uint32_t)k); // the reference's `int x` overflows for k >= 13: the same bits, without the undefined behaviour
		sum += elambda * y / x;
		if (1.0 - sum < thres) return k;
	}
	return 2;
}

struct MaxDiffTable {
	float fnr = -1.f;
	std::vector<int> tab;
	int get(int len, const gap_opt_t *opt)
	{
		if (!(opt->fnr > 0.0)) return opt->max_diff;
		if (fnr != opt->fnr) { fnr = opt->fnr; tab.assign(65536, -1); }
		if (tab[len] < 0) tab[len] = cal_maxdiff(len, 0.02 /* BWA_AVG_ERR bwtaln.h:27 */, opt->fnr);
		return tab[len];
	}
};

static inline GapOpt to_gapopt(const gap_opt_t *o)
{
	GapOpt g;
	g.s_mm = o->s_mm; g.s_gapo = o->s_gapo; g.s_gape = o->s_gape; g.mode = o->mode;
	g.indel_end_skip = o->indel_end_skip; g.max_del_occ = o->max_del_occ; g.max_entries = o->max_entries;
	g.max_gape = o->max_gape; g.max_seed_diff = o->max_seed_diff; g.seed_len = o->seed_len; g.max_top2 = o->max_top2;
	return g;
}

// width-arena entries of one read: [w0 | w1 | seed_w0 | seed_w1], each padded to 8 entries (WSTRIDE)
static inline uint64_t width_entries(int len, int seed_len)
{
	return len > 0 ? 2 * (uint64_t)WSTRIDE(len) + (len > seed_len ? 2 * (uint64_t)WSTRIDE(seed_len) : 0) : 0;
}

// per-read meta (max_diff, clamped max_gapo: bwtaln.c:102-103 with n_seqs = 1) + offsets
static inline int fill_meta(int len, uint64_t seq_off, uint64_t w_off, const gap_opt_t *opt, MaxDiffTable &mdt, ReadMeta &m,
                     uint64_t &w_entries, uint32_t &n_stacks)
{
	if (len < 0 || len > 32766) return hostprep_fail("read length %d not supported (max 32766)", len);
	int md = len > 0 ? mdt.get(len, opt) : 0;
	int go = opt->max_gapo;
	if (md < go) go = md;
	// max_diff <= 126 and max_seed_diff <= 14: the compact context entries of k_search saturate their bid fields at 127 / 15
	if (md < 0 || md > 126 || go < 0 || go > 255 || opt->max_gape < 0 || opt->max_gape > 255 || opt->max_seed_diff < 0 || opt->max_seed_diff > 14)
		return hostprep_fail("option range not supported on device: max_diff=%d (<= 126) max_gapo=%d max_gape=%d max_seed_diff=%d (<= 14)", md, go,
		                     opt->max_gape, opt->max_seed_diff);
	uint32_t ns = (uint32_t)((md + 1) * opt->s_mm + (go + 1) * opt->s_gapo + (opt->max_gape + 1) * opt->s_gape);
	if (ns > 256) return hostprep_fail("score range %u exceeds 256 buckets (s_mm/s_gapo/s_gape/max_diff too large)", ns);
	if (ns > n_stacks) n_stacks = ns;
	m.seq_off = (uint32_t)seq_off;
	m.w_off = (uint32_t)w_off;
	m.len = (uint16_t)len;
	m.max_diff = (uint8_t)md;
	m.max_gapo = (uint8_t)go;
	m.n_amb = 0; // filled by pack_*
	w_entries = width_entries(len, opt->seed_len);
	return 0;
}


// byte j of a packed read = seq[0][j] | seq[1][j] << 4 (values 0..4)
// Both return the number of ambiguous bases in seq[0] (bwtgap.c:118-119 counts seq[0]).
static inline uint32_t pack_seq_pair(uint8_t *dst, const uint8_t *seq, const uint8_t *rseq, int len)
{
	uint32_t n_amb = 0;
	for (int j = 0; j < len; ++j) {
		dst[j] = (uint8_t)((seq[j] > 3 ? 4 : seq[j]) | (rseq[j] > 3 ? 4 : rseq[j]) << 4);
		n_amb += seq[j] > 3;
	}
	return n_amb;
}

// read in sequencing orientation -> seq = reversed read, rseq = reverse complement, i.e.
// rseq[j] = complement(seq[j])  (bam1_to_seq, bwaseqio.c:294-297 with is_comp = 1)
static inline uint32_t pack_read(uint8_t *dst, const uint8_t *read, int len)
{
	uint32_t n_amb = 0;
	for (int j = 0; j < len; ++j) {
		const uint8_t b = read[len - 1 - j];
		dst[j] = (uint8_t)((b > 3 ? 4 : b) | (b > 3 ? 4 : 3 - b) << 4);
		n_amb += b > 3;
	}
	return n_amb;
}

} // namespace bwagpu
