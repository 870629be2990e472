// kernel_emu.cpp -- TEST INFRASTRUCTURE ONLY (see host_emu_shim.h).
// Drives the kernel bodies of kernels.cuh on the CPU, one emulated thread at a time, with
// the same host-side preparation (hostprep.h) the library uses.  Built by
// tests/test_kernel_logic.py into tests/host_emu/libkernel_emu.so.
#define BWAGPU_HOST_EMU 1
#include <stdlib.h>
#include <stdio.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../network-aware-bwa_b200/csrc/hostprep.h"

using namespace bwagpu;

static std::string g_err;
namespace bwagpu {
int hostprep_fail(const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_err = buf;
	return 1;
}
}

struct EmuIndex {
	std::vector<uint4> blk[2];
	std::vector<uint32_t> sa[2];
	DevIndex ix[2];
};

extern "C" const char *emu_last_error(void) { return g_err.c_str(); }

extern "C" void *emu_load_index(bwt_t *const bwt[2])
{
	EmuIndex *E = new EmuIndex();
	for (int s = 0; s < 2; ++s) {
		const bwt_t *b = bwt[s];
		const uint32_t n_blk = (b->seq_len >> 6) + 1;
		E->blk[s].resize(2 * (size_t)n_blk);
		blockDim.x = 1; threadIdx.x = 0;
		for (uint32_t t = 0; t < n_blk; ++t) {
			blockIdx.x = t;
			k_relayout(b->bwt, b->seq_len, n_blk, E->blk[s].data(), b->L2[1] - b->L2[0], b->L2[2] - b->L2[1],
			           b->L2[3] - b->L2[2], b->L2[4] - b->L2[3]);
		}
		DevIndex &ix = E->ix[s];
		ix.blk = E->blk[s].data();
		ix.primary = b->primary; ix.seq_len = b->seq_len; ix.n_blk = n_blk;
		for (int j = 0; j < 5; ++j) ix.L2[j] = b->L2[j];
		ix.sa = nullptr; ix.n_sa = 0; ix.sa_intv = 32;
		if (b->sa) {
			E->sa[s].assign(b->sa, b->sa + b->n_sa);
			E->sa[s][0] = 0xffffffffu;
			ix.sa = E->sa[s].data(); ix.n_sa = b->n_sa; ix.sa_intv = (uint32_t)b->sa_intv;
		}
	}
	return E;
}

extern "C" void emu_free_index(void *h) { delete (EmuIndex *)h; }

// same contract as bwa_gpu_aln_flat, pool returned through a malloc'd array the caller frees
// with emu_free; cap/aln_cap let tests exercise the overflow tiers.  stats4 (optional):
// ref fetches, own fetches, pops, pushes.
extern "C" int emu_aln_flat(void *h, int n, const uint8_t *bases, const int64_t *offs, const gap_opt_t *opt,
                            int32_t *n_aln, int32_t *max_entries, int64_t *aln_off, uint4 **pool_out,
                            uint32_t cap1, uint32_t aln_cap1, int n_slots, unsigned long long *stats8, uint32_t pool_chunks)
{
	EmuIndex *E = (EmuIndex *)h;
	MaxDiffTable mdt;
	std::vector<uint8_t> seq((size_t)(offs[n] - offs[0]) + 1);
	std::vector<ReadMeta> meta(n);
	uint64_t so = 0, wo = 0;
	uint32_t n_stacks = 1;
	for (int i = 0; i < n; ++i) {
		const int len = (int)(offs[i + 1] - offs[i]);
		const uint32_t n_amb = pack_read(seq.data() + so, bases + offs[i], len);
		uint64_t we = 0;
		if (fill_meta(len, so, wo, opt, mdt, meta[i], we, n_stacks)) return 1;
		meta[i].n_amb = n_amb;
		so += len; wo += we;
	}
	std::vector<uint32_t> w(wo + 1);
	std::vector<uint16_t> bid(wo + 1);
	std::vector<uint2> ctx(wo + 8);
	std::vector<uint16_t> ctx16(wo + 32); // k_search reads whole 16-entry sectors
	std::vector<uint32_t> pool_off(n);
	std::vector<uint4> pool((size_t)n * 64 + (1 << 16));
	std::vector<int32_t> jobs_a(n), jobs_b(n);
	int counters[4] = {0, 0, 0, 0};
	unsigned long long stats[96] = {0};
	Batch B;
	B.ix[0] = E->ix[0]; B.ix[1] = E->ix[1];
	B.opt = to_gapopt(opt);
	B.n_reads = n;
	B.seq = seq.data(); B.meta = meta.data();
	B.w = w.data(); B.bid = bid.data(); B.ctx = ctx.data(); B.ctx16 = ctx16.data();
	B.n_aln = n_aln; B.max_entries = max_entries; B.pool_off = pool_off.data();
	B.pool = pool.data(); B.pool_cap = (uint32_t)pool.size();
	B.pool_count = (unsigned int *)&counters[2];
	B.work_counter = &counters[0]; B.overflow_count = &counters[1];
	B.stats = stats;
	B.n_stacks = n_stacks;
	B.pop_cap = getenv("EMU_POP_CAP") ? (uint32_t)atoi(getenv("EMU_POP_CAP")) : 0;
	// K2
	blockDim.x = 1; threadIdx.x = 0;
	B.jobs = nullptr; B.n_jobs = n;
	for (long long t = 0; t < 4ll * n; ++t) { blockIdx.x = (unsigned)t; k_width<true>(B); }
	for (int t = 0; t < n; ++t) { blockIdx.x = (unsigned)t; k_ctx16(B); }
	if (stats8) { stats8[0] = stats[0]; stats8[1] = stats[1]; }
	stats[0] = stats[1] = 0;
	// K3 passes, as bwagpu.cu runs them: 0 = private arena only (k_search<.., false, ..>: small bucket mask, no free
	// list), 1 = private arena + pool_chunks chunks of the shared pool, 2 = guaranteed
	const uint32_t need = (uint32_t)opt->max_entries + 16u;
	const uint32_t caps[3] = {cap1, cap1, 1024u};
	(void)aln_cap1; // hits live in the arena now: no separate list capacity
	const uint32_t pool_chunks_of[3] = {1, pool_chunks, (need >> ARENA_CHUNK_LOG) + 2};
	const bool stdmode = (opt->mode & 0x15) == 0x01 && !getenv("EMU_GENERIC_MODE");
	int n_jobs = n;
	const int32_t *jobs = nullptr;
	for (int t = 0; t < 3 && n_jobs > 0; ++t) {
		const int slots = t < 2 ? n_slots : 1;
		std::vector<uint4> ent((size_t)slots * ARENA_ALLOC(caps[t]));
		std::vector<uint32_t> nxt((size_t)slots * ARENA_ALLOC(caps[t]));
		B.ent = ent.data(); B.nxt = nxt.data(); B.heads = nullptr;
		B.cap = caps[t];
		const uint32_t stride = ((need > caps[t] ? need - caps[t] : 0) >> ARENA_CHUNK_LOG) + 2;
		std::vector<uint4> xent((size_t)pool_chunks_of[t] << ARENA_CHUNK_LOG);
		std::vector<uint32_t> xnxt((size_t)pool_chunks_of[t] << ARENA_CHUNK_LOG), ctab((size_t)slots * stride);
		unsigned int x_next = 0;
		B.xent = xent.data(); B.xnxt = xnxt.data(); B.ctab = ctab.data(); B.ctab_stride = stride;
		unsigned long long x_free_top = 0xffffffffull;
		std::vector<uint32_t> x_free_next(pool_chunks_of[t] + 1);
		B.x_chunks = pool_chunks_of[t]; B.x_next = &x_next; B.x_free_top = &x_free_top; B.x_free_next = x_free_next.data();
		B.jobs = jobs; B.n_jobs = n_jobs;
		int32_t *ovf = (t & 1) ? jobs_b.data() : jobs_a.data();
		B.overflow_ids = ovf;
		counters[0] = counters[1] = 0;
		if (t > 0) {
			for (long long q = 0; q < 4ll * n_jobs; ++q) { blockIdx.x = (unsigned)q; k_width<true>(B); }
			for (int q = 0; q < n_jobs; ++q) { blockIdx.x = (unsigned)q; k_ctx16(B); }
		}
		gridDim.x = (unsigned)slots;
		for (int s = 0; s < slots; ++s) {
			blockIdx.x = (unsigned)s;
			if (t == 0) { if (stdmode) k_search<true, false, true>(B); else k_search<true, false, false>(B); }
			else { if (stdmode) k_search<true, true, true>(B); else k_search<true, true, false>(B); }
		}
		if (stats8 && t > 0) stats8[3 + t] = (unsigned long long)counters[1]; // [4] reads that needed the guaranteed pass
		if (counters[1] > 0 && t == 2) { g_err = "reads exceeded the largest tier"; return 1; }
		jobs = ovf; n_jobs = counters[1];
	}
	if (getenv("EMU_PRINT")) fprintf(stderr, "[emu] reads %d pops %llu memory pops %llu stored %llu trips %llu\n", n, stats[2], stats[10], stats[4], stats[9]);
	if (stats8) { stats8[2] = stats[0]; stats8[3] = stats[1]; stats8[7] = stats[2]; stats8[6] = stats[4]; }
	int64_t acc = 0;
	for (int i = 0; i < n; ++i) { aln_off[i] = acc; acc += n_aln[i]; }
	aln_off[n] = acc;
	uint4 *out = (uint4 *)malloc((size_t)(acc + 1) * sizeof(uint4));
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n_aln[i]; ++j) out[aln_off[i] + j] = pool[pool_off[i] + j];
	*pool_out = out;
	return 0;
}

extern "C" void emu_free(void *p) { free(p); }

extern "C" int emu_sa(void *h, long long n, const uint32_t *q, const uint8_t *which, uint32_t *out)
{
	EmuIndex *E = (EmuIndex *)h;
	IndexPair P; P.ix[0] = E->ix[0]; P.ix[1] = E->ix[1];
	blockDim.x = 1; threadIdx.x = 0;
	for (long long t = 0; t < n; ++t) { blockIdx.x = (unsigned)t; k_sa(P, n, q, which, out); }
	return 0;
}

// the library's restatement of bwa_cal_maxdiff (hostprep.h), for the long-read check of tests/test_kernel_logic.py
extern "C" int emu_cal_maxdiff(int l, double err, double thres) { return cal_maxdiff(l, err, thres); }
