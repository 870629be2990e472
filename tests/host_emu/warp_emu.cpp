// warp_emu.cpp -- TEST INFRASTRUCTURE ONLY (see host_emu_shim.h).
// Runs the body of k_search_warp (network-aware-bwa_b200/csrc/search_warp.cuh) on the CPU with a REAL 32-lane warp:
// every lane is a coroutine (ucontext) with its own stack; a lane runs until it reaches a warp collective
// (__shfl*_sync, __ballot_sync, __syncwarp, __reduce_*_sync), deposits its operand and yields; when all 32 lanes have
// arrived the scheduler lets them pick up their results.  One OS thread, deterministic lane order, so plain loads and
// stores stand in for atomics.  The `-m "not gpu"` suite uses it to check the kernel's logic -- chains per lane, the
// count / scan / store passes, rollback after a hit, the chunked buckets and the chunk cache -- against the same
// golden vectors as the thread-per-read kernel.  Nothing in the product includes this file.
#define BWAGPU_HOST_EMU 1
#define BWAGPU_WARP_EMU 1
#ifndef EMU_TEAM
#define EMU_TEAM 1 // warps of the emulated block; > 1: they share one read (k_search_warp<.., TEAM = EMU_TEAM>)
#endif
#define WK_WARPS_PER_BLOCK EMU_TEAM
#include <stdlib.h>
#include <stdio.h>
#include <ucontext.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include "host_emu_shim.h"

// ---- the block: EMU_TEAM warps of 32 lanes
static const int NL = 32 * EMU_TEAM;
static ucontext_t g_sched, g_lane_ctx[NL];
static int g_cur = 0;                 // lane running now
static bool g_done[NL];
static unsigned long long g_slot[NL]; // operands of the collectives in progress
// rendezvous groups: [w] = the lanes of warp w (shuffles, ballots, __syncwarp), [EMU_TEAM] = the whole block (__syncthreads)
static int g_arrived[EMU_TEAM + 1];
static unsigned g_gen[EMU_TEAM + 1];

static void lane_yield() { swapcontext(&g_lane_ctx[g_cur], &g_sched); }

static void rendezvous(int grp)
{
	const unsigned gen = g_gen[grp];
	++g_arrived[grp];
	while (g_gen[grp] == gen) lane_yield(); // the scheduler bumps the generation once every member is here
}

// every lane of the warp deposits v, waits for the others, and gets a copy of all 32 operands
static void collective(unsigned long long v, unsigned long long out[32])
{
	const int w = g_cur >> 5;
	g_slot[g_cur] = v;
	rendezvous(w);
	for (int i = 0; i < 32; ++i) out[i] = g_slot[w * 32 + i];
	rendezvous(w); // nobody may overwrite the slots before everyone has read them
}

static inline void __syncthreads() { rendezvous(EMU_TEAM); }

static inline int emu_lane() { return g_cur & 31; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{
	unsigned long long o[32]; collective((unsigned long long)(long long)v, o); return (T)o[src & 31];
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d)
{
	unsigned long long o[32]; collective((unsigned long long)(long long)v, o); const int me = emu_lane(); return me >= d ? (T)o[me - d] : v;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m)
{
	unsigned long long o[32]; collective((unsigned long long)(long long)v, o); return (T)o[(emu_lane() ^ m) & 31];
}
static inline unsigned __ballot_sync(unsigned, bool p)
{
	unsigned long long o[32]; collective(p ? 1ull : 0ull, o); unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (unsigned)o[i] << i; return r;
}
static inline void __syncwarp() { unsigned long long o[32]; collective(0ull, o); }
static inline int __reduce_add_sync(unsigned, int v) { unsigned long long o[32]; collective((unsigned long long)(long long)v, o); int r = 0; for (int i = 0; i < 32; ++i) r += (int)o[i]; return r; }
static inline int __reduce_max_sync(unsigned, int v) { unsigned long long o[32]; collective((unsigned long long)(long long)v, o); int r = (int)o[0]; for (int i = 1; i < 32; ++i) r = std::max(r, (int)o[i]); return r; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline long long clock64() { return 0; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T, typename U> static inline T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
static inline void __threadfence_block() {}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
#define __shared__
#define __align__(x)
static uint32_t wk_smem[1 << 16];

#include "../../network-aware-bwa_b200/csrc/hostprep.h"
#include "../../network-aware-bwa_b200/csrc/search_warp.cuh"

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

static const Batch *g_batch;
static bool g_std;
static void lane_main(int lane)
{
	threadIdx.x = (unsigned)lane;
	if (g_std) k_search_warp<true, false, EMU_TEAM>(*g_batch); else k_search_warp<false, false, EMU_TEAM>(*g_batch);
	g_done[lane] = true;
	g_cur = lane;
	swapcontext(&g_lane_ctx[lane], &g_sched);
}

// one launch of a one-warp block
static void run_warp(const Batch &B, bool stdmode)
{
	static std::vector<char> stacks;
	const size_t STK = 1 << 20;
	stacks.resize(STK * NL);
	g_batch = &B; g_std = stdmode;
	for (int g = 0; g <= EMU_TEAM; ++g) { g_arrived[g] = 0; g_gen[g] = 0; }
	for (int l = 0; l < NL; ++l) {
		g_done[l] = false;
		getcontext(&g_lane_ctx[l]);
		g_lane_ctx[l].uc_stack.ss_sp = stacks.data() + STK * l;
		g_lane_ctx[l].uc_stack.ss_size = STK;
		g_lane_ctx[l].uc_link = &g_sched;
		makecontext(&g_lane_ctx[l], (void (*)())lane_main, 1, l);
	}
	blockDim.x = NL; blockIdx.x = 0; gridDim.x = 1;
	for (;;) {
		int alive = 0, alive_w[EMU_TEAM] = {0};
		for (int l = 0; l < NL; ++l) {
			if (g_done[l]) continue;
			g_cur = l;
			threadIdx.x = (unsigned)l;
			swapcontext(&g_sched, &g_lane_ctx[l]);
		}
		for (int l = 0; l < NL; ++l) if (!g_done[l]) { ++alive; ++alive_w[l >> 5]; }
		if (!alive) break;
		// a group whose every live member is at its rendezvous is released
		for (int w = 0; w < EMU_TEAM; ++w)
			if (alive_w[w] > 0 && g_arrived[w] == alive_w[w]) { g_arrived[w] = 0; ++g_gen[w]; }
		if (g_arrived[EMU_TEAM] == alive) { g_arrived[EMU_TEAM] = 0; ++g_gen[EMU_TEAM]; }
	}
}

struct EmuIndex {
	std::vector<uint4> blk[2];
	DevIndex ix[2];
};

extern "C" const char *wemu_last_error(void) { return g_err.c_str(); }

extern "C" void *wemu_load_index(bwt_t *const bwt[2])
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
	}
	return E;
}

extern "C" void wemu_free_index(void *h) { delete (EmuIndex *)h; }

// bwa_gpu_aln_flat's contract, every read through k_search_warp on one emulated warp.  pool_chunks bounds the shared
// chunk pool (reads that find it dry are counted in *n_dry and left with n_aln = -1).
extern "C" int wemu_aln_flat(void *h, int n, const uint8_t *bases, const int64_t *offs, const gap_opt_t *opt,
                             int32_t *n_aln, int32_t *max_entries, int64_t *aln_off, uint4 **pool_out, uint32_t pool_chunks, int *n_dry)
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
	std::vector<uint32_t> pool_off(n);
	std::vector<uint4> pool((size_t)n * 64 + (1 << 16));
	std::vector<int32_t> ovf(n);
	int counters[4] = {0, 0, 0, 0};
	unsigned long long stats[32] = {0};
	Batch B;
	memset(&B, 0, sizeof(B));
	B.ix[0] = E->ix[0]; B.ix[1] = E->ix[1];
	B.opt = to_gapopt(opt);
	B.n_reads = n;
	B.seq = seq.data(); B.meta = meta.data();
	B.w = w.data(); B.bid = bid.data(); B.ctx = ctx.data();
	B.n_aln = n_aln; B.max_entries = max_entries; B.pool_off = pool_off.data();
	B.pool = pool.data(); B.pool_cap = (uint32_t)pool.size();
	B.pool_count = (unsigned int *)&counters[2];
	B.work_counter = &counters[0]; B.overflow_count = &counters[1];
	B.overflow_ids = ovf.data();
	B.stats = stats;
	B.n_stacks = n_stacks;
	blockDim.x = 1; threadIdx.x = 0;
	B.jobs = nullptr; B.n_jobs = n;
	for (long long t = 0; t < 4ll * n; ++t) { blockIdx.x = (unsigned)t; k_width<false>(B); }
	for (int t = 0; t < n; ++t) { blockIdx.x = (unsigned)t; k_ctx(B); }
	std::vector<uint4> xent((size_t)pool_chunks << ARENA_CHUNK_LOG);
	std::vector<uint32_t> xnxt((size_t)pool_chunks << ARENA_CHUNK_LOG), x_free_next(pool_chunks + 1);
	unsigned int x_next = 0;
	unsigned long long x_free_top = 0xffffffffull;
	B.xent = xent.data(); B.xnxt = xnxt.data();
	B.x_chunks = pool_chunks; B.x_next = &x_next; B.x_free_top = &x_free_top; B.x_free_next = x_free_next.data();
	const bool stdmode = (opt->mode & 0x15) == 0x01;
	run_warp(B, stdmode);
	*n_dry = counters[1];
	int64_t acc = 0;
	for (int i = 0; i < n; ++i) { aln_off[i] = acc; acc += n_aln[i] > 0 ? n_aln[i] : 0; }
	aln_off[n] = acc;
	uint4 *out = (uint4 *)malloc((size_t)(acc + 1) * sizeof(uint4));
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n_aln[i]; ++j) out[aln_off[i] + j] = pool[pool_off[i] + j];
	*pool_out = out;
	return 0;
}

extern "C" void wemu_free(void *p) { free(p); }
