// bgzf_emu.cpp -- TEST INFRASTRUCTURE ONLY (see host_emu_shim.h).
// Runs bgzf::deflate_block (network-aware-bwa_b200/csrc/bgzf.cuh) on the CPU with a REAL 256-thread block: every thread
// is a coroutine (ucontext) that runs until it reaches a barrier or a warp collective, deposits its operand and yields;
// when the whole group has arrived the scheduler lets it pick up the results.  One OS thread, deterministic order.  The
// `-m "not gpu"` suite inflates the members with zlib and compares with the input.  Nothing in the product includes this.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__

#ifndef BGZF_T
#define BGZF_T 512
#endif
static const int NT = BGZF_T, NW = NT / 32;
struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 threadIdx = {0, 0, 0};
static ucontext_t g_sched, g_ctx[NT];
static int g_cur;
static bool g_done[NT];
static unsigned long long g_slot[NT];
static int g_arrived[NW + 1];
static unsigned g_gen[NW + 1];

static void lane_yield() { swapcontext(&g_ctx[g_cur], &g_sched); }
static void rendezvous(int grp)
{
	const unsigned gen = g_gen[grp];
	++g_arrived[grp];
	while (g_gen[grp] == gen) lane_yield();
}
static void collective(unsigned long long v, unsigned long long out[32])
{
	const int w = g_cur >> 5;
	g_slot[g_cur] = v;
	rendezvous(w);
	for (int i = 0; i < 32; ++i) out[i] = g_slot[w * 32 + i];
	rendezvous(w);
}
static inline void __syncthreads() { rendezvous(NW); }
static inline void __syncwarp() { unsigned long long o[32]; collective(0, o); }
static inline int emu_lane() { return g_cur & 31; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d)
{
	unsigned long long o[32]; collective((unsigned long long)v, o); const int me = emu_lane(); return me >= d ? (T)o[me - d] : v;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{
	unsigned long long o[32]; collective((unsigned long long)v, o); return (T)o[src & 31];
}
static inline unsigned __match_any_sync(unsigned, uint32_t v)
{
	unsigned long long o[32]; collective(v, o);
	unsigned r = 0;
	for (int i = 0; i < 32; ++i) if ((uint32_t)o[i] == v) r |= 1u << i;
	return r;
}
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline uint32_t __brev(uint32_t x)
{
	uint32_t r = 0;
	for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
	return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
	sh &= 31;
	return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }

#include "../../network-aware-bwa_b200/csrc/bgzf.cuh"

struct Job { const uint8_t *in; int len, level; uint8_t *out; int32_t *clen; uint32_t *tok; const uint32_t *x2n; };
static Job g_job;
static bgzf::Smem g_smem;

// the warp-per-member inflate: the block's 8 warps take the members in turn
struct WJob { const uint8_t *in; const bgzf::InfJob *jobs; int n; uint8_t *out; int *status; };
static WJob g_wjob;
static bgzf::InfWarp g_infw[NW];
static int g_mode; // 0: deflate_block, 1: inflate_member_warp

static void lane_main(int lane)
{
	threadIdx.x = (unsigned)lane;
	if (g_mode == 1) {
		const int warp = lane >> 5;
		for (int j = warp; j < g_wjob.n; j += NW) {
			const bgzf::InfJob J = g_wjob.jobs[j];
			const int st = bgzf::inflate_member_warp(g_wjob.in + J.in_off, J.in_len, g_wjob.out + J.out_off, J.out_len, g_infw[warp]);
			if ((lane & 31) == 0) g_wjob.status[j] = st;
			__syncwarp();
		}
	} else
	bgzf::deflate_block(g_smem, g_job.in, g_job.len, g_job.level, g_job.out, g_job.clen, g_job.tok, g_job.x2n);
	g_done[lane] = true;
	g_cur = lane;
	swapcontext(&g_ctx[lane], &g_sched);
}

static void run_block()
{
	static std::vector<char> stacks;
	const size_t STK = 256 << 10;
	stacks.resize(STK * NT);
	for (int g = 0; g <= NW; ++g) { g_arrived[g] = 0; g_gen[g] = 0; }
	for (int l = 0; l < NT; ++l) {
		g_done[l] = false;
		getcontext(&g_ctx[l]);
		g_ctx[l].uc_stack.ss_sp = stacks.data() + STK * l;
		g_ctx[l].uc_stack.ss_size = STK;
		g_ctx[l].uc_link = &g_sched;
		makecontext(&g_ctx[l], (void (*)())lane_main, 1, l);
	}
	for (;;) {
		int alive = 0, alive_w[NW] = {0};
		for (int l = 0; l < NT; ++l) {
			if (g_done[l]) continue;
			g_cur = l;
			threadIdx.x = (unsigned)l;
			swapcontext(&g_sched, &g_ctx[l]);
		}
		for (int l = 0; l < NT; ++l) if (!g_done[l]) { ++alive; ++alive_w[l >> 5]; }
		if (!alive) break;
		for (int w = 0; w < NW; ++w)
			if (alive_w[w] > 0 && g_arrived[w] == alive_w[w]) { g_arrived[w] = 0; ++g_gen[w]; }
		if (g_arrived[NW] == alive) { g_arrived[NW] = 0; ++g_gen[NW]; }
	}
}

// bwa_gpu_bgzf_deflate's contract on the emulated block: members back to back into out (capacity n_blocks * 65536)
extern "C" long long bgzf_emu_deflate(const uint8_t *in, long long n_bytes, int level, uint8_t *out, int32_t *clen)
{
	uint32_t x2n[32];
	bgzf::crc_x2n_table(x2n);
	std::vector<uint32_t> tok(65536);
	std::vector<uint32_t> member(bgzf::OUT_STRIDE / 4 + 16);
	std::vector<uint32_t> aligned_in(bgzf::IN_MAX / 4 + 16);
	long long total = 0;
	int k = 0;
	for (long long off = 0; off < n_bytes; off += bgzf::IN_MAX, ++k) {
		const int len = (int)(n_bytes - off < bgzf::IN_MAX ? n_bytes - off : bgzf::IN_MAX);
		memcpy(aligned_in.data(), in + off, (size_t)len);
		g_job.in = (const uint8_t *)aligned_in.data(); g_job.len = len; g_job.level = level;
		g_job.out = (uint8_t *)member.data(); g_job.clen = &clen[k]; g_job.tok = tok.data(); g_job.x2n = x2n;
		g_mode = 0;
		run_block();
		memcpy(out + total, member.data(), (size_t)clen[k]);
		total += clen[k];
	}
	return total;
}

// bwa_gpu_bgzf_inflate's kernel body on the CPU: inflate_member is one thread's serial walk, so it runs as it is.
// status[k] = its return code per member.
extern "C" int bgzf_emu_inflate(const uint8_t *in, long long n_bytes, int n_members, const long long *member_off, uint8_t *out, long long *out_off,
                                int *status)
{
	static bgzf::InfTables T;
	long long total = 0;
	int bad = 0;
	(void)n_bytes;
	for (int k = 0; k < n_members; ++k) {
		const long long a = member_off[k], b = member_off[k + 1];
		uint32_t isize;
		memcpy(&isize, in + b - 4, 4);
		out_off[k] = total;
		status[k] = isize > 65536 ? 99 : bgzf::inflate_member(in + a, (int)(b - a), out + total, (int)isize, T);
		bad += status[k] != 0;
		total += isize > 65536 ? 0 : isize;
	}
	out_off[n_members] = total;
	return bad;
}

// the product kernel's body (one warp per member) on the emulated block
extern "C" int bgzf_emu_inflate_warp(const uint8_t *in, long long n_bytes, int n_members, const long long *member_off, uint8_t *out, long long *out_off,
                                     int *status)
{
	std::vector<uint8_t> padded((size_t)n_bytes + 64, 0); // the word reader looks a few bytes past a member (the device buffer has slack too)
	std::vector<bgzf::InfJob> jobs((size_t)n_members);
	long long total = 0;
	memcpy(padded.data(), in, (size_t)n_bytes);
	for (int k = 0; k < n_members; ++k) {
		uint32_t isize;
		memcpy(&isize, in + member_off[k + 1] - 4, 4);
		if (isize > 65536) isize = 0;
		jobs[k].in_off = member_off[k]; jobs[k].in_len = (int)(member_off[k + 1] - member_off[k]);
		jobs[k].out_off = total; jobs[k].out_len = (int)isize;
		out_off[k] = total;
		total += isize;
	}
	out_off[n_members] = total;
	g_wjob.in = padded.data(); g_wjob.jobs = jobs.data(); g_wjob.n = n_members; g_wjob.out = out; g_wjob.status = status;
	g_mode = 1;
	run_block();
	int bad = 0;
	for (int k = 0; k < n_members; ++k) bad += status[k] != 0;
	return bad;
}
