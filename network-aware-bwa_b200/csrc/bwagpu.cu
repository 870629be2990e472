// bwagpu.cu -- host side of libbwagpu.so: the C-ABI of include/bwa_gpu.h.
//
// One Ctx per CUDA device (its own stream, scratch arenas, pinned staging).  A batch call
// splits the reads into contiguous per-device ranges (index replicated, no collective:
// SURVEY.md §8e), each range into chunks, and runs per chunk
//     H2D -> K2 width -> K3 search (pass 0: thread per read on private arenas [-> pass 1: warp per read
//     on the shared chunk pool (search_warp.cuh) -> pass 2: thread per read with guaranteed memory] for
//     the reads whose search went deeper) -> scan + gather (read-ordered aln pool) -> D2H.
// There is no CPU fallback: without a usable device every entry point returns an error.
#include <cuda_runtime.h>
#include <malloc.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <condition_variable>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bwa_gpu.h"
#include "kernels.cuh"
#include "search_warp.cuh"
#include "sw.cuh"
#include "hostprep.h"

using namespace bwagpu;

// ------------------------------------------------------------------ errors
static thread_local std::string t_err;
static std::string g_err;
static std::mutex g_err_mu;

static int fail(const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	t_err = buf;
	std::lock_guard<std::mutex> g(g_err_mu);
	g_err = buf;
	return 1;
}

#define CK(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

namespace bwagpu {
int hostprep_fail(const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	return fail("%s", buf);
}
}

extern "C" const char *bwa_gpu_last_error(void)
{
	static thread_local std::string copy; // the caller's own copy: another thread may fail (and rewrite g_err) at any time
	std::lock_guard<std::mutex> g(g_err_mu);
	copy = g_err;
	return copy.c_str();
}

// ------------------------------------------------------------------ buffers
// BWAGPU_TRACE=1: every (re)allocation with its wall time on stderr -- cudaFree / cudaMallocHost synchronise with the device
// and with every other thread's CUDA calls, so a growing buffer in mid-run stalls all lanes.
static bool trace_on() { static const bool on = getenv("BWAGPU_TRACE") != nullptr && atoi(getenv("BWAGPU_TRACE")) != 0; return on; }
struct AllocTrace {
	const char *what; size_t bytes; std::chrono::steady_clock::time_point t0;
	AllocTrace(const char *w, size_t b) : what(w), bytes(b), t0(std::chrono::steady_clock::now()) {}
	~AllocTrace()
	{
		if (!trace_on()) return;
		static const auto T0 = std::chrono::steady_clock::now();
		const auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[trace] %s %.1f MB: %.1f ms (at %.0f ms)\n", what, (double)bytes / 1048576.0,
		        std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t1 - T0).count());
	}
};

template <typename T> struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	int reserve(size_t n)
	{
		if (n <= cap) return 0;
		size_t want = n + n / 8 + 256;
		AllocTrace tr(p ? "device buffer regrown" : "device buffer", want * sizeof(T));
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T> struct PinBuf {
	T *p = nullptr;
	size_t cap = 0;
	int reserve(size_t n, bool keep = false)
	{
		if (n <= cap) return 0;
		size_t want = n + n / 4 + 256;
		T *q = nullptr;
		AllocTrace tr(p ? "pinned buffer regrown" : "pinned buffer", want * sizeof(T));
		cudaError_t e = cudaMallocHost((void **)&q, want * sizeof(T));
		if (e != cudaSuccess) return fail("cudaMallocHost(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
		if (p) { if (keep) memcpy(q, p, cap * sizeof(T)); cudaFreeHost(p); }
		p = q; cap = want;
		return 0;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Pass { // per-pass scratch of k_search
	uint32_t cap, slots_blocks, ctab_stride; // slots = blocks * 128
	DevBuf<uint4> ent;
	DevBuf<uint32_t> nxt, heads;
};

// One Ctx = one LANE: a host thread's worth of state on one device (stream, scratch arenas,
// pinned staging).  A device gets BWAGPU_LANES lanes (default 3) that share its index; a
// batch call gives every lane a contiguous range of reads, so host-side packing/unpacking of
// one lane overlaps the kernels of another and a kernel's straggler tail overlaps the next
// kernel's start.
struct Ctx {
	int dev = 0;
	int lane = 0;
	Ctx *owner = nullptr; // lane 0 of the same device owns the index allocations
	int n_sm = 0;
	cudaStream_t st = nullptr;
	cudaEvent_t ev[16] = {};
	cudaEvent_t ev_sync = nullptr; // cudaEventBlockingSync: a waiting host thread sleeps instead of spinning on a core the host code needs
	// index
	DevIndex ix[2] = {};
	DevBuf<uint4> blk_all; // both strands' occ blocks in ONE allocation (one L2 persistence window)
	DevBuf<uint32_t> sa[2];
	DevBuf<uint8_t> pac;
	int64_t l_pac = 0;
	bool has_index = false, has_sa = false, has_pac = false;
	// chunk inputs / outputs on device
	DevBuf<uint8_t> d_seq;
	DevBuf<ReadMeta> d_meta;
	DevBuf<uint32_t> d_w;
	DevBuf<uint16_t> d_bid;
	DevBuf<uint2> d_ctx;       // 8-byte context words: only the warp pass reads them (allocated when it first runs)
	DevBuf<uint16_t> d_ctx16;  // compact per-position context of k_search
	DevBuf<int32_t> d_naln, d_maxent, d_jobs_a, d_jobs_b, d_ids, d_order;
	DevBuf<uint8_t> d_keys, d_keys2;
	DevBuf<uint32_t> d_pooloff, d_outoff;
	DevBuf<uint4> d_pool, d_out;
	DevBuf<int> d_counters;            // [0] work [1] retry-list length [2] hit-pool fill (u32) [3] chunk-pool bump
	DevBuf<unsigned long long> d_stats; // 16 counters of the STATS kernels (kernels.cuh: Batch::stats)
	DevBuf<uint8_t> d_cubtmp;
	Pass pass_buf[3];
	DevBuf<uint4> xent;           // shared overflow pool of this lane (records)
	DevBuf<uint32_t> xnxt, ctab, x_free_next;
	DevBuf<unsigned long long> x_free_top;
	uint32_t x_chunks = 0;
	// pinned staging
	PinBuf<uint8_t> h_seq;
	PinBuf<ReadMeta> h_meta;
	PinBuf<int32_t> h_naln, h_maxent;
	PinBuf<uint32_t> h_pooloff;
	PinBuf<uint4> h_out;
	PinBuf<int> h_counters;
	// K4 / K5 / K6 staging
	SwScratch sw;
	DevBuf<uint32_t> d_q, d_qo;
	DevBuf<uint8_t> d_which;
	// resident batch
	int res_n = 0;
	size_t res_w_entries = 0;
	GapOpt res_opt = {};
	uint32_t res_nstacks = 0;
	bool res_valid = false;
	int64_t res_total_aln = 0;
	// per-call stats
	bwa_gpu_stats_t stats = {};
};

static bwa_gpu_totals_t g_tot = {};
static std::mutex g_tot_mu;
#define TOT(stmt) do { std::lock_guard<std::mutex> tg_(g_tot_mu); stmt; } while (0)

// wait for everything queued on the lane's stream, without burning a host core while it runs
static cudaError_t lane_sync(Ctx *c)
{
	const cudaError_t e = cudaEventRecord(c->ev_sync, c->st);
	return e != cudaSuccess ? e : cudaEventSynchronize(c->ev_sync);
}

static std::vector<Ctx *> g_ctx; // search lanes, device by device
static std::vector<Ctx *> g_svc; // one service lane per device: K4, K5, K6
static bool g_stats_enabled = false;
// Locks.  g_life: the contexts exist and keep their index while a call holds it shared; init / destroy / load take it
// exclusively.  Calls of different kinds do not wait for each other: the search calls take a GROUP of lanes (BWAGPU_CALL_GROUPS
// groups per device, default 1: one search call at a time), K4 / K5 / K6 run on a service lane of their own under g_svc_mu,
// the BGZF codec has its own stream (bgzf.cu).  The measurement entry points (resident_*, probe) take g_life exclusively.
static std::shared_mutex g_life;
static std::mutex g_svc_mu, g_flat_mu, g_grp_mu;
static std::condition_variable g_grp_cv;
static int g_groups = 1;
static std::vector<char> g_grp_busy;
#define LIFE_SHARED std::shared_lock<std::shared_mutex> life_(g_life)
#define LIFE_EXCLUSIVE std::unique_lock<std::shared_mutex> life_(g_life)
static std::vector<uint4> g_flat_pool; // result pool of the last flat call (multi-device concat)

static uint32_t env_u32(const char *name, uint32_t dflt)
{
	const char *e = getenv(name);
	return e && atoll(e) > 0 ? (uint32_t)atoll(e) : dflt;
}

namespace bwagpu {
void bgzf_release(); // bgzf.cu
int primary_device() { LIFE_SHARED; return g_ctx.empty() ? -1 : g_ctx[0]->dev; }
void count_bgzf(int launches, double ms, int64_t bytes_in, int64_t bytes_out, int inflate)
{
	TOT(g_tot.launches += launches; g_tot.h2d_bytes += bytes_in; g_tot.d2h_bytes += bytes_out;
	    if (inflate) { g_tot.ms_inflate += ms; g_tot.inflate_bytes_in += bytes_in; g_tot.inflate_bytes_out += bytes_out; }
	    else { g_tot.ms_bgzf += ms; g_tot.bgzf_bytes_in += bytes_in; g_tot.bgzf_bytes_out += bytes_out; });
}
}

// ------------------------------------------------------------------ lifetime
extern "C" void bwa_gpu_destroy(void)
{
	bwagpu::bgzf_release();
	LIFE_EXCLUSIVE;
	std::vector<Ctx *> all(g_ctx);
	all.insert(all.end(), g_svc.begin(), g_svc.end());
	for (Ctx *c : all) {
		cudaSetDevice(c->dev);
		cudaDeviceSynchronize();
		if (!c->owner) {
			c->blk_all.release();
			for (int s = 0; s < 2; ++s) c->sa[s].release();
			c->pac.release();
		}
		c->d_seq.release(); c->d_meta.release(); c->d_w.release(); c->d_bid.release(); c->d_ctx.release(); c->d_ctx16.release();
		c->d_naln.release(); c->d_maxent.release(); c->d_jobs_a.release(); c->d_jobs_b.release(); c->d_ids.release(); c->d_order.release(); c->d_keys.release(); c->d_keys2.release();
		c->d_pooloff.release(); c->d_outoff.release(); c->d_pool.release(); c->d_out.release();
		c->d_counters.release(); c->d_stats.release(); c->d_cubtmp.release();
		for (Pass &t : c->pass_buf) { t.ent.release(); t.nxt.release(); t.heads.release(); }
		c->xent.release(); c->xnxt.release(); c->ctab.release(); c->x_free_next.release(); c->x_free_top.release();
		c->h_seq.release(); c->h_meta.release(); c->h_naln.release(); c->h_maxent.release(); c->h_pooloff.release(); c->h_out.release();
		c->h_counters.release();
		c->d_q.release(); c->d_qo.release(); c->d_which.release(); c->sw.release();
		for (auto &e : c->ev) if (e) cudaEventDestroy(e);
		if (c->ev_sync) cudaEventDestroy(c->ev_sync);
		if (c->st) cudaStreamDestroy(c->st);
		delete c;
	}
	g_ctx.clear();
	g_svc.clear();
	std::vector<uint4>().swap(g_flat_pool);
}

extern "C" int bwa_gpu_init(int n_devices, const int *device_ids)
{
	bwa_gpu_destroy();
	LIFE_EXCLUSIVE;
	{
		// The struct API hands every read a libc-allocated aln[] (the caller free()s it, bwaseqio.c:259): 10 M small blocks per call.
		// With glibc's defaults the heaps are trimmed when the caller frees them and grown again 128 KB at a time on the next
		// call (page faults + mprotect on the unpack threads).  BWAGPU_MALLOPT=1 keeps freed memory and grows in larger steps: a
		// process-wide setting, hence the HOST's choice (integration/bwa_gpu_batch.c and bench.py opt in); off by default.
		const char *e = getenv("BWAGPU_MALLOPT");
		if (e && atoi(e) != 0) {
			mallopt(M_TRIM_THRESHOLD, 1 << 30);
			mallopt(M_TOP_PAD, 64 << 20);
		}
	}
	int have = 0;
	cudaError_t e = cudaGetDeviceCount(&have);
	if (e != cudaSuccess || have == 0)
		return fail("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
	std::vector<int> ids;
	if (n_devices <= 0 || !device_ids) ids.push_back(0);
	else ids.assign(device_ids, device_ids + n_devices);
	for (int id : ids) {
		if (id < 0 || id >= have) return fail("device id %d out of range (have %d)", id, have);
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, id));
		if (prop.major < 10) return fail("device %d is sm_%d%d; libbwagpu is built for sm_100a only", id, prop.major, prop.minor);
		CK(cudaSetDevice(id));
		{
			// The path is random 32-byte sector gathers: ask L2 to fetch from DRAM at sector granularity instead of the
			// default 64 bytes (BWAGPU_L2_FETCH=64|128 restores / widens it for A/B runs).  A hint the B200 does not act on:
			// the call succeeds and reads back as set, but a load that misses still fills the whole 128-byte line (3.9 DRAM
			// sectors per random block at 32, 64 and 128 alike, profiles/r2_fetch_probe.md) -- what narrows the fill is the
			// .L2::64B qualifier on the index loads themselves (fmindex.cuh)
			const uint32_t gran = env_u32("BWAGPU_L2_FETCH", 32);
			cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
			cudaGetLastError();
		}
		const int lanes = (int)env_u32("BWAGPU_LANES", 3);
		Ctx *first = nullptr;
		for (int ln = 0; ln <= lanes; ++ln) { // the last one is the device's service lane
			Ctx *c = new Ctx();
			c->dev = id;
			c->lane = ln;
			c->owner = first;
			if (!first) first = c;
			c->n_sm = prop.multiProcessorCount;
			CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
			for (auto &ev : c->ev) CK(cudaEventCreate(&ev));
			CK(cudaEventCreateWithFlags(&c->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
			if (ln < lanes) g_ctx.push_back(c); else g_svc.push_back(c);
		}
		g_groups = (int)std::min<uint32_t>(env_u32("BWAGPU_CALL_GROUPS", 1), (uint32_t)lanes);
	}
	g_grp_busy.assign((size_t)g_groups, 0);
	return 0;
}

static int upload_index_one(Ctx *c, bwt_t *const bwt[2], const ubyte_t *pac, int64_t l_pac)
{
	CK(cudaSetDevice(c->dev));
	c->has_sa = true;
	size_t blk_off[2] = {0, 0}, blk_total = 0;
	for (int s = 0; s < 2; ++s) {
		if (!bwt[s] || !bwt[s]->bwt) return fail("bwa_gpu_load_index: bwt[%d] is not loaded", s);
		blk_off[s] = blk_total;
		blk_total += 2 * ((size_t)(bwt[s]->seq_len >> 6) + 1);
		blk_total = (blk_total + 7) & ~(size_t)7; // keep each strand 128-byte aligned
	}
	if (c->blk_all.reserve(blk_total)) return 1;
	for (int s = 0; s < 2; ++s) {
		const bwt_t *b = bwt[s];
		const uint32_t n_blk = (b->seq_len >> 6) + 1;
		DevBuf<uint32_t> raw;
		if (raw.reserve(b->bwt_size)) return 1;
		CK(cudaMemcpyAsync(raw.p, b->bwt, (size_t)b->bwt_size * 4, cudaMemcpyHostToDevice, c->st));
		k_relayout<<<(n_blk + 255) / 256, 256, 0, c->st>>>(raw.p, b->seq_len, n_blk, c->blk_all.p + blk_off[s], b->L2[1] - b->L2[0],
		                                                   b->L2[2] - b->L2[1], b->L2[3] - b->L2[2], b->L2[4] - b->L2[3]);
		CK(cudaGetLastError());
		DevIndex &ix = c->ix[s];
		ix.blk = c->blk_all.p + blk_off[s];
		ix.primary = b->primary; ix.seq_len = b->seq_len; ix.n_blk = n_blk;
		for (int j = 0; j < 5; ++j) ix.L2[j] = b->L2[j];
		ix.sa = nullptr; ix.n_sa = 0; ix.sa_intv = 32;
		if (b->sa) {
			if (b->sa_intv <= 0) return fail("bwt[%d].sa_intv = %d", s, b->sa_intv);
			if (c->sa[s].reserve(b->n_sa)) return 1;
			CK(cudaMemcpyAsync(c->sa[s].p, b->sa, (size_t)b->n_sa * 4, cudaMemcpyHostToDevice, c->st));
			// sa[0] is never read on the host path either (bwt.c:80); make it well defined
			const uint32_t m1 = 0xffffffffu;
			CK(cudaMemcpyAsync(c->sa[s].p, &m1, 4, cudaMemcpyHostToDevice, c->st));
			ix.sa = c->sa[s].p; ix.n_sa = b->n_sa; ix.sa_intv = (uint32_t)b->sa_intv;
		} else c->has_sa = false;
		CK(lane_sync(c));
		raw.release();
	}
	c->has_pac = false;
	if (pac && l_pac > 0) {
		const size_t nb = (size_t)(l_pac / 4 + 1);
		if (c->pac.reserve(nb)) return 1;
		CK(cudaMemcpyAsync(c->pac.p, pac, nb, cudaMemcpyHostToDevice, c->st));
		CK(lane_sync(c));
		c->l_pac = l_pac;
		c->has_pac = true;
	}
	// L2 persistence for the occ blocks (B200: 126 MB L2).  The window covers both strands;
	// when the index is larger than the persisting carve-out, hitRatio keeps a random
	// subset resident instead of thrashing.  Stack/width traffic is marked streaming.
	{
		int max_persist = 0, max_window = 0;
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->dev);
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->dev);
		const char *env = getenv("BWAGPU_L2_PERSIST");
		const bool on = env && atoi(env) != 0; // measured slower on B200 for a 100 MB index (profiles/): opt-in
		const size_t bytes = blk_total * sizeof(uint4);
		if (on && max_persist > 0 && max_window > 0) {
			const size_t carve = std::min((size_t)max_persist, bytes);
			if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
				cudaStreamAttrValue av;
				memset(&av, 0, sizeof av);
				av.accessPolicyWindow.base_ptr = c->blk_all.p;
				av.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)max_window);
				av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)av.accessPolicyWindow.num_bytes);
				av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
				av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
				cudaStreamSetAttribute(c->st, cudaStreamAttributeAccessPolicyWindow, &av);
			}
			cudaGetLastError(); // persistence is a hint: never fatal
		}
	}
	c->has_index = true;
	c->res_valid = false;
	return 0;
}

extern "C" int bwa_gpu_load_index(bwt_t *const bwt[2], const ubyte_t *pac, int64_t l_pac)
{
	LIFE_EXCLUSIVE;
	if (g_ctx.empty()) return fail("bwa_gpu_load_index: call bwa_gpu_init first");
	if (!bwt || !bwt[0] || !bwt[1]) return fail("bwa_gpu_load_index: bwt[0] / bwt[1] is NULL");
	if (bwt[0]->seq_len != bwt[1]->seq_len) return fail("forward and reverse index differ in seq_len");
	std::vector<Ctx *> all(g_ctx);
	all.insert(all.end(), g_svc.begin(), g_svc.end());
	for (Ctx *c : all) {
		if (!c->owner) { if (upload_index_one(c, bwt, pac, l_pac)) return 1; }
		else { // sibling lane: same device, same index arrays
			const Ctx *o = c->owner;
			c->ix[0] = o->ix[0]; c->ix[1] = o->ix[1];
			c->pac.p = o->pac.p; c->pac.cap = 0; c->l_pac = o->l_pac;
			c->has_index = o->has_index; c->has_sa = o->has_sa; c->has_pac = o->has_pac;
			c->res_valid = false;
		}
	}
	return 0;
}

extern "C" int bwa_gpu_load_pac(const ubyte_t *pac, int64_t l_pac)
{
	LIFE_EXCLUSIVE;
	if (g_ctx.empty()) return fail("bwa_gpu_load_pac: call bwa_gpu_init first");
	if (!pac || l_pac <= 0) return fail("bwa_gpu_load_pac: bad argument");
	std::vector<Ctx *> all(g_ctx);
	all.insert(all.end(), g_svc.begin(), g_svc.end());
	for (Ctx *c : all) {
		if (c->owner) { c->pac.p = c->owner->pac.p; c->l_pac = l_pac; c->has_pac = true; continue; }
		CK(cudaSetDevice(c->dev));
		const size_t nb = (size_t)(l_pac / 4 + 1);
		if (c->pac.reserve(nb)) return 1;
		CK(cudaMemcpy(c->pac.p, pac, nb, cudaMemcpyHostToDevice));
		c->l_pac = l_pac;
		c->has_pac = true;
	}
	return 0;
}

extern "C" int bwa_gpu_set_stats(int enabled) { g_stats_enabled = enabled != 0; return 0; }

extern "C" int bwa_gpu_get_totals(bwa_gpu_totals_t *out)
{
	if (!out) return fail("bwa_gpu_get_totals: null");
	std::lock_guard<std::mutex> tg(g_tot_mu);
	*out = g_tot;
	return 0;
}

extern "C" void bwa_gpu_reset_totals(void)
{
	std::lock_guard<std::mutex> tg(g_tot_mu);
	g_tot = bwa_gpu_totals_t();
}

// ------------------------------------------------------------------ device pipeline for one resident chunk
static const int N_PASSES = 3;

// Pass 0: k_search, private arenas only (BWAGPU_T1_CAP records per thread).  Pass 1 (optimistic): k_search_warp, one warp per read, every
// entry in 1024-record chunks of a pool shared by the launch (BWAGPU_WARP_PASS=0: the pooled thread-per-read instantiation of k_search
// instead); the pool can run dry -- the reads it fails are retried from scratch in pass 2 (guaranteed): k_search with few enough threads
// that each can own opt->max_entries + 16 records, which the search can never exceed (bwtgap.c:140 stops it first).
typedef void (*search_fn)(const Batch);
static search_fn search_kernel(bool stats, bool pooled, bool stdmode)
{
	static const search_fn tab[8] = {k_search<false, false, false>, k_search<false, false, true>, k_search<false, true, false>,
	                                 k_search<false, true, true>,   k_search<true, false, false>, k_search<true, false, true>,
	                                 k_search<true, true, false>,   k_search<true, true, true>};
	return tab[(stats ? 4 : 0) | (pooled ? 2 : 0) | (stdmode ? 1 : 0)];
}
// bucket heads in shared memory: u16 per (bucket, thread) in pass 0 (private arena), u32 in the pooled passes
static size_t search_smem(bool pooled, uint32_t n_stacks)
{
	return SEARCH_SMEM(n_stacks, pooled ? sizeof(uint32_t) : sizeof(uint16_t)); // bucket heads + the hit list's two ends, context sectors
}
static bool is_stdmode(int mode) { return (mode & 0x01) && !(mode & 0x04) && !(mode & 0x10); }
// pass 1 = the warp-per-read kernel (search_warp.cuh) unless BWAGPU_WARP_PASS=0 (then: thread-per-read on the pool, as pass 0 but pooled)
static bool warp_pass_enabled()
{
	const char *e = getenv("BWAGPU_WARP_PASS");
	return !e || atoi(e) != 0;
}
static bool warp_stats_enabled()
{
	const char *e = getenv("BWAGPU_WARP_STATS");
	return e && atoi(e) != 0;
}
// reads per block of the warp pass: WK_WARPS with one warp per read, 1 when the whole block works on one read (BWAGPU_WARP_TEAM=1:
// rounds of 32 * WK_WARPS chains; same lanes per SM, a deep read finishes WK_WARPS times sooner)
// Measured (profiles/r1_ab_experiments.md): with ~10^5 deep reads the launch lasts as long as its deepest read and the team form is
// 13 % faster; with 4 x 10^5 it is throughput-bound and the warp form is 7 % faster (a team's round waits for the slowest of 128
// chains).  BWAGPU_WARP_TEAM=0/1 forces one form; unset: by the number of reads in the pass.
static bool warp_team_enabled(int n_jobs)
{
	const char *e = getenv("BWAGPU_WARP_TEAM");
	return e ? atoi(e) != 0 : n_jobs < 250000;
}
static int warp_reads_per_block(int n_jobs) { return warp_team_enabled(n_jobs) ? 1 : WK_WARPS; }
static search_fn warp_kernel(bool stdmode, int n_jobs = 0)
{
	if (warp_team_enabled(n_jobs)) return stdmode ? k_search_warp<true, false, WK_WARPS> : k_search_warp<false, false, WK_WARPS>;
	if (warp_stats_enabled()) return stdmode ? k_search_warp<true, true, 1> : k_search_warp<false, true, 1>;
	return stdmode ? k_search_warp<true, false, 1> : k_search_warp<false, false, 1>;
}
static const size_t WARP_SMEM = (size_t)WK_WARPS * WK_WORDS_PER_WARP * sizeof(uint32_t);

static int pass_setup(Ctx *c, int t, uint32_t n_stacks, uint32_t max_entries_opt, bool stdmode)
{
	Pass &T = c->pass_buf[t];
	// the shared pool: BWAGPU_POOL_MB (default 16384) per lane, never more than a quarter of what is free
	if (!c->xent.p) {
		size_t free_b = 0, total_b = 0;
		CK(cudaMemGetInfo(&free_b, &total_b));
		size_t want = (size_t)env_u32("BWAGPU_POOL_MB", 16384) << 20;
		if (want > free_b / 4) want = free_b / 4;
		size_t chunks = want / ((size_t)ARENA_CHUNK * 20);
		if (chunks < 64) chunks = 64;
		if (c->xent.reserve(chunks << ARENA_CHUNK_LOG) || c->xnxt.reserve(chunks << ARENA_CHUNK_LOG)) return 1;
		if (c->x_free_next.reserve(chunks + 1) || c->x_free_top.reserve(1)) return 1;
		c->x_chunks = (uint32_t)chunks;
	}
	const uint32_t need = max_entries_opt + 16; // records one search can hold at most
	if (t == 1 && warp_pass_enabled()) {
		// one warp per read, every entry in pool chunks: no private arenas, no chunk table
		int bps = 0;
		// both forms (warp per read / block per read) have the same block size, shared memory and register bound
		CK(cudaFuncSetAttribute(warp_kernel(stdmode, 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WARP_SMEM));
		CK(cudaFuncSetAttribute(warp_kernel(stdmode, 1 << 30), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WARP_SMEM));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, warp_kernel(stdmode, 1 << 30), WK_WARPS * 32, WARP_SMEM));
		if (bps < 1) bps = 1;
		bps = (int)std::min<uint32_t>((uint32_t)bps, env_u32("BWAGPU_WARP_BLOCKS_PER_SM", 64));
		T.slots_blocks = (uint32_t)(bps * c->n_sm);
		T.cap = 0; T.ctab_stride = 1;
		if (c->ctab.reserve(1)) return 1;
		return 0;
	}
	if (t <= 1) {
		// pass 0: private arenas only (k_search<.., false>); pass 1: same occupancy, arenas continue in the pool
		int bps = 0, bps_p = 0;
		for (int v = 0; v < 8; ++v)
			CK(cudaFuncSetAttribute(search_kernel(v & 4, v & 2, v & 1), cudaFuncAttributeMaxDynamicSharedMemorySize,
			                        (int)search_smem(v & 2, n_stacks)));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, search_kernel(g_stats_enabled, false, stdmode), 128, search_smem(false, n_stacks)));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_p, search_kernel(g_stats_enabled, true, stdmode), 128, search_smem(true, n_stacks)));
		if (t == 1) bps = bps_p;
		if (bps < 1) bps = 1;
		bps = (int)std::min<uint32_t>((uint32_t)bps, env_u32("BWAGPU_T1_BLOCKS_PER_SM", 64));
		T.slots_blocks = (uint32_t)(bps * c->n_sm);
		T.cap = std::min<uint32_t>(env_u32("BWAGPU_T1_CAP", 2048), 65535u); // pass-0 heads are u16
	} else {
		// guaranteed: threads x (chunks one search may need) <= pool
		const uint32_t per_thread = (need + ARENA_CHUNK - 1) >> ARENA_CHUNK_LOG;
		if (c->x_chunks < per_thread * 128u) { // a pool configured too small for one block of full-depth searches: grow it
			const size_t chunks = (size_t)per_thread * 128u;
			c->xent.release(); c->xnxt.release(); c->x_free_next.release();
			if (c->xent.reserve(chunks << ARENA_CHUNK_LOG) || c->xnxt.reserve(chunks << ARENA_CHUNK_LOG) || c->x_free_next.reserve(chunks + 1))
				return 1;
			c->x_chunks = (uint32_t)chunks;
		}
		uint32_t threads = c->x_chunks / std::max(1u, per_thread);
		if (threads > 4096) threads = 4096;
		T.slots_blocks = std::max(1u, threads / 128);
		T.cap = 1024;
	}
	if (T.cap > need) T.cap = need;
	const size_t slots = (size_t)T.slots_blocks * 128;
	const uint32_t stride = (need > T.cap ? (need - T.cap + ARENA_CHUNK - 1) >> ARENA_CHUNK_LOG : 0) + 1;
	if (T.ent.reserve(slots * ARENA_ALLOC(T.cap)) || T.nxt.reserve(slots * ARENA_ALLOC(T.cap))) return 1;
	if (c->ctab.reserve(slots * stride)) return 1;
	T.ctab_stride = stride;
#if !BWAGPU_SMEM_HEADS
	if (T.heads.reserve(slots * (n_stacks + 2))) return 1;
#endif
	return 0;
}

// Runs K2 + K3 (all passes) + ordered compaction on a chunk whose seq/meta are on the device.
// On return (stream synchronised): d_naln, d_maxent, d_outoff (exclusive scan), d_out
// hold the results, *total_aln the pool size.
static int run_chunk_device(Ctx *c, int n, size_t w_entries, const GapOpt &opt, uint32_t n_stacks, int64_t *total_aln,
                            bool device_compact)
{
	const bool stats = g_stats_enabled;
	if (c->d_w.reserve(w_entries + 1) || c->d_bid.reserve(w_entries + 1) || c->d_ctx16.reserve(w_entries + 32)) return 1;
	if (c->d_naln.reserve(n + 1) || c->d_maxent.reserve(n) || c->d_pooloff.reserve(n) || c->d_outoff.reserve(n + 1))
		return 1;
	if (c->d_jobs_a.reserve(n) || c->d_jobs_b.reserve(n)) return 1;
	if (c->d_counters.reserve(4) || c->h_counters.reserve(4) || c->d_stats.reserve(96)) return 1;
	// hit pool of the chunk (completion order): 16 hits per read on average to start with; grown and the
	// affected reads retried when a batch of short, repetitive reads needs more
	size_t pool_cap = std::max<size_t>((size_t)n * env_u32("BWAGPU_HITS_PER_READ", 16), 1u << 16);
	if (pool_cap > 0xfffffff0ull) pool_cap = 0xfffffff0ull;
	if (c->d_pool.reserve(pool_cap)) return 1;
	pool_cap = std::min<size_t>(c->d_pool.cap, 0xfffffff0ull);

	Batch B;
	B.ix[0] = c->ix[0]; B.ix[1] = c->ix[1];
	B.opt = opt;
	B.n_reads = n;
	B.seq = c->d_seq.p; B.meta = c->d_meta.p;
	B.w = c->d_w.p; B.bid = c->d_bid.p; B.ctx = c->d_ctx.p; B.ctx16 = c->d_ctx16.p;
	B.n_aln = c->d_naln.p; B.max_entries = c->d_maxent.p; B.pool_off = c->d_pooloff.p;
	B.pool = c->d_pool.p; B.pool_cap = (uint32_t)pool_cap;
	B.pool_count = (unsigned int *)(c->d_counters.p + 2);
	B.work_counter = c->d_counters.p; B.overflow_count = c->d_counters.p + 1;
	B.stats = c->d_stats.p;
	B.n_stacks = n_stacks;
	B.pop_cap = warp_pass_enabled() ? env_u32("BWAGPU_POP_CAP", 0) : 0; // without the warp pass a straggler is best left where it is

	CK(cudaMemsetAsync(c->d_counters.p, 0, 4 * sizeof(int), c->st));
	CK(cudaMemsetAsync(c->d_stats.p + 16, 0, 80 * sizeof(unsigned long long), c->st)); // k_search_warp diagnostics + the STATS histograms
	if (stats) {
		unsigned long long init[16] = {0};
		init[11] = init[12] = ~0ull;
		CK(cudaMemcpyAsync(c->d_stats.p, init, sizeof init, cudaMemcpyHostToDevice, c->st));
		CK(lane_sync(c));
	}

	// K2
	B.jobs = nullptr; B.n_jobs = n;
	CK(cudaEventRecord(c->ev[1], c->st));
	{
		const long long threads = 4ll * n;
		const int blocks = (int)((threads + 127) / 128);
		if (blocks > 0) {
			if (stats) k_width<true><<<blocks, 128, 0, c->st>>>(B);
			else k_width<false><<<blocks, 128, 0, c->st>>>(B);
			CK(cudaGetLastError());
			k_ctx16<<<(int)((32ll * n + 255) / 256), 256, 0, c->st>>>(B);
			CK(cudaGetLastError());
			c->stats.launches += 2;
		}
	}
	if (stats) {
		unsigned long long hs[16];
		CK(cudaMemcpyAsync(hs, c->d_stats.p, sizeof hs, cudaMemcpyDeviceToHost, c->st));
		CK(lane_sync(c));
		c->stats.occ_fetches_width += (int64_t)hs[0];
		c->stats.own_fetches_width += (int64_t)hs[1];
		TOT(g_tot.occ_fetches_width += (int64_t)hs[0]);
		unsigned long long init[16] = {0};
		init[11] = init[12] = ~0ull;
		CK(cudaMemcpyAsync(c->d_stats.p, init, sizeof init, cudaMemcpyHostToDevice, c->st));
		CK(lane_sync(c));
	}

	// job order for pass 0: longest-looking searches first (k_job_keys)
	const int32_t *jobs = nullptr;
	{
		const char *env = getenv("BWAGPU_SORT_JOBS");
		if (env && atoi(env) != 0 && n > 1) { // opt-in: helps 1-2M-read batches (shorter tail), not the 10M workload
			if (c->d_keys.reserve(n) || c->d_keys2.reserve(n) || c->d_ids.reserve(n) || c->d_order.reserve(n)) return 1;
			k_job_keys<<<(n + 255) / 256, 256, 0, c->st>>>(B, c->d_keys.p, c->d_ids.p);
			CK(cudaGetLastError());
			size_t tb = 0;
			CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, c->d_keys.p, c->d_keys2.p, c->d_ids.p, c->d_order.p, n, 0, 8, c->st));
			if (c->d_cubtmp.reserve(tb + 16)) return 1;
			CK(cub::DeviceRadixSort::SortPairs(c->d_cubtmp.p, tb, c->d_keys.p, c->d_keys2.p, c->d_ids.p, c->d_order.p, n, 0, 8, c->st));
			c->stats.launches += 3;
			jobs = c->d_order.p;
		}
	}
	CK(cudaEventRecord(c->ev[2], c->st)); // the sort is accounted to the width phase

	// K3, pass by pass
	int n_jobs = n;
	for (int pass = 0; n_jobs > 0; ++pass) {
		const int t = pass < N_PASSES ? pass : N_PASSES - 1;
		if (pass > 8) return fail("%d reads still unfinished after %d passes", n_jobs, pass);
		if (pass_setup(c, t, n_stacks, (uint32_t)opt.max_entries, is_stdmode(opt.mode))) return 1;
		Pass &T = c->pass_buf[t];
		B.ent = T.ent.p; B.nxt = T.nxt.p; B.heads = T.heads.p;
		B.cap = T.cap;
		B.xent = c->xent.p; B.xnxt = c->xnxt.p; B.ctab = c->ctab.p; B.ctab_stride = T.ctab_stride;
		B.x_chunks = c->x_chunks; B.x_next = (unsigned int *)(c->d_counters.p + 3);
		B.x_free_top = c->x_free_top.p; B.x_free_next = c->x_free_next.p;
		B.jobs = jobs; B.n_jobs = n_jobs;
		int32_t *ovf = (pass & 1) ? c->d_jobs_b.p : c->d_jobs_a.p;
		B.overflow_ids = ovf;
		CK(cudaMemsetAsync(c->d_counters.p, 0, 2 * sizeof(int), c->st)); // work + overflow counters
		CK(cudaMemsetAsync(c->d_counters.p + 3, 0, sizeof(int), c->st)); // overflow-pool bump counter
		{
			static const unsigned long long empty_stack = 0xffffffffull; // tag 0, head NIL
			CK(cudaMemcpyAsync(c->x_free_top.p, &empty_stack, sizeof empty_stack, cudaMemcpyHostToDevice, c->st));
		}
		CK(cudaEventRecord(c->ev[8 + 2 * t], c->st));
		if (pass > 0) { // pristine widths for the reads being retried
			const int wb = (int)((4ll * n_jobs + 127) / 128);
			if (stats) k_width<true><<<wb, 128, 0, c->st>>>(B);
			else k_width<false><<<wb, 128, 0, c->st>>>(B);
			CK(cudaGetLastError());
			if (t == 1 && warp_pass_enabled()) { // k_search_warp reads the 8-byte context words
				if (c->d_ctx.reserve(w_entries + 8)) return 1;
				B.ctx = c->d_ctx.p;
				k_ctx<<<(int)((32ll * n_jobs + 255) / 256), 256, 0, c->st>>>(B);
			} else k_ctx16<<<(int)((32ll * n_jobs + 255) / 256), 256, 0, c->st>>>(B);
			CK(cudaGetLastError());
			c->stats.launches += 2;
		}
		uint32_t blocks = T.slots_blocks;
		if (t == 1 && warp_pass_enabled()) {
			const int rpb = warp_reads_per_block(n_jobs);
			const uint32_t need = (uint32_t)((n_jobs + rpb - 1) / rpb);
			if (blocks > need) blocks = need;
			warp_kernel(is_stdmode(opt.mode), n_jobs)<<<blocks, WK_WARPS * 32, WARP_SMEM, c->st>>>(B);
		} else {
			const uint32_t need = (uint32_t)((n_jobs + 127) / 128);
			if (blocks > need) blocks = need;
			search_kernel(stats, t != 0, is_stdmode(opt.mode))<<<blocks, 128, search_smem(t != 0, n_stacks), c->st>>>(B);
		}
		CK(cudaGetLastError());
		c->stats.launches++;
		CK(cudaEventRecord(c->ev[9 + 2 * t], c->st));
		if (t == 1 && warp_pass_enabled() && warp_stats_enabled()) { // diagnostics of k_search_warp (search_warp.cuh)
			unsigned long long w[16];
			CK(cudaMemcpyAsync(w, c->d_stats.p + 16, sizeof w, cudaMemcpyDeviceToHost, c->st));
			CK(lane_sync(c));
			CK(cudaMemsetAsync(c->d_stats.p + 16, 0, sizeof w, c->st));
			const double r = (double)std::max<unsigned long long>(w[0], 1);
			fprintf(stderr, "[k_search_warp] reads %llu rounds %llu taken/round %.2f committed/round %.2f steps/lane %.2f longest chain/round %.2f | "
			                "clocks/round: take %.0f passA %.0f scan+commit+alloc %.0f passB %.0f hit+rest %.0f\n",
			        w[10], w[0], w[1] / r, w[2] / r, (double)w[3] / (double)std::max<unsigned long long>(w[1], 1), w[4] / r,
			        w[5] / r, w[6] / r, w[7] / r, w[8] / r, w[9] / r);
		}
		CK(cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
		CK(lane_sync(c));
		{
			float tms = 0;
			CK(cudaEventElapsedTime(&tms, c->ev[8 + 2 * t], c->ev[9 + 2 * t]));
			c->stats.ms_tier[t] += tms;
			TOT(g_tot.ms_search_pass[t] += tms);
		}
		const int n_over = c->h_counters.p[1];
		c->stats.x_chunks_used = std::max<int64_t>(c->stats.x_chunks_used, (int64_t)(unsigned int)c->h_counters.p[3]);
		const unsigned int pool_used = (unsigned int)c->h_counters.p[2];
		if (pool_used > pool_cap) {
			// The hit pool ran out: reads that could not place their hits were flagged like arena
			// overflows.  Grow the pool (keeping what finished reads wrote) and rewind the counter to the
			// old capacity: everything below it is either valid or an unreferenced reservation.
			size_t want = std::min<size_t>((size_t)pool_used * 2 + (1u << 20), 0xfffffff0ull);
			if (want <= pool_cap) return fail("hit pool cannot grow beyond %zu records in one chunk; lower BWAGPU_CHUNK", pool_cap);
			uint4 *bigger = nullptr;
			cudaError_t e = cudaMalloc((void **)&bigger, want * sizeof(uint4));
			if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) for the hit pool: %s", want * sizeof(uint4), cudaGetErrorString(e));
			CK(cudaMemcpyAsync(bigger, c->d_pool.p, pool_cap * sizeof(uint4), cudaMemcpyDeviceToDevice, c->st));
			const unsigned int rewind = (unsigned int)pool_cap;
			CK(cudaMemcpyAsync(c->d_counters.p + 2, &rewind, sizeof rewind, cudaMemcpyHostToDevice, c->st));
			CK(lane_sync(c));
			cudaFree(c->d_pool.p);
			c->d_pool.p = bigger; c->d_pool.cap = want;
			pool_cap = want;
			B.pool = bigger; B.pool_cap = (uint32_t)pool_cap;
		}
		if (pass == 0) c->stats.n_overflow_t2 += n_over;
		else if (pass == 1) c->stats.n_overflow_t3 += n_over;
		jobs = ovf;
		n_jobs = n_over;
	}
	CK(cudaEventRecord(c->ev[3], c->st));
	if (stats) {
		unsigned long long hs[16];
		CK(cudaMemcpyAsync(hs, c->d_stats.p, sizeof hs, cudaMemcpyDeviceToHost, c->st));
		CK(lane_sync(c));
		c->stats.occ_fetches_search += (int64_t)hs[0];
		c->stats.own_fetches_search += (int64_t)hs[1];
		TOT(g_tot.occ_fetches_search += (int64_t)hs[0]; g_tot.own_fetches_search += (int64_t)hs[1]);
		c->stats.n_pops += (int64_t)hs[2];
		c->stats.n_pushes += (int64_t)hs[3];
		c->stats.n_stored += (int64_t)hs[4];
		c->stats.n_pruned += (int64_t)hs[5];
		c->stats.n_expand += (int64_t)hs[6];
		c->stats.n_exact += (int64_t)hs[7];
		c->stats.n_derive += (int64_t)hs[8];
		c->stats.n_trips += (int64_t)hs[9];
		c->stats.ns_queue_empty = (int64_t)(hs[11] - hs[12]);
		c->stats.ns_kernel = (int64_t)(hs[13] - hs[12]);
		if (getenv("BWAGPU_PRINT_HIST")) { // diagnostics: where in the read the lookups happen, and how the work per read is distributed
			unsigned long long h[64];
			CK(cudaMemcpyAsync(h, c->d_stats.p + 32, sizeof h, cudaMemcpyDeviceToHost, c->st));
			CK(lane_sync(c));
			fprintf(stderr, "[k_search] lookups by depth (len - i; last = 31+):");
			for (int q = 0; q < 32; ++q) fprintf(stderr, " %llu", h[q]);
			fprintf(stderr, "\n[k_search] reads by log2(pops):");
			for (int q = 0; q < 32; ++q) fprintf(stderr, " %llu", h[32 + q]);
			fprintf(stderr, "\n");
		}
	}

	if (!device_compact) {
		// Chunked host path: the hits stay in completion order; the host reads pool_off[] and puts them in
		// read order while it unpacks.  No kernel after k_search: a small kernel of this lane would have to
		// wait for SM slots behind another lane's persistent k_search, and the lanes would fall into step.
		CK(cudaEventRecord(c->ev[4], c->st));
		CK(lane_sync(c));
		float ms0 = 0;
		CK(cudaEventElapsedTime(&ms0, c->ev[1], c->ev[2])); c->stats.ms_width += ms0; TOT(g_tot.ms_width += ms0);
		CK(cudaEventElapsedTime(&ms0, c->ev[2], c->ev[3])); c->stats.ms_search += ms0; TOT(g_tot.ms_search += ms0);
		*total_aln = (int64_t)std::min<size_t>((size_t)(unsigned int)c->h_counters.p[2], pool_cap); // hit-pool fill
		return 0;
	}
	// ordered compaction: exclusive scan of n_aln (n+1 items so that out_off[n] = total)
	CK(cudaMemsetAsync(c->d_naln.p + n, 0, sizeof(int32_t), c->st));
	size_t tmp_bytes = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->d_naln.p, (int32_t *)c->d_outoff.p, n + 1, c->st));
	if (c->d_cubtmp.reserve(tmp_bytes + 16)) return 1;
	CK(cub::DeviceScan::ExclusiveSum(c->d_cubtmp.p, tmp_bytes, c->d_naln.p, (int32_t *)c->d_outoff.p, n + 1, c->st));
	c->stats.launches += 2;
	uint32_t tot = 0;
	CK(cudaMemcpyAsync(&tot, c->d_outoff.p + n, 4, cudaMemcpyDeviceToHost, c->st));
	CK(lane_sync(c));
	if (c->d_out.reserve((size_t)tot + 1)) return 1;
	if (n > 0) {
		k_gather_aln<<<(n + 255) / 256, 256, 0, c->st>>>(n, c->d_naln.p, c->d_pooloff.p, c->d_outoff.p, c->d_pool.p, c->d_out.p);
		CK(cudaGetLastError());
		c->stats.launches++;
	}
	CK(cudaEventRecord(c->ev[4], c->st));
	CK(lane_sync(c));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats.ms_width += ms; TOT(g_tot.ms_width += ms);
	CK(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); c->stats.ms_search += ms; TOT(g_tot.ms_search += ms);
	CK(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); c->stats.ms_compact += ms;
	*total_aln = tot;
	return 0;
}

// ------------------------------------------------------------------ flat batch on one device
// The reads one DEVICE handles in a batch call, cut into chunks that the device's lanes pull from a
// shared queue.  The first chunks are deliberately unequal (1/L, 2/L, ... of a chunk for L lanes) so
// that the lanes fall out of step at once: while one lane packs or unpacks, another has its kernels
// on the device.
struct FlatJob {
	int n = 0;
	const uint8_t *bases = nullptr;      // flat API: sequencing orientation
	const int64_t *offs = nullptr;
	bwa_seq_t *seqs = nullptr;           // struct API instead
	const gap_opt_t *opt = nullptr;
	int32_t *n_aln = nullptr, *max_entries = nullptr;
	std::vector<std::pair<size_t, int>> chunks;   // (first read, reads)
	std::vector<std::vector<uint4>> pools;        // flat API: read-ordered alns of each chunk
	std::atomic<int> next{0}, failed{0};
	std::mutex err_mu;
	std::string err;
	FlatJob() {}
	FlatJob(const FlatJob &) = delete;
};

// Host-side helper threads of one lane: packing reads and handing results back are embarrassingly
// parallel over reads once the per-read offsets are known.
static const int MAX_HOST_THREADS = 64;
static int host_threads()
{
	static int n = 0;
	if (!n) {
		const unsigned hw = std::thread::hardware_concurrency();
		const uint32_t lanes = env_u32("BWAGPU_LANES", 3);
		n = (int)env_u32("BWAGPU_HOST_THREADS", std::max(1u, std::min(8u, (hw ? hw : 8u) / std::max(1u, lanes))));
		if (n > MAX_HOST_THREADS) n = MAX_HOST_THREADS; // the per-thread scratch of run_range is sized by it
	}
	return n;
}

template <typename F> static void parallel_for(int n, F fn)
{
	const int nt = std::min(host_threads(), std::max(1, n / 4096));
	if (nt <= 1) { fn(0, n, 0); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < nt; ++t) {
		const int lo = (int)((int64_t)n * t / nt), hi = (int)((int64_t)n * (t + 1) / nt);
		th.emplace_back([=]() { fn(lo, hi, t); });
	}
	for (auto &x : th) x.join();
}

static size_t chunk_reads()
{
	const char *env = getenv("BWAGPU_CHUNK");
	size_t c = env ? (size_t)atoll(env) : 0;
	return c > 0 ? c : (size_t)1 << 21; // 2 M reads: long kernels (the straggler tail of k_search is per launch), two lanes alternate
}

static int run_range(Ctx *c, FlatJob &J)
{
	CK(cudaSetDevice(c->dev));
	if (!c->has_index) return fail("no index loaded (bwa_gpu_load_index)");
	const gap_opt_t *opt = J.opt;
	const GapOpt gopt = to_gapopt(opt);
	MaxDiffTable mdt;
	for (;;) {
		const int ci = J.next.fetch_add(1);
		if (ci >= (int)J.chunks.size() || J.failed.load()) break;
		const size_t r0 = J.chunks[ci].first;
		const int n = J.chunks[ci].second;
		std::vector<uint4> &pool_out = J.pools[ci];
		auto t0 = std::chrono::steady_clock::now();
		// ---- marshal, pass 1 (serial, cheap): lengths -> offsets into the packed arrays
		std::vector<uint64_t> so_of(n + 1), wo_of(n + 1);
		int max_len = 0;
		{
			// per-read sizes in parallel (the bwa_seq_t array alone is 200 B per read), then one serial scan
			std::vector<int> bad_t(64, -1), max_t(64, 0);
			parallel_for(n, [&](int lo, int hi, int tid) {
				for (int i = lo; i < hi; ++i) {
					const int len = J.seqs ? (int)J.seqs[r0 + i].len : (int)(J.offs[r0 + i + 1] - J.offs[r0 + i]);
					if (len < 0 || len > 32766) { bad_t[tid] = len; return; }
					so_of[i + 1] = (uint64_t)len;
					wo_of[i + 1] = width_entries(len, opt->seed_len);
					if (len > max_t[tid]) max_t[tid] = len;
				}
			});
			for (int t = 0; t < 64; ++t) {
				if (bad_t[t] != -1) return fail("read length %d not supported (max 32766)", bad_t[t]);
				if (max_t[t] > max_len) max_len = max_t[t];
			}
			so_of[0] = wo_of[0] = 0;
			for (int i = 0; i < n; ++i) { so_of[i + 1] += so_of[i]; wo_of[i + 1] += wo_of[i]; }
		}
		const uint64_t n_bases = so_of[n], wo = wo_of[n];
		if (n_bases >= 0xffffffffull) return fail("chunk holds %llu bases; lower BWAGPU_CHUNK", (unsigned long long)n_bases);
		if (wo >= 0xffffffffull) return fail("width arena exceeds 2^32 entries; lower BWAGPU_CHUNK");
		if (c->h_seq.reserve(n_bases + 1) || c->h_meta.reserve(n)) return 1;
		for (int len = 0; len <= max_len; ++len) (void)mdt.get(len, opt); // fill the table: read-only below
		// ---- pass 2 (parallel over reads): pack bases, per-read meta
		uint32_t n_stacks = 1;
		{
			std::vector<uint32_t> ns_t(64, 1);
			std::vector<int> rc_t(64, 0);
			std::vector<std::string> err_t(64);
			parallel_for(n, [&](int lo, int hi, int tid) {
				MaxDiffTable local = mdt; // private copy: get() may not mutate shared state
				for (int i = lo; i < hi; ++i) {
					int len;
					uint32_t n_amb;
					uint8_t *dst = c->h_seq.p + so_of[i];
					if (J.seqs) {
						const bwa_seq_t *p = J.seqs + r0 + i;
						len = (int)p->len;
						n_amb = pack_seq_pair(dst, p->seq, p->rseq, len);
					} else {
						const uint8_t *src = J.bases + J.offs[r0 + i];
						len = (int)(J.offs[r0 + i + 1] - J.offs[r0 + i]);
						n_amb = pack_read(dst, src, len);
					}
					uint64_t we = 0;
					if (fill_meta(len, so_of[i], wo_of[i], opt, local, c->h_meta.p[i], we, ns_t[tid])) { rc_t[tid] = 1; err_t[tid] = t_err; return; }
					c->h_meta.p[i].n_amb = n_amb;
				}
			});
			for (int t = 0; t < 64; ++t) {
				if (rc_t[t]) return fail("%s", err_t[t].c_str());
				if (ns_t[t] > n_stacks) n_stacks = ns_t[t];
			}
		}
		auto t1 = std::chrono::steady_clock::now();
		c->stats.ms_host_marshal += std::chrono::duration<double, std::milli>(t1 - t0).count();
		// ---- H2D
		if (c->d_seq.reserve(n_bases + 1) || c->d_meta.reserve(n)) return 1;
		CK(cudaEventRecord(c->ev[0], c->st));
		CK(cudaMemcpyAsync(c->d_seq.p, c->h_seq.p, n_bases, cudaMemcpyHostToDevice, c->st));
		CK(cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, (size_t)n * sizeof(ReadMeta), cudaMemcpyHostToDevice, c->st));
		int64_t tot = 0; // hit-pool fill: records to copy back (completion order)
		if (run_chunk_device(c, n, (size_t)wo, gopt, n_stacks, &tot, false)) return 1;
		// ---- D2H
		if (c->h_naln.reserve(n) || c->h_maxent.reserve(n) || c->h_pooloff.reserve(n) || c->h_out.reserve((size_t)tot + 1)) return 1;
		CK(cudaEventRecord(c->ev[5], c->st));
		CK(cudaMemcpyAsync(c->h_naln.p, c->d_naln.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
		CK(cudaMemcpyAsync(c->h_maxent.p, c->d_maxent.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
		CK(cudaMemcpyAsync(c->h_pooloff.p, c->d_pooloff.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
		if (tot) CK(cudaMemcpyAsync(c->h_out.p, c->d_pool.p, (size_t)tot * 16, cudaMemcpyDeviceToHost, c->st));
		CK(cudaEventRecord(c->ev[6], c->st));
		CK(lane_sync(c));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats.ms_h2d += ms;
		CK(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6])); c->stats.ms_d2h += ms;
		CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[6])); c->stats.ms_total_device += ms;
		auto t2 = std::chrono::steady_clock::now();
		if (getenv("BWAGPU_TRACE")) {
			static const auto T0 = std::chrono::steady_clock::now();
			auto ms = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(t - T0).count(); };
			fprintf(stderr, "[trace] lane %d.%d chunk %d (%d reads): pack %.0f-%.0f device %.0f-%.0f\n", c->dev, c->lane, ci, n, ms(t0), ms(t1), ms(t1), ms(t2));
		}
		memcpy(J.n_aln + r0, c->h_naln.p, (size_t)n * 4);
		memcpy(J.max_entries + r0, c->h_maxent.p, (size_t)n * 4);
		if (J.seqs) { // struct API: hand the hits back the way bwa_cal_sa_reg_gap does
			std::vector<int> rc_t(64, 0);
			parallel_for(n, [&](int lo, int hi, int tid) {
			for (int i = lo; i < hi; ++i) {
				bwa_seq_t *p = J.seqs + r0 + i;
				const int na = c->h_naln.p[i];
				const uint4 *src = c->h_out.p + c->h_pooloff.p[i];
				// bwtaln.c:113 resets these before the search
				p->sa = 0; p->type = 0 /* BWA_TYPE_NO_MATCH */; p->c1 = p->c2 = 0;
				p->n_aln = na;
				if (p->len == 0) { p->aln = 0; continue; } // bwtaln.c:134
				// bwt_match_gap always returns a calloc'd array of m_aln >= 4 records, grown by
				// doubling (bwtgap.c:114-115,188-191); consumers free() it.
				int m_aln = 4;
				while (m_aln < na) m_aln <<= 1;
				p->aln = (bwt_aln1_t *)calloc((size_t)m_aln, sizeof(bwt_aln1_t));
				if (!p->aln) { rc_t[tid] = 1; return; }
				if (na) memcpy(p->aln, src, (size_t)na * 16);
				// untouched when the search never ran (bwtgap.c:120-123)
				if (c->h_maxent.p[i] > 0) p->max_entries = c->h_maxent.p[i];
			}
			});
			for (int t = 0; t < 64; ++t) if (rc_t[t]) return fail("calloc failed");
		} else { // flat API: read order
			std::vector<int64_t> aoff(n + 1);
			aoff[0] = 0;
			for (int i = 0; i < n; ++i) aoff[i + 1] = aoff[i] + c->h_naln.p[i];
			pool_out.resize((size_t)aoff[n]);
			parallel_for(n, [&](int lo, int hi, int) {
				for (int i = lo; i < hi; ++i)
					if (c->h_naln.p[i]) memcpy(pool_out.data() + aoff[i], c->h_out.p + c->h_pooloff.p[i], (size_t)c->h_naln.p[i] * 16);
			});
		}
		auto t3 = std::chrono::steady_clock::now();
		if (getenv("BWAGPU_TRACE")) fprintf(stderr, "[trace] lane %d.%d chunk %d unpack %.0f ms\n", c->dev, c->lane, ci, std::chrono::duration<double, std::milli>(t3 - t2).count());
		c->stats.ms_host_marshal += std::chrono::duration<double, std::milli>(t3 - t2).count();
		c->stats.n_reads += n;
		int64_t na_sum = 0;
		for (int i = 0; i < n; ++i) na_sum += c->h_naln.p[i];
		c->stats.n_aln += na_sum;
		TOT(g_tot.reads += n; g_tot.alns += na_sum; g_tot.h2d_bytes += (int64_t)n_bases + (int64_t)n * (int64_t)sizeof(ReadMeta);
		    g_tot.d2h_bytes += (int64_t)n * 12 + tot * 16);
	}
	return 0;
}

// splits [0,n) over the devices, cuts each device's range into chunks, one host thread per lane
// a search call's lanes: group g of every device (lane l of a device with L lanes belongs to group l * G / L)
struct LaneGroup {
	int g = -1;
	LaneGroup()
	{
		std::unique_lock<std::mutex> lk(g_grp_mu);
		for (;;) {
			for (int k = 0; k < g_groups; ++k)
				if (!g_grp_busy[(size_t)k]) { g = k; break; }
			if (g >= 0) break;
			g_grp_cv.wait(lk);
		}
		g_grp_busy[(size_t)g] = 1;
	}
	~LaneGroup()
	{
		{ std::lock_guard<std::mutex> lk(g_grp_mu); g_grp_busy[(size_t)g] = 0; }
		g_grp_cv.notify_one();
	}
};

static bool lane_in_group(const Ctx *c, int g)
{
	int lanes = 0;
	for (const Ctx *o : g_ctx) if (o->dev == c->dev) ++lanes;
	return c->lane * g_groups / lanes == g;
}

static int run_all(int n, const uint8_t *bases, const int64_t *offs, bwa_seq_t *seqs, const gap_opt_t *opt,
                   int32_t *n_aln, int32_t *max_entries, std::vector<std::unique_ptr<FlatJob>> &jobs, int group)
{
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called (no CPU fallback)");
	std::vector<Ctx *> owners, mine;
	for (Ctx *c : g_ctx) if (!c->owner) owners.push_back(c);
	for (Ctx *c : g_ctx) if (lane_in_group(c, group)) mine.push_back(c);
	const int nd = (int)owners.size();
	jobs.clear();
	for (Ctx *c : mine) { c->stats = bwa_gpu_stats_t(); c->stats.n_devices = nd; }
	const size_t CH = chunk_reads();
	std::vector<std::thread> th;
	for (int d = 0; d < nd; ++d) {
		const int64_t lo = (int64_t)n * d / nd, hi = (int64_t)n * (d + 1) / nd;
		jobs.emplace_back(new FlatJob());
		FlatJob &J = *jobs.back();
		J.n = (int)(hi - lo);
		J.opt = opt;
		J.n_aln = n_aln + lo; J.max_entries = max_entries + lo;
		if (seqs) J.seqs = seqs + lo;
		else { J.bases = bases; J.offs = offs + lo; }
		std::vector<Ctx *> lanes;
		for (Ctx *c : mine) if (c == owners[d] || c->owner == owners[d]) lanes.push_back(c);
		const size_t L = lanes.size();
		// ramp up (the device starts working after a short first pack) and ramp down (a short last unpack)
		const size_t floor_sz = std::min<size_t>(CH, 65536);
		std::vector<size_t> tail_sizes; // sizes of the last chunks, outermost last
		for (size_t k = 0, left = (size_t)J.n / 2; k < L && left > 0; ++k) {
			const size_t m = std::min(left, std::max<size_t>(CH * (k + 1) / (L + 1), floor_sz));
			tail_sizes.push_back(m); left -= m;
		}
		size_t tail_total = 0;
		for (size_t m : tail_sizes) tail_total += m;
		for (size_t r0 = 0, k = 0; r0 < (size_t)J.n - tail_total; ++k) {
			const size_t want = k < L ? std::max<size_t>(CH * (k + 1) / (L + 1), floor_sz) : CH;
			const size_t m = std::min(want, (size_t)J.n - tail_total - r0);
			J.chunks.emplace_back(r0, (int)m);
			r0 += m;
		}
		for (size_t k = tail_sizes.size(), r0 = (size_t)J.n - tail_total; k-- > 0;) {
			J.chunks.emplace_back(r0, (int)tail_sizes[k]);
			r0 += tail_sizes[k];
		}
		J.pools.resize(J.chunks.size());
		for (Ctx *c : lanes)
			th.emplace_back([c, &J]() {
				if (run_range(c, J)) {
					std::lock_guard<std::mutex> g(J.err_mu);
					if (!J.failed.exchange(1)) J.err = t_err;
				}
			});
	}
	for (auto &t : th) t.join();
	{
		int64_t l = 0;
		for (Ctx *c : mine) l += c->stats.launches;
		TOT(g_tot.launches += l);
	}
	for (int d = 0; d < nd; ++d)
		if (jobs[d]->failed.load()) return fail("device %d: %s", owners[d]->dev, jobs[d]->err.c_str());
	return 0;
}

extern "C" int bwa_gpu_aln_flat(int n, const uint8_t *bases, const int64_t *offs, const gap_opt_t *opt, int32_t *n_aln,
                                int32_t *max_entries, int64_t *aln_off, const bwt_aln1_t **aln_pool)
{
	LIFE_SHARED;
	std::lock_guard<std::mutex> fg(g_flat_mu); // the result pool of the flat call is one per process
	if (n < 0 || !opt || !n_aln || !max_entries || !aln_off || !aln_pool) return fail("bwa_gpu_aln_flat: bad argument");
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called (no CPU fallback)");
	std::vector<std::unique_ptr<FlatJob>> jobs;
	LaneGroup grp;
	if (run_all(n, bases, offs, nullptr, opt, n_aln, max_entries, jobs, grp.g)) return 1;
	int64_t acc = 0;
	for (int i = 0; i < n; ++i) { aln_off[i] = acc; acc += n_aln[i]; }
	aln_off[n] = acc;
	g_flat_pool.clear();
	g_flat_pool.reserve((size_t)acc);
	for (auto &J : jobs) // devices in read order, chunks in read order
		for (auto &P : J->pools) g_flat_pool.insert(g_flat_pool.end(), P.begin(), P.end());
	if ((int64_t)g_flat_pool.size() != acc) return fail("internal: pool size %zu != %lld", g_flat_pool.size(), (long long)acc);
	*aln_pool = (const bwt_aln1_t *)g_flat_pool.data();
	return 0;
}

extern "C" int bwa_gpu_cal_sa_reads_gap(int n_seqs, bwa_seq_t *seqs, const gap_opt_t *opt)
{
	LIFE_SHARED;
	if (n_seqs < 0 || !opt || (n_seqs && !seqs)) return fail("bwa_gpu_cal_sa_reads_gap: bad argument");
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called (no CPU fallback)");
	// what bwa_cal_sa_reg_gap does to every read up front (bwtaln.c:113): a caller that cleans up after an error must not find
	// stale pointers
	for (int i = 0; i < n_seqs; ++i) { seqs[i].aln = 0; seqs[i].n_aln = 0; }
	std::vector<int32_t> n_aln(n_seqs), max_entries(n_seqs);
	std::vector<std::unique_ptr<FlatJob>> jobs;
	LaneGroup grp;
	if (run_all(n_seqs, nullptr, nullptr, seqs, opt, n_aln.data(), max_entries.data(), jobs, grp.g)) {
		for (int i = 0; i < n_seqs; ++i) { free(seqs[i].aln); seqs[i].aln = 0; seqs[i].n_aln = 0; } // hand nothing half-done back
		return 1;
	}
	return 0;
}

extern "C" void bwa_gpu_free_alns(int n_seqs, bwa_seq_t *seqs)
{
	for (int i = 0; i < n_seqs; ++i) { free(seqs[i].aln); seqs[i].aln = 0; seqs[i].n_aln = 0; }
}

extern "C" int bwa_gpu_get_stats(bwa_gpu_stats_t *out)
{
	if (!out) return fail("bwa_gpu_get_stats: null");
	LIFE_SHARED;
	bwa_gpu_stats_t s = bwa_gpu_stats_t();
	for (Ctx *c : g_ctx) {
		const bwa_gpu_stats_t &t = c->stats;
		// devices run concurrently: times are the max over devices, counters the sum
		s.ms_h2d = std::max(s.ms_h2d, t.ms_h2d); s.ms_width = std::max(s.ms_width, t.ms_width);
		s.ms_search = std::max(s.ms_search, t.ms_search); s.ms_compact = std::max(s.ms_compact, t.ms_compact);
		s.ms_d2h = std::max(s.ms_d2h, t.ms_d2h); s.ms_total_device = std::max(s.ms_total_device, t.ms_total_device);
		s.ms_host_marshal = std::max(s.ms_host_marshal, t.ms_host_marshal);
		s.n_reads += t.n_reads; s.n_aln += t.n_aln; s.n_overflow_t2 += t.n_overflow_t2; s.n_overflow_t3 += t.n_overflow_t3;
		s.occ_fetches_width += t.occ_fetches_width; s.occ_fetches_search += t.occ_fetches_search;
		s.own_fetches_width += t.own_fetches_width; s.own_fetches_search += t.own_fetches_search;
		s.n_pops += t.n_pops; s.n_pushes += t.n_pushes; s.n_stored += t.n_stored; s.n_pruned += t.n_pruned; s.n_expand += t.n_expand; s.n_exact += t.n_exact; s.n_derive += t.n_derive; s.ms_sw_kernel = std::max(s.ms_sw_kernel, t.ms_sw_kernel); s.n_trips += t.n_trips; s.ns_queue_empty = std::max(s.ns_queue_empty, t.ns_queue_empty); s.ns_kernel = std::max(s.ns_kernel, t.ns_kernel); s.launches += t.launches;
		for (int q = 0; q < 4; ++q) s.ms_tier[q] = std::max(s.ms_tier[q], t.ms_tier[q]);
		s.x_chunks_used = std::max(s.x_chunks_used, t.x_chunks_used);
	}
	for (Ctx *c : g_svc) s.ms_sw_kernel = std::max(s.ms_sw_kernel, c->stats.ms_sw_kernel);
	s.n_devices = 0;
	for (Ctx *c : g_ctx) if (!c->owner) ++s.n_devices;
	*out = s;
	return 0;
}

// ------------------------------------------------------------------ measurement: random-sector gather ceiling
extern "C" int bwa_gpu_probe_random_sectors(int64_t buffer_bytes, int chains, int steps, double *gb_per_s)
{
	LIFE_EXCLUSIVE; // measurement: alone on the device
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called (no CPU fallback)");
	if (!gb_per_s || buffer_bytes < (1 << 20) || steps < 1) return fail("bwa_gpu_probe_random_sectors: bad arguments");
	Ctx *c = g_ctx[0];
	CK(cudaSetDevice(c->dev));
	const uint64_t n_blk64 = (uint64_t)buffer_bytes / 32;
	if (n_blk64 > 0xffffffffull) return fail("probe buffer too large");
	const uint32_t n_blk = (uint32_t)n_blk64;
	uint4 *buf = nullptr;
	uint32_t *sink = nullptr;
	const int blocks = c->n_sm * 8, threads = 256;
	CK(cudaMalloc((void **)&buf, (size_t)n_blk * 32));
	CK(cudaMalloc((void **)&sink, (size_t)blocks * threads * 4));
	CK(cudaMemsetAsync(buf, 0x5a, (size_t)n_blk * 32, c->st));
	float best = 0;
	for (int rep = 0; rep < 4; ++rep) { // rep 0 warms up
		CK(cudaEventRecord(c->ev[5], c->st));
		if (chains <= 1) k_probe_gather<1><<<blocks, threads, 0, c->st>>>(buf, n_blk, steps, 17u * rep, sink);
		else if (chains == 2) k_probe_gather<2><<<blocks, threads, 0, c->st>>>(buf, n_blk, steps, 17u * rep, sink);
		else if (chains <= 4) k_probe_gather<4><<<blocks, threads, 0, c->st>>>(buf, n_blk, steps, 17u * rep, sink);
		else k_probe_gather<8><<<blocks, threads, 0, c->st>>>(buf, n_blk, steps, 17u * rep, sink);
		CK(cudaGetLastError());
		CK(cudaEventRecord(c->ev[6], c->st));
		CK(lane_sync(c));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]));
		const int ch = chains <= 1 ? 1 : chains == 2 ? 2 : chains <= 4 ? 4 : 8;
		const float gbs = (float)((double)blocks * threads * ch * steps * 32.0 / (ms * 1e-3) / 1e9);
		if (rep > 0 && gbs > best) best = gbs;
	}
	cudaFree(buf); cudaFree(sink);
	*gb_per_s = best;
	return 0;
}

// ------------------------------------------------------------------ resident batch (kernel-only timing)
extern "C" int bwa_gpu_resident_stage(int n, const uint8_t *bases, const int64_t *offs, const gap_opt_t *opt)
{
	LIFE_EXCLUSIVE; // measurement: alone on the device
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called (no CPU fallback)");
	Ctx *c = g_ctx[0];
	CK(cudaSetDevice(c->dev));
	if (!c->has_index) return fail("no index loaded");
	MaxDiffTable mdt;
	const uint64_t n_bases = (uint64_t)(offs[n] - offs[0]);
	if (n_bases >= 0xffffffffull) return fail("resident batch too large (%llu bases)", (unsigned long long)n_bases);
	if (c->h_seq.reserve(n_bases + 1) || c->h_meta.reserve(n)) return 1;
	uint64_t so = 0, wo = 0;
	uint32_t n_stacks = 1;
	for (int i = 0; i < n; ++i) {
		const uint8_t *src = bases + offs[i];
		const int len = (int)(offs[i + 1] - offs[i]);
		uint8_t *dst = c->h_seq.p + so;
		const uint32_t n_amb = pack_read(dst, src, len);
		uint64_t we = 0;
		if (fill_meta(len, so, wo, opt, mdt, c->h_meta.p[i], we, n_stacks)) return 1;
		c->h_meta.p[i].n_amb = n_amb;
		so += (uint64_t)len; wo += we;
		if (wo >= 0xffffffffull) return fail("resident batch too large (width arena)");
	}
	if (c->d_seq.reserve(n_bases + 1) || c->d_meta.reserve(n)) return 1;
	CK(cudaMemcpyAsync(c->d_seq.p, c->h_seq.p, n_bases, cudaMemcpyHostToDevice, c->st));
	CK(cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, (size_t)n * sizeof(ReadMeta), cudaMemcpyHostToDevice, c->st));
	CK(lane_sync(c));
	c->res_n = n; c->res_w_entries = (size_t)wo; c->res_opt = to_gapopt(opt); c->res_nstacks = n_stacks;
	c->res_valid = true;
	return 0;
}

extern "C" int bwa_gpu_resident_run(double *ms)
{
	LIFE_EXCLUSIVE; // measurement: alone on the device
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called");
	Ctx *c = g_ctx[0];
	if (!c->res_valid) return fail("bwa_gpu_resident_run: nothing staged");
	CK(cudaSetDevice(c->dev));
	c->stats = bwa_gpu_stats_t();
	c->stats.n_devices = 1;
	CK(cudaEventRecord(c->ev[0], c->st));
	int64_t tot = 0;
	if (run_chunk_device(c, c->res_n, c->res_w_entries, c->res_opt, c->res_nstacks, &tot, true)) return 1;
	CK(cudaEventRecord(c->ev[7], c->st));
	CK(lane_sync(c));
	float t = 0;
	CK(cudaEventElapsedTime(&t, c->ev[0], c->ev[7]));
	c->stats.ms_total_device = t;
	c->stats.n_reads = c->res_n;
	c->stats.n_aln = tot;
	TOT(g_tot.launches += c->stats.launches; g_tot.reads += c->res_n; g_tot.alns += tot);
	c->res_total_aln = tot;
	if (ms) *ms = t;
	return 0;
}

extern "C" int bwa_gpu_resident_fetch(int32_t *n_aln, int32_t *max_entries, int64_t *aln_off, const bwt_aln1_t **aln_pool)
{
	LIFE_EXCLUSIVE; // measurement: alone on the device
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called");
	Ctx *c = g_ctx[0];
	if (!c->res_valid) return fail("bwa_gpu_resident_fetch: nothing staged");
	CK(cudaSetDevice(c->dev));
	const int n = c->res_n;
	const int64_t tot = c->res_total_aln;
	g_flat_pool.resize((size_t)tot);
	CK(cudaMemcpy(n_aln, c->d_naln.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(max_entries, c->d_maxent.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
	if (tot) CK(cudaMemcpy(g_flat_pool.data(), c->d_out.p, (size_t)tot * 16, cudaMemcpyDeviceToHost));
	int64_t acc = 0;
	for (int i = 0; i < n; ++i) { aln_off[i] = acc; acc += n_aln[i]; }
	aln_off[n] = acc;
	*aln_pool = (const bwt_aln1_t *)g_flat_pool.data();
	return 0;
}

// ------------------------------------------------------------------ K4
extern "C" int bwa_gpu_cal_pac_pos(int64_t n, const bwtint_t *sa_idx, const uint8_t *which, bwtint_t *out_sa)
{
	LIFE_SHARED;
	std::lock_guard<std::mutex> sg(g_svc_mu);
	if (g_ctx.empty()) return fail("bwa_gpu_init has not been called (no CPU fallback)");
	if (n < 0 || (n && (!sa_idx || !which || !out_sa))) return fail("bwa_gpu_cal_pac_pos: bad argument");
	const int nd = (int)g_svc.size(); // the queries are split over the devices' service lanes
	std::vector<int> rc(nd, 0);
	std::vector<std::string> errs(nd);
	auto work = [&](int d) -> int {
		Ctx *c = g_svc[d];
		CK(cudaSetDevice(c->dev));
		if (!c->has_index || !c->has_sa) return fail("no suffix array loaded (bwt->sa was NULL in bwa_gpu_load_index)");
		const int64_t lo = n * d / nd, hi = n * (d + 1) / nd;
		const int64_t CH = 1ll << 26;
		for (int64_t r0 = lo; r0 < hi; r0 += CH) {
			const int64_t m = std::min(CH, hi - r0);
			if (c->d_q.reserve(m) || c->d_qo.reserve(m) || c->d_which.reserve(m)) return 1;
			for (int64_t i = 0; i < m; ++i) {
				const DevIndex &ix = c->ix[which[r0 + i] ? 0 : 1];
				if (sa_idx[r0 + i] > ix.seq_len) return fail("sa_idx[%lld] = %u out of range", (long long)(r0 + i), sa_idx[r0 + i]);
			}
			CK(cudaMemcpyAsync(c->d_q.p, sa_idx + r0, (size_t)m * 4, cudaMemcpyHostToDevice, c->st));
			CK(cudaMemcpyAsync(c->d_which.p, which + r0, (size_t)m, cudaMemcpyHostToDevice, c->st));
			IndexPair P; P.ix[0] = c->ix[0]; P.ix[1] = c->ix[1];
			CK(cudaEventRecord(c->ev[14], c->st));
			k_sa<<<(unsigned)((m + 255) / 256), 256, 0, c->st>>>(P, m, c->d_q.p, c->d_which.p, c->d_qo.p);
			CK(cudaGetLastError());
			CK(cudaEventRecord(c->ev[15], c->st));
			CK(cudaMemcpyAsync(out_sa + r0, c->d_qo.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->st));
			CK(lane_sync(c));
			float ms = 0;
			CK(cudaEventElapsedTime(&ms, c->ev[14], c->ev[15]));
			TOT(g_tot.ms_sa += ms; g_tot.launches += 1; g_tot.sa_queries += m; g_tot.h2d_bytes += m * 5; g_tot.d2h_bytes += m * 4);
		}
		return 0;
	};
	if (nd == 1) { if (work(0)) return 1; return 0; }
	std::vector<std::thread> th;
	for (int d = 0; d < nd; ++d) th.emplace_back([&, d]() { rc[d] = work(d); if (rc[d]) errs[d] = t_err; });
	for (auto &t : th) t.join();
	for (int d = 0; d < nd; ++d) if (rc[d]) return fail("device %d: %s", g_svc[d]->dev, errs[d].c_str());
	return 0;
}

// ------------------------------------------------------------------ K5
// result pools of the last K6 calls: one per entry-point family, so that a host thread reading the CIGARs of its
// bwa_gpu_mate_sw_path call is not disturbed by another thread's bwa_gpu_global_align* call (and vice versa)
static std::vector<uint16_t> g_cigar_pool_sw, g_cigar_pool_ga;

static int sw_entry(int n, const bwa_gpu_sw_job_t *jobs, int mode, int gap_end, int band, bwa_gpu_sw_res_t *res,
                    bwa_gpu_path_res_t *pres, const bwa_cigar_t **cigar_pool, const char *who)
{
	LIFE_SHARED;
	std::lock_guard<std::mutex> sg(g_svc_mu);
	if (g_ctx.empty()) return fail("%s: bwa_gpu_init has not been called (no CPU fallback)", who);
	if (n < 0 || (n && (!jobs || (mode == 0 && !res) || (mode != 0 && (!pres || !cigar_pool))))) return fail("%s: bad argument", who);
	Ctx *c = g_svc[0];
	CK(cudaSetDevice(c->dev));
	if (!c->has_pac) return fail("%s: no packed reference loaded (pac was NULL in bwa_gpu_load_index)", who);
	double ms[2] = {0, 0};
	SwCounts cnt;
	std::vector<bwa_gpu_sw_res_t> tmp;
	if (mode == 1 && !res) { tmp.resize(n); res = tmp.data(); }
	std::vector<uint16_t> &pool = mode == 1 ? g_cigar_pool_sw : g_cigar_pool_ga;
	const int rc = sw_batch(c->sw, c->st, c->pac.p, c->l_pac, n, jobs, mode, gap_end, band, res, pres, mode ? &pool : nullptr, fail, ms, &cnt);
	c->stats.ms_sw_kernel = ms[0] + ms[1];
	TOT(g_tot.ms_sw += ms[0]; g_tot.ms_global += ms[1]; g_tot.launches += cnt.launches; g_tot.sw_cells_fwd += cnt.cells_fwd;
	    g_tot.h2d_bytes += cnt.h2d; g_tot.d2h_bytes += cnt.d2h; if (mode == 2) g_tot.ga_jobs += n; else g_tot.sw_jobs += n);
	if (mode && cigar_pool) *cigar_pool = pool.data();
	return rc;
}

extern "C" int bwa_gpu_mate_sw(int n, const bwa_gpu_sw_job_t *jobs, bwa_gpu_sw_res_t *res)
{
	return sw_entry(n, jobs, 0, 0, 0, res, nullptr, nullptr, "bwa_gpu_mate_sw");
}

extern "C" int bwa_gpu_mate_sw_path(int n, const bwa_gpu_sw_job_t *jobs, bwa_gpu_path_res_t *res, const bwa_cigar_t **cigar_pool)
{
	return sw_entry(n, jobs, 1, -1, 50, nullptr, res, cigar_pool, "bwa_gpu_mate_sw_path");
}

extern "C" int bwa_gpu_global_align(int n, const bwa_gpu_sw_job_t *jobs, int gap_end, int band, bwa_gpu_path_res_t *res,
                                    const bwa_cigar_t **cigar_pool)
{
	if (band < 1) return fail("bwa_gpu_global_align: band must be >= 1");
	return sw_entry(n, jobs, 2, gap_end, band, nullptr, res, cigar_pool, "bwa_gpu_global_align");
}

extern "C" int bwa_gpu_global_align_seqs(int n, const bwa_gpu_ga_job_t *jobs, int gap_end, int band, bwa_gpu_path_res_t *res,
                                         const bwa_cigar_t **cigar_pool)
{
	LIFE_SHARED;
	std::lock_guard<std::mutex> sg(g_svc_mu);
	if (g_ctx.empty()) return fail("bwa_gpu_global_align_seqs: bwa_gpu_init has not been called (no CPU fallback)");
	if (band < 1) return fail("bwa_gpu_global_align_seqs: band must be >= 1");
	if (n < 0 || (n && (!jobs || !res || !cigar_pool))) return fail("bwa_gpu_global_align_seqs: bad argument");
	Ctx *c = g_svc[0];
	CK(cudaSetDevice(c->dev));
	double ms = 0;
	SwCounts cnt;
	const int rc = ga_seqs_batch(c->sw, c->st, n, jobs, gap_end, band, res, &g_cigar_pool_ga, fail, &ms, &cnt);
	TOT(g_tot.ms_global += ms; g_tot.launches += cnt.launches; g_tot.ga_jobs += n; g_tot.h2d_bytes += cnt.h2d; g_tot.d2h_bytes += cnt.d2h);
	*cigar_pool = g_cigar_pool_ga.data();
	return rc;
}
