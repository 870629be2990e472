/* bwa_gpu_batch.c -- the BATCHED drop-in: `bwa bam2bam -t 1` with its hot path on the GPU, one device
 * call per phase and batch instead of one per record.
 *
 * SURVEY.md §8(f) rank 1 / INTEGRATION.md: the reference's sequential driver (bam2bam.c:1143-1219)
 * handles one record at a time -- read_bam_pair -> pair_aln -> pair_posn -> ... -- so its calls into the
 * alignment layer carry one read each.  This shim REPLACES the two loop functions
 *
 *     sequential_loop_pass1 (bam2bam.c:1143)      sequential_loop_pass2 (bam2bam.c:1178)
 *
 * (both are plain global functions of the reference, reached through the PLT of a shared-library build,
 * so LD_PRELOAD can supply them) with versions that gather a batch of records and make ONE call per phase:
 *
 *   pass 1   bam1_to_seq x n  ->  bwa_gpu_cal_sa_reads_gap   (replaces bwa_cal_sa_reg_gap, bam2bam.c:616,676)
 *            bwa_aln2seq[_core] in record order (it consumes drand48: bwase.c:33,36,78)
 *            -> bwa_gpu_cal_pac_pos                           (replaces bwt_sa at bwase.c:145,152; bam2bam.c:635-636)
 *   pass 2   hit enumeration of all pairs -> bwa_gpu_cal_pac_pos   (bwt_sa at bam2bam.c:752,761)
 *            pairing + bwa_aln2seq_core in record order -> bwa_gpu_cal_pac_pos (bam2bam.c:786)
 *            bwa_paired_sw1 -> bwa_gpu_mate_sw_path           (aln_local_core inside bwa_sw_core, bwape.c:456)
 *            bwa_refine_gapped, bwa_update_bam1, BAM output: the reference's own code, per record.
 *
 * Everything that is not the hot path IS the reference: this file calls its exported functions
 * (read_bam_pair, bam1_to_seq, bwa_aln2seq_core, bwa_cal_pac_pos_core, pairing, bwa_paired_sw1,
 * bwa_refine_gapped, bwa_update_bam1, pair_print_custom, ...) and never restates their arithmetic.
 * Where a reference function interleaves host logic with a hot call (bwa_cal_pac_pos_core around bwt_sa,
 * bwa_paired_sw1/bwa_sw_core around aln_local_core) it is run in RECORD / REPLAY fashion: a first run with
 * the hot call interposed to note its arguments (and return "no result", which leaves the record
 * untouched), one device call for the whole batch, then the real run with the hot call interposed to
 * hand back the device's answers in the same order.
 *
 * Order-sensitive state is kept exactly as `-t 1` has it: drand48 is consumed by bwa_aln2seq_core only
 * (bwase.c), which runs in record order in both passes; the pass-2 position cache for intervals >= 1000
 * wide (bam2bam.c:743-758) is filled by the first record, in record order, that touches a key.
 * The output BAM is therefore record-identical to `bam2bam -t 1` (tests/test_batched_bam2bam.py).
 *
 * The reference's static globals (bwt, bns, pac, gap_opt, pe_opt, the three input flags) are captured
 * by interposing the loaders / option parser that produce them.  Not supported here: `.sai` inputs
 * (-0/-1/-2; broken in the reference itself) -- such records take the per-record path.
 *
 * Build (integration/Makefile): gcc -I$(REF) -I../oracle/zmq_shim -I../include ... -lbwagpu
 * Use:   LD_PRELOAD=integration/libbwa_gpu_batch.so oracle/_ref/ref_driver bam2bam -g idx -t 1 -f out.bam in.bam
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <getopt.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#include <zlib.h>

#include "bamlite.h"
#include "bwtaln.h"
#include "bwase.h"
#include "bwape.h"
#include "khash.h"
#include "bgzf.h"
#include "zmq.h" /* oracle/zmq_shim: the libzmq ABI the reference is built against */
#include "bwa_gpu.h" /* after bwtaln.h: re-uses the reference's own types */

KHASH_MAP_INIT_INT64(64, poslist_t) /* the position cache's type, as bam2bam.c:38 declares it */

/* reference functions this driver calls that no header declares (bam2bam.c) */
void pair_aln(bam_pair_t *p);
void pair_print_custom(gzFile f, bam_pair_t *p);
int read_pair_custom(gzFile f, bam_pair_t *p);
void pair_print_bam(BGZF *output, bam_pair_t *p);
void bwa_update_bam1(bam1_t *out, const bntseq_t *bns, bwa_seq_t *p, const bwa_seq_t *mate, int mode, int max_top2);
void bwa_cal_pac_pos_core(const bwt_t *forward_bwt, const bwt_t *reverse_bwt, bwa_seq_t *seq, const int max_mm, const float fnr);
void bwa_aln2seq(int n_aln, const bwt_aln1_t *aln, bwa_seq_t *s);

#define REAL(ret, name, ...) \
	static ret (*real_##name)(__VA_ARGS__); \
	if (!real_##name) real_##name = (ret (*)(__VA_ARGS__))dlsym(RTLD_NEXT, #name)

static void die(const char *what)
{
	fprintf(stderr, "[bwa_gpu_batch] %s: %s\n", what, bwa_gpu_last_error());
	abort(); /* the reference's convention on this path: xassert -> abort (utils.c:68-83) */
}

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* ------------------------------------------------------------------ the reference's static globals, captured */
static bwt_t *g_bwt[2];        /* bam2bam.c:88  (init_genome_index 855-856) */
static const bntseq_t *g_bns;  /* bam2bam.c:89 */
static ubyte_t *g_pac;         /* bam2bam.c:91 */
static gap_opt_t *g_gap;       /* bam2bam.c:94 */
static pe_opt_t *g_pe;         /* bam2bam.c:95 */
static int g_broken_input, g_skip_duplicates, g_drop_aligned, g_only_aligned; /* bam2bam.c:96-101, options 130 / 131 / 133 / 128 */
static isize_info_t g_null_ii; /* bam2bam.c:106 */

bwt_t *bwt_restore_bwt(const char *fn, int touch)
{
	REAL(bwt_t *, bwt_restore_bwt, const char *, int);
	bwt_t *b = real_bwt_restore_bwt(fn, touch);
	const size_t n = strlen(fn);
	g_bwt[n >= 5 && strcmp(fn + n - 5, ".rbwt") == 0 ? 1 : 0] = b;
	return b;
}

ubyte_t *bwt_restore_pac(const bntseq_t *bns, int touch)
{
	REAL(ubyte_t *, bwt_restore_pac, const bntseq_t *, int);
	g_bns = bns;
	g_pac = real_bwt_restore_pac(bns, touch);
	return g_pac;
}

gap_opt_t *gap_init_opt(void)
{
	REAL(gap_opt_t *, gap_init_opt, void);
	return g_gap = real_gap_init_opt();
}

pe_opt_t *bwa_init_pe_opt(void)
{
	REAL(pe_opt_t *, bwa_init_pe_opt, void);
	return g_pe = real_bwa_init_pe_opt();
}

int getopt_long(int argc, char *const argv[], const char *optstring, const struct option *longopts, int *longindex)
{
	REAL(int, getopt_long, int, char *const *, const char *, const struct option *, int *);
	const int c = real_getopt_long(argc, argv, optstring, longopts, longindex);
	if (c == 128) g_only_aligned = 1;      /* bam2bam.c:1988 */
	if (c == 130) g_broken_input = 1;      /* bam2bam.c:1991 */
	if (c == 131) g_skip_duplicates = 1;
	if (c == 133) g_drop_aligned = 1;
	return c;
}

static int unique_rec(const bam_pair_t *p) /* bam2bam.c:595-606 */
{
	int i;
	if (!g_skip_duplicates) return 1;
	if (p->kind == eof_marker) return 0;
	for (i = 0; i != (int)p->kind; ++i)
		if (p->bam_rec[i].core.flag & SAM_FDP) return 0;
	return 1;
}


/* ------------------------------------------------------------------ host threads
 * With the hot path on the device, what is left of a bam2bam run is per-record host work of the reference
 * (bam1_to_seq, bwa_refine_gapped, bwa_update_bam1, the temp-file codec, deflate).  The reference itself runs these
 * functions concurrently in its worker threads (run_worker_thread, bam2bam.c:1387), so they are re-entrant; the shim
 * spreads each such phase of a batch over BWAGPU_SHIM_THREADS threads (default: the host's cores, at most 32).
 * Everything order-sensitive (drand48 in bwa_aln2seq_core, the position cache, isize statistics) stays serial. */
typedef void (*pf_fn)(size_t i, void *ctx);
typedef struct { size_t n, grain; size_t next; pf_fn fn; void *ctx; } pf_job_t;

static int shim_threads(void)
{
	static int n;
	if (!n) {
		const char *e = getenv("BWAGPU_SHIM_THREADS");
		n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
		if (n < 1) n = 1;
		if (n > 32) n = 32;
	}
	return n;
}

static void *pf_worker(void *arg)
{
	pf_job_t *j = (pf_job_t *)arg;
	for (;;) {
		const size_t lo = __sync_fetch_and_add(&j->next, j->grain);
		size_t i, hi;
		if (lo >= j->n) break;
		hi = lo + j->grain < j->n ? lo + j->grain : j->n;
		for (i = lo; i < hi; ++i) j->fn(i, j->ctx);
	}
	return 0;
}

static void parallel_for(size_t n, size_t grain, pf_fn fn, void *ctx)
{
	pf_job_t j = {n, grain ? grain : 1, 0, fn, ctx};
	pthread_t th[32];
	int t, nt = shim_threads();
	if ((size_t)nt > (n + j.grain - 1) / j.grain) nt = (int)((n + j.grain - 1) / j.grain);
	if (nt <= 1) { pf_worker(&j); return; }
	for (t = 1; t < nt; ++t) pthread_create(&th[t], 0, pf_worker, &j);
	pf_worker(&j);
	for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
}


/* ------------------------------------------------------------------ input BAM: inflate on a side thread
 * bamlite reads the input with gzread (bamlite.h:8-11), three small calls per record, on the thread that also parses
 * and allocates the records; inflate is about half of that thread's time and it is the slowest stage of pass 1.  The
 * first file opened with gzopen(.., "r") -- the input BAM (bwa_bam_open) -- gets a read-ahead thread that gzread()s it in
 * 256 KB pieces into a ring; the interposed gzread serves bamlite's calls from the ring.  Every other gzFile (the
 * temporary file) goes straight to zlib.  BWAGPU_READAHEAD=0 turns it off. */
#define RA_CAP ((size_t)16 << 20)
#define RA_PIECE ((size_t)256 << 10)
typedef struct {
	gzFile real;
	pthread_t th;
	uint8_t *ring;
	volatile size_t head, tail; /* bytes produced / consumed so far */
	volatile int eof, stop;
} readahead_t;
static readahead_t g_ra;

static void *ra_main(void *arg)
{
	REAL(int, gzread, gzFile, voidp, unsigned);
	uint8_t *piece = (uint8_t *)malloc(RA_PIECE);
	(void)arg;
	while (!g_ra.stop) {
		int got, done = 0;
		while (!g_ra.stop && RA_CAP - (g_ra.head - __atomic_load_n(&g_ra.tail, __ATOMIC_ACQUIRE)) < RA_PIECE) usleep(100);
		if (g_ra.stop) break;
		got = real_gzread(g_ra.real, piece, (unsigned)RA_PIECE);
		if (got <= 0) break;
		while (done < got) {
			const size_t at = (g_ra.head + (size_t)done) % RA_CAP;
			const size_t run = RA_CAP - at < (size_t)(got - done) ? RA_CAP - at : (size_t)(got - done);
			memcpy(g_ra.ring + at, piece + done, run);
			done += (int)run;
		}
		__atomic_store_n(&g_ra.head, g_ra.head + (size_t)got, __ATOMIC_RELEASE);
	}
	__atomic_store_n(&g_ra.eof, 1, __ATOMIC_RELEASE);
	free(piece);
	return 0;
}

gzFile gzopen(const char *path, const char *mode)
{
	REAL(gzFile, gzopen, const char *, const char *);
	gzFile f = real_gzopen(path, mode);
	const char *e = getenv("BWAGPU_READAHEAD");
	if (f && !g_ra.real && mode && mode[0] == 'r' && !(e && atoi(e) == 0)) {
		memset(&g_ra, 0, sizeof(g_ra));
		g_ra.ring = (uint8_t *)malloc(RA_CAP);
		if (g_ra.ring) {
			g_ra.real = f;
			pthread_create(&g_ra.th, 0, ra_main, 0);
		}
	}
	return f;
}

int gzread(gzFile f, voidp buf, unsigned len)
{
	REAL(int, gzread, gzFile, voidp, unsigned);
	unsigned done = 0;
	if (!g_ra.real || f != g_ra.real) return real_gzread(f, buf, len);
	while (done < len) {
		size_t avail = __atomic_load_n(&g_ra.head, __ATOMIC_ACQUIRE) - g_ra.tail;
		if (avail == 0) {
			if (__atomic_load_n(&g_ra.eof, __ATOMIC_ACQUIRE) && __atomic_load_n(&g_ra.head, __ATOMIC_ACQUIRE) == g_ra.tail) break;
			usleep(50);
			continue;
		}
		{
			const size_t at = g_ra.tail % RA_CAP;
			size_t run = avail < (size_t)(len - done) ? avail : (size_t)(len - done);
			if (RA_CAP - at < run) run = RA_CAP - at;
			memcpy((uint8_t *)buf + done, g_ra.ring + at, run);
			done += (unsigned)run;
			__atomic_store_n(&g_ra.tail, g_ra.tail + run, __ATOMIC_RELEASE);
		}
	}
	return (int)done;
}

int gzclose(gzFile f)
{
	REAL(int, gzclose, gzFile);
	if (g_ra.real && f == g_ra.real) {
		g_ra.stop = 1;
		pthread_join(g_ra.th, 0);
		free(g_ra.ring);
		g_ra.real = 0;
	}
	return real_gzclose(f);
}

/* ------------------------------------------------------------------ device context */
static int g_ready;
static long g_calls_aln, g_calls_sa, g_calls_sw, g_reads_aln, g_q_sa, g_jobs_sw;
static double g_t_aln, g_t_sa, g_t_sw;

static void report(void)
{
	fprintf(stderr, "[bwa_gpu_batch] device calls: cal_sa_reads_gap=%ld (%ld reads, %.2f s)  cal_pac_pos=%ld (%ld queries, %.2f s)  "
	                "mate_sw_path=%ld (%ld jobs, %.2f s)\n", g_calls_aln, g_reads_aln, g_t_aln, g_calls_sa, g_q_sa, g_t_sa,
	        g_calls_sw, g_jobs_sw, g_t_sw);
}

static void ensure_gpu(void);
static void *ensure_gpu_thread(void *arg) { (void)arg; ensure_gpu(); return 0; }

static void ensure_gpu(void)
{
	if (g_ready) return;
	if (!g_bwt[0] || !g_bwt[1] || !g_bns || !g_pac || !g_gap || !g_pe) {
		fprintf(stderr, "[bwa_gpu_batch] the reference's index/option globals were not seen being loaded\n");
		abort();
	}
	{
		const char *e = getenv("BWAGPU_NDEV");
		int ids[16], n = e ? atoi(e) : 1, i;
		if (n < 1) n = 1;
		if (n > 16) n = 16;
		for (i = 0; i < n; ++i) ids[i] = i;
		if (bwa_gpu_init(n, ids)) die("bwa_gpu_init");
	}
	if (bwa_gpu_load_index(g_bwt, g_pac, g_bns->l_pac)) die("bwa_gpu_load_index");
	g_ready = 1;
	atexit(report);
}

static size_t batch_records(void)
{
	const char *e = getenv("BWAGPU_BATCH_RECORDS");
	const long v = e ? atol(e) : 0;
	return v > 0 ? (size_t)v : (size_t)1 << 18;
}

/* ------------------------------------------------------------------ bwt_sa: real / replay */
typedef struct { size_t n, m; bwtint_t *k; uint8_t *which; bwtint_t *out; } saq_t;
static saq_t g_q;
static int g_sa_replay;
static size_t g_sa_pos;

static void saq_push(saq_t *q, bwtint_t k, int forward)
{
	if (q->n == q->m) {
		q->m = q->m ? q->m << 1 : 1 << 16;
		q->k = (bwtint_t *)realloc(q->k, q->m * sizeof(bwtint_t));
		q->which = (uint8_t *)realloc(q->which, q->m);
	}
	q->k[q->n] = k; q->which[q->n] = (uint8_t)(forward != 0); ++q->n;
}

static void saq_run(saq_t *q)
{
	const double t0 = now();
	q->out = (bwtint_t *)realloc(q->out, (q->n + 1) * sizeof(bwtint_t));
	if (q->n && bwa_gpu_cal_pac_pos((int64_t)q->n, q->k, q->which, q->out)) die("bwa_gpu_cal_pac_pos");
	++g_calls_sa; g_q_sa += (long)q->n; g_t_sa += now() - t0;
	g_sa_pos = 0;
}

static bwtint_t saq_next(saq_t *q, const bwt_t *bwt, bwtint_t k)
{
	if (g_sa_pos >= q->n || q->k[g_sa_pos] != k || q->which[g_sa_pos] != (uint8_t)(bwt == g_bwt[0])) {
		fprintf(stderr, "[bwa_gpu_batch] bwt_sa replay out of step at query %zu\n", g_sa_pos);
		abort();
	}
	return q->out[g_sa_pos++];
}

bwtint_t bwt_sa(const bwt_t *bwt, bwtint_t k)
{
	REAL(bwtint_t, bwt_sa, const bwt_t *, bwtint_t);
	if (g_sa_replay) return saq_next(&g_q, bwt, k);
	return real_bwt_sa(bwt, k); /* index construction, per-record fallbacks */
}

/* ------------------------------------------------------------------ aln_local_core: real / record / replay */
enum { SW_REAL = 0, SW_RECORD = 1, SW_REPLAY = 2 };
static int g_sw_mode;
static int64_t g_sw_beg; /* *beg of the bwa_sw_core call in progress */
typedef struct { size_t n, m; bwa_gpu_sw_job_t *job; size_t *seq_off; size_t sn, sm; ubyte_t *seqs; bwa_gpu_path_res_t *res; const bwa_cigar_t *pool; } swq_t;
static swq_t g_sw;
static size_t g_sw_pos;

bwa_cigar_t *bwa_sw_core(bwtint_t l_pac, const ubyte_t *pacseq, int len, const ubyte_t *seq, int64_t *beg, int reglen,
                         int *n_cigar, uint32_t *cnt)
{
	REAL(bwa_cigar_t *, bwa_sw_core, bwtint_t, const ubyte_t *, int, const ubyte_t *, int64_t *, int, int *, uint32_t *);
	g_sw_beg = *beg; /* the window the reference is about to unpack (bwape.c:447-450) */
	return real_bwa_sw_core(l_pac, pacseq, len, seq, beg, reglen, n_cigar, cnt);
}

int aln_local_core(unsigned char *seq1, int len1, unsigned char *seq2, int len2, const AlnParam *ap, path_t *path, int *path_len,
                   int thres, int *subo)
{
	REAL(int, aln_local_core, unsigned char *, int, unsigned char *, int, const AlnParam *, path_t *, int *, int, int *);
	if (g_sw_mode == SW_RECORD) {
		swq_t *q = &g_sw;
		if (q->n == q->m) {
			q->m = q->m ? q->m << 1 : 1 << 12;
			q->job = (bwa_gpu_sw_job_t *)realloc(q->job, q->m * sizeof(*q->job));
			q->seq_off = (size_t *)realloc(q->seq_off, q->m * sizeof(size_t));
		}
		if (q->sn + (size_t)len2 > q->sm) {
			q->sm = (q->sn + (size_t)len2) * 2 + 4096;
			q->seqs = (ubyte_t *)realloc(q->seqs, q->sm);
		}
		memcpy(q->seqs + q->sn, seq2, (size_t)len2); /* bwa_paired_sw1 un-reverses the read right after the call */
		q->job[q->n].beg = g_sw_beg; q->job[q->n].reglen = len1; q->job[q->n].len = len2; q->job[q->n].seq = 0;
		q->seq_off[q->n] = q->sn; q->sn += (size_t)len2; ++q->n;
		return -1; /* "no alignment": bwa_sw_core returns 0 and bwa_paired_sw1 leaves both reads untouched (bwape.c:457-460) */
	}
	if (g_sw_mode == SW_REPLAY) {
		swq_t *q = &g_sw;
		const bwa_gpu_path_res_t *r;
		const bwa_cigar_t *cg;
		int n = 0, c, t, i, j;
		if (g_sw_pos >= q->n || q->job[g_sw_pos].reglen != len1 || q->job[g_sw_pos].len != len2) {
			fprintf(stderr, "[bwa_gpu_batch] aln_local_core replay out of step at job %zu\n", g_sw_pos);
			abort();
		}
		r = &q->res[g_sw_pos++];
		if (r->score < 0) return r->score;
		/* path_t[] as aln_global_core's backtrace leaves it (stdaln.c:496-512, shifted at 741-744): path[0] = the end
		 * cell, path[path_len-1] = the start cell; an element's ctype says how its cell was entered */
		cg = q->pool + r->cigar_off;
		for (c = 0; c < r->n_cigar; ++c) n += cg[c] & 0x3fff;
		*path_len = n;
		if (n == 0) return r->score;
		i = r->start_i; j = r->start_j;
		for (c = 0, t = n - 1; c < r->n_cigar; ++c) {
			const int op = cg[c] >> 14, run = cg[c] & 0x3fff;
			int u;
			for (u = 0; u < run; ++u, --t) {
				if (t != n - 1) { if (op != FROM_I) ++i; if (op != FROM_D) ++j; }
				path[t].i = i; path[t].j = j; path[t].ctype = (unsigned char)op;
			}
		}
		if (path[0].i != r->end_i || path[0].j != r->end_j) {
			fprintf(stderr, "[bwa_gpu_batch] device path does not end at its end cell (%d,%d) vs (%d,%d)\n", path[0].i, path[0].j, r->end_i, r->end_j);
			abort();
		}
		if (subo) *subo = 0;
		return r->score;
	}
	return real_aln_local_core(seq1, len1, seq2, len2, ap, path, path_len, thres, subo);
}

static void swq_run(swq_t *q)
{
	size_t i;
	const double t0 = now();
	for (i = 0; i < q->n; ++i) q->job[i].seq = q->seqs + q->seq_off[i];
	q->res = (bwa_gpu_path_res_t *)realloc(q->res, (q->n + 1) * sizeof(*q->res));
	q->pool = 0;
	if (q->n && bwa_gpu_mate_sw_path((int)q->n, q->job, q->res, &q->pool)) die("bwa_gpu_mate_sw_path");
	++g_calls_sw; g_jobs_sw += (long)q->n; g_t_sw += now() - t0;
	g_sw_pos = 0;
}


/* ------------------------------------------------------------------ the intermediate file, kept in memory
 * Pass 1 hands its records to pass 2 through a gzip'ed temporary file in the reference (pair_print_custom /
 * read_pair_custom, bam2bam.c:1099-1137: the 0MQ message encoding of a record, length-prefixed).  Deflate and inflate of
 * that file were 6 of the 21 seconds of a 2 M-read run once the hot path was on the device.  The shim keeps the SAME
 * encoded messages (the reference's msg_init_from_pair / pair_init_from_msg, so a record makes the same round trip) in
 * memory, up to BWAGPU_MEMTEMP_MB (default: a quarter of physical memory, at most 64 GB), and spills whatever comes
 * after that to the reference's temporary file in the reference's format.  Encoding and decoding run on the host threads. */
void msg_init_from_pair(zmq_msg_t *m, bam_pair_t *p);
void pair_init_from_msg(bam_pair_t *p, zmq_msg_t *m);

#define MT_CHUNK ((size_t)64 << 20)
typedef struct {
	uint8_t **chunk; size_t n_chunk, m_chunk, used;
	uint8_t **rec; uint32_t *len; size_t n_rec, m_rec, rd;
	size_t bytes, cap;
	int spilled;
} memtemp_t;
static memtemp_t g_mt;

static size_t memtemp_cap(void)
{
	const char *e = getenv("BWAGPU_MEMTEMP_MB"), *eb = getenv("BWAGPU_MEMTEMP_BYTES");
	size_t cap;
	if (eb) return (size_t)atoll(eb);
	if (e) return (size_t)atoll(e) << 20;
	cap = (size_t)sysconf(_SC_PHYS_PAGES) * (size_t)sysconf(_SC_PAGESIZE) / 4;
	return cap > ((size_t)64 << 30) ? (size_t)64 << 30 : cap;
}

static int memtemp_put(const void *data, uint32_t len)
{
	memtemp_t *t = &g_mt;
	if (t->spilled || t->bytes + len > t->cap || len > MT_CHUNK) { t->spilled = 1; return 0; }
	if (t->n_chunk == 0 || t->used + len > MT_CHUNK) {
		if (t->n_chunk == t->m_chunk) { t->m_chunk = t->m_chunk ? t->m_chunk << 1 : 16; t->chunk = (uint8_t **)realloc(t->chunk, t->m_chunk * sizeof(*t->chunk)); }
		t->chunk[t->n_chunk] = (uint8_t *)malloc(MT_CHUNK);
		if (!t->chunk[t->n_chunk]) { t->spilled = 1; return 0; }
		++t->n_chunk; t->used = 0;
	}
	if (t->n_rec == t->m_rec) {
		t->m_rec = t->m_rec ? t->m_rec << 1 : 1 << 20;
		t->rec = (uint8_t **)realloc(t->rec, t->m_rec * sizeof(*t->rec));
		t->len = (uint32_t *)realloc(t->len, t->m_rec * sizeof(*t->len));
	}
	t->rec[t->n_rec] = t->chunk[t->n_chunk - 1] + t->used;
	t->len[t->n_rec] = len;
	memcpy(t->rec[t->n_rec], data, len);
	++t->n_rec; t->used += len; t->bytes += len;
	return 1;
}

static void memtemp_free(void)
{
	size_t i;
	for (i = 0; i < g_mt.n_chunk; ++i) free(g_mt.chunk[i]);
	free(g_mt.chunk); free(g_mt.rec); free(g_mt.len);
	memset(&g_mt, 0, sizeof(g_mt));
}

typedef struct { bam_pair_t *recs; zmq_msg_t *msgs; size_t base; } codec_ctx_t;
static void decode_one(size_t i, void *ctx)
{
	codec_ctx_t *c = (codec_ctx_t *)ctx;
	zmq_msg_t m;
	zmq_msg_init_data(&m, g_mt.rec[c->base + i], g_mt.len[c->base + i], 0, 0);
	pair_init_from_msg(&c->recs[i], &m);
	zmq_msg_close(&m);
}

/* pass 1: records [0, n) -> memory (or, once that is full, the reference's temporary file), then destroyed.  Serial on purpose:
 * encode and destroy are malloc/free of blocks other threads allocated, and spreading them over threads is slower (measured:
 * 0.29 s per 300 k records on one thread, 0.74 s on eight) */
static void destroy_records(bam_pair_t *recs, size_t n)
{
	size_t i;
	for (i = 0; i < n; ++i) bam_destroy_pair(&recs[i]);
}

static void store_records(gzFile temporary, bam_pair_t *recs, size_t n, int destroy)
{
	size_t i;
	for (i = 0; i < n; ++i) {
		zmq_msg_t m;
		uint32_t len;
		msg_init_from_pair(&m, &recs[i]);
		len = (uint32_t)zmq_msg_size(&m);
		if (!memtemp_put(zmq_msg_data(&m), len)) { /* pair_print_custom's two writes (bam2bam.c:1104-1105) */
			if (gzwrite(temporary, &len, sizeof(len)) != (int)sizeof(len) || gzwrite(temporary, zmq_msg_data(&m), len) != (int)len) {
				fprintf(stderr, "[bwa_gpu_batch] error writing temporary file\n");
				exit(1);
			}
		}
		zmq_msg_close(&m);
		if (destroy) bam_destroy_pair(&recs[i]);
	}
}

/* pass 2: up to B records, the ones kept in memory first */
static size_t load_records(gzFile temporary, bam_pair_t *recs, size_t B, long *tot_seqs)
{
	size_t n = 0, i;
	if (g_mt.rd < g_mt.n_rec) {
		codec_ctx_t c = {recs, 0, g_mt.rd};
		n = g_mt.n_rec - g_mt.rd < B ? g_mt.n_rec - g_mt.rd : B;
		parallel_for(n, 1024, decode_one, &c);
		g_mt.rd += n;
		for (i = 0; i < n; ++i) *tot_seqs += recs[i].kind;
	}
	while (n < B) {
		const int rc = read_pair_custom(temporary, &recs[n]);
		if (rc < 0) { fprintf(stderr, "[bwa_gpu_batch] error reading intermediate file\n"); exit(1); }
		if (rc == 0) break;
		*tot_seqs += recs[n].kind;
		++n;
	}
	return n;
}

/* ------------------------------------------------------------------ BAM output: BGZF blocks deflated on the host threads
 * pair_print_bam -> bwa_print_bam1 -> bgzf_write (bam2bam.c:304-321, 908-924; bgzf.c:594-623) deflates one 64 KB block at a
 * time on the calling thread.  Here the batch's records are laid out as the same byte stream, cut into BGZF blocks, the
 * blocks are deflated in parallel (same level, same header and footer as deflate_block, bgzf.c:265-340) and written in order
 * to the BGZF handle's own FILE.  The decompressed stream -- what any BAM reader sees -- is identical; block boundaries are
 * not (the reference cuts at 65536 bytes of input, this writer at 65280 so that a block always fits). */
#define OB_IN 65280
#define OB_OUT 65536
typedef struct { bam_pair_t *recs; size_t *off; uint8_t *ubuf; size_t total; uint8_t *cbuf; int *clen; int level; int failed; } ob_ctx_t;

static size_t rec_bytes(const bam_pair_t *p) /* pair_print_bam's filter and bwa_print_bam1's record size */
{
	size_t n = 0;
	int i;
	if (g_only_aligned)
		for (i = 0; i != (int)p->kind; ++i)
			if (p->bam_rec[i].core.flag & SAM_FSU) return 0;
	for (i = 0; i != (int)p->kind; ++i) n += 4 + sizeof(bam1_core_t) + (size_t)p->bam_rec[i].data_len;
	return n;
}

static void ob_fill_one(size_t i, void *ctx)
{
	ob_ctx_t *c = (ob_ctx_t *)ctx;
	const bam_pair_t *p = &c->recs[i];
	uint8_t *q = c->ubuf + c->off[i];
	int j;
	if (c->off[i + 1] == c->off[i]) return;
	for (j = 0; j != (int)p->kind; ++j) { /* bwa_print_bam1 (bam2bam.c:304-321) */
		const bam1_t *b = &p->bam_rec[j];
		uint32_t w[9];
		w[0] = (uint32_t)(sizeof(bam1_core_t) + b->data_len);
		w[1] = (uint32_t)b->core.tid; w[2] = (uint32_t)b->core.pos;
		w[3] = (uint32_t)((int)b->core.bin << 16 | (int)b->core.qual << 8 | (int)b->core.l_qname);
		w[4] = (uint32_t)((int)b->core.flag << 16 | (int)b->core.n_cigar);
		w[5] = (uint32_t)b->core.l_qseq; w[6] = (uint32_t)b->core.mtid; w[7] = (uint32_t)b->core.mpos; w[8] = (uint32_t)b->core.isize;
		memcpy(q, w, 36); q += 36;
		memcpy(q, b->data, (size_t)b->data_len); q += b->data_len;
	}
}

static void ob_deflate_one(size_t k, void *ctx)
{
	ob_ctx_t *c = (ob_ctx_t *)ctx;
	const uint8_t *in = c->ubuf + k * OB_IN;
	const int in_len = (int)(c->total - k * OB_IN < OB_IN ? c->total - k * OB_IN : OB_IN);
	uint8_t *out = c->cbuf + k * OB_OUT;
	static const uint8_t hdr[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 0, 0}; /* bgzf.c:274-291 */
	z_stream zs;
	uint32_t crc, len;
	memcpy(out, hdr, 18);
	memset(&zs, 0, sizeof(zs));
	zs.next_in = (Bytef *)in; zs.avail_in = (uInt)in_len;
	zs.next_out = out + 18; zs.avail_out = OB_OUT - 18 - 8;
	if (deflateInit2(&zs, c->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK || deflate(&zs, Z_FINISH) != Z_STREAM_END) { c->failed = 1; return; }
	deflateEnd(&zs);
	len = (uint32_t)zs.total_out + 18 + 8;
	out[16] = (uint8_t)((len - 1) & 0xff); out[17] = (uint8_t)((len - 1) >> 8);
	crc = (uint32_t)crc32(crc32(0L, 0, 0), in, (uInt)in_len);
	memcpy(out + 18 + zs.total_out, &crc, 4);
	memcpy(out + 18 + zs.total_out + 4, &in_len, 4);
	c->clen[k] = (int)len;
}

static void write_records_bam(BGZF *output, bam_pair_t *recs, size_t n)
{
	static size_t *off; static size_t m_off;
	static uint8_t *ubuf, *cbuf; static size_t m_ubuf, m_cbuf;
	static int *clen; static size_t m_clen;
	ob_ctx_t c;
	size_t i, nblk;
	if (n + 1 > m_off) { m_off = n + 1; off = (size_t *)realloc(off, m_off * sizeof(*off)); }
	off[0] = 0;
	for (i = 0; i < n; ++i) off[i + 1] = off[i] + rec_bytes(&recs[i]);
	memset(&c, 0, sizeof(c));
	c.recs = recs; c.off = off; c.total = off[n]; c.level = output->compress_level;
	if (c.total) {
		if (bgzf_flush(output) != 0) { fprintf(stderr, "[bwa_gpu_batch] BAM write failed\n"); exit(1); } /* what bgzf_write buffered so far (the header) */
		nblk = (c.total + OB_IN - 1) / OB_IN;
		if (c.total > m_ubuf) { m_ubuf = c.total + c.total / 4; ubuf = (uint8_t *)realloc(ubuf, m_ubuf); }
		if (nblk * OB_OUT > m_cbuf) { m_cbuf = nblk * OB_OUT + (nblk / 4) * OB_OUT; cbuf = (uint8_t *)realloc(cbuf, m_cbuf); }
		if (nblk > m_clen) { m_clen = nblk + nblk / 4; clen = (int *)realloc(clen, m_clen * sizeof(int)); }
		c.ubuf = ubuf; c.cbuf = cbuf; c.clen = clen;
		parallel_for(n, 2048, ob_fill_one, &c);
		parallel_for(nblk, 4, ob_deflate_one, &c);
		if (c.failed) { fprintf(stderr, "[bwa_gpu_batch] deflate failed\n"); exit(1); }
		for (i = 0; i < nblk; ++i) {
			if (fwrite(cbuf + i * OB_OUT, 1, (size_t)clen[i], output->file) != (size_t)clen[i]) { fprintf(stderr, "[bwa_gpu_batch] BAM write failed\n"); exit(1); }
			output->block_address += clen[i];
		}
	}
	for (i = 0; i < n; ++i) bam_destroy_pair(&recs[i]);
}

/* ------------------------------------------------------------------ pass 1 */
static void gpu_align(int m, bwa_seq_t *flat)
{
	const double t0 = now();
	if (m && bwa_gpu_cal_sa_reads_gap(m, flat, g_gap)) die("bwa_gpu_cal_sa_reads_gap");
	++g_calls_aln; g_reads_aln += m; g_t_aln += now() - t0;
}

static int is_mapped(const bwa_seq_t *p) { return p->type == BWA_TYPE_UNIQUE || p->type == BWA_TYPE_REPEAT; }

static void to_seq_one(size_t i, void *ctx) /* bam1_to_seq of aln_singleton / aln_pair (bam2bam.c:615, 674-675) */
{
	bam_pair_t *r = (bam_pair_t *)ctx + i;
	int j;
	if (r->phase != pristine || !unique_rec(r)) return;
	for (j = 0; j != (int)r->kind; ++j) bam1_to_seq(&r->bam_rec[j], &r->bwa_seq[j], 1, g_gap->trim_qual);
}

/* aln_* (bam2bam.c:608-620, 660-681) for records [0, n): what sequential_loop_pass1 and a worker thread do first to a pristine
 * record, with the search hoisted out of the per-record loop.  `flat` holds >= 2 n elements. */
static void align_range(bam_pair_t *recs, size_t n, bwa_seq_t *flat, double *t_toseq)
{
	size_t i;
	int m = 0, j;
	double t1 = now();
	/* aln_singleton / aln_pair without the search ... */
	parallel_for(n, 2048, to_seq_one, recs);
	for (i = 0; i < n; ++i) {
		bam_pair_t *r = &recs[i];
		if (r->phase != pristine) continue;
		if (unique_rec(r))
			for (j = 0; j != (int)r->kind; ++j) flat[m++] = r->bwa_seq[j];
	}
	*t_toseq += now() - t1;
	/* ... which is ONE device call for the batch */
	gpu_align(m, flat);
	t1 = now();
	for (i = 0, m = 0; i < n; ++i) {
		bam_pair_t *r = &recs[i];
		if (r->phase != pristine) continue;
		if (unique_rec(r))
			for (j = 0; j != (int)r->kind; ++j) r->bwa_seq[j] = flat[m++];
		r->phase = aligned;
	}
	*t_toseq += now() - t1;
}

/* posn_* (bam2bam.c:622-641, 683-703) for records [0, n): primary-hit selection on the host, in record order (drand48),
 * one device call for the SA rows whose coordinates are wanted, then the reference's own bwa_cal_pac_pos_core fed from the answers */
static void position_range(bam_pair_t *recs, size_t n, double *t_host)
{
	size_t i;
	int j;
	double t1 = now();
	g_q.n = 0;
	for (i = 0; i < n; ++i) {
		bam_pair_t *r = &recs[i];
		if (r->phase != aligned || !unique_rec(r)) continue;
		for (j = 0; j != (int)r->kind; ++j) {
			bwa_seq_t *p = &r->bwa_seq[j];
			int k;
			if (r->kind == singleton) bwa_aln2seq_core(p->n_aln, p->aln, p, 1, g_pe->max_occ_se);
			else { p->n_multi = 0; bwa_aln2seq(p->n_aln, p->aln, p); }
			if (is_mapped(p)) saq_push(&g_q, p->sa, p->strand);
			if (r->kind == singleton)
				for (k = 0; k < p->n_multi; ++k) saq_push(&g_q, p->multi[k].pos, p->multi[k].strand);
		}
	}
	*t_host += now() - t1;
	saq_run(&g_q);
	t1 = now();
	g_sa_replay = 1;
	for (i = 0; i < n; ++i) {
		bam_pair_t *r = &recs[i];
		if (r->phase != aligned) continue;
		if (unique_rec(r))
			for (j = 0; j != (int)r->kind; ++j) {
				bwa_seq_t *p = &r->bwa_seq[j];
				int k;
				bwa_cal_pac_pos_core(g_bwt[0], g_bwt[1], p, g_gap->max_diff, g_gap->fnr);
				if (r->kind == singleton)
					for (k = 0; k < p->n_multi; ++k) { /* bam2bam.c:633-637 */
						bwt_multi1_t *q = p->multi + k;
						if (q->strand) q->pos = bwt_sa(g_bwt[0], q->pos);
						else q->pos = g_bwt[1]->seq_len - (bwt_sa(g_bwt[1], q->pos) + p->len);
					}
			}
		r->phase = positioned;
	}
	g_sa_replay = 0;
	if (g_sa_pos != g_q.n) { fprintf(stderr, "[bwa_gpu_batch] %zu of %zu SA answers unused\n", g_q.n - g_sa_pos, g_q.n); abort(); }
	*t_host += now() - t1;
}

static void align_position_range(bam_pair_t *recs, size_t n, bwa_seq_t *flat, double *t_toseq, double *t_host)
{
	align_range(recs, n, flat, t_toseq);
	position_range(recs, n, t_host);
}

/* Pass 1 as a pipeline of five stages, each on its own thread, batches handed on in order through a ring of slots:
 *   read     read_bam_pair (inflate + parse; bamlite is a sequential gzread)
 *   align    bam1_to_seq on the host threads + the search on the device            (the calling thread)
 *   position bwa_aln2seq_core in record order (drand48), SA rows on the device, bwa_cal_pac_pos_core, improve_isize_est
 *   store    the reference's record encoding into memory / its temporary file
 *   destroy  bam_destroy_pair of the batch (a dozen free()s per record)
 * Every order-sensitive piece of state belongs to exactly one stage, and a stage sees the batches in input order, so the
 * result is what the one-record-at-a-time loop (bam2bam.c:1143-1176) produces. */
#define P1_SLOTS 8 /* the reader may run this many batches ahead (it does while the device context is created) */
enum { SL_FREE = 0, SL_READ, SL_ALIGNED, SL_POSITIONED, SL_STORED };
typedef struct {
	pthread_mutex_t mu;
	pthread_cond_t cv;
	int state[P1_SLOTS];
	bam_pair_t *recs[P1_SLOTS];
	size_t n[P1_SLOTS];
	long seqs[P1_SLOTS];
	size_t B;
	bwa_seqio_t *ks;
	gzFile temporary;
	khash_t(isize_infos) *iinfos;
	double t0, t_read, t_host, t_write, t_destroy;
	long tot_seqs;
} pipe1_t;

static void slot_wait(pipe1_t *P, int slot, int want)
{
	pthread_mutex_lock(&P->mu);
	while (P->state[slot] != want) pthread_cond_wait(&P->cv, &P->mu);
	pthread_mutex_unlock(&P->mu);
}

static void slot_set(pipe1_t *P, int slot, int st)
{
	pthread_mutex_lock(&P->mu);
	P->state[slot] = st;
	pthread_cond_broadcast(&P->cv);
	pthread_mutex_unlock(&P->mu);
}

static void *stage_read(void *arg)
{
	pipe1_t *P = (pipe1_t *)arg;
	unsigned q;
	for (q = 0;; ++q) {
		const int slot = (int)(q % P1_SLOTS);
		bam_pair_t *recs = P->recs[slot];
		size_t n = 0;
		long seqs = 0;
		double t;
		slot_wait(P, slot, SL_FREE);
		t = now();
		while (n < P->B) {
			const int rc = read_bam_pair(P->ks, &recs[n], g_broken_input, g_drop_aligned);
			if (rc < 0) {
				fprintf(stderr, "[sequential_loop_pass1] error reading input BAM%s\n", rc == -2 ? " (lone mate)" : "");
				exit(1);
			}
			if (rc == 0) break;
			seqs += recs[n].kind;
			++n;
		}
		P->t_read += now() - t;
		P->n[slot] = n; P->seqs[slot] = seqs;
		slot_set(P, slot, SL_READ);
		if (n == 0) break; /* an empty batch is the end marker; it travels through every stage */
	}
	return 0;
}

static void *stage_position(void *arg)
{
	pipe1_t *P = (pipe1_t *)arg;
	unsigned q;
	for (q = 0;; ++q) {
		const int slot = (int)(q % P1_SLOTS);
		bam_pair_t *recs;
		size_t n, i;
		double t1;
		slot_wait(P, slot, SL_ALIGNED);
		recs = P->recs[slot]; n = P->n[slot];
		if (n) {
			position_range(recs, n, &P->t_host);
			t1 = now();
			for (i = 0; i < n; ++i) /* the unchanged tail of the loop (bam2bam.c:1167-1170) */
				if (unique_rec(&recs[i])) improve_isize_est(P->iinfos, &recs[i], g_pe->ap_prior, g_bwt[0]->seq_len);
			P->t_host += now() - t1;
		}
		slot_set(P, slot, SL_POSITIONED);
		if (n == 0) break;
	}
	return 0;
}

static void *stage_store(void *arg)
{
	pipe1_t *P = (pipe1_t *)arg;
	unsigned q;
	for (q = 0;; ++q) {
		const int slot = (int)(q % P1_SLOTS);
		size_t n;
		double t1;
		slot_wait(P, slot, SL_POSITIONED);
		n = P->n[slot];
		if (n == 0) { slot_set(P, slot, SL_STORED); break; }
		t1 = now();
		store_records(P->temporary, P->recs[slot], n, 0);
		P->t_write += now() - t1;
		P->tot_seqs += P->seqs[slot];
		fprintf(stderr, "[sequential_loop_pass1] %ld sequences processed in %.2f sec\n", P->tot_seqs, now() - P->t0);
		slot_set(P, slot, SL_STORED);
	}
	return 0;
}

static void *stage_destroy(void *arg)
{
	pipe1_t *P = (pipe1_t *)arg;
	unsigned q;
	for (q = 0;; ++q) {
		const int slot = (int)(q % P1_SLOTS);
		double t1;
		slot_wait(P, slot, SL_STORED);
		if (P->n[slot] == 0) break;
		t1 = now();
		destroy_records(P->recs[slot], P->n[slot]);
		P->t_destroy += now() - t1;
		slot_set(P, slot, SL_FREE);
	}
	return 0;
}

void sequential_loop_pass1(bwa_seqio_t *ks, gzFile temporary, khash_t(isize_infos) *iinfos)
{
	const size_t B = batch_records();
	pipe1_t P;
	double t_toseq = 0, t_init;
	bwa_seq_t *flat = (bwa_seq_t *)malloc(2 * B * sizeof(bwa_seq_t));
	pthread_t init_th, read_th, pos_th, store_th, destroy_th;
	unsigned q;
	int s;
	memset(&P, 0, sizeof(P));
	pthread_mutex_init(&P.mu, 0); pthread_cond_init(&P.cv, 0);
	P.B = B; P.ks = ks; P.temporary = temporary; P.iinfos = iinfos; P.t0 = now();
	for (s = 0; s < P1_SLOTS; ++s) P.recs[s] = (bam_pair_t *)calloc(B, sizeof(bam_pair_t));
	g_mt.cap = memtemp_cap();
	/* device context + index upload (seconds) overlap the reading of the first batches */
	pthread_create(&init_th, 0, ensure_gpu_thread, 0);
	pthread_create(&read_th, 0, stage_read, &P);
	pthread_create(&pos_th, 0, stage_position, &P);
	pthread_create(&store_th, 0, stage_store, &P);
	pthread_create(&destroy_th, 0, stage_destroy, &P);
	pthread_join(init_th, 0);
	t_init = now() - P.t0;
	for (q = 0;; ++q) { /* the align stage */
		const int slot = (int)(q % P1_SLOTS);
		slot_wait(&P, slot, SL_READ);
		if (P.n[slot]) align_range(P.recs[slot], P.n[slot], flat, &t_toseq);
		slot_set(&P, slot, SL_ALIGNED);
		if (P.n[slot] == 0) break;
	}
	pthread_join(read_th, 0); pthread_join(pos_th, 0); pthread_join(store_th, 0); pthread_join(destroy_th, 0);
	for (s = 0; s < P1_SLOTS; ++s) free(P.recs[s]);
	free(flat);
	pthread_mutex_destroy(&P.mu); pthread_cond_destroy(&P.cv);
	fprintf(stderr, "[%s] %zu records (%.0f MB) kept in memory for pass 2%s\n", __func__, g_mt.n_rec, g_mt.bytes / 1048576.0,
	        g_mt.spilled ? ", the rest in the temporary file" : "");
	fprintf(stderr, "[%s] %ld sequences processed in %.2f sec (pipelined stages, busy seconds each: device init %.2f, read %.2f, bam1_to_seq %.2f, aln2seq/posn/isize %.2f, device calls %.2f, temp write %.2f, destroy %.2f)\n",
	        __func__, P.tot_seqs, now() - P.t0, t_init, P.t_read, t_toseq, P.t_host, g_t_aln + g_t_sa, P.t_write, P.t_destroy);
	fprintf(stderr, "[%s] finished cleanly.\n", __func__);
}

/* ------------------------------------------------------------------ pass 2 */
typedef struct { size_t n, m; uint8_t *a; } bits_t;
static void bits_push(bits_t *b, int v)
{
	if (b->n == b->m) { b->m = b->m ? b->m << 1 : 1 << 12; b->a = (uint8_t *)realloc(b->a, b->m); }
	b->a[b->n++] = (uint8_t)v;
}

static const isize_info_t *ii_of(khash_t(isize_infos) *iinfos, const bam_pair_t *r) /* bam2bam.c:715-716 */
{
	khiter_t it = kh_get(isize_infos, iinfos, bam_get_rg(r->bam_rec));
	return it == kh_end(iinfos) ? &g_null_ii : &kh_val(iinfos, it);
}

static int wants_pairing(const bam_pair_t *r) /* bam2bam.c:726-735 */
{
	const bwa_seq_t *p0 = &r->bwa_seq[0], *p1 = &r->bwa_seq[1];
	long long n_occ[2];
	int j, k;
	if (!is_mapped(p0) || !is_mapped(p1)) return 0;
	for (j = 0; j < 2; ++j) {
		const bwa_seq_t *p = &r->bwa_seq[j];
		n_occ[j] = 0;
		for (k = 0; k < p->n_aln; ++k) n_occ[j] += p->aln[k].l - p->aln[k].k + 1;
	}
	return n_occ[0] <= g_pe->max_occ && n_occ[1] <= g_pe->max_occ;
}

static double g_t_stageA, g_t_stageB, g_t_stageC, g_t_stageD;
static int is_pair_job(const bam_pair_t *r) { return r->kind == proper_pair && r->phase == positioned && unique_rec(r); }

static void pair_to_seq_one(size_t i, void *ctx)
{
	bam_pair_t *r = (bam_pair_t *)ctx + i;
	int j;
	if (!is_pair_job(r)) return;
	for (j = 0; j < 2; ++j)
		if (!r->bwa_seq[j].seq) bam1_to_seq(&r->bam_rec[j], &r->bwa_seq[j], 1, g_gap->trim_qual);
}

static void finish_tail_one(size_t i, void *ctx)
{
	bam_pair_t *r = (bam_pair_t *)ctx + i;
	if (is_pair_job(r)) {
		bwa_refine_gapped(g_bns, 1, &r->bwa_seq[0], g_pac, 0);
		bwa_refine_gapped(g_bns, 1, &r->bwa_seq[1], g_pac, 0);
		bwa_update_bam1(&r->bam_rec[0], g_bns, &r->bwa_seq[0], &r->bwa_seq[1], g_gap->mode, g_gap->max_top2);
		bwa_update_bam1(&r->bam_rec[1], g_bns, &r->bwa_seq[1], &r->bwa_seq[0], g_gap->mode, g_gap->max_top2);
		bwa_free_read_seq1(&r->bwa_seq[1]);
		bwa_free_read_seq1(&r->bwa_seq[0]);
	} else if (r->kind == singleton && r->phase == positioned && unique_rec(r)) {
		bwa_seq_t *p = &r->bwa_seq[0];
		if (!p->seq) bam1_to_seq(&r->bam_rec[0], p, 1, g_gap->trim_qual);
		bwa_refine_gapped(g_bns, 1, p, g_pac, 0);
		bwa_update_bam1(&r->bam_rec[0], g_bns, p, 0, g_gap->mode, g_gap->max_top2);
		bwa_free_read_seq1(p);
	}
	if (r->kind != eof_marker && r->phase == positioned) r->phase = finished;
}

/* finish_pair (bam2bam.c:705-811) for records [lo, hi), phase by phase */
static void finish_range(bam_pair_t *recs, size_t lo, size_t hi, khash_t(isize_infos) *iinfos, uint64_t n_tot[2], uint64_t n_mapped[2],
                         kh_64_t *my_hash)
{
	static bits_t fresh; /* per wide interval visited: did THIS record create its cache entry? */
	size_t i, fpos = 0;
	int j, k;
	bwtint_t l;

	double t1 = now();
	parallel_for(hi - lo, 1024, pair_to_seq_one, recs + lo); /* finish_pair's bam1_to_seq (bam2bam.c:717-718) */
	/* A: every SA row whose coordinate pairing will want (bam2bam.c:736-765), in the reference's visiting order */
	g_q.n = 0; fresh.n = 0;
	for (i = lo; i < hi; ++i) {
		bam_pair_t *r = &recs[i];
		if (!is_pair_job(r)) continue;
		for (j = 0; j < 2; ++j)
			if (!r->bwa_seq[j].seq) bam1_to_seq(&r->bam_rec[j], &r->bwa_seq[j], 1, g_gap->trim_qual);
		if (!wants_pairing(r)) continue;
		for (j = 0; j < 2; ++j)
			for (k = 0; k < r->bwa_seq[j].n_aln; ++k) {
				const bwt_aln1_t *a = r->bwa_seq[j].aln + k;
				int is_new = 1;
				if (a->l - a->k + 1 >= MIN_HASH_WIDTH) { /* cached by (k,l) only; the first record to touch a key fills it */
					int ret;
					khint_t it = kh_put(64, my_hash, (uint64_t)a->k << 32 | a->l, &ret);
					is_new = ret != 0;
					if (is_new) {
						poslist_t *z = &kh_val(my_hash, it);
						z->n = a->l - a->k + 1;
						z->a = (bwtint_t *)malloc(sizeof(bwtint_t) * z->n);
					}
					bits_push(&fresh, is_new);
				}
				if (is_new)
					for (l = a->k; l <= a->l; ++l) { saq_push(&g_q, l, a->a); if (l == a->l) break; }
			}
	}
	g_t_stageA += now() - t1;
	saq_run(&g_q);
	t1 = now();

	/* B: pairing on the host, in record order, then the hit lists for XA (consumes drand48), collecting their rows */
	{
		static saq_t q2;
		size_t qpos = 0;
		q2.n = 0;
		for (i = lo; i < hi; ++i) {
			bam_pair_t *r = &recs[i];
			bwa_seq_t *p[2];
			pe_data_t d;
			if (!is_pair_job(r)) continue;
			p[0] = &r->bwa_seq[0]; p[1] = &r->bwa_seq[1];
			memset(&d, 0, sizeof(pe_data_t));
			for (j = 0; j < 2; ++j) { d.aln[j].a = p[j]->aln; d.aln[j].n = p[j]->n_aln; }
			if (wants_pairing(r)) {
				d.arr.n = 0;
				for (j = 0; j < 2; ++j)
					for (k = 0; k < (int)d.aln[j].n; ++k) {
						const bwt_aln1_t *a = d.aln[j].a + k;
						const bwtint_t w = a->l - a->k + 1;
						bwtint_t t;
						uint64_t x;
						if (w >= MIN_HASH_WIDTH) {
							khint_t it = kh_get(64, my_hash, (uint64_t)a->k << 32 | a->l);
							poslist_t *z = &kh_val(my_hash, it);
							if (fresh.a[fpos++]) /* this record created the entry: its strand and length define the values */
								for (t = 0; t < w; ++t, ++qpos)
									z->a[t] = a->a ? g_q.out[qpos] : g_bwt[1]->seq_len - (g_q.out[qpos] + p[j]->len);
							for (t = 0; t < (bwtint_t)z->n; ++t) {
								x = z->a[t];
								x = x << 32 | k << 1 | j;
								kv_push(uint64_t, d.arr, x);
							}
						} else
							for (t = 0; t < w; ++t, ++qpos) {
								x = a->a ? g_q.out[qpos] : g_bwt[1]->seq_len - (g_q.out[qpos] + p[j]->len);
								x = x << 32 | k << 1 | j;
								kv_push(uint64_t, d.arr, x);
							}
					}
				pairing(p, &d, g_pe, g_gap->s_mm, ii_of(iinfos, r));
			}
			if (g_pe->N_multi || g_pe->n_multi) /* bam2bam.c:771-791 */
				for (j = 0; j < 2; ++j)
					if (p[j]->type != BWA_TYPE_NO_MATCH) {
						if (!(p[j]->extra_flag & SAM_FPP) && p[1 - j]->type != BWA_TYPE_NO_MATCH)
							bwa_aln2seq_core(d.aln[j].n, d.aln[j].a, p[j], 0,
							                 p[j]->c1 + p[j]->c2 - 1 > g_pe->N_multi ? g_pe->n_multi : g_pe->N_multi);
						else bwa_aln2seq_core(d.aln[j].n, d.aln[j].a, p[j], 0, g_pe->n_multi);
						for (k = 0; k < p[j]->n_multi; ++k) saq_push(&q2, p[j]->multi[k].pos, p[j]->multi[k].strand);
					}
			kv_destroy(d.arr);
			kv_destroy(d.pos[0]); kv_destroy(d.pos[1]);
		}
		if (qpos != g_q.n || fpos != fresh.n) { fprintf(stderr, "[bwa_gpu_batch] pass-2 enumeration out of step\n"); abort(); }
		saq_run(&q2);
		qpos = 0;
		for (i = lo; i < hi; ++i) {
			bam_pair_t *r = &recs[i];
			if (!is_pair_job(r) || !(g_pe->N_multi || g_pe->n_multi)) continue;
			for (j = 0; j < 2; ++j) {
				bwa_seq_t *p = &r->bwa_seq[j];
				if (p->type == BWA_TYPE_NO_MATCH) continue;
				for (k = 0; k < p->n_multi; ++k, ++qpos) {
					bwt_multi1_t *q = p->multi + k;
					q->pos = q->strand ? q2.out[qpos] : g_bwt[1]->seq_len - (q2.out[qpos] + p->len);
				}
			}
		}
	}

	g_t_stageB += now() - t1; t1 = now();
	/* C: mate rescue.  bwa_paired_sw1 (bwape.c:519-633) decides the windows in floating point and judges the
	 * alignments; only aln_local_core moves.  First run: note the jobs.  Device.  Second run: the real one. */
	{
		uint64_t dummy_tot[2] = {0, 0}, dummy_mapped[2] = {0, 0};
		g_sw.n = 0; g_sw.sn = 0;
		g_sw_mode = SW_RECORD;
		for (i = lo; i < hi; ++i)
			if (is_pair_job(&recs[i])) {
				bwa_seq_t *p[2] = {&recs[i].bwa_seq[0], &recs[i].bwa_seq[1]};
				bwa_paired_sw1(g_bns, g_pac, p, g_pe, ii_of(iinfos, &recs[i]), dummy_tot, dummy_mapped);
			}
		g_sw_mode = SW_REAL;
		swq_run(&g_sw);
		g_sw_mode = SW_REPLAY;
		for (i = lo; i < hi; ++i)
			if (is_pair_job(&recs[i])) {
				bwa_seq_t *p[2] = {&recs[i].bwa_seq[0], &recs[i].bwa_seq[1]};
				bwa_paired_sw1(g_bns, g_pac, p, g_pe, ii_of(iinfos, &recs[i]), n_tot, n_mapped);
			}
		g_sw_mode = SW_REAL;
		if (g_sw_pos != g_sw.n) { fprintf(stderr, "[bwa_gpu_batch] %zu of %zu SW answers unused\n", g_sw.n - g_sw_pos, g_sw.n); abort(); }
	}

	g_t_stageC += now() - t1; t1 = now();
	/* D: the reference's own per-record tail (bam2bam.c:643-658, 798-810), records spread over the host threads */
	parallel_for(hi - lo, 512, finish_tail_one, recs + lo);
	g_t_stageD += now() - t1;
}

/* A batch's BAM records are laid out, deflated and written while the next batch is loaded and finished */
typedef struct { BGZF *output; bam_pair_t *recs; size_t n; double secs; } writer_t;
static void *write_batch(void *arg)
{
	writer_t *w = (writer_t *)arg;
	const double t = now();
	write_records_bam(w->output, w->recs, w->n);
	w->secs = now() - t;
	return 0;
}

void sequential_loop_pass2(gzFile temporary, BGZF *output, khash_t(isize_infos) *iinfos)
{
	const size_t B = batch_records();
	const long long max_q = getenv("BWAGPU_BATCH_SA") ? atoll(getenv("BWAGPU_BATCH_SA")) : 1ll << 25; /* SA rows per device call */
	const double t0 = now();
	double t_read = 0, t_fin = 0, t_write = 0, t_wait = 0, t1;
	uint64_t n_tot[2] = {0, 0}, n_mapped[2] = {0, 0};
	bam_pair_t *buf[2] = {(bam_pair_t *)calloc(B, sizeof(bam_pair_t)), (bam_pair_t *)calloc(B, sizeof(bam_pair_t))};
	kh_64_t *my_hash = kh_init(64);
	khiter_t it;
	long tot_seqs = 0;
	pthread_t write_th;
	writer_t wr;
	int cur = 0, writing = 0;
	ensure_gpu();
	for (;;) {
		bam_pair_t *recs = buf[cur];
		size_t n, lo;
		t1 = now();
		n = load_records(temporary, recs, B, &tot_seqs);
		t_read += now() - t1;
		if (n == 0) break;
		t1 = now();
		for (lo = 0; lo < n;) { /* sub-ranges bounded by the SA rows their hit lists expand to */
			size_t hi = lo;
			long long q = 0;
			while (hi < n && (hi == lo || q < max_q)) {
				const bam_pair_t *r = &recs[hi];
				if (is_pair_job(r) && wants_pairing(r)) {
					int j, k;
					for (j = 0; j < 2; ++j)
						for (k = 0; k < r->bwa_seq[j].n_aln; ++k) q += r->bwa_seq[j].aln[k].l - r->bwa_seq[j].aln[k].k + 1;
				}
				++hi;
			}
			finish_range(recs, lo, hi, iinfos, n_tot, n_mapped, my_hash);
			lo = hi;
		}
		t_fin += now() - t1;
		t1 = now();
		if (writing) { pthread_join(write_th, 0); t_write += wr.secs; }
		t_wait += now() - t1;
		wr.output = output; wr.recs = recs; wr.n = n;
		pthread_create(&write_th, 0, write_batch, &wr);
		writing = 1;
		cur ^= 1;
		fprintf(stderr, "[%s] %ld sequences processed in %.2f sec\n", __func__, tot_seqs, now() - t0);
	}
	t1 = now();
	if (writing) { pthread_join(write_th, 0); t_write += wr.secs; }
	t_wait += now() - t1;
	fprintf(stderr, "[%s] finish = enumerate %.2f + pairing/XA %.2f + mate rescue (host side, both runs) %.2f + refine/update %.2f + device calls\n",
	        __func__, g_t_stageA, g_t_stageB, g_t_stageC, g_t_stageD);
	fprintf(stderr, "[%s] %ld sequences processed in %.2f sec (temp read %.2f, finish %.2f, BAM write %.2f of which %.2f not hidden behind the next batch)\n"
	                "[%s] finished cleanly, shutting down.\n"
	                "[bwa_paired_sw] %lld out of %lld Q%d singletons are mated.\n"
	                "[bwa_paired_sw] %lld out of %lld Q%d discordant pairs are fixed.\n",
	        __func__, tot_seqs, now() - t0, t_read, t_fin, t_write, t_wait, __func__, (long long)n_mapped[1], (long long)n_tot[1], SW_MIN_MAPQ,
	        (long long)n_mapped[0], (long long)n_tot[0], SW_MIN_MAPQ);
	for (it = kh_begin(my_hash); it != kh_end(my_hash); ++it)
		if (kh_exist(my_hash, it)) free(kh_val(my_hash, it).a);
	kh_destroy(64, my_hash);
	free(buf[0]); free(buf[1]);
	memtemp_free();
}

/* ------------------------------------------------------------------ 0MQ worker
 * run_worker_thread (bam2bam.c:1387-1442) handles one message = one record per loop trip.  This replacement speaks
 * the same protocol on the same socket (DEALER on inproc://work_io; wire format bam2bam.c:951-1097 through the
 * reference's own msg_init_from_pair / pair_init_from_msg) but DRAINS messages into a batch, makes the batch calls
 * above, and answers every record.  It serves both `bam2bam -t 1 -p PORT` (local thread behind the multiplexor) and
 * `bwa worker -t 1` (remote process behind the streamer device, bam2bam.c:2099-2180); the mux, the reader / output
 * threads and remote CPU workers are untouched.
 *
 * What the mux's behaviour (bam2bam.c:1577-1601) demands of a batching worker:
 *   * whenever a record is un-acked and the crowd socket is writable it sends -- new records first, else it RE-SENDS
 *     outstanding ones round-robin.  So a duplicate `recno` means "nothing new right now": the batch is closed early
 *     (once it holds BWAGPU_WORKER_MIN records) instead of waiting for the deadline, and duplicates of records this
 *     worker took recently are dropped (re-sends that crossed their answer; the mux would discard a repeated answer
 *     anyway, 1610-1623);
 *   * the worker stops receiving while a batch is on the device, so the HWM of 64 (bam2bam.c:35) back-pressures
 *     the mux instead of flooding the worker.
 * First arrivals are in recno order (new records are sent in order), so with ONE batching worker drand48 is consumed
 * in record order in both passes and the BAM equals `bam2bam -t 1`'s (tests/test_batched_worker.py), which the
 * reference's own `-t N` does not guarantee.
 */

void msg_init_from_pair(zmq_msg_t *m, bam_pair_t *p);
void pair_init_from_msg(bam_pair_t *p, zmq_msg_t *m);
void pair_posn(bam_pair_t *p);
void set_sockopts(void *socket);

static void *g_zmq_ctx;              /* bam2bam.c:103 (static there): captured where it is created */
static void *volatile g_iinfos_seen; /* bam2bam.c:107 (static there): what g_iinfos is set to at 1769 / 1857 / 2092 / 2298 */

void *zmq_init(int io_threads)
{
	REAL(void *, zmq_init, int);
	return g_zmq_ctx = real_zmq_init(io_threads);
}

void infer_all_isizes(khash_t(isize_infos) *iinfos, double ap_prior, int64_t L)
{
	REAL(void, infer_all_isizes, khash_t(isize_infos) *, double, int64_t);
	real_infer_all_isizes(iinfos, ap_prior, L);
	g_iinfos_seen = iinfos; /* the caller assigns g_iinfos = iinfos right after (bam2bam.c:1768-1769, 1856-1857) */
}

khash_t(isize_infos) *decode_iinfo(char *p, char *q)
{
	REAL(khash_t(isize_infos) *, decode_iinfo, char *, char *);
	khash_t(isize_infos) *r = real_decode_iinfo(p, q);
	g_iinfos_seen = r; /* bwa worker: g_iinfos = decode_iinfo(...) (bam2bam.c:2092, 2298) */
	return r;
}

static long env_long(const char *name, long dflt)
{
	const char *e = getenv(name);
	const long v = e ? atol(e) : 0;
	return v > 0 ? v : dflt;
}

/* Which records this worker has taken: the mux keeps at most ring_size = 524288 records in flight (bam2bam.c:9), so a
 * direct-mapped table of twice that many slots, indexed by recno, never has two live records in one slot.  A record
 * taken less than BWAGPU_WORKER_REANSWER_S seconds ago is a re-send that crossed its answer and is dropped; an older
 * one is processed again (its answer may have been lost with a TCP reconnect), as the reference's worker would. */
#define TAKEN_SLOTS (1u << 20)
typedef struct { uint64_t recno_p1; float t; uint8_t phase; } taken_t;

void *run_worker_thread(void *arg)
{
	const size_t B = (size_t)env_long("BWAGPU_WORKER_RECORDS", 1 << 17);
	const size_t min_batch = (size_t)env_long("BWAGPU_WORKER_MIN", 1024);
	const double wait_s = 1e-3 * (double)env_long("BWAGPU_WORKER_WAIT_MS", 50);
	const long long max_q = getenv("BWAGPU_BATCH_SA") ? atoll(getenv("BWAGPU_BATCH_SA")) : 1ll << 25;
	bam_pair_t *recs = (bam_pair_t *)calloc(B, sizeof(bam_pair_t));
	bwa_seq_t *flat = (bwa_seq_t *)malloc(2 * B * sizeof(bwa_seq_t));
	const double reanswer_s = (double)env_long("BWAGPU_WORKER_REANSWER_S", 30);
	const double t_start = now();
	taken_t *taken = (taken_t *)calloc(TAKEN_SLOTS, sizeof(taken_t));
	kh_64_t *my_hash = kh_init(64);
	khiter_t it;
	uint64_t n_tot[2] = {0, 0}, n_mapped[2] = {0, 0};
	long n_batches = 0, n_records = 0, n_dupes = 0, failure_count = 0;
	double t_toseq = 0, t_host = 0, t_fin = 0, t_wait = 0, t_send = 0;
	int done = 0, have_pending = 0;
	bam_pair_t pending;
	void *upstream;
	static int claimed;
	(void)arg;

	/* the batch state above (SA / SW queues, record-replay cursors) is one set of globals and there is one device
	 * queue: ONE thread batches, whatever -t says; further threads have nothing to add and leave */
	if (__sync_lock_test_and_set(&claimed, 1)) {
		fprintf(stderr, "[run_worker_thread] a batching GPU worker is already running in this process; extra thread exits.\n");
		free(recs); free(flat); free(taken); kh_destroy(64, my_hash);
		return 0;
	}
	if (!g_zmq_ctx) { fprintf(stderr, "[bwa_gpu_batch] run_worker_thread: the 0MQ context was not seen being created\n"); abort(); }
	upstream = zmq_socket(g_zmq_ctx, ZMQ_DEALER);
	if (!upstream) { fprintf(stderr, "[bwa_gpu_batch] error creating socket\n"); abort(); }
	set_sockopts(upstream);
	if (zmq_connect(upstream, "inproc://work_io") != 0) { fprintf(stderr, "[bwa_gpu_batch] zmq_connect failed: %s\n", zmq_strerror(zmq_errno())); exit(1); }
	ensure_gpu();

	while (!done) {
		size_t n = 0, i, lo;
		int saw_dupe = 0, phase_of_batch = -1, ret;
		double t_first = 0, t1 = now();
		/* ---- drain messages into a batch */
		if (have_pending) { /* the record that closed the previous batch (it was in another phase) opens this one */
			recs[0] = pending; have_pending = 0;
			phase_of_batch = (int)recs[0].phase;
			n = 1; t_first = now();
		}
		for (;;) {
			zmq_msg_t msg;
			bam_pair_t *r = &recs[n];
			taken_t *tk;
			if (n > 0) { /* the first message is waited for indefinitely, later ones until the deadline */
				zmq_pollitem_t item = {upstream, 0, ZMQ_POLLIN, 0};
				const double left = t_first + wait_s - now();
				if (left <= 0 || (saw_dupe && n >= min_batch)) break;
				ret = zmq_poll(&item, 1, (long)(left * 1e3) + 1);
				if (ret < 0) { if (zmq_errno() == ETERM || zmq_errno() == EINTR) done = 1; break; }
				if (ret == 0) break;
			}
			zmq_msg_init(&msg);
			if (zmq_msg_recv(&msg, upstream, 0) < 0) {
				zmq_msg_close(&msg);
				if (zmq_errno() == ETERM || zmq_errno() == EINTR) { done = 1; break; } /* clean exit, as bam2bam.c:1406 */
				fprintf(stderr, "zmq_msg_recv failed: %s\n", zmq_strerror(zmq_errno()));
				exit(1);
			}
			if (n == 0) t_first = now();
			pair_init_from_msg(r, &msg);
			zmq_msg_close(&msg);
			tk = &taken[r->recno & (TAKEN_SLOTS - 1)];
			if (tk->recno_p1 == r->recno + 1 && tk->phase == (uint8_t)r->phase && (now() - t_start) - tk->t < reanswer_s) {
				saw_dupe = 1; ++n_dupes; bam_destroy_pair(r); continue;
			}
			tk->recno_p1 = r->recno + 1; tk->phase = (uint8_t)r->phase; tk->t = (float)(now() - t_start);
			if (phase_of_batch >= 0 && (int)r->phase != phase_of_batch) { /* a batch holds one phase: this record opens the next */
				pending = *r; have_pending = 1;
				memset(r, 0, sizeof(*r));
				break;
			}
			phase_of_batch = (int)r->phase;
			if (++n == B) break;
		}
		t_wait += now() - t1;
		if (n == 0) continue;
		++n_batches; n_records += (long)n;

		/* ---- the switch of bam2bam.c:1414-1422, for the whole batch */
		t1 = now();
		if (phase_of_batch == pristine) align_position_range(recs, n, flat, &t_toseq, &t_host);
		else if (phase_of_batch == aligned) for (i = 0; i < n; ++i) pair_posn(&recs[i]);
		else if (phase_of_batch == positioned) {
			khash_t(isize_infos) *iinfos = (khash_t(isize_infos) *)g_iinfos_seen;
			if (!iinfos) failure_count += (long)n; /* sent back as they came (bam2bam.c:1419-1420) */
			else
				for (lo = 0; lo < n;) { /* sub-ranges bounded by the SA rows their hit lists expand to, as sequential_loop_pass2 */
					size_t hi = lo;
					long long q = 0;
					while (hi < n && (hi == lo || q < max_q)) {
						const bam_pair_t *r = &recs[hi];
						if (is_pair_job(r) && wants_pairing(r)) {
							int j, k;
							for (j = 0; j < 2; ++j)
								for (k = 0; k < r->bwa_seq[j].n_aln; ++k) q += r->bwa_seq[j].aln[k].l - r->bwa_seq[j].aln[k].k + 1;
						}
						++hi;
					}
					finish_range(recs, lo, hi, iinfos, n_tot, n_mapped, my_hash);
					lo = hi;
				}
		}
		t_fin += now() - t1;

		/* ---- answer every record */
		t1 = now();
		for (i = 0; i < n; ++i) {
			zmq_msg_t msg;
			msg_init_from_pair(&msg, &recs[i]);
			bam_destroy_pair(&recs[i]);
			ret = zmq_msg_send(&msg, upstream, 0);
			zmq_msg_close(&msg);
			if (ret < 0) {
				if (zmq_errno() == ETERM || zmq_errno() == EINTR) { done = 1; for (++i; i < n; ++i) bam_destroy_pair(&recs[i]); break; }
				fprintf(stderr, "zmq_msg_send failed: %s\n", zmq_strerror(zmq_errno()));
				exit(1);
			}
		}
		t_send += now() - t1;
		if (failure_count >= 1024) {
			fprintf(stderr, "[run_worker_thread] Lots of failures due to missing insert size information.\n");
			fprintf(stderr, "[run_worker_thread] Terminating due to suspected communication problem.\n");
			raise(SIGINT); /* the reference sets its static s_interrupted; its own handler does the same (bam2bam.c:129-132) */
			break;
		}
	}
	fprintf(stderr, "[run_worker_thread] exiting: %ld batches, %ld records, %ld duplicates dropped; wait/recv %.2f s, bam1_to_seq %.2f, "
	                "host phases %.2f, batch work incl. device %.2f, send %.2f\n", n_batches, n_records, n_dupes, t_wait, t_toseq, t_host, t_fin, t_send);
	for (it = kh_begin(my_hash); it != kh_end(my_hash); ++it)
		if (kh_exist(my_hash, it)) free(kh_val(my_hash, it).a);
	kh_destroy(64, my_hash);
	free(recs); free(flat); free(taken);
	zmq_close(upstream);
	return 0;
}
