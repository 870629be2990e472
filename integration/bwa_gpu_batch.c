/* bwa_gpu_batch.c -- the BATCHED drop-in: `bwa bam2bam -t 1` (and the 0MQ worker) with the hot path on the GPU, one device
 * call per phase and batch instead of one per record.
 *
 * SURVEY.md §8(f) rank 1 / INTEGRATION.md: the reference's sequential driver (bam2bam.c:1143-1219) handles one record at
 * a time -- read_bam_pair -> pair_aln -> pair_posn -> ... -- so its calls into the alignment layer carry one read each.
 * This library REPLACES the two loop functions
 *
 *     sequential_loop_pass1 (bam2bam.c:1143)      sequential_loop_pass2 (bam2bam.c:1178)
 *
 * and the worker thread run_worker_thread (bam2bam.c:1387) -- plain global functions of the reference, reached through the
 * PLT of a shared-library build of it (integration/Makefile: _host/libbwahost.so) -- with versions that gather a batch of
 * records and make ONE call per phase:
 *
 *   pass 1   bam1_to_seq x n  ->  bwa_gpu_cal_sa_reads_gap   (replaces bwa_cal_sa_reg_gap, bam2bam.c:616,676)
 *            bwa_aln2seq[_core] in record order (it consumes drand48: bwase.c:33,36,78)
 *            -> bwa_gpu_cal_pac_pos                           (replaces bwt_sa at bwase.c:145,152; bam2bam.c:635-636)
 *   pass 2   hit enumeration of all pairs -> bwa_gpu_cal_pac_pos   (bwt_sa at bam2bam.c:752,761)
 *            pairing; bwa_aln2seq_core in record order -> bwa_gpu_cal_pac_pos (bam2bam.c:786)
 *            bwa_paired_sw1 -> bwa_gpu_mate_sw_path           (aln_local_core inside bwa_sw_core, bwape.c:456)
 *            bwa_refine_gapped -> bwa_gpu_global_align_seqs   (aln_global_core inside refine_gapped_core, bwase.c:212)
 *            bwa_update_bam1, BAM output: the reference's own code per record, spread over the host threads.
 *
 * Everything that is not the hot path IS the reference: this file calls its exported functions (read_bam_pair,
 * bam1_to_seq, bwa_aln2seq_core, bwa_cal_pac_pos_core, pairing, bwa_paired_sw1, bwa_refine_gapped, bwa_update_bam1, ...)
 * and never restates their arithmetic.  Where a reference function interleaves host logic with a hot call
 * (bwa_cal_pac_pos_core around bwt_sa, bwa_paired_sw1/bwa_sw_core around aln_local_core, bwa_refine_gapped around
 * aln_global_core) it is run in RECORD / REPLAY fashion: a first run with the hot call interposed to note its arguments,
 * one device call for the whole batch, then the real run with the hot call interposed to hand back the device's answers
 * in the same order.
 *
 * Both passes are pipelines of stages, each stage a thread, batches handed on in order; the stages that are functions of
 * one record (pairing, mate rescue, refinement, BAM record rewrite) run on all host threads.  Order-sensitive state is
 * kept exactly as `-t 1` has it: drand48 is consumed by bwa_aln2seq_core only (bwase.c), which runs in record order in
 * both passes; the pass-2 position cache for intervals >= 1000 wide (bam2bam.c:743-758) is filled by the first record, in
 * record order, that touches a key.  The output BAM is therefore record-identical to `bam2bam -t 1`
 * (tests/test_batched_bam2bam.py).
 *
 * There is no CPU path in here: `.sai` side inputs (-0/-1/-2), which would bypass the device search, are refused.
 *
 * Two ways in:
 *   LD_PRELOAD=integration/libbwa_gpu_batch.so integration/_host/bwa_host bam2bam -g idx -t 1 -f out.bam in.bam
 *   dlopen("libbwa_gpu_batch.so") and call bwa_bam_to_bam() in-process (bench.py; the library links libbwahost.so itself).
 * No libc / zlib symbol is interposed -- only functions of the reference -- so both ways behave the same.
 */
#include "shim.h"
#include <getopt.h>
#include <signal.h>
#include <sys/resource.h>
#include <time.h>
#include <unistd.h>
#include "bwa_gpu_batch.h"

KHASH_MAP_INIT_INT64(64, poslist_t) /* the position cache's type, as bam2bam.c:38 declares it */

/* reference functions this driver calls that no header declares (bam2bam.c) */
void pair_aln(bam_pair_t *p);
void pair_posn(bam_pair_t *p);
void pair_print_custom(gzFile f, bam_pair_t *p);
int read_pair_custom(gzFile f, bam_pair_t *p);
void pair_print_bam(BGZF *output, bam_pair_t *p);
void bwa_update_bam1(bam1_t *out, const bntseq_t *bns, bwa_seq_t *p, const bwa_seq_t *mate, int mode, int max_top2);
void bwa_cal_pac_pos_core(const bwt_t *forward_bwt, const bwt_t *reverse_bwt, bwa_seq_t *seq, const int max_mm, const float fnr);
void bwa_aln2seq(int n_aln, const bwt_aln1_t *aln, bwa_seq_t *s);
void msg_init_from_pair(zmq_msg_t *m, bam_pair_t *p);
void pair_init_from_msg(bam_pair_t *p, zmq_msg_t *m);
void set_sockopts(void *socket);
extern struct option longopts[]; /* bam2bam.c:40 */

static void die(const char *what)
{
	fprintf(stderr, "[bwa_gpu_batch] %s: %s\n", what, bwa_gpu_last_error());
	abort(); /* the reference's convention on this path: xassert -> abort (utils.c:68-83) */
}

#define now shim_now

/* ------------------------------------------------------------------ host threads
 * With the hot path on the device, what is left of a bam2bam run is per-record host work of the reference (bam1_to_seq,
 * pairing, bwa_paired_sw1, bwa_refine_gapped, bwa_update_bam1, the intermediate-record codec, deflate).  The reference
 * itself runs these functions concurrently in its worker threads (run_worker_thread, bam2bam.c:1387), so they are
 * re-entrant; the shim spreads each such phase of a batch over BWAGPU_SHIM_THREADS threads (default: the host's cores, at
 * most 64).  Everything order-sensitive (drand48 in bwa_aln2seq_core, the position cache, isize statistics) stays serial. */
typedef struct { size_t n, grain; volatile size_t next; pf_fn fn; void *ctx; int bucket; int helpers; } pf_job_t;
#define MAX_THREADS 64

__thread int t_cpu_bucket = CPU_OTHER;
static int64_t g_cpu_ns[CPU_N];
static const char *const g_cpu_name[CPU_N] = {"inflate", "parse", "bam1_to_seq", "aln2seq (serial)", "cal_pac_pos_core", "isize", "store/encode", "destroy",
	"load/decode", "enumerate", "pairing", "XA aln2seq", "rescue record", "rescue replay", "refine record", "refine replay", "update_bam1",
	"BAM layout", "deflate", "fwrite", "other"};

double thread_cpu_now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void cpu_add(int bucket, double seconds) { __sync_fetch_and_add(&g_cpu_ns[bucket], (int64_t)(seconds * 1e9)); }

static void cpu_report(const char *what)
{
	char line[2048];
	int b, o = 0;
	double tot = 0;
	for (b = 0; b < CPU_N; ++b) tot += 1e-9 * (double)g_cpu_ns[b];
	o += snprintf(line + o, sizeof(line) - (size_t)o, "[bwa_gpu_batch] host CPU seconds by activity (%s, %.2f in all):", what, tot);
	for (b = 0; b < CPU_N; ++b)
		if (g_cpu_ns[b]) o += snprintf(line + o, sizeof(line) - (size_t)o, " %s %.2f;", g_cpu_name[b], 1e-9 * (double)g_cpu_ns[b]);
	fprintf(stderr, "%s\n", line);
}

int shim_threads(void)
{
	static int n;
	if (!n) {
		const char *e = getenv("BWAGPU_SHIM_THREADS");
		n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
		if (n < 1) n = 1;
		if (n > MAX_THREADS) n = MAX_THREADS;
	}
	return n;
}

/* Threads are made per call on purpose.  A persistent pool was tried (round 2): the same loops then cost 50 % more CPU
 * (refine/update 4.2 -> 6.3 CPU-seconds per 2 M reads, pass 2 0.67 -> 0.98 s) -- long-lived workers keep allocating from
 * arenas that the serial destroy stages are freeing into, whereas a fresh thread attaches to an idle arena. */
static void *pf_worker(void *arg)
{
	pf_job_t *j = (pf_job_t *)arg;
	const double c0 = thread_cpu_now();
	for (;;) {
		const size_t lo = __sync_fetch_and_add(&j->next, j->grain);
		size_t i, hi;
		if (lo >= j->n) break;
		hi = lo + j->grain < j->n ? lo + j->grain : j->n;
		for (i = lo; i < hi; ++i) j->fn(i, j->ctx);
	}
	cpu_add(j->bucket, thread_cpu_now() - c0);
	return 0;
}

void parallel_for(size_t n, size_t grain, pf_fn fn, void *ctx)
{
	pf_job_t j = {n, grain ? grain : 1, 0, fn, ctx, t_cpu_bucket, 0};
	pthread_t th[MAX_THREADS];
	int t, nt = shim_threads();
	if (n == 0) return;
	if ((size_t)nt > (n + j.grain - 1) / j.grain) nt = (int)((n + j.grain - 1) / j.grain);
	if (nt <= 1) { pf_worker(&j); return; }
	for (t = 1; t < nt; ++t) pthread_create(&th[t], 0, pf_worker, &j);
	pf_worker(&j);
	for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

int slice_count(size_t n, size_t min_per_slice)
{
	size_t nt = (size_t)shim_threads();
	if (min_per_slice < 1) min_per_slice = 1;
	if (nt > (n + min_per_slice - 1) / min_per_slice) nt = (n + min_per_slice - 1) / min_per_slice;
	return nt < 1 ? 1 : (int)nt;
}

/* n items cut into slice_count() contiguous slices, fn(slice, lo, hi) once per slice on the pool; returns the slice count */
typedef struct { ps_fn fn; void *ctx; size_t n; int nt; } ps_arg_t;
static void ps_one(size_t s, void *arg)
{
	const ps_arg_t *a = (const ps_arg_t *)arg;
	a->fn((int)s, a->n * s / (size_t)a->nt, a->n * (s + 1) / (size_t)a->nt, a->ctx);
}

int parallel_slices(size_t n, size_t min_per_slice, ps_fn fn, void *ctx)
{
	ps_arg_t a = {fn, ctx, n, slice_count(n, min_per_slice)};
	parallel_for((size_t)a.nt, 1, ps_one, &a);
	return a.nt;
}

/* ------------------------------------------------------------------ the reference's static globals, captured */
static bwt_t *g_bwt[2];        /* bam2bam.c:88  (init_genome_index 855-856) */
static const bntseq_t *g_bns;  /* bam2bam.c:89 */
static ubyte_t *g_pac;         /* bam2bam.c:91 */
static gap_opt_t *g_gap;       /* bam2bam.c:94 */
static pe_opt_t *g_pe;         /* bam2bam.c:95 */
int g_broken_input, g_skip_duplicates, g_drop_aligned, g_only_aligned; /* bam2bam.c:96-101, options 130 / 131 / 133 / 128 */
static isize_info_t g_null_ii; /* bam2bam.c:106 */
static bwa_gpu_batch_report_t g_rep; /* the run in progress / the last run (bwa_gpu_batch.h) */
static double g_cpu0;
static int g_trace, g_trace_pass; /* BWAGPU_TRACE=1: every stage hand-over on stderr (seconds since the call, pass, slot, new state) */
static double g_t_call, g_t_pass1_begin, g_t_pass1_end, g_t_pass2_begin, g_t_pass2_end; /* where a run's wall time goes outside the passes */
void shim_trace_pt(const char *what) { if (g_trace) fprintf(stderr, "[trace] %.4f %s\n", now() - g_t_call, what); }

/* The index of a long-lived host process (bench.py runs bam2bam several times in one process): with keep_index on, the
 * loaders hand back what an earlier run loaded from the same files, the matching destroy calls leave it alone, and the
 * device copy stays valid.  Off (the default), every run loads and frees like the reference. */
static int g_keep;
static struct { char *fn; bwt_t *b; } g_kept_bwt[2];
static struct { ubyte_t *pac; int64_t l_pac; } g_kept_pac;
static int g_ready; /* the device holds g_bwt / g_pac */
static bwt_t *g_dev_bwt[2];
static ubyte_t *g_dev_pac;

void bwa_gpu_batch_keep_index(int on) { g_keep = on != 0; }

static int is_kept_bwt(const bwt_t *b) { return b && (b == g_kept_bwt[0].b || b == g_kept_bwt[1].b); }

bwt_t *bwt_restore_bwt(const char *fn, int touch)
{
	REAL(bwt_t *, bwt_restore_bwt, const char *, int);
	const size_t n = strlen(fn);
	const int s = n >= 5 && strcmp(fn + n - 5, ".rbwt") == 0 ? 1 : 0;
	const double t0 = now();
	shim_trace_pt("bwt_restore_bwt");
	bwt_t *b;
	if (g_keep && g_kept_bwt[s].b && strcmp(g_kept_bwt[s].fn, fn) == 0) b = g_kept_bwt[s].b;
	else {
		b = real_bwt_restore_bwt(fn, touch);
		if (g_keep) {
			if (g_kept_bwt[s].b) { /* another genome: the old one goes */
				REAL(void, bwt_destroy, bwt_t *);
				bwt_t *old = g_kept_bwt[s].b;
				g_kept_bwt[s].b = 0;
				real_bwt_destroy(old);
				free(g_kept_bwt[s].fn);
			}
			g_kept_bwt[s].b = b; g_kept_bwt[s].fn = strdup(fn);
			if (g_kept_pac.pac) { free(g_kept_pac.pac); g_kept_pac.pac = 0; }
		}
	}
	g_bwt[s] = b;
	g_rep.index_load_s += now() - t0;
	return b;
}

void bwt_restore_sa(const char *fn, bwt_t *bwt, int touch)
{
	REAL(void, bwt_restore_sa, const char *, bwt_t *, int);
	const double t0 = now();
	if (!(is_kept_bwt(bwt) && bwt->sa)) real_bwt_restore_sa(fn, bwt, touch);
	g_rep.index_load_s += now() - t0;
}

ubyte_t *bwt_restore_pac(const bntseq_t *bns, int touch)
{
	REAL(ubyte_t *, bwt_restore_pac, const bntseq_t *, int);
	const double t0 = now();
	shim_trace_pt("bwt_restore_pac");
	g_bns = bns;
	if (g_keep && g_kept_pac.pac && g_kept_pac.l_pac == (int64_t)bns->l_pac) g_pac = g_kept_pac.pac;
	else {
		g_pac = real_bwt_restore_pac(bns, touch);
		if (g_keep) { g_kept_pac.pac = g_pac; g_kept_pac.l_pac = (int64_t)bns->l_pac; }
	}
	g_rep.index_load_s += now() - t0;
	return g_pac;
}

void bwt_destroy(bwt_t *bwt)
{
	REAL(void, bwt_destroy, bwt_t *);
	if (is_kept_bwt(bwt)) return;
	if (bwt && (bwt == g_dev_bwt[0] || bwt == g_dev_bwt[1])) g_ready = 0; /* the device copy no longer has a host twin to be matched against */
	real_bwt_destroy(bwt);
}

void bwt_destroy_pac(ubyte_t *pac, const bntseq_t *bns)
{
	REAL(void, bwt_destroy_pac, ubyte_t *, const bntseq_t *);
	if (pac && pac == g_kept_pac.pac) return;
	real_bwt_destroy_pac(pac, bns);
}

/* the device context is torn down and set up again by the next run: the library reads its BWAGPU_* settings when it is
 * initialised (bench.py measures the same job in two configurations of one process) */
void bwa_gpu_batch_reset_device(void)
{
	if (g_ready) { bwa_gpu_destroy(); g_ready = 0; }
}

void bwa_gpu_batch_drop_index(void)
{
	REAL(void, bwt_destroy, bwt_t *);
	int s;
	for (s = 0; s < 2; ++s)
		if (g_kept_bwt[s].b) {
			bwt_t *b = g_kept_bwt[s].b;
			g_kept_bwt[s].b = 0;
			real_bwt_destroy(b);
			free(g_kept_bwt[s].fn); g_kept_bwt[s].fn = 0;
		}
	free(g_kept_pac.pac); g_kept_pac.pac = 0;
	memtemp_release();
	if (g_ready) { bwa_gpu_destroy(); g_ready = 0; }
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

/* bwa_bam_to_bam (bam2bam.c:1942): the reference keeps four of its long options in statics nobody can see; parse the same
 * command line with the reference's own option table first, then hand over.  Also the clock of the whole run. */
static void env_defaults(void);
int bwa_bam_to_bam(int argc, char *argv[], char *version)
{
	REAL(int, bwa_bam_to_bam, int, char **, char *);
	char **av = (char **)malloc(((size_t)argc + 1) * sizeof(char *));
	int c, rc;
	double t0;
	memcpy(av, argv, (size_t)argc * sizeof(char *));
	av[argc] = 0;
	env_defaults(); /* no stage thread exists yet */
	{ const char *e = getenv("BWAGPU_TRACE"); g_trace = e && atoi(e) != 0; }
	g_only_aligned = g_broken_input = g_skip_duplicates = g_drop_aligned = 0;
	memset(&g_rep, 0, sizeof(g_rep));
	memset(g_cpu_ns, 0, sizeof(g_cpu_ns));
	optind = 0; opterr = 0; /* glibc: a full re-initialisation, quietly -- the real parse below reports what is wrong */
	while ((c = getopt_long(argc, av, "g:n:o:e:i:d:l:k:LR:m:t:NM:O:E:q:f:C:D:a:sc:h:H:Ap:0:1:2:", longopts, 0)) >= 0) {
		if (c == 128) g_only_aligned = 1;      /* bam2bam.c:1988 */
		if (c == 130) g_broken_input = 1;      /* bam2bam.c:1991 */
		if (c == 131) g_skip_duplicates = 1;
		if (c == 133) g_drop_aligned = 1;
	}
	free(av);
	optind = 0; opterr = 1;
	{
		struct rusage ru;
		getrusage(RUSAGE_SELF, &ru);
		g_cpu0 = (double)ru.ru_utime.tv_sec + 1e-6 * (double)ru.ru_utime.tv_usec + (double)ru.ru_stime.tv_sec + 1e-6 * (double)ru.ru_stime.tv_usec;
	}
	t0 = now();
	g_t_call = t0;
	rc = real_bwa_bam_to_bam(argc, argv, version);
	shim_trace_pt("bwa_bam_to_bam returned");
	g_rep.wall_s = now() - t0;
	fprintf(stderr, "[bwa_gpu_batch] outside the passes: %.3f s before pass 1 (options, header, index), %.3f s between the passes (isize inference), %.3f s after pass 2 (close)\n",
	        g_t_pass1_begin - g_t_call, g_t_pass2_begin - g_t_pass1_end, now() - g_t_pass2_end);
	g_rep.inflate_cpu_s = 1e-9 * (double)g_cpu_ns[CPU_INFLATE];
	{
		struct rusage ru;
		getrusage(RUSAGE_SELF, &ru);
		g_rep.process_cpu_s = (double)ru.ru_utime.tv_sec + 1e-6 * (double)ru.ru_utime.tv_usec + (double)ru.ru_stime.tv_sec + 1e-6 * (double)ru.ru_stime.tv_usec - g_cpu0;
	}
	cpu_report("whole run");
	return rc;
}

/* bwa_bam_open (bwaseqio.c:22-64): note which file the records will come from (shim_io.c inflates its blocks in
 * parallel); `.sai` side inputs would put alignments made elsewhere in place of the device search: refused. */
bwa_seqio_t *bwa_bam_open(const char *fn, int which, char **saif, gap_opt_t *o0, bam_header_t **hh)
{
	REAL(bwa_seqio_t *, bwa_bam_open, const char *, int, char **, gap_opt_t *, bam_header_t **);
	if (saif && (saif[0] || saif[1] || saif[2])) {
		fprintf(stderr, "[bwa_gpu_batch] .sai inputs (-0/-1/-2) bypass the alignment this library exists to run on the device; "
		                "they are not supported here (run the plain reference for them)\n");
		exit(1);
	}
	fastin_set_path(fn);
	shim_trace_pt("bwa_bam_open");
	{
		bwa_seqio_t *r = real_bwa_bam_open(fn, which, saif, o0, hh);
		shim_trace_pt("bwa_bam_open done");
		return r;
	}
}

void bwa_seq_close(bwa_seqio_t *bs)
{
	REAL(void, bwa_seq_close, bwa_seqio_t *);
	shim_trace_pt("bwa_seq_close");
	fastin_close();
	shim_trace_pt("fastin_close done");
	real_bwa_seq_close(bs);
	shim_trace_pt("bwa_seq_close done");
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

/* ------------------------------------------------------------------ device context */
static pthread_mutex_t g_rep_mu = PTHREAD_MUTEX_INITIALIZER;
#define REP_ADD(calls_f, units_f, secs_f, units, t0) do { pthread_mutex_lock(&g_rep_mu); ++g_rep.calls_f; g_rep.units_f += (int64_t)(units); \
		g_rep.secs_f += now() - (t0); pthread_mutex_unlock(&g_rep_mu); } while (0)

int shim_device_ready(void) { return __atomic_load_n(&g_ready, __ATOMIC_ACQUIRE); }

void shim_count_bgzf(int64_t bytes, double seconds)
{
	pthread_mutex_lock(&g_rep_mu); ++g_rep.calls_bgzf; g_rep.bytes_bgzf += bytes; g_rep.dev_bgzf_s += seconds; pthread_mutex_unlock(&g_rep_mu);
}

static void report(void)
{
	fprintf(stderr, "[bwa_gpu_batch] device calls: cal_sa_reads_gap=%ld (%ld reads, %.2f s)  cal_pac_pos=%ld (%ld queries, %.2f s)  "
	                "mate_sw_path=%ld (%ld jobs, %.2f s)  global_align=%ld (%ld jobs, %.2f s)  bgzf_deflate=%ld (%ld bytes, %.2f s)\n", (long)g_rep.calls_aln, (long)g_rep.reads_aln,
	        g_rep.dev_aln_s, (long)g_rep.calls_sa, (long)g_rep.q_sa, g_rep.dev_sa_s, (long)g_rep.calls_sw, (long)g_rep.jobs_sw, g_rep.dev_sw_s,
	        (long)g_rep.calls_ga, (long)g_rep.jobs_ga, g_rep.dev_ga_s, (long)g_rep.calls_bgzf, (long)g_rep.bytes_bgzf, g_rep.dev_bgzf_s);
	{ /* the kernels inside those calls (CUDA events in the library; kernels of different lanes overlap) */
		bwa_gpu_totals_t t;
		if (bwa_gpu_get_totals(&t) == 0)
			fprintf(stderr, "[bwa_gpu_batch] kernel ms since the device was set up: widths %.1f, search %.1f (passes %.1f / %.1f / %.1f), SA %.1f, SW %.1f, "
			                "global %.1f, BGZF deflate %.1f, inflate %.1f; %ld launches; H2D %.0f MB, D2H %.0f MB\n", t.ms_width, t.ms_search,
			        t.ms_search_pass[0], t.ms_search_pass[1], t.ms_search_pass[2], t.ms_sa, t.ms_sw, t.ms_global, t.ms_bgzf, t.ms_inflate,
			        (long)t.launches, (double)t.h2d_bytes / 1e6, (double)t.d2h_bytes / 1e6);
	}
}

/* The library's configuration this host wants, unless the user said otherwise.  setenv() may move `environ`, so it must not
 * run beside a getenv() of another thread (ensure_gpu runs on a thread of its own next to the reader, whose getenv() calls
 * crashed once in ~60 loaded runs when the setenv()s were there): it runs when the shim is loaded and at the top of every
 * bwa_bam_to_bam, before any stage thread exists. */
__attribute__((constructor)) static void env_defaults(void)
{
	setenv("BWAGPU_LANES", "4", 0);       /* two search calls in flight (the two align threads of pass 1), two lanes each */
	setenv("BWAGPU_CALL_GROUPS", "2", 0);
	setenv("BWAGPU_MALLOPT", "1", 0); /* this host frees millions of aln[] per batch: keep the heaps (a process-wide choice, ours to make) */
}

static void ensure_gpu(void)
{
	static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	static int exit_hook;
	const double t0 = now();
	pthread_mutex_lock(&mu);
	if (g_ready && g_dev_bwt[0] == g_bwt[0] && g_dev_bwt[1] == g_bwt[1] && g_dev_pac == g_pac) { pthread_mutex_unlock(&mu); return; }
	if (!g_bwt[0] || !g_bwt[1] || !g_bns || !g_pac || !g_gap || !g_pe) {
		fprintf(stderr, "[bwa_gpu_batch] the reference's index/option globals were not seen being loaded\n");
		abort();
	}
	{
		const char *e = getenv("BWAGPU_NDEV"), *f = getenv("BWAGPU_DEVICE");
		int ids[16], n = e ? atoi(e) : 1, first = f ? atoi(f) : 0, i;
		if (n < 1) n = 1;
		if (n > 16) n = 16;
		for (i = 0; i < n; ++i) ids[i] = first + i;
		if (bwa_gpu_init(n, ids)) die("bwa_gpu_init");
	}
	if (bwa_gpu_load_index(g_bwt, g_pac, g_bns->l_pac)) die("bwa_gpu_load_index");
	g_dev_bwt[0] = g_bwt[0]; g_dev_bwt[1] = g_bwt[1]; g_dev_pac = g_pac;
	g_ready = 1;
	if (!exit_hook) { exit_hook = 1; atexit(report); }
	g_rep.device_init_s += now() - t0;
	pthread_mutex_unlock(&mu);
}
static void *ensure_gpu_thread(void *arg) { (void)arg; ensure_gpu(); return 0; }

/* The first batches of a pass are small and double up to the full size: the pipeline's stages all have work after a
 * fraction of a full batch's latency (BWAGPU_BATCH_RAMP=0: every batch full size). */
static size_t ramp_records(size_t B, unsigned q)
{
	const char *e = getenv("BWAGPU_BATCH_RAMP");
	size_t b;
	if ((e && atoi(e) == 0) || q >= 4) return B;
	b = B >> (4 - q);
	return b < 8192 ? (B < 8192 ? B : 8192) : b;
}

/* ... and the last ones shrink again, so that the pipeline drains as fast as it fills: with `left` records to go (0 = not
 * known) the next batch takes a third of them once fewer than two full batches remain */
static size_t tail_records(size_t want, size_t B, size_t left)
{
	const char *e = getenv("BWAGPU_BATCH_RAMP");
	size_t b;
	if ((e && atoi(e) == 0) || left == 0 || left >= 2 * B) return want;
	b = left / 3;
	if (b < B / 16) b = B / 16;
	if (b < 4096) b = 4096;
	return b < want ? b : want;
}

static size_t batch_records(void)
{
	const char *e = getenv("BWAGPU_BATCH_RECORDS");
	const long v = e ? atol(e) : 0;
	return v > 0 ? (size_t)v : (size_t)1 << 17;
}

/* ------------------------------------------------------------------ bwt_sa: real / replay
 * A queue of SA rows whose coordinates are wanted, one device call, then the reference's own code run again with bwt_sa
 * answering from the results in the same order (per thread: the replaying thread owns the cursor). */
typedef struct { size_t n, m; bwtint_t *k; uint8_t *which; bwtint_t *out; size_t m_out; } saq_t;
static __thread const saq_t *t_sa_q;  /* non-NULL: bwt_sa replays from it */
static __thread size_t t_sa_pos;

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
	if (q->n + 1 > q->m_out) { q->m_out = q->n + 1 + q->n / 4; q->out = (bwtint_t *)realloc(q->out, q->m_out * sizeof(bwtint_t)); }
	if (q->n && bwa_gpu_cal_pac_pos((int64_t)q->n, q->k, q->which, q->out)) die("bwa_gpu_cal_pac_pos");
	REP_ADD(calls_sa, q_sa, dev_sa_s, q->n, t0);
}

static void saq_free(saq_t *q) { free(q->k); free(q->which); free(q->out); memset(q, 0, sizeof(*q)); }

bwtint_t bwt_sa(const bwt_t *bwt, bwtint_t k)
{
	REAL(bwtint_t, bwt_sa, const bwt_t *, bwtint_t);
	if (t_sa_q) {
		const saq_t *q = t_sa_q;
		if (t_sa_pos >= q->n || q->k[t_sa_pos] != k || q->which[t_sa_pos] != (uint8_t)(bwt == g_bwt[0])) {
			fprintf(stderr, "[bwa_gpu_batch] bwt_sa replay out of step at query %zu\n", t_sa_pos);
			abort();
		}
		return q->out[t_sa_pos++];
	}
	return real_bwt_sa(bwt, k); /* index construction and other uses outside the batched loops */
}

/* ------------------------------------------------------------------ aln_local_core: real / record / replay (per thread) */
enum { RR_REAL = 0, RR_RECORD = 1, RR_REPLAY = 2 };
typedef struct { size_t n, m; bwa_gpu_sw_job_t *job; size_t *seq_off; size_t sn, sm; ubyte_t *seqs; } swq_t;
static __thread int t_sw_mode;
static __thread int64_t t_sw_beg;       /* *beg of the bwa_sw_core call in progress */
static __thread swq_t *t_swq;           /* RECORD: where the jobs go */
static __thread const bwa_gpu_sw_job_t *t_sw_jobs; /* REPLAY: this thread's part of the batch's jobs / results */
static __thread const bwa_gpu_path_res_t *t_sw_res;
static __thread const bwa_cigar_t *t_sw_pool;
static __thread size_t t_sw_pos, t_sw_n;

bwa_cigar_t *bwa_sw_core(bwtint_t l_pac, const ubyte_t *pacseq, int len, const ubyte_t *seq, int64_t *beg, int reglen,
                         int *n_cigar, uint32_t *cnt)
{
	REAL(bwa_cigar_t *, bwa_sw_core, bwtint_t, const ubyte_t *, int, const ubyte_t *, int64_t *, int, int *, uint32_t *);
	t_sw_beg = *beg; /* the window the reference is about to unpack (bwape.c:447-450) */
	return real_bwa_sw_core(l_pac, pacseq, len, seq, beg, reglen, n_cigar, cnt);
}

/* path_t[] as aln_global_core's backtrace leaves it (stdaln.c:496-512): path[0] = the end cell, path[path_len-1] = the
 * start cell; an element's ctype says how its cell was entered.  Rebuilt from the device's CIGAR and start cell. */
static int path_from_cigar(const bwa_cigar_t *cg, int n_cigar, int start_i, int start_j, path_t *path)
{
	int n = 0, c, t, i = start_i, j = start_j;
	for (c = 0; c < n_cigar; ++c) n += cg[c] & 0x3fff;
	for (c = 0, t = n - 1; c < n_cigar; ++c) {
		const int op = cg[c] >> 14, run = cg[c] & 0x3fff;
		int u;
		for (u = 0; u < run; ++u, --t) {
			if (t != n - 1) { if (op != FROM_I) ++i; if (op != FROM_D) ++j; }
			path[t].i = i; path[t].j = j; path[t].ctype = (unsigned char)op;
		}
	}
	return n;
}

int aln_local_core(unsigned char *seq1, int len1, unsigned char *seq2, int len2, const AlnParam *ap, path_t *path, int *path_len,
                   int thres, int *subo)
{
	REAL(int, aln_local_core, unsigned char *, int, unsigned char *, int, const AlnParam *, path_t *, int *, int, int *);
	if (t_sw_mode == RR_RECORD) {
		swq_t *q = t_swq;
		if (q->n == q->m) {
			q->m = q->m ? q->m << 1 : 1 << 10;
			q->job = (bwa_gpu_sw_job_t *)realloc(q->job, q->m * sizeof(*q->job));
			q->seq_off = (size_t *)realloc(q->seq_off, q->m * sizeof(size_t));
		}
		if (q->sn + (size_t)len2 > q->sm) {
			q->sm = (q->sn + (size_t)len2) * 2 + 4096;
			q->seqs = (ubyte_t *)realloc(q->seqs, q->sm);
		}
		memcpy(q->seqs + q->sn, seq2, (size_t)len2); /* bwa_paired_sw1 un-reverses the read right after the call */
		q->job[q->n].beg = t_sw_beg; q->job[q->n].reglen = len1; q->job[q->n].len = len2; q->job[q->n].seq = 0;
		q->seq_off[q->n] = q->sn; q->sn += (size_t)len2; ++q->n;
		return -1; /* "no alignment": bwa_sw_core returns 0 and bwa_paired_sw1 leaves both reads untouched (bwape.c:457-460) */
	}
	if (t_sw_mode == RR_REPLAY) {
		const bwa_gpu_path_res_t *r;
		if (t_sw_pos >= t_sw_n || t_sw_jobs[t_sw_pos].reglen != len1 || t_sw_jobs[t_sw_pos].len != len2) {
			fprintf(stderr, "[bwa_gpu_batch] aln_local_core replay out of step at job %zu\n", t_sw_pos);
			abort();
		}
		r = &t_sw_res[t_sw_pos++];
		if (r->score < 0) return r->score;
		*path_len = path_from_cigar(t_sw_pool + r->cigar_off, r->n_cigar, r->start_i, r->start_j, path);
		if (*path_len && (path[0].i != r->end_i || path[0].j != r->end_j)) {
			fprintf(stderr, "[bwa_gpu_batch] device path does not end at its end cell (%d,%d) vs (%d,%d)\n", path[0].i, path[0].j, r->end_i, r->end_j);
			abort();
		}
		if (subo) *subo = 0;
		return r->score;
	}
	return real_aln_local_core(seq1, len1, seq2, len2, ap, path, path_len, thres, subo);
}

/* ------------------------------------------------------------------ aln_global_core inside bwa_refine_gapped: real / record / replay
 * refine_gapped_core (bwase.c:189-237, static) unpacks its reference window and calls aln_global_core (bwase.c:212).
 * RECORD runs bwa_refine_gapped on a scratch copy of the read with aln_global_core noting both sequences (and bwa_cal_md1
 * switched off: the copy's MD is not wanted); REPLAY is the real run, aln_global_core answering with the device's path. */
typedef struct { size_t n, m; int *len1, *len2; size_t *off1, *off2; size_t bn, bm; ubyte_t *bytes; } gaq_t;
static __thread int t_ga_mode;
static __thread gaq_t *t_gaq;
static __thread const bwa_gpu_path_res_t *t_ga_res;
static __thread const bwa_cigar_t *t_ga_pool;
static __thread const int *t_ga_len1, *t_ga_len2;
static __thread size_t t_ga_pos, t_ga_n;

static int is_bwa_param(const AlnParam *ap)
{
	return ap->gap_open == aln_param_bwa.gap_open && ap->gap_ext == aln_param_bwa.gap_ext && ap->matrix == aln_param_bwa.matrix &&
	       ap->row == aln_param_bwa.row && ap->band_width == aln_param_bwa.band_width && ap->gap_end == aln_param_bwa.gap_end;
}

int aln_global_core(unsigned char *seq1, int len1, unsigned char *seq2, int len2, const AlnParam *ap, path_t *path, int *path_len)
{
	REAL(int, aln_global_core, unsigned char *, int, unsigned char *, int, const AlnParam *, path_t *, int *);
	if (t_ga_mode == RR_RECORD && is_bwa_param(ap) && len1 > 0 && len2 > 0) {
		gaq_t *q = t_gaq;
		if (q->n == q->m) {
			q->m = q->m ? q->m << 1 : 1 << 10;
			q->len1 = (int *)realloc(q->len1, q->m * sizeof(int)); q->len2 = (int *)realloc(q->len2, q->m * sizeof(int));
			q->off1 = (size_t *)realloc(q->off1, q->m * sizeof(size_t)); q->off2 = (size_t *)realloc(q->off2, q->m * sizeof(size_t));
		}
		if (q->bn + (size_t)len1 + (size_t)len2 > q->bm) {
			q->bm = (q->bn + (size_t)len1 + (size_t)len2) * 2 + 4096;
			q->bytes = (ubyte_t *)realloc(q->bytes, q->bm);
		}
		q->len1[q->n] = len1; q->len2[q->n] = len2;
		q->off1[q->n] = q->bn; memcpy(q->bytes + q->bn, seq1, (size_t)len1); q->bn += (size_t)len1;
		q->off2[q->n] = q->bn; memcpy(q->bytes + q->bn, seq2, (size_t)len2); q->bn += (size_t)len2;
		++q->n;
		path[0].i = len1; path[0].j = len2; path[0].ctype = FROM_M; /* a one-cell path keeps refine_gapped_core's CIGAR fix-ups in bounds */
		*path_len = 1;
		return 0;
	}
	if (t_ga_mode == RR_REPLAY && is_bwa_param(ap) && len1 > 0 && len2 > 0) {
		const bwa_gpu_path_res_t *r;
		if (t_ga_pos >= t_ga_n || t_ga_len1[t_ga_pos] != len1 || t_ga_len2[t_ga_pos] != len2) {
			fprintf(stderr, "[bwa_gpu_batch] aln_global_core replay out of step at job %zu\n", t_ga_pos);
			abort();
		}
		r = &t_ga_res[t_ga_pos++];
		*path_len = path_from_cigar(t_ga_pool + r->cigar_off, r->n_cigar, r->start_i, r->start_j, path);
		return r->score;
	}
	return real_aln_global_core(seq1, len1, seq2, len2, ap, path, path_len);
}

char *bwa_cal_md1(int n_cigar, bwa_cigar_t *cigar, int len, bwtint_t pos, ubyte_t *seq, const bntseq_t *bns, ubyte_t *pacseq,
                  kstring_t *str, int *_nm)
{
	REAL(char *, bwa_cal_md1, int, bwa_cigar_t *, int, bwtint_t, ubyte_t *, const bntseq_t *, ubyte_t *, kstring_t *, int *);
	if (t_ga_mode == RR_RECORD) { *_nm = 0; return 0; }
	return real_bwa_cal_md1(n_cigar, cigar, len, pos, seq, bns, pacseq, str, _nm);
}

/* ------------------------------------------------------------------ intermediate records (shim_io.c keeps them) */
typedef struct { bam_pair_t *recs; size_t base; } codec_ctx_t;
static void decode_one(size_t i, void *ctx)
{
	codec_ctx_t *c = (codec_ctx_t *)ctx;
	const uint8_t *data;
	uint32_t len;
	zmq_msg_t m;
	memtemp_get(c->base + i, &data, &len);
	zmq_msg_init_data(&m, (void *)data, len, 0, 0);
	pair_init_from_msg(&c->recs[i], &m);
	zmq_msg_close(&m);
}

/* a dozen free()s per record.  Serial on purpose: freeing from several threads what other threads allocated contends on
 * glibc's arenas (measured: 7 CPU-seconds per million reads instead of 0.4, and the parse thread's mallocs slow down too) */
static void destroy_records(bam_pair_t *recs, size_t n)
{
	const double c0 = thread_cpu_now();
	size_t i;
	for (i = 0; i < n; ++i) bam_destroy_pair(&recs[i]);
	cpu_add(CPU_DESTROY, thread_cpu_now() - c0);
}

/* pass 1: records [0, n) -> memory (or, once that is full, the reference's temporary file).  Serial on purpose: the encoder
 * mallocs one message per record and the chunks are appended in order */
static void store_records(gzFile temporary, bam_pair_t *recs, size_t n)
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
	}
}

/* pass 2: up to B records, the ones kept in memory first */
static size_t load_records(gzFile temporary, bam_pair_t *recs, size_t B, long *tot_seqs)
{
	size_t n, i, first;
	n = memtemp_take(B, &first);
	if (n) {
		codec_ctx_t c = {recs, first};
		t_cpu_bucket = CPU_LOAD;
		parallel_for(n, 1024, decode_one, &c);
		for (i = 0; i < n; ++i) *tot_seqs += recs[i].kind;
	}
	while (n < B && memtemp_spilled()) {
		const int rc = read_pair_custom(temporary, &recs[n]);
		if (rc < 0) { fprintf(stderr, "[bwa_gpu_batch] error reading intermediate file\n"); exit(1); }
		if (rc == 0) break;
		*tot_seqs += recs[n].kind;
		++n;
	}
	return n;
}

/* ------------------------------------------------------------------ pass 1 */
static void gpu_align(int m, bwa_seq_t *flat)
{
	const double t0 = now();
	if (m && bwa_gpu_cal_sa_reads_gap(m, flat, g_gap)) die("bwa_gpu_cal_sa_reads_gap");
	REP_ADD(calls_aln, reads_aln, dev_aln_s, m, t0);
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
static void align_range(bam_pair_t *recs, size_t n, bwa_seq_t *flat, double *t_toseq, int seq_ready)
{
	size_t i;
	int m = 0, j;
	double t1 = now();
	/* aln_singleton / aln_pair without the search ... (seq_ready: the read stage has run bam1_to_seq already) */
	if (!seq_ready) {
		t_cpu_bucket = CPU_TOSEQ;
		parallel_for(n, 2048, to_seq_one, recs);
	}
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
 * one device call for the SA rows whose coordinates are wanted, then the reference's own bwa_cal_pac_pos_core fed from the
 * answers (slices of records on the host threads, each replaying its own part of the queue) */
typedef struct { bam_pair_t *recs; const saq_t *q; const size_t *qoff; } posn_ctx_t;
static void posn_slice(int s, size_t lo, size_t hi, void *ctx)
{
	posn_ctx_t *c = (posn_ctx_t *)ctx;
	size_t i;
	int j, k;
	(void)s;
	t_sa_q = c->q; t_sa_pos = c->qoff[lo];
	for (i = lo; i < hi; ++i) {
		bam_pair_t *r = &c->recs[i];
		if (r->phase != aligned) continue;
		if (unique_rec(r))
			for (j = 0; j != (int)r->kind; ++j) {
				bwa_seq_t *p = &r->bwa_seq[j];
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
	if (t_sa_pos != c->qoff[hi]) { fprintf(stderr, "[bwa_gpu_batch] SA answers of records %zu..%zu out of step\n", lo, hi); abort(); }
	t_sa_q = 0;
}

static void position_range(bam_pair_t *recs, size_t n, saq_t *q, size_t **qoff_buf, size_t *qoff_cap, double *t_host)
{
	size_t i;
	int j;
	double t1 = now(), c0 = thread_cpu_now();
	posn_ctx_t c;
	if (n + 1 > *qoff_cap) { *qoff_cap = n + 1; *qoff_buf = (size_t *)realloc(*qoff_buf, (n + 1) * sizeof(size_t)); }
	q->n = 0;
	for (i = 0; i < n; ++i) {
		bam_pair_t *r = &recs[i];
		(*qoff_buf)[i] = q->n;
		if (r->phase != aligned || !unique_rec(r)) continue;
		for (j = 0; j != (int)r->kind; ++j) {
			bwa_seq_t *p = &r->bwa_seq[j];
			int k;
			if (r->kind == singleton) bwa_aln2seq_core(p->n_aln, p->aln, p, 1, g_pe->max_occ_se);
			else { p->n_multi = 0; bwa_aln2seq(p->n_aln, p->aln, p); }
			if (is_mapped(p)) saq_push(q, p->sa, p->strand);
			if (r->kind == singleton)
				for (k = 0; k < p->n_multi; ++k) saq_push(q, p->multi[k].pos, p->multi[k].strand);
		}
	}
	(*qoff_buf)[n] = q->n;
	*t_host += now() - t1;
	cpu_add(CPU_POSN_SERIAL, thread_cpu_now() - c0);
	saq_run(q);
	t1 = now();
	c.recs = recs; c.q = q; c.qoff = *qoff_buf;
	t_cpu_bucket = CPU_POSN_PAR;
	parallel_slices(n, 4096, posn_slice, &c);
	*t_host += now() - t1;
}

/* Pass 1 as a pipeline of five stages, each on its own thread, batches handed on in order through a ring of slots:
 *   read     read_bam_pair (parse; the inflate runs on shim_io.c's threads)
 *   align    bam1_to_seq on the host threads + the search on the device            (the calling thread)
 *   position bwa_aln2seq_core in record order (drand48), SA rows on the device, bwa_cal_pac_pos_core, improve_isize_est
 *   store    the reference's record encoding into memory / its temporary file
 *   destroy  bam_destroy_pair of the batch (a dozen free()s per record)
 * Every order-sensitive piece of state belongs to exactly one stage, and a stage sees the batches in input order, so the
 * result is what the one-record-at-a-time loop (bam2bam.c:1143-1176) produces. */
#define P1_SLOTS 12 /* the reader may run this many batches ahead (it does while the device context is created) */
enum { SL_FREE = 0, SL_PARSED, SL_READ, SL_ALIGNED, SL_POSITIONED, SL_STORED };
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
	double t0, t_read, t_toseq, t_host, t_write, t_destroy[2];
	long tot_seqs;
	size_t n_read; /* records the read stage has taken so far */
	int destroy_done;
	int align_done; /* the end marker has reached the align stage */
} pipe1_t;

static void slot_wait(pthread_mutex_t *mu, pthread_cond_t *cv, const int *state, int want)
{
	pthread_mutex_lock(mu);
	while (*(volatile const int *)state != want) pthread_cond_wait(cv, mu);
	pthread_mutex_unlock(mu);
}

static void slot_set(pthread_mutex_t *mu, pthread_cond_t *cv, int *state, int st)
{
	if (g_trace) fprintf(stderr, "[trace] %.4f pass %d slot %p -> %d\n", now() - g_t_call, g_trace_pass, (void *)state, st);
	pthread_mutex_lock(mu);
	*state = st;
	pthread_cond_broadcast(cv);
	pthread_mutex_unlock(mu);
}
#define P1_WAIT(P, slot, want) slot_wait(&(P)->mu, &(P)->cv, &(P)->state[slot], want)
#define P1_SET(P, slot, st) slot_set(&(P)->mu, &(P)->cv, &(P)->state[slot], st)

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
		P1_WAIT(P, slot, SL_FREE);
		t = now();
		const double c0 = thread_cpu_now();
		{
			const double done = fastin_progress();
			const size_t left = done > 0.02 && done < 1.0 ? (size_t)((double)P->n_read * (1.0 - done) / done) : 0;
			n = fastin_read_pairs(P->ks, recs, tail_records(ramp_records(P->B, q), P->B, left), &seqs, g_broken_input, g_drop_aligned);
			P->n_read += n;
		}
		P->t_read += now() - t;
		cpu_add(CPU_PARSE, thread_cpu_now() - c0);
		P->n[slot] = n; P->seqs[slot] = seqs;
		P1_SET(P, slot, SL_PARSED);
		if (n == 0) break; /* an empty batch is the end marker; it travels through every stage */
	}
	return 0;
}

static void *stage_toseq(void *arg) /* bam1_to_seq, so that the align stage is the device call and nothing else */
{
	pipe1_t *P = (pipe1_t *)arg;
	unsigned q;
	for (q = 0;; ++q) {
		const int slot = (int)(q % P1_SLOTS);
		double t;
		P1_WAIT(P, slot, SL_PARSED);
		t = now();
		if (P->n[slot]) {
			t_cpu_bucket = CPU_TOSEQ;
			parallel_for(P->n[slot], 2048, to_seq_one, P->recs[slot]);
		}
		P->t_toseq += now() - t;
		P1_SET(P, slot, SL_READ);
		if (P->n[slot] == 0) break;
	}
	return 0;
}

static void *stage_position(void *arg)
{
	pipe1_t *P = (pipe1_t *)arg;
	saq_t q;
	size_t *qoff = 0, qoff_cap = 0;
	unsigned b;
	memset(&q, 0, sizeof(q));
	for (b = 0;; ++b) {
		const int slot = (int)(b % P1_SLOTS);
		bam_pair_t *recs;
		size_t n, i;
		double t1;
		P1_WAIT(P, slot, SL_ALIGNED);
		recs = P->recs[slot]; n = P->n[slot];
		if (n) {
			position_range(recs, n, &q, &qoff, &qoff_cap, &P->t_host);
			t1 = now();
			{
				const double c0 = thread_cpu_now();
				for (i = 0; i < n; ++i) /* the unchanged tail of the loop (bam2bam.c:1167-1170) */
					if (unique_rec(&recs[i])) improve_isize_est(P->iinfos, &recs[i], g_pe->ap_prior, g_bwt[0]->seq_len);
				cpu_add(CPU_ISIZE, thread_cpu_now() - c0);
			}
			P->t_host += now() - t1;
		}
		P1_SET(P, slot, SL_POSITIONED);
		if (n == 0) break;
	}
	saq_free(&q); free(qoff);
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
		P1_WAIT(P, slot, SL_POSITIONED);
		n = P->n[slot];
		if (n == 0) { P1_SET(P, slot, SL_STORED); break; }
		t1 = now();
		{
			const double c0 = thread_cpu_now();
			store_records(P->temporary, P->recs[slot], n);
			cpu_add(CPU_STORE, thread_cpu_now() - c0);
		}
		P->t_write += now() - t1;
		P->tot_seqs += P->seqs[slot];
		fprintf(stderr, "[sequential_loop_pass1] %ld sequences processed in %.2f sec\n", P->tot_seqs, now() - P->t0);
		P1_SET(P, slot, SL_STORED);
	}
	return 0;
}

typedef struct { void *P; int id; } destroy1_arg_t;
static void *stage_destroy(void *arg) /* two threads, whole batches in turn */
{
	destroy1_arg_t *A = (destroy1_arg_t *)arg;
	pipe1_t *P = (pipe1_t *)A->P;
	unsigned q;
	for (q = (unsigned)A->id;; q += 2) {
		const int slot = (int)(q % P1_SLOTS);
		double t1;
		pthread_mutex_lock(&P->mu);
		while (P->state[slot] != SL_STORED && !P->destroy_done) pthread_cond_wait(&P->cv, &P->mu);
		if (P->state[slot] != SL_STORED) { pthread_mutex_unlock(&P->mu); break; }
		if (P->n[slot] == 0) { P->destroy_done = 1; pthread_cond_broadcast(&P->cv); pthread_mutex_unlock(&P->mu); break; }
		pthread_mutex_unlock(&P->mu);
		t1 = now();
		destroy_records(P->recs[slot], P->n[slot]);
		P->t_destroy[A->id] += now() - t1;
		P1_SET(P, slot, SL_FREE);
	}
	return 0;
}

#define MAX_ALIGN 4
typedef struct { pipe1_t *P; int id, n; bwa_seq_t *flat; double t_toseq; } align_arg_t;

static int align_threads(void)
{
	const char *e = getenv("BWAGPU_CALL_GROUPS");
	int n = e ? atoi(e) : 2;
	return n < 1 ? 1 : n > MAX_ALIGN ? MAX_ALIGN : n;
}

static void *stage_align(void *arg)
{
	align_arg_t *A = (align_arg_t *)arg;
	pipe1_t *P = A->P;
	unsigned q;
	for (q = (unsigned)A->id;; q += (unsigned)A->n) {
		const int slot = (int)(q % P1_SLOTS);
		pthread_mutex_lock(&P->mu);
		while (P->state[slot] != SL_READ && !P->align_done) pthread_cond_wait(&P->cv, &P->mu);
		if (P->state[slot] != SL_READ) { pthread_mutex_unlock(&P->mu); break; } /* another align thread met the end marker */
		pthread_mutex_unlock(&P->mu);
		if (P->n[slot]) align_range(P->recs[slot], P->n[slot], A->flat, &A->t_toseq, 1);
		if (P->n[slot] == 0) { pthread_mutex_lock(&P->mu); P->align_done = 1; pthread_mutex_unlock(&P->mu); }
		P1_SET(P, slot, SL_ALIGNED);
		if (P->n[slot] == 0) break;
	}
	return 0;
}

void sequential_loop_pass1(bwa_seqio_t *ks, gzFile temporary, khash_t(isize_infos) *iinfos)
{
	const size_t B = batch_records();
	pipe1_t P;
	double t_toseq = 0, t_init;
	bwa_seq_t *flat = (bwa_seq_t *)malloc(2 * B * sizeof(bwa_seq_t));
	pthread_t init_th, read_th, seq_th, pos_th, store_th, destroy_th[2];
	destroy1_arg_t da[2];
	int s;
	memset(&P, 0, sizeof(P));
	pthread_mutex_init(&P.mu, 0); pthread_cond_init(&P.cv, 0);
	P.B = B; P.ks = ks; P.temporary = temporary; P.iinfos = iinfos; P.t0 = now();
	g_t_pass1_begin = P.t0; g_trace_pass = 1;
	for (s = 0; s < P1_SLOTS; ++s) P.recs[s] = (bam_pair_t *)calloc(B, sizeof(bam_pair_t));
	memtemp_begin();
	/* device context + index upload (seconds) overlap the reading of the first batches */
	pthread_create(&init_th, 0, ensure_gpu_thread, 0);
	pthread_create(&read_th, 0, stage_read, &P);
	pthread_create(&seq_th, 0, stage_toseq, &P);
	pthread_create(&pos_th, 0, stage_position, &P);
	pthread_create(&store_th, 0, stage_store, &P);
	for (s = 0; s < 2; ++s) { da[s].P = &P; da[s].id = s; pthread_create(&destroy_th[s], 0, stage_destroy, &da[s]); }
	pthread_join(init_th, 0);
	t_init = now() - P.t0;
	{ /* the align stage: N_ALIGN threads take the batches in turn, so that one batch's device call (its host-side packing, its
	   * kernels' straggler tail, its unpacking) overlaps the next one's -- the library runs them on different lane groups */
		align_arg_t aa[MAX_ALIGN];
		pthread_t ath[MAX_ALIGN];
		const int na = align_threads();
		int a;
		for (a = 0; a < na; ++a) {
			aa[a].P = &P; aa[a].id = a; aa[a].n = na; aa[a].t_toseq = 0;
			aa[a].flat = a == 0 ? flat : (bwa_seq_t *)malloc(2 * B * sizeof(bwa_seq_t));
			pthread_create(&ath[a], 0, stage_align, &aa[a]);
		}
		for (a = 0; a < na; ++a) { pthread_join(ath[a], 0); t_toseq += aa[a].t_toseq; if (a) free(aa[a].flat); }
	}
	pthread_join(read_th, 0); pthread_join(seq_th, 0); pthread_join(pos_th, 0); pthread_join(store_th, 0); pthread_join(destroy_th[0], 0); pthread_join(destroy_th[1], 0);
	for (s = 0; s < P1_SLOTS; ++s) free(P.recs[s]);
	free(flat);
	pthread_mutex_destroy(&P.mu); pthread_cond_destroy(&P.cv);
	g_rep.pass1_s = now() - P.t0;
	g_t_pass1_end = now();
	g_rep.sequences = P.tot_seqs;
	fprintf(stderr, "[%s] %zu records (%.0f MB) kept in memory for pass 2%s\n", __func__, memtemp_records(), memtemp_bytes() / 1048576.0,
	        memtemp_spilled() ? ", the rest in the temporary file" : "");
	fprintf(stderr, "[%s] %ld sequences processed in %.2f sec (pipelined stages, busy seconds each: device init %.2f, read %.2f, bam1_to_seq %.2f, aln2seq/posn/isize %.2f, device calls %.2f, temp write %.2f, destroy %.2f + %.2f)\n",
	        __func__, P.tot_seqs, now() - P.t0, t_init, P.t_read, P.t_toseq + t_toseq, P.t_host, g_rep.dev_aln_s + g_rep.dev_sa_s, P.t_write, P.t_destroy[0], P.t_destroy[1]);
	fprintf(stderr, "[%s] finished cleanly.\n", __func__);
}

/* ------------------------------------------------------------------ pass 2: finish_pair / finish_singleton, phase by phase */
typedef struct { bwtint_t *arr; uint32_t n; uint8_t fresh, strand; int len; size_t qpos; } visit_t; /* one wide interval, as a record meets it */

typedef struct { /* everything one batch carries through the stages */
	bam_pair_t *recs;
	size_t n;
	long seqs;
	khash_t(isize_infos) *iinfos;
	/* enumerate */
	saq_t q1;
	visit_t *visit; size_t n_visit, m_visit;
	size_t *qoff, *voff; uint8_t *want; size_t m_rec;
	/* XA rows */
	saq_t q2;
	/* mate rescue / refinement: per-slice queues, merged job lists, results */
	swq_t swq[MAX_THREADS];
	gaq_t gaq[MAX_THREADS];
	bwa_gpu_sw_job_t *sw_jobs; size_t m_sw_jobs; size_t sw_off[MAX_THREADS + 1];
	bwa_gpu_path_res_t *sw_res; size_t m_sw_res; const bwa_cigar_t *sw_pool;
	bwa_cigar_t *sw_pool_copy; size_t m_sw_pool;
	bwa_gpu_ga_job_t *ga_jobs; size_t m_ga_jobs; size_t ga_off[MAX_THREADS + 1];
	int *ga_len1, *ga_len2; size_t m_ga_len;
	bwa_gpu_path_res_t *ga_res; size_t m_ga_res; const bwa_cigar_t *ga_pool;
	bwa_cigar_t *ga_pool_copy; size_t m_ga_pool;
	uint64_t n_tot[MAX_THREADS][2], n_mapped[MAX_THREADS][2];
	int n_slices;
} batch2_t;

static void batch2_free(batch2_t *b)
{
	int s;
	saq_free(&b->q1); saq_free(&b->q2);
	free(b->visit); free(b->qoff); free(b->voff); free(b->want);
	for (s = 0; s < MAX_THREADS; ++s) {
		free(b->swq[s].job); free(b->swq[s].seq_off); free(b->swq[s].seqs);
		free(b->gaq[s].len1); free(b->gaq[s].len2); free(b->gaq[s].off1); free(b->gaq[s].off2); free(b->gaq[s].bytes);
	}
	free(b->sw_jobs); free(b->sw_res); free(b->sw_pool_copy); free(b->ga_pool_copy); free(b->ga_jobs); free(b->ga_len1); free(b->ga_len2); free(b->ga_res);
	memset(b, 0, sizeof(*b));
}

static const isize_info_t *ii_of(khash_t(isize_infos) *iinfos, const bam_pair_t *r) /* bam2bam.c:715-716 */
{
	khiter_t it = kh_get(isize_infos, iinfos, bam_get_rg(r->bam_rec));
	return it == kh_end(iinfos) ? &g_null_ii : &kh_val(iinfos, it);
}

static long long n_rows(const bam_pair_t *r)
{
	long long q = 0;
	int j, k;
	for (j = 0; j < 2; ++j)
		for (k = 0; k < r->bwa_seq[j].n_aln; ++k) q += (long long)r->bwa_seq[j].aln[k].l - r->bwa_seq[j].aln[k].k + 1;
	return q;
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

static int is_pair_job(const bam_pair_t *r) { return r->kind == proper_pair && r->phase == positioned && unique_rec(r); }
static int is_single_job(const bam_pair_t *r) { return r->kind == singleton && r->phase == positioned && unique_rec(r); }

static void pair_to_seq_one(size_t i, void *ctx) /* finish_*'s bam1_to_seq (bam2bam.c:646, 717-718) */
{
	bam_pair_t *r = (bam_pair_t *)ctx + i;
	int j;
	if (!is_pair_job(r) && !is_single_job(r)) return;
	for (j = 0; j < (int)r->kind; ++j)
		if (!r->bwa_seq[j].seq) bam1_to_seq(&r->bam_rec[j], &r->bwa_seq[j], 1, g_gap->trim_qual);
}

/* A: every SA row whose coordinate pairing will want (bam2bam.c:736-765), in the reference's visiting order.  Owns the
 * position cache: a wide interval is looked up by (k,l) only; the first record to touch a key creates the entry and its
 * rows are fetched, later records reuse the values (which are filled in stage B of the batch that created them). */
static void stage_enumerate(batch2_t *b, kh_64_t *my_hash)
{
	size_t i;
	int j, k;
	bwtint_t l;
	if (b->n + 1 > b->m_rec) {
		b->m_rec = b->n + 1;
		b->qoff = (size_t *)realloc(b->qoff, b->m_rec * sizeof(size_t)); b->voff = (size_t *)realloc(b->voff, b->m_rec * sizeof(size_t));
		b->want = (uint8_t *)realloc(b->want, b->m_rec);
	}
	t_cpu_bucket = CPU_TOSEQ;
	parallel_for(b->n, 1024, pair_to_seq_one, b->recs);
	b->q1.n = 0; b->n_visit = 0;
	const double c0 = thread_cpu_now();
	for (i = 0; i < b->n; ++i) {
		bam_pair_t *r = &b->recs[i];
		b->qoff[i] = b->q1.n; b->voff[i] = b->n_visit; b->want[i] = 0;
		if (!is_pair_job(r) || !wants_pairing(r)) continue;
		b->want[i] = 1;
		for (j = 0; j < 2; ++j)
			for (k = 0; k < r->bwa_seq[j].n_aln; ++k) {
				const bwt_aln1_t *a = r->bwa_seq[j].aln + k;
				int is_new = 1;
				if (a->l - a->k + 1 >= MIN_HASH_WIDTH) {
					int ret;
					khint_t it = kh_put(64, my_hash, (uint64_t)a->k << 32 | a->l, &ret);
					poslist_t *z = &kh_val(my_hash, it);
					visit_t *v;
					is_new = ret != 0;
					if (is_new) {
						z->n = a->l - a->k + 1;
						z->a = (bwtint_t *)malloc(sizeof(bwtint_t) * z->n);
					}
					if (b->n_visit == b->m_visit) { b->m_visit = b->m_visit ? b->m_visit << 1 : 256; b->visit = (visit_t *)realloc(b->visit, b->m_visit * sizeof(visit_t)); }
					v = &b->visit[b->n_visit++];
					v->arr = z->a; v->n = (uint32_t)z->n; v->fresh = (uint8_t)is_new; v->strand = (uint8_t)a->a; v->len = (int)r->bwa_seq[j].len; v->qpos = b->q1.n;
				}
				if (is_new)
					for (l = a->k; l <= a->l; ++l) { saq_push(&b->q1, l, a->a); if (l == a->l) break; }
			}
	}
	b->qoff[b->n] = b->q1.n; b->voff[b->n] = b->n_visit;
	cpu_add(CPU_ENUM, thread_cpu_now() - c0);
	saq_run(&b->q1);
}

/* B1: the values of the cache entries this batch created (their creator's strand and length define them, bam2bam.c:751-757) */
static void fill_visit_one(size_t vi, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	const visit_t *v = &b->visit[vi];
	uint32_t t;
	if (!v->fresh) return;
	for (t = 0; t < v->n; ++t) v->arr[t] = v->strand ? b->q1.out[v->qpos + t] : g_bwt[1]->seq_len - (b->q1.out[v->qpos + t] + v->len);
}

/* B2: pairing (bwape.c:180-293) of one record, on the coordinates stage A fetched */
static void pairing_one(size_t i, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	bam_pair_t *r = &b->recs[i];
	bwa_seq_t *p[2];
	pe_data_t d;
	size_t qpos, vpos;
	int j, k;
	if (!b->want[i]) return;
	p[0] = &r->bwa_seq[0]; p[1] = &r->bwa_seq[1];
	memset(&d, 0, sizeof(pe_data_t));
	for (j = 0; j < 2; ++j) { d.aln[j].a = p[j]->aln; d.aln[j].n = p[j]->n_aln; }
	qpos = b->qoff[i]; vpos = b->voff[i];
	for (j = 0; j < 2; ++j)
		for (k = 0; k < (int)d.aln[j].n; ++k) {
			const bwt_aln1_t *a = d.aln[j].a + k;
			const bwtint_t w = a->l - a->k + 1;
			bwtint_t t;
			uint64_t x;
			if (w >= MIN_HASH_WIDTH) {
				const visit_t *v = &b->visit[vpos++];
				if (v->fresh) qpos += w;
				for (t = 0; t < (bwtint_t)v->n; ++t) {
					x = v->arr[t];
					x = x << 32 | k << 1 | j;
					kv_push(uint64_t, d.arr, x);
				}
			} else
				for (t = 0; t < w; ++t, ++qpos) {
					x = a->a ? b->q1.out[qpos] : g_bwt[1]->seq_len - (b->q1.out[qpos] + p[j]->len);
					x = x << 32 | k << 1 | j;
					kv_push(uint64_t, d.arr, x);
				}
		}
	if (qpos != b->qoff[i + 1] || vpos != b->voff[i + 1]) { fprintf(stderr, "[bwa_gpu_batch] pass-2 enumeration out of step at record %zu\n", i); abort(); }
	pairing(p, &d, g_pe, g_gap->s_mm, ii_of(b->iinfos, r));
	kv_destroy(d.arr);
}

typedef struct { batch2_t *b; } xa_ctx_t;
static void xa_assign_slice(int s, size_t lo, size_t hi, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	size_t i, qpos = b->voff[lo]; /* voff is reused as the per-record offset into q2 by stage_pairing */
	int j, k;
	(void)s;
	for (i = lo; i < hi; ++i) {
		bam_pair_t *r = &b->recs[i];
		if (!is_pair_job(r)) continue;
		for (j = 0; j < 2; ++j) {
			bwa_seq_t *p = &r->bwa_seq[j];
			if (p->type == BWA_TYPE_NO_MATCH) continue;
			for (k = 0; k < p->n_multi; ++k, ++qpos) {
				bwt_multi1_t *q = p->multi + k;
				q->pos = q->strand ? b->q2.out[qpos] : g_bwt[1]->seq_len - (b->q2.out[qpos] + p->len);
			}
		}
	}
}

/* the XA hit lists of one record (bam2bam.c:771-784).  bwa_aln2seq_core is called with set_main = 0 here, and in that form
 * it consumes no random numbers: its sampling branch (bwase.c:74-86) needs a hit group larger than `rest`, and `rest` starts
 * as the sum of all groups (n_occ <= n_multi + 1, else the function has returned at bwase.c:57).  So, unlike the primary-hit
 * selection of pass 1, these calls do not have to run in record order. */
static void xa_core_one(size_t i, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	bam_pair_t *r = &b->recs[i];
	bwa_seq_t *p[2];
	size_t cnt = 0;
	int j;
	b->voff[i] = 0;
	if (!is_pair_job(r) || !(g_pe->N_multi || g_pe->n_multi)) return;
	p[0] = &r->bwa_seq[0]; p[1] = &r->bwa_seq[1];
	for (j = 0; j < 2; ++j)
		if (p[j]->type != BWA_TYPE_NO_MATCH) {
			if (!(p[j]->extra_flag & SAM_FPP) && p[1 - j]->type != BWA_TYPE_NO_MATCH)
				bwa_aln2seq_core(p[j]->n_aln, p[j]->aln, p[j], 0, p[j]->c1 + p[j]->c2 - 1 > g_pe->N_multi ? g_pe->n_multi : g_pe->N_multi);
			else bwa_aln2seq_core(p[j]->n_aln, p[j]->aln, p[j], 0, g_pe->n_multi);
			cnt += (size_t)p[j]->n_multi;
		}
	b->voff[i] = cnt; /* turned into an offset by stage_pairing */
}

static void xa_rows_slice(int s, size_t lo, size_t hi, void *ctx) /* the SA rows of the lists, at their records' offsets in q2 */
{
	batch2_t *b = (batch2_t *)ctx;
	size_t i;
	int j, k;
	(void)s;
	for (i = lo; i < hi; ++i) {
		bam_pair_t *r = &b->recs[i];
		size_t at = b->voff[i];
		if (b->voff[i + 1] == at) continue;
		for (j = 0; j < 2; ++j) {
			const bwa_seq_t *p = &r->bwa_seq[j];
			if (p->type == BWA_TYPE_NO_MATCH) continue;
			for (k = 0; k < p->n_multi; ++k, ++at) { b->q2.k[at] = p->multi[k].pos; b->q2.which[at] = (uint8_t)(p->multi[k].strand != 0); }
		}
	}
}

/* B: pairing, the hit lists for XA and their rows, all on the host threads */
static void stage_pairing(batch2_t *b)
{
	size_t i, total = 0;
	t_cpu_bucket = CPU_PAIRING;
	parallel_for(b->n_visit, 64, fill_visit_one, b);
	parallel_for(b->n, 256, pairing_one, b);
	t_cpu_bucket = CPU_XA_SERIAL;
	parallel_for(b->n, 256, xa_core_one, b);
	for (i = 0; i < b->n; ++i) { const size_t c = b->voff[i]; b->voff[i] = total; total += c; }
	b->voff[b->n] = total;
	if (total > b->q2.m) {
		b->q2.m = total + total / 4 + 1024;
		b->q2.k = (bwtint_t *)realloc(b->q2.k, b->q2.m * sizeof(bwtint_t));
		b->q2.which = (uint8_t *)realloc(b->q2.which, b->q2.m);
	}
	b->q2.n = total;
	if (total) parallel_slices(b->n, 4096, xa_rows_slice, b);
	saq_run(&b->q2);
	if (b->q2.n) parallel_slices(b->n, 4096, xa_assign_slice, b);
}

/* C: mate rescue.  bwa_paired_sw1 (bwape.c:519-633) decides the windows in floating point and judges the alignments; only
 * aln_local_core moves.  First run (host threads, a slice of records each): note the jobs.  Device.  Second run: the real one. */
static void rescue_record_slice(int s, size_t lo, size_t hi, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	uint64_t dummy_tot[2] = {0, 0}, dummy_mapped[2] = {0, 0};
	size_t i;
	b->swq[s].n = 0; b->swq[s].sn = 0;
	t_swq = &b->swq[s];
	t_sw_mode = RR_RECORD;
	for (i = lo; i < hi; ++i)
		if (is_pair_job(&b->recs[i])) {
			bwa_seq_t *p[2] = {&b->recs[i].bwa_seq[0], &b->recs[i].bwa_seq[1]};
			bwa_paired_sw1(g_bns, g_pac, p, g_pe, ii_of(b->iinfos, &b->recs[i]), dummy_tot, dummy_mapped);
		}
	t_sw_mode = RR_REAL;
	t_swq = 0;
}

static void rescue_replay_slice(int s, size_t lo, size_t hi, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	size_t i;
	t_sw_jobs = b->sw_jobs + b->sw_off[s]; t_sw_res = b->sw_res + b->sw_off[s]; t_sw_pool = b->sw_pool;
	t_sw_n = b->sw_off[s + 1] - b->sw_off[s]; t_sw_pos = 0;
	t_sw_mode = RR_REPLAY;
	b->n_tot[s][0] = b->n_tot[s][1] = b->n_mapped[s][0] = b->n_mapped[s][1] = 0;
	for (i = lo; i < hi; ++i)
		if (is_pair_job(&b->recs[i])) {
			bwa_seq_t *p[2] = {&b->recs[i].bwa_seq[0], &b->recs[i].bwa_seq[1]};
			bwa_paired_sw1(g_bns, g_pac, p, g_pe, ii_of(b->iinfos, &b->recs[i]), b->n_tot[s], b->n_mapped[s]);
		}
	t_sw_mode = RR_REAL;
	if (t_sw_pos != t_sw_n) { fprintf(stderr, "[bwa_gpu_batch] %zu of %zu SW answers unused\n", t_sw_n - t_sw_pos, t_sw_n); abort(); }
}

static void stage_rescue(batch2_t *b, uint64_t n_tot[2], uint64_t n_mapped[2])
{
	size_t total = 0, i;
	int s, ns;
	double t0;
	t_cpu_bucket = CPU_RESCUE_RECORD;
	ns = parallel_slices(b->n, 2048, rescue_record_slice, b);
	b->n_slices = ns;
	for (s = 0; s < ns; ++s) { b->sw_off[s] = total; total += b->swq[s].n; }
	b->sw_off[ns] = total;
	if (total + 1 > b->m_sw_jobs) { b->m_sw_jobs = total + 1 + total / 4; b->sw_jobs = (bwa_gpu_sw_job_t *)realloc(b->sw_jobs, b->m_sw_jobs * sizeof(*b->sw_jobs)); }
	if (total + 1 > b->m_sw_res) { b->m_sw_res = total + 1 + total / 4; b->sw_res = (bwa_gpu_path_res_t *)realloc(b->sw_res, b->m_sw_res * sizeof(*b->sw_res)); }
	for (s = 0; s < ns; ++s)
		for (i = 0; i < b->swq[s].n; ++i) {
			b->sw_jobs[b->sw_off[s] + i] = b->swq[s].job[i];
			b->sw_jobs[b->sw_off[s] + i].seq = b->swq[s].seqs + b->swq[s].seq_off[i];
		}
	t0 = now();
	b->sw_pool = 0;
	if (total) {
		size_t pool_n = 0;
		if (bwa_gpu_mate_sw_path((int)total, b->sw_jobs, b->sw_res, &b->sw_pool)) die("bwa_gpu_mate_sw_path");
		/* the pool belongs to the library until its next call, which another stage's thread may make: keep a copy */
		for (i = 0; i < total; ++i)
			if (b->sw_res[i].n_cigar && (size_t)(b->sw_res[i].cigar_off + b->sw_res[i].n_cigar) > pool_n) pool_n = (size_t)(b->sw_res[i].cigar_off + b->sw_res[i].n_cigar);
		if (pool_n + 1 > b->m_sw_pool) { b->m_sw_pool = pool_n + 1 + pool_n / 4; b->sw_pool_copy = (bwa_cigar_t *)realloc(b->sw_pool_copy, b->m_sw_pool * sizeof(bwa_cigar_t)); }
		memcpy(b->sw_pool_copy, b->sw_pool, pool_n * sizeof(bwa_cigar_t));
		b->sw_pool = b->sw_pool_copy;
	}
	REP_ADD(calls_sw, jobs_sw, dev_sw_s, total, t0);
	t_cpu_bucket = CPU_RESCUE_REPLAY;
	parallel_slices(b->n, 2048, rescue_replay_slice, b);
	for (s = 0; s < ns; ++s) { n_tot[0] += b->n_tot[s][0]; n_tot[1] += b->n_tot[s][1]; n_mapped[0] += b->n_mapped[s][0]; n_mapped[1] += b->n_mapped[s][1]; }
}

/* D: bwa_refine_gapped + bwa_update_bam1 (bam2bam.c:643-658, 798-810).  The banded global alignments of the gapped hits
 * (refine_gapped_core -> aln_global_core, bwase.c:212) go to the device in one call: RECORD on a scratch copy of each read
 * (bwa_refine_gapped edits pos / cigar / multi[] and reverses seq in place; all of that is undone or thrown away), REPLAY = the
 * reference's own tail of finish_*. */
static void refine_record_read(bwa_seq_t *s)
{
	bwa_seq_t tmp;
	bwt_multi1_t *multi_copy = 0;
	int j, gapped = 0;
	if (s->type != BWA_TYPE_NO_MATCH && s->type != BWA_TYPE_MATESW && s->n_gapo) gapped = 1;
	for (j = 0; j < s->n_multi && !gapped; ++j) gapped = s->multi[j].gap != 0;
	if (!gapped) return; /* bwa_refine_gapped would not call aln_global_core (bwase.c:371-380) */
	tmp = *s;
	if (s->cigar && s->n_cigar > 0) { /* bwa_correct_trimmed may realloc the copy's CIGAR (bwase.c:330-352): never the original's block */
		tmp.cigar = (bwa_cigar_t *)malloc((size_t)s->n_cigar * sizeof(bwa_cigar_t));
		memcpy(tmp.cigar, s->cigar, (size_t)s->n_cigar * sizeof(bwa_cigar_t));
	}
	if (s->n_multi) {
		multi_copy = (bwt_multi1_t *)malloc((size_t)s->n_multi * sizeof(bwt_multi1_t));
		memcpy(multi_copy, s->multi, (size_t)s->n_multi * sizeof(bwt_multi1_t));
		tmp.multi = multi_copy;
	}
	bwa_refine_gapped(g_bns, 1, &tmp, g_pac, 0);
	seq_reverse(s->len, s->seq, 0); /* bwase.c:369 reversed the shared buffer: put it back for the real run */
	for (j = 0; j < s->n_multi; ++j)
		if (multi_copy[j].cigar && multi_copy[j].cigar != s->multi[j].cigar) free(multi_copy[j].cigar);
	free(multi_copy);
	if (tmp.cigar != s->cigar) free(tmp.cigar);
	if (tmp.md != s->md) free(tmp.md);
}

static void refine_record_slice(int s, size_t lo, size_t hi, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	size_t i;
	b->gaq[s].n = 0; b->gaq[s].bn = 0;
	t_gaq = &b->gaq[s];
	t_ga_mode = RR_RECORD;
	for (i = lo; i < hi; ++i) {
		bam_pair_t *r = &b->recs[i];
		if (is_pair_job(r)) { refine_record_read(&r->bwa_seq[0]); refine_record_read(&r->bwa_seq[1]); }
		else if (is_single_job(r)) refine_record_read(&r->bwa_seq[0]);
	}
	t_ga_mode = RR_REAL;
	t_gaq = 0;
}

static void refine_replay_slice(int s, size_t lo, size_t hi, void *ctx)
{
	batch2_t *b = (batch2_t *)ctx;
	size_t i;
	t_ga_res = b->ga_res + b->ga_off[s]; t_ga_pool = b->ga_pool;
	t_ga_len1 = b->ga_len1 + b->ga_off[s]; t_ga_len2 = b->ga_len2 + b->ga_off[s];
	t_ga_n = b->ga_off[s + 1] - b->ga_off[s]; t_ga_pos = 0;
	t_ga_mode = RR_REPLAY;
	for (i = lo; i < hi; ++i) {
		bam_pair_t *r = &b->recs[i];
		if (is_pair_job(r)) {
			bwa_refine_gapped(g_bns, 1, &r->bwa_seq[0], g_pac, 0);
			bwa_refine_gapped(g_bns, 1, &r->bwa_seq[1], g_pac, 0);
		} else if (is_single_job(r)) bwa_refine_gapped(g_bns, 1, &r->bwa_seq[0], g_pac, 0);
	}
	t_ga_mode = RR_REAL;
	if (t_ga_pos != t_ga_n) { fprintf(stderr, "[bwa_gpu_batch] %zu of %zu global-alignment answers unused\n", t_ga_n - t_ga_pos, t_ga_n); abort(); }
}

/* E: bwa_update_bam1 (bam2bam.c:804-810), a stage of its own so that it overlaps the next batch's refine */
static void update_one(size_t i, void *ctx)
{
	bam_pair_t *r = (bam_pair_t *)ctx + i;
	if (is_pair_job(r)) {
		bwa_update_bam1(&r->bam_rec[0], g_bns, &r->bwa_seq[0], &r->bwa_seq[1], g_gap->mode, g_gap->max_top2);
		bwa_update_bam1(&r->bam_rec[1], g_bns, &r->bwa_seq[1], &r->bwa_seq[0], g_gap->mode, g_gap->max_top2);
		bwa_free_read_seq1(&r->bwa_seq[1]);
		bwa_free_read_seq1(&r->bwa_seq[0]);
	} else if (is_single_job(r)) {
		bwa_seq_t *p = &r->bwa_seq[0];
		bwa_update_bam1(&r->bam_rec[0], g_bns, p, 0, g_gap->mode, g_gap->max_top2);
		bwa_free_read_seq1(p);
	}
	if (r->kind != eof_marker && r->phase == positioned) r->phase = finished;
}

static void stage_update(batch2_t *b)
{
	t_cpu_bucket = CPU_UPDATE;
	parallel_for(b->n, 1024, update_one, b->recs);
}

static void stage_refine(batch2_t *b)
{
	size_t total = 0, i;
	int s, ns;
	double t0;
	t_cpu_bucket = CPU_TOSEQ;
	parallel_for(b->n, 1024, pair_to_seq_one, b->recs); /* singletons reach this stage first */
	t_cpu_bucket = CPU_REFINE_RECORD;
	ns = parallel_slices(b->n, 2048, refine_record_slice, b);
	for (s = 0; s < ns; ++s) { b->ga_off[s] = total; total += b->gaq[s].n; }
	b->ga_off[ns] = total;
	if (total + 1 > b->m_ga_jobs) { b->m_ga_jobs = total + 1 + total / 4; b->ga_jobs = (bwa_gpu_ga_job_t *)realloc(b->ga_jobs, b->m_ga_jobs * sizeof(*b->ga_jobs)); }
	if (total + 1 > b->m_ga_res) { b->m_ga_res = total + 1 + total / 4; b->ga_res = (bwa_gpu_path_res_t *)realloc(b->ga_res, b->m_ga_res * sizeof(*b->ga_res)); }
	if (total + 1 > b->m_ga_len) { b->m_ga_len = total + 1 + total / 4; b->ga_len1 = (int *)realloc(b->ga_len1, b->m_ga_len * sizeof(int)); b->ga_len2 = (int *)realloc(b->ga_len2, b->m_ga_len * sizeof(int)); }
	for (s = 0; s < ns; ++s)
		for (i = 0; i < b->gaq[s].n; ++i) {
			bwa_gpu_ga_job_t *jb = &b->ga_jobs[b->ga_off[s] + i];
			jb->ref = b->gaq[s].bytes + b->gaq[s].off1[i]; jb->reflen = b->gaq[s].len1[i];
			jb->seq = b->gaq[s].bytes + b->gaq[s].off2[i]; jb->len = b->gaq[s].len2[i];
			b->ga_len1[b->ga_off[s] + i] = jb->reflen; b->ga_len2[b->ga_off[s] + i] = jb->len;
		}
	t0 = now();
	b->ga_pool = 0;
	if (total) {
		size_t pool_n = 0;
		if (bwa_gpu_global_align_seqs((int)total, b->ga_jobs, aln_param_bwa.gap_end, aln_param_bwa.band_width, b->ga_res, &b->ga_pool))
			die("bwa_gpu_global_align_seqs");
		/* the pool is the library's until its next SW / global-alignment call, which the rescue stage of the next batch may make */
		for (i = 0; i < total; ++i)
			if (b->ga_res[i].n_cigar && (size_t)(b->ga_res[i].cigar_off + b->ga_res[i].n_cigar) > pool_n) pool_n = (size_t)(b->ga_res[i].cigar_off + b->ga_res[i].n_cigar);
		if (pool_n + 1 > b->m_ga_pool) { b->m_ga_pool = pool_n + 1 + pool_n / 4; b->ga_pool_copy = (bwa_cigar_t *)realloc(b->ga_pool_copy, b->m_ga_pool * sizeof(bwa_cigar_t)); }
		memcpy(b->ga_pool_copy, b->ga_pool, pool_n * sizeof(bwa_cigar_t));
		b->ga_pool = b->ga_pool_copy;
	}
	REP_ADD(calls_ga, jobs_ga, dev_ga_s, total, t0);
	t_cpu_bucket = CPU_REFINE_REPLAY;
	parallel_slices(b->n, 2048, refine_replay_slice, b);
}

/* Pass 2 as a pipeline: load -> enumerate (A) -> pairing + XA (B) -> mate rescue (C) -> refine + update (D) -> write,
 * one thread per stage, batches in order.  A owns the position cache, B the random numbers, the writer the output file. */
#define P2_SLOTS 10
#define N_DESTROY 2 /* threads freeing finished batches, whole batches in turn (free()s inside ONE batch on several threads fight over the arenas) */
enum { S2_FREE = 0, S2_LOADED, S2_ENUM, S2_PAIRED, S2_RESCUED, S2_REFINED, S2_UPDATED, S2_WRITTEN };
typedef struct {
	pthread_mutex_t mu;
	pthread_cond_t cv;
	int state[P2_SLOTS];
	batch2_t b[P2_SLOTS];
	size_t B;
	gzFile temporary;
	BGZF *output;
	khash_t(isize_infos) *iinfos;
	kh_64_t *my_hash;
	uint64_t n_tot[2], n_mapped[2];
	double t0, t_load, t_enum, t_pair, t_rescue, t_refine, t_update, t_write, t_destroy[N_DESTROY];
	long tot_seqs;
	int destroy_done;
} pipe2_t;
#define P2_WAIT(P, slot, want) slot_wait(&(P)->mu, &(P)->cv, &(P)->state[slot], want)
#define P2_SET(P, slot, st) slot_set(&(P)->mu, &(P)->cv, &(P)->state[slot], st)

#define P2_STAGE(name, from, to, timer, body) \
	static void *name(void *arg) \
	{ \
		pipe2_t *P = (pipe2_t *)arg; \
		unsigned q; \
		for (q = 0;; ++q) { \
			const int slot = (int)(q % P2_SLOTS); \
			batch2_t *b = &P->b[slot]; \
			double t1; \
			P2_WAIT(P, slot, from); \
			t1 = now(); \
			if (b->n) { body; } \
			P->timer += now() - t1; \
			P2_SET(P, slot, to); \
			if (b->n == 0) break; \
		} \
		return 0; \
	}

P2_STAGE(stage2_enum, S2_LOADED, S2_ENUM, t_enum, stage_enumerate(b, P->my_hash))
P2_STAGE(stage2_pair, S2_ENUM, S2_PAIRED, t_pair, stage_pairing(b))
P2_STAGE(stage2_rescue, S2_PAIRED, S2_RESCUED, t_rescue, stage_rescue(b, P->n_tot, P->n_mapped))
P2_STAGE(stage2_refine, S2_RESCUED, S2_REFINED, t_refine, stage_refine(b))
P2_STAGE(stage2_update, S2_REFINED, S2_UPDATED, t_update, stage_update(b))

static void *stage2_write(void *arg)
{
	pipe2_t *P = (pipe2_t *)arg;
	unsigned q;
	for (q = 0;; ++q) {
		const int slot = (int)(q % P2_SLOTS);
		batch2_t *b = &P->b[slot];
		double t1;
		P2_WAIT(P, slot, S2_UPDATED);
		if (b->n == 0) { P2_SET(P, slot, S2_WRITTEN); break; }
		t1 = now();
		write_records_bam(P->output, b->recs, b->n);
		P->t_write += now() - t1;
		P->tot_seqs += b->seqs;
		fprintf(stderr, "[sequential_loop_pass2] %ld sequences processed in %.2f sec\n", P->tot_seqs, now() - P->t0);
		P2_SET(P, slot, S2_WRITTEN);
	}
	return 0;
}

typedef struct { pipe2_t *P; int id; } destroy2_arg_t;
static void *stage2_destroy(void *arg) /* the free()s of a written batch, off the writer's thread */
{
	destroy2_arg_t *A = (destroy2_arg_t *)arg;
	pipe2_t *P = A->P;
	unsigned q;
	for (q = (unsigned)A->id;; q += N_DESTROY) {
		const int slot = (int)(q % P2_SLOTS);
		batch2_t *b = &P->b[slot];
		double t1;
		pthread_mutex_lock(&P->mu);
		while (P->state[slot] != S2_WRITTEN && !P->destroy_done) pthread_cond_wait(&P->cv, &P->mu);
		if (P->state[slot] != S2_WRITTEN) { pthread_mutex_unlock(&P->mu); break; } /* the other thread met the end marker */
		if (b->n == 0) { P->destroy_done = 1; pthread_cond_broadcast(&P->cv); pthread_mutex_unlock(&P->mu); break; }
		pthread_mutex_unlock(&P->mu);
		t1 = now();
		destroy_records(b->recs, b->n);
		P->t_destroy[A->id] += now() - t1;
		P2_SET(P, slot, S2_FREE);
	}
	return 0;
}

void sequential_loop_pass2(gzFile temporary, BGZF *output, khash_t(isize_infos) *iinfos)
{
	const double t_begin_ = (g_t_pass2_begin = now());
	g_trace_pass = 2;
	const size_t B = batch_records();
	const long long max_q = getenv("BWAGPU_BATCH_SA") ? atoll(getenv("BWAGPU_BATCH_SA")) : 1ll << 25; /* SA rows per device call */
	pipe2_t *P = (pipe2_t *)calloc(1, sizeof(pipe2_t));
	bam_pair_t *stash = (bam_pair_t *)calloc(B, sizeof(bam_pair_t)); /* loaded records not yet handed to a batch */
	size_t stash_n = 0, stash_at = 0;
	pthread_t th[6 + N_DESTROY];
	destroy2_arg_t da[N_DESTROY];
	khiter_t it;
	unsigned q;
	int s, eof = 0;
	long dummy = 0;
	pthread_mutex_init(&P->mu, 0); pthread_cond_init(&P->cv, 0);
	P->B = B; P->temporary = temporary; P->output = output; P->iinfos = iinfos; P->my_hash = kh_init(64); P->t0 = now();
	for (s = 0; s < P2_SLOTS; ++s) P->b[s].recs = (bam_pair_t *)calloc(B, sizeof(bam_pair_t));
	ensure_gpu();
	pthread_create(&th[0], 0, stage2_enum, P);
	pthread_create(&th[1], 0, stage2_pair, P);
	pthread_create(&th[2], 0, stage2_rescue, P);
	pthread_create(&th[3], 0, stage2_refine, P);
	pthread_create(&th[4], 0, stage2_update, P);
	pthread_create(&th[5], 0, stage2_write, P);
	for (s = 0; s < N_DESTROY; ++s) { da[s].P = P; da[s].id = s; pthread_create(&th[6 + s], 0, stage2_destroy, &da[s]); }
	for (q = 0;; ++q) { /* the load stage: batches bounded by records and by the SA rows their hit lists expand to */
		const int slot = (int)(q % P2_SLOTS);
		batch2_t *b = &P->b[slot];
		long long rows = 0;
		double t1;
		P2_WAIT(P, slot, S2_FREE);
		t1 = now();
		b->n = 0; b->seqs = 0; b->iinfos = iinfos;
		const size_t Bq = tail_records(ramp_records(B, q), B, memtemp_spilled() ? 0 : memtemp_left() + (stash_n - stash_at));
		while (b->n < Bq) {
			if (stash_at == stash_n) {
				if (eof) break;
				stash_n = load_records(temporary, stash, B, &dummy);
				stash_at = 0;
				if (stash_n == 0) { eof = 1; break; }
			}
			{
				const bam_pair_t *r = &stash[stash_at];
				const long long w = is_pair_job(r) && wants_pairing(r) ? n_rows(r) : 0;
				if (b->n && rows + w > max_q) break;
				rows += w;
				b->seqs += r->kind;
				b->recs[b->n++] = stash[stash_at++];
			}
		}
		P->t_load += now() - t1;
		P2_SET(P, slot, S2_LOADED);
		if (b->n == 0) break;
	}
	for (s = 0; s < 6 + N_DESTROY; ++s) pthread_join(th[s], 0);
	g_rep.pass2_s = now() - P->t0;
	fprintf(stderr, "[%s] %ld sequences processed in %.2f sec (pipelined stages, busy seconds each incl. their device calls: load %.2f, enumerate %.2f, "
	                "pairing/XA %.2f, mate rescue %.2f, refine %.2f, update %.2f, BAM write %.2f, destroy %.2f + %.2f)\n"
	                "[%s] finished cleanly, shutting down.\n"
	                "[bwa_paired_sw] %lld out of %lld Q%d singletons are mated.\n"
	                "[bwa_paired_sw] %lld out of %lld Q%d discordant pairs are fixed.\n",
	        __func__, P->tot_seqs, now() - P->t0, P->t_load, P->t_enum, P->t_pair, P->t_rescue, P->t_refine, P->t_update, P->t_write, P->t_destroy[0], P->t_destroy[N_DESTROY - 1], __func__,
	        (long long)P->n_mapped[1], (long long)P->n_tot[1], SW_MIN_MAPQ, (long long)P->n_mapped[0], (long long)P->n_tot[0], SW_MIN_MAPQ);
	for (it = kh_begin(P->my_hash); it != kh_end(P->my_hash); ++it)
		if (kh_exist(P->my_hash, it)) free(kh_val(P->my_hash, it).a);
	kh_destroy(64, P->my_hash);
	for (s = 0; s < P2_SLOTS; ++s) { bam_pair_t *recs = P->b[s].recs; batch2_free(&P->b[s]); free(recs); }
	free(stash);
	pthread_mutex_destroy(&P->mu); pthread_cond_destroy(&P->cv);
	free(P);
	memtemp_free();
	if (!g_keep) memtemp_release(); /* a host that keeps its index between runs keeps the record chunks too */
	(void)t_begin_;
	g_t_pass2_end = now();
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
 * in record order in both passes and the BAM equals `bam2bam -t 1`'s (tests/test_batched_bam2bam.py), which the
 * reference's own `-t N` does not guarantee.
 */
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
	batch2_t *b2 = (batch2_t *)calloc(1, sizeof(batch2_t));
	saq_t q1;
	size_t *qoff = 0, qoff_cap = 0;
	khiter_t it;
	uint64_t n_tot[2] = {0, 0}, n_mapped[2] = {0, 0};
	long n_batches = 0, n_records = 0, n_dupes = 0, failure_count = 0;
	double t_toseq = 0, t_host = 0, t_fin = 0, t_wait = 0, t_send = 0;
	int done = 0, have_pending = 0;
	bam_pair_t pending;
	void *upstream;
	static int claimed;
	(void)arg;
	memset(&q1, 0, sizeof(q1));

	/* the device queue is one: ONE thread batches, whatever -t says; further threads have nothing to add and leave */
	if (__sync_lock_test_and_set(&claimed, 1)) {
		fprintf(stderr, "[run_worker_thread] a batching GPU worker is already running in this process; extra thread exits.\n");
		free(recs); free(flat); free(taken); kh_destroy(64, my_hash); free(b2);
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
		if (phase_of_batch == pristine) {
			align_range(recs, n, flat, &t_toseq, 0);
			position_range(recs, n, &q1, &qoff, &qoff_cap, &t_host);
		} else if (phase_of_batch == aligned) for (i = 0; i < n; ++i) pair_posn(&recs[i]);
		else if (phase_of_batch == positioned) {
			khash_t(isize_infos) *iinfos = (khash_t(isize_infos) *)g_iinfos_seen;
			if (!iinfos) failure_count += (long)n; /* sent back as they came (bam2bam.c:1419-1420) */
			else
				for (lo = 0; lo < n;) { /* sub-ranges bounded by the SA rows their hit lists expand to, as sequential_loop_pass2 */
					size_t hi = lo;
					long long rows = 0;
					while (hi < n) {
						const bam_pair_t *r = &recs[hi];
						const long long w = is_pair_job(r) && wants_pairing(r) ? n_rows(r) : 0;
						if (hi > lo && rows + w > max_q) break;
						rows += w;
						++hi;
					}
					b2->recs = recs + lo; b2->n = hi - lo; b2->iinfos = iinfos;
					stage_enumerate(b2, my_hash);
					stage_pairing(b2);
					stage_rescue(b2, n_tot, n_mapped);
					stage_refine(b2);
					stage_update(b2);
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
	b2->recs = 0; batch2_free(b2); free(b2);
	saq_free(&q1); free(qoff);
	free(recs); free(flat); free(taken);
	zmq_close(upstream);
	return 0;
}

/* ------------------------------------------------------------------ for hosts that run bam2bam in-process */
int bwa_gpu_batch_report_size(void) { return (int)sizeof(bwa_gpu_batch_report_t); }

int bwa_gpu_batch_last_report(bwa_gpu_batch_report_t *out)
{
	if (!out) return 1;
	*out = g_rep;
	return 0;
}
