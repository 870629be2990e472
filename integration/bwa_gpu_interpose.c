/* bwa_gpu_interpose.c -- the drop-in seam in its most literal form.
 *
 * A shared object that defines the reference's OWN function names at the hot-path boundary
 *  and forwards them to libbwagpu.so.  Pre-loaded into an unmodified build of the
 * reference (LD_PRELOAD, or linked ahead of its objects), it makes `bwa bam2bam` / `bwa worker`
 * run their alignment hot path on the GPU without touching a line of bam2bam.c:
 *
 *   bwa_cal_sa_reg_gap (bwtaln.c:93, called at bam2bam.c:616,676)  -> bwa_gpu_cal_sa_reads_gap
 *   bwt_sa             (bwt.c:72, called at bam2bam.c:635-636,752,761,786; bwase.c:145,152)
 *                                                                   -> bwa_gpu_cal_pac_pos
 *   bwa_sw_core        (bwape.c:433, called at bwape.c:577)         -> the reference's own code runs
 *        and the device result of bwa_gpu_mate_sw_path (K5 + K6) for the same job is CHECKED against the
 *        reference's aln_local_core: score, path end points and CIGAR; mismatches are counted and
 *        reported at exit (the reference's own CIGAR is what goes into the BAM).
 *
 * Because the reference calls these per record (n_seqs = 1), this is the CORRECTNESS path -- every
 * call is a device round trip.  Throughput needs the batched driver of INTEGRATION.md.  Used by
 * tests/test_dropin_bam2bam.py to show record-identical BAM output.
 *
 * Build: gcc -O2 -fPIC -shared -I ../include -o libbwa_gpu_interpose.so bwa_gpu_interpose.c -ldl \
 *            -L../network-aware-bwa_b200 -lbwagpu
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include "bwa_gpu.h"

int bwa_gpu_load_pac(const ubyte_t *pac, int64_t l_pac);

static bwt_t *g_bwt[2];
static int g_ready, g_pac_ready;
static long g_n_aln_calls, g_n_sa_calls, g_n_sw_calls, g_n_sw_checked, g_n_sw_mismatch;

static void die(const char *what)
{
	fprintf(stderr, "[bwa_gpu_interpose] %s: %s\n", what, bwa_gpu_last_error());
	abort(); /* the reference's own convention on this path: xassert -> abort (utils.c:68-83) */
}

static void report(void)
{
	fprintf(stderr, "[bwa_gpu_interpose] calls: cal_sa_reg_gap=%ld bwt_sa=%ld sw_core=%ld; SW checked=%ld mismatches=%ld\n",
	        g_n_aln_calls, g_n_sa_calls, g_n_sw_calls, g_n_sw_checked, g_n_sw_mismatch);
	if (g_n_sw_mismatch) _exit(97);
}

static void ensure_index(bwt_t *const bwt[2])
{
	if (g_ready) return;
	if (bwa_gpu_init(0, 0)) die("bwa_gpu_init");
	if (bwa_gpu_load_index(bwt, 0, 0)) die("bwa_gpu_load_index");
	g_bwt[0] = bwt[0]; g_bwt[1] = bwt[1];
	g_ready = 1;
	atexit(report);
}

void bwa_cal_sa_reg_gap(bwt_t *const bwt[2], int n_seqs, bwa_seq_t *seqs, const gap_opt_t *opt)
{
	int i;
	ensure_index(bwt);
	++g_n_aln_calls;
	/* The reference derives max_diff and the max_gapo clamp from the LONGEST read of the call
	 * (bwtaln.c:100-103); bam2bam always passes n_seqs = 1.  The batch entry point has per-read
	 * semantics, so a multi-read call is forwarded read by read only if lengths differ. */
	for (i = 1; i < n_seqs; ++i)
		if (seqs[i].len != seqs[0].len) break;
	if (i >= n_seqs) { if (bwa_gpu_cal_sa_reads_gap(n_seqs, seqs, opt)) die("bwa_gpu_cal_sa_reads_gap"); }
	else { fprintf(stderr, "[bwa_gpu_interpose] mixed-length multi-read call is not supported\n"); abort(); }
}

bwtint_t bwt_sa(const bwt_t *bwt, bwtint_t k)
{
	static bwtint_t (*real)(const bwt_t *, bwtint_t);
	uint8_t which;
	bwtint_t out;
	if (!g_ready || (bwt != g_bwt[0] && bwt != g_bwt[1])) { /* e.g. index construction: not our index */
		if (!real) real = (bwtint_t(*)(const bwt_t *, bwtint_t))dlsym(RTLD_NEXT, "bwt_sa");
		return real(bwt, k);
	}
	++g_n_sa_calls;
	which = bwt == g_bwt[0];
	if (bwa_gpu_cal_pac_pos(1, &k, &which, &out)) die("bwa_gpu_cal_pac_pos");
	return out;
}

/* stdaln.h path_t / AlnParam are not needed here: aln_local_core is wrapped through void pointers. */
#define MAX_CHECK_CIGAR 1024
static __thread int t_last_score, t_last_end_i, t_last_end_j, t_last_start_i, t_last_start_j, t_have_last, t_last_n_cigar;
static __thread bwa_cigar_t t_last_cigar[MAX_CHECK_CIGAR];

int aln_local_core(unsigned char *seq1, int len1, unsigned char *seq2, int len2, const void *ap, void *path, int *path_len,
                   int thres, int *subo)
{
	static int (*real)(unsigned char *, int, unsigned char *, int, const void *, void *, int *, int, int *);
	int score;
	if (!real) real = (int (*)(unsigned char *, int, unsigned char *, int, const void *, void *, int *, int, int *))dlsym(RTLD_NEXT, "aln_local_core");
	score = real(seq1, len1, seq2, len2, ap, path, path_len, thres, subo);
	t_have_last = 0;
	if (path && path_len && *path_len > 0) { /* path_t = {int i, j; uchar ctype}: 12 bytes; path[0] = the end cell */
		const int *p = (const int *)path;
		const int n = *path_len;
		int k, nc = 0, last = -1;
		t_last_score = score; t_last_end_i = p[0]; t_last_end_j = p[1];
		t_last_start_i = p[3 * (n - 1)]; t_last_start_j = p[3 * (n - 1) + 1];
		for (k = n - 1; k >= 0; --k) { /* aln_path2cigar32 (stdaln.c:1009-1039): runs of equal ctype, start -> end */
			const int ct = *(const unsigned char *)(p + 3 * k + 2);
			if (ct == last) ++t_last_cigar[nc - 1];
			else if (nc < MAX_CHECK_CIGAR) { t_last_cigar[nc++] = (bwa_cigar_t)(ct << 14 | 1); last = ct; }
		}
		t_last_n_cigar = nc;
		t_have_last = 1;
	}
	return score;
}

bwa_cigar_t *bwa_sw_core(bwtint_t l_pac, const ubyte_t *pacseq, int len, const ubyte_t *seq, int64_t *beg, int reglen,
                         int *n_cigar, uint32_t *cnt)
{
	static bwa_cigar_t *(*real)(bwtint_t, const ubyte_t *, int, const ubyte_t *, int64_t *, int, int *, uint32_t *);
	const int64_t beg0 = *beg;
	bwa_cigar_t *ret;
	if (!real) real = (bwa_cigar_t * (*)(bwtint_t, const ubyte_t *, int, const ubyte_t *, int64_t *, int, int *, uint32_t *))dlsym(RTLD_NEXT, "bwa_sw_core");
	++g_n_sw_calls;
	t_have_last = 0;
	ret = real(l_pac, pacseq, len, seq, beg, reglen, n_cigar, cnt);
	if (g_ready && t_have_last) { /* the reference did run aln_local_core on this job: check the device against it */
		bwa_gpu_sw_job_t job;
		bwa_gpu_path_res_t res;
		const bwa_cigar_t *pool = 0;
		int bad;
		if (!g_pac_ready) { if (bwa_gpu_load_pac(pacseq, (int64_t)l_pac)) die("bwa_gpu_load_pac"); g_pac_ready = 1; }
		job.beg = beg0; job.reglen = reglen; job.len = len; job.seq = seq;
		if (bwa_gpu_mate_sw_path(1, &job, &res, &pool)) die("bwa_gpu_mate_sw_path");
		++g_n_sw_checked;
		/* K5 + K6 against the reference's own aln_local_core on this job: score, path end points, CIGAR */
		bad = res.score != t_last_score || res.end_i != t_last_end_i || res.end_j != t_last_end_j ||
		      res.start_i != t_last_start_i || res.start_j != t_last_start_j || res.n_cigar != t_last_n_cigar ||
		      memcmp(pool + res.cigar_off, t_last_cigar, sizeof(bwa_cigar_t) * (size_t)t_last_n_cigar) != 0;
		if (bad) {
			++g_n_sw_mismatch;
			fprintf(stderr, "[bwa_gpu_interpose] SW mismatch: device (%d; %d,%d-%d,%d; %d ops) reference (%d; %d,%d-%d,%d; %d ops)\n",
			        res.score, res.start_i, res.start_j, res.end_i, res.end_j, res.n_cigar, t_last_score, t_last_start_i,
			        t_last_start_j, t_last_end_i, t_last_end_j, t_last_n_cigar);
		}
	}
	return ret;
}
