/* TEST INFRASTRUCTURE ONLY -- multi-threaded driver around the UNMODIFIED reference
 * functions, linked against oracle/_ref/libbwaref.so.  Used (a) by bench.py's
 * cpu_baseline / --impl reference legs to time the reference's CPU path on all host
 * cores without the 0MQ per-record mux, and (b) by tests as a batch front-end.
 *
 * Built both as a program (prints its usage) and, with -DREFH_SHARED, as
 * libref_harness.so for ctypes.  It only *calls* reference functions:
 *   bwa_cal_sa_reg_gap  (bwtaln.c:93)   with n_seqs = 1, as bam2bam.c:616/676 does
 *   bwt_sa              (bwt.c:72)
 *   aln_local_core      (stdaln.c:529)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include "bwtaln.h"
#include "bwt.h"
#include "stdaln.h"

typedef struct {
	bwt_t *const *bwt;
	int n, tid, nthreads;
	bwa_seq_t *seqs;
	const gap_opt_t *opt;
} aln_job_t;

static void *aln_worker(void *p)
{
	aln_job_t *j = (aln_job_t *)p;
	int i;
	/* interleaved assignment: neighbouring reads go to different threads, so a
	 * sorted/clustered input still balances */
	for (i = j->tid; i < j->n; i += j->nthreads)
		bwa_cal_sa_reg_gap(j->bwt, 1, j->seqs + i, j->opt);
	return 0;
}

/* per-read semantics of bam2bam (n_seqs = 1 per call), spread over nthreads */
int refh_aln_batch(bwt_t *const bwt[2], int n, bwa_seq_t *seqs, const gap_opt_t *opt, int nthreads)
{
	int t;
	pthread_t *th;
	aln_job_t *jobs;
	if (nthreads < 1) nthreads = 1;
	th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
	jobs = (aln_job_t *)calloc(nthreads, sizeof(aln_job_t));
	for (t = 0; t < nthreads; ++t) {
		jobs[t].bwt = bwt; jobs[t].n = n; jobs[t].tid = t; jobs[t].nthreads = nthreads;
		jobs[t].seqs = seqs; jobs[t].opt = opt;
		pthread_create(&th[t], 0, aln_worker, &jobs[t]);
	}
	for (t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
	free(th); free(jobs);
	return 0;
}

/* free() the aln arrays a batch left behind (the aln part of bwa_free_read_seq1, bwaseqio.c:259) */
void refh_free_alns(int n, bwa_seq_t *seqs)
{
	int i;
	for (i = 0; i < n; ++i) { free(seqs[i].aln); seqs[i].aln = 0; seqs[i].n_aln = 0; }
}

typedef struct {
	bwt_t *const *bwt;
	int64_t n;
	int tid, nthreads;
	const bwtint_t *k;
	const uint8_t *which;
	bwtint_t *out;
} sa_job_t;

static void *sa_worker(void *p)
{
	sa_job_t *j = (sa_job_t *)p;
	int64_t i;
	for (i = j->tid; i < j->n; i += j->nthreads)
		j->out[i] = bwt_sa(j->bwt[j->which[i] ? 0 : 1], j->k[i]);
	return 0;
}

/* which[i] != 0 -> forward index bwt[0] (.bwt/.sa), else the reverse index bwt[1] */
int refh_sa_batch(bwt_t *const bwt[2], int64_t n, const bwtint_t *k, const uint8_t *which, bwtint_t *out, int nthreads)
{
	int t;
	pthread_t *th;
	sa_job_t *jobs;
	if (nthreads < 1) nthreads = 1;
	th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
	jobs = (sa_job_t *)calloc(nthreads, sizeof(sa_job_t));
	for (t = 0; t < nthreads; ++t) {
		jobs[t].bwt = bwt; jobs[t].n = n; jobs[t].tid = t; jobs[t].nthreads = nthreads;
		jobs[t].k = k; jobs[t].which = which; jobs[t].out = out;
		pthread_create(&th[t], 0, sa_worker, &jobs[t]);
	}
	for (t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
	free(th); free(jobs);
	return 0;
}

/* One aln_local_core call exactly as bwa_sw_core makes it (bwape.c:456):
 * returns score; fills start/end (1-based path coordinates) and the path. */
int refh_sw1(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
             int *path_len, int32_t *path_ijc /* 3*(l_ref+l_query) ints */)
{
	path_t *path = (path_t *)calloc(l_ref + l_query + 2, sizeof(path_t));
	int i, score;
	*path_len = 0;
	score = aln_local_core((unsigned char *)ref, l_ref, (unsigned char *)query, l_query,
	                       &aln_param_bwa, path, path_len, 1, 0);
	for (i = 0; i < *path_len; ++i) {
		path_ijc[3*i] = path[i].i; path_ijc[3*i+1] = path[i].j; path_ijc[3*i+2] = path[i].ctype;
	}
	free(path);
	return score;
}

/* Forward score + start/end cells of aln_local_core without the path-filling third pass:
 * path_len == NULL makes it store (start_i,start_j),(end_i,end_j) in path[0..1]
 * (stdaln.c:708-712); _thres = 0 keeps the NULL path_len from being dereferenced at
 * stdaln.c:634.  Passes 1 and 2 do not depend on _thres. */
int refh_sw_ends(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query, int32_t out[4])
{
	path_t path[2];
	int score;
	memset(path, 0, sizeof path);
	score = aln_local_core((unsigned char *)ref, l_ref, (unsigned char *)query, l_query, &aln_param_bwa, path, 0, 0, 0);
	out[0] = path[0].i; out[1] = path[0].j; out[2] = path[1].i; out[3] = path[1].j;
	return score;
}

/* aln_global_core with aln_param_bwa's scores and the given gap_end / band (stdaln.c:345) */
int refh_global(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query, int gap_end, int band,
                int *path_len, int32_t *path_ijc)
{
	path_t *path = (path_t *)calloc(l_ref + l_query + 2, sizeof(path_t));
	AlnParam ap = aln_param_bwa;
	int i, score;
	ap.gap_end = gap_end; ap.band_width = band;
	*path_len = 0;
	score = aln_global_core((unsigned char *)ref, l_ref, (unsigned char *)query, l_query, &ap, path, path_len);
	for (i = 0; i < *path_len; ++i) {
		path_ijc[3*i] = path[i].i; path_ijc[3*i+1] = path[i].j; path_ijc[3*i+2] = path[i].ctype;
	}
	free(path);
	return score;
}

typedef struct {
	int n, tid, nthreads;
	const uint8_t *refs, *queries;
	const int64_t *ref_off, *q_off;
	int32_t *out; /* 5 per job: score, start_i, start_j, end_i, end_j */
} sw_job_t;

static void *sw_worker(void *p)
{
	sw_job_t *j = (sw_job_t *)p;
	int i;
	for (i = j->tid; i < j->n; i += j->nthreads)
		j->out[5 * i] = refh_sw_ends(j->refs + j->ref_off[i], (int)(j->ref_off[i + 1] - j->ref_off[i]),
		                             j->queries + j->q_off[i], (int)(j->q_off[i + 1] - j->q_off[i]), j->out + 5 * i + 1);
	return 0;
}

int refh_sw_batch(int n, const uint8_t *refs, const int64_t *ref_off, const uint8_t *queries, const int64_t *q_off,
                  int32_t *out, int nthreads)
{
	int t;
	pthread_t *th;
	sw_job_t *jobs;
	if (nthreads < 1) nthreads = 1;
	th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
	jobs = (sw_job_t *)calloc(nthreads, sizeof(sw_job_t));
	for (t = 0; t < nthreads; ++t) {
		jobs[t].n = n; jobs[t].tid = t; jobs[t].nthreads = nthreads;
		jobs[t].refs = refs; jobs[t].queries = queries; jobs[t].ref_off = ref_off; jobs[t].q_off = q_off; jobs[t].out = out;
		pthread_create(&th[t], 0, sw_worker, &jobs[t]);
	}
	for (t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
	free(th); free(jobs);
	return 0;
}

#ifndef REFH_SHARED
int main(void)
{
	fprintf(stderr, "ref_harness: build with -DREFH_SHARED and load libref_harness.so via ctypes\n"
	                "(see tests/refload.py); entry points refh_aln_batch, refh_sa_batch, refh_sw1\n");
	return 0;
}
#endif
