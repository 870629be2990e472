/* shim.h -- internal declarations shared by the two translation units of libbwa_gpu_batch.so
 * (bwa_gpu_batch.c: the batched drivers; shim_io.c: BAM input / intermediate records / BAM output). */
#ifndef BWA_GPU_SHIM_H
#define BWA_GPU_SHIM_H

#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <zlib.h>

#include "bamlite.h"
#include "bwtaln.h"
#include "bwase.h"
#include "bwape.h"
#include "khash.h"
#include "kstring.h"
#include "bgzf.h"
#include "zmq.h" /* oracle/zmq_shim: the libzmq ABI the reference is built against */
#include "bwa_gpu.h" /* after bwtaln.h: re-uses the reference's own types */

#define REAL(ret, name, ...) \
	static ret (*real_##name)(__VA_ARGS__); \
	if (!real_##name) real_##name = (ret (*)(__VA_ARGS__))dlsym(RTLD_NEXT, #name)

static inline double shim_now(void)
{
	struct timeval tv;
	gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* ---- CPU accounting: seconds of thread CPU time by activity (a profile of the host side without a profiler) */
enum { CPU_INFLATE = 0, CPU_PARSE, CPU_TOSEQ, CPU_POSN_SERIAL, CPU_POSN_PAR, CPU_ISIZE, CPU_STORE, CPU_DESTROY, CPU_LOAD, CPU_ENUM, CPU_PAIRING,
       CPU_XA_SERIAL, CPU_RESCUE_RECORD, CPU_RESCUE_REPLAY, CPU_REFINE_RECORD, CPU_REFINE_REPLAY, CPU_UPDATE, CPU_BAM_LAYOUT, CPU_DEFLATE, CPU_WRITE, CPU_OTHER, CPU_N };
extern __thread int t_cpu_bucket; /* where the calling thread's parallel_for / parallel_slices workers (and cpu_add) book their time */
double thread_cpu_now(void);
void cpu_add(int bucket, double seconds);

/* ---- host threads (bwa_gpu_batch.c) */
typedef void (*pf_fn)(size_t i, void *ctx);
typedef void (*ps_fn)(int slice, size_t lo, size_t hi, void *ctx);
int shim_threads(void);
void parallel_for(size_t n, size_t grain, pf_fn fn, void *ctx);
/* [0, n) cut into at most shim_threads() contiguous slices, slice s = [lo, hi) handled by one thread; returns the number of slices */
int parallel_slices(size_t n, size_t min_per_slice, ps_fn fn, void *ctx);
int slice_count(size_t n, size_t min_per_slice);

/* ---- the reference's option flags the shim needs (captured in bwa_bam_to_bam, bwa_gpu_batch.c) */
extern int g_broken_input, g_skip_duplicates, g_drop_aligned, g_only_aligned;

/* ---- shim_io.c */
void fastin_set_path(const char *path);   /* bwa_bam_open: the file the next bam_read1 stream comes from */
void fastin_close(void);                  /* bwa_seq_close */
size_t fastin_read_pairs(bwa_seqio_t *ks, bam_pair_t *recs, size_t B, long *seqs, int broken_input, int drop_aligned); /* a batch of read_bam_pair's */
double fastin_progress(void);             /* fraction of the input file consumed so far, 0 if unknown */
double fastin_inflate_seconds(void);      /* CPU seconds the inflate threads spent (all threads summed) */

void memtemp_begin(void);
int memtemp_put(const void *data, uint32_t len); /* 1 = kept in memory, 0 = caller writes it to the temporary file */
size_t memtemp_records(void);
size_t memtemp_left(void);               /* records not yet taken by pass 2 */
size_t memtemp_bytes(void);
int memtemp_spilled(void);
/* next record kept in memory (pass 2), 0 when all were handed out */
int memtemp_next(const uint8_t **data, uint32_t *len);
size_t memtemp_take(size_t want, size_t *first); /* reserves up to `want` records: returns how many, *first = index of the first */
void memtemp_get(size_t idx, const uint8_t **data, uint32_t *len);
void memtemp_free(void);                 /* the chunks go to a pool for the next run ... */
void memtemp_release(void);              /* ... which this frees */

void write_records_bam(BGZF *output, bam_pair_t *recs, size_t n);
int shim_device_ready(void);          /* bwa_gpu_batch.c: the device context of this run exists */
void shim_count_bgzf(int64_t bytes, double seconds); /* bwa_gpu_batch.c: the run report */

#endif
