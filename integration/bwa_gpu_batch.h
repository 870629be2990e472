/* bwa_gpu_batch.h -- what libbwa_gpu_batch.so offers a host that runs `bam2bam` IN-PROCESS (dlopen the library, call the
 * reference's own entry point bwa_bam_to_bam(argc, argv, version) through it) rather than as a command with LD_PRELOAD.
 * Nothing here is needed for the command-line use. */
#ifndef BWA_GPU_BATCH_H
#define BWA_GPU_BATCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Clocks and counters of the bwa_bam_to_bam call in progress / the last one. */
typedef struct {
	double wall_s;        /* the whole bwa_bam_to_bam call */
	double index_load_s;  /* inside it: bwt_restore_bwt/_sa/_pac (what the reference prints as "loading index...") */
	double device_init_s; /* device context + index upload + re-layout (overlaps the reading of the first batches) */
	double pass1_s, pass2_s;
	double dev_aln_s, dev_sa_s, dev_sw_s, dev_ga_s; /* host-side seconds inside the four device calls */
	double inflate_cpu_s; /* CPU seconds of the input inflate threads, summed */
	double process_cpu_s; /* CPU seconds of the whole process during the call (user + system, every thread) */
	int64_t calls_aln, reads_aln, calls_sa, q_sa, calls_sw, jobs_sw, calls_ga, jobs_ga;
	int64_t sequences;    /* reads that went through pass 1 */
	double dev_bgzf_s;    /* host-side seconds inside bwa_gpu_bgzf_deflate */
	int64_t calls_bgzf, bytes_bgzf;
} bwa_gpu_batch_report_t;

int bwa_gpu_batch_last_report(bwa_gpu_batch_report_t *out);
int bwa_gpu_batch_report_size(void); /* sizeof(bwa_gpu_batch_report_t) as this library was built: lets a foreign-language binding check its mirror */

/* on: the index files a run loads (and their device copy) stay loaded for the next run on the same files; the reference's
 * destroy calls at the end of a run leave them alone.  bwa_gpu_batch_drop_index() frees them. */
void bwa_gpu_batch_keep_index(int on);
void bwa_gpu_batch_drop_index(void);
/* tear the device context down; the next run sets it up again (and re-reads the library's BWAGPU_* settings) */
void bwa_gpu_batch_reset_device(void);

#ifdef __cplusplus
}
#endif
#endif
