/* inproc_host.c -- TEST INFRASTRUCTURE ONLY: a host process that runs `bam2bam` several times IN-PROCESS through
 * libbwa_gpu_batch.so (the way bench.py does with ctypes on the GPU box), linked so that tests/cpu_stub answers the
 * bwa_gpu_* calls.  usage: inproc_host <runs> <keep_index 0|1> <prefix> <in.bam> <out-stem>   -> <out-stem>.<run>.bam */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bwa_gpu_batch.h"
int bwa_bam_to_bam(int argc, char *argv[], char *vn);
int main(int argc, char *argv[])
{
	int runs, keep, r;
	if (argc != 6) { fprintf(stderr, "usage: inproc_host runs keep prefix in.bam out-stem\n"); return 2; }
	runs = atoi(argv[1]); keep = atoi(argv[2]);
	bwa_gpu_batch_keep_index(keep);
	for (r = 0; r < runs; ++r) {
		char out[4096];
		char *av[] = {"bam2bam", "-g", argv[3], "-t", "1", "-f", out, argv[4], 0};
		bwa_gpu_batch_report_t rep;
		snprintf(out, sizeof out, "%s.%d.bam", argv[5], r);
		if (bwa_bam_to_bam(8, av, "inproc")) return 1;
		if (bwa_gpu_batch_last_report(&rep)) return 1;
		printf("run %d wall %.3f index_load %.3f pass1 %.3f pass2 %.3f cpu %.3f reads_aln %ld sequences %ld\n", r, rep.wall_s, rep.index_load_s,
		       rep.pass1_s, rep.pass2_s, rep.process_cpu_s, (long)rep.reads_aln, (long)rep.sequences);
	}
	bwa_gpu_batch_drop_index();
	return 0;
}
