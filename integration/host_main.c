/* host_main.c -- the `bwa` sub-commands of the alignment workflow, dispatched the way the reference's main.c:48-76 does,
 * on top of the shared-library build of the unmodified reference (integration/Makefile: _host/libbwahost.so).
 *   bwa_host bam2bam <options as for `bwa bam2bam`>      (LD_PRELOAD=libbwa_gpu_batch.so puts the hot path on the GPU)
 *   bwa_host worker  <options as for `bwa worker`>
 *   bwa_host index   <options as for `bwa index`> */
#include <stdio.h>
#include <string.h>
int bwa_bam_to_bam(int argc, char *argv[], char *vn);
int bwa_index(int argc, char *argv[]);
int bwa_worker(int argc, char *argv[]);
int main(int argc, char *argv[])
{
	if (argc >= 2 && strcmp(argv[1], "bam2bam") == 0) return bwa_bam_to_bam(argc - 1, argv + 1, "bwa_host");
	if (argc >= 2 && strcmp(argv[1], "index") == 0) return bwa_index(argc - 1, argv + 1);
	if (argc >= 2 && strcmp(argv[1], "worker") == 0) return bwa_worker(argc - 1, argv + 1);
	fprintf(stderr, "usage: bwa_host bam2bam|index|worker ...\n");
	return 1;
}
