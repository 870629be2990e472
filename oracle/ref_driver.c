/* TEST INFRASTRUCTURE ONLY: `bwa` sub-commands of the unmodified reference reached through
 * libbwaref.so instead of the static binary, so that its cross-file calls go through the PLT and the
 * boundary functions can be interposed (integration/bwa_gpu_interpose.c, LD_PRELOAD).
 *   ref_driver bam2bam <options as for `bwa bam2bam`>
 *   ref_driver worker  <options as for `bwa worker`>  (main.c:60 dispatches it the same way) */
#include <stdio.h>
#include <string.h>
int bwa_bam_to_bam(int argc, char *argv[], char *vn);
int bwa_index(int argc, char *argv[]);
int bwa_worker(int argc, char *argv[]);
int main(int argc, char *argv[])
{
	if (argc >= 2 && strcmp(argv[1], "bam2bam") == 0) return bwa_bam_to_bam(argc - 1, argv + 1, "oracle-ref");
	if (argc >= 2 && strcmp(argv[1], "index") == 0) return bwa_index(argc - 1, argv + 1);
	if (argc >= 2 && strcmp(argv[1], "worker") == 0) return bwa_worker(argc - 1, argv + 1);
	fprintf(stderr, "usage: ref_driver bam2bam|index|worker ...\n");
	return 1;
}
