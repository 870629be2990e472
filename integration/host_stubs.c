/* host_stubs.c -- libbwahost.so leaves out the reference's main.c (a shared library should not define main); this supplies
 * the one other symbol main.c owns (main.c:43-46) so that the library has no undefined references. */
#include <stdio.h>
void bwa_print_sam_PG(void) { printf("@PG\tID:bwa\tPN:bwa\tVN:bwa_host\n"); }
