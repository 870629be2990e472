/* bwagpu_cpu_stub.c -- TEST INFRASTRUCTURE ONLY.  Never built into, linked with or loaded by the product.
 *
 * Lets the `-m "not gpu"` suite exercise the HOST LOGIC of integration/bwa_gpu_batch.c (batching, record /
 * replay, ordering of drand48 and of the pass-2 position cache) on a box without a GPU: it defines the
 * handful of bwa_gpu_* entry points that shim calls and answers them with the reference's own per-record
 * functions from oracle/_ref/libbwaref.so.  Pre-loaded AHEAD of the shim so that its symbols win:
 *
 *   LD_PRELOAD=tests/cpu_stub/libbwagpu_cpu_stub.so:integration/libbwa_gpu_batch.so ref_driver bam2bam ...
 *
 * The `-m gpu` twin of the test runs the same command without this stub, i.e. on libbwagpu.so.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bwtaln.h"
#include "bwt.h"
#include "stdaln.h"
#include "bwase.h"
#include "bwa_gpu.h"

static bwt_t *s_bwt[2];
static const ubyte_t *s_pac;
static int64_t s_lpac;
static bwa_cigar_t *s_pool;
static size_t s_pool_n, s_pool_m;

int bwa_gpu_init(int n, const int *ids) { (void)n; (void)ids; fprintf(stderr, "[cpu_stub] bwa_gpu_* answered by the reference's CPU functions\n"); return 0; }
void bwa_gpu_destroy(void) {}
const char *bwa_gpu_last_error(void) { return "cpu stub"; }
int bwa_gpu_load_index(bwt_t *const bwt[2], const ubyte_t *pac, int64_t l_pac)
{
	s_bwt[0] = bwt[0]; s_bwt[1] = bwt[1]; s_pac = pac; s_lpac = l_pac;
	return 0;
}
int bwa_gpu_cal_sa_reads_gap(int n, bwa_seq_t *seqs, const gap_opt_t *opt)
{
	int i;
	for (i = 0; i < n; ++i) bwa_cal_sa_reg_gap(s_bwt, 1, &seqs[i], opt);
	return 0;
}
int bwa_gpu_cal_pac_pos(int64_t n, const bwtint_t *k, const uint8_t *which, bwtint_t *out)
{
	int64_t i;
	for (i = 0; i < n; ++i) out[i] = bwt_sa(s_bwt[which[i] ? 0 : 1], k[i]);
	return 0;
}
int bwa_gpu_mate_sw_path(int n, const bwa_gpu_sw_job_t *jobs, bwa_gpu_path_res_t *res, const bwa_cigar_t **cigar_pool)
{
	int i;
	s_pool_n = 0;
	for (i = 0; i < n; ++i) {
		const bwa_gpu_sw_job_t *jb = &jobs[i];
		ubyte_t *ref = (ubyte_t *)calloc(jb->reglen + 1, 1);
		path_t *path = (path_t *)calloc(jb->reglen + jb->len + 2, sizeof(path_t));
		int64_t k;
		int l = 0, path_len = 0, n_cigar = 0, c;
		bwa_cigar_t *cg;
		for (k = jb->beg; l < jb->reglen && k < s_lpac; ++k) ref[l++] = s_pac[k >> 2] >> ((~k & 3) << 1) & 3;
		memset(&res[i], 0, sizeof(res[i]));
		res[i].score = aln_local_core(ref, l, (ubyte_t *)jb->seq, jb->len, &aln_param_bwa, path, &path_len, 1, 0);
		res[i].cigar_off = (int64_t)s_pool_n;
		if (res[i].score >= 0 && path_len > 0) {
			cg = bwa_aln_path2cigar(path, path_len, &n_cigar);
			if (s_pool_n + n_cigar > s_pool_m) { s_pool_m = (s_pool_n + n_cigar) * 2 + 1024; s_pool = (bwa_cigar_t *)realloc(s_pool, s_pool_m * sizeof(bwa_cigar_t)); }
			for (c = 0; c < n_cigar; ++c) s_pool[s_pool_n++] = cg[c];
			free(cg);
			res[i].n_cigar = n_cigar;
			res[i].start_i = path[path_len - 1].i; res[i].start_j = path[path_len - 1].j;
			res[i].end_i = path[0].i; res[i].end_j = path[0].j;
		}
		free(ref); free(path);
	}
	*cigar_pool = s_pool;
	return 0;
}

static bwa_cigar_t *g_pool;
static size_t g_pool_n, g_pool_m;
int bwa_gpu_global_align_seqs(int n, const bwa_gpu_ga_job_t *jobs, int gap_end, int band, bwa_gpu_path_res_t *res, const bwa_cigar_t **cigar_pool)
{
	int i;
	AlnParam ap = aln_param_bwa;
	ap.gap_end = gap_end; ap.band_width = band;
	g_pool_n = 0;
	for (i = 0; i < n; ++i) {
		const bwa_gpu_ga_job_t *jb = &jobs[i];
		path_t *path = (path_t *)calloc(jb->reflen + jb->len + 2, sizeof(path_t));
		int path_len = 0, n_cigar = 0, c;
		bwa_cigar_t *cg;
		memset(&res[i], 0, sizeof(res[i]));
		res[i].score = aln_global_core((ubyte_t *)jb->ref, jb->reflen, (ubyte_t *)jb->seq, jb->len, &ap, path, &path_len);
		res[i].cigar_off = (int64_t)g_pool_n;
		cg = bwa_aln_path2cigar(path, path_len, &n_cigar);
		if (g_pool_n + n_cigar > g_pool_m) { g_pool_m = (g_pool_n + n_cigar) * 2 + 1024; g_pool = (bwa_cigar_t *)realloc(g_pool, g_pool_m * sizeof(bwa_cigar_t)); }
		for (c = 0; c < n_cigar; ++c) g_pool[g_pool_n++] = cg[c];
		free(cg);
		res[i].n_cigar = n_cigar;
		if (path_len > 0) {
			res[i].start_i = path[path_len - 1].i; res[i].start_j = path[path_len - 1].j;
			res[i].end_i = path[0].i; res[i].end_j = path[0].j;
		}
		free(path);
	}
	*cigar_pool = g_pool;
	return 0;
}

/* the BGZF codec: zlib, one member per 65280 bytes (what bgzf.c's deflate_block does per block) */
#include <zlib.h>
void *bwa_gpu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void bwa_gpu_host_free(void *p) { free(p); }
int bwa_gpu_bgzf_deflate(const uint8_t *in, int64_t n_bytes, int level, const uint8_t **out, int64_t *out_bytes,
                         const int32_t **member_len, int32_t *n_members, double *kernel_ms)
{
	static uint8_t *buf; static int32_t *lens; static size_t m_blk;
	static const uint8_t hdr[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 0, 0};
	const size_t nblk = ((size_t)n_bytes + 65279) / 65280;
	size_t k, at = 0;
	if (nblk > m_blk) { m_blk = nblk + 16; buf = (uint8_t *)realloc(buf, m_blk * 65536); lens = (int32_t *)realloc(lens, m_blk * sizeof(int32_t)); }
	for (k = 0; k < nblk; ++k) {
		const uint8_t *src = in + k * 65280;
		const uint32_t len = (uint32_t)((size_t)n_bytes - k * 65280 < 65280 ? (size_t)n_bytes - k * 65280 : 65280);
		uint8_t *dst = buf + at;
		z_stream zs;
		uint32_t crc, total;
		memcpy(dst, hdr, 18);
		memset(&zs, 0, sizeof(zs));
		zs.next_in = (Bytef *)src; zs.avail_in = len; zs.next_out = dst + 18; zs.avail_out = 65536 - 26;
		if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK || deflate(&zs, Z_FINISH) != Z_STREAM_END) return 1;
		deflateEnd(&zs);
		total = (uint32_t)zs.total_out + 26;
		dst[16] = (uint8_t)((total - 1) & 0xff); dst[17] = (uint8_t)((total - 1) >> 8);
		crc = (uint32_t)crc32(crc32(0L, 0, 0), src, len);
		memcpy(dst + 18 + zs.total_out, &crc, 4); memcpy(dst + 22 + zs.total_out, &len, 4);
		lens[k] = (int32_t)total; at += total;
	}
	*out = buf; *out_bytes = (int64_t)at;
	if (member_len) *member_len = lens;
	if (n_members) *n_members = (int32_t)nblk;
	if (kernel_ms) *kernel_ms = 0;
	return 0;
}

int bwa_gpu_bgzf_inflate(const uint8_t *in, int64_t n_bytes, int32_t n_members, const int64_t *member_off, uint8_t *out,
                         int64_t out_cap, int64_t *out_off, double *kernel_ms)
{
	int64_t total = 0;
	int k;
	(void)n_bytes;
	for (k = 0; k < n_members; ++k) {
		const uint8_t *m = in + member_off[k];
		const size_t len = (size_t)(member_off[k + 1] - member_off[k]);
		uint32_t isize;
		z_stream zs;
		memcpy(&isize, m + len - 4, 4);
		out_off[k] = total;
		if (total + isize > out_cap) return 1;
		memset(&zs, 0, sizeof(zs));
		if (inflateInit2(&zs, -15) != Z_OK) return 1;
		zs.next_in = (Bytef *)(m + 18); zs.avail_in = (uInt)(len - 26);
		zs.next_out = out + total; zs.avail_out = isize;
		if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != isize) { inflateEnd(&zs); return 1; }
		inflateEnd(&zs);
		total += isize;
	}
	out_off[n_members] = total;
	if (kernel_ms) *kernel_ms = 0;
	return 0;
}
