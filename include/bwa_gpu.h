/* bwa_gpu.h -- C-ABI of the B200-native `bwa aln` hot path (libbwagpu.so).
 *
 * The reference (mpieva/network-aware-bwa) has no plugin/FFI interface: the seam is the
 * set of plain C calls bam2bam.c's phase functions make into the alignment layer, one
 * record at a time.  Each entry point below is the BATCH form of one of those calls and
 * cites the call it replaces.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Conventions:
 *   - return 0 on success, non-zero on error (bwa_gpu_last_error() has the text); the
 *     reference itself never returns errors on this path (xassert -> abort, utils.c:68-83),
 *     so a reference-style caller should abort() on non-zero;
 *   - there is NO CPU fallback: every entry point fails if no sm_100 device is usable;
 *   - arrays handed back inside reference structs (bwa_seq_t.aln) are libc calloc()'d,
 *     because bwa_free_read_seq1 (bwaseqio.c:253-261) free()s them.
 *
 * Including this header AFTER the reference's bwtaln.h re-uses the reference's own
 * types; stand-alone it declares layout-identical mirrors (sizes checked at compile
 * time below and against the real headers in tests/test_abi.py).
 */
#ifndef BWA_GPU_H
#define BWA_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(BWTALN_H)
/* ---- mirrors of the reference's types (only when its headers are not in scope) ---- */
#ifndef BWA_UBYTE
#define BWA_UBYTE
typedef unsigned char ubyte_t;
#endif
#ifndef BWA_BWT_H
typedef uint32_t bwtint_t; /* bwt.h:41 */

/* bwt.h:43-59 (non-mmap build, the Makefile default) */
typedef struct {
	bwtint_t primary;
	bwtint_t L2[5];
	bwtint_t seq_len;
	bwtint_t bwt_size;
	uint32_t *bwt;
	uint32_t cnt_table[256];
	int sa_intv;
	bwtint_t n_sa;
	bwtint_t *sa;
} bwt_t;
#endif

/* bwtaln.h:43-47 -- also the .sai record and a 0MQ wire format (bam2bam.c:1001) */
typedef struct {
	uint32_t n_mm : 8, n_gapo : 8, n_gape : 8, a : 1;
	bwtint_t k, l;
	int score;
} bwt_aln1_t;

typedef uint16_t bwa_cigar_t; /* bwtaln.h:49 */

/* bwtaln.h:58-62 */
typedef struct {
	uint32_t pos;
	uint32_t n_cigar : 15, gap : 8, mm : 8, strand : 1;
	bwa_cigar_t *cigar;
} bwt_multi1_t;

/* bwtaln.h:64-90 */
typedef struct {
	char *name;
	ubyte_t *seq, *rseq, *qual;
	uint32_t len : 20, strand : 1, type : 2, dummy : 1, extra_flag : 8;
	uint32_t n_mm : 8, n_gapo : 8, n_gape : 8, mapQ : 8;
	int score;
	int clip_len;
	int n_aln;
	bwt_aln1_t *aln;
	int n_multi;
	bwt_multi1_t *multi;
	bwtint_t sa, pos;
	uint64_t c1 : 28, c2 : 28, seQ : 8;
	int n_cigar;
	bwa_cigar_t *cigar;
	int tid;
	char bc[64];
	uint32_t full_len : 20, nm : 12;
	char *md;
	int max_entries;
} bwa_seq_t;

/* bwtaln.h:143-153; defaults gap_init_opt bwtaln.c:19-35 */
typedef struct {
	int s_mm, s_gapo, s_gape;
	int mode;
	int indel_end_skip, max_del_occ, max_entries;
	float fnr;
	int max_diff, max_gapo, max_gape;
	int max_seed_diff, seed_len;
	int n_threads;
	int max_top2;
	int trim_qual;
} gap_opt_t;

#define BWA_MODE_GAPE 0x01
#define BWA_MODE_COMPREAD 0x02
#define BWA_MODE_LOGGAP 0x04
#define BWA_MODE_NONSTOP 0x10
#endif /* !BWTALN_H */

#if defined(__cplusplus) && __cplusplus >= 201103L
static_assert(sizeof(bwt_aln1_t) == 16, "bwt_aln1_t ABI");
static_assert(sizeof(bwt_multi1_t) == 16, "bwt_multi1_t ABI");
static_assert(sizeof(bwa_seq_t) == 200, "bwa_seq_t ABI");
static_assert(sizeof(gap_opt_t) == 64, "gap_opt_t ABI");
#endif

/* ------------------------------------------------------------------ lifetime */

/* One context (stream set, scratch arenas) per listed CUDA device.  n_devices <= 0 or
 * device_ids == NULL means "device 0 only".  Replaces nothing in the reference (it has
 * no device); sits where init_genome_index (bam2bam.c:844-868) is called. */
int bwa_gpu_init(int n_devices, const int *device_ids);

/* Upload both strands' FM-index and re-lay it out for HBM (32-byte blocks of 4 counts +
 * 64 bases as two bit planes), replicated on every device of the context.
 * bwt[0] = forward (.bwt/.sa), bwt[1] = reverse (.rbwt/.rsa), exactly the globals of
 * bam2bam.c:88-91 after bwt_restore_bwt/bwt_restore_sa (bam2bam.c:848-861).  sa may be
 * NULL (then bwa_gpu_cal_pac_pos fails); pac may be NULL (then bwa_gpu_mate_sw fails).
 * The host arrays are not referenced after the call returns. */
int bwa_gpu_load_index(bwt_t *const bwt[2], const ubyte_t *pac, int64_t l_pac);

/* (Re)load only the packed forward sequence (bwt_restore_pac, bwtio.c:145-152; 4 bases per byte,
 * l_pac/4+1 bytes), for hosts that load it later than the BWTs, as bam2bam's memory plan does
 * (bam2bam.txt:74-78).  Needed by bwa_gpu_mate_sw only. */
int bwa_gpu_load_pac(const ubyte_t *pac, int64_t l_pac);

void bwa_gpu_destroy(void);

const char *bwa_gpu_last_error(void);

/* ------------------------------------------------------------------ index construction
 * The compute of `bwa index` (bwtindex.c:102-190) on one device: from the packed forward sequence (the .pac
 * bns_fasta2bntseq wrote, bntseq.c:169-250: 4 bases per byte, l_pac/4+1 bytes) to both strands' bwt_t --
 * BWT (bwt_pac2bwt, bwtmisc.c:56-101; is.c / bwt_gen) with the occurrence counts interleaved every 128 bases
 * (bwt_bwtupdate_core, bwtmisc.c:125-152) and the suffix array sampled every 32 rows (bwt_cal_sa, bwt.c:48-70); rev is
 * the index of the reversed, not complemented, sequence (bwa_pac_rev_core, bwtmisc.c:168-193).  The structs are filled
 * the way bwt_restore_bwt + bwt_restore_sa fill them (bwtio.c:157-200; bwt and sa malloc()'d, cnt_table set), so they can go
 * straight into bwa_gpu_load_index or to bwt_dump_bwt / bwt_dump_sa.  A full suffix sort in HBM (prefix doubling on radix
 * sorts) replaces induced sorting / BWT-SW; the result is the same bytes.  Needs no bwa_gpu_init.  l_pac < 2^32 - 64. */
int bwa_gpu_index_build(const ubyte_t *pac, int64_t l_pac, int device, bwt_t *fwd, bwt_t *rev);
/* free()s what bwa_gpu_index_build allocated inside *b (bwt_destroy, bwt.c:226-235, without freeing b itself) */
void bwa_gpu_index_free(bwt_t *b);
/* bwt_dump_bwt + bwt_dump_sa (bwtio.c:17-38) for both strands: <prefix>.bwt / .rbwt / .sa / .rsa */
int bwa_gpu_index_write(const char *prefix, const bwt_t *fwd, const bwt_t *rev);

/* ------------------------------------------------------------------ K2 + K3: gapped search
 * Batch form of   bwa_cal_sa_reg_gap(bwt, 1, &seqs[i], opt)   for i in [0, n_seqs)
 * (bwtaln.c:93-142 as called from bam2bam.c:616 and 676: PER-READ semantics -- max_diff,
 * the max_gapo clamp and the seed switch are derived from each read's own length).
 * Reads  seqs[i].{seq, rseq, len};  writes {n_aln, aln (calloc'd), max_entries} and resets
 * sa/type/c1/c2 like bwtaln.c:113.  The aln list is byte-identical to the reference's,
 * order included. */
int bwa_gpu_cal_sa_reads_gap(int n_seqs, bwa_seq_t *seqs, const gap_opt_t *opt);

/* Convenience for batch drivers: free() every seqs[i].aln and zero the field -- the part of
 * bwa_free_read_seq1 (bwaseqio.c:253-261) that concerns what the call above allocated. */
void bwa_gpu_free_alns(int n_seqs, bwa_seq_t *seqs);

/* Flat form of the same call for batch drivers that do not keep bwa_seq_t around.
 * bases[offs[i] .. offs[i+1]) is read i in sequencing orientation, codes 0..3 = ACGT,
 * 4 = N (bwaseqio.c:10); the library derives seq (reversed) and rseq (reverse complement)
 * as bam1_to_seq does (bwaseqio.c:294-297).  Outputs (caller-allocated): n_aln[n],
 * max_entries[n] (0 when the search never started: len == 0 or too many N), aln_off[n+1]
 * (prefix sums of n_aln).  *aln_pool receives a library-owned pinned host array of
 * aln_off[n] records in read order, valid until the next call on this context. */
int bwa_gpu_aln_flat(int n, const uint8_t *bases, const int64_t *offs, const gap_opt_t *opt,
                     int32_t *n_aln, int32_t *max_entries, int64_t *aln_off,
                     const bwt_aln1_t **aln_pool);

/* ------------------------------------------------------------------ K4: SA -> coordinate
 * Batch form of the raw bwt_sa() calls (bwt.c:72-81) at bwase.c:145,152 and
 * bam2bam.c:635-636, 752, 761, 786.  which[i] != 0 selects the forward index (bwt[0]),
 * 0 the reverse index (bwt[1]).  out[i] = bwt_sa(bwt[which ? 0 : 1], sa_idx[i]), the raw
 * u32 value; the caller applies  pos = strand ? out : seq_len - (out + len)  (bwase.c:144-153). */
int bwa_gpu_cal_pac_pos(int64_t n, const bwtint_t *sa_idx, const uint8_t *which, bwtint_t *out_sa);

/* ------------------------------------------------------------------ K5: mate-rescue SW
 * Batch form of the aln_local_core(ref, l, seq, len, &aln_param_bwa, path, &path_len, 1, 0)
 * call inside bwa_sw_core (bwape.c:456; stdaln.c:529-761): local affine-gap alignment of
 * seq[0..len) against the pac window [beg, beg+reglen).  score/end_* are the forward
 * pass's first arg-max in (read row, ref column) scan order (stdaln.c:623-625); start_*
 * come from the reverse pass (stdaln.c:651-696).  Coordinates are 1-based like path_t. */
typedef struct {
	int64_t beg;        /* first reference base of the window (pac coordinate) */
	int32_t reglen;     /* window length (l in bwape.c:447-456) */
	int32_t len;        /* read length */
	const ubyte_t *seq; /* read bases 0..4, orientation as bwa_sw_core receives it */
} bwa_gpu_sw_job_t;

typedef struct {
	int32_t score;
	int32_t start_i, start_j; /* ref, read: 1-based start of the local alignment */
	int32_t end_i, end_j;     /* ref, read: 1-based end */
} bwa_gpu_sw_res_t;

int bwa_gpu_mate_sw(int n, const bwa_gpu_sw_job_t *jobs, bwa_gpu_sw_res_t *res);

/* ------------------------------------------------------------------ K6: banded global alignment + CIGAR
 * One result per job: the score, the path's first and last cell (1-based, window / read coordinates;
 * p = path + path_len - 1 and path[0] in the reference's terms) and the raw CIGAR of the path as
 * bwa_aln_path2cigar gives it (bwtaln.c:396-406: op << 14 | len, ops FROM_M 0 / FROM_I 1 / FROM_D 2),
 * at cigar_pool[cigar_off .. cigar_off + n_cigar).  *cigar_pool is library-owned, valid until the next
 * call of the same family (bwa_gpu_mate_sw_path has one pool, the bwa_gpu_global_align* calls another). */
typedef struct {
	int32_t score;
	int32_t n_cigar;
	int32_t start_i, start_j, end_i, end_j;
	int64_t cigar_off;
} bwa_gpu_path_res_t;

/* Batch form of the whole aln_local_core(ref, l, seq, len, &aln_param_bwa, path, &path_len, 1, 0) call of
 * bwa_sw_core (bwape.c:456): K5 (passes 1-2) followed by the third pass, aln_global_core on the box with
 * band 50 doubled until the global score equals the local one (stdaln.c:723-745).  score is what
 * aln_local_core returns (-1 for its "Potential bug" branch, or for empty input); n_cigar = 0 when the
 * local score is below 1 (bwa_sw_core's thres).  The host keeps bwa_sw_core's post-processing: the
 * >= 20 matched bases test, soft clips, cnt, and the accept/reject arithmetic of bwa_paired_sw1. */
int bwa_gpu_mate_sw_path(int n, const bwa_gpu_sw_job_t *jobs, bwa_gpu_path_res_t *res, const bwa_cigar_t **cigar_pool);

/* Batch form of the aln_global_core(ref_seq, l, seq, len, &ap, path, &path_len) call of refine_gapped_core
 * (bwase.c:212; ap = aln_param_bwa: gap_end 5, band 50) -- and of any other use with aln_param_bwa's
 * scores: jobs[i].reglen reference bases from jobs[i].beg against the read.  gap_end < 0 = end gaps cost
 * like inner gaps.  The host keeps refine_gapped_core's coordinate fix-ups (bwase.c:215-234). */
int bwa_gpu_global_align(int n, const bwa_gpu_sw_job_t *jobs, int gap_end, int band, bwa_gpu_path_res_t *res,
                         const bwa_cigar_t **cigar_pool);

/* The same call on explicit sequences -- aln_global_core(seq1, len1, seq2, len2, &aln_param_bwa-like ap, path, &path_len)
 * (stdaln.c:345) exactly as refine_gapped_core issues it after unpacking its window (bwase.c:199-212): ref[0 .. reflen) and
 * seq[0 .. len), one base per byte (0..3, anything above = N).  For callers that interpose aln_global_core itself and never see
 * the window's pac coordinate (integration/bwa_gpu_batch.c). */
typedef struct {
	const ubyte_t *ref;
	int32_t reflen;
	int32_t len;
	const ubyte_t *seq;
} bwa_gpu_ga_job_t;
int bwa_gpu_global_align_seqs(int n, const bwa_gpu_ga_job_t *jobs, int gap_end, int band, bwa_gpu_path_res_t *res,
                              const bwa_cigar_t **cigar_pool);

/* ------------------------------------------------------------------ BGZF output on the device
 * Replaces the compute of bgzf.c:265-330 (deflate_block, one zlib deflate per <= 64 KB of BAM stream; bam2bam.c:2061
 * opens its output at level 2 = zlib's greedy deflate_fast) and of bgzf_write's blocking (bgzf.c:533-556): `n_bytes` of
 * BAM stream are cut every 65280 bytes and each piece becomes one complete BGZF member (18-byte header with BSIZE, raw
 * deflate stream, CRC-32, ISIZE); *out is the members back to back -- write it to the file as it is.  Inflating the
 * members gives back the input byte for byte (that stream is what the reference writes too); the compressed bytes are
 * this codec's own and the same on every run.  level 0 = stored blocks (bgzf's "u" mode), anything else = compress.
 * *out / *member_len point into pinned memory owned by the library, valid until the next bwa_gpu_bgzf_deflate call;
 * one call at a time (calls are serialised inside).  `in` is copied fastest from bwa_gpu_host_alloc memory. */
int bwa_gpu_bgzf_deflate(const uint8_t *in, int64_t n_bytes, int level, const uint8_t **out, int64_t *out_bytes,
                         const int32_t **member_len, int32_t *n_members, double *kernel_ms);
/* The input side: replaces the compute of bamlite's gzread (bamlite.h:7-11; bam_read1, bamlite.c:125-155 reads the records
 * through one zlib inflate stream).  `n_members` BGZF members lie at in[member_off[k] .. member_off[k+1]); they are inflated on
 * the device (one thread per member, all of them at once) and their bytes placed back to back in `out` (host memory, ideally
 * from bwa_gpu_host_alloc; room for out_cap bytes); out_off[k] .. out_off[k+1] = where member k's bytes are (n_members + 1
 * entries).  A member that is not well-formed deflate, or whose length differs from its ISIZE, fails the call with a message. */
int bwa_gpu_bgzf_inflate(const uint8_t *in, int64_t n_bytes, int32_t n_members, const int64_t *member_off, uint8_t *out,
                         int64_t out_cap, int64_t *out_off, double *kernel_ms);
/* page-locked host memory for the buffers handed to the calls above (NULL + bwa_gpu_last_error on failure) */
void *bwa_gpu_host_alloc(size_t bytes);
void bwa_gpu_host_free(void *p);

/* ------------------------------------------------------------------ measurement hooks
 * Device-side timing (CUDA events on the library's own streams) and work counters of the
 * most recent batch call, summed over its chunks.  occ_fetches counts occurrence-block
 * fetches the way the REFERENCE layout needs them (1 when both ends fall in one 128-base
 * block, else 2: bwt.c:127,189) -- the algorithmic-bytes unit of SURVEY.md §8(d). */
typedef struct {
	double ms_h2d, ms_width, ms_search, ms_compact, ms_d2h, ms_total_device;
	double ms_host_marshal;
	int64_t n_reads, n_aln;
	int64_t n_overflow_t2, n_overflow_t3; /* reads retried after pass 0 / after pass 1 */
	int64_t occ_fetches_width, occ_fetches_search; /* filled only when stats are enabled */
	int64_t own_fetches_width, own_fetches_search; /* 32-byte blocks this layout touched */
	int64_t n_pops, n_pushes;
	int32_t launches; /* kernels launched by the call */
	int32_t n_devices;
	double ms_tier[4]; /* search kernels (+ their width refresh) per pass: [0] k_search on private arenas, [1] k_search_warp on the shared chunk pool, [2] k_search with guaranteed memory */
	int64_t n_stored;  /* records that reached the in-memory stack (stats builds) */
	int64_t n_pruned, n_expand, n_exact, n_derive; /* pops pruned / nodes expanded / exact-tail steps / group-child derivations */
	int64_t n_trips;          /* loop trips summed over threads (stats builds) */
	double ms_sw_kernel;      /* last bwa_gpu_mate_sw call: k_sw device time */
	int64_t x_chunks_used;    /* most 1024-record overflow chunks a k_search launch took from the shared pool */
	int64_t ns_queue_empty;   /* last k_search launch: time from start until the work queue ran dry */
	int64_t ns_kernel;        /* last k_search launch: start to last thread exit (globaltimer) */
} bwa_gpu_stats_t;

int bwa_gpu_get_stats(bwa_gpu_stats_t *out);
/* 0 = off (default, the timed configuration), 1 = count fetches/pops/pushes in-kernel */
int bwa_gpu_set_stats(int enabled);

/* Running totals over every call since bwa_gpu_init / bwa_gpu_reset_totals, for hosts that make many calls from several
 * threads (a bam2bam run through integration/bwa_gpu_batch.c): kernel-only device time (CUDA events on the library's
 * streams; the lanes of a device overlap, so the sum can exceed wall time), launches, units and copy volumes. */
typedef struct {
	double ms_width, ms_search, ms_sa, ms_sw, ms_global; /* K2 (+K2b), K3 (all passes), K4, K5, K6 */
	double ms_search_pass[3];
	int64_t launches;                                  /* kernels launched (cub passes included) */
	int64_t reads, alns, sa_queries, sw_jobs, ga_jobs; /* units: reads searched, hits returned, SA rows, K5 jobs, K6-only jobs */
	int64_t sw_cells_fwd;                              /* sum of window x read cells of K5's forward pass */
	int64_t h2d_bytes, d2h_bytes;
	int64_t occ_fetches_width, occ_fetches_search, own_fetches_search; /* from calls made while stats were enabled */
	double ms_bgzf;                                    /* BGZF deflate kernels (deflate + offsets + pack) */
	int64_t bgzf_bytes_in, bgzf_bytes_out;             /* BAM stream bytes in / member bytes out */
	double ms_inflate;                                 /* BGZF inflate kernel */
	int64_t inflate_bytes_in, inflate_bytes_out;       /* member bytes in / BAM stream bytes out */
} bwa_gpu_totals_t;
int bwa_gpu_get_totals(bwa_gpu_totals_t *out);
void bwa_gpu_reset_totals(void);

/* Roofline denominator for the occurrence-lookup kernels: the rate, in GB/s of 32-byte sectors, that a
 * kernel of nothing but dependent random sector loads over a buffer_bytes buffer sustains on device 0, with `chains`
 * (1, 2, 4 or 8) independent chains per thread and `steps` loads per chain.  Measurement only; replaces no reference call. */
int bwa_gpu_probe_random_sectors(int64_t buffer_bytes, int chains, int steps, double *gb_per_s);

/* Device-resident variant used for kernel-only throughput: stage a flat batch in HBM
 * once, then run K2+K3 over it repeatedly without host traffic.  run returns the device
 * time of that pass in *ms (CUDA events).  Results stay on the device; fetch copies the
 * last pass's results out in bwa_gpu_aln_flat's output format. */
int bwa_gpu_resident_stage(int n, const uint8_t *bases, const int64_t *offs, const gap_opt_t *opt);
int bwa_gpu_resident_run(double *ms);
int bwa_gpu_resident_fetch(int32_t *n_aln, int32_t *max_entries, int64_t *aln_off,
                           const bwt_aln1_t **aln_pool);

#ifdef __cplusplus
}
#endif
#endif /* BWA_GPU_H */
