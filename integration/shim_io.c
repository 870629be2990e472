/* shim_io.c -- record I/O of the batched drop-in: everything around the hot path that moves
 * bytes in and out of a bam2bam run once the alignment itself is on the device.
 *
 *   input BAM     bam_read1 (bamlite.c:125-155) reads through zlib's gzread: one inflate stream on the thread that also
 *                 parses and allocates the records.  Here bam_read1 is REPLACED (it is a plain global of the reference,
 *                 reached through the PLT): the input file's BGZF blocks -- the framing every BAM writer produces -- are
 *                 inflated into a ring, in block order, by the device (bwa_gpu_bgzf_inflate, 1024 members per call, a warp
 *                 per member) or, until the device context exists, by a pool of zlib threads; records are parsed out of the
 *                 ring a batch at a time.
 *                 A plain (single-stream) gzip input cannot be cut into blocks: it is announced on stderr and read ahead
 *                 by ONE zlib thread instead.
 *   intermediate  pass 1 hands its records to pass 2 through a gzip'ed temporary file (pair_print_custom /
 *                 read_pair_custom, bam2bam.c:1099-1137).  The shim keeps the same encoded messages in memory up to a cap
 *                 and spills the rest to that file in the reference's format.
 *   output BAM    pair_print_bam -> bgzf_write deflates one 64 KB block at a time on the calling thread (bgzf.c:594-623);
 *                 here a batch's records are laid out as the same byte stream and its blocks deflated in parallel.
 */
#include "shim.h"
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* ================================================================== input BAM */
#define FI_SLOTS 4096           /* ring of inflated blocks: up to 256 MB (touched only as far as the reader runs ahead) */
#define FI_DEV_MEMBERS 1024     /* members per device inflate call: the kernel gives each a warp, so a call wants many */
#define FI_BLOCK 65536
#define FI_MAXTH 16
enum { FS_FREE = 0, FS_BUSY, FS_FILLED };

typedef struct {
	uint8_t *buf;
	int len, state;
	int clen; /* the member's size in the file */
	uint64_t seq;
} fi_slot_t;

typedef struct {
	int active;              /* a stream is set up (either kind) */
	int bgzf;                /* 1: block-parallel, 0: one read-ahead thread on the zlib handle */
	char *path;
	bamFile fp;              /* the reference's handle: its position tells how long the header was */
	/* BGZF */
	const uint8_t *map;
	size_t map_len, scan;
	uint64_t next_job, total; /* total = number of blocks, known once the scan hit the end (else ~0) */
	fi_slot_t slot[FI_SLOTS];
	pthread_mutex_t mu;
	pthread_cond_t cv_filled, cv_free;
	pthread_t th[FI_MAXTH];
	int nth, stop, error;
	int dev_active;          /* the device worker has taken over: the zlib workers retire */
	pthread_t dev_th;
	int dev_started;
	double cpu_s;
	/* consumer */
	uint64_t cons_seq;
	size_t cons_file_bytes; /* file bytes of the members handed to the consumer so far */
	int cons_off, have_slot;
	size_t skip;
	/* the batch reader's staging area: decompressed stream bytes [stage_pos, stage_len) come before what the ring holds */
	uint8_t *stage;
	size_t stage_pos, stage_len, stage_cap;
	/* plain gzip: ring of bytes filled by one thread */
	uint8_t *ring;
	volatile size_t head, tail;
	volatile int eof;
} fastin_t;

static fastin_t F;
static char *g_next_path;
static double g_inflate_s;

#define RA_CAP ((size_t)32 << 20)
#define RA_PIECE ((size_t)256 << 10)

void fastin_set_path(const char *path)
{
	free(g_next_path);
	g_next_path = path ? strdup(path) : 0;
}

double fastin_inflate_seconds(void) { return g_inflate_s; }

/* how far through the input file the reader is (0..1; 0 when that is not known: plain gzip, a pipe) */
double fastin_progress(void) { return F.active && F.bgzf && F.map_len ? (double)F.cons_file_bytes / (double)F.map_len : 0.0; }

static int bgzf_header_ok(const uint8_t *p, size_t left, size_t *bsize)
{
	/* 1f 8b 08 04 | mtime xfl os | xlen = 6 | 'B' 'C' 2 0 | BSIZE (bgzf.c:274-291) */
	if (left < 28 || p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return 0;
	if (p[10] != 6 || p[11] != 0 || p[12] != 66 || p[13] != 67 || p[14] != 2 || p[15] != 0) return 0;
	*bsize = (size_t)(p[16] | p[17] << 8) + 1;
	return *bsize >= 26 && *bsize <= left;
}

static void *fi_worker(void *arg)
{
	z_stream zs;
	double cpu = 0;
	(void)arg;
	memset(&zs, 0, sizeof(zs));
	if (inflateInit2(&zs, -15) != Z_OK) { pthread_mutex_lock(&F.mu); F.error = 1; pthread_cond_broadcast(&F.cv_filled); pthread_mutex_unlock(&F.mu); return 0; }
	for (;;) {
		size_t off, bsize = 0;
		uint64_t j;
		fi_slot_t *s;
		double t0;
		uint32_t isize;
		pthread_mutex_lock(&F.mu);
		if (F.stop || F.error || F.dev_active) { pthread_mutex_unlock(&F.mu); break; }
		if (F.scan >= F.map_len) { /* clean end of file */
			F.total = F.next_job;
			pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
			break;
		}
		if (!bgzf_header_ok(F.map + F.scan, F.map_len - F.scan, &bsize)) {
			fprintf(stderr, "[bwa_gpu_batch] %s: damaged BGZF block at offset %zu\n", F.path, F.scan);
			F.error = 1;
			pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
			break;
		}
		off = F.scan; F.scan += bsize; j = F.next_job++;
		s = &F.slot[j % FI_SLOTS];
		while (s->state != FS_FREE && !F.stop) pthread_cond_wait(&F.cv_free, &F.mu);
		if (F.stop) { pthread_mutex_unlock(&F.mu); break; }
		s->state = FS_BUSY;
		pthread_mutex_unlock(&F.mu);
		t0 = thread_cpu_now();
		memcpy(&isize, F.map + off + bsize - 4, 4);
		if (isize > FI_BLOCK) isize = FI_BLOCK + 1; /* forces the error below */
		zs.next_in = (Bytef *)(F.map + off + 18); zs.avail_in = (uInt)(bsize - 18 - 8);
		zs.next_out = s->buf; zs.avail_out = FI_BLOCK;
		{
			const int rc = isize <= FI_BLOCK ? inflate(&zs, Z_FINISH) : Z_DATA_ERROR;
			const int bad = rc != Z_STREAM_END || zs.total_out != isize;
			inflateReset(&zs);
			cpu += thread_cpu_now() - t0;
			pthread_mutex_lock(&F.mu);
			if (bad) { fprintf(stderr, "[bwa_gpu_batch] %s: inflate failed in the block at offset %zu\n", F.path, off); F.error = 1; }
			s->len = (int)isize; s->clen = (int)bsize; s->seq = j; s->state = FS_FILLED;
			pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
		}
	}
	inflateEnd(&zs);
	cpu_add(CPU_INFLATE, cpu);
	return 0;
}

/* The device takes the inflate over as soon as its context exists (bwa_gpu_bgzf_inflate: one thread per BGZF member,
 * FI_DEV_MEMBERS members per call); until then -- the first second of a fresh process, while the index is uploaded -- and with
 * BWAGPU_HOST_INFLATE=1 the zlib workers above do it.  Both fill the same ring in member order. */
static void *fi_device_worker(void *arg)
{
	uint8_t *dbuf = 0;
	const char *em = getenv("BWAGPU_INFLATE_MEMBERS");
	const int M = em && atoi(em) > 0 ? (atoi(em) > (1 << 16) ? 1 << 16 : atoi(em)) : FI_DEV_MEMBERS; /* members per call */
	int64_t *moff = (int64_t *)malloc(((size_t)M + 1) * sizeof(int64_t)), *ooff = (int64_t *)malloc(((size_t)M + 1) * sizeof(int64_t));
	const char *er = getenv("BWAGPU_BATCH_RAMP");
	const int ramp = !(er && atoi(er) == 0);
	int call_no;
	(void)arg;
	while (!shim_device_ready()) {
		if (F.stop || F.error) { free(moff); free(ooff); return 0; }
		usleep(2000);
	}
	{ /* the page-locked landing area is kept between runs of one process (allocating it costs as much as inflating into it) */
		static uint8_t *kept; static size_t kept_bytes;
		if ((size_t)M * FI_BLOCK > kept_bytes) {
			if (kept) bwa_gpu_host_free(kept);
			kept = (uint8_t *)bwa_gpu_host_alloc((size_t)M * FI_BLOCK);
			kept_bytes = kept ? (size_t)M * FI_BLOCK : 0;
		}
		dbuf = kept;
	}
	if (!dbuf) { free(moff); free(ooff); return 0; } /* the zlib workers carry on */
	pthread_mutex_lock(&F.mu);
	F.dev_active = 1;
	pthread_mutex_unlock(&F.mu);
	for (call_no = 0;; ++call_no) {
		int n = 0, k;
		uint64_t j0;
		/* the first calls are short (128, 256, 512 members), like the first record batches: the reader has something to parse
		 * after a few milliseconds instead of after a whole 64 MB call */
		const int lim = ramp && call_no < 3 && (128 << call_no) < M ? 128 << call_no : M;
		pthread_mutex_lock(&F.mu);
		if (F.stop || F.error) { pthread_mutex_unlock(&F.mu); break; }
		if (F.scan >= F.map_len) { /* clean end of file */
			F.total = F.next_job;
			pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
			break;
		}
		while (n < lim && F.scan < F.map_len) {
			size_t bsize = 0;
			if (!bgzf_header_ok(F.map + F.scan, F.map_len - F.scan, &bsize)) {
				if (n) break; /* hand over what is good first */
				fprintf(stderr, "[bwa_gpu_batch] %s: damaged BGZF block at offset %zu\n", F.path, F.scan);
				F.error = 1;
				pthread_cond_broadcast(&F.cv_filled);
				break;
			}
			moff[n++] = (int64_t)F.scan;
			F.scan += bsize;
		}
		if (F.error) { pthread_mutex_unlock(&F.mu); break; }
		moff[n] = (int64_t)F.scan;
		j0 = F.next_job; F.next_job += (uint64_t)n;
		pthread_mutex_unlock(&F.mu);
		if (bwa_gpu_bgzf_inflate(F.map, (int64_t)F.map_len, n, moff, dbuf, (int64_t)M * FI_BLOCK, ooff, 0)) {
			fprintf(stderr, "[bwa_gpu_batch] %s: %s\n", F.path, bwa_gpu_last_error());
			pthread_mutex_lock(&F.mu);
			F.error = 1;
			pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
			break;
		}
		for (k = 0; k < n; ++k) {
			fi_slot_t *s = &F.slot[(j0 + (uint64_t)k) % FI_SLOTS];
			const double c0 = thread_cpu_now();
			pthread_mutex_lock(&F.mu);
			while (s->state != FS_FREE && !F.stop) pthread_cond_wait(&F.cv_free, &F.mu);
			if (F.stop) { pthread_mutex_unlock(&F.mu); goto out; }
			s->state = FS_BUSY;
			pthread_mutex_unlock(&F.mu);
			memcpy(s->buf, dbuf + ooff[k], (size_t)(ooff[k + 1] - ooff[k]));
			pthread_mutex_lock(&F.mu);
			s->len = (int)(ooff[k + 1] - ooff[k]); s->clen = (int)(moff[k + 1] - moff[k]); s->seq = j0 + (uint64_t)k; s->state = FS_FILLED;
			pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
			cpu_add(CPU_INFLATE, thread_cpu_now() - c0);
		}
	}
out:
	free(moff); free(ooff);
	return 0;
}

static void *ra_main(void *arg) /* plain gzip: the one thread that may touch the zlib handle from now on */
{
	uint8_t *piece = (uint8_t *)malloc(RA_PIECE);
	double cpu = 0;
	(void)arg;
	while (!F.stop) {
		int got, done = 0;
		double t0;
		while (!F.stop && RA_CAP - (F.head - __atomic_load_n(&F.tail, __ATOMIC_ACQUIRE)) < RA_PIECE) usleep(100);
		if (F.stop) break;
		t0 = thread_cpu_now();
		got = gzread(F.fp, piece, (unsigned)RA_PIECE);
		cpu += thread_cpu_now() - t0;
		if (got <= 0) break;
		while (done < got) {
			const size_t at = (F.head + (size_t)done) % RA_CAP;
			const size_t run = RA_CAP - at < (size_t)(got - done) ? RA_CAP - at : (size_t)(got - done);
			memcpy(F.ring + at, piece + done, run);
			done += (int)run;
		}
		__atomic_store_n(&F.head, F.head + (size_t)got, __ATOMIC_RELEASE);
	}
	__atomic_store_n(&F.eof, 1, __ATOMIC_RELEASE);
	free(piece);
	cpu_add(CPU_INFLATE, cpu);
	return 0;
}

static int fi_threads(void)
{
	const char *e = getenv("BWAGPU_INFLATE_THREADS");
	int n = e ? atoi(e) : shim_threads() / 2;
	if (n < 1) n = 1;
	if (n > FI_MAXTH) n = FI_MAXTH;
	return n;
}

/* first bam_read1 on a handle: decide how its bytes will be produced */
static void fastin_start(bamFile fp)
{
	const char *e = getenv("BWAGPU_READAHEAD");
	int fd = -1, i;
	struct stat st;
	size_t bsize;
	memset(&F, 0, sizeof(F));
	F.fp = fp;
	F.total = ~(uint64_t)0;
	if (e && atoi(e) == 0) return; /* leave the stream to the reference's bam_read1 */
	F.path = g_next_path; g_next_path = 0;
	if (F.path && strcmp(F.path, "-") != 0) fd = open(F.path, O_RDONLY);
	if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 28) {
		void *m = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
		if (m != MAP_FAILED) {
			if (bgzf_header_ok((const uint8_t *)m, (size_t)st.st_size, &bsize)) {
				F.map = (const uint8_t *)m; F.map_len = (size_t)st.st_size; F.bgzf = 1;
				madvise(m, F.map_len, MADV_SEQUENTIAL);
			} else munmap(m, (size_t)st.st_size);
		}
	}
	if (fd >= 0) close(fd);
	if (F.bgzf) {
		F.skip = (size_t)gztell(fp); /* the header the reference already read through its own handle */
		pthread_mutex_init(&F.mu, 0); pthread_cond_init(&F.cv_filled, 0); pthread_cond_init(&F.cv_free, 0);
		for (i = 0; i < FI_SLOTS; ++i) { F.slot[i].buf = (uint8_t *)malloc(FI_BLOCK); F.slot[i].state = FS_FREE; }
		F.nth = fi_threads();
		for (i = 0; i < F.nth; ++i) pthread_create(&F.th[i], 0, fi_worker, 0);
		{
			const char *h = getenv("BWAGPU_HOST_INFLATE");
			if (!(h && atoi(h) != 0)) F.dev_started = pthread_create(&F.dev_th, 0, fi_device_worker, 0) == 0;
		}
		F.active = 1;
		return;
	}
	fprintf(stderr, "[bwa_gpu_batch] input is not a BGZF file (plain gzip or a pipe): its inflate cannot be spread over threads; "
	                "one read-ahead thread is used\n");
	F.ring = (uint8_t *)malloc(RA_CAP);
	if (!F.ring) return;
	F.nth = 1;
	pthread_create(&F.th[0], 0, ra_main, 0);
	F.active = 1;
}

void fastin_close(void)
{
	int i;
	if (!F.fp) return;
	if (F.active) {
		if (F.bgzf) {
			pthread_mutex_lock(&F.mu);
			F.stop = 1;
			pthread_cond_broadcast(&F.cv_free); pthread_cond_broadcast(&F.cv_filled);
			pthread_mutex_unlock(&F.mu);
		} else F.stop = 1;
		for (i = 0; i < F.nth; ++i) pthread_join(F.th[i], 0);
		if (F.bgzf && F.dev_started) pthread_join(F.dev_th, 0);
		if (F.bgzf) {
			for (i = 0; i < FI_SLOTS; ++i) free(F.slot[i].buf);
			munmap((void *)F.map, F.map_len);
			pthread_mutex_destroy(&F.mu); pthread_cond_destroy(&F.cv_filled); pthread_cond_destroy(&F.cv_free);
		} else free(F.ring);
	}
	g_inflate_s += F.cpu_s;
	free(F.path);
	free(F.stage);
	memset(&F, 0, sizeof(F));
}

/* the consumer's next block: 1 = F.slot[cons_seq % FI_SLOTS] is readable, 0 = end of file, -1 = error */
static int fi_next_block(void)
{
	int rc;
	pthread_mutex_lock(&F.mu);
	if (F.have_slot) {
		F.cons_file_bytes += (size_t)F.slot[F.cons_seq % FI_SLOTS].clen;
		F.slot[F.cons_seq % FI_SLOTS].state = FS_FREE;
		pthread_cond_broadcast(&F.cv_free);
		++F.cons_seq;
		F.have_slot = 0;
	}
	for (;;) {
		fi_slot_t *s = &F.slot[F.cons_seq % FI_SLOTS];
		if (F.error) { rc = -1; break; }
		if (s->state == FS_FILLED && s->seq == F.cons_seq) { F.have_slot = 1; F.cons_off = 0; rc = 1; break; }
		if (F.total != ~(uint64_t)0 && F.cons_seq >= F.total) { rc = 0; break; }
		pthread_cond_wait(&F.cv_filled, &F.mu);
	}
	pthread_mutex_unlock(&F.mu);
	return rc;
}

/* up to len bytes of the decompressed stream; returns how many were available (fewer only at the end), -1 on error */
static int fastin_read(void *dst, int len)
{
	int done = 0;
	if (F.stage_pos < F.stage_len) { /* left over from the batch reader */
		size_t run = F.stage_len - F.stage_pos;
		if (run > (size_t)len) run = (size_t)len;
		memcpy(dst, F.stage + F.stage_pos, run);
		F.stage_pos += run; done = (int)run;
		if (done == len) return done;
	}
	if (F.bgzf) {
		while (done < len) {
			fi_slot_t *s;
			int run;
			if (!F.have_slot || F.cons_off == F.slot[F.cons_seq % FI_SLOTS].len) {
				const int rc = fi_next_block();
				if (rc < 0) return -1;
				if (rc == 0) break;
			}
			s = &F.slot[F.cons_seq % FI_SLOTS];
			run = s->len - F.cons_off;
			if (F.skip) { /* the BAM header, already consumed through the zlib handle */
				if ((size_t)run > F.skip) run = (int)F.skip;
				F.skip -= (size_t)run; F.cons_off += run;
				continue;
			}
			if (run > len - done) run = len - done;
			memcpy((uint8_t *)dst + done, s->buf + F.cons_off, (size_t)run);
			F.cons_off += run; done += run;
		}
		return done;
	}
	while (done < len) {
		size_t avail = __atomic_load_n(&F.head, __ATOMIC_ACQUIRE) - F.tail;
		if (avail == 0) {
			if (__atomic_load_n(&F.eof, __ATOMIC_ACQUIRE) && __atomic_load_n(&F.head, __ATOMIC_ACQUIRE) == F.tail) break;
			usleep(50);
			continue;
		}
		{
			const size_t at = F.tail % RA_CAP;
			size_t run = avail < (size_t)(len - done) ? avail : (size_t)(len - done);
			if (RA_CAP - at < run) run = RA_CAP - at;
			memcpy((uint8_t *)dst + done, F.ring + at, run);
			done += (int)run;
			__atomic_store_n(&F.tail, F.tail + run, __ATOMIC_RELEASE);
		}
	}
	return done;
}

/* bam_read1 (bamlite.c:125-155) on the shim's stream: same fields, same return values (little-endian hosts) */
int bam_read1(bamFile fp, bam1_t *b)
{
	REAL(int, bam_read1, bamFile, bam1_t *);
	bam1_core_t *c = &b->core;
	int32_t block_len;
	uint32_t x[8];
	int ret;
	if (F.fp != fp) { /* a handle not seen before */
		if (F.fp) fastin_close();
		fastin_start(fp);
	}
	if (!F.active) return real_bam_read1(fp, b);
	if ((ret = fastin_read(&block_len, 4)) != 4) return ret == 0 ? -1 : -2; /* normal end of file / truncated */
	if (fastin_read(x, 32) != 32) return -3;
	c->tid = (int32_t)x[0]; c->pos = (int32_t)x[1];
	c->bin = x[2] >> 16; c->qual = x[2] >> 8 & 0xff; c->l_qname = x[2] & 0xff;
	c->flag = x[3] >> 16; c->n_cigar = x[3] & 0xffff;
	c->l_qseq = (int32_t)x[4];
	c->mtid = (int32_t)x[5]; c->mpos = (int32_t)x[6]; c->isize = (int32_t)x[7];
	b->data_len = block_len - (int32_t)sizeof(bam1_core_t);
	if (b->data_len < 0) return -4;
	if (b->m_data < b->data_len) {
		b->m_data = b->data_len;
		kroundup32(b->m_data);
		b->data = (uint8_t *)realloc(b->data, (size_t)b->m_data);
	}
	if (fastin_read(b->data, b->data_len) != b->data_len) return -4;
	b->l_aux = b->data_len - c->n_cigar * 4 - c->l_qname - c->l_qseq - (c->l_qseq + 1) / 2;
	return 4 + block_len;
}

/* ------------------------------------------------------------------ a batch of records at a time
 * read_bam_pair (bwaseqio.c:343-407, 466-496) on one thread is the slowest stage of pass 1 once the search is on the
 * device: three reads, a malloc and a tag scan per record.  fastin_read_pairs cuts the decompressed stream into records
 * on the calling thread (a length field and a name comparison each) and builds the bam_pair_t's on the host threads.
 * Only the plain cases are taken this way -- a single read, or two mates with equal names and READ1/READ2 flags in either
 * order; anything else (lone mates, odd flags, a truncated file, the --broken-input / --drop-aligned modes) goes to the
 * reference's own read_bam_pair, which continues from the same stream position. */
static int stage_need(size_t upto) /* 1: the staging area holds stream bytes up to `upto` */
{
	while (F.stage_len < upto) {
		fi_slot_t *sl;
		int off, run;
		if (!F.have_slot || F.cons_off == F.slot[F.cons_seq % FI_SLOTS].len) {
			if (fi_next_block() <= 0) return 0;
		}
		sl = &F.slot[F.cons_seq % FI_SLOTS];
		off = F.cons_off; run = sl->len - off; /* what fastin_read has not taken from the block */
		if (F.skip) { /* the BAM header, already consumed through the zlib handle */
			const int h = (size_t)run > F.skip ? (int)F.skip : run;
			F.skip -= (size_t)h; off += h; run -= h;
		}
		if (F.stage_len + (size_t)run > F.stage_cap) {
			F.stage_cap = (F.stage_len + (size_t)run) * 2 + ((size_t)1 << 20);
			F.stage = (uint8_t *)realloc(F.stage, F.stage_cap);
			if (!F.stage) { fprintf(stderr, "[bwa_gpu_batch] out of memory staging the input\n"); exit(1); }
		}
		memcpy(F.stage + F.stage_len, sl->buf + off, (size_t)run);
		F.stage_len += (size_t)run;
		F.cons_off = sl->len; /* the block is used up */
	}
	return 1;
}

typedef struct { size_t off[2]; int kind; } fr_desc_t; /* kind 0: already built by read_bam_pair */
typedef struct { bam_pair_t *recs; const fr_desc_t *d; const uint8_t *stage; } fr_ctx_t;

/* erase_unwanted_tags (bwaseqio.c:411-464): AM NM CM SM MD X0 X1 XA XC XG XM XN XO XT YQ leave the record; like the
 * reference, l_aux is left as it was */
static void fr_strip_tags(bam1_t *b)
{
	uint8_t *end = b->data + b->data_len;
	uint8_t *rd = b->data + b->core.n_cigar * 4 + b->core.l_qname + b->core.l_qseq + (b->core.l_qseq + 1) / 2, *wr = rd;
	int kept = (int)(rd - b->data);
	while (rd < end) {
		const int a = rd[0], c = rd[1], ty = rd[2] & ~32;
		int drop = 0, len = 3;
		if (a == 'A' || a == 'S' || a == 'C' || a == 'N') drop = c == 'M';
		else if (a == 'M') drop = c == 'D';
		else if (a == 'X') drop = c == 0 || strchr("01ACGMNOT", c) != 0; /* strchr finds the terminator too */
		else if (a == 'Y') drop = c == 'Q';
		if (ty == 'C' || ty == 'A') len += 1;
		else if (ty == 'S') len += 2;
		else if (ty == 'I' || ty == 'F') len += 4;
		else if (ty == 'D') len += 8;
		else if (ty == 'Z' || ty == 'H') { while (rd[len]) ++len; ++len; }
		else if (ty == 'B') {
			const int count = (int)rd[4] | (int)rd[5] << 8 | (int)rd[6] << 16 | (int)rd[7] << 24, el = rd[3] & ~32;
			len += 5;
			if (el == 'C' || el == 'A') len += count;
			else if (el == 'S') len += 2 * count;
			else if (el == 'I' || el == 'F') len += 4 * count;
			else if (el == 'D') len += 8 * count;
		}
		if (!drop) { memmove(wr, rd, (size_t)len); wr += len; kept += len; }
		rd += len;
	}
	b->data_len = kept;
}

static void fr_fill(bam1_t *b, const uint8_t *rec) /* bam_read1's fields from the record at rec (its length word first) */
{
	bam1_core_t *c = &b->core;
	int32_t block_len;
	uint32_t x[8];
	memcpy(&block_len, rec, 4); memcpy(x, rec + 4, 32);
	c->tid = (int32_t)x[0]; c->pos = (int32_t)x[1];
	c->bin = x[2] >> 16; c->qual = x[2] >> 8 & 0xff; c->l_qname = x[2] & 0xff;
	c->flag = x[3] >> 16; c->n_cigar = x[3] & 0xffff;
	c->l_qseq = (int32_t)x[4];
	c->mtid = (int32_t)x[5]; c->mpos = (int32_t)x[6]; c->isize = (int32_t)x[7];
	b->data_len = block_len - 32;
	if (b->data_len > 0) {
		b->m_data = b->data_len;
		kroundup32(b->m_data);
		b->data = (uint8_t *)malloc((size_t)b->m_data);
		memcpy(b->data, rec + 36, (size_t)b->data_len);
	}
	b->l_aux = b->data_len - c->n_cigar * 4 - c->l_qname - c->l_qseq - (c->l_qseq + 1) / 2;
}

static void fr_build_one(size_t i, void *ctx)
{
	const fr_ctx_t *c = (const fr_ctx_t *)ctx;
	const fr_desc_t *d = &c->d[i];
	bam_pair_t *p = &c->recs[i];
	int k;
	if (d->kind == 0) return;
	memset(p, 0, sizeof(*p));
	for (k = 0; k < d->kind; ++k) fr_fill(&p->bam_rec[k], c->stage + d->off[k]);
	p->kind = d->kind;
	if (d->kind == 2) { /* either both mates fail QC or none (bwaseqio.c:488-491) */
		p->bam_rec[0].core.flag |= p->bam_rec[1].core.flag & SAM_FQC;
		p->bam_rec[1].core.flag |= p->bam_rec[0].core.flag & SAM_FQC;
	}
	for (k = 0; k < d->kind; ++k) fr_strip_tags(&p->bam_rec[k]);
}

/* a well-formed record starts at p: its length in *len (0 = not usable here) */
static int fr_peek(size_t p, int32_t *len)
{
	int32_t bl;
	const uint8_t *r;
	int l_qname, n_cigar, l_qseq;
	*len = 0;
	if (!stage_need(p + 36)) return 0;
	memcpy(&bl, F.stage + p, 4);
	if (bl < 32 || !stage_need(p + 4 + (size_t)bl)) return 0;
	r = F.stage + p + 4;
	l_qname = r[8]; n_cigar = r[12] | r[13] << 8; memcpy(&l_qseq, r + 16, 4);
	if (l_qname < 1 || l_qseq < 0 || (long long)l_qname + 4ll * n_cigar + l_qseq + (l_qseq + 1) / 2 > (long long)bl - 32) return 0;
	if (r[32 + l_qname - 1] != 0) return 0; /* the name is not terminated */
	*len = bl;
	return 1;
}

/* up to B records of the input into recs[]; returns how many (fewer than B only at the end of the input); the number
 * of reads among them in *seqs.  Errors are fatal, with the reference's message. */
size_t fastin_read_pairs(bwa_seqio_t *ks, bam_pair_t *recs, size_t B, long *seqs, int broken_input, int drop_aligned)
{
	static fr_desc_t *desc; static size_t m_desc;
	size_t n = 0, n_fast = 0;
	int at_end = 0;
	*seqs = 0;
	if (B > m_desc) { m_desc = B; desc = (fr_desc_t *)realloc(desc, m_desc * sizeof(*desc)); }
	if (F.stage_pos && F.stage_pos == F.stage_len) F.stage_pos = F.stage_len = 0;
	else if (F.stage_pos) { memmove(F.stage, F.stage + F.stage_pos, F.stage_len - F.stage_pos); F.stage_len -= F.stage_pos; F.stage_pos = 0; }
	while (n < B && !at_end) {
		int fast = F.active && F.bgzf && !broken_input && !drop_aligned;
		if (fast) {
			const size_t p0 = F.stage_pos;
			int32_t l0, l1;
			fast = 0;
			if (fr_peek(p0, &l0)) {
				const uint8_t *r0 = F.stage + p0 + 4;
				const unsigned f0 = (unsigned)(r0[14] | r0[15] << 8);
				if (!(f0 & BAM_FPAIRED)) {
					desc[n].kind = 1; desc[n].off[0] = p0;
					F.stage_pos = p0 + 4 + (size_t)l0;
					fast = 1;
				} else if (fr_peek(p0 + 4 + (size_t)l0, &l1)) {
					const size_t p1 = p0 + 4 + (size_t)l0;
					const uint8_t *r1;
					unsigned f1, m0, m1;
					r0 = F.stage + p0 + 4; r1 = F.stage + p1 + 4; /* the staging area may have moved */
					f1 = (unsigned)(r1[14] | r1[15] << 8);
					m0 = f0 & (BAM_FPAIRED | BAM_FREAD1 | BAM_FREAD2); m1 = f1 & (BAM_FPAIRED | BAM_FREAD1 | BAM_FREAD2);
					if (strcmp((const char *)r0 + 32, (const char *)r1 + 32) == 0) {
						if (m0 == (BAM_FPAIRED | BAM_FREAD1) && m1 == (BAM_FPAIRED | BAM_FREAD2)) { desc[n].off[0] = p0; desc[n].off[1] = p1; fast = 1; }
						else if (m1 == (BAM_FPAIRED | BAM_FREAD1) && m0 == (BAM_FPAIRED | BAM_FREAD2)) { desc[n].off[0] = p1; desc[n].off[1] = p0; fast = 1; }
						if (fast) { desc[n].kind = 2; F.stage_pos = p1 + 4 + (size_t)l1; }
					}
				}
			}
			if (fast) { *seqs += desc[n].kind; ++n; ++n_fast; continue; }
		}
		{ /* the reference's reader, from wherever the stream stands */
			const int rc = read_bam_pair(ks, &recs[n], broken_input, drop_aligned);
			if (rc < 0) {
				fprintf(stderr, "[sequential_loop_pass1] error reading input BAM%s\n", rc == -2 ? " (lone mate)" : "");
				exit(1);
			}
			if (rc == 0) { at_end = 1; break; }
			desc[n].kind = 0;
			*seqs += recs[n].kind;
			++n;
		}
	}
	if (n_fast) {
		fr_ctx_t c = {recs, desc, F.stage};
		t_cpu_bucket = CPU_PARSE;
		parallel_for(n, 1024, fr_build_one, &c);
	}
	return n;
}

/* ================================================================== the intermediate records, kept in memory
 * The SAME encoded messages (the reference's msg_init_from_pair / pair_init_from_msg, so a record makes the same round
 * trip), up to BWAGPU_MEMTEMP_MB (default: a quarter of physical memory, at most 64 GB); whatever comes after that goes to
 * the reference's temporary file in the reference's format. */
#define MT_CHUNK ((size_t)64 << 20)
typedef struct {
	uint8_t **chunk; size_t n_chunk, m_chunk, used;
	uint8_t **rec; uint32_t *len; size_t n_rec, m_rec, rd;
	size_t bytes, cap;
	int spilled;
} memtemp_t;
static memtemp_t g_mt;

static size_t memtemp_cap(void)
{
	const char *e = getenv("BWAGPU_MEMTEMP_MB"), *eb = getenv("BWAGPU_MEMTEMP_BYTES");
	size_t cap;
	if (eb) return (size_t)atoll(eb);
	if (e) return (size_t)atoll(e) << 20;
	cap = (size_t)sysconf(_SC_PHYS_PAGES) * (size_t)sysconf(_SC_PAGESIZE) / 4;
	return cap > ((size_t)64 << 30) ? (size_t)64 << 30 : cap;
}

/* Chunks go back to a small pool instead of to the system: a process that runs bam2bam again (bench.py; a long-lived host)
 * appends the next run's records into pages it already owns -- a fresh 64 MB mapping costs 16 K page faults on the thread
 * that stores the records, a stage near the critical path.  memtemp_release() (the index is dropped) frees the pool. */
#define MT_POOL_MAX 64 /* x 64 MB */
static uint8_t *g_mt_pool[MT_POOL_MAX];
static size_t g_mt_pool_n;

static uint8_t *mt_chunk_get(void)
{
	if (g_mt_pool_n) return g_mt_pool[--g_mt_pool_n];
	return (uint8_t *)malloc(MT_CHUNK);
}

static void mt_chunk_put(uint8_t *c)
{
	if (g_mt_pool_n < MT_POOL_MAX) g_mt_pool[g_mt_pool_n++] = c;
	else free(c);
}

void memtemp_release(void)
{
	while (g_mt_pool_n) free(g_mt_pool[--g_mt_pool_n]);
}

void memtemp_begin(void)
{
	memtemp_free();
	g_mt.cap = memtemp_cap();
}

int memtemp_put(const void *data, uint32_t len)
{
	memtemp_t *t = &g_mt;
	if (t->spilled || t->bytes + len > t->cap || len > MT_CHUNK) { t->spilled = 1; return 0; }
	if (t->n_chunk == 0 || t->used + len > MT_CHUNK) {
		if (t->n_chunk == t->m_chunk) { t->m_chunk = t->m_chunk ? t->m_chunk << 1 : 16; t->chunk = (uint8_t **)realloc(t->chunk, t->m_chunk * sizeof(*t->chunk)); }
		t->chunk[t->n_chunk] = mt_chunk_get();
		if (!t->chunk[t->n_chunk]) { t->spilled = 1; return 0; }
		++t->n_chunk; t->used = 0;
	}
	if (t->n_rec == t->m_rec) {
		t->m_rec = t->m_rec ? t->m_rec << 1 : 1 << 20;
		t->rec = (uint8_t **)realloc(t->rec, t->m_rec * sizeof(*t->rec));
		t->len = (uint32_t *)realloc(t->len, t->m_rec * sizeof(*t->len));
	}
	t->rec[t->n_rec] = t->chunk[t->n_chunk - 1] + t->used;
	t->len[t->n_rec] = len;
	memcpy(t->rec[t->n_rec], data, len);
	++t->n_rec; t->used += len; t->bytes += len;
	return 1;
}

size_t memtemp_records(void) { return g_mt.n_rec; }
size_t memtemp_left(void) { return g_mt.n_rec - g_mt.rd; }
size_t memtemp_bytes(void) { return g_mt.bytes; }
int memtemp_spilled(void) { return g_mt.spilled; }

size_t memtemp_take(size_t want, size_t *first)
{
	size_t n = g_mt.n_rec - g_mt.rd < want ? g_mt.n_rec - g_mt.rd : want;
	*first = g_mt.rd;
	g_mt.rd += n;
	return n;
}

void memtemp_get(size_t idx, const uint8_t **data, uint32_t *len)
{
	*data = g_mt.rec[idx]; *len = g_mt.len[idx];
}

int memtemp_next(const uint8_t **data, uint32_t *len)
{
	if (g_mt.rd >= g_mt.n_rec) return 0;
	memtemp_get(g_mt.rd++, data, len);
	return 1;
}

void memtemp_free(void)
{
	size_t i;
	for (i = 0; i < g_mt.n_chunk; ++i) mt_chunk_put(g_mt.chunk[i]);
	free(g_mt.chunk); free(g_mt.rec); free(g_mt.len);
	memset(&g_mt, 0, sizeof(g_mt));
}

/* ================================================================== BAM output: BGZF blocks deflated on the host threads
 * The decompressed stream -- what any BAM reader sees -- is identical to the reference's; block boundaries are not (the
 * reference cuts at 65536 bytes of input, this writer at 65280 so that a block always fits). */
#define OB_IN 65280
#define OB_OUT 65536
typedef struct { bam_pair_t *recs; size_t *off; uint8_t *ubuf; size_t total; uint8_t *cbuf; int *clen; int level; int failed; } ob_ctx_t;

static size_t rec_bytes(const bam_pair_t *p) /* pair_print_bam's filter and bwa_print_bam1's record size */
{
	size_t n = 0;
	int i;
	if (g_only_aligned)
		for (i = 0; i != (int)p->kind; ++i)
			if (p->bam_rec[i].core.flag & SAM_FSU) return 0;
	for (i = 0; i != (int)p->kind; ++i) n += 4 + sizeof(bam1_core_t) + (size_t)p->bam_rec[i].data_len;
	return n;
}

static void ob_fill_one(size_t i, void *ctx)
{
	ob_ctx_t *c = (ob_ctx_t *)ctx;
	const bam_pair_t *p = &c->recs[i];
	uint8_t *q = c->ubuf + c->off[i];
	int j;
	if (c->off[i + 1] == c->off[i]) return;
	for (j = 0; j != (int)p->kind; ++j) { /* bwa_print_bam1 (bam2bam.c:304-321) */
		const bam1_t *b = &p->bam_rec[j];
		uint32_t w[9];
		w[0] = (uint32_t)(sizeof(bam1_core_t) + b->data_len);
		w[1] = (uint32_t)b->core.tid; w[2] = (uint32_t)b->core.pos;
		w[3] = (uint32_t)((int)b->core.bin << 16 | (int)b->core.qual << 8 | (int)b->core.l_qname);
		w[4] = (uint32_t)((int)b->core.flag << 16 | (int)b->core.n_cigar);
		w[5] = (uint32_t)b->core.l_qseq; w[6] = (uint32_t)b->core.mtid; w[7] = (uint32_t)b->core.mpos; w[8] = (uint32_t)b->core.isize;
		memcpy(q, w, 36); q += 36;
		memcpy(q, b->data, (size_t)b->data_len); q += b->data_len;
	}
}

static void ob_deflate_one(size_t k, void *ctx)
{
	ob_ctx_t *c = (ob_ctx_t *)ctx;
	const uint8_t *in = c->ubuf + k * OB_IN;
	const int in_len = (int)(c->total - k * OB_IN < OB_IN ? c->total - k * OB_IN : OB_IN);
	uint8_t *out = c->cbuf + k * OB_OUT;
	static const uint8_t hdr[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 0, 0}; /* bgzf.c:274-291 */
	z_stream zs;
	uint32_t crc, len;
	memcpy(out, hdr, 18);
	memset(&zs, 0, sizeof(zs));
	zs.next_in = (Bytef *)in; zs.avail_in = (uInt)in_len;
	zs.next_out = out + 18; zs.avail_out = OB_OUT - 18 - 8;
	if (deflateInit2(&zs, c->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK || deflate(&zs, Z_FINISH) != Z_STREAM_END) { c->failed = 1; return; }
	deflateEnd(&zs);
	len = (uint32_t)zs.total_out + 18 + 8;
	out[16] = (uint8_t)((len - 1) & 0xff); out[17] = (uint8_t)((len - 1) >> 8);
	crc = (uint32_t)crc32(crc32(0L, 0, 0), in, (uInt)in_len);
	memcpy(out + 18 + zs.total_out, &crc, 4);
	memcpy(out + 18 + zs.total_out + 4, &in_len, 4);
	c->clen[k] = (int)len;
}

/* the zlib writer (BWAGPU_HOST_DEFLATE=1; the A/B partner of the device codec) */
static void deflate_on_host(BGZF *output, ob_ctx_t *c, size_t nblk)
{
	static uint8_t *cbuf; static size_t m_cbuf;
	static int *clen; static size_t m_clen;
	size_t i;
	if (nblk * OB_OUT > m_cbuf) { m_cbuf = nblk * OB_OUT + (nblk / 4) * OB_OUT; cbuf = (uint8_t *)realloc(cbuf, m_cbuf); }
	if (nblk > m_clen) { m_clen = nblk + nblk / 4; clen = (int *)realloc(clen, m_clen * sizeof(int)); }
	c->cbuf = cbuf; c->clen = clen;
	t_cpu_bucket = CPU_DEFLATE;
	parallel_for(nblk, 4, ob_deflate_one, c);
	if (c->failed) { fprintf(stderr, "[bwa_gpu_batch] deflate failed\n"); exit(1); }
	{
		const double c0 = thread_cpu_now();
		for (i = 0; i < nblk; ++i) {
			if (fwrite(cbuf + i * OB_OUT, 1, (size_t)clen[i], output->file) != (size_t)clen[i]) { fprintf(stderr, "[bwa_gpu_batch] BAM write failed\n"); exit(1); }
			output->block_address += clen[i];
		}
		cpu_add(CPU_WRITE, thread_cpu_now() - c0);
	}
}

static int host_deflate_wanted(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("BWAGPU_HOST_DEFLATE"); v = e && atoi(e) != 0; }
	return v;
}

/* The batch's records laid out as BAM stream (host threads, straight into page-locked memory), cut into BGZF blocks and
 * deflated on the device (bwa_gpu_bgzf_deflate: what bgzf_write + deflate_block do, bgzf.c:265-330, 533-556), written with
 * one fwrite. */
void write_records_bam(BGZF *output, bam_pair_t *recs, size_t n)
{
	static size_t *off; static size_t m_off;
	static uint8_t *ubuf; static size_t m_ubuf; static int ubuf_pinned;
	ob_ctx_t c;
	size_t i, nblk;
	if (n + 1 > m_off) { m_off = n + 1; off = (size_t *)realloc(off, m_off * sizeof(*off)); }
	off[0] = 0;
	for (i = 0; i < n; ++i) off[i + 1] = off[i] + rec_bytes(&recs[i]);
	memset(&c, 0, sizeof(c));
	c.recs = recs; c.off = off; c.total = off[n]; c.level = output->compress_level;
	if (c.total == 0) return;
	if (bgzf_flush(output) != 0) { fprintf(stderr, "[bwa_gpu_batch] BAM write failed\n"); exit(1); } /* what bgzf_write buffered so far (the header) */
	nblk = (c.total + OB_IN - 1) / OB_IN;
	if (c.total > m_ubuf) {
		if (ubuf_pinned) bwa_gpu_host_free(ubuf); else free(ubuf);
		m_ubuf = c.total + c.total / 4;
		ubuf_pinned = !host_deflate_wanted();
		ubuf = ubuf_pinned ? (uint8_t *)bwa_gpu_host_alloc(m_ubuf) : (uint8_t *)malloc(m_ubuf);
		if (!ubuf) { fprintf(stderr, "[bwa_gpu_batch] output buffer of %zu bytes: %s\n", m_ubuf, ubuf_pinned ? bwa_gpu_last_error() : "out of memory"); exit(1); }
	}
	c.ubuf = ubuf;
	t_cpu_bucket = CPU_BAM_LAYOUT;
	parallel_for(n, 2048, ob_fill_one, &c);
	if (host_deflate_wanted()) { deflate_on_host(output, &c, nblk); return; }
	{
		const uint8_t *packed;
		int64_t packed_bytes;
		double t0 = shim_now(), c0;
		if (bwa_gpu_bgzf_deflate(ubuf, (int64_t)c.total, c.level, &packed, &packed_bytes, 0, 0, 0)) {
			fprintf(stderr, "[bwa_gpu_batch] bwa_gpu_bgzf_deflate: %s\n", bwa_gpu_last_error());
			exit(1);
		}
		shim_count_bgzf((int64_t)c.total, shim_now() - t0);
		c0 = thread_cpu_now();
		if (fwrite(packed, 1, (size_t)packed_bytes, output->file) != (size_t)packed_bytes) { fprintf(stderr, "[bwa_gpu_batch] BAM write failed\n"); exit(1); }
		output->block_address += packed_bytes;
		cpu_add(CPU_WRITE, thread_cpu_now() - c0);
	}
}
