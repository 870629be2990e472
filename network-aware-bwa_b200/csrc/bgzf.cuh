// bgzf.cuh -- BGZF (blocked gzip) deflate on the device: the BAM output side of bam2bam.
//
// Replaces the compute of bgzf.c:265-330 (deflate_block: one zlib deflate per <= 64 KB of BAM stream, level 2 as
// bam2bam.c:2061 opens its output "w2") -- the largest single item of host CPU time in a bam2bam run once the
// alignment itself is on the GPU.  The decompressed stream, which is what every BAM reader sees, is byte-identical to the
// reference's; the compressed bytes are this codec's own (deflate does not prescribe them).
//
// One CTA of 512 threads per BGZF block, everything in shared memory (positions are taken 512 at a time; BGZF_T=256 is the
// smaller form, 12 % slower):
//   0. the block's input (<= 65280 bytes) is loaded once; CRC-32 by T partial CRCs combined with x^(8 L 2^k) mod P;
//   1. LZ77, greedy (the reference's level 2 is zlib's deflate_fast): positions are taken T = 512 at a time --
//      A  warp 0 walks the batch's positions in order, 32 per step, through a 16 K-entry hash-head table of 4-byte strings:
//         candidate = the most recent earlier position with the same hash (__match_any_sync finds it inside the step);
//         the hashes themselves were computed by all threads while the previous batch was being emitted,
//      B  every thread extends its candidate (4 bytes per compare) -> match length / distance of its position,
//      C  the greedy parse -- "emit a token at p, continue at p + max(1, len)" -- is a pointer chain; the positions on
//         the chain from the carried-in start are found by pointer doubling (log2 T rounds),
//      D  the marked positions become tokens (literal | length, distance) in position order (ballot + scan), counted into
//         the literal/length and distance histograms;
//   2. length-limited canonical Huffman codes for both alphabets and for the code-length alphabet (rank sort in parallel,
//      two-queue merge on one thread, zlib's overflow repair for the 15 / 7 bit limits), dynamic-block header with
//      run-length coded code lengths; a stored block when that is smaller (incompressible input always fits 64 KB);
//   3. every token's bits are placed at their prefix-summed bit offset with shared-memory atomicOr;
//   4. the finished member (18-byte BGZF header, deflate stream, CRC-32, ISIZE) leaves shared memory in one aligned copy.
// All of it is deterministic: the same input gives the same bytes on every run.
#pragma once
#include <stdint.h>

namespace bgzf {

#ifndef BGZF_T
#define BGZF_T 512
#endif
constexpr int T = BGZF_T;          // threads per CTA = positions per LZ77 batch (256 or 512)
constexpr int LOG_T = T == 512 ? 9 : 8;
static_assert(T == 256 || T == 512, "the CRC tree and the pointer doubling are written for 256 or 512 threads");
constexpr int IN_MAX = 65280;      // input bytes per BGZF block (a stored block of this size still fits 64 KB)
constexpr int OUT_STRIDE = 65536;  // a finished member is at most this long (BSIZE is 16 bits)
constexpr int HBITS = 14;
constexpr int MIN_MATCH = 4, MAX_MATCH = 258, MAX_DIST = 32768;
constexpr int NLIT = 286, NDIST = 30, NCL = 19;
constexpr uint32_t NONE = 0xffffu;
constexpr uint32_t CRC_POLY = 0xedb88320u;

struct Smem {
	uint32_t buf[OUT_STRIDE / 4 + 4]; // the input bytes, later the output member
	uint16_t head[1 << HBITS];
	uint32_t lit_freq[288], dist_freq[32], cl_freq[20];
	uint16_t lit_code[288], dist_code[32], cl_code[20];
	uint8_t lit_len[288], dist_len[32], cl_len[20];
	uint16_t cand[T], hash[T], jump[2][T + 2];
	uint8_t mark[T + 4];
	uint32_t wsum[T / 32];
	// Huffman construction
	uint16_t sorted[288], parent[576];
	uint32_t weight[576];
	uint8_t depth[288];
	uint32_t bl_count[16], next_code[16];
	int n_used;
	// dynamic header
	uint8_t cl_sym[320], cl_ext[320];
	int n_cl, hlit, hdist;
	// CRC
	uint32_t crc_tab[256], crc_part[T], crc_pow[LOG_T];
	// block state
	int cur;
	uint32_t bitpos;
	int stored;
};

struct Params {
	const uint8_t *in;   // the whole input stream
	long long n_bytes;
	int n_blocks;        // ceil(n_bytes / IN_MAX)
	int level;           // 0: stored blocks only
	uint8_t *out;        // n_blocks x OUT_STRIDE
	int32_t *clen;       // member size per block
	uint32_t *tok;       // gridDim.x x 65536 token words
	uint32_t x2n[32];    // x^(2^k) mod P, k = 0..31 (reflected), for the CRC combination
};

// ---------------------------------------------------------------- CRC-32 (the gzip one), combination as zlib's multmodp
__host__ __device__ inline uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
	uint32_t m = 1u << 31, p = 0;
	for (;;) {
		if (a & m) {
			p ^= b;
			if ((a & (m - 1)) == 0) break;
		}
		m >>= 1;
		b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
	}
	return p;
}

// x^(n * 2^k) mod P
__host__ __device__ inline uint32_t crc_x2n(const uint32_t *x2n, uint32_t n, unsigned k)
{
	uint32_t p = 1u << 31;
	while (n) {
		if (n & 1) p = crc_mulmod(x2n[k & 31], p);
		n >>= 1;
		++k;
	}
	return p;
}

inline void crc_x2n_table(uint32_t *x2n)
{
	uint32_t p = 1u << 30; // x^1
	x2n[0] = p;
	for (int n = 1; n < 32; ++n) x2n[n] = p = crc_mulmod(p, p);
}

// ---------------------------------------------------------------- deflate's symbol maps (RFC 1951 3.2.5), closed form
__device__ __forceinline__ void len_symbol(int len, int &sym, int &ebits, int &eval)
{
	const int x = len - 3;
	if (len == 258) { sym = 285; ebits = 0; eval = 0; }
	else if (x < 8) { sym = 257 + x; ebits = 0; eval = 0; }
	else {
		const int e = 29 - __clz(x); // floor(log2 x) - 2
		sym = 261 + 4 * e + ((x >> e) & 3);
		ebits = e; eval = x & ((1 << e) - 1);
	}
}

__device__ __forceinline__ void dist_symbol(int dist, int &sym, int &ebits, int &eval)
{
	const int y = dist - 1;
	if (y < 4) { sym = y; ebits = 0; eval = 0; }
	else {
		const int e = 30 - __clz(y); // floor(log2 y) - 1
		sym = 2 * (e + 1) + ((y >> e) & 1);
		ebits = e; eval = y & ((1 << e) - 1);
	}
}

__device__ __forceinline__ int len_extra_bits(int sym) { return sym < 265 || sym == 285 ? 0 : (sym - 261) >> 2; }
__device__ __forceinline__ int dist_extra_bits(int sym) { return sym < 4 ? 0 : (sym >> 1) - 1; }

__device__ __forceinline__ uint32_t ld4(const uint32_t *w, int i) // the 4 bytes at byte offset i
{
	return __funnelshift_r(w[i >> 2], w[(i >> 2) + 1], (i & 3) << 3);
}

__device__ __forceinline__ void put_bits(uint32_t *w, uint32_t bitpos, unsigned long long v, int n)
{
	if (n == 0) return;
	const unsigned sh = bitpos & 31;
	uint32_t *q = w + (bitpos >> 5);
	const uint32_t w0 = (uint32_t)(v << sh);
	const unsigned long long rest = sh ? v >> (32 - sh) : v >> 32;
	if (w0) atomicOr(q, w0);
	if ((uint32_t)rest) atomicOr(q + 1, (uint32_t)rest);
	if ((uint32_t)(rest >> 32)) atomicOr(q + 2, (uint32_t)(rest >> 32));
}

// ---------------------------------------------------------------- length-limited canonical Huffman code (all threads)
// freq[0..n) -> len[0..n) (0 = unused) and the bit-reversed codes; at least two symbols get a code (zlib's build_tree does
// the same: a decoder wants a complete code).
__device__ void build_code(Smem &s, uint32_t *freq, int n, int maxbits, uint8_t *len, uint16_t *code)
{
	const int t = threadIdx.x;
	if (t == 0) {
		int nz = 0;
		for (int i = 0; i < n; ++i) nz += freq[i] != 0;
		for (int i = 0; nz < 2 && i < n; ++i)
			if (!freq[i]) { freq[i] = 1; ++nz; }
		s.n_used = nz;
	}
	__syncthreads();
	for (int i = t; i < n; i += T) { // rank sort by (frequency, symbol)
		const uint32_t f = freq[i];
		len[i] = 0;
		if (f) {
			int r = 0;
			for (int j = 0; j < n; ++j) {
				const uint32_t g = freq[j];
				r += (g != 0) && (g < f || (g == f && j < i));
			}
			s.sorted[r] = (uint16_t)i;
			s.weight[r] = f;
		}
	}
	__syncthreads();
	const int m = s.n_used, root = 2 * m - 2;
	if (t == 0) { // leaves 0..m-1 in rising weight, internal nodes m..2m-2 in the order they are made (rising weight too)
		int lq = 0, iq = m, in = m;
		for (int k = 0; k < m - 1; ++k) {
			int a, b;
			if (lq < m && (iq >= in || s.weight[lq] <= s.weight[iq])) a = lq++; else a = iq++;
			if (lq < m && (iq >= in || s.weight[lq] <= s.weight[iq])) b = lq++; else b = iq++;
			s.weight[in] = s.weight[a] + s.weight[b];
			s.parent[a] = (uint16_t)in; s.parent[b] = (uint16_t)in;
			++in;
		}
	}
	__syncthreads();
	for (int i = t; i < m; i += T) {
		int d = 0, id = i;
		while (id != root) { id = s.parent[id]; ++d; }
		s.depth[i] = (uint8_t)(d > maxbits ? maxbits : d);
	}
	__syncthreads();
	if (t == 0) {
		unsigned long long kraft = 0;
		for (int b = 0; b <= maxbits; ++b) s.bl_count[b] = 0;
		for (int i = 0; i < m; ++i) { ++s.bl_count[s.depth[i]]; kraft += 1ull << (maxbits - s.depth[i]); }
		// every leaf clamped to maxbits takes one slot too many: zlib's repair (trees.c gen_bitlen) -- split the deepest
		// leaf above the limit, hang one of the surplus leaves next to it
		for (long long excess = (long long)(kraft - (1ull << maxbits)); excess > 0; --excess) {
			int bits = maxbits - 1;
			while (s.bl_count[bits] == 0) --bits;
			--s.bl_count[bits]; s.bl_count[bits + 1] += 2; --s.bl_count[maxbits];
		}
		int i = 0;
		for (int bits = maxbits; bits >= 1; --bits) // rarest symbols get the longest codes
			for (uint32_t c = s.bl_count[bits]; c; --c) len[s.sorted[i++]] = (uint8_t)bits;
		uint32_t c = 0;
		s.bl_count[0] = 0;
		for (int bits = 1; bits <= maxbits; ++bits) { c = (c + s.bl_count[bits - 1]) << 1; s.next_code[bits] = c; }
	}
	__syncthreads();
	for (int i = t; i < n; i += T) {
		const int l = len[i];
		uint32_t c = 0;
		if (l) {
			int r = 0;
			for (int j = 0; j < i; ++j) r += len[j] == l;
			c = __brev(s.next_code[l] + r) >> (32 - l);
		}
		code[i] = (uint16_t)c;
	}
	__syncthreads();
}

// exclusive prefix of v over the block (by thread index) and the block total; one barrier, s.wsum is free again after the
// caller's next barrier
__device__ __forceinline__ uint32_t block_scan(Smem &s, uint32_t v, uint32_t &total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += o;
	}
	if (lane == 31) s.wsum[warp] = inc;
	__syncthreads();
	uint32_t before = 0, all = 0;
#pragma unroll
	for (int w = 0; w < T / 32; ++w) {
		const uint32_t x = s.wsum[w];
		if (w < warp) before += x;
		all += x;
	}
	total = all;
	return before + inc - v;
}

// hash of the 4 bytes at p for the head table; positions too close to the end get a value no table entry and no other lane has
__device__ __forceinline__ uint32_t pos_hash(const Smem &s, int p, int len)
{
	if (p + MIN_MATCH > len) return 0x4000u + (threadIdx.x & 31);
	return (ld4(s.buf, p) * 2654435761u) >> (32 - HBITS);
}

// ---------------------------------------------------------------- one BGZF block
__device__ void deflate_block(Smem &s, const uint8_t *__restrict__ in, int len, int level, uint8_t *__restrict__ out, int32_t *clen,
                              uint32_t *__restrict__ tok, const uint32_t *x2n)
{
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	uint8_t *bytes = (uint8_t *)s.buf;

	// ---- 0. load, clear, CRC
	__syncthreads(); // the previous block's copy-out is done
	for (int i = t; i < OUT_STRIDE / 4 + 4; i += T) s.buf[i] = 0;
	for (int i = t; i < (1 << HBITS); i += T) s.head[i] = (uint16_t)NONE;
	for (int i = t; i < 288; i += T) s.lit_freq[i] = 0;
	if (t < 32) s.dist_freq[t] = 0;
	if (t < 20) s.cl_freq[t] = 0;
	if (t < 256) {
		uint32_t c = (uint32_t)t;
		for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
		s.crc_tab[t] = c;
	}
	if (t == 0) { s.cur = 0; s.stored = 0; }
	__syncthreads();
	if ((((uintptr_t)in) & 3) == 0) {
		const uint32_t *in4 = (const uint32_t *)in;
		for (int i = t; i < (len >> 2); i += T) s.buf[i] = in4[i];
		for (int i = (len & ~3) + t; i < len; i += T) bytes[i] = in[i];
	} else {
		for (int i = t; i < len; i += T) bytes[i] = in[i];
	}
	__syncthreads();
	uint32_t crc;
	{
		// the message with its first four bytes inverted, zero-padded IN FRONT to 256 L bytes, remainder of division by P
		// (leading zeros do not change it); thread t takes the t-th run of L bytes
		const int L = (len + T - 1) / T, pad = L * T - len;
		uint32_t c = 0;
		int lo = t * L - pad, hi = lo + L;
		if (lo < 0) lo = 0;
		for (int i = lo; i < hi; ++i) {
			uint32_t b = bytes[i];
			if (i < 4) b ^= 0xffu;
			c = s.crc_tab[(c ^ b) & 0xff] ^ (c >> 8);
		}
		s.crc_part[t] = c;
		s.hash[t] = (uint16_t)pos_hash(s, t, len); // the first batch's hashes (phase A reads them after the barriers below)
		if (t == 0) {
			uint32_t p = crc_x2n(x2n, (uint32_t)L, 3); // x^(8 L)
			for (int k = 0; k < LOG_T; ++k) { s.crc_pow[k] = p; p = crc_mulmod(p, p); }
		}
		__syncthreads();
		for (int k = 0; k < LOG_T; ++k) {
			const int span = 1 << k;
			if ((t & (2 * span - 1)) == 0) s.crc_part[t] = crc_mulmod(s.crc_pow[k], s.crc_part[t]) ^ s.crc_part[t + span];
			__syncthreads();
		}
		crc = ~s.crc_part[0];
		if (len < 4) { // too short for the inverted-prefix form: plainly
			uint32_t c2 = 0xffffffffu;
			for (int i = 0; i < len; ++i) c2 = s.crc_tab[(c2 ^ bytes[i]) & 0xff] ^ (c2 >> 8);
			crc = ~c2;
		}
	}

	// ---- 1. LZ77
	uint32_t ntok = 0;
	int cur = 0;
	if (level > 0)
		for (int base = 0; base < len; base += T) {
			if (warp == 0) { // A: the hashes were computed by all threads during the previous batch; only the table walk is serial
				uint32_t hq[T / 32];
				unsigned sameq[T / 32];
#pragma unroll
				for (int sub = 0; sub < T / 32; ++sub) hq[sub] = s.hash[sub * 32 + lane];
#pragma unroll
				for (int sub = 0; sub < T / 32; ++sub) sameq[sub] = __match_any_sync(0xffffffffu, hq[sub]);
#pragma unroll
				for (int sub = 0; sub < T / 32; ++sub) {
					const int p = base + sub * 32 + lane;
					const uint32_t h = hq[sub];
					const bool valid = h < (1u << HBITS);
					const unsigned same = sameq[sub], lower = same & ((1u << lane) - 1);
					uint32_t c = valid ? s.head[h] : NONE;
					if (valid && lower) c = (uint32_t)(p - lane + (31 - __clz(lower)));
					if (valid && (same >> lane) == 1u) s.head[h] = (uint16_t)p; // the last position of the step with this hash
					s.cand[sub * 32 + lane] = (uint16_t)c;
					__syncwarp();
				}
			}
			__syncthreads();
			// B
			const int p = base + t;
			const int c = s.cand[t];
			int ml = 0;
			if (c != (int)NONE && p - c <= MAX_DIST) {
				const int maxlen = len - p < MAX_MATCH ? len - p : MAX_MATCH;
				bool open = true;
				while (open && ml + 4 <= maxlen) {
					const uint32_t x = ld4(s.buf, p + ml) ^ ld4(s.buf, c + ml);
					if (x) { ml += (__ffs((int)x) - 1) >> 3; open = false; }
					else ml += 4;
				}
				while (open && ml < maxlen && bytes[p + ml] == bytes[c + ml]) ++ml;
				if (ml < MIN_MATCH) ml = 0;
			}
			const int step = ml ? ml : 1;
			s.jump[0][t] = (uint16_t)(t + step < T ? t + step : T);
			s.mark[t] = (uint8_t)(p == cur);
			if (t == 0) { s.jump[0][T] = T; s.jump[1][T] = T; s.mark[T] = 0; }
			__syncthreads();
			// C: after round r the first 2^(r+1) positions of the chain are marked, jump = 2^(r+1) steps along it
#pragma unroll 1
			for (int r = 0; r < LOG_T; ++r) {
				const int j = s.jump[r & 1][t];
				if (s.mark[t]) s.mark[j] = 1;
				s.jump[(r + 1) & 1][t] = s.jump[r & 1][j];
				__syncthreads();
			}
			// D
			const bool is_tok = s.mark[t] != 0 && p < len;
			if (is_tok && t + step >= T) s.cur = p + step;
			s.hash[t] = (uint16_t)pos_hash(s, p + T, len); // the next batch's (warp 0 is done with this batch's since phase A)
			uint32_t total;
			const uint32_t at = ntok + block_scan(s, is_tok ? 1u : 0u, total);
			ntok += total;
			cur = s.cur;
			if (is_tok) {
				if (ml) {
					int sym, eb, ev;
					tok[at] = 0x80000000u | (uint32_t)(ml - 3) << 16 | (uint32_t)(p - c - 1);
					len_symbol(ml, sym, eb, ev);
					atomicAdd(&s.lit_freq[sym], 1u);
					dist_symbol(p - c, sym, eb, ev);
					atomicAdd(&s.dist_freq[sym], 1u);
				} else {
					const uint32_t b = bytes[p];
					tok[at] = b;
					atomicAdd(&s.lit_freq[b], 1u);
				}
			}
		}
	if (t == 0) s.lit_freq[256] = 1;
	__syncthreads();

	// ---- 2. codes, header
	if (level > 0) {
		build_code(s, s.lit_freq, NLIT, 15, s.lit_len, s.lit_code);
		build_code(s, s.dist_freq, NDIST, 15, s.dist_len, s.dist_code);
		if (t == 0) { // the code lengths, run-length coded (RFC 1951 3.2.7)
			int hlit = NLIT, hdist = NDIST, n = 0;
			while (hlit > 257 && s.lit_len[hlit - 1] == 0) --hlit;
			while (hdist > 1 && s.dist_len[hdist - 1] == 0) --hdist;
			s.hlit = hlit; s.hdist = hdist;
			const int tot = hlit + hdist;
			int i = 0;
			while (i < tot) {
				const int v = i < hlit ? s.lit_len[i] : s.dist_len[i - hlit];
				int run = 1;
				while (i + run < tot && (i + run < hlit ? s.lit_len[i + run] : s.dist_len[i + run - hlit]) == v) ++run;
				i += run;
				if (v == 0) {
					while (run >= 11) { const int r = run < 138 ? run : 138; s.cl_sym[n] = 18; s.cl_ext[n++] = (uint8_t)(r - 11); run -= r; }
					if (run >= 3) { s.cl_sym[n] = 17; s.cl_ext[n++] = (uint8_t)(run - 3); run = 0; }
					while (run-- > 0) { s.cl_sym[n] = 0; s.cl_ext[n++] = 0; }
				} else {
					s.cl_sym[n] = (uint8_t)v; s.cl_ext[n++] = 0; --run;
					while (run >= 3) { const int r = run < 6 ? run : 6; s.cl_sym[n] = 16; s.cl_ext[n++] = (uint8_t)(r - 3); run -= r; }
					while (run-- > 0) { s.cl_sym[n] = (uint8_t)v; s.cl_ext[n++] = 0; }
				}
			}
			s.n_cl = n;
			for (int k = 0; k < n; ++k) ++s.cl_freq[s.cl_sym[k]];
		}
		__syncthreads();
		build_code(s, s.cl_freq, NCL, 7, s.cl_len, s.cl_code);
	}
	for (int i = t; i < OUT_STRIDE / 4 + 4; i += T) s.buf[i] = 0; // the input is no longer needed: the member is assembled here
	__syncthreads();
	if (t == 0) {
		const uint8_t order[NCL] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
		uint32_t bp = 18 * 8;
		int stored = level <= 0;
		if (!stored) {
			int hclen = NCL;
			while (hclen > 4 && s.cl_len[order[hclen - 1]] == 0) --hclen;
			unsigned long long bits = 3 + 5 + 5 + 4 + 3 * (unsigned long long)hclen;
			for (int k = 0; k < s.n_cl; ++k) bits += s.cl_len[s.cl_sym[k]] + (s.cl_sym[k] == 16 ? 2 : s.cl_sym[k] == 17 ? 3 : s.cl_sym[k] == 18 ? 7 : 0);
			for (int i = 0; i < NLIT; ++i) bits += (unsigned long long)s.lit_freq[i] * (s.lit_len[i] + (i > 256 ? len_extra_bits(i) : 0));
			for (int i = 0; i < NDIST; ++i) bits += (unsigned long long)s.dist_freq[i] * (s.dist_len[i] + dist_extra_bits(i));
			if ((bits + 7) / 8 >= (unsigned long long)len + 5) stored = 1;
			else {
				put_bits(s.buf, bp, 1 | 2 << 1, 3); bp += 3;
				put_bits(s.buf, bp, (unsigned)(s.hlit - 257), 5); bp += 5;
				put_bits(s.buf, bp, (unsigned)(s.hdist - 1), 5); bp += 5;
				put_bits(s.buf, bp, (unsigned)(hclen - 4), 4); bp += 4;
				for (int k = 0; k < hclen; ++k) { put_bits(s.buf, bp, s.cl_len[order[k]], 3); bp += 3; }
				for (int k = 0; k < s.n_cl; ++k) {
					const int sym = s.cl_sym[k];
					put_bits(s.buf, bp, s.cl_code[sym], s.cl_len[sym]); bp += s.cl_len[sym];
					const int eb = sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0;
					put_bits(s.buf, bp, s.cl_ext[k], eb); bp += eb;
				}
			}
		}
		s.stored = stored;
		s.bitpos = bp;
	}
	__syncthreads();

	// ---- 3. the stream
	uint32_t payload; // bytes of deflate stream
	if (s.stored) {
		if (t == 0) {
			bytes[18] = 1; // BFINAL = 1, BTYPE = 00, padding
			bytes[19] = (uint8_t)(len & 0xff); bytes[20] = (uint8_t)(len >> 8);
			bytes[21] = (uint8_t)(~len & 0xff); bytes[22] = (uint8_t)((~len >> 8) & 0xff);
		}
		for (int i = t; i < len; i += T) bytes[23 + i] = in[i];
		payload = (uint32_t)len + 5;
	} else {
		uint32_t bp = s.bitpos;
		for (uint32_t k0 = 0; k0 < ntok; k0 += 4 * T) {
			unsigned long long v[4];
			int nb[4];
			uint32_t mine = 0;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t k = k0 + 4 * t + q;
				v[q] = 0; nb[q] = 0;
				if (k < ntok) {
					const uint32_t w = tok[k];
					if (w & 0x80000000u) {
						int sym, eb, ev, n;
						len_symbol((int)((w >> 16) & 0xff) + 3, sym, eb, ev);
						v[q] = s.lit_code[sym]; n = s.lit_len[sym];
						v[q] |= (unsigned long long)ev << n; n += eb;
						dist_symbol((int)(w & 0x7fff) + 1, sym, eb, ev);
						v[q] |= (unsigned long long)s.dist_code[sym] << n; n += s.dist_len[sym];
						v[q] |= (unsigned long long)ev << n; n += eb;
						nb[q] = n;
					} else {
						v[q] = s.lit_code[w]; nb[q] = s.lit_len[w];
					}
				}
				mine += (uint32_t)nb[q];
			}
			uint32_t total;
			uint32_t at = bp + block_scan(s, mine, total);
			bp += total;
#pragma unroll
			for (int q = 0; q < 4; ++q) { put_bits(s.buf, at, v[q], nb[q]); at += (uint32_t)nb[q]; }
			__syncthreads(); // s.wsum is reused by the next round
		}
		if (t == 0) put_bits(s.buf, bp, s.lit_code[256], s.lit_len[256]);
		bp += s.lit_len[256];
		payload = (bp - 18 * 8 + 7) >> 3;
	}
	__syncthreads();

	// ---- 4. member header / trailer (bgzf.c:274-291, 318-326), copy-out
	const uint32_t total = 18 + payload + 8;
	if (t == 0) {
		const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};
		for (int i = 0; i < 16; ++i) bytes[i] = hdr[i];
		bytes[16] = (uint8_t)((total - 1) & 0xff); bytes[17] = (uint8_t)((total - 1) >> 8);
		uint8_t *q = bytes + 18 + payload;
		for (int i = 0; i < 4; ++i) { q[i] = (uint8_t)(crc >> (8 * i)); q[4 + i] = (uint8_t)((uint32_t)len >> (8 * i)); }
		*clen = (int32_t)total;
	}
	__syncthreads();
	uint32_t *out4 = (uint32_t *)out;
	for (uint32_t i = t; i < (total + 3) / 4; i += T) out4[i] = s.buf[i];
}

} // namespace bgzf

// ================================================================ inflate: the BAM input side
// Replaces the compute of bamlite's gzread (bamlite.h:7-11, bamlite.c:125-155: one zlib inflate stream on the thread
// that parses the records).  The members of a BGZF file are independent, so the parallelism is across members: ONE THREAD
// per member, thousands of members in flight -- a member is a serial bit stream (every code's position depends on the
// lengths of all codes before it), and a thread walks it with the canonical-code decoder of RFC 1951 3.2.2 (count of
// codes per length + symbols in code order; a code is recognised the moment its value falls inside its length's range).
// The per-thread tables (lit/len: 16 counts + 288 symbols; distance: 16 + 30) live in shared memory, the bit buffer in
// a register; literals go out as byte stores into the member's slice of the output stream, matches are byte copies from
// that slice (overlap = run-length semantics for free).  Stored, fixed and dynamic blocks are all handled; a malformed
// member sets a status the host turns into an error.
namespace bgzf {

constexpr int INF_T = 64; // threads per CTA (700 B of tables each)

struct InfTables {
	uint16_t lcnt[16], lsym[288], dcnt[16], dsym[32];
};

struct InfJob {
	long long in_off, out_off;
	int in_len, out_len; // the whole member (header .. ISIZE); its ISIZE
};

struct BitIn {
	const uint8_t *p, *end;
	unsigned long long buf;
	int n;
	int fake; // zero bits appended after the end of the stream (a code is looked at 15 bits at a time, so some may be needed)
};
__device__ __forceinline__ bool bits_over(const BitIn &b) { return b.n < b.fake; } // bits past the end were CONSUMED

__device__ __forceinline__ void bits_need(BitIn &b, int want) // want <= 32
{
	while (b.n < want) {
		unsigned long long v = 0;
		if (b.p < b.end) v = *b.p++; else b.fake += 8;
		b.buf |= v << b.n;
		b.n += 8;
	}
}
__device__ __forceinline__ uint32_t bits_take(BitIn &b, int k) // k <= 16
{
	bits_need(b, k);
	const uint32_t v = (uint32_t)(b.buf & ((1ull << k) - 1));
	b.buf >>= k; b.n -= k;
	return v;
}

// one symbol of a canonical code: cnt[len] codes of each length, sym[] in (length, symbol) order; -1 = no such code
__device__ __forceinline__ int inf_decode(BitIn &b, const uint16_t *cnt, const uint16_t *sym)
{
	bits_need(b, 15);
	int code = 0, first = 0, index = 0;
	unsigned long long w = b.buf;
#pragma unroll 1
	for (int len = 1; len <= 15; ++len) {
		code |= (int)(w & 1); w >>= 1;
		const int c = cnt[len];
		if (code - c < first) { b.buf >>= len; b.n -= len; return sym[index + (code - first)]; }
		index += c; first += c;
		first <<= 1; code <<= 1;
	}
	return -1;
}

// code lengths -> decoder tables; false: over-subscribed (an incomplete code is fine: zlib accepts it for a lone distance code)
__device__ bool inf_build(const uint8_t *lens, int n, uint16_t *cnt, uint16_t *sym)
{
	uint16_t offs[16];
	for (int l = 0; l < 16; ++l) cnt[l] = 0;
	for (int s = 0; s < n; ++s) ++cnt[lens[s]];
	int left = 1;
	for (int l = 1; l < 16; ++l) { left <<= 1; left -= cnt[l]; if (left < 0) return false; }
	offs[1] = 0;
	for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
	for (int s = 0; s < n; ++s)
		if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
	cnt[0] = 0;
	return true;
}

// returns 0, or a code saying what was wrong
__device__ int inflate_member(const uint8_t *in, int in_len, uint8_t *out, int out_len, InfTables &T) // (out is read back by the match copies: no __restrict__)
{
	if (in_len < 18 || in[0] != 31 || in[1] != 139 || in[2] != 8) return 1;
	const int flg = in[3];
	int h = 10;
	if (flg & 4) { if (h + 2 > in_len) return 1; h += 2 + (in[h] | in[h + 1] << 8); }
	if (flg & 8) { while (h < in_len && in[h]) ++h; ++h; }
	if (flg & 16) { while (h < in_len && in[h]) ++h; ++h; }
	if (flg & 2) h += 2;
	if (h + 8 > in_len) return 1;
	BitIn b;
	b.p = in + h; b.end = in + in_len - 8; b.buf = 0; b.n = 0; b.fake = 0;
	int pos = 0;
	for (;;) {
		const uint32_t last = bits_take(b, 1), type = bits_take(b, 2);
		if (type == 0) { // stored
			b.buf >>= (b.n & 7); b.n -= (b.n & 7);
			const uint32_t len = bits_take(b, 16), nlen = bits_take(b, 16);
			if ((len ^ nlen) != 0xffffu) return 2;
			if (pos + (int)len > out_len) return 3;
			for (uint32_t i = 0; i < len; ++i) out[pos++] = (uint8_t)bits_take(b, 8);
		} else if (type == 3) return 4;
		else {
			uint8_t lens[320];
			if (type == 1) { // fixed code (RFC 1951 3.2.6)
				for (int s = 0; s < 288; ++s) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
				if (!inf_build(lens, 288, T.lcnt, T.lsym)) return 5;
				for (int s = 0; s < 30; ++s) lens[s] = 5;
				if (!inf_build(lens, 30, T.dcnt, T.dsym)) return 5;
			} else {
				const int hlit = (int)bits_take(b, 5) + 257, hdist = (int)bits_take(b, 5) + 1, hclen = (int)bits_take(b, 4) + 4;
				if (hlit > 286 || hdist > 30) return 6;
				const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
				for (int k = 0; k < 19; ++k) lens[order[k]] = k < hclen ? (uint8_t)bits_take(b, 3) : 0;
				if (!inf_build(lens, 19, T.dcnt, T.dsym)) return 7; // the code-length code borrows the distance tables
				int k = 0;
				while (k < hlit + hdist) {
					const int s = inf_decode(b, T.dcnt, T.dsym);
					if (s < 0) return 8;
					if (s < 16) lens[k++] = (uint8_t)s;
					else {
						int rep, v = 0;
						if (s == 16) { if (k == 0) return 8; v = lens[k - 1]; rep = 3 + (int)bits_take(b, 2); }
						else if (s == 17) rep = 3 + (int)bits_take(b, 3);
						else rep = 11 + (int)bits_take(b, 7);
						if (k + rep > hlit + hdist) return 8;
						while (rep--) lens[k++] = (uint8_t)v;
					}
				}
				if (lens[256] == 0) return 9; // no end-of-block code
				if (!inf_build(lens, hlit, T.lcnt, T.lsym)) return 10;
				if (!inf_build(lens + hlit, hdist, T.dcnt, T.dsym)) return 10;
			}
			for (;;) {
				const int s = inf_decode(b, T.lcnt, T.lsym);
				if (s < 0) return 11;
				if (bits_over(b)) return 15;
				if (s < 256) { if (pos >= out_len) return 3; out[pos++] = (uint8_t)s; }
				else if (s == 256) break;
				else {
					if (s > 285) return 12;
					int len;
					if (s < 265) len = s - 254;
					else if (s == 285) len = 258;
					else { const int e = (s - 261) >> 2; len = 3 + ((4 + ((s - 261) & 3)) << e) + (int)bits_take(b, e); }
					const int ds = inf_decode(b, T.dcnt, T.dsym);
					if (ds < 0 || ds > 29) return 12;
					int dist;
					if (ds < 4) dist = ds + 1;
					else { const int e = (ds >> 1) - 1; dist = 1 + ((2 + (ds & 1)) << e) + (int)bits_take(b, e); }
					if (dist > pos) return 13;
					if (pos + len > out_len) return 3;
					const uint8_t *src = out + pos - dist;
					for (int i = 0; i < len; ++i) out[pos + i] = src[i];
					pos += len;
				}
			}
		}
		if (bits_over(b)) return 15;
		if (last) break;
	}
	if (pos != out_len) return 14;
	uint32_t isize;
	const uint8_t *t = in + in_len - 4;
	isize = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
	return isize == (uint32_t)out_len ? 0 : 14;
}


// ---------------------------------------------------------------- one WARP per member
// inflate_member above is the plain statement of the decoder (and what the emulated tests pin first); as a kernel it is slow --
// 32 members in a warp diverge at every symbol, and every match byte is an L2 round trip.  The product kernel gives a member
// a warp: lane 0 walks the bit stream (a 10-bit lookup table per alphabet, built by all lanes from the canonical tables;
// longer codes fall back to the canonical walk; the bit buffer is refilled 32 bits at a time from a word loaded one refill
// earlier) and stores the literals; at a match the warp copies together, 32 bytes per step.
constexpr int INFW_WARPS = 8;
constexpr int INF_LB = 10, INF_DB = 9;

struct InfWarp {
	uint16_t lit[1 << INF_LB], dist[1 << INF_DB]; // symbol | code length << 9; 0 = no code this short
	InfTables T;
	uint16_t first[2][16], base[2][16]; // canonical code of the first symbol of each length / its index in T.*sym
	uint8_t lens[320];
};

struct WordIn { // lane 0's view of the stream: aligned 32-bit loads, one word ahead
	const uint32_t *w;
	const uint8_t *start;
	long long total_bits; // bits the stream really has
	unsigned long long buf;
	uint32_t pre, pre2; // the next two words of the stream, loaded two refills (64 bits of codes) before they are needed
	int n;
	long long loaded_bits;
};

__device__ __forceinline__ void win_refill(WordIn &b)
{
	if (b.n <= 32) {
		b.buf |= (unsigned long long)b.pre << b.n;
		b.n += 32; b.loaded_bits += 32;
		b.pre = b.pre2;
		b.pre2 = *b.w++;
	}
}
__device__ __forceinline__ uint32_t win_take(WordIn &b, int k) // k <= 16
{
	win_refill(b);
	const uint32_t v = (uint32_t)(b.buf & ((1ull << k) - 1));
	b.buf >>= k; b.n -= k;
	return v;
}
__device__ __forceinline__ bool win_over(const WordIn &b) { return b.loaded_bits - b.n > b.total_bits; }

__device__ __forceinline__ int win_decode_slow(WordIn &b, const uint16_t *cnt, const uint16_t *sym)
{
	win_refill(b);
	int code = 0, first = 0, index = 0;
	unsigned long long w = b.buf;
#pragma unroll 1
	for (int len = 1; len <= 15; ++len) {
		code |= (int)(w & 1); w >>= 1;
		const int c = cnt[len];
		if (code - c < first) { b.buf >>= len; b.n -= len; return sym[index + (code - first)]; }
		index += c; first += c;
		first <<= 1; code <<= 1;
	}
	return -1;
}

// lookup table of one alphabet from its canonical tables, all lanes (which = 0: literal/length, 1: distance)
__device__ void infw_fill(InfWarp &W, int which, int lane)
{
	uint16_t *tab = which ? W.dist : W.lit;
	const int bits = which ? INF_DB : INF_LB;
	const uint16_t *cnt = which ? W.T.dcnt : W.T.lcnt, *sym = which ? W.T.dsym : W.T.lsym;
	for (int i = lane; i < (1 << bits); i += 32) tab[i] = 0;
	if (lane == 0) {
		int code = 0, idx = 0;
		for (int l = 1; l < 16; ++l) { W.first[which][l] = (uint16_t)code; W.base[which][l] = (uint16_t)idx; code = (code + cnt[l]) << 1; idx += cnt[l]; }
		W.base[which][0] = (uint16_t)idx; // number of coded symbols
	}
	__syncwarp();
	const int n_coded = W.base[which][0];
	for (int idx = lane; idx < n_coded; idx += 32) {
		int l = 1;
		while (l < 15 && idx >= W.base[which][l] + cnt[l]) ++l;
		if (l > bits) continue;
		const uint32_t code = (uint32_t)W.first[which][l] + (uint32_t)(idx - W.base[which][l]);
		const uint32_t r = __brev(code) >> (32 - l);
		const uint16_t e = (uint16_t)(sym[idx] | l << 9);
		for (uint32_t k = r; k < (1u << bits); k += 1u << l) tab[k] = e;
	}
	__syncwarp();
}

// the whole warp calls this; returns the member's status (the same on every lane)
__device__ int inflate_member_warp(const uint8_t *in, int in_len, uint8_t *out, int out_len, InfWarp &W)
{
	const int lane = threadIdx.x & 31;
	const unsigned full = 0xffffffffu;
	int h = 10, err = 0;
	if (in_len < 18 || in[0] != 31 || in[1] != 139 || in[2] != 8) return 1;
	{
		const int flg = in[3];
		if (flg & 4) { if (h + 2 > in_len) return 1; h += 2 + (in[h] | in[h + 1] << 8); }
		if (flg & 8) { while (h < in_len && in[h]) ++h; ++h; }
		if (flg & 16) { while (h < in_len && in[h]) ++h; ++h; }
		if (flg & 2) h += 2;
		if (h + 8 > in_len) return 1;
	}
	WordIn b; // meaningful on lane 0
	{
		const uint8_t *p = in + h;
		b.start = p; b.total_bits = 8ll * (in_len - 8 - h);
		b.buf = 0; b.n = 0; b.loaded_bits = 0;
		while (((uintptr_t)p & 3) != 0) { b.buf |= (unsigned long long)*p++ << b.n; b.n += 8; b.loaded_bits += 8; } // up to 3 bytes: to a word boundary
		b.w = (const uint32_t *)p;
		b.pre = *b.w++;
		b.pre2 = *b.w++;
	}
	int pos = 0;
	for (;;) { // one deflate block per trip
		int type = 0, last = 0, slen = 0;
		long long soff = 0;
		if (lane == 0) {
			last = (int)win_take(b, 1); type = (int)win_take(b, 2);
			if (type == 0) { // stored: where its bytes are
				const int drop = b.n & 7;
				b.buf >>= drop; b.n -= drop;
				const uint32_t len = win_take(b, 16), nlen = win_take(b, 16);
				if ((len ^ nlen) != 0xffffu) err = 2;
				else if (pos + (int)len > out_len) err = 3;
				slen = (int)len;
				soff = (b.loaded_bits - b.n) >> 3; // bytes of the stream consumed so far
			} else if (type == 3) err = 4;
			else if (type == 2) { // the code lengths
				const int hlit = (int)win_take(b, 5) + 257, hdist = (int)win_take(b, 5) + 1, hclen = (int)win_take(b, 4) + 4;
				if (hlit > 286 || hdist > 30) err = 6;
				else {
					const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
					for (int k = 0; k < 19; ++k) W.lens[order[k]] = k < hclen ? (uint8_t)win_take(b, 3) : 0;
					if (!inf_build(W.lens, 19, W.T.dcnt, W.T.dsym)) err = 7;
					int k = 0;
					while (!err && k < hlit + hdist) {
						const int s = win_decode_slow(b, W.T.dcnt, W.T.dsym);
						if (s < 0) err = 8;
						else if (s < 16) W.lens[k++] = (uint8_t)s;
						else {
							int rep, v = 0;
							if (s == 16) { if (k == 0) { err = 8; break; } v = W.lens[k - 1]; rep = 3 + (int)win_take(b, 2); }
							else if (s == 17) rep = 3 + (int)win_take(b, 3);
							else rep = 11 + (int)win_take(b, 7);
							if (k + rep > hlit + hdist) { err = 8; break; }
							while (rep--) W.lens[k++] = (uint8_t)v;
						}
					}
					if (!err && W.lens[256] == 0) err = 9;
					if (!err && !inf_build(W.lens, hlit, W.T.lcnt, W.T.lsym)) err = 10;
					if (!err && !inf_build(W.lens + hlit, hdist, W.T.dcnt, W.T.dsym)) err = 10;
				}
			} else { // fixed code (RFC 1951 3.2.6)
				for (int s = 0; s < 288; ++s) W.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
				inf_build(W.lens, 288, W.T.lcnt, W.T.lsym);
				for (int s = 0; s < 30; ++s) W.lens[s] = 5;
				inf_build(W.lens, 30, W.T.dcnt, W.T.dsym);
			}
			if (win_over(b)) err = 15;
		}
		err = __shfl_sync(full, err, 0);
		if (err) return err;
		type = __shfl_sync(full, type, 0); last = __shfl_sync(full, last, 0);
		if (type == 0) {
			slen = __shfl_sync(full, slen, 0);
			soff = __shfl_sync(full, soff, 0);
			const uint8_t *src = in + h + soff;
			if (soff + slen > (long long)(in_len - 8 - h)) return 15;
			for (int i = lane; i < slen; i += 32) out[pos + i] = src[i];
			pos += slen;
			if (lane == 0) { // restart the word reader after the copied bytes
				const uint8_t *p = src + slen;
				b.buf = 0; b.n = 0; b.loaded_bits = 8ll * (soff + slen);
				while (((uintptr_t)p & 3) != 0) { b.buf |= (unsigned long long)*p++ << b.n; b.n += 8; b.loaded_bits += 8; }
				b.w = (const uint32_t *)p;
				b.pre = *b.w++;
				b.pre2 = *b.w++;
			}
			__syncwarp();
		} else {
			__syncwarp();
			infw_fill(W, 0, lane);
			infw_fill(W, 1, lane);
			for (;;) { // lane 0 runs to the next match (or the end of the block), the warp copies
				int ev = 0, len = 0, dist = 0; // ev: 1 match, 2 end of block, else an error code << 2
				if (lane == 0) {
					for (;;) {
						win_refill(b);
						const uint32_t e = W.lit[b.buf & ((1u << INF_LB) - 1)];
						int s;
						if (e) { s = (int)(e & 511u); b.buf >>= (e >> 9); b.n -= (int)(e >> 9); }
						else s = win_decode_slow(b, W.T.lcnt, W.T.lsym);
						if (s < 256) {
							if (s < 0) { ev = 11 << 2; break; }
							if (pos >= out_len) { ev = 3 << 2; break; }
							out[pos++] = (uint8_t)s;
						} else if (s == 256) { ev = 2; break; }
						else {
							if (s > 285) { ev = 12 << 2; break; }
							if (s < 265) len = s - 254;
							else if (s == 285) len = 258;
							else { const int x = (s - 261) >> 2; len = 3 + ((4 + ((s - 261) & 3)) << x) + (int)win_take(b, x); }
							win_refill(b);
							const uint32_t d = W.dist[b.buf & ((1u << INF_DB) - 1)];
							int ds;
							if (d) { ds = (int)(d & 511u); b.buf >>= (d >> 9); b.n -= (int)(d >> 9); }
							else ds = win_decode_slow(b, W.T.dcnt, W.T.dsym);
							if (ds < 0 || ds > 29) { ev = 12 << 2; break; }
							if (ds < 4) dist = ds + 1;
							else { const int x = (ds >> 1) - 1; dist = 1 + ((2 + (ds & 1)) << x) + (int)win_take(b, x); }
							if (dist > pos) { ev = 13 << 2; break; }
							if (pos + len > out_len) { ev = 3 << 2; break; }
							ev = 1;
							break;
						}
					}
					if (win_over(b)) ev = 15 << 2;
				}
				ev = __shfl_sync(full, ev, 0);
				pos = __shfl_sync(full, pos, 0);
				if (ev == 1) {
					len = __shfl_sync(full, len, 0); dist = __shfl_sync(full, dist, 0);
					__syncwarp(); // lane 0's literal stores are visible to the lanes that copy from them
					const uint8_t *src = out + pos - dist;
					if (dist >= len) { for (int i = lane; i < len; i += 32) out[pos + i] = src[i]; }
					else if (dist >= 32) { // each step of 32 bytes reads only bytes of earlier steps
						for (int i0 = 0; i0 < len; i0 += 32) { const int i = i0 + lane; if (i < len) out[pos + i] = src[i]; __syncwarp(); }
					} else { for (int i = lane; i < len; i += 32) out[pos + i] = src[i % dist]; } // a run: period dist, all of it already written
					pos += len;
					__syncwarp();
				} else if (ev == 2) break;
				else return ev >> 2;
			}
		}
		if (last) break;
	}
	if (pos != out_len) return 14;
	const uint8_t *t = in + in_len - 4;
	const uint32_t isize = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
	return isize == (uint32_t)out_len ? 0 : 14;
}

} // namespace bgzf
