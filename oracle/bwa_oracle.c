/* bwa_oracle.c -- TEST INFRASTRUCTURE ONLY (see bwa_oracle.h).
 *
 * CPU restatement of the reference's alignment hot path, written for clarity rather than
 * speed: where the reference uses SWAR popcounts, byte look-up tables and packed 16:16
 * integers, this file counts symbol by symbol and keeps H and E in separate arrays.  The
 * results are required to be IDENTICAL to the reference's (tests/test_oracle.py).
 */
#include "bwa_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- FM-index primitives */

/* Symbol p of the '$'-less BWT in the reference layout: block p/128 = 4 counts then 8
 * words, 16 symbols per word, first symbol in the top bits (bwt.h:61-68). */
static inline int bwt_sym(const orc_index_t *ix, uint32_t p)
{
	const uint32_t *blk = ix->bwt + (size_t)(p >> 7) * 12;
	uint32_t w = blk[4 + ((p & 127) >> 4)];
	return (w >> ((15 - (p & 15)) << 1)) & 3;
}

/* occ(k, c): number of c in BWT[0..k] where k indexes the BWT WITH its '$' row; the row is
 * not stored, so positions at or after `primary` shift down by one (bwt.c:99).  k == -1
 * gives 0 (bwt.c:98); k == seq_len is the total count (bwt.c:97). */
uint32_t orc_occ(const orc_index_t *ix, uint32_t k, int c)
{
	uint32_t n, p, q;
	if (k == ix->seq_len) return ix->L2[c + 1] - ix->L2[c];
	if (k == (uint32_t)-1) return 0;
	if (k >= ix->primary) --k;
	n = ix->bwt[(size_t)(k >> 7) * 12 + c]; /* checkpoint: count before this 128-block */
	for (p = k & ~127u, q = k; p <= q; ++p) n += bwt_sym(ix, p) == c;
	return n;
}

/* bwt.c:159-176.  Unlike orc_occ there is no k == seq_len shortcut in the reference; the
 * generic path gives the same totals. */
void orc_occ4(const orc_index_t *ix, uint32_t k, uint32_t cnt[4])
{
	uint32_t p;
	if (k == (uint32_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	if (k >= ix->primary) --k;
	memcpy(cnt, ix->bwt + (size_t)(k >> 7) * 12, 16);
	for (p = k & ~127u; p <= k; ++p) ++cnt[bwt_sym(ix, p)];
}

/* bwt_2occ / bwt_2occ4 (bwt.c:118-153, 179-216) only share work between two lookups; their
 * values are those of two independent lookups. */
static void occ_pair(const orc_index_t *ix, uint32_t k, uint32_t l, int c, uint32_t *ok, uint32_t *ol)
{
	*ok = orc_occ(ix, k, c);
	*ol = orc_occ(ix, l, c);
}

/* inverse Psi (bwt.h:71-75) and the sampled-SA walk (bwt.c:72-81) */
static uint32_t inv_psi(const orc_index_t *ix, uint32_t k)
{
	int c;
	if (k == ix->primary) return 0;
	c = bwt_sym(ix, k < ix->primary ? k : k - 1);
	return ix->L2[c] + orc_occ(ix, k, c);
}

uint32_t orc_sa(const orc_index_t *ix, uint32_t k)
{
	uint32_t steps = 0;
	while (k % (uint32_t)ix->sa_intv != 0) { ++steps; k = inv_psi(ix, k); }
	k /= (uint32_t)ix->sa_intv;
	return steps + (k ? ix->sa[k] : (uint32_t)-1);
}

/* ---------------------------------------------------------------- per-read budget */

/* bwtaln.c:37-49: smallest k such that P[Poisson(l*err) > k] < thres.  `x` is an int in the
 * reference too (it overflows for k > 12 exactly the same way). */
int orc_cal_maxdiff(int l, double err, double thres)
{
	double elambda = exp(-l * err), sum = elambda, y = 1.0;
	int k, x = 1;
	for (k = 1; k < 1000; ++k) {
		y *= l * err;
		x *= k;
		sum += elambda * y / x;
		if (1.0 - sum < thres) return k;
	}
	return 2;
}

/* bwtaln.c:52-76: width[i].bid = lower bound on the differences needed to align str[0..i],
 * from exact backward search that restarts at every miss (or N). */
int orc_cal_width(const orc_index_t *ix, int len, const uint8_t *str, orc_width_t *width)
{
	uint32_t k = 0, l = ix->seq_len, ok, ol;
	int i, bid = 0;
	for (i = 0; i < len; ++i) {
		int c = str[i];
		if (c < 4) {
			occ_pair(ix, k - 1, l, c, &ok, &ol);
			k = ix->L2[c] + ok + 1;
			l = ix->L2[c] + ol;
		}
		if (k > l || c > 3) { k = 0; l = ix->seq_len; ++bid; }
		width[i].w = l - k + 1;
		width[i].bid = bid;
	}
	width[len].w = 0;
	width[len].bid = ++bid;
	return bid;
}

/* ---------------------------------------------------------------- bounded best-first search */

typedef struct {
	uint32_t k, l;
	int i, a, n_mm, n_gapo, n_gape, state, last_diff_pos;
} node_t;

typedef struct { node_t *v; int n, m; } bucket_t;

typedef struct {
	bucket_t *b;
	int n_buckets, best, n_entries;
} heap_t; /* gap_stack_t: one LIFO per score, lowest non-empty score first (bwtgap.c:13-79) */

static int score_of(const orc_opt_t *o, int mm, int go, int ge) { return mm * o->s_mm + go * o->s_gapo + ge * o->s_gape; }

static void heap_push(heap_t *h, const orc_opt_t *o, int a, int i, uint32_t k, uint32_t l, int mm, int go, int ge,
                      int state, int is_diff)
{
	int s = score_of(o, mm, go, ge);
	bucket_t *q = h->b + s;
	node_t *p;
	if (q->n == q->m) { q->m = q->m ? q->m << 1 : 4; q->v = (node_t *)realloc(q->v, sizeof(node_t) * q->m); }
	p = q->v + q->n++;
	p->k = k; p->l = l; p->i = i; p->a = a; p->n_mm = mm; p->n_gapo = go; p->n_gape = ge; p->state = state;
	p->last_diff_pos = is_diff ? i : 0;
	++h->n_entries;
	if (s < h->best) h->best = s;
}

static node_t heap_pop(heap_t *h)
{
	bucket_t *q = h->b + h->best;
	node_t e = q->v[--q->n];
	--h->n_entries;
	if (h->n_entries == 0) h->best = h->n_buckets;
	else while (h->b[h->best].n == 0) ++h->best;
	return e;
}

/* bwt_match_exact_alt (bwt.c:237-252) */
static int match_exact(const orc_index_t *ix, int len, const uint8_t *str, uint32_t *k0, uint32_t *l0)
{
	uint32_t k = *k0, l = *l0, ok, ol;
	int i;
	for (i = len - 1; i >= 0; --i) {
		int c = str[i];
		if (c > 3) return 0;
		occ_pair(ix, k - 1, l, c, &ok, &ol);
		k = ix->L2[c] + ok + 1;
		l = ix->L2[c] + ol;
		if (k > l) return 0;
	}
	*k0 = k; *l0 = l;
	return 1;
}

/* gap_shadow (bwtgap.c:81-91): after accepting a hit of x occurrences, take them out of the
 * width bounds left of the last difference. */
static void shadow(uint32_t x, uint32_t max, int last_diff_pos, orc_width_t *w)
{
	int i;
	uint32_t j = 0;
	for (i = 0; i < last_diff_pos; ++i) {
		if (w[i].w > x) w[i].w -= x;
		else if (w[i].w == x) { w[i].bid = 1; w[i].w = max - (++j); }
	}
}

static int ilog2(uint32_t v) { int c = 0; while (v >>= 1) ++c; return c; } /* int_log2, bwtgap.c:93-102 */

/* bwt_match_gap (bwtgap.c:104-266).  o->max_diff / o->max_gapo / o->seed_len are the
 * per-read local_opt values of bwa_cal_sa_reg_gap. */
static orc_aln_t *match_gap(const orc_index_t ix[2], int len, const uint8_t *seq[2], orc_width_t *w[2],
                            orc_width_t *seed_w[2], const orc_opt_t *o, int *n_aln_out, int *max_entries_out)
{
	const int gape = o->mode & 0x01, loggap = o->mode & 0x04, nonstop = o->mode & 0x10;
	int best_score = score_of(o, o->max_diff + 1, o->max_gapo + 1, o->max_gape + 1);
	int max_diff = o->max_diff, best_cnt = 0, max_entries = 0, n_aln = 0, m_aln = 4, j, n_amb = 0;
	orc_aln_t *aln = (orc_aln_t *)calloc(m_aln, sizeof(orc_aln_t));
	heap_t h;

	for (j = 0; j < len; ++j) n_amb += seq[0][j] > 3;
	if (n_amb > max_diff) { *n_aln_out = 0; return aln; } /* bwtgap.c:118-123 */

	h.n_buckets = best_score; /* gap_init_stack, bwtgap.c:18 */
	h.b = (bucket_t *)calloc(h.n_buckets, sizeof(bucket_t));
	h.best = h.n_buckets; h.n_entries = 0;
	heap_push(&h, o, 0, len, 0, ix[0].seq_len, 0, 0, 0, 0, 0);
	heap_push(&h, o, 1, len, 0, ix[0].seq_len, 0, 0, 0, 0, 0);

	while (h.n_entries) {
		node_t e;
		const orc_index_t *bwt;
		const uint8_t *str;
		orc_width_t *width;
		uint32_t k, l, ck[4], cl[4], occ;
		int a, i, m, m_seed = 0, hit = 0, allow_diff = 1, allow_M = 1, tmp, c;

		if (max_entries < h.n_entries) max_entries = h.n_entries;
		if (h.n_entries > o->max_entries) break;
		e = heap_pop(&h);
		k = e.k; l = e.l; a = e.a; i = e.i;
		if (!nonstop && score_of(o, e.n_mm, e.n_gapo, e.n_gape) > best_score + o->s_mm) break;
		m = max_diff - (e.n_mm + e.n_gapo) - (gape ? e.n_gape : 0);
		if (m < 0) continue;
		bwt = &ix[1 - a]; str = seq[a]; width = w[a]; /* strand a is searched in the OTHER index, bwtgap.c:149 */
		if (seed_w) m_seed = o->max_seed_diff - (e.n_mm + e.n_gapo) - (gape ? e.n_gape : 0);
		if (i > 0 && m < width[i - 1].bid) continue;

		if (i == 0) hit = 1;
		else if (m == 0 && (e.state == 0 || gape || e.n_gape == o->max_gape)) {
			if (match_exact(bwt, i, str, &k, &l)) hit = 1;
			else continue;
		}
		if (hit) { /* bwtgap.c:167-200 */
			int score = score_of(o, e.n_mm, e.n_gapo, e.n_gape), add = 1;
			if (n_aln == 0) {
				int best_diff = e.n_mm + e.n_gapo + (gape ? e.n_gape : 0);
				best_score = score;
				if (!nonstop) max_diff = best_diff + 1 > o->max_diff ? o->max_diff : best_diff + 1;
			}
			if (score == best_score) best_cnt += (int)(l - k + 1);
			else if (best_cnt > o->max_top2) break;
			if (e.n_gapo)
				for (j = 0; j < n_aln; ++j)
					if (aln[j].k == k && aln[j].l == l) { add = 0; break; }
			if (add) {
				shadow(l - k + 1, bwt->seq_len, e.last_diff_pos, width);
				if (n_aln == m_aln) {
					m_aln <<= 1;
					aln = (orc_aln_t *)realloc(aln, m_aln * sizeof(orc_aln_t));
					memset(aln + m_aln / 2, 0, m_aln / 2 * sizeof(orc_aln_t));
				}
				aln[n_aln].info = (uint32_t)e.n_mm | (uint32_t)e.n_gapo << 8 | (uint32_t)e.n_gape << 16 | (uint32_t)a << 24;
				aln[n_aln].k = k; aln[n_aln].l = l; aln[n_aln].score = score;
				++n_aln;
			}
			continue;
		}

		--i;
		orc_occ4(bwt, k - 1, ck);
		orc_occ4(bwt, l, cl);
		occ = l - k + 1;
		if (i > 0) { /* may a difference / a mismatch still pay off here?  bwtgap.c:206-216 */
			int ii = i - (len - o->seed_len);
			if (width[i - 1].bid > m - 1) allow_diff = 0;
			else if (width[i - 1].bid == m - 1 && width[i].bid == m - 1 && width[i - 1].w == width[i].w) allow_M = 0;
			if (seed_w && ii > 0) {
				const orc_width_t *sw = seed_w[a];
				if (sw[ii - 1].bid > m_seed - 1) allow_diff = 0;
				else if (sw[ii - 1].bid == m_seed - 1 && sw[ii].bid == m_seed - 1 && sw[ii - 1].w == sw[ii].w) allow_M = 0;
			}
		}
		tmp = loggap ? ilog2((uint32_t)(e.n_gape + e.n_gapo)) / 2 + 1 : e.n_gapo + e.n_gape;
		if (allow_diff && i >= o->indel_end_skip + tmp && len - i >= o->indel_end_skip + tmp) {
			if (e.state == 0) {
				if (e.n_gapo < o->max_gapo) {
					heap_push(&h, o, a, i, k, l, e.n_mm, e.n_gapo + 1, e.n_gape, 1, 1);
					for (c = 0; c < 4; ++c) {
						uint32_t nk = bwt->L2[c] + ck[c] + 1, nl = bwt->L2[c] + cl[c];
						if (nk <= nl) heap_push(&h, o, a, i + 1, nk, nl, e.n_mm, e.n_gapo + 1, e.n_gape, 2, 1);
					}
				}
			} else if (e.state == 1) {
				if (e.n_gape < o->max_gape) heap_push(&h, o, a, i, k, l, e.n_mm, e.n_gapo, e.n_gape + 1, 1, 1);
			} else if (e.n_gape < o->max_gape && (e.n_gape + e.n_gapo < max_diff || occ < (uint32_t)o->max_del_occ)) {
				for (c = 0; c < 4; ++c) {
					uint32_t nk = bwt->L2[c] + ck[c] + 1, nl = bwt->L2[c] + cl[c];
					if (nk <= nl) heap_push(&h, o, a, i + 1, nk, nl, e.n_mm, e.n_gapo, e.n_gape + 1, 2, 1);
				}
			}
		}
		if (allow_diff && allow_M) {
			for (j = 1; j <= 4; ++j) {
				int is_mm = j != 4 || str[i] > 3;
				uint32_t nk, nl;
				c = (str[i] + j) & 3;
				nk = bwt->L2[c] + ck[c] + 1; nl = bwt->L2[c] + cl[c];
				if (nk <= nl) heap_push(&h, o, a, i, nk, nl, e.n_mm + is_mm, e.n_gapo, e.n_gape, 0, is_mm);
			}
		} else if (str[i] < 4) {
			uint32_t nk, nl;
			c = str[i];
			nk = bwt->L2[c] + ck[c] + 1; nl = bwt->L2[c] + cl[c];
			if (nk <= nl) heap_push(&h, o, a, i, nk, nl, e.n_mm, e.n_gapo, e.n_gape, 0, 0);
		}
	}
	for (j = 0; j < h.n_buckets; ++j) free(h.b[j].v);
	free(h.b);
	*n_aln_out = n_aln;
	*max_entries_out = max_entries;
	return aln;
}

int orc_cal_sa_reg_gap1(const orc_index_t ix[2], int len, const uint8_t *seq, const uint8_t *rseq,
                        const orc_opt_t *opt, orc_aln_t **aln, int *max_entries)
{
	orc_opt_t lo = *opt;
	orc_width_t *w[2], *sw[2];
	const uint8_t *s[2];
	int n_aln = 0, use_seed;
	*aln = 0;
	if (len <= 0) return 0; /* bwtaln.c:134 */
	/* n_seqs = 1: the stack-sizing max_diff, the max_gapo clamp and the search budget all
	 * come from this read's own length (bwtaln.c:100-103,126) */
	if (opt->fnr > 0.0) lo.max_diff = orc_cal_maxdiff(len, 0.02, opt->fnr);
	if (lo.max_diff < lo.max_gapo) lo.max_gapo = lo.max_diff;
	lo.seed_len = opt->seed_len < len ? opt->seed_len : INT_MAX;
	use_seed = len > opt->seed_len;
	s[0] = seq; s[1] = rseq;
	w[0] = (orc_width_t *)calloc(len + 1, sizeof(orc_width_t));
	w[1] = (orc_width_t *)calloc(len + 1, sizeof(orc_width_t));
	sw[0] = (orc_width_t *)calloc((use_seed ? opt->seed_len : 0) + 1, sizeof(orc_width_t));
	sw[1] = (orc_width_t *)calloc((use_seed ? opt->seed_len : 0) + 1, sizeof(orc_width_t));
	orc_cal_width(&ix[0], len, s[0], w[0]);
	orc_cal_width(&ix[1], len, s[1], w[1]);
	if (use_seed) {
		orc_cal_width(&ix[0], opt->seed_len, s[0] + (len - opt->seed_len), sw[0]);
		orc_cal_width(&ix[1], opt->seed_len, s[1] + (len - opt->seed_len), sw[1]);
	}
	*aln = match_gap(ix, len, s, w, use_seed ? sw : 0, &lo, &n_aln, max_entries);
	free(w[0]); free(w[1]); free(sw[0]); free(sw[1]);
	return n_aln;
}

int orc_aln_flat(const orc_index_t ix[2], int n, const uint8_t *bases, const int64_t *offs, const orc_opt_t *opt,
                 int32_t *n_aln, int32_t *max_entries, orc_aln_t **pool_out)
{
	size_t cap = (size_t)n * 2 + 16, used = 0;
	orc_aln_t *pool = (orc_aln_t *)malloc(cap * sizeof(orc_aln_t));
	int i, j;
	for (i = 0; i < n; ++i) {
		int len = (int)(offs[i + 1] - offs[i]), me = 0, na;
		const uint8_t *r = bases + offs[i];
		uint8_t *seq = (uint8_t *)malloc(len + 1), *rseq = (uint8_t *)malloc(len + 1);
		orc_aln_t *aln = 0;
		for (j = 0; j < len; ++j) { /* bam1_to_seq, bwaseqio.c:294-297 */
			uint8_t b = r[len - 1 - j];
			seq[j] = b > 3 ? 4 : b;
			rseq[j] = b > 3 ? 4 : 3 - b;
		}
		na = orc_cal_sa_reg_gap1(ix, len, seq, rseq, opt, &aln, &me);
		if (used + (size_t)na > cap) { cap = (used + na) * 2; pool = (orc_aln_t *)realloc(pool, cap * sizeof(orc_aln_t)); }
		if (na) memcpy(pool + used, aln, (size_t)na * sizeof(orc_aln_t));
		used += (size_t)na;
		n_aln[i] = na; max_entries[i] = me;
		free(aln); free(seq); free(rseq);
	}
	*pool_out = pool;
	return 0;
}

void orc_free(void *p) { free(p); }

/* ---------------------------------------------------------------- mate-rescue Smith-Waterman */

/* aln_sm_maq (stdaln.c:206-212) and aln_param_bwa = {26, 9, ...} (stdaln.c:227) */
static int sw_score(int a, int b) { return (a > 3 || b > 3) ? -13 : (a == b ? 11 : -19); }
enum { SW_Q = 26, SW_R = 9, SW_QR = 35, SW_MAXSC = 11 };

/* aln_local_core passes 1 and 2 (stdaln.c:578-696).  i runs over ref (1..len1), j over the
 * query (1..len2).  The reference packs H (high 16 bits) and E (low 16 bits) of a column in
 * one int and recycles one array; here H and E are separate arrays with the same update
 * order, which makes two of its rules visible:
 *   forward:  E(i,j) is carried only while H(i,j-1) >= q+r+1 (stdaln.c:615); otherwise 0;
 *             F is refreshed only while the cell to the left is positive (stdaln.c:611);
 *   the maximum is the FIRST cell in (j, then i) order that strictly exceeds the running
 *             maximum (stdaln.c:623-625).
 * The overflow rescaling (stdaln.c:587-606, 654-667) cannot trigger below 32000, i.e. for
 * queries shorter than ~2900 bases; it is not restated and -2 is returned if it would. */
int orc_sw_local(const uint8_t *ref, int len1, const uint8_t *query, int len2, int res[4])
{
	int *H, *E; /* H[i] = H(i, j-1) while row j is being computed, updated in place; E likewise */
	int i, j, score_f = 0, end_i = 0, end_j = 0, start_i = 0, start_j = 0;
	res[0] = res[1] = res[2] = res[3] = 0;
	if (len1 == 0 || len2 == 0) return -1;
	H = (int *)calloc(len1 + 2, sizeof(int));
	E = (int *)calloc(len1 + 2, sizeof(int));
	for (j = 1; j <= len2; ++j) {
		int last_h = 0, f = 0, diag = 0; /* diag = H(i-1, j-1); H(0, *) = 0 */
		for (i = 1; i <= len1; ++i) {
			int up = H[i]; /* H(i, j-1) */
			int h = diag + sw_score(ref[i - 1], query[j - 1]), e = 0;
			if (h < 0) h = 0;
			if (last_h > 0) {
				f = f > last_h - SW_Q ? f - SW_R : last_h - SW_QR;
				if (h < f) h = f;
			}
			if (up >= SW_QR + 1) {
				e = E[i] > up - SW_Q ? E[i] - SW_R : up - SW_QR;
				if (h < e) h = e;
			}
			E[i] = e;
			diag = up;
			H[i] = h;
			last_h = h;
			if (score_f < h) { score_f = h; end_i = i; end_j = j; }
		}
		if (score_f > 32000) { free(H); free(E); return -2; }
	}
	res[2] = end_i; res[3] = end_j;
	if (score_f < 1 || end_i == 0 || end_j == 0) { free(H); free(E); return score_f; }

	/* reverse pass from the end cell in an adaptive band (stdaln.c:638-696).  Array cell x of
	 * the reference's eh[] holds H in its high half and E in its low half; h[x], e[x] here. */
	{
		int *h = H, *e = E, score_r, start, end;
		for (i = 0; i <= end_i; ++i) h[i] = e[i] = 0;
		score_r = sw_score(ref[end_i - 1], query[end_j - 1]);
		start_i = end_i; start_j = end_j;
		h[end_i] = SW_QR + score_r; e[end_i] = 0;
		start = end_i - 1;
		end = end_i - 3;
		if (end <= 0) end = 0;
		for (j = end_j - 1; j != 0; --j) {
			int last_h = 0, f = 0, x, stop = 0;
			if (start < end) { free(H); free(E); return -3; } /* never observed; the reference would run off its array */
			for (i = start, x = start + 1; i != end; --i, --x) {
				int cur = h[x] + sw_score(ref[i - 1], query[j - 1]), ee, left;
				if (cur < 0) cur = 0;
				if (last_h > 0) {
					f = f > last_h - SW_Q ? f - SW_R : last_h - SW_QR;
					if (cur < f) cur = f;
				}
				left = h[x - 1];
				ee = e[x] > left - SW_Q ? e[x] - SW_R : left - SW_QR;
				if (ee < 0) ee = 0;
				if (cur < ee) cur = ee;
				h[x] = last_h; e[x] = ee;
				last_h = cur;
				if (score_r < cur) {
					score_r = cur; start_i = i; start_j = j;
					if (score_r - SW_QR == score_f) { stop = 1; break; }
				}
			}
			h[x] = last_h; e[x] = 0;
			if (stop) break;
			if (h[start] <= SW_QR) --start;
			if (start <= 0) start = 0;
			end = start_i - (start_j - j) - (score_r + (start_j - j) * SW_MAXSC) / SW_R - 1;
			if (end <= 0) end = 0;
		}
	}
	res[0] = start_i; res[1] = start_j;
	free(H); free(E);
	return score_f;
}

/* ---------------------------------------------------------------- banded global alignment */

#define G_INF (-1073741823) /* MINOR_INF, stdaln.h:84 */
enum { G_M = 0, G_I = 1, G_D = 2 };
typedef struct { int M, I, D; } gscore_t;            /* dpscore_t, stdaln.c:326-329 */
typedef struct { unsigned char Mt, It, Dt; } gcell_t; /* dpcell_t, stdaln.c:321-324 */

/* the set_* macros of stdaln.c:260-319 as functions; `p` is the cell the value comes from */
static int g_setM(gcell_t *c, const gscore_t *p, int sc)
{
	if (p->M >= p->I) {
		if (p->M >= p->D) { c->Mt = G_M; return p->M + sc; }
		c->Mt = G_D; return p->D + sc;
	}
	if (p->I > p->D) { c->Mt = G_I; return p->I + sc; }
	c->Mt = G_D; return p->D + sc;
}
static int g_setI(gcell_t *c, const gscore_t *p, int ext)
{
	if (p->M - SW_Q > p->I) { c->It = G_M; return p->M - SW_Q - ext; }
	c->It = G_I; return p->I - ext;
}
static int g_setD(gcell_t *c, const gscore_t *p, int ext)
{
	if (p->M - SW_Q > p->D) { c->Dt = G_M; return p->M - SW_Q - ext; }
	c->Dt = G_D; return p->D - ext;
}

int orc_global(const uint8_t *ref, int len1, const uint8_t *query, int len2, int gap_end, int band,
               int32_t *path_ijc, int *path_len)
{
	const int eend = gap_end >= 0 ? gap_end : SW_R; /* set_end_* fall back to set_* when gap_end < 0 */
	int b1, b2, width, i, j, end, tmp_end, max, type, ctype, n;
	gcell_t *cells;  /* (len2+1) rows of `width` cells; row j > b2 is shifted left by j - b2 */
	gscore_t *curr, *last, *sw;
#define CELL(j, i) (cells + (size_t)(j) * width + ((j) > b2 ? (i) - ((j) - b2) : (i)))
#define SC(i) sw_score(ref[(i) - 1], query[j - 1])
	if (len1 == 0 || len2 == 0) { *path_len = 0; return 0; }
	if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; }
	else { b1 = band; b2 = len2 - len1 + band; }
	if (b1 > len1) b1 = len1;
	if (b2 > len2) b2 = len2;
	width = (b1 + b2 <= len1) ? b1 + b2 + 1 : len1 + 1;
	cells = (gcell_t *)calloc((size_t)(len2 + 1) * width + 8, sizeof(gcell_t));
	curr = (gscore_t *)calloc(len1 + 2, sizeof(gscore_t));
	last = (gscore_t *)calloc(len1 + 2, sizeof(gscore_t));

	/* first row (stdaln.c:393-399) */
	curr[0].M = 0; curr[0].I = curr[0].D = G_INF;
	for (i = 1; i < b1; ++i) {
		curr[i].M = curr[i].I = G_INF;
		curr[i].D = g_setD(CELL(0, i), &curr[i - 1], eend);
	}
	sw = curr; curr = last; last = sw;

	/* part 1: rows whose band still starts at column 0 (stdaln.c:401-440) */
	tmp_end = b2 < len2 ? b2 : len2 - 1;
	for (j = 1; j <= tmp_end + 1; ++j) {
		const int last_row = j == tmp_end + 1; /* the "last row for part 1" variant: end gaps in D */
		if (last_row && !(j == len2 && b2 != len2 - 1)) break;
		curr[0].M = curr[0].D = G_INF;
		curr[0].I = g_setI(CELL(j, 0), &last[0], eend);
		end = (j + b1 <= len1 + 1) ? j + b1 - 1 : len1;
		for (i = 1; i != end; ++i) {
			curr[i].M = g_setM(CELL(j, i), &last[i - 1], SC(i));
			curr[i].I = g_setI(CELL(j, i), &last[i], SW_R);
			curr[i].D = g_setD(CELL(j, i), &curr[i - 1], last_row ? eend : SW_R);
		}
		curr[i].M = g_setM(CELL(j, i), &last[i - 1], SC(i));
		curr[i].D = g_setD(CELL(j, i), &curr[i - 1], last_row ? eend : SW_R);
		if (j + b1 - 1 > len1) curr[i].I = g_setI(CELL(j, i), &last[i], eend);
		else curr[i].I = G_INF;
		sw = curr; curr = last; last = sw;
	}
	/* j now = first row not done by part 1 (the reference's ++j after its special row included) */

	/* part 2: band strictly inside (stdaln.c:442-456) */
	for (; j <= len2 - b2 + 1; ++j) {
		curr[j - b2].M = curr[j - b2].I = curr[j - b2].D = G_INF;
		end = j + b1 - 1;
		for (i = j - b2 + 1; i != end; ++i) {
			curr[i].M = g_setM(CELL(j, i), &last[i - 1], SC(i));
			curr[i].I = g_setI(CELL(j, i), &last[i], SW_R);
			curr[i].D = g_setD(CELL(j, i), &curr[i - 1], SW_R);
		}
		curr[i].M = g_setM(CELL(j, i), &last[i - 1], SC(i));
		curr[i].D = g_setD(CELL(j, i), &curr[i - 1], SW_R);
		curr[i].I = G_INF;
		sw = curr; curr = last; last = sw;
	}
	/* part 3: band reaches the last column (stdaln.c:458-471), then the last row (472-487) */
	for (; j <= len2; ++j) {
		const int last_row = j == len2;
		curr[j - b2].M = curr[j - b2].I = curr[j - b2].D = G_INF;
		for (i = j - b2 + 1; i < len1; ++i) {
			curr[i].M = g_setM(CELL(j, i), &last[i - 1], SC(i));
			curr[i].I = g_setI(CELL(j, i), &last[i], SW_R);
			curr[i].D = g_setD(CELL(j, i), &curr[i - 1], last_row ? eend : SW_R);
		}
		curr[i].M = g_setM(CELL(j, i), &last[len1 - 1], SC(i));
		curr[i].I = g_setI(CELL(j, i), &last[i], eend);
		curr[i].D = g_setD(CELL(j, i), &curr[i - 1], last_row ? eend : SW_R);
		sw = curr; curr = last; last = sw;
	}

	/* backtrace (stdaln.c:489-513) */
	i = len1; j = len2;
	max = last[len1].M; type = CELL(j, i)->Mt; ctype = G_M;
	if (last[len1].I > max) { max = last[len1].I; type = CELL(j, i)->It; ctype = G_I; }
	if (last[len1].D > max) { max = last[len1].D; type = CELL(j, i)->Dt; ctype = G_D; }
	n = 0;
	path_ijc[0] = i; path_ijc[1] = j; path_ijc[2] = ctype; ++n;
	do {
		const gcell_t *q;
		if (ctype == G_M) { --i; --j; } else if (ctype == G_I) --j; else --i;
		q = CELL(j, i);
		ctype = type;
		type = type == G_M ? q->Mt : type == G_I ? q->It : q->Dt;
		path_ijc[3 * n] = i; path_ijc[3 * n + 1] = j; path_ijc[3 * n + 2] = ctype; ++n;
	} while (i || j);
	*path_len = n - 1;
	free(cells); free(curr); free(last);
#undef CELL
#undef SC
	return max;
}

int orc_path2cigar(const int32_t *path_ijc, int path_len, uint16_t *cigar)
{
	int i, n = 0, last;
	if (path_len == 0) return 0;
	last = path_ijc[3 * (path_len - 1) + 2];
	cigar[0] = (uint16_t)(last << 14 | 1);
	for (i = path_len - 2; i >= 0; --i) {
		int t = path_ijc[3 * i + 2];
		if (t == last) ++cigar[n];
		else { cigar[++n] = (uint16_t)(t << 14 | 1); last = t; }
	}
	return n + 1;
}

int orc_sw_local_path(const uint8_t *ref, int len1, const uint8_t *query, int len2, int32_t *path_ijc, int *path_len)
{
	int res[4], score_f, score_g, band, span, t;
	*path_len = 0;
	score_f = orc_sw_local(ref, len1, query, len2, res);
	if (score_f < 1 || res[2] == 0 || res[3] == 0) return score_f; /* stdaln.c:633-640 */
	/* score_r - qr == score_f always held in the tests; the reference compares both (stdaln.c:732) */
	span = res[2] - res[0] > res[3] - res[1] ? res[2] - res[0] : res[3] - res[1];
	++span;
	for (band = 50;; band <<= 1) { /* aln_param_bwa.band_width = 50 */
		score_g = orc_global(ref + res[0] - 1, res[2] - res[0] + 1, query + res[1] - 1, res[3] - res[1] + 1, -1, band,
		                     path_ijc, path_len);
		if (score_g == score_f) break;
		if (band > span) break;
	}
	if (score_f > score_g) return -1; /* "Potential bug" branch, stdaln.c:736-739 */
	for (t = 0; t < *path_len; ++t) { path_ijc[3 * t] += res[0] - 1; path_ijc[3 * t + 1] += res[1] - 1; }
	return score_g;
}
