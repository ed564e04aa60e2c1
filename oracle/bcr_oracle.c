/*
 * bcr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of ropebwt2's batched multi-string
 * insertion (mr_insert_multi, reference mrope.c:258-345 and mr_insert_multi_aux,
 * mrope.c:184-233).  It exists only so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check the CUDA engine; the product path under
 * ropebwt2_b200/ never links, imports or executes it.
 *
 * It is deliberately NOT structured like the reference: there is no rope, no
 * B+-tree and no run-length codec.  Each of the six buckets is a flat array with
 * one byte per BWT symbol, and one BCR column is restated as a batch operation in
 * "pre-column" coordinates:
 *
 *   entries of a bucket arrive sorted by position; a group is a maximal run of
 *   entries with equal `u` (mrope.c:191-192).  With h = index of the group's first
 *   entry inside the bucket segment, the group's interval in the bucket as it was
 *   BEFORE this column is  L = l - h, U = u - h.  For a member whose next symbol
 *   is a:
 *       l' = occ(a, L) + ins(a) + AC[a]          (return value of rope_insert_run,
 *       u' = l' + occ(a, U) - occ(a, L)           rope.c:114-148, plus mrope.c:332-340)
 *   where ins(a) counts the a-symbols inserted by earlier groups of this column and
 *   AC[a] the post-column count of a in the buckets in front of this one.  The
 *   group's new symbols go in at  L + sum_{a' before a}(occ(a',U)-occ(a',L))  in the
 *   order $,A,C,G,T,N (RLO / input order) or $,T,G,C,A,N (RCLO), mrope.c:206-224.
 *
 * Parity pinning: tests/test_oracle.py checks this file against (1) the unmodified
 * reference compiled into oracle/_ref (libref.so and the ropebwt2 binary), (2) the
 * golden fixtures in tests/golden/ generated from that binary by
 * oracle/gen_golden.py, (3) the README identities (README.md:18-25) and (4) a naive
 * suffix-sort definition (oracle/naive_bwt.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>

#define ORC_SO_IO   0   /* mrope.h:6-8 */
#define ORC_SO_RLO  1
#define ORC_SO_RCLO 2

typedef struct {
	uint8_t *sym;       /* one nt6 code (0..5) per BWT symbol */
	int64_t n, cap;
	int64_t tot[6];     /* marginal counts, the analogue of rope_t::c (rope.h:19) */
} orc_bucket_t;

typedef struct {
	int so;
	orc_bucket_t b[6];  /* bucket i holds symbols whose FOLLOWING symbol is i (README.md:78-80) */
} orc_t;

typedef struct {        /* the analogue of triple64_t, mrope.c:174-178 */
	int64_t l, u;
	int c;
	const uint8_t *p;
} orc_ent_t;

orc_t *orc_init(int so)
{
	orc_t *o;
	assert(so >= 0 && so <= 2); /* mrope.c:18 */
	o = (orc_t*)calloc(1, sizeof(orc_t));
	o->so = so;
	return o;
}

void orc_destroy(orc_t *o)
{
	int i;
	if (o == 0) return;
	for (i = 0; i < 6; ++i) free(o->b[i].sym);
	free(o);
}

static void bucket_reserve(orc_bucket_t *b, int64_t n)
{
	if (n > b->cap) {
		b->cap = n + (n >> 1) + 16;
		b->sym = (uint8_t*)realloc(b->sym, b->cap);
	}
}

/*
 * One column for one bucket: restates mr_insert_multi_aux (mrope.c:184-233).
 * `e[0..m)` are this bucket's entries, already sorted by position.  On return
 * e[k].l/u hold the new interval WITHOUT the cross-bucket offset (the caller adds
 * it, as mrope.c:332-340 does) and e[k].c the symbol just inserted.
 */
static void orc_column(orc_t *o, orc_bucket_t *bk, int64_t m, orc_ent_t *e)
{
	static const int ord_fwd[6] = { 0, 1, 2, 3, 4, 5 };   /* mrope.c:210 */
	static const int ord_cmp[6] = { 0, 4, 3, 2, 1, 5 };   /* mrope.c:209 */
	const int *ord = o->so == ORC_SO_RCLO? ord_cmp : ord_fwd;
	uint8_t *old = bk->sym, *out;
	int64_t n_old = bk->n, x = 0, w = 0, k, beg, a;
	int64_t run[6] = { 0, 0, 0, 0, 0, 0 }; /* occ(., x) while sweeping the old bucket */
	int64_t ins[6] = { 0, 0, 0, 0, 0, 0 }; /* symbols inserted so far in this column */

	if (m == 0) return;
	for (k = 0; k < m; ++k) e[k].c = *e[k].p++;               /* mrope.c:189-190 */
	out = (uint8_t*)malloc(n_old + m + 1);
	for (k = 1, beg = 0; k <= m; ++k) {
		int64_t L, U, occL[6], occU[6], cnt[6], at[6], i;
		int s;
		if (k != m && e[k].u == e[k-1].u) continue;             /* mrope.c:192 */
		L = e[beg].l - beg; U = e[beg].u - beg;                 /* pre-column coordinates */
		assert(L >= x && U >= L && U <= n_old);
		while (x < L) { ++run[old[x]]; out[w++] = old[x++]; }   /* advance the sweep to L */
		memcpy(occL, run, sizeof(run));
		memcpy(occU, run, sizeof(run));
		for (i = L; i < U; ++i) ++occU[old[i]];                 /* rope_rank2a, mrope.c:202 */
		memset(cnt, 0, sizeof(cnt));
		for (i = beg; i < k; ++i) ++cnt[e[i].c];                /* mrope.c:203-204 */
		/* physical insertion in $,A,C,G,T,N or $,T,G,C,A,N order (mrope.c:206-224): new
		 * symbols of slot s go in front of the old symbols of slot s inside [L,U) */
		for (s = 0; s < 6; ++s) {
			int64_t stop;
			a = ord[s];
			at[a] = run[a];  /* rank(a, insertion point) before inserting: rope.c:115,147 */
			for (i = 0; i < cnt[a]; ++i) out[w++] = (uint8_t)a;
			stop = x + (occU[a] - occL[a]);
			while (x < stop) { ++run[old[x]]; out[w++] = old[x++]; }
		}
		for (i = beg; i < k; ++i) {                              /* mrope.c:226-229 */
			a = e[i].c;
			e[i].l = at[a] + ins[a];
			e[i].u = e[i].l + (occU[a] - occL[a]);
		}
		for (a = 0; a < 6; ++a) ins[a] += cnt[a];
		beg = k;
	}
	while (x < n_old) out[w++] = old[x++];
	for (a = 0; a < 6; ++a) bk->tot[a] += ins[a];
	free(old);
	bk->sym = out; bk->n = w; bk->cap = n_old + m + 1;
}

/*
 * Restates the driver loop of mr_insert_multi (mrope.c:258-345): `s` is `len` bytes
 * of nt6 codes, each string reversed and NUL-terminated (mrope.c:268).
 */
void orc_insert_multi(orc_t *o, int64_t len, const uint8_t *s)
{
	int64_t m = 0, k, n0 = 0, n_live, i;
	orc_ent_t *cur, *nxt;
	const uint8_t *q;
	int b, is_srt = (o->so != ORC_SO_IO);

	assert(len > 0 && s[len-1] == 0);
	for (i = 0; i < len; ++i) m += (s[i] == 0);                /* mrope.c:271-272 */
	cur = (orc_ent_t*)malloc(m * sizeof(orc_ent_t));
	nxt = (orc_ent_t*)malloc(m * sizeof(orc_ent_t));
	for (i = 0, k = 0, q = s; i < len; ++i)                     /* mrope.c:275-276 */
		if (s[i] == 0) cur[k++].p = q, q = s + i + 1;
	for (b = 0; b < 6; ++b) n0 += o->b[b].tot[0];               /* mrope.c:279 */
	for (k = 0; k < m; ++k) {                                   /* mrope.c:280-284 */
		if (is_srt) cur[k].l = 0, cur[k].u = n0;
		else cur[k].l = cur[k].u = n0 + k;
		cur[k].c = 0;
	}
	orc_column(o, &o->b[0], m, cur);                            /* mrope.c:285 */
	n_live = m;
	while (n_live) {
		int64_t c[6], off[6], ac[6];
		memset(c, 0, sizeof(c));
		for (k = 0; k < n_live; ++k) ++c[cur[k].c];             /* mrope.c:303-309 */
		for (b = 1, off[0] = 0; b < 6; ++b) off[b] = off[b-1] + c[b-1];
		for (k = 0; k < n_live; ++k) nxt[off[cur[k].c]++] = cur[k];
		/* strings whose last symbol was the sentinel are done (mrope.c:310): keep only
		 * the part of the sorted array behind them */
		n_live -= c[0];
		memcpy(cur, nxt + c[0], n_live * sizeof(orc_ent_t));
		if (n_live == 0) break;
		for (b = 1, i = 0; b < 6; ++b) {                          /* mrope.c:327-328 */
			orc_column(o, &o->b[b], c[b], cur + i);
			i += c[b];
		}
		memset(ac, 0, sizeof(ac));
		for (b = 1, i = 0; b < 6; ++b) {                          /* mrope.c:332-340 */
			int a;
			for (a = 0; a < 6; ++a) ac[a] += o->b[b-1].tot[a];
			for (k = 0; k < c[b]; ++k, ++i)
				cur[i].l += ac[cur[i].c], cur[i].u += ac[cur[i].c];
		}
	}
	free(cur); free(nxt);
}

/* total number of BWT symbols (mr_get_tot, mrope.h:108-116) */
int64_t orc_total(const orc_t *o)
{
	int b;
	int64_t t = 0;
	for (b = 0; b < 6; ++b) t += o->b[b].n;
	return t;
}

/* per-bucket marginal counts, c[b*6+a] = rope b's c[a] (rope.h:19) */
void orc_counts(const orc_t *o, int64_t c[36])
{
	int b;
	for (b = 0; b < 6; ++b) memcpy(c + b * 6, o->b[b].tot, 48);
}

/* the BWT as nt6 codes, buckets 0..5 concatenated: what main.c:308-313 prints */
void orc_text(const orc_t *o, uint8_t *out)
{
	int b;
	for (b = 0; b < 6; ++b) {
		memcpy(out, o->b[b].sym, o->b[b].n);
		out += o->b[b].n;
	}
}

/* whole-index rank: cx[a] = #a in BWT[0,x) -- mr_rank1a semantics (mrope.c:70-105) */
void orc_rank1a(const orc_t *o, int64_t x, int64_t cx[6])
{
	int b;
	int64_t i;
	memset(cx, 0, 48);
	for (b = 0; b < 6 && x > 0; ++b) {
		int64_t n = o->b[b].n < x? o->b[b].n : x;
		for (i = 0; i < n; ++i) ++cx[o->b[b].sym[i]];
		x -= n;
	}
}
