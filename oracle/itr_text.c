/*
 * itr_text.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Walks a multi-rope block iterator (mr_itr_first / mr_itr_next_block, reference
 * mrope.c:111-130) through function pointers and decodes every leaf block
 * `[uint16 nbytes][runs...]` (reference rle.h:36-51) the way main.c:290-313 does.
 * Because it only uses the two iterator entry points it works unchanged on the
 * reference (oracle/_ref/libref.so) and on libropebwt2_b200.so, which is what makes
 * the parity tests symmetric.
 */
#include <stdint.h>
#include <string.h>

typedef void (*itr_first_fn)(void *mr, void *itr, int to_free);
typedef const uint8_t *(*itr_next_fn)(void *itr);

/* one run of the "43+3" byte codec (rle.h:39-51), restated as a function */
static inline const uint8_t *dec_run(const uint8_t *p, int *c, int64_t *l)
{
	uint8_t h = *p;
	*c = h & 7;
	if (h < 0x80) { *l = h >> 3; return p + 1; }
	if ((h >> 5) == 6) { *l = ((int64_t)(h & 0x18) << 3) | (p[1] & 0x3f); return p + 2; }
	{
		int i, n = (h & 0x10)? 8 : 4;
		int64_t x = (h >> 3) & 1;
		for (i = 1; i < n; ++i) x = x << 6 | (p[i] & 0x3f);
		*l = x;
		return p + n;
	}
}

/*
 * Decode the whole index into `out` (one nt6 code per symbol when ascii==0, or the
 * characters "$ACGTN" when ascii!=0).  Returns the number of symbols, or -(needed)
 * if `cap` is too small (nothing past cap is written).  n_blocks/n_runs are
 * optional statistics.
 */
int64_t itr_text(void *mr, itr_first_fn first, itr_next_fn next, int to_free,
                 uint8_t *out, int64_t cap, int ascii, int64_t *n_blocks, int64_t *n_runs)
{
	uint64_t itr_space[512]; /* mritr_t is ~1 KB (mrope.h:16-20, rope.h:24-29); 4 KB is plenty */
	const uint8_t *blk;
	int64_t n = 0, nb = 0, nr = 0;
	memset(itr_space, 0, sizeof(itr_space));
	first(mr, itr_space, to_free);
	while ((blk = next(itr_space)) != 0) {
		const uint8_t *q = blk + 2, *end = blk + 2 + *(const uint16_t*)blk;
		++nb;
		while (q < end) {
			int c; int64_t l, j;
			q = dec_run(q, &c, &l);
			++nr;
			if (n + l <= cap) {
				uint8_t v = ascii? (uint8_t)"$ACGTN"[c] : (uint8_t)c;
				for (j = 0; j < l; ++j) out[n + j] = v;
			}
			n += l;
		}
	}
	if (n_blocks) *n_blocks = nb;
	if (n_runs) *n_runs = nr;
	return n <= cap? n : -n;
}

/*
 * Same walk, but emit maximal runs (adjacent equal symbols merged, as rld_enc does,
 * rld0.c:153-161): syms[i], lens[i].  Returns the number of merged runs or -(needed).
 */
int64_t itr_runs(void *mr, itr_first_fn first, itr_next_fn next, int to_free,
                 uint8_t *syms, int64_t *lens, int64_t cap)
{
	uint64_t itr_space[512];
	const uint8_t *blk;
	int64_t n = 0, cur_l = 0;
	int cur_c = -1;
	memset(itr_space, 0, sizeof(itr_space));
	first(mr, itr_space, to_free);
	while ((blk = next(itr_space)) != 0) {
		const uint8_t *q = blk + 2, *end = blk + 2 + *(const uint16_t*)blk;
		while (q < end) {
			int c; int64_t l;
			q = dec_run(q, &c, &l);
			if (l == 0) continue;
			if (c == cur_c) { cur_l += l; continue; }
			if (cur_c >= 0) { if (n < cap) syms[n] = (uint8_t)cur_c, lens[n] = cur_l; ++n; }
			cur_c = c; cur_l = l;
		}
	}
	if (cur_c >= 0) { if (n < cap) syms[n] = (uint8_t)cur_c, lens[n] = cur_l; ++n; }
	return n <= cap? n : -n;
}

/*
 * Decode `n` leaf blocks laid out back to back (512 bytes each, [uint16 nbytes][runs...]) into one
 * nt6 code per symbol.  Used for indexes that are fetched block-wise (the sharded build, whose
 * sub-buckets live on different ranks).  Returns the number of symbols or -(needed).
 */
int64_t blocks_text(const uint8_t *blocks, int64_t n_blocks, uint8_t *out, int64_t cap)
{
	int64_t n = 0, b;
	for (b = 0; b < n_blocks; ++b) {
		const uint8_t *blk = blocks + b * 512;
		const uint8_t *q = blk + 2, *end = blk + 2 + *(const uint16_t*)blk;
		while (q < end) {
			int c; int64_t l, j;
			q = dec_run(q, &c, &l);
			if (n + l <= cap) for (j = 0; j < l; ++j) out[n + j] = (uint8_t)c;
			n += l;
		}
	}
	return n <= cap? n : -n;
}

/*
 * Streaming decode for indexes too large to hold as text (bench.py's md5 check of a 10 - 120 G-symbol
 * index against the reference's md5): open, then read the text chunk by chunk.
 */
#include <stdlib.h>
typedef struct { uint64_t itr_space[512]; itr_next_fn next; const uint8_t *q, *end; int c; int64_t rem; int done; } itr_stream_t;

void *itr_stream_open(void *mr, itr_first_fn first, itr_next_fn next, int to_free)
{
	itr_stream_t *s = (itr_stream_t*)calloc(1, sizeof(itr_stream_t));
	s->next = next;
	first(mr, s->itr_space, to_free);
	return s;
}

/* up to `cap` symbols into out (nt6 codes, or "$ACGTN" characters when ascii != 0); 0 at the end */
int64_t itr_stream_read(void *s_, uint8_t *out, int64_t cap, int ascii)
{
	itr_stream_t *s = (itr_stream_t*)s_;
	int64_t n = 0;
	while (n < cap && !s->done) {
		if (s->rem == 0) {
			if (s->q == s->end) {
				const uint8_t *blk = s->next(s->itr_space);
				if (blk == 0) { s->done = 1; break; }
				s->q = blk + 2; s->end = blk + 2 + *(const uint16_t*)blk;
				continue;
			}
			s->q = dec_run(s->q, &s->c, &s->rem);
			continue;
		}
		{
			int64_t l = s->rem < cap - n? s->rem : cap - n;
			memset(out + n, ascii? "$ACGTN"[s->c] : s->c, (size_t)l);
			n += l; s->rem -= l;
		}
	}
	return n;
}

void itr_stream_close(void *s) { free(s); }

/* blocks_text with a choice of output alphabet: nt6 codes (ascii == 0) or the characters "$ACGTN" */
int64_t blocks_text2(const uint8_t *blocks, int64_t n_blocks, uint8_t *out, int64_t cap, int ascii)
{
	int64_t n = 0, b;
	for (b = 0; b < n_blocks; ++b) {
		const uint8_t *blk = blocks + b * 512;
		const uint8_t *q = blk + 2, *end = blk + 2 + *(const uint16_t*)blk;
		while (q < end) {
			int c; int64_t l;
			q = dec_run(q, &c, &l);
			if (n + l <= cap) memset(out + n, ascii? "$ACGTN"[c] : c, (size_t)l);
			n += l;
		}
	}
	return n <= cap? n : -n;
}
