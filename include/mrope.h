/*
 * mrope.h -- multi-rope C API, source compatible with lh3/ropebwt2 (reference mrope.h).
 *
 * This is the drop-in boundary: the reference driver (main.c) includes this header and
 * links libropebwt2_b200.so instead of mrope.c / rope.c / rle.c.  Same names, argument
 * meaning and error behaviour (assert-style aborts, nothing returned); the work behind
 * mr_insert_multi runs as sm_100a CUDA kernels (ropebwt2_b200/csrc/rb2_engine.cu).
 *
 * Six buckets: bucket i holds the BWT symbols whose following symbol is i
 * ($=0 A=1 C=2 G=3 T=4 N=5).
 */
#ifndef RB2_MROPE_H_
#define RB2_MROPE_H_

#include "rope.h"

#define MR_SO_IO    0 /* strings stay in input order */
#define MR_SO_RLO   1 /* reverse lexicographical order */
#define MR_SO_RCLO  2 /* reverse-complement lexicographical order */

typedef struct {
	uint8_t so;     /* sorting order, fixed at mr_init or inherited from the .fmr (reference mrope.h:11) */
	int thr_min;    /* accepted for compatibility; the GPU path has no serial-tail switch */
	rope_t *r[6];   /* per-bucket handles; r[a]->c[] are the marginal counts (read by mr_get_c) */
	void *priv;     /* engine handle -- new trailing field, invisible to reference-era callers */
} mrope_t;

typedef struct {
	mrope_t *r;
	int a, to_free;
	rpitr_t i;
} mritr_t;

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lifetime ------------------------------------------------------------------------------- */

/*
 * mr_init: an empty index of six buckets on CUDA device $RB2_DEVICE (else $LOCAL_RANK, else 0).
 * sorting_order is one of MR_SO_*; anything else aborts (reference mrope.c:14-25).  max_nodes and
 * block_len only shape the .fmr written by mr_dump -- the device leaf size is fixed at 512 bytes.
 */
mrope_t *mr_init(int max_nodes, int block_len, int sorting_order);

/* mr_destroy: releases the engine and the bucket handles that a freeing iteration left (mrope.c:27-33) */
void mr_destroy(mrope_t *r);

/* mr_thr_min: stored and returned like the reference does (mrope.c:35-40); no effect on the GPU path */
int mr_thr_min(mrope_t *r, int thr_min);

/* ---- insertion ------------------------------------------------------------------------------ */

/*
 * mr_insert_multi: the hot path (reference mrope.c:258-345).  s holds len bytes of nt6 codes, every
 * string REVERSED and NUL terminated, strings concatenated, s[len-1] == 0 (violations abort, as the
 * reference asserts).  The buffer is only read and may be reused on return.  is_thr is accepted for
 * compatibility; the GPU is always used.
 */
void mr_insert_multi(mrope_t *mr, int64_t len, const uint8_t *s, int is_thr);

/*
 * mr_insert1: one REVERSED, NUL-terminated string (reference mrope.c:42-68), served by the same
 * kernels as a one-string batch.  Returns the rank of the string's sentinel inside its bucket.
 */
int64_t mr_insert1(mrope_t *r, const uint8_t *str);

/* ---- queries and traversal ---------------------------------------------------------------- */

/* mr_rank2a: symbol counts in BWT[0,x) and BWT[0,y) over the concatenated buckets (mrope.c:70-105) */
void mr_rank2a(const mrope_t *mr, int64_t x, int64_t y, int64_t *cx, int64_t *cy);
#define mr_rank1a(mr, x, cx) mr_rank2a(mr, x, -1, cx, 0)

/*
 * Block iterator (reference mrope.c:111-130): buckets 0..5, leaves left to right.  Each block is
 * [uint16 nbytes][runs] -- decode with rle_dec1 from rle.h.  With to_free != 0 a bucket's host
 * handle is released once it has been walked (mr->r[a] becomes NULL), as in the reference.
 */
void mr_itr_first(mrope_t *r, mritr_t *i, int to_free);
const uint8_t *mr_itr_next_block(mritr_t *i);

/* ---- persistence ----------------------------------------------------------------------------- */

void mr_dump(mrope_t *mr, FILE *fp);   /* .fmr ("RB\2") writer; the reference's -i reads it */
mrope_t *mr_restore(FILE *fp);         /* .fmr reader; accepts files written by the reference's -b */
void mr_print_tree(const mrope_t *mr); /* Newick-style debug print of the tree mr_dump would write */

#ifdef __cplusplus
}
#endif

/* marginal counts over all buckets; returns the total (cf. reference mrope.h:86-97, without its c[6] over-read) */
static inline int64_t mr_get_c(const mrope_t *mr, int64_t c[6])
{
	int a, b;
	int64_t tot = 0;
	for (b = 0; b < 6; ++b) c[b] = 0;
	for (a = 0; a < 6; ++a)
		for (b = 0; b < 6; ++b)
			c[b] += mr->r[a]->c[b], tot += mr->r[a]->c[b];
	return tot;
}

/* accumulated counts: ac[a] = number of symbols smaller than a */
static inline int64_t mr_get_ac(const mrope_t *mr, int64_t ac[7])
{
	int a;
	int64_t c[6], tot = mr_get_c(mr, c);
	for (a = 0, ac[0] = 0; a < 6; ++a) ac[a + 1] = ac[a] + c[a];
	return tot;
}

static inline int64_t mr_get_tot(const mrope_t *mr)
{
	int64_t c[6];
	return mr_get_c(mr, c);
}

#endif
