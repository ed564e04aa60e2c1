/*
 * mrope_b200.c -- the reference's mrope.h / rope.h host API on top of the B200 engine.
 *
 * Plain C, like the reference's own host side.  Everything that touches the index goes
 * through the C-ABI in include/ropebwt2_b200.h; this file only does what stays on the host
 * in any design: argument checks, the block iterator's staging buffer, and the .fmr
 * reader/writer (reference mrope.c:136-160, rope.c:253-318).
 *
 * .fmr layout (little endian, raw structs), as written by the reference:
 *   "RB\2", uint8 so;  then for each of the 6 ropes:
 *   int32 max_nodes, int32 block_len, then nodes in pre-order:
 *     uint8 is_bottom, int16 n;
 *       bottom:   n x { int64 c[6]; uint16 nbytes; nbytes run bytes }
 *       internal: n child nodes
 * The tree shape is not canonical in the reference either (it depends on insertion
 * history), so mr_dump lays the flat leaf sequence out as a balanced
 * tree with fan-out 3/4 * max_nodes, which satisfies everything the reference's restore and
 * later insertions rely on: n <= max_nodes, nbytes + 18 <= block_len, block_len % 8 == 0.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../../include/mrope.h"
#include "../../include/rle.h"
#include "../../include/ropebwt2_b200.h"

#define ITR_CHUNK 8192 /* leaf blocks fetched from the GPU per iterator refill (4 MB) */
#define RESTORE_FILL 400 /* run bytes per leaf when re-blocking a restored index (leaves room to grow) */
#define DEV_MAXRUN ((1 << 19) - 1) /* device invariant: no run longer than the 4-byte form holds */

#define mr_fatal(...) do { fprintf(stderr, "[ropebwt2_b200] fatal: " __VA_ARGS__); fputc('\n', stderr); abort(); } while (0)

typedef struct {
	rb2_engine_t *eng;
	int max_nodes, block_len;
	uint8_t *itbuf;    /* ITR_CHUNK * 512 bytes */
	int64_t *itcnt;    /* ITR_CHUNK * 6 */
	const void *it_owner; int it_bucket; int64_t it_first; /* which iterator / bucket / first block the staging buffer holds */
	int standalone;    /* 1: owned by a rope_t created with rope_init/rope_restore */
} rb2_priv_t;

typedef struct { /* iterator cursor, overlaid on rpitr_t::pa (640 bytes) */
	int64_t nblk, next;        /* blocks in the bucket; next logical block to hand out */
	int64_t chunk_first, chunk_n;
} itr_state_t;

/* multi: the engine behind an mrope_t may be a proxy over several GPUs (RB2_GPUS=P); a bare rope_t never is */
static rb2_priv_t *priv_new(int device, int so, int max_nodes, int block_len, int multi)
{
	rb2_priv_t *p = (rb2_priv_t*)calloc(1, sizeof(rb2_priv_t));
	p->eng = multi? rb2_create_auto(device, so) : rb2_create(device, so);
	p->max_nodes = (max_nodes + 1) >> 1 << 1;        /* reference rope.c:60 */
	if (p->max_nodes < 4) p->max_nodes = 4;
	if (p->max_nodes > 510) p->max_nodes = 510;      /* rpnode_t::n has 9 bits (reference rope.h:13) */
	if (block_len < RB2_BLOCK_BYTES) block_len = RB2_BLOCK_BYTES; /* device leaves are 512 bytes */
	p->block_len = (block_len + 7) >> 3 << 3;        /* reference rope.c:61 */
	return p;
}

static void priv_free(rb2_priv_t *p)
{
	if (p == 0) return;
	rb2_destroy(p->eng);
	free(p->itbuf); free(p->itcnt);
	free(p);
}

static int env_device(void)
{
	const char *s = getenv("RB2_DEVICE");
	if (s && *s) return atoi(s);
	s = getenv("LOCAL_RANK");
	return s && *s? atoi(s) : 0;
}

static rope_t *rope_handle(rb2_priv_t *p, int bucket)
{
	rope_t *r = (rope_t*)calloc(1, sizeof(rope_t));
	r->max_nodes = p->max_nodes; r->block_len = p->block_len;
	r->node = p; r->leaf = (void*)(intptr_t)bucket;
	return r;
}

static void refresh_counts(mrope_t *mr)
{
	int64_t c[36];
	int a;
	rb2_priv_t *p = (rb2_priv_t*)mr->priv;
	rb2_counts(p->eng, c);
	for (a = 0; a < 6; ++a)
		if (mr->r[a]) memcpy(mr->r[a]->c, c + a * 6, 48);
}

/****************
 * multi-rope   *
 ****************/

mrope_t *mr_init(int max_nodes, int block_len, int sorting_order)
{
	mrope_t *mr;
	int a;
	if (sorting_order < 0 || sorting_order > 2) mr_fatal("mr_init: sorting order must be 0..2"); /* reference mrope.c:18 */
	mr = (mrope_t*)calloc(1, sizeof(mrope_t));
	mr->so = (uint8_t)sorting_order;
	mr->thr_min = 1000; /* reference mrope.c:21 */
	mr->priv = priv_new(env_device(), sorting_order, max_nodes, block_len, 1);
	for (a = 0; a < 6; ++a) mr->r[a] = rope_handle((rb2_priv_t*)mr->priv, a);
	return mr;
}

void mr_destroy(mrope_t *mr)
{
	int a;
	if (mr == 0) return;
	for (a = 0; a < 6; ++a) free(mr->r[a]); /* NULL after a freeing iteration (reference mrope.c:31) */
	priv_free((rb2_priv_t*)mr->priv);
	free(mr);
}

int mr_thr_min(mrope_t *mr, int thr_min)
{
	if (thr_min > 0) mr->thr_min = thr_min;
	return mr->thr_min;
}

void mr_insert_multi(mrope_t *mr, int64_t len, const uint8_t *s, int is_thr)
{
	rb2_priv_t *p = (rb2_priv_t*)mr->priv;
	(void)is_thr;
	if (mr->thr_min < 0) mr->thr_min = 0; /* reference mrope.c:267 */
	if (!(len > 0 && s[len-1] == 0)) mr_fatal("mr_insert_multi: len > 0 && s[len-1] == 0 violated"); /* reference mrope.c:268 */
	rb2_insert_multi(p->eng, len, s);
	refresh_counts(mr);
}

/* from rb2_engine.cu: bucket-local rank of the sentinel of the last single-string batch */
int64_t rb2_last_sentinel_rank(rb2_engine_t *e);

int64_t mr_insert1(mrope_t *mr, const uint8_t *str)
{
	rb2_priv_t *p = (rb2_priv_t*)mr->priv;
	rb2_insert_multi(p->eng, (int64_t)strlen((const char*)str) + 1, str);
	refresh_counts(mr);
	return rb2_last_sentinel_rank(p->eng);
}

void mr_rank2a(const mrope_t *mr, int64_t x, int64_t y, int64_t *cx, int64_t *cy)
{
	rb2_priv_t *p = (rb2_priv_t*)mr->priv;
	rb2_rank2a(p->eng, x, cy? y : -1, cx, cy);
}

/*************
 * iterators *
 *************/

static void itr_begin(rb2_priv_t *p, int bucket, rpitr_t *i)
{
	itr_state_t *s = (itr_state_t*)i->pa;
	memset(s, 0, sizeof(itr_state_t));
	s->nblk = rb2_num_blocks(p->eng, bucket);
	if (p->itbuf == 0) {
		p->itbuf = (uint8_t*)malloc((size_t)ITR_CHUNK * RB2_BLOCK_BYTES);
		p->itcnt = (int64_t*)malloc((size_t)ITR_CHUNK * 48);
	}
}

/* next block of `bucket` (pointer into the staging buffer, valid until the next call), or 0 */
static const uint8_t *itr_step(rb2_priv_t *p, int bucket, rpitr_t *i, const int64_t **cnt)
{
	itr_state_t *s = (itr_state_t*)i->pa;
	const uint8_t *ret;
	if (s->next >= s->nblk) return 0;
	/* all iterators of an index share one staging buffer: (re)fetch when this iterator's chunk is exhausted or
	 * another iterator (or a dump / print running in between) has used the buffer since */
	if (s->next >= s->chunk_first + s->chunk_n || p->it_owner != (const void*)i || p->it_bucket != bucket || p->it_first != s->chunk_first) {
		s->chunk_first = s->next;
		s->chunk_n = rb2_fetch_blocks(p->eng, bucket, s->chunk_first, ITR_CHUNK, p->itbuf, p->itcnt);
		p->it_owner = i; p->it_bucket = bucket; p->it_first = s->chunk_first;
	}
	ret = p->itbuf + (size_t)(s->next - s->chunk_first) * RB2_BLOCK_BYTES;
	if (cnt) *cnt = p->itcnt + (size_t)(s->next - s->chunk_first) * 6;
	++s->next;
	return ret;
}

void mr_itr_first(mrope_t *mr, mritr_t *i, int to_free)
{
	i->a = 0; i->r = mr; i->to_free = to_free;
	memset(&i->i, 0, sizeof(rpitr_t));
	i->i.rope = mr->r[0];
	itr_begin((rb2_priv_t*)mr->priv, 0, &i->i);
}

const uint8_t *mr_itr_next_block(mritr_t *i)
{
	rb2_priv_t *p = (rb2_priv_t*)i->r->priv;
	const uint8_t *s;
	if (i->a >= 6) return 0;
	while ((s = itr_step(p, i->a, &i->i, 0)) == 0) {
		if (i->to_free) { free(i->r->r[i->a]); i->r->r[i->a] = 0; } /* reference mrope.c:122-125 */
		if (++i->a == 6) return 0;
		i->i.rope = i->r->r[i->a];
		itr_begin(p, i->a, &i->i);
	}
	return s;
}

/***********************
 * .fmr dump / restore *
 ***********************/

typedef struct { rb2_priv_t *p; int bucket; rpitr_t it; FILE *fp; int fan; } dump_t;

/* write the subtree of depth d covering the next n leaves, pre-order */
static void dump_subtree(dump_t *D, int d, int64_t n)
{
	uint8_t is_bottom = (d == 1);
	int16_t k;
	int64_t cap = 1, j;
	int t;
	for (t = 1; t < d; ++t) cap *= D->fan;
	k = (int16_t)((n + cap - 1) / cap);
	fwrite(&is_bottom, 1, 1, D->fp);
	fwrite(&k, 2, 1, D->fp);
	if (is_bottom) {
		for (j = 0; j < n; ++j) {
			const int64_t *c = 0;
			const uint8_t *blk = itr_step(D->p, D->bucket, &D->it, &c);
			if (blk == 0 || c == 0) mr_fatal("mr_dump: bucket %d ended after %lld of %lld leaf blocks", D->bucket, (long long)j, (long long)n);
			fwrite(c, 8, 6, D->fp);
			fwrite(blk, 1, *rle_nptr(blk) + 2, D->fp);
		}
	} else {
		for (j = 0; j < k; ++j)
			dump_subtree(D, d - 1, n - j * cap < cap? n - j * cap : cap);
	}
}

static void dump_bucket(rb2_priv_t *p, int bucket, FILE *fp)
{
	dump_t D;
	int32_t mn = p->max_nodes, bl = p->block_len;
	int64_t n = rb2_num_blocks(p->eng, bucket), cap;
	int d = 1;
	D.p = p; D.bucket = bucket; D.fp = fp;
	D.fan = p->max_nodes * 3 / 4; if (D.fan < 2) D.fan = 2;
	for (cap = D.fan; cap < n; cap *= D.fan) ++d;
	memset(&D.it, 0, sizeof(D.it));
	itr_begin(p, bucket, &D.it);
	fwrite(&mn, 4, 1, fp);
	fwrite(&bl, 4, 1, fp);
	dump_subtree(&D, d, n);
}

void mr_dump(mrope_t *mr, FILE *fp)
{
	int a;
	fwrite("RB\2", 1, 3, fp);
	fwrite(&mr->so, 1, 1, fp);
	for (a = 0; a < 6; ++a) dump_bucket((rb2_priv_t*)mr->priv, a, fp);
}

typedef struct { /* re-blocker: run stream -> device-legal 512-byte leaves */
	uint8_t *blk; int64_t *cnt; int64_t n, cap;
	int c; int64_t l;       /* pending run */
	int fill;               /* run bytes in the open block */
} reblock_t;

static void rb_open(reblock_t *R)
{
	if (R->n == R->cap) {
		R->cap = R->cap? R->cap << 1 : 1024;
		R->blk = (uint8_t*)realloc(R->blk, (size_t)R->cap * RB2_BLOCK_BYTES);
		R->cnt = (int64_t*)realloc(R->cnt, (size_t)R->cap * 48);
	}
	memset(R->blk + (size_t)R->n * RB2_BLOCK_BYTES, 0, RB2_BLOCK_BYTES);
	memset(R->cnt + (size_t)R->n * 6, 0, 48);
	++R->n; R->fill = 0;
}

static void rb_flush(reblock_t *R)
{
	while (R->l > 0) {
		int64_t l = R->l < DEV_MAXRUN? R->l : DEV_MAXRUN;
		uint8_t *b;
		if (R->n == 0 || R->fill + 4 > RESTORE_FILL) rb_open(R);
		b = R->blk + (size_t)(R->n - 1) * RB2_BLOCK_BYTES;
		R->fill += rle_enc1(b + 2 + R->fill, R->c, l);
		*rle_nptr(b) = (uint16_t)R->fill;
		R->cnt[(size_t)(R->n - 1) * 6 + R->c] += l;
		R->l -= l;
	}
}

static void rb_push(reblock_t *R, int c, int64_t l)
{
	if (l <= 0) return;
	if (c == R->c) { R->l += l; return; }
	rb_flush(R);
	R->c = c; R->l = l;
}

static void restore_node(FILE *fp, reblock_t *R, int depth)
{
	uint8_t is_bottom;
	int16_t i, n;
	if (depth > ROPE_MAX_DEPTH) mr_fatal("mr_restore: tree deeper than %d", ROPE_MAX_DEPTH);
	if (fread(&is_bottom, 1, 1, fp) != 1 || fread(&n, 2, 1, fp) != 1) mr_fatal("mr_restore: truncated .fmr");
	if (is_bottom) {
		for (i = 0; i < n; ++i) {
			int64_t c[6];
			uint16_t nb;
			static uint8_t buf[65536];
			const uint8_t *q, *end;
			if (fread(c, 8, 6, fp) != 6 || fread(&nb, 2, 1, fp) != 1 || fread(buf, 1, nb, fp) != nb) mr_fatal("mr_restore: truncated .fmr");
			for (q = buf, end = buf + nb; q < end;) {
				int sym; int64_t l;
				rle_dec1(q, sym, l);
				rb_push(R, sym, l);
			}
		}
	} else for (i = 0; i < n; ++i) restore_node(fp, R, depth + 1);
}

static void restore_bucket(rb2_priv_t *p, int bucket, FILE *fp, int32_t *max_nodes, int32_t *block_len)
{
	reblock_t R;
	memset(&R, 0, sizeof(R));
	R.c = -1;
	if (fread(max_nodes, 4, 1, fp) != 1 || fread(block_len, 4, 1, fp) != 1) mr_fatal("mr_restore: truncated .fmr");
	restore_node(fp, &R, 0);
	rb_flush(&R);
	if (R.n) rb2_load_blocks(p->eng, bucket, R.n, R.blk, R.cnt);
	free(R.blk); free(R.cnt);
}

mrope_t *mr_restore(FILE *fp)
{
	mrope_t *mr;
	uint8_t magic[4];
	int64_t c[6];
	int a;
	int32_t mn = ROPE_DEF_MAX_NODES, bl = ROPE_DEF_BLOCK_LEN;
	rb2_priv_t *p;
	if (fread(magic, 1, 4, fp) != 4) mr_fatal("mr_restore: truncated .fmr");
	if (magic[3] > 2) mr_fatal("mr_restore: bad sorting order byte %d", magic[3]);
	mr = (mrope_t*)calloc(1, sizeof(mrope_t));
	mr->so = magic[3]; /* thr_min stays 0, as in the reference (mrope.c:152) */
	mr->priv = p = priv_new(env_device(), mr->so, ROPE_DEF_MAX_NODES, ROPE_DEF_BLOCK_LEN, 1);
	for (a = 0; a < 6; ++a) {
		restore_bucket(p, a, fp, &mn, &bl);
		if (a == 0) {
			rb2_priv_t *q = p;
			q->max_nodes = mn < 4? 4 : (mn > 510? 510 : mn);
			q->block_len = bl < RB2_BLOCK_BYTES? RB2_BLOCK_BYTES : bl;
		}
	}
	for (a = 0; a < 6; ++a) mr->r[a] = rope_handle(p, a);
	refresh_counts(mr);
	mr_get_c(mr, c);
	fprintf(stderr, "[M::%s] ($, A, C, G, T, N) = (%ld, %ld, %ld, %ld, %ld, %ld)\n", __func__,
			(long)c[0], (long)c[1], (long)c[2], (long)c[3], (long)c[4], (long)c[5]); /* reference mrope.c:157-158 */
	return mr;
}

/* Newick-like debug print (reference mrope.c:162-168, rope.c:225-251); shape = mr_dump's tree */
static void print_subtree(dump_t *D, int d, int64_t n)
{
	int64_t cap = 1, j, k;
	int t;
	for (t = 1; t < d; ++t) cap *= D->fan;
	k = (n + cap - 1) / cap;
	putchar('(');
	if (d == 1) {
		for (j = 0; j < n; ++j) {
			const uint8_t *blk = itr_step(D->p, D->bucket, &D->it, 0), *q = blk + 2, *end = blk + 2 + *rle_nptr(blk);
			if (j) putchar(',');
			while (q < end) {
				int c; int64_t l, x;
				rle_dec1(q, c, l);
				for (x = 0; x < l; ++x) putchar("$ACGTN"[c]);
			}
		}
	} else for (j = 0; j < k; ++j) {
		if (j) putchar(',');
		print_subtree(D, d - 1, n - j * cap < cap? n - j * cap : cap);
	}
	putchar(')');
}

void mr_print_tree(const mrope_t *mr)
{
	int a;
	for (a = 0; a < 6; ++a) {
		dump_t D;
		int64_t n, cap;
		int d = 1;
		D.p = (rb2_priv_t*)mr->priv; D.bucket = a; D.fp = 0;
		D.fan = D.p->max_nodes * 3 / 4; if (D.fan < 2) D.fan = 2;
		n = rb2_num_blocks(D.p->eng, a);
		for (cap = D.fan; cap < n; cap *= D.fan) ++d;
		memset(&D.it, 0, sizeof(D.it));
		itr_begin(D.p, a, &D.it);
		print_subtree(&D, d, n);
	}
	putchar('\n');
}

/***************
 * single rope *
 ***************/

/* from rb2_engine.cu: insert rl copies of symbol a behind the first x symbols of `bucket`;
 * returns the bucket-local rank(a, x) before the insertion */
int64_t rb2_insert_run(rb2_engine_t *e, int bucket, int64_t x, int a, int64_t rl);
void rb2_bucket_rank2a(rb2_engine_t *e, int bucket, int64_t x, int64_t y, int64_t cx[6], int64_t cy[6]);

rope_t *rope_init(int max_nodes, int block_len)
{
	rb2_priv_t *p = priv_new(env_device(), RB2_SO_IO, max_nodes, block_len, 0);
	p->standalone = 1;
	return rope_handle(p, 0);
}

void rope_destroy(rope_t *rope)
{
	rb2_priv_t *p;
	if (rope == 0) return;
	p = (rb2_priv_t*)rope->node;
	if (p && p->standalone) priv_free(p);
	free(rope);
}

int64_t rope_insert_run(rope_t *rope, int64_t x, int a, int64_t rl, rpcache_t *cache)
{
	rb2_priv_t *p = (rb2_priv_t*)rope->node;
	int bucket = (int)(intptr_t)rope->leaf;
	int64_t z;
	(void)cache;
	if (a < 0 || a > 5 || rl <= 0) mr_fatal("rope_insert_run: bad symbol or run length");
	z = rb2_insert_run(p->eng, bucket, x, a, rl);
	rope->c[a] += rl; /* reference rope.c:135 */
	return z;
}

void rope_rank2a(const rope_t *rope, int64_t x, int64_t y, int64_t *cx, int64_t *cy)
{
	rb2_priv_t *p = (rb2_priv_t*)rope->node;
	rb2_bucket_rank2a(p->eng, (int)(intptr_t)rope->leaf, x, (cy && y >= x)? y : -1, cx, cy);
}

void rope_itr_first(const rope_t *rope, rpitr_t *i)
{
	memset(i, 0, sizeof(rpitr_t));
	i->rope = rope;
	itr_begin((rb2_priv_t*)rope->node, (int)(intptr_t)rope->leaf, i);
}

const uint8_t *rope_itr_next_block(rpitr_t *i)
{
	return itr_step((rb2_priv_t*)i->rope->node, (int)(intptr_t)i->rope->leaf, i, 0);
}

void rope_print_node(const rpnode_t *p) { (void)p; }

void rope_dump(const rope_t *r, FILE *fp)
{
	dump_bucket((rb2_priv_t*)r->node, (int)(intptr_t)r->leaf, fp);
}

rope_t *rope_restore(FILE *fp)
{
	rb2_priv_t *p = priv_new(env_device(), RB2_SO_IO, ROPE_DEF_MAX_NODES, ROPE_DEF_BLOCK_LEN, 0);
	int32_t mn, bl;
	int64_t c[36];
	rope_t *r;
	p->standalone = 1;
	restore_bucket(p, 0, fp, &mn, &bl);
	p->max_nodes = mn < 4? 4 : (mn > 510? 510 : mn);
	p->block_len = bl < RB2_BLOCK_BYTES? RB2_BLOCK_BYTES : bl;
	r = rope_handle(p, 0);
	rb2_counts(p->eng, c);
	memcpy(r->c, c, 48);
	return r;
}

/**************
 * rle.h bits *
 **************/

void rle_count(const uint8_t *block, int64_t cnt[6])
{
	const uint8_t *q = block + 2, *end = q + *rle_nptr(block);
	while (q < end) {
		int c; int64_t l;
		rle_dec1(q, c, l);
		cnt[c] += l;
	}
}

void rle_print(const uint8_t *block, int expand)
{
	const uint8_t *q = block + 2, *end = q + *rle_nptr(block);
	while (q < end) {
		int c; int64_t l, x;
		rle_dec1(q, c, l);
		if (expand) for (x = 0; x < l; ++x) putchar("$ACGTN"[c]);
		else printf("%c%ld", "$ACGTN"[c], (long)l);
	}
	putchar('\n');
}
