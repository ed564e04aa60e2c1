/*
 * rope.h -- single-rope C API, source compatible with lh3/ropebwt2 (reference rope.h).
 *
 * In the reference a rope_t is a pointer-linked B+-tree of run-length-coded leaves
 * (rope.c).  Here a rope_t is a thin host handle: the leaves live in GPU memory as a pool
 * of 512-byte blocks behind a flat directory and every call below is
 * served by CUDA kernels.  The struct layouts are kept field-for-field so that code which
 * peeks into them keeps working -- notably mr_get_c() in mrope.h reads rope_t::c, which
 * this library keeps current after every mutating call.
 */
#ifndef RB2_ROPE_H_
#define RB2_ROPE_H_

#include <stdint.h>
#include <stdio.h>

#define ROPE_MAX_DEPTH 80
#define ROPE_DEF_MAX_NODES 64
#define ROPE_DEF_BLOCK_LEN 512

/* kept for layout compatibility (reference rope.h:11-15); the engine has no tree nodes */
typedef struct rpnode_s {
	struct rpnode_s *p;
	uint64_t l:54, n:9, is_bottom:1;
	int64_t c[6];
} rpnode_t;

typedef struct {
	int32_t max_nodes, block_len; /* recorded for .fmr output; the device leaf size is fixed at 512 */
	int64_t c[6];                 /* marginal symbol counts, always current (reference rope.h:19) */
	rpnode_t *root;               /* unused (NULL) */
	void *node, *leaf;            /* node: private handle of this library; leaf: unused */
} rope_t;

/* iterator state; same size as the reference's (rope.h:24-29) because callers allocate it */
typedef struct {
	const rope_t *rope;
	const rpnode_t *pa[ROPE_MAX_DEPTH];
	int ia[ROPE_MAX_DEPTH];
	int d;
} rpitr_t;

/* accepted and ignored: the reference uses it to resume leaf scans (rope.h:31-35) */
typedef struct {
	int beg;
	int64_t bc[6];
	uint8_t *p;
} rpcache_t;

#ifdef __cplusplus
extern "C" {
#endif

rope_t *rope_init(int max_nodes, int block_len);
void rope_destroy(rope_t *rope);
/* insert rl copies of symbol a behind the first x symbols; returns rank(a, x) before the insertion (reference rope.c:114-148) */
int64_t rope_insert_run(rope_t *rope, int64_t x, int a, int64_t rl, rpcache_t *cache);
/* cx[a] = #a in [0,x), cy[a] = #a in [0,y); y < x or cy == NULL: only cx (reference rope.c:179-194) */
void rope_rank2a(const rope_t *rope, int64_t x, int64_t y, int64_t *cx, int64_t *cy);
#define rope_rank1a(rope, x, cx) rope_rank2a(rope, x, -1, cx, 0)

void rope_itr_first(const rope_t *rope, rpitr_t *i);
const uint8_t *rope_itr_next_block(rpitr_t *i);

void rope_print_node(const rpnode_t *p); /* no tree nodes exist: prints nothing; use mr_print_tree */
void rope_dump(const rope_t *r, FILE *fp);
rope_t *rope_restore(FILE *fp);

#ifdef __cplusplus
}
#endif

#endif
