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

/*
 * Layout-compatibility types.  Nothing in this library walks a tree: the three structs below keep
 * the reference's field order and sizes only because callers allocate them (iterators live on the
 * caller's stack) or read them directly (mr_get_c reads rope_t::c).
 */

/* one B+-tree entry in the reference (rope.h:11-15): 8-byte pointer, 64-bit packed word, six counts */
typedef struct rpnode_s {
	struct rpnode_s *p;
	uint64_t l:54, n:9, is_bottom:1;
	int64_t c[6];
} rpnode_t;

/* per-bucket handle; byte offsets: 0 max_nodes, 4 block_len, 8 c[6], 56 root, 64 node, 72 leaf */
typedef struct {
	int32_t max_nodes, block_len; /* recorded for .fmr output only; device leaves are always 512 bytes */
	int64_t c[6];                 /* marginal symbol counts of the bucket, refreshed after every mutating call */
	rpnode_t *root;               /* always NULL here */
	void *node, *leaf;            /* node: private engine handle; leaf: bucket number */
} rope_t;

/* iterator cursor; the library stores its own state in pa[] (the struct is caller-allocated) */
typedef struct {
	const rope_t *rope;
	const rpnode_t *pa[ROPE_MAX_DEPTH];
	int ia[ROPE_MAX_DEPTH];
	int d;
} rpitr_t;

/* leaf-scan resume state of the reference (rope.h:31-35); accepted by rope_insert_run and ignored */
typedef struct {
	int beg;
	int64_t bc[6];
	uint8_t *p;
} rpcache_t;

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Lifetime.  rope_init creates a private engine (CUDA device $RB2_DEVICE / $LOCAL_RANK / 0) whose
 * bucket 0 backs the rope; rope_destroy releases it.  Ropes that belong to an mrope_t are created
 * and destroyed by mr_init / mr_destroy.
 */
rope_t *rope_init(int max_nodes, int block_len);
void rope_destroy(rope_t *rope);

/*
 * rope_insert_run: put `rl` copies of symbol `a` behind the first `x` symbols and return
 * rank(a, x) as it was before the insertion (reference rope.c:114-148).  One record through the
 * merge kernels; meant for API completeness, not throughput -- batches go through mr_insert_multi.
 */
int64_t rope_insert_run(rope_t *rope, int64_t x, int a, int64_t rl, rpcache_t *cache);

/*
 * rope_rank2a: cx[s] = number of s in [0,x), cy[s] = number of s in [0,y).  With y < x or
 * cy == NULL only cx is produced (reference rope.c:179-194).
 */
void rope_rank2a(const rope_t *rope, int64_t x, int64_t y, int64_t *cx, int64_t *cy);
#define rope_rank1a(rope, x, cx) rope_rank2a(rope, x, -1, cx, 0)

/* leaf blocks left to right; each returned pointer is valid until the next call (reference rope.c:200-219) */
void rope_itr_first(const rope_t *rope, rpitr_t *i);
const uint8_t *rope_itr_next_block(rpitr_t *i);

/* one rope's share of a .fmr file: int32 max_nodes, int32 block_len, nodes in pre-order (reference rope.c:253-318) */
void rope_dump(const rope_t *r, FILE *fp);
rope_t *rope_restore(FILE *fp);

/* there are no tree nodes to print: a no-op kept for link compatibility; see mr_print_tree */
void rope_print_node(const rpnode_t *p);

#ifdef __cplusplus
}
#endif

#endif
