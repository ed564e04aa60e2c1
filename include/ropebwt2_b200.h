/*
 * ropebwt2_b200.h -- C-ABI of the B200 (sm_100a) BCR insertion engine.
 *
 * This is the device shim under the reference-compatible host API in mrope.h / rope.h.
 * Plain C types only: no CUDA, torch or C++ types cross this boundary, so the library
 * can be bound from C (the reference's own language), ctypes, cgo, JNI, ...
 *
 * What each entry point replaces in lh3/ropebwt2 (file:line into the reference):
 *
 *   rb2_create / rb2_destroy     mr_init / mr_destroy              mrope.c:14-33
 *   rb2_insert_multi             the body of mr_insert_multi       mrope.c:258-345
 *                                incl. mr_insert_multi_aux         mrope.c:184-233
 *                                rope_insert_run / rope_rank2a     rope.c:114-194
 *                                rle_insert_cached / rle_rank2a    rle.c:10-89,134-191
 *   rb2_insert_multi_dev         same, input already resident in HBM (bench "value" leg)
 *   rb2_create_sharded / rb2_insert_multi_sharded   the same path on ONE index spread over several GPUs
 *                                (the reference's per-bucket worker threads, mrope.c:287-325, across devices)
 *   rb2_rank2a                   mr_rank2a                         mrope.c:70-105
 *   rb2_counts                   the rope_t::c[6] marginals        rope.h:19, mrope.h:86-116
 *   rb2_num_blocks/fetch_blocks  rope_itr_first/next_block         rope.c:200-219
 *   rb2_load_blocks              rope_restore (leaf upload)        rope.c:277-318
 *
 * Every function aborts the process with a message on a CUDA failure; there is no CPU
 * fallback anywhere behind this header.
 */
#ifndef ROPEBWT2_B200_H_
#define ROPEBWT2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB2_SO_IO   0  /* input order                     (MR_SO_IO,   mrope.h:6) */
#define RB2_SO_RLO  1  /* reverse lexicographical order   (MR_SO_RLO,  mrope.h:7) */
#define RB2_SO_RCLO 2  /* reverse-complement lex. order   (MR_SO_RCLO, mrope.h:8) */

#define RB2_BLOCK_BYTES 512 /* one leaf block in HBM: [uint16 nbytes][runs...], the rle.h layout (rle.h:36) */
#define RB2_BLOCK_FILL  494 /* max run bytes per block: block_len - RLE_MIN_SPACE (rope.c:143, rle.h:35) */

typedef struct rb2_engine rb2_engine_t;

/* counters and CUDA-event timings accumulated since the last rb2_reset_stats() */
typedef struct {
	int64_t n_strings, n_symbols;       /* inserted strings / symbols (sentinels included) */
	int64_t n_columns;                  /* BCR columns executed */
	int64_t n_launches;                 /* kernels launched by this library */
	int64_t n_merge_launches;           /* launches of the dominant kernel (k_flat_merge in the dense regime, k_merge_half in the sparse one) */
	int64_t merge_blocks;               /* work items of the merge kernels, all launches: 8192-symbol tiles (dense) / leaf blocks read (sparse) */
	int64_t merge_bytes_rw;             /* algorithmic HBM bytes of the merge kernels: arrays read + written + records (dense) / blocks read + written, x512 (sparse) */
	int64_t n_records;                  /* (position, symbol, count) insertion records merged */
	int64_t pool_blocks, pool_capacity; /* leaf blocks in use / allocated */
	double  ms_total;                   /* device time of whole rb2_insert_multi* calls (CUDA events) */
	double  ms_h2d;                     /* host->device copy of the batch */
	double  ms_transpose;               /* string split + column-major transpose */
	double  ms_members;                 /* symbol fetch, radix partition of the string set */
	double  ms_groups;                  /* group scan, record emission, rank pre-pass */
	double  ms_merge;                   /* the dominant kernel: k_flat_merge with its two planning kernels (dense) / k_merge_half (sparse) */
	double  ms_directory;               /* item planning + directory rebuild */
	double  ms_merge_general;           /* sparse regime: k_merge_fast + k_merge_general (over-full / multi-item / empty blocks) */
	int64_t general_items;              /* sparse regime: work items k_merge_half handed on */
	double  ms_exchange;                /* sharded build: ncclSend/ncclRecv of the interval starts behind the merge (0 with direct delivery) */
	int64_t exch_bytes;                 /* sharded build: bytes of string state this rank received */
	double  ms_convert;                 /* dense regime: leaf blocks <-> flat symbol array at the ends of a batch */
	int64_t flat_batches;               /* batches that ran in the dense regime (k_flat_merge instead of the block merges) */
	int64_t p2p_batches;                /* sharded build: batches whose merge kernels stored the new interval starts straight into the peer GPUs' memory */
} rb2_stats_t;

int  rb2_device_count(void);

/* Create an engine on CUDA device `device` with an empty six-bucket index. */
rb2_engine_t *rb2_create(int device, int sorting_order);
/* The engine behind an mrope_t (mr_init): like rb2_create, but with RB2_GPUS=P (P > 1) in the environment it is a
 * proxy over P sharded engines, one per GPU, so the unmodified reference driver uses every GPU of the node. */
rb2_engine_t *rb2_create_auto(int device, int sorting_order);
void rb2_destroy(rb2_engine_t *e);
int  rb2_sorting_order(const rb2_engine_t *e);
/* Empty the index (six empty buckets, as after rb2_create) but keep all HBM allocations. */
void rb2_reset(rb2_engine_t *e);

/*
 * Insert a batch.  `s` holds `len` bytes of nt6 codes ($=0 A=1 C=2 G=3 T=4 N=5), each string
 * REVERSED and NUL-terminated, strings concatenated; len > 0 and s[len-1] == 0 (mrope.c:268).
 * The buffer is only read and may be reused as soon as the call returns (main.c:243).
 */
void rb2_insert_multi(rb2_engine_t *e, int64_t len, const uint8_t *s_host);
void rb2_insert_multi_dev(rb2_engine_t *e, int64_t len, const uint8_t *s_dev);

/* c[b*6+a] = number of symbols a in bucket b (bucket b = symbols followed by b).
 * rb2_insert_multi on a one-GPU engine returns as soon as the batch has been copied to the device (the caller's
 * buffer is free again, main.c:243); the insertion runs on a worker thread while the caller prepares the next
 * batch.  The counts cover everything submitted (they follow from the batches alone); every other call first
 * waits for the queued batches.  RB2_ASYNC=0 makes rb2_insert_multi wait for the insertion itself. */
void rb2_counts(rb2_engine_t *e, int64_t c[36]);
/* wait for all queued batches (what every call except rb2_insert_multi / rb2_counts / rb2_reset does implicitly) */
void rb2_sync(rb2_engine_t *e);
/* device-timed span of a stream of calls: milliseconds (CUDA events) from rb2_span_begin to the end of the last
 * insertion queued before rb2_span_ms */
void rb2_span_begin(rb2_engine_t *e);
double rb2_span_ms(rb2_engine_t *e);
/* per-batch statistics of the last (at most `max`) rb2_insert_multi calls since rb2_reset_stats, oldest first */
int rb2_job_history(rb2_engine_t *e, rb2_stats_t *out, int max);

/* cx[a] = #a in BWT[0,x), cy[a] = #a in BWT[0,y) over the concatenated buckets; y < 0 or
 * cy == NULL skips the second query (mr_rank1a) */
void rb2_rank2a(rb2_engine_t *e, int64_t x, int64_t y, int64_t cx[6], int64_t cy[6]);

/* n rank queries in one call (one warp per position): out[i*6+a] = #a in BWT[0, x[i]) */
void rb2_rank_batch(rb2_engine_t *e, int64_t n, const int64_t *x, int64_t *out);

/* Leaf blocks of one bucket in logical (left-to-right) order. */
int64_t rb2_num_blocks(rb2_engine_t *e, int bucket);
/* Copy blocks [first, first+n) of `bucket` to host: dst gets n*512 bytes, cnt (optional)
 * n*6 per-block symbol counts.  Returns the number of blocks copied. */
int64_t rb2_fetch_blocks(rb2_engine_t *e, int bucket, int64_t first, int64_t n, uint8_t *dst, int64_t *cnt);
/* Append n host blocks (512 B each, rle.h layout, every run < 2^19, nbytes <= 494) to the
 * right end of `bucket`; cnt = n*6 per-block symbol counts.  Used by mr_restore. */
void rb2_load_blocks(rb2_engine_t *e, int bucket, int64_t n, const uint8_t *src, const int64_t *cnt);

void rb2_get_stats(rb2_engine_t *e, rb2_stats_t *st);
void rb2_reset_stats(rb2_engine_t *e);
/* stream the engine launches on (a cudaStream_t), for callers that time with their own events */
void *rb2_stream(rb2_engine_t *e);
/* device scratch helpers for callers that stage inputs in HBM themselves (bench, tests) */
void *rb2_dev_alloc(rb2_engine_t *e, int64_t bytes);
void  rb2_dev_free(rb2_engine_t *e, void *p);
void  rb2_dev_upload(rb2_engine_t *e, void *dst_dev, const void *src_host, int64_t bytes);
/* pinned host memory for batches handed to rb2_insert_multi / mr_insert_multi (faster H2D) */
void *rb2_host_alloc(int64_t bytes);
void  rb2_host_free(void *p);
/* rope.h back ends: one run into one bucket (rope_insert_run, rope.c:114-148) and bucket-local
 * rank (rope_rank2a, rope.c:179-194); mr_insert1's return value (mrope.c:67) */
int64_t rb2_insert_run(rb2_engine_t *e, int bucket, int64_t x, int a, int64_t rl);
void    rb2_bucket_rank2a(rb2_engine_t *e, int bucket, int64_t x, int64_t y, int64_t cx[6], int64_t cy[6]);
int64_t rb2_last_sentinel_rank(rb2_engine_t *e);

/*
 * Sharded build: ONE index spread over several GPUs.  The six buckets of the
 * reference's worker threads (mrope.c:312-325) are cut once more by the second symbol of the suffix
 * into 36 sub-buckets, each owned by one rank; per column the ranks all-gather the per-sub-bucket
 * symbol counts (the cross-bucket offsets of mrope.c:332-340) and hand every string to the owner of
 * the sub-bucket it inserts into next.
 *
 * Ranks are either processes (one per GPU, `nccl_uid` = 128 bytes from rb2_nccl_unique_id() on rank
 * 0, distributed by the caller) or threads of one process (`group` from rb2_group_create()).
 * rb2_insert_multi_sharded* is collective: every rank calls it with ITS share of the batch (len may
 * be 0); the strings are ordered by (rank, position in the rank's buffer).  rb2_counts, rb2_reset,
 * rb2_get_stats work as before; rb2_num_blocks / rb2_fetch_blocks address sub-bucket x*6+y (blocks
 * exist only on its owner, rb2_shard_owner()); the whole BWT is the concatenation over sub-buckets.
 *
 * Dense batches deliver the new interval starts directly: every rank maps the other ranks' state buffers (CUDA IPC
 * between processes) and the merge kernel stores into them; the mappings are kept from batch to batch.
 * rb2_sharded_quiesce (collective) closes them in an ordered way -- call it on every rank before the ranks destroy
 * their engines at different times; the next dense batch maps again.  RB2_P2P=0 keeps to ncclSend/ncclRecv.
 */
typedef struct rb2_group rb2_group_t;
rb2_group_t *rb2_group_create(int nranks);
void rb2_group_destroy(rb2_group_t *g);
void rb2_nccl_unique_id(uint8_t out[128]);
rb2_engine_t *rb2_create_sharded(int device, int sorting_order, int rank, int nranks, rb2_group_t *group, const uint8_t *nccl_uid);
void rb2_insert_multi_sharded(rb2_engine_t *e, int64_t len, const uint8_t *s_host);
void rb2_insert_multi_sharded_dev(rb2_engine_t *e, int64_t len, const uint8_t *s_dev);
void rb2_sharded_quiesce(rb2_engine_t *e);
int  rb2_shard_owner(int nranks, int subbucket);
int  rb2_num_buckets(const rb2_engine_t *e); /* 6, or 36 for a sharded engine */

#ifdef __cplusplus
}
#endif
#endif
