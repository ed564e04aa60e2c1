// rb2_engine.cu -- B200 (sm_100a) engine for ropebwt2's batched multi-string insertion.
//
// What it replaces: mr_insert_multi (reference mrope.c:258-345) and everything below it:
// mr_insert_multi_aux (mrope.c:184-233), rope_insert_run / rope_rank2a (rope.c:114-194),
// rle_insert_cached / rle_rank2a (rle.c:10-89, 134-191).  See DESIGN.md for the data layout
// and the derivation; the short version:
//
//  * The BWT lives in HBM as a pool of 512-byte leaf blocks in the reference's own leaf
//    format ([uint16 nbytes][43+3 coded runs], rle.h:36-75).  A flat directory replaces the
//    B+-tree: `order` (logical -> physical block), `cumLen` / `cumCnt` (exclusive prefix of
//    block lengths / per-symbol counts over ALL six buckets, so a directory lookup plus an
//    in-block decode is a whole-index rank, i.e. occ() including the cross-bucket offsets of
//    mrope.c:332-340).
//  * The string set of a batch is a column-major symbol matrix T[column][string] plus two
//    sorted structures: GROUPS (one per suffix-array interval: start position, interval
//    size, member range) and MEMBERS (string ids, grouped).  One BCR column =
//      members : fetch next symbol, 6-way stable partition               (mrope.c:189-190, 303-309)
//      groups  : per-group symbol histogram -> one insertion RECORD per   (mrope.c:191-224)
//                (group, symbol) in $,A,C,G,T,N / $,T,G,C,A,N order; every record with a
//                symbol != $ is also the next column's group
//      blocks  : k_merge_fast / k_merge_general: one warp per touched leaf block decodes it, merges its
//                records into the run stream, re-encodes, splits overfull blocks, and returns
//                rank(a, position) for every record = the next interval start (rope.c:114-148)
//      directory rebuild (prefix sums).
//  * No CPU fallback: every failure aborts.
#include <string.h>
#include <time.h>
#include <algorithm>
#include <vector>
#include <deque>
#include <thread>
#include <mutex>
#include <condition_variable>
#include "rb2_codec.cuh"
#include "rb2_comm.h"
#include "../../include/ropebwt2_b200.h"

#define MEM_TILE    1024   // members per CTA in the fetch / partition kernels (256 threads x 4)
#define RMAX        256    // max records merged by one work item (bounds the staging buffer)
#define STAGE_BYTES 2560   // >= 510 + 8*RMAX: old block bytes + <=8 new bytes per record
#define SPLIT_T     488    // piece size target when a block overflows (pieces are < SPLIT_T+4 <= 494)
#define MAXPIECES   16
#ifndef MERGE_WARPS
#define MERGE_WARPS 4
#endif
#ifndef MERGE_MINCTA
#define MERGE_MINCTA 10
#endif
#ifndef HALF_MINCTA
#define HALF_MINCTA 8
#endif
#define NONE32      0xffffffffu
#define SMALL_GROUP 32     // groups up to this size are histogrammed by one thread
#define NGC         13     // group-scan counters: has[6], hist[6], nrec

#define NBMAX 36          // buckets: 6 (one GPU: bucket = following symbol) or 36 (sharded: sub-bucket = following two symbols)
#define NBA   (NBMAX + 4)  // bucket-range arrays hold nb+1 boundaries plus padding

struct Ctl { // device control block (one per engine), mirrored through pinned host memory
	uint32_t poolUsed, err, nItems, nlogNew;
	uint32_t overflow, failBase;   // pool exhausted during a merge kernel: first block id that did not fit
	uint32_t nTodo, todoNext;     // items the fast kernel left for k_merge_general, and its work counter
	uint32_t nTodoA, todoANext;   // items the half-warp kernel left for k_merge_fast, and its work counter
	uint32_t nb, tables;          // number of buckets; sharded engines also fill grpPre / memPre
	uint32_t nrec, Gnext, Mnext, poolCap;
	uint32_t blkBkt[NBA];   // logical block range of bucket b: [blkBkt[b], blkBkt[b+1])
	uint32_t blkBktNew[NBA];
	uint32_t gBkt[NBA], mBkt[NBA];     // this column: group / member index range per bucket
	uint32_t recBkt[NBA];   // record index range per bucket (this column)
	uint32_t gSymBase[8], mSymBase[8]; // next column: first group / member index of the strings that insert symbol a now
	int64_t  cpost[8];      // global start position of bucket a AFTER this column's insertions
	uint32_t memTot[8];     // members per next symbol
	uint32_t grpTot[16];    // grand totals of the NGC group counters
	// sharded engines: per-symbol exclusive prefix of next groups / members at the first group of
	// every bucket (row nb = totals); the differences of two rows are what a bucket sends on
	uint32_t grpPre[(NBMAX + 1) * 6], memPre[(NBMAX + 1) * 6];
};

struct Dir {
	uint32_t *order;   // [cap]      logical -> physical block id
	int64_t  *cumLen;  // [cap+1]    symbols in front of logical block i (all buckets)
	int64_t  *cumCnt;  // [cap+1][6] per-symbol counts in front of logical block i
	size_t cap;
};

// =====================================================================================
// Batch setup: split the NUL-delimited buffer into strings, transpose to column-major
// =====================================================================================

// 256 threads x 16 bytes: count NUL bytes per 4096-byte tile
__global__ void __launch_bounds__(256) k_count_nul(const uint8_t *s, int64_t len, uint32_t *tileCnt)
{
	__shared__ uint32_t sm[8];
	int64_t off = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
	uint32_t c = 0;
	if (off + 16 <= len) {
		uint4 v = *reinterpret_cast<const uint4*>(s + off);
		c = (__popc(__vcmpeq4(v.x, 0)) + __popc(__vcmpeq4(v.y, 0)) + __popc(__vcmpeq4(v.z, 0)) + __popc(__vcmpeq4(v.w, 0))) >> 3;
	} else for (int64_t i = off; i < len; ++i) c += s[i] == 0;
	c = warp_sum(c);
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < 8; ++w) t += sm[w];
		tileCnt[blockIdx.x] = t;
	}
}

// strEnd[k] = byte offset of the k-th NUL
__global__ void __launch_bounds__(256) k_string_ends(const uint8_t *s, int64_t len, const uint32_t *tilePre, int64_t *strEnd)
{
	__shared__ uint32_t sm[8];
	int64_t off = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
	uint32_t nul = 0; // bit i: byte off+i is a NUL
	if (off + 16 <= len) {
		const uint4 v = *reinterpret_cast<const uint4*>(s + off);
		const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t z = __vcmpeq4(w[j], 0) & 0x01010101u;      // one flag bit per byte
			nul |= (((z * 0x00204081u) >> 21) & 0xfu) << (4 * j);     // gather the four flags (bits 0, 8, 16, 24) into a nibble
		}
	} else for (int i = 0; i < 16; ++i) if (off + i < len && s[off + i] == 0) nul |= 1u << i;
	uint32_t v[1] = { (uint32_t)__popc(nul) }, tot[1];
	cta_excl_scan<1, 256, uint32_t>(v, tot, sm);
	uint32_t k = tilePre[blockIdx.x] + v[0];
	while (nul) { strEnd[k++] = off + (__ffs(nul) - 1); nul &= nul - 1; }
}

// (strings kBase .. kBase+m-1 of the batch: a batch whose symbol matrix would not fit is processed in ranges)
__global__ void k_maxlen(const int64_t *strEnd, uint32_t kBase, uint32_t m, unsigned long long *maxlen)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long l = 0;
	if (k < m) { k += kBase; l = (unsigned long long)(strEnd[k] - (k ? strEnd[k-1] + 1 : 0)); }
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { unsigned long long y = __shfl_xor_sync(FULLMASK, l, o); l = y > l ? y : l; }
	if ((threadIdx.x & 31) == 0 && l) atomicMax(maxlen, l);
}

// hist[b*6+a] += #positions i with s[i-1] == b (0 in front of the batch) and s[i] == a.  The BWT symbol a = s[i] of a
// (reversed, NUL-terminated) string sits in bucket b = s[i-1], and a string's sentinel is the pair (last symbol, NUL):
// these are exactly the marginal counts mr_get_c reports (mrope.h:86-97), available as soon as the batch is on the
// device -- the insertion itself may still be running (rb2_insert_multi returns once the copy is done).
// Persistent CTAs; lane-private shared-memory bins (no atomics, no bank conflicts).
__global__ void __launch_bounds__(256) k_pair_hist(const uint8_t *s, int64_t len, unsigned long long *hist)
{
	__shared__ uint32_t bins[36][256];
	const int tid = threadIdx.x;
	for (int k = 0; k < 36; ++k) bins[k][tid] = 0;
	const int64_t nChunk = (len + 15) / 16;
	for (int64_t c = (int64_t)blockIdx.x * 256 + tid; c < nChunk; c += (int64_t)gridDim.x * 256) {
		const int64_t off = c * 16;
		uint32_t prev = off ? s[off - 1] : 0u;
		uint8_t b[16];
		if (off + 16 <= len) { const uint4 v = *reinterpret_cast<const uint4*>(s + off); memcpy(b, &v, 16); }
		else for (int i = 0; i < 16; ++i) b[i] = off + i < len ? s[off + i] : 255;
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			const uint32_t a = b[i];
			if (a < 6u && prev < 6u) ++bins[prev * 6 + a][tid];
			prev = a;
		}
	}
	__syncthreads();
	if (tid < 36) {
		unsigned long long t = 0;
		for (int k = 0; k < 256; ++k) t += bins[tid][(k + tid) & 255];
		if (t) atomicAdd(hist + tid, t);
	}
}

// Column-major symbol matrix of a batch, 4 bits per symbol: byte T[j*tstride + (k >> 1)], nibble k & 1,
// holds the j-th symbol of (reversed) string k, including its terminating NUL (tstride = bytes per
// column, rounded up to 16).  One CTA transposes 128 strings, 32 symbols at a time: reads walk along the
// strings (aligned words + funnel shift, 16 bytes per thread), shared memory holds the slab as
// [4 symbols][string] words, and half a warp stores the same symbol of 128 consecutive strings (64 bytes).
#define TR_S 128
__host__ __device__ __forceinline__ uint64_t t_stride(uint64_t m) { return (((m + 1) >> 1) + 15) & ~(uint64_t)15; }
__global__ void __launch_bounds__(256) k_transpose(const uint8_t *s, const int64_t *strEnd, uint32_t kBase, uint32_t m, int64_t ncol, uint8_t *T)
{
	__shared__ __align__(16) uint32_t tile[8][TR_S + 4];
	__shared__ int64_t sStart[TR_S];
	__shared__ int32_t sLen[TR_S];
	__shared__ int32_t sMax;
	const uint64_t mstride = t_stride(m);
	const uint32_t k0 = blockIdx.x * TR_S;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) sMax = -1;
	__syncthreads();
	if (tid < TR_S) {
		const uint32_t k = k0 + tid;
		int64_t st = 0; int32_t ln = -1;
		if (k < m) { const uint32_t kk = kBase + k; st = kk ? strEnd[kk-1] + 1 : 0; ln = (int32_t)(strEnd[kk] - st); }
		sStart[tid] = st; sLen[tid] = ln;
		atomicMax(&sMax, ln);
	}
	__syncthreads();
	const int32_t maxl = sMax;
	const int r = tid >> 1, half = tid & 1;
	for (int64_t j0 = 0; j0 <= maxl; j0 += 32) {
		{ // 16 symbols of string r from index j0 + 16*half on (zero behind the terminating NUL)
			const int64_t jb = j0 + half * 16;
			const int64_t left = (int64_t)sLen[r] + 1 - jb;
			uint32_t w[4] = { 0, 0, 0, 0 };
			if (left > 0) {
				const int64_t a = sStart[r] + jb;
				const uint32_t *wp = reinterpret_cast<const uint32_t*>(s + (a & ~(int64_t)3));
				const uint32_t sh = (uint32_t)(a & 3) * 8;
				const uint32_t x0 = wp[0], x1 = wp[1], x2 = wp[2], x3 = wp[3], x4 = wp[4];
				w[0] = __funnelshift_r(x0, x1, sh); w[1] = __funnelshift_r(x1, x2, sh); w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
				if (left < 16) {
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						const int64_t kb = left - i * 4;
						w[i] = kb >= 4 ? w[i] : (kb > 0 ? w[i] & ((1u << (kb * 8)) - 1u) : 0u);
					}
				}
			}
#pragma unroll
			for (int i = 0; i < 4; ++i) tile[half * 4 + i][r] = w[i];
		}
		__syncthreads();
		for (int cp = wid; cp < 16; cp += 8) { // two columns per warp pass: lanes 0-15 / 16-31; a lane packs 8 strings
			const int c = cp * 2 + (lane >> 4), hl = lane & 15;
			const int64_t j = j0 + c;
			if (j < ncol && (uint64_t)(k0 >> 1) + 4 * hl < mstride) {
				const uint4 v0 = *reinterpret_cast<const uint4*>(&tile[c >> 2][8 * hl]), v1 = *reinterpret_cast<const uint4*>(&tile[c >> 2][8 * hl + 4]);
				const uint32_t sel = 0x0040 + (c & 3) * 0x0011; // result bytes 0,1 = byte (c&3) of the first / second operand
				uint32_t a = (__byte_perm(v0.x, v0.y, sel) & 0xffffu) | (__byte_perm(v0.z, v0.w, sel) << 16); // symbols of strings 0..3, one per byte
				uint32_t b = (__byte_perm(v1.x, v1.y, sel) & 0xffffu) | (__byte_perm(v1.z, v1.w, sel) << 16); // strings 4..7
				a = (a | (a >> 4)) & 0x00ff00ffu; a = (a | (a >> 8)) & 0xffffu;   // four nibbles
				b = (b | (b >> 4)) & 0x00ff00ffu; b = (b | (b >> 8)) & 0xffffu;
				*reinterpret_cast<uint32_t*>(T + j * mstride + (k0 >> 1) + 4 * hl) = a | (b << 16);
			}
		}
		__syncthreads();
	}
}

// Column 0 state (mrope.c:279-284): sorted modes start with one group [0, n0) holding every
// string; input order starts with one empty-interval group per string at the end of bucket $.
__global__ void k_init_state(int sorted, uint32_t m, int64_t n0, int64_t *gL, int64_t *gSize, uint32_t *gOff, uint32_t *sid)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < m) sid[k] = k;
	if (sorted) {
		if (k == 0) { gL[0] = 0; gSize[0] = n0; gOff[0] = 0; gOff[1] = m; }
	} else {
		if (k < m) { gL[k] = n0; gSize[k] = 0; gOff[k] = k; }
		if (k == 0) gOff[m] = m;
	}
}

// =====================================================================================
// Members: fetch the next symbol of every live string; stable 6-way partition
// =====================================================================================

// The column-major symbol matrix of a batch.  A sharded build keeps one matrix per rank (the rank's
// own strings), replicated on every GPU; string ids are global, rank r owns [off[r], off[r+1]).
#define TV_MAX 8
struct TView { int n; uint32_t off[TV_MAX + 1]; const uint8_t *col[TV_MAX]; }; // col[r] = rank r's current column

__device__ __forceinline__ uint32_t tview_fetch(const TView &tv, uint32_t id)
{
	int r = 0;
#pragma unroll
	for (int x = 1; x < TV_MAX; ++x) r += x < tv.n && id >= tv.off[x];
	const uint8_t *c = tv.col[0]; uint32_t o = tv.off[0];
#pragma unroll
	for (int x = 1; x < TV_MAX; ++x) if (r == x) { c = tv.col[x]; o = tv.off[x]; }
	return (c[(id - o) >> 1] >> (((id - o) & 1) * 4)) & 15u;
}

__global__ void __launch_bounds__(256) k_member_fetch(const TView tv, const uint32_t *sid, uint32_t M, uint8_t *asym, uint32_t *tileTot)
{
	__shared__ uint32_t sm[6 * 8];
	const uint32_t k = blockIdx.x * MEM_TILE + threadIdx.x * 4;
	uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
	uint32_t packed = 0;
	if (k < M) {
		uint32_t id[4];
		if (k + 4 <= M) { uint4 v = *reinterpret_cast<const uint4*>(sid + k); id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w; }
		else for (int i = 0; i < 4; ++i) id[i] = k + i < M ? sid[k + i] : 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (k + i < M) {
				uint32_t a = tv.n == 1 ? (tv.col[0][id[i] >> 1] >> ((id[i] & 1) * 4)) & 15u : tview_fetch(tv, id[i]);
				packed |= a << (8 * i);
#pragma unroll
				for (int x = 0; x < 6; ++x) c[x] += a == x;
			}
		}
		*reinterpret_cast<uint32_t*>(asym + k) = packed; // asym is padded to a multiple of 4
	}
#pragma unroll
	for (int x = 0; x < 6; ++x) {
		uint32_t t = warp_sum(c[x]);
		if ((threadIdx.x & 31) == 0) sm[x * 8 + (threadIdx.x >> 5)] = t;
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		uint32_t t = 0;
		for (int w = 0; w < 8; ++w) t += sm[threadIdx.x * 8 + w];
		tileTot[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
	}
}

// dest = start of next bucket a + #earlier members with symbol a (mrope.c:303-309); members
// whose symbol is the sentinel are finished and dropped (mrope.c:310)
__global__ void __launch_bounds__(256) k_partition(const uint32_t *sid, const uint8_t *asym, uint32_t M, const uint32_t *tilePre,
                                                   const Ctl *ctl, uint32_t *sidNext)
{
	__shared__ uint32_t sm[6 * 8];
	const uint32_t k = blockIdx.x * MEM_TILE + threadIdx.x * 4;
	uint32_t a4 = 0, id[4] = { 0, 0, 0, 0 };
	uint32_t c[6] = { 0, 0, 0, 0, 0, 0 }, tot[6];
	if (k < M) {
		a4 = *reinterpret_cast<const uint32_t*>(asym + k);
		for (int i = 0; i < 4; ++i) if (k + i < M) {
			id[i] = sid[k + i];
			uint32_t a = (a4 >> (8 * i)) & 0xff;
#pragma unroll
			for (int x = 0; x < 6; ++x) c[x] += a == x;
		}
	}
	cta_excl_scan<6, 256, uint32_t>(c, tot, sm);
	if (k < M) {
		uint32_t base[6];
#pragma unroll
		for (int x = 0; x < 6; ++x) base[x] = ctl->mSymBase[x] + tilePre[(size_t)blockIdx.x * 6 + x] + c[x];
		for (int i = 0; i < 4; ++i) if (k + i < M) {
			uint32_t a = (a4 >> (8 * i)) & 0xff;
			uint32_t d = 0;
#pragma unroll
			for (int x = 1; x < 6; ++x) if (a == x) d = base[x]++;
			if (a) sidNext[d] = id[i];
		}
	}
}

// =====================================================================================
// Rank: whole-index occ(., x) from the flat directory + one warp-decoded block
// =====================================================================================

// logical block whose range (cumLen[i], cumLen[i+1]] contains x (block 0 for x == 0)
__device__ __forceinline__ uint32_t find_block(const int64_t *cumLen, uint32_t lo, uint32_t hi, int64_t x)
{
	// smallest i in [lo,hi) with cumLen[i+1] >= x; cumLen[hi] >= x is guaranteed by the caller
	while (lo < hi) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		if (cumLen[mid + 1] >= x) hi = mid; else lo = mid + 1;
	}
	return lo;
}

// occ(a, x) for all six symbols; warp-cooperative, result valid in every lane.
// scratch: img[RB2_IMG_BYTES], cnt[32*7] words, res[6] int64 -- private to the warp.
__device__ void warp_rank6(const uint8_t *pool, Dir dir, uint32_t nlog, int64_t x, int lane,
                           uint8_t *img, uint32_t *cntScratch, int64_t *res, int64_t (&out)[6], uint32_t &err)
{
	uint32_t i = find_block(dir.cumLen, 0, nlog - 1, x);
	if (i >= nlog) i = nlog - 1;
	// x at the very end of block i: the answer is the directory entry of block i+1.  In a sharded
	// engine this also steps over the symbols that other ranks hold between the two blocks.
	if (i + 1 < nlog && x >= dir.cumLen[i + 1]) {
#pragma unroll
		for (int a = 0; a < 6; ++a) out[a] = dir.cumCnt[(size_t)(i + 1) * 6 + a];
		return;
	}
	const uint32_t xrel = (uint32_t)(x - dir.cumLen[i]);
	LaneDec d; uint32_t basePos, baseCnt[6], blkLen, blkCnt[6], nbytes; uint4 own;
	warp_decode_block(pool + (size_t)dir.order[i] * RB2_BLK, lane, img, cntScratch, d, basePos, baseCnt, blkLen, blkCnt, nbytes, err, own);
	const bool mine = xrel > basePos && xrel <= basePos + d.len;
	if (xrel == 0) { if (lane < 6) res[lane] = dir.cumCnt[(size_t)i * 6 + lane]; }
	else if (mine) {
		uint32_t pc[6] = { baseCnt[0], baseCnt[1], baseCnt[2], baseCnt[3], baseCnt[4], baseCnt[5] };
		uint32_t pos = basePos, bp = lane * 16 + d.fb;
		for (uint32_t q = 0; q < d.nr && pos < xrel; ++q) {
			uint32_t l, s, nb;
			parse_run(img, bp, s, l, nb);
			const uint32_t take = xrel - pos < l ? xrel - pos : l;
#pragma unroll
			for (int a = 0; a < 6; ++a) pc[a] += s == a ? take : 0;
			pos += l; bp += nb;
		}
#pragma unroll
		for (int a = 0; a < 6; ++a) res[a] = dir.cumCnt[(size_t)i * 6 + a] + pc[a];
	}
	__syncwarp();
#pragma unroll
	for (int a = 0; a < 6; ++a) out[a] = res[a];
	__syncwarp();
}

// sizes6[g][a] = #a in [gL, gL+gSize) for every group with a non-empty interval (rope_rank2a, mrope.c:202)
__global__ void __launch_bounds__(128) k_rank_groups(const uint8_t *pool, Dir dir, uint32_t nlog, uint32_t G,
                                                     const int64_t *gL, const int64_t *gSize, int64_t *sizes6, Ctl *ctl)
{
	__shared__ __align__(16) uint8_t sRuns[4][RB2_IMG_BYTES];
	__shared__ uint32_t sCnt[4][32 * 7];
	__shared__ int64_t sRes[4][6];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t g0 = (blockIdx.x * 4 + wid) * 32;
	if (g0 >= G) return;
	const uint32_t g = g0 + lane;
	const int64_t myL = g < G ? gL[g] : 0, mySz = g < G ? gSize[g] : 0;
	uint32_t todo = __ballot_sync(FULLMASK, mySz > 0), err = 0;
	while (todo) {
		const int src = __ffs(todo) - 1; todo &= todo - 1;
		const int64_t L = __shfl_sync(FULLMASK, myL, src), sz = __shfl_sync(FULLMASK, mySz, src);
		int64_t cl[6], cu[6];
		warp_rank6(pool, dir, nlog, L, lane, sRuns[wid], sCnt[wid], sRes[wid], cl, err);
		warp_rank6(pool, dir, nlog, L + sz, lane, sRuns[wid], sCnt[wid], sRes[wid], cu, err);
		if (lane < 6) {
			int64_t v = 0;
#pragma unroll
			for (int a = 0; a < 6; ++a) if (lane == a) v = cu[a] - cl[a];
			sizes6[(size_t)(g0 + src) * 6 + lane] = v;
		}
	}
	if (err) atomicOr(&ctl->err, err);
}

// API rank (mr_rank2a): one warp, up to two queries
__global__ void __launch_bounds__(32) k_rank_query(const uint8_t *pool, Dir dir, uint32_t nlog, int64_t x, int64_t y, int64_t *out, Ctl *ctl)
{
	__shared__ __align__(16) uint8_t sRuns[RB2_IMG_BYTES];
	__shared__ uint32_t sCnt[32 * 7];
	__shared__ int64_t sRes[6];
	const int lane = threadIdx.x;
	uint32_t err = 0;
	int64_t c[6];
	warp_rank6(pool, dir, nlog, x, lane, sRuns, sCnt, sRes, c, err);
	if (lane == 0) for (int a = 0; a < 6; ++a) out[a] = c[a];
	if (y >= 0) {
		warp_rank6(pool, dir, nlog, y, lane, sRuns, sCnt, sRes, c, err);
		if (lane == 0) for (int a = 0; a < 6; ++a) out[6 + a] = c[a];
	}
	if (err && lane == 0) atomicOr(&ctl->err, err);
}

// batched API rank (a query service on top of mr_rank1a, mrope.c:70-105): one warp per position
__global__ void __launch_bounds__(128) k_rank_batch(const uint8_t *pool, Dir dir, uint32_t nlog, uint32_t n, const int64_t *x, int64_t *out, Ctl *ctl)
{
	__shared__ __align__(16) uint8_t sRuns[4][RB2_IMG_BYTES];
	__shared__ uint32_t sCnt[4][32 * 7];
	__shared__ int64_t sRes[4][6];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t err = 0;
	for (uint32_t q = blockIdx.x * 4 + wid; q < n; q += gridDim.x * 4) {
		int64_t c[6];
		warp_rank6(pool, dir, nlog, x[q], lane, sRuns[wid], sCnt[wid], sRes[wid], c, err);
		if (lane < 6) {
			int64_t v = 0;
#pragma unroll
			for (int a = 0; a < 6; ++a) if (lane == a) v = c[a];
			out[(size_t)q * 6 + lane] = v;
		}
	}
	if (err && lane == 0) atomicOr(&ctl->err, err);
}

// =====================================================================================
// Groups: per-group histogram of next symbols, scan, record + next-group emission
// =====================================================================================

// cooperative #symbol counts of asym[S, E) for one warp; result in every lane
__device__ __forceinline__ void warp_count_range(const uint8_t *asym, uint32_t S, uint32_t E, int lane, uint32_t (&h)[6])
{
#pragma unroll
	for (int a = 0; a < 6; ++a) h[a] = 0;
	for (uint32_t k = S + lane; k < E; k += 32) {
		uint32_t s = asym[k];
#pragma unroll
		for (int a = 0; a < 6; ++a) h[a] += s == a;
	}
#pragma unroll
	for (int a = 0; a < 6; ++a) h[a] = warp_sum(h[a]);
}

// histogram of a big group [S,E): direct count if short, else differences of the sampled
// global prefix counts (tilePre has one entry per MEM_TILE members plus a terminal one)
__device__ __forceinline__ void warp_group_hist(const uint8_t *asym, const uint32_t *tilePre, uint32_t S, uint32_t E, int lane, uint32_t (&h)[6])
{
	if (E - S <= 4 * MEM_TILE) { warp_count_range(asym, S, E, lane, h); return; }
	uint32_t hs[6], he[6];
	const uint32_t ts = S / MEM_TILE, te = E / MEM_TILE;
	warp_count_range(asym, ts * MEM_TILE, S, lane, hs);
	warp_count_range(asym, te * MEM_TILE, E, lane, he);
#pragma unroll
	for (int a = 0; a < 6; ++a)
		h[a] = (tilePre[(size_t)te * 6 + a] + he[a]) - (tilePre[(size_t)ts * 6 + a] + hs[a]);
}

struct GroupArgs {
	const uint32_t *gOff; const int64_t *gL, *gSize; const int64_t *sizes6; // current column
	const uint8_t *asym; const uint32_t *tilePre;
	uint32_t G;
	uint32_t *ctaTot;        // [nCta][NGC]: totals (MODE 0) / exclusive prefix (MODE 1)
	int64_t *gSizeNext; uint32_t *gOffNext;
	int64_t *recP; uint32_t *recSC, *recDst; // recSC = count << 3 | symbol
	Ctl *ctl;
	uint32_t *recPre;        // dense regime: members in front of the record = symbols inserted in front of it (or null)
};

// MODE 0: reduce (per-CTA totals).  MODE 1: emit.  COMP: RCLO insertion order $,T,G,C,A,N.
template <int MODE, bool COMP>
__global__ void __launch_bounds__(256) k_group_pass(GroupArgs A)
{
	__shared__ uint32_t sm[NGC * 8];
	__shared__ uint64_t sm64[3 * 8];
	__shared__ uint32_t sCtl[24]; // [0,8) gSymBase, [8,16) mSymBase, [21] a bucket starts inside this CTA, [22] all singletons, [23] member base
	__shared__ uint32_t sBkt[NBA]; // first group of bucket b if it lies inside this CTA
	const int lane = threadIdx.x & 31;
	const uint32_t g0 = blockIdx.x * 256, g = g0 + threadIdx.x;
	const bool valid = g < A.G;
	const uint32_t nv = A.G - g0 < 256 ? A.G - g0 : 256;
	if (threadIdx.x == 0) {
		const uint32_t mb = A.gOff[g0];
		sCtl[23] = mb;
		sCtl[22] = A.gOff[g0 + nv] - mb == nv; // every group of this CTA is a singleton
	}
	if (MODE == 1) {
		if (threadIdx.x < 8) { sCtl[threadIdx.x] = A.ctl->gSymBase[threadIdx.x]; sCtl[8 + threadIdx.x] = A.ctl->mSymBase[threadIdx.x]; }
		if (threadIdx.x >= 32 && threadIdx.x < 96) { // two warps cover up to 64 buckets
			const uint32_t b = threadIdx.x - 32;
			bool in = false;
			if (b < A.ctl->nb) {
				const uint32_t gb = A.ctl->gBkt[b];
				in = gb >= g0 && gb < g0 + 256;
				sBkt[b] = in ? gb : NONE32;
			}
			const uint32_t any = __ballot_sync(FULLMASK, in);
			if (lane == 0) sCtl[20 + (b >> 5)] = any != 0; // [20], [21]
		}
	}
	__syncthreads();
	const bool allSingle = sCtl[22] != 0;
	uint32_t S, E;
	if (allSingle) { S = sCtl[23] + threadIdx.x; E = S + (valid ? 1 : 0); }
	else { S = valid ? A.gOff[g] : 0; E = valid ? A.gOff[g + 1] : 0; }
	const uint32_t cnt = E - S;
	uint32_t h[6] = { 0, 0, 0, 0, 0, 0 };
	if (valid && cnt <= SMALL_GROUP) {
		for (uint32_t k = S; k < E; ++k) {
			uint32_t s = A.asym[k];
#pragma unroll
			for (int a = 0; a < 6; ++a) h[a] += s == a;
		}
	}
	uint32_t big = allSingle ? 0u : __ballot_sync(FULLMASK, valid && cnt > SMALL_GROUP);
	while (big) {
		const int src = __ffs(big) - 1; big &= big - 1;
		uint32_t hh[6];
		warp_group_hist(A.asym, A.tilePre, __shfl_sync(FULLMASK, S, src), __shfl_sync(FULLMASK, E, src), lane, hh);
		if (lane == src) {
#pragma unroll
			for (int a = 0; a < 6; ++a) h[a] = hh[a];
		}
	}
	uint32_t v[NGC], tot[NGC];
	uint32_t nrec = 0;
#pragma unroll
	for (int a = 0; a < 6; ++a) {
		v[a] = h[a] != 0;
		v[6 + a] = h[a];
		nrec += (h[a] + RB2_MAXRUN - 1) / RB2_MAXRUN;
	}
	v[12] = nrec;
	// All 13 counters of a CTA of small groups fit 16-bit fields (256 threads x <= 32 members), so
	// three 64-bit scans replace thirteen 32-bit ones; a CTA holding a big group takes the wide path.
	const int anyBig = allSingle ? 0 : __syncthreads_or(valid && cnt > SMALL_GROUP);
	if (!anyBig) {
		uint64_t pk[3], pt[3];
		pk[0] = (uint64_t)v[0] | (uint64_t)v[1] << 10 | (uint64_t)v[2] << 20 | (uint64_t)v[3] << 30 | (uint64_t)v[4] << 40 | (uint64_t)v[5] << 50;
		pk[1] = (uint64_t)v[6] | (uint64_t)v[7] << 16 | (uint64_t)v[8] << 32 | (uint64_t)v[9] << 48;
		pk[2] = (uint64_t)v[10] | (uint64_t)v[11] << 16 | (uint64_t)v[12] << 32;
		cta_excl_scan<3, 256, uint64_t>(pk, pt, sm64);
#pragma unroll
		for (int a = 0; a < 6; ++a) { v[a] = (uint32_t)(pk[0] >> (10 * a)) & 0x3ffu; tot[a] = (uint32_t)(pt[0] >> (10 * a)) & 0x3ffu; }
#pragma unroll
		for (int a = 0; a < 4; ++a) { v[6 + a] = (uint32_t)(pk[1] >> (16 * a)) & 0xffffu; tot[6 + a] = (uint32_t)(pt[1] >> (16 * a)) & 0xffffu; }
#pragma unroll
		for (int a = 0; a < 3; ++a) { v[10 + a] = (uint32_t)(pk[2] >> (16 * a)) & 0xffffu; tot[10 + a] = (uint32_t)(pt[2] >> (16 * a)) & 0xffffu; }
	} else cta_excl_scan<NGC, 256, uint32_t>(v, tot, sm);
	if (MODE == 0) {
		if (threadIdx.x == 0) {
#pragma unroll
			for (int k = 0; k < NGC; ++k) A.ctaTot[(size_t)blockIdx.x * NGC + k] = tot[k];
		}
		return;
	}
	if (!valid) return;
#pragma unroll
	for (int k = 0; k < NGC; ++k) v[k] += A.ctaTot[(size_t)blockIdx.x * NGC + k];
	// records of bucket b start at the record prefix of the bucket's first group
	if (sCtl[20] | sCtl[21]) {
		const uint32_t nb = A.ctl->nb, tables = A.ctl->tables;
		for (uint32_t b = 0; b < nb; ++b) if (g == sBkt[b]) {
			A.ctl->recBkt[b] = v[12];
			if (tables) {
#pragma unroll
				for (int a = 0; a < 6; ++a) { A.ctl->grpPre[b * 6 + a] = v[a]; A.ctl->memPre[b * 6 + a] = v[6 + a]; }
			}
		}
	}
	const bool useSizes = A.sizes6 != 0;
	const bool nonempty = useSizes && A.gSize[g] > 0;
	int64_t P = A.gL[g];
	uint32_t r = v[12], mpre = S;
	constexpr int ord[6] = { 0, COMP ? 4 : 1, COMP ? 3 : 2, COMP ? 2 : 3, COMP ? 1 : 4, 5 }; // mrope.c:209-210
#pragma unroll
	for (int slot = 0; slot < 6; ++slot) {
		const int a = ord[slot];
		const int64_t sza = nonempty ? A.sizes6[(size_t)g * 6 + a] : 0;
		if (h[a]) {
			uint32_t dst = NONE32;
			if (a > 0) { // a child that continues: it is a group of the next column (bucket a)
				dst = sCtl[a] + v[a];
				if (useSizes) A.gSizeNext[dst] = sza; // without an old index every interval stays empty and gSize is never read
				A.gOffNext[dst] = sCtl[8 + a] + v[6 + a];
			}
			uint32_t rem = h[a];
			while (rem) { // counts above the 4-byte run limit become several records at the same position
				uint32_t c = rem < RB2_MAXRUN ? rem : RB2_MAXRUN;
				A.recP[r] = P; A.recSC[r] = (c << 3) | (uint32_t)a; A.recDst[r] = dst;
				if (A.recPre) { A.recPre[r] = mpre; mpre += c; }
				dst = NONE32; rem -= c; ++r;
			}
		}
		P += sza; // new symbols of slot a go in front of the old a's of the interval (mrope.c:206-218)
	}
}

// one thread: derive the next column's bucket ranges from the scan totals
__global__ void k_col_bases(Ctl *ctl, uint32_t *gOffNext, uint32_t *recPre, uint32_t M)
{
	uint32_t g = 0, m = 0, bad = 0;
	ctl->gSymBase[0] = 0; ctl->mSymBase[0] = 0;
	for (int a = 1; a <= 6; ++a) {
		ctl->gSymBase[a] = g; ctl->mSymBase[a] = m;
		if (a < 6) { g += ctl->grpTot[a]; m += ctl->grpTot[6 + a]; bad |= ctl->grpTot[6 + a] != ctl->memTot[a]; }
	}
	ctl->gSymBase[7] = g; ctl->mSymBase[7] = m;
	ctl->Gnext = g; ctl->Mnext = m; ctl->nrec = ctl->grpTot[12];
	// buckets whose first group is never visited (empty buckets at the very end) start at the totals
	for (uint32_t b = 0; b < ctl->nb + 2; ++b) ctl->recBkt[b] = ctl->grpTot[12];
	if (ctl->tables)
		for (uint32_t b = 0; b <= ctl->nb; ++b)
			for (int a = 0; a < 6; ++a) { ctl->grpPre[b * 6 + a] = ctl->grpTot[a]; ctl->memPre[b * 6 + a] = ctl->grpTot[6 + a]; }
	gOffNext[g] = m;
	if (recPre) recPre[ctl->grpTot[12]] = M;
	if (bad) ctl->err |= RB2_ERR_ORDER;
}

// ---- all-singleton regime ----------------------------------------------------------------------
// When every group has exactly one member (always in input order, mrope.c:282; in the sorted modes
// as soon as all suffixes read so far are distinct) group index = member index = record index, and
// the next group index of a string is its partition destination.  One kernel then does the work of
// both group passes, their scans and the partition (mrope.c:195-198 is the reference's own
// special case for this situation).
__global__ void k_col_bases_single(Ctl *ctl, uint32_t *gOffNext, uint32_t M, uint32_t *recPre)
{
	uint32_t m = 0;
	ctl->gSymBase[0] = 0; ctl->mSymBase[0] = 0;
	for (int a = 1; a <= 6; ++a) {
		ctl->gSymBase[a] = m; ctl->mSymBase[a] = m;
		if (a < 6) m += ctl->memTot[a];
	}
	ctl->gSymBase[7] = m; ctl->mSymBase[7] = m;
	ctl->Gnext = m; ctl->Mnext = m; ctl->nrec = M;
	for (uint32_t b = 0; b < ctl->nb + 2; ++b) ctl->recBkt[b] = ctl->mBkt[b]; // one record per member, same order
	if (ctl->tables)
		for (uint32_t b = 0; b <= ctl->nb; ++b)
			for (int a = 0; a < 6; ++a) ctl->memPre[b * 6 + a] = ctl->memTot[a];
	gOffNext[m] = m;
	if (recPre) recPre[M] = M;
}

struct SingleArgs {
	const uint32_t *sid; const uint8_t *asym; uint32_t M; const uint32_t *tilePre;
	const int64_t *gL, *gSize, *sizes6; Ctl *ctl;
	uint32_t *sidNext; int64_t *gSizeNext; uint32_t *gOffNext;
	int64_t *recP; uint32_t *recSC, *recDst;
	uint32_t *recPre;
	int lean;   // dense regime, no interval sizes: only sidNext / recDst are written (RecView in rb2_flat.cuh)
};

template <bool COMP>
__global__ void __launch_bounds__(256) k_column_singletons(SingleArgs A)
{
	__shared__ uint32_t sm[6 * 8];
	__shared__ uint32_t sBkt[NBA], sAny;
	const uint32_t k = blockIdx.x * MEM_TILE + threadIdx.x * 4;
	uint32_t a4 = 0, id[4] = { 0, 0, 0, 0 };
	uint32_t c[6] = { 0, 0, 0, 0, 0, 0 }, tot[6];
	if (A.ctl->tables) { // sharded engines: which buckets start inside this tile
		if (threadIdx.x == 0) sAny = 0;
		__syncthreads();
		if (threadIdx.x < A.ctl->nb) {
			const uint32_t mb = A.ctl->mBkt[threadIdx.x];
			const bool in = mb >= blockIdx.x * MEM_TILE && mb < (blockIdx.x + 1) * MEM_TILE && mb < A.M;
			sBkt[threadIdx.x] = in ? mb : NONE32;
			if (in) sAny = 1;
		}
	} else if (threadIdx.x == 0) sAny = 0;
	if (k < A.M) {
		a4 = *reinterpret_cast<const uint32_t*>(A.asym + k);
		for (int i = 0; i < 4; ++i) if (k + i < A.M) {
			if (A.sidNext) id[i] = A.sid[k + i];
			const uint32_t a = (a4 >> (8 * i)) & 0xff;
#pragma unroll
			for (int x = 0; x < 6; ++x) c[x] += a == x;
		}
	}
	cta_excl_scan<6, 256, uint32_t>(c, tot, sm);
	if (k >= A.M) return;
	uint32_t base[6];
#pragma unroll
	for (int x = 0; x < 6; ++x) base[x] = A.ctl->mSymBase[x] + A.tilePre[(size_t)blockIdx.x * 6 + x] + c[x];
	constexpr int ord[6] = { 0, COMP ? 4 : 1, COMP ? 3 : 2, COMP ? 2 : 3, COMP ? 1 : 4, 5 }; // mrope.c:209-210
	const bool marks = sAny != 0; // (written before the scan's barriers)
	for (int i = 0; i < 4; ++i) if (k + i < A.M) {
		const uint32_t g = k + i, a = (a4 >> (8 * i)) & 0xff;
		if (marks) {
			const uint32_t nb = A.ctl->nb;
			for (uint32_t b = 0; b < nb; ++b) if (g == sBkt[b]) {
#pragma unroll
				for (int x = 0; x < 6; ++x) A.ctl->memPre[b * 6 + x] = base[x] - A.ctl->mSymBase[x];
			}
		}
		uint32_t d = NONE32;
#pragma unroll
		for (int x = 1; x < 6; ++x) if (a == x) d = base[x]++;
		if (a == 0) ++base[0]; // only the bucket-start prefixes of a sharded engine look at it
		int64_t P = A.lean ? 0 : A.gL[g], sza = 0;
		if (A.sizes6 && A.gSize[g] > 0) { // insertion point: behind the old symbols of the earlier slots
#pragma unroll
			for (int slot = 0; slot < 6; ++slot) {
				const int64_t z = A.sizes6[(size_t)g * 6 + ord[slot]];
				if ((uint32_t)ord[slot] == a) { sza = z; break; }
				P += z;
			}
		}
		if (a) {
			if (A.sidNext) A.sidNext[d] = id[i]; // (null: the merge kernel delivers the ids, rb2_flat.cuh PeerRoute)
			if (!A.lean) A.gOffNext[d] = d;
			if (A.sizes6) A.gSizeNext[d] = sza;
		}
		A.recDst[g] = d;
		if (!A.lean) {
			A.recP[g] = P; A.recSC[g] = (1u << 3) | a;
			if (A.recPre) A.recPre[g] = g;
		}
	}
}

// =====================================================================================
// Blocks: plan work items, merge records into leaf blocks, rebuild the directory
// =====================================================================================

// bucket of logical block i: the last b with bkt[b] <= i (empty buckets share their start with the
// next non-empty one, which is the one that is found)
__device__ __forceinline__ int bucket_of(const uint32_t *bkt, uint32_t nb, uint32_t i)
{
	if (nb == 6) {
		int b = 0;
#pragma unroll
		for (int x = 1; x < 6; ++x) b += i >= bkt[x];
		return b;
	}
	uint32_t lo = 0, hi = nb - 1;
	while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (bkt[mid] <= i) lo = mid; else hi = mid - 1; }
	return (int)lo;
}

// recHi[i] = #records (global index) at positions <= end of logical block i, inside the
// block's bucket.  A record at position P goes to the block with start < P <= end; P at the
// very start of a bucket goes to the bucket's first block.
__global__ void __launch_bounds__(256) k_rec_hi(Dir dir, uint32_t nlog, const Ctl *ctl, const int64_t *recP, uint32_t *recHi)
{
	const uint32_t i = blockIdx.x * 256 + threadIdx.x;
	if (i >= nlog) return;
	const int b = bucket_of(ctl->blkBkt, ctl->nb, i);
	uint32_t lo = ctl->recBkt[b], hi = ctl->recBkt[b + 1];
	if (i + 1 != ctl->blkBkt[b + 1]) {
		const int64_t key = dir.cumLen[i + 1];
		while (lo < hi) { // first record with P > key
			uint32_t mid = lo + ((hi - lo) >> 1);
			if (recP[mid] <= key) lo = mid + 1; else hi = mid;
		}
	} else lo = hi;
	recHi[i] = lo;
}

struct alignas(16) ItemMeta { // everything a merge warp needs to start, in two 128-bit loads
	uint32_t i, phys, r0, r1;   // logical block, physical block, record range [r0, r1)
	int64_t blkStart;           // global position of the block's first symbol
	uint32_t nIt, sub;          // items of this block, index of this item among them
};

struct ItemScan { // K=1: work items per logical block
	const Ctl *ctl; const uint32_t *recHi; uint32_t nlog;
	uint32_t *itemOff; ItemMeta *itemMeta; Ctl *ctlw;
	const uint32_t *order; const int64_t *cumLen;
	__device__ uint32_t rec_lo(uint32_t i) const {
		const int b = bucket_of(ctl->blkBkt, ctl->nb, i);
		return i == ctl->blkBkt[b] ? ctl->recBkt[b] : recHi[i - 1];
	}
	__device__ void load(uint64_t i, uint32_t (&v)[1]) const {
		uint32_t n = recHi[i] - rec_lo((uint32_t)i);
		v[0] = (n + RMAX - 1) / RMAX;
	}
	__device__ void store(uint64_t i, const uint32_t (&own)[1], const uint32_t (&pre)[1]) const {
		itemOff[i] = pre[0];
		if (own[0]) {
			const uint32_t lo = rec_lo((uint32_t)i), hi = recHi[i], ph = order[i];
			const int64_t bs = cumLen[i];
			for (uint32_t s = 0; s < own[0]; ++s) {
				ItemMeta m;
				m.i = (uint32_t)i; m.phys = ph; m.r0 = lo + s * RMAX; m.r1 = m.r0 + RMAX < hi ? m.r0 + RMAX : hi;
				m.blkStart = bs; m.nIt = own[0]; m.sub = s;
				itemMeta[pre[0] + s] = m;
			}
		}
		if (i + 1 == nlog) { itemOff[nlog] = pre[0] + own[0]; ctlw->nItems = pre[0] + own[0]; }
	}
};

struct MergeArgs {
	uint8_t *pool; uint32_t *blkCnt; Dir dir; uint32_t nlog;
	const uint32_t *recHi, *itemOff; const ItemMeta *itemMeta;
	const int64_t *recP; const uint32_t *recSC, *recDst; // recSC = count << 3 | symbol
	int64_t *gLNext;
	uint32_t *itemPieces, *itemFirst, *itemRest;
	uint32_t *todoA;  // items k_merge_half left for k_merge_fast
	uint32_t *todo;   // items left for k_merge_general
	Ctl *ctl;
};

#define FAST_MAXREC 32  // records per item the edit-based fast path handles (one per lane)

struct GenScratch {                 // general path
	uint32_t lcnt[32 * 7];          // per-lane running old-symbol counts (lane-private, stride 7)
	uint32_t pcl[32 * 7];           // per-lane counts of the piece currently being written
	uint32_t cut[MAXPIECES + 1];
	uint32_t pcnt[MAXPIECES * 6];
};
struct FastScratch {                // edit-based fast path
	uint32_t laneBase[32 * 7];      // per-lane exclusive per-symbol counts
	uint32_t laneEnd[32], laneNr[32], laneRunPre[32], laneFb[32];
	uint32_t eStart[FAST_MAXREC + 1], eEnd[FAST_MAXREC + 1], eNew[FAST_MAXREC + 1], eBuf[FAST_MAXREC + 1];
	int32_t  eCum[FAST_MAXREC + 2];   // exclusive prefix of (new - old) byte deltas
	uint32_t cntAdd[8];               // symbols added by the item's records
	int64_t  cumBase[6];              // cumCnt of the block (per-symbol counts in front of it, all buckets)
};
#define OUT_OFF  0     // fast path: the output block image is assembled at stage[0, 1024)
#define EBUF_OFF 1536  // fast path: replacement bytes of the edits, 16 per record, at stage[1536, 2048)
#define FAST_STAGE 2048
struct RecPre { int64_t P; uint32_t sc, dst; }; // one record, loaded ahead of the block decode

struct alignas(16) FastSmem {       // per warp, k_merge_fast
	uint8_t  img[RB2_IMG_BYTES];    // the input block, byte-addressable (+16 zero bytes)
	uint8_t  stage[FAST_STAGE];     // output image + replacement bytes
	FastScratch f;
};
struct alignas(16) GenSmem {        // per warp, k_merge_general
	uint8_t  img[RB2_IMG_BYTES];
	uint8_t  stage[STAGE_BYTES];    // output run bytes
	GenScratch g;
};

// Sequential merge of one lane's old runs with the records that fall into them.
// EMIT=false: only count output bytes.  EMIT=true: write bytes into `stage` from offset `o`,
// note piece cuts / per-piece counts, and deliver rank(a, P) of every record.
template <bool EMIT>
__device__ __forceinline__ uint32_t lane_merge(const uint8_t *img, uint32_t bp, uint32_t nr, uint32_t pos, uint32_t rlo, uint32_t rhi,
                                               const MergeArgs &A, int64_t blkStart, uint32_t posLo, uint32_t posHi,
                                               const int64_t *cumCntBlk, uint32_t *lc, uint32_t o, uint32_t T,
                                               uint8_t *stage, uint32_t *cut, uint32_t *pcnt, uint32_t *pl)
{
	uint32_t psym = 8, plen = 0, r = rlo;
	uint32_t curPiece = NONE32;
	const uint32_t o0 = o;
	uint32_t nextP = r < rhi ? (uint32_t)(A.recP[r] - blkStart) : 0xffffffffu;

	auto flush = [&]() {
		while (plen) {
			const uint32_t l = plen < RB2_MAXRUN ? plen : RB2_MAXRUN;
			const int nb = run_nbytes(l);
			if (EMIT) {
				enc_run(stage + o, psym, l);
				const uint32_t p = (o + nb - 1) / T;
				if (p && o <= p * T) cut[p] = o;      // this run is the first one of piece p
				if (p != curPiece) {
					if (curPiece != NONE32) {
#pragma unroll
						for (int a = 0; a < 6; ++a) if (pl[a]) { atomicAdd(&pcnt[curPiece * 6 + a], pl[a]); pl[a] = 0; }
					}
					curPiece = p;
				}
				pl[psym] += l;
			}
			o += nb; plen -= l;
		}
	};
	auto emit = [&](uint32_t s, uint32_t l) {
		if (l == 0) return;
		if (s != psym) { flush(); psym = s; }
		plen += l;
	};
	auto emit_old = [&](uint32_t s, uint32_t from, uint32_t to) { // old symbols [from,to), clipped to the item's window
		const uint32_t a = from > posLo ? from : posLo, b = to < posHi ? to : posHi;
		if (b > a) emit(s, b - a);
	};
	auto do_record = [&]() {
		const uint32_t sc = A.recSC[r], a = sc & 7u;
		if (EMIT) {
			const uint32_t dst = A.recDst[r];
			if (dst != NONE32) A.gLNext[dst] = A.ctl->cpost[a] + cumCntBlk[a] + lc[a];
		}
		emit(a, sc >> 3);
		++r;
		nextP = r < rhi ? (uint32_t)(A.recP[r] - blkStart) : 0xffffffffu;
	};

	for (uint32_t q = 0; q < nr; ++q) {
		uint32_t s, rl, nb;
		parse_run(img, bp, s, rl, nb);
		bp += nb;
		const uint32_t end = pos + rl;
		uint32_t cur = pos;
		while (nextP < end) {              // records in front of or inside this run
			if (nextP > cur) { emit_old(s, cur, nextP); lc[s] += nextP - cur; cur = nextP; }
			do_record();
		}
		emit_old(s, cur, end); lc[s] += end - cur;
		pos = end;
	}
	while (r < rhi) do_record();           // records right behind the lane's last run
	flush();
	if (EMIT && curPiece != NONE32) {
#pragma unroll
		for (int a = 0; a < 6; ++a) if (pl[a]) { atomicAdd(&pcnt[curPiece * 6 + a], pl[a]); pl[a] = 0; }
	}
	return o - o0;
}

// Everything a warp knows about its work item after the common prologue.
struct ItemCtx {
	uint32_t w, i, phys, nIt, sub, r0, r1;
	int64_t blkStart;
	const int64_t *cumCntBlk;
	LaneDec d;
	uint32_t basePos, baseCnt[6], blkLen, blkCnt[6], nbytes, pureMask;
	uint4 own;
};

// Allocate nNew fresh leaf blocks for one item (lane 0 asks, everybody gets the answer).
// Returns NONE32 when the pool is exhausted; nothing has been written at that point, so the
// host can grow the pool and run the item again.
__device__ __forceinline__ uint32_t alloc_blocks(const MergeArgs &A, int lane, uint32_t nNew)
{
	uint32_t base = 0;
	if (lane == 0 && nNew) {
		base = atomicAdd(&A.ctl->poolUsed, nNew);
		if (base + nNew > A.ctl->poolCap) { atomicMin(&A.ctl->failBase, base); A.ctl->overflow = 1; base = NONE32; }
	}
	return __shfl_sync(FULLMASK, base, 0);
}

// ---- general path: any number of records, sub-items, empty blocks ---------------------------
__device__ __forceinline__ void merge_general(const MergeArgs &A, GenSmem &S, int lane, const ItemCtx &C)
{
	GenScratch &G = S.g;
	const uint32_t r0 = C.r0, r1 = C.r1;
	const int64_t blkStart = C.blkStart;
	// the slice of old symbols this item re-emits: [posLo, posHi) relative to the block start
	const uint32_t posLo = C.sub == 0 ? 0 : (uint32_t)(A.recP[r0] - blkStart);
	const uint32_t posHi = C.sub + 1 == C.nIt ? C.blkLen : (uint32_t)(A.recP[r1] - blkStart);
	// lane owns the records with position in (c_{lane-1}, c_lane], c = end of the lane's runs capped
	// at posHi; lane 0 also takes position 0.  Lanes that end in front of posLo own nothing (every
	// record of the item is >= posLo), so a record is always ranked by the lane that contains it.
	uint32_t c = C.basePos + C.d.len;
	c = c > posHi ? posHi : c;
	uint32_t lo = r0, hi = r1;
	{
		const int64_t key = blkStart + c;
		while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (A.recP[mid] <= key) lo = mid + 1; else hi = mid; }
	}
	const uint32_t rhi = lane == 31 ? r1 : lo; // the last lane sweeps up whatever is left (positions == posHi)
	uint32_t rlo = __shfl_up_sync(FULLMASK, rhi, 1);
	if (lane == 0) rlo = r0;

	uint32_t *lc = G.lcnt + lane * 7, *pl = G.pcl + lane * 7;
	const uint32_t bp0 = lane * 16 + C.d.fb;
#pragma unroll
	for (int a = 0; a < 6; ++a) { lc[a] = C.baseCnt[a]; pl[a] = 0; }

	// pass 1: output bytes per lane
	const uint32_t myBytes = lane_merge<false>(S.img, bp0, C.d.nr, C.basePos, rlo, rhi, A, blkStart, posLo, posHi, C.cumCntBlk, lc, 0, 1u << 30, S.stage, G.cut, G.pcnt, pl);
	const uint32_t incl = warp_incl_scan(myBytes, lane);
	const uint32_t out = __shfl_sync(FULLMASK, incl, 31);
	if (out > STAGE_BYTES) { if (lane == 0) atomicOr(&A.ctl->err, RB2_ERR_STAGE); return; }
	const uint32_t K = out <= RB2_FILL ? 1 : (out + SPLIT_T - 1) / SPLIT_T;
	const uint32_t T = K == 1 ? (1u << 30) : (out + K - 1) / K;
	const bool inplace = C.nIt == 1;
	const uint32_t nNew = inplace ? K - 1 : K;
	if (K > MAXPIECES) { if (lane == 0) atomicOr(&A.ctl->err, RB2_ERR_PIECES); return; }
	const uint32_t newBase = alloc_blocks(A, lane, nNew);
	if (newBase == NONE32) return;

	for (int k = lane; k < MAXPIECES * 6; k += 32) G.pcnt[k] = 0;
	if (lane <= MAXPIECES) G.cut[lane] = 0;
#pragma unroll
	for (int a = 0; a < 6; ++a) { lc[a] = C.baseCnt[a]; pl[a] = 0; }
	__syncwarp();

	// pass 2: emit bytes, piece cuts, piece counts, ranks
	lane_merge<true>(S.img, bp0, C.d.nr, C.basePos, rlo, rhi, A, blkStart, posLo, posHi, C.cumCntBlk, lc, incl - myBytes, T, S.stage, G.cut, G.pcnt, pl);
	__syncwarp();
	if (lane == 0) G.cut[K] = out;
	__syncwarp();

	// copy the pieces out as leaf blocks: [uint16 nbytes][runs...], zero padded
	const uint32_t firstPhys = inplace ? C.phys : newBase;
	const uint32_t restPhys = inplace ? newBase : newBase + 1;
	for (uint32_t k = 0; k < K; ++k) {
		const uint32_t p = k == 0 ? firstPhys : restPhys + (k - 1);
		const uint32_t st = G.cut[k], n = G.cut[k + 1] - st;
		uint32_t wv[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			uint32_t x = 0;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int bp = lane * 16 + j * 4 + q;
				uint32_t byte;
				if (bp < 2) byte = bp == 0 ? (n & 0xff) : (n >> 8);
				else byte = (uint32_t)(bp - 2) < n ? S.stage[st + bp - 2] : 0;
				x |= byte << (8 * q);
			}
			wv[j] = x;
		}
		*(reinterpret_cast<uint4*>(A.pool + (size_t)p * RB2_BLK) + lane) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
		if (lane < 6) A.blkCnt[(size_t)p * 6 + lane] = G.pcnt[k * 6 + lane];
	}
	if (lane == 0) { A.itemPieces[C.w] = K; A.itemFirst[C.w] = firstPhys; A.itemRest[C.w] = restPhys; }
}

// ---- fast path: one item per block, <= 32 records, non-empty block ---------------------------
// Only the runs a record touches are re-encoded.  Lane j locates record j (target lane by binary
// search over the per-lane end positions, then a walk over that lane's <= 16 runs, parsed from the
// shared-memory block image), which also yields rank(a, P).  Records whose touched runs overlap
// form a group; the group's first lane re-encodes that short span with the records merged in (an
// "edit": old byte range -> replacement bytes).  The output image is assembled by pushing: every
// lane stores its 16 input bytes at their shifted position (bytes inside an edited span are
// dropped), the group heads store their replacement bytes, then each lane reads back 16 aligned
// bytes.  Returns false (no side effects besides idempotent rank writes) if it cannot place a split.
__device__ __forceinline__ bool merge_fast(const MergeArgs &A, FastSmem &S, int lane, const ItemCtx &C, const RecPre &rp)
{
	FastScratch &F = S.f;
	const uint8_t *img = S.img;
	const uint32_t nrec = C.r1 - C.r0, nbytes = C.nbytes, endBp = 2 + nbytes;
	F.laneEnd[lane] = C.basePos + C.d.len;
	F.laneNr[lane] = C.d.nr;
	F.laneFb[lane] = C.d.fb;
	const uint32_t runIncl = warp_incl_scan(C.d.nr, lane);
	F.laneRunPre[lane] = runIncl - C.d.nr;
	const uint32_t nRuns = __shfl_sync(FULLMASK, runIncl, 31);
#pragma unroll
	for (int a = 0; a < 6; ++a) F.laneBase[lane * 7 + a] = C.baseCnt[a];
	if (lane < 8) F.cntAdd[lane] = 0;
	__syncwarp();

	// ---- locate record `lane` ----------------------------------------------------------
	const bool act = (uint32_t)lane < nrec;
	uint32_t P = 0, a = 0, cnt = 0, bpq = 0, off = 0, len = 0, sym = 0, pos = 0, s = 0, e = 0;
	if (act) {
		P = (uint32_t)(rp.P - C.blkStart);
		a = rp.sc & 7u; cnt = rp.sc >> 3;
		uint32_t lo = 0, hi = 31;
		while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (F.laneEnd[mid] >= P) hi = mid; else lo = mid + 1; }
		const uint32_t t = lo;
		pos = t ? F.laneEnd[t - 1] : 0;
		uint32_t ca = F.laneBase[t * 7 + a];
		const uint32_t nrt = F.laneNr[t];
		uint32_t q = 0, nb;
		if (((C.pureMask >> t) & 1u) && P > 0) {
			// target lane holds only 1-byte runs: find the run with 4-byte SIMD steps.  Lengths are
			// (byte >> 3); a multiply by 0x01010101 gives the inclusive prefix inside a word.
			const uint32_t *tw = reinterpret_cast<const uint32_t*>(img) + t * 4;
			const uint32_t prel = P - pos; // 1 .. symbols in the lane
			const uint32_t tlo = a < 4 ? 1u << (8 * a) : 0u, thi = a >= 4 ? 1u << (8 * (a - 4)) : 0u;
			uint32_t acc = 0, idx = 0;
#pragma unroll 1
			for (int j = 0; j < 4; ++j) { // rolled on purpose (instruction footprint)
				uint32_t wj = tw[j];
				if (t == 0 && j == 0) wj &= 0xffff0000u;
				const uint32_t lens = (wj >> 3) & 0x0f0f0f0fu;
				const uint32_t sy = wj & 0x07070707u;
				const uint32_t tt = sy | (sy >> 4);
				const uint32_t wt = __byte_perm(tlo, thi, (tt & 0xffu) | ((tt >> 8) & 0xff00u)); // 1 where symbol == a
				const uint32_t wsum = __dp4a(lens, 0x01010101u, 0u);
				if (acc + wsum >= prel) {
					const uint32_t pre = lens * 0x01010101u;
					int i = 0;
#pragma unroll
					for (int x = 2; x >= 0; --x) if (acc + ((pre >> (8 * x)) & 0xffu) < prel) { i = x + 1; break; }
					idx = 4 * j + i;
					len = (lens >> (8 * i)) & 0xffu; sym = (sy >> (8 * i)) & 7u;
					const uint32_t before = i ? (1u << (8 * i)) - 1u : 0u;
					ca += __dp4a(lens & before, wt, 0u);
					pos += acc + ((pre >> (8 * i)) & 0xffu) - len;
					break;
				}
				acc += wsum; ca += __dp4a(lens, wt, 0u);
			}
			q = idx - F.laneFb[t];
			bpq = t * 16 + idx;
		} else {
			bpq = t * 16 + F.laneFb[t];
			for (;; ++q) { // run that contains symbol P-1 (the first run for P == 0)
				parse_run(img, bpq, sym, len, nb);
				if (pos + len >= P || q + 1 >= nrt) break;
				ca += sym == a ? len : 0;
				pos += len; bpq += nb;
			}
		}
		off = P - pos;
		if (rp.dst != NONE32) A.gLNext[rp.dst] = A.ctl->cpost[a] + F.cumBase[a] + ca + (sym == a ? off : 0);
		atomicAdd(&F.cntAdd[a], cnt);
		s = F.laneRunPre[t] + q;                                   // global index of that run
		e = s + ((off == len && s + 1 < nRuns) ? 1 : 0);           // at a run boundary the next run may absorb the record
	}
	// ---- group records whose touched runs overlap ----------------------------------------------
	const uint32_t ePrev = __shfl_up_sync(FULLMASK, e, 1);
	const bool head = act && (lane == 0 || s > ePrev);
	const uint32_t H = __ballot_sync(FULLMASK, head);
	const uint32_t ng = __popc(H);
	const uint32_t above = lane == 31 ? 0u : (H & ~((2u << lane) - 1u));
	const uint32_t gend = above ? (uint32_t)(__ffs(above) - 1) : nrec;  // one past the group's last record
	const uint32_t eLast = __shfl_sync(FULLMASK, e, (gend - 1) & 31);
	const uint32_t k = __popc(H & ((1u << lane) - 1u));
	// ---- group heads re-encode their span ------------------------------------------------
	if (head) {
		uint8_t *ob = S.stage + EBUF_OFF + lane * 16;
		uint32_t o = 0, psym = 8, plen = 0;
		uint32_t rr_ = C.r0 + lane;
		const uint32_t rend = C.r0 + gend;
		uint32_t nextP = P, na = a, nc = cnt;
		uint32_t bp = bpq, p0 = pos;
		auto flush = [&]() {
			while (plen) {
				const uint32_t l = plen < RB2_MAXRUN ? plen : RB2_MAXRUN;
				o += enc_run(ob + o, psym, l);
				plen -= l;
			}
		};
		auto emit = [&](uint32_t sy, uint32_t l) {
			if (l == 0) return;
			if (sy != psym) { flush(); psym = sy; }
			plen += l;
		};
		auto next_rec = [&]() {
			++rr_;
			if (rr_ < rend) { nextP = (uint32_t)(A.recP[rr_] - C.blkStart); const uint32_t sc = A.recSC[rr_]; na = sc & 7u; nc = sc >> 3; }
		};
		for (uint32_t g = s; g <= eLast; ++g) {
			uint32_t sy, rl, nb;
			parse_run(img, bp, sy, rl, nb);
			bp += nb;
			const uint32_t end = p0 + rl;
			uint32_t cur = p0;
			while (rr_ < rend && nextP < end) {
				if (nextP > cur) { emit(sy, nextP - cur); cur = nextP; }
				emit(na, nc);
				next_rec();
			}
			emit(sy, end - cur);
			p0 = end;
		}
		while (rr_ < rend) { emit(na, nc); next_rec(); }
		flush();
		F.eStart[k] = bpq;   // the span's first byte ...
		F.eEnd[k] = bp;      // ... and one past its last byte in the input image
		F.eNew[k] = o;
		F.eBuf[k] = EBUF_OFF + lane * 16;
	}
	__syncwarp();
	// ---- output geometry ---------------------------------------------------------------
	int32_t delta = 0;
	if ((uint32_t)lane < ng) delta = (int32_t)F.eNew[lane] - (int32_t)(F.eEnd[lane] - F.eStart[lane]);
	const int32_t dIncl = warp_incl_scan(delta, lane);
	if ((uint32_t)lane < ng) F.eCum[lane] = dIncl - delta;
	const int32_t totalDelta = __shfl_sync(FULLMASK, dIncl, 31);
	if (lane == 0) F.eCum[ng] = totalDelta;
	__syncwarp();
	const uint32_t outEnd = endBp + totalDelta;   // end of the output image (header included)
	const uint32_t outBytes = outEnd - 2;
	// edits in front of this lane's input window, and whether the window is one verbatim stretch
	const uint32_t bp0 = lane * 16;
	uint32_t kA = 0;
	{ uint32_t lo = 0, hi = ng; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (F.eStart[mid] <= bp0) lo = mid + 1; else hi = mid; } kA = lo; }
	const bool insideFirst = kA && bp0 < F.eEnd[kA - 1];
	const bool simple = !insideFirst && (kA >= ng || F.eStart[kA] >= bp0 + 16);
	uint32_t K = 1, cutImg = outEnd, cutLane = 32;
	if (outBytes > RB2_FILL) {
		// split in two at the first run of some lane (never inside an edited span)
		K = 2;
		uint32_t score = 0xffffffffu;
		if (C.d.nr) {
			const uint32_t bpF = bp0 + C.d.fb;
			uint32_t kk = kA;
			while (kk < ng && F.eStart[kk] <= bpF) ++kk;
			const bool inside = kk && bpF < F.eEnd[kk - 1];
			const uint32_t cand = bpF + F.eCum[kk];
			if (!inside && cand > 2 && cand - 2 <= RB2_FILL && outEnd - cand <= RB2_FILL && cand < outEnd) {
				const uint32_t mid = 2 + outBytes / 2;
				score = ((cand > mid ? cand - mid : mid - cand) << 5) | lane;
				cutImg = cand;
			}
		}
		uint32_t best = score;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { const uint32_t y = __shfl_xor_sync(FULLMASK, best, o); best = y < best ? y : best; }
		if (best == 0xffffffffu) return false;
		cutLane = best & 31;
		cutImg = __shfl_sync(FULLMASK, cutImg, cutLane);
	}
	const uint32_t newBase = alloc_blocks(A, lane, K - 1);
	if (newBase == NONE32) return true; // pool exhausted: item stays unmerged, the host retries

	// ---- new per-block symbol counts ------------------------------------------------------
	if (K == 1) {
		if (lane < 6) {
			uint32_t v = 0;
#pragma unroll
			for (int x = 0; x < 6; ++x) if (lane == x) v = C.blkCnt[x];
			A.blkCnt[(size_t)C.phys * 6 + lane] = v + F.cntAdd[lane];
		}
	} else {
		const uint32_t cutPos = __shfl_sync(FULLMASK, C.basePos, cutLane);
		uint32_t v0 = 0, v1 = 0;
#pragma unroll
		for (int x = 0; x < 6; ++x) {
			const uint32_t mine = act && a == (uint32_t)x ? cnt : 0;
			const uint32_t add0 = warp_sum(act && P < cutPos ? mine : 0u);
			const uint32_t before = __shfl_sync(FULLMASK, C.baseCnt[x], cutLane);
			if (lane == x) { v0 = before + add0; v1 = C.blkCnt[x] - before + F.cntAdd[x] - add0; }
		}
		if (lane < 6) { A.blkCnt[(size_t)C.phys * 6 + lane] = v0; A.blkCnt[(size_t)newBase * 6 + lane] = v1; }
	}
	// ---- assemble the output image: push input bytes and replacement bytes to their place ----------
	uint8_t *out = S.stage + OUT_OFF;
	{
		const uint32_t w4[4] = { C.own.x, C.own.y, C.own.z, C.own.w };
		if (simple) {
			const uint32_t d0 = bp0 + (uint32_t)F.eCum[kA];
#pragma unroll
			for (int i = 0; i < 16; ++i) out[d0 + i] = (uint8_t)(w4[i >> 2] >> ((i & 3) * 8));
		} else {
			// walk the window with the current edit's bounds cached in registers
			uint32_t kk = kA;
			uint32_t curEnd = kk ? F.eEnd[kk - 1] : 0;                  // bytes below this belong to an edited span
			uint32_t nextStart = kk < ng ? F.eStart[kk] : 0xffffffffu;  // first byte of the next edited span
			uint32_t cum = (uint32_t)F.eCum[kk];
#pragma unroll 1
			for (int i = 0; i < 16; ++i) { // rolled on purpose (instruction footprint); bytes come from the image
				const uint32_t bp = bp0 + i;
				if (bp >= nextStart) { // distinct starts: at most one edit begins per byte
					++kk;
					curEnd = F.eEnd[kk - 1];
					nextStart = kk < ng ? F.eStart[kk] : 0xffffffffu;
					cum = (uint32_t)F.eCum[kk];
				}
				if (bp >= curEnd) out[bp + cum] = img[bp];
			}
		}
		if (head) { // my group's replacement bytes
			const uint32_t d0 = F.eStart[k] + (uint32_t)F.eCum[k], n = F.eNew[k];
			const uint8_t *src = S.stage + EBUF_OFF + lane * 16;
			for (uint32_t i = 0; i < n; ++i) out[d0 + i] = src[i];
		}
	}
	// The zero bytes behind the input's last run were pushed too, so the image is zero up to byte
	// 512 + totalDelta; only a shrinking block leaves a few stale bytes in front of byte 512.
	if (totalDelta < 0 && lane == 31) for (int32_t i = totalDelta; i < 0; ++i) out[RB2_BLK + i] = 0;
	__syncwarp();
	auto tail_mask = [&](uint4 v, uint32_t base, uint32_t end) -> uint4 { // zero the bytes at image index >= end
		if (base + 16 <= end) return v;
		uint32_t wv[4] = { v.x, v.y, v.z, v.w };
		const uint32_t keep = end > base ? end - base : 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t kb = keep > (uint32_t)j * 4 ? keep - j * 4 : 0;
			wv[j] = kb >= 4 ? wv[j] : (kb ? wv[j] & ((1u << (kb * 8)) - 1u) : 0u);
		}
		return make_uint4(wv[0], wv[1], wv[2], wv[3]);
	};
	const uint32_t end0 = K == 1 ? outEnd : cutImg;
	uint4 v0 = reinterpret_cast<const uint4*>(out)[lane];
	if (K == 2) v0 = tail_mask(v0, bp0, end0);
	if (lane == 0) v0.x = (v0.x & 0xffff0000u) | (end0 - 2);
	*(reinterpret_cast<uint4*>(A.pool + (size_t)C.phys * RB2_BLK) + lane) = v0;
	if (K == 2) { // second piece: image bytes [cutImg, outEnd) behind a fresh header
		const uint32_t src = OUT_OFF + cutImg - 2 + bp0;
		const uint32_t *wp = reinterpret_cast<const uint32_t*>(S.stage) + (src >> 2);
		const uint32_t sh = (src & 3) * 8;
		const uint32_t a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
		uint4 v1 = make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh));
		v1 = tail_mask(v1, cutImg - 2 + bp0, outEnd);
		if (lane == 0) v1.x = (v1.x & 0xffff0000u) | (outEnd - cutImg);
		*(reinterpret_cast<uint4*>(A.pool + (size_t)newBase * RB2_BLK) + lane) = v1;
	}
	if (lane == 0) { A.itemPieces[C.w] = K; A.itemFirst[C.w] = C.phys; A.itemRest[C.w] = newBase; }
	return true;
}

// ---- half-warp path: two items per warp, 16 lanes x 32 bytes each -------------------------------
// Same algorithm as merge_fast, for items with <= 16 records (the bulk of all block visits).  The
// per-record sections of merge_fast keep only the few lanes that hold a record busy; with two
// items per warp every instruction of those sections serves both.  All collectives use the
// half's own lane mask, so the two halves may diverge freely.
#define HALF_MAXREC 16
#define HALF_OUT    1024  // output image at stage[0, 1024)
struct HalfScratch {
	uint32_t laneBase[16 * 7];
	uint32_t laneEnd[16], laneNr[16], laneRunPre[16], laneFb[16];
	uint32_t eStart[HALF_MAXREC + 1], eEnd[HALF_MAXREC + 1], eNew[HALF_MAXREC + 1];
	int32_t  eCum[HALF_MAXREC + 2];
	uint32_t cntAdd[8];
	int64_t  cumBase[6];
};
struct alignas(16) HalfSmem {
	uint8_t img[RB2_IMG_BYTES];
	uint8_t stage[HALF_OUT + HALF_MAXREC * 16];
	HalfScratch h;
};

template <typename T>
__device__ __forceinline__ T half_incl_scan(T v, int hl, uint32_t hmask)
{
#pragma unroll
	for (int o = 1; o < 16; o <<= 1) {
		T y = __shfl_up_sync(hmask, v, o, 16);
		if (hl >= o) v += y;
	}
	return v;
}

__global__ void __launch_bounds__(MERGE_WARPS * 32, HALF_MINCTA) k_merge_half(MergeArgs A)
{
	RB2_DYN_SMEM(smraw);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
	const uint32_t hmask = half ? 0xffff0000u : 0x0000ffffu;
	const uint32_t w = (blockIdx.x * MERGE_WARPS + wid) * 2 + half;
	if (w >= A.ctl->nItems) return;
	if (A.itemPieces[w] != 0) return; // merged by an earlier launch (retry after pool growth)
	HalfSmem &S = reinterpret_cast<HalfSmem*>(smraw)[wid * 2 + half];
	HalfScratch &F = S.h;
	const ItemMeta m = A.itemMeta[w];
	const uint32_t nrec = m.r1 - m.r0;
	auto defer = [&]() { if (hl == 0) A.todoA[atomicAdd(&A.ctl->nTodoA, 1u)] = w; };
	if (m.nIt != 1 || nrec > HALF_MAXREC) { defer(); return; }

	// ---- loads: my record, the block's directory counts, the block itself --------------------
	const bool act = (uint32_t)hl < nrec;
	int64_t recP = 0; uint32_t recSC = 0, recDst = NONE32;
	if (act) { const uint32_t r = m.r0 + hl; recP = A.recP[r]; recSC = A.recSC[r]; recDst = A.recDst[r]; }
	if (hl < 6) F.cumBase[hl] = A.dir.cumCnt[(size_t)m.i * 6 + hl];
	const uint4 *blk = reinterpret_cast<const uint4*>(A.pool + (size_t)m.phys * RB2_BLK);
	const uint4 o0 = blk[hl * 2], o1 = blk[hl * 2 + 1];
	uint8_t *img = S.img;
	reinterpret_cast<uint4*>(img)[hl * 2] = o0;
	reinterpret_cast<uint4*>(img)[hl * 2 + 1] = o1;
	if (hl == 0) reinterpret_cast<uint4*>(img)[32] = make_uint4(0, 0, 0, 0);
	if (hl < 8) F.cntAdd[hl] = 0;
	const uint32_t nbytes = __shfl_sync(hmask, o0.x, 0, 16) & 0xffffu, endBp = 2 + nbytes;
	__syncwarp(hmask);
	if (nbytes == 0) { defer(); return; }

	// ---- decode: 32 bytes per lane ---------------------------------------------------------
	uint32_t wv[8] = { o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w };
	if (hl == 0) wv[0] &= 0xffff0000u;
	const int lim = (int)nbytes + 2 - hl * 32;
	const uint32_t vhi = lim > 32 ? 32u : (lim < 0 ? 0u : (uint32_t)lim), vlo = hl == 0 ? 2u : 0u;
	uint32_t any = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) any |= wv[j];
	const bool pure = (any & 0x80808080u) == 0 || lim <= 0;
	const uint32_t pureMask = (__ballot_sync(hmask, pure) >> (16 * half)) & 0xffffu;
	uint32_t dnr, dlen, dfb, dc[6] = { 0, 0, 0, 0, 0, 0 };
	if (pure) {
		if (vhi < 32) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const int keep = (int)vhi - 4 * j;
				wv[j] = keep >= 4 ? wv[j] : (keep <= 0 ? 0u : wv[j] & ((1u << (8 * keep)) - 1u));
			}
		}
		dnr = vhi > vlo ? vhi - vlo : 0; dfb = vhi > vlo ? vlo : 32u; dlen = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t lens = (wv[j] >> 3) & 0x0f0f0f0fu;
			const uint32_t sy = wv[j] & 0x07070707u;
			const uint32_t t = sy | (sy >> 4);
			const uint32_t sel = (t & 0xffu) | ((t >> 8) & 0xff00u);
			dlen = __dp4a(lens, 0x01010101u, dlen);
			dc[0] = __dp4a(lens, __byte_perm(0x00000001u, 0u, sel), dc[0]);
			dc[1] = __dp4a(lens, __byte_perm(0x00000100u, 0u, sel), dc[1]);
			dc[2] = __dp4a(lens, __byte_perm(0x00010000u, 0u, sel), dc[2]);
			dc[3] = __dp4a(lens, __byte_perm(0x01000000u, 0u, sel), dc[3]);
			dc[4] = __dp4a(lens, __byte_perm(0u, 0x00000001u, sel), dc[4]);
			dc[5] = __dp4a(lens, __byte_perm(0u, 0x00000100u, sel), dc[5]);
		}
	} else {
		uint32_t *lcs = F.laneBase + hl * 7;
		const uint64_t r = decode_span_serial(img, hl * 32, vlo, vhi, 32u, lcs);
		dnr = (uint32_t)r & 0xffu; dfb = ((uint32_t)r >> 8) & 0xffu; dlen = (uint32_t)(r >> 32);
		if (((uint32_t)r >> 16) & 0xffu) atomicOr(&A.ctl->err, ((uint32_t)r >> 16) & 0xffu);
#pragma unroll
		for (int a = 0; a < 6; ++a) dc[a] = lcs[a];
	}
	const uint32_t lenIncl = half_incl_scan(dlen, hl, hmask);
	const uint32_t basePos = lenIncl - dlen, blkLen = __shfl_sync(hmask, lenIncl, 15, 16);
	if (blkLen >= 65536u) { defer(); return; } // counts would not fit the packed scans: full-warp kernel
	uint32_t baseCnt[6], blkCnt[6];
#pragma unroll
	for (int a = 0; a < 6; a += 2) {
		const uint32_t v = dc[a] | (dc[a + 1] << 16);
		const uint32_t x = half_incl_scan(v, hl, hmask);
		const uint32_t ex = x - v, t = __shfl_sync(hmask, x, 15, 16);
		baseCnt[a] = ex & 0xffffu; baseCnt[a + 1] = ex >> 16;
		blkCnt[a] = t & 0xffffu; blkCnt[a + 1] = t >> 16;
	}
	const uint32_t runIncl = half_incl_scan(dnr, hl, hmask);
	const uint32_t nRuns = __shfl_sync(hmask, runIncl, 15, 16);
	__syncwarp(hmask); // serial lanes are done with their laneBase scratch
	F.laneEnd[hl] = basePos + dlen; F.laneNr[hl] = dnr; F.laneFb[hl] = dfb; F.laneRunPre[hl] = runIncl - dnr;
#pragma unroll
	for (int a = 0; a < 6; ++a) F.laneBase[hl * 7 + a] = baseCnt[a];
	__syncwarp(hmask);

	// ---- locate record `hl` ----------------------------------------------------------------
	uint32_t P = 0, a = 0, cnt = 0, bpq = 0, off = 0, len = 0, sym = 0, pos = 0, s = 0, e = 0;
	if (act) {
		P = (uint32_t)(recP - m.blkStart);
		a = recSC & 7u; cnt = recSC >> 3;
		uint32_t lo = 0, hi = 15;
		while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (F.laneEnd[mid] >= P) hi = mid; else lo = mid + 1; }
		const uint32_t t = lo;
		pos = t ? F.laneEnd[t - 1] : 0;
		uint32_t ca = F.laneBase[t * 7 + a];
		const uint32_t nrt = F.laneNr[t];
		uint32_t q = 0, nb;
		if (((pureMask >> t) & 1u) && P > 0) {
			const uint32_t *tw = reinterpret_cast<const uint32_t*>(img) + t * 8;
			const uint32_t prel = P - pos;
			const uint32_t tlo = a < 4 ? 1u << (8 * a) : 0u, thi = a >= 4 ? 1u << (8 * (a - 4)) : 0u;
			uint32_t acc = 0, idx = 0;
#pragma unroll 1
			for (int j = 0; j < 8; ++j) {
				uint32_t wj = tw[j];
				if (t == 0 && j == 0) wj &= 0xffff0000u;
				const uint32_t lens = (wj >> 3) & 0x0f0f0f0fu;
				const uint32_t sy = wj & 0x07070707u;
				const uint32_t tt = sy | (sy >> 4);
				const uint32_t wt = __byte_perm(tlo, thi, (tt & 0xffu) | ((tt >> 8) & 0xff00u));
				const uint32_t wsum = __dp4a(lens, 0x01010101u, 0u);
				if (acc + wsum >= prel) {
					const uint32_t pre = lens * 0x01010101u;
					int i = 0;
#pragma unroll
					for (int x = 2; x >= 0; --x) if (acc + ((pre >> (8 * x)) & 0xffu) < prel) { i = x + 1; break; }
					idx = 4 * j + i;
					len = (lens >> (8 * i)) & 0xffu; sym = (sy >> (8 * i)) & 7u;
					const uint32_t before = i ? (1u << (8 * i)) - 1u : 0u;
					ca += __dp4a(lens & before, wt, 0u);
					pos += acc + ((pre >> (8 * i)) & 0xffu) - len;
					break;
				}
				acc += wsum; ca += __dp4a(lens, wt, 0u);
			}
			q = idx - F.laneFb[t];
			bpq = t * 32 + idx;
		} else {
			bpq = t * 32 + F.laneFb[t];
			for (;; ++q) {
				parse_run(img, bpq, sym, len, nb);
				if (pos + len >= P || q + 1 >= nrt) break;
				ca += sym == a ? len : 0;
				pos += len; bpq += nb;
			}
		}
		off = P - pos;
		if (recDst != NONE32) A.gLNext[recDst] = A.ctl->cpost[a] + F.cumBase[a] + ca + (sym == a ? off : 0);
		atomicAdd(&F.cntAdd[a], cnt);
		s = F.laneRunPre[t] + q;
		e = s + ((off == len && s + 1 < nRuns) ? 1 : 0);
	}
	// ---- groups of records with overlapping spans -----------------------------------------------
	const uint32_t ePrev = __shfl_up_sync(hmask, e, 1, 16);
	const bool head = act && (hl == 0 || s > ePrev);
	const uint32_t H = (__ballot_sync(hmask, head) >> (16 * half)) & 0xffffu;
	const uint32_t ng = __popc(H);
	const uint32_t above = H & ~((2u << hl) - 1u);
	const uint32_t gend = above ? (uint32_t)(__ffs(above) - 1) : nrec;
	const uint32_t eLast = __shfl_sync(hmask, e, (gend - 1) & 15, 16);
	const uint32_t k = __popc(H & ((1u << hl) - 1u));
	uint8_t *ebuf = S.stage + HALF_OUT + hl * 16;
	if (head) {
		uint32_t o = 0, psym = 8, plen = 0;
		uint32_t rr_ = m.r0 + hl;
		const uint32_t rend = m.r0 + gend;
		uint32_t nextP = P, na = a, nc = cnt;
		uint32_t bp = bpq, p0 = pos;
		auto flush = [&]() {
			while (plen) {
				const uint32_t l = plen < RB2_MAXRUN ? plen : RB2_MAXRUN;
				o += enc_run(ebuf + o, psym, l);
				plen -= l;
			}
		};
		auto emit = [&](uint32_t sy, uint32_t l) {
			if (l == 0) return;
			if (sy != psym) { flush(); psym = sy; }
			plen += l;
		};
		auto next_rec = [&]() {
			++rr_;
			if (rr_ < rend) { nextP = (uint32_t)(A.recP[rr_] - m.blkStart); const uint32_t sc = A.recSC[rr_]; na = sc & 7u; nc = sc >> 3; }
		};
		for (uint32_t g = s; g <= eLast; ++g) {
			uint32_t sy, rl, nb;
			parse_run(img, bp, sy, rl, nb);
			bp += nb;
			const uint32_t end = p0 + rl;
			uint32_t cur = p0;
			while (rr_ < rend && nextP < end) {
				if (nextP > cur) { emit(sy, nextP - cur); cur = nextP; }
				emit(na, nc);
				next_rec();
			}
			emit(sy, end - cur);
			p0 = end;
		}
		while (rr_ < rend) { emit(na, nc); next_rec(); }
		flush();
		F.eStart[k] = bpq; F.eEnd[k] = bp; F.eNew[k] = o;
	}
	__syncwarp(hmask);
	// ---- output geometry ---------------------------------------------------------------
	int32_t delta = 0;
	if ((uint32_t)hl < ng) delta = (int32_t)F.eNew[hl] - (int32_t)(F.eEnd[hl] - F.eStart[hl]);
	const int32_t dIncl = half_incl_scan(delta, hl, hmask);
	if ((uint32_t)hl < ng) F.eCum[hl] = dIncl - delta;
	const int32_t totalDelta = __shfl_sync(hmask, dIncl, 15, 16);
	if (hl == 0) F.eCum[ng] = totalDelta;
	__syncwarp(hmask);
	const uint32_t outEnd = endBp + totalDelta, outBytes = outEnd - 2;
	const uint32_t bp0 = hl * 32;
	uint32_t kA = 0;
	{ uint32_t lo = 0, hi = ng; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (F.eStart[mid] <= bp0) lo = mid + 1; else hi = mid; } kA = lo; }
	const bool insideFirst = kA && bp0 < F.eEnd[kA - 1];
	const bool simple = !insideFirst && (kA >= ng || F.eStart[kA] >= bp0 + 32);
	uint32_t K = 1, cutImg = outEnd, cutLane = 16;
	if (outBytes > RB2_FILL) {
		K = 2;
		uint32_t score = 0xffffffffu;
		if (dnr) {
			const uint32_t bpF = bp0 + dfb;
			uint32_t kk = kA;
			while (kk < ng && F.eStart[kk] <= bpF) ++kk;
			const bool inside = kk && bpF < F.eEnd[kk - 1];
			const uint32_t cand = bpF + F.eCum[kk];
			if (!inside && cand > 2 && cand - 2 <= RB2_FILL && outEnd - cand <= RB2_FILL && cand < outEnd) {
				const uint32_t mid = 2 + outBytes / 2;
				score = ((cand > mid ? cand - mid : mid - cand) << 4) | hl;
				cutImg = cand;
			}
		}
		uint32_t best = score;
#pragma unroll
		for (int o = 8; o > 0; o >>= 1) { const uint32_t y = __shfl_xor_sync(hmask, best, o, 16); best = y < best ? y : best; }
		if (best == 0xffffffffu) { defer(); return; } // ranks already written are idempotent
		cutLane = best & 15;
		cutImg = __shfl_sync(hmask, cutImg, cutLane, 16);
	}
	uint32_t newBase = 0;
	if (K == 2) {
		if (hl == 0) {
			newBase = atomicAdd(&A.ctl->poolUsed, 1u);
			if (newBase + 1 > A.ctl->poolCap) { atomicMin(&A.ctl->failBase, newBase); A.ctl->overflow = 1; newBase = NONE32; }
		}
		newBase = __shfl_sync(hmask, newBase, 0, 16);
		if (newBase == NONE32) return; // pool exhausted: item stays unmerged, the host retries
	}
	// ---- per-block symbol counts -------------------------------------------------------------
	if (K == 1) {
		if (hl < 6) {
			uint32_t v = 0;
#pragma unroll
			for (int x = 0; x < 6; ++x) if (hl == x) v = blkCnt[x];
			A.blkCnt[(size_t)m.phys * 6 + hl] = v + F.cntAdd[hl];
		}
	} else {
		const uint32_t cutPos = __shfl_sync(hmask, basePos, cutLane, 16);
		uint32_t v0 = 0, v1 = 0;
#pragma unroll
		for (int x = 0; x < 6; ++x) {
			uint32_t add0 = act && a == (uint32_t)x && P < cutPos ? cnt : 0u;
#pragma unroll
			for (int o = 8; o > 0; o >>= 1) add0 += __shfl_xor_sync(hmask, add0, o, 16);
			const uint32_t before = __shfl_sync(hmask, baseCnt[x], cutLane, 16);
			if (hl == x) { v0 = before + add0; v1 = blkCnt[x] - before + F.cntAdd[x] - add0; }
		}
		if (hl < 6) { A.blkCnt[(size_t)m.phys * 6 + hl] = v0; A.blkCnt[(size_t)newBase * 6 + hl] = v1; }
	}
	// ---- assemble: push input bytes and replacement bytes -----------------------------------------
	uint8_t *out = S.stage;
	if (simple) {
		const uint32_t d0 = bp0 + (uint32_t)F.eCum[kA];
		const uint32_t ow[8] = { o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w };
#pragma unroll
		for (int i = 0; i < 32; ++i) out[d0 + i] = (uint8_t)(ow[i >> 2] >> ((i & 3) * 8));
	} else {
		uint32_t kk = kA;
		uint32_t curEnd = kk ? F.eEnd[kk - 1] : 0;
		uint32_t nextStart = kk < ng ? F.eStart[kk] : 0xffffffffu;
		uint32_t cum = (uint32_t)F.eCum[kk];
#pragma unroll 1
		for (int i = 0; i < 32; ++i) {
			const uint32_t bp = bp0 + i;
			if (bp >= nextStart) {
				++kk;
				curEnd = F.eEnd[kk - 1];
				nextStart = kk < ng ? F.eStart[kk] : 0xffffffffu;
				cum = (uint32_t)F.eCum[kk];
			}
			if (bp >= curEnd) out[bp + cum] = img[bp];
		}
	}
	if (head) {
		const uint32_t d0 = F.eStart[k] + (uint32_t)F.eCum[k], n = F.eNew[k];
		for (uint32_t i = 0; i < n; ++i) out[d0 + i] = ebuf[i];
	}
	if (totalDelta < 0 && hl == 15) for (int32_t i = totalDelta; i < 0; ++i) out[RB2_BLK + i] = 0;
	__syncwarp(hmask);
	auto tail_mask = [&](uint4 v, uint32_t base, uint32_t end) -> uint4 {
		if (base + 16 <= end) return v;
		uint32_t x[4] = { v.x, v.y, v.z, v.w };
		const uint32_t keep = end > base ? end - base : 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t kb = keep > (uint32_t)j * 4 ? keep - j * 4 : 0;
			x[j] = kb >= 4 ? x[j] : (kb ? x[j] & ((1u << (kb * 8)) - 1u) : 0u);
		}
		return make_uint4(x[0], x[1], x[2], x[3]);
	};
	const uint32_t end0 = K == 1 ? outEnd : cutImg;
	uint4 a0 = reinterpret_cast<const uint4*>(out)[hl * 2], a1 = reinterpret_cast<const uint4*>(out)[hl * 2 + 1];
	if (K == 2) { a0 = tail_mask(a0, bp0, end0); a1 = tail_mask(a1, bp0 + 16, end0); }
	if (hl == 0) a0.x = (a0.x & 0xffff0000u) | (end0 - 2);
	uint4 *dst0 = reinterpret_cast<uint4*>(A.pool + (size_t)m.phys * RB2_BLK);
	dst0[hl * 2] = a0; dst0[hl * 2 + 1] = a1;
	if (K == 2) {
		auto fetch = [&](uint32_t src) -> uint4 {
			const uint32_t *wp = reinterpret_cast<const uint32_t*>(S.stage) + (src >> 2);
			const uint32_t sh = (src & 3) * 8;
			const uint32_t x0 = wp[0], x1 = wp[1], x2 = wp[2], x3 = wp[3], x4 = wp[4];
			return make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
		};
		const uint32_t b1 = cutImg - 2 + bp0;
		uint4 c0 = tail_mask(fetch(b1), b1, outEnd), c1 = tail_mask(fetch(b1 + 16), b1 + 16, outEnd);
		if (hl == 0) c0.x = (c0.x & 0xffff0000u) | (outEnd - cutImg);
		uint4 *dst1 = reinterpret_cast<uint4*>(A.pool + (size_t)newBase * RB2_BLK);
		dst1[hl * 2] = c0; dst1[hl * 2 + 1] = c1;
	}
	if (hl == 0) { A.itemPieces[w] = K; A.itemFirst[w] = m.phys; A.itemRest[w] = newBase; }
}

// Common prologue of both merge kernels: which block, which records, decode the block.
__device__ __forceinline__ void item_prologue(const MergeArgs &A, ItemCtx &C, int lane, uint8_t *img, uint32_t *cntScratch)
{
	const ItemMeta m = A.itemMeta[C.w];
	C.i = m.i; C.phys = m.phys; C.r0 = m.r0; C.r1 = m.r1; C.blkStart = m.blkStart; C.nIt = m.nIt; C.sub = m.sub;
	C.cumCntBlk = A.dir.cumCnt + (size_t)C.i * 6;
	uint32_t err = 0;
	warp_decode_block(A.pool + (size_t)C.phys * RB2_BLK, lane, img, cntScratch, C.d, C.basePos, C.baseCnt, C.blkLen, C.blkCnt, C.nbytes, err, C.own, &C.pureMask);
	if (err && lane == 0) atomicOr(&A.ctl->err, err);
}

// One warp per work item = (logical block, slice of <= RMAX of its records).  Items the fast path
// cannot take (several items per block, > 32 records, empty block, unplaceable split) are queued
// for k_merge_general.
__global__ void __launch_bounds__(MERGE_WARPS * 32, MERGE_MINCTA) k_merge_fast(MergeArgs A)
{
	// persistent: warps pull the items k_merge_half deferred
	RB2_DYN_SMEM(smraw);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	FastSmem &S = reinterpret_cast<FastSmem*>(smraw)[wid];
	const uint32_t nTodoA = A.ctl->nTodoA;
	for (;;) {
		uint32_t q = 0;
		if (lane == 0) q = atomicAdd(&A.ctl->todoANext, 1u);
		q = __shfl_sync(FULLMASK, q, 0);
		if (q >= nTodoA) break;
		ItemCtx C;
		C.w = A.todoA[q];
		bool done = false;
		const ItemMeta m = A.itemMeta[C.w];
		if (m.nIt == 1) { // eligibility is known before touching the block
			// issue the record loads before the block decode so that their latency overlaps it
			RecPre rp = { 0, 0, NONE32 };
			if (m.r1 - m.r0 <= FAST_MAXREC && m.r0 + lane < m.r1) {
				const uint32_t r = m.r0 + lane;
				rp.P = A.recP[r]; rp.sc = A.recSC[r]; rp.dst = A.recDst[r];
			}
			if (lane < 6) S.f.cumBase[lane] = A.dir.cumCnt[(size_t)m.i * 6 + lane];
			item_prologue(A, C, lane, S.img, S.f.laneBase);
			if (C.r1 - C.r0 <= FAST_MAXREC && C.nbytes > 0) done = merge_fast(A, S, lane, C, rp);
		}
		if (!done && lane == 0) A.todo[atomicAdd(&A.ctl->nTodo, 1u)] = C.w;
		__syncwarp();
	}
}

// Persistent kernel: warps pull queued items until the queue is empty.
__global__ void __launch_bounds__(MERGE_WARPS * 32) k_merge_general(MergeArgs A)
{
	RB2_DYN_SMEM(smraw);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	GenSmem &S = reinterpret_cast<GenSmem*>(smraw)[wid];
	const uint32_t nTodo = A.ctl->nTodo;
	for (;;) {
		uint32_t k = 0;
		if (lane == 0) k = atomicAdd(&A.ctl->todoNext, 1u);
		k = __shfl_sync(FULLMASK, k, 0);
		if (k >= nTodo) break;
		ItemCtx C;
		C.w = A.todo[k];
		item_prologue(A, C, lane, S.img, S.g.lcnt);
		merge_general(A, S, lane, C);
		__syncwarp();
	}
}

struct RebuildScan { // K=1: pieces per old logical block -> new logical order
	const Ctl *ctl; Ctl *ctlw; uint32_t nlog;
	const uint32_t *order, *itemOff, *itemPieces, *itemFirst, *itemRest;
	uint32_t *orderNew;
	__device__ void load(uint64_t i, uint32_t (&v)[1]) const {
		uint32_t n = 0;
		for (uint32_t it = itemOff[i]; it < itemOff[i + 1]; ++it) n += itemPieces[it];
		v[0] = n ? n : 1;
	}
	__device__ void store(uint64_t i, const uint32_t (&own)[1], const uint32_t (&pre)[1]) const {
		uint32_t o = pre[0];
		if (itemOff[i] == itemOff[i + 1]) orderNew[o++] = order[i];
		else for (uint32_t it = itemOff[i]; it < itemOff[i + 1]; ++it) {
			orderNew[o++] = itemFirst[it];
			for (uint32_t k = 1; k < itemPieces[it]; ++k) orderNew[o++] = itemRest[it] + (k - 1);
		}
		const uint32_t nb = ctl->nb;
		if (nb == 6) {
#pragma unroll
			for (int b = 0; b < 6; ++b) if (i == ctl->blkBkt[b]) ctlw->blkBktNew[b] = pre[0];
		} else if (i == 0 || bucket_of(ctl->blkBkt, nb, (uint32_t)i) != bucket_of(ctl->blkBkt, nb, (uint32_t)i - 1)) {
			for (uint32_t b = 0; b < nb; ++b) if (i == ctl->blkBkt[b]) ctlw->blkBktNew[b] = pre[0];
		}
		if (i + 1 == nlog) {
			ctlw->nlogNew = pre[0] + own[0];
			for (uint32_t b = 0; b < nb + 2; ++b) if (b >= nb || ctl->blkBkt[b] >= nlog) ctlw->blkBktNew[b] = pre[0] + own[0];
		}
	}
};

struct DirScan { // K=7 (int64): per-symbol counts + length of every logical block -> cumCnt / cumLen
	const uint32_t *order, *blkCnt; uint32_t nlog;
	int64_t *cumLen, *cumCnt;
	// sharded engines: off[b][0..5] / off[b][6] = symbols of the index that sit in front of bucket b on
	// OTHER ranks (counts / length), so that the directory is in whole-index coordinates
	const int64_t *off; const uint32_t *bkt; uint32_t nb;
	__device__ void load(uint64_t i, int64_t (&v)[7]) const {
		const uint32_t *c = blkCnt + (size_t)order[i] * 6;
		int64_t t = 0;
#pragma unroll
		for (int a = 0; a < 6; ++a) { v[a] = c[a]; t += c[a]; }
		v[6] = t;
	}
	__device__ void store(uint64_t i, const int64_t (&own)[7], const int64_t (&pre)[7]) const {
		int64_t o[7] = { 0, 0, 0, 0, 0, 0, 0 };
		if (off) {
			const int64_t *ob = off + (size_t)bucket_of(bkt, nb, (uint32_t)i) * 7;
#pragma unroll
			for (int a = 0; a < 7; ++a) o[a] = ob[a];
		}
#pragma unroll
		for (int a = 0; a < 6; ++a) cumCnt[i * 6 + a] = pre[a] + o[a];
		cumLen[i] = pre[6] + o[6];
		if (i + 1 == nlog) {
#pragma unroll
			for (int a = 0; a < 6; ++a) cumCnt[(i + 1) * 6 + a] = pre[a] + own[a] + o[a];
			cumLen[i + 1] = pre[6] + own[6] + o[6];
		}
	}
};

// export: gather logical blocks [first, first+n) of the index into a contiguous staging area
__global__ void k_gather_blocks(const uint8_t *pool, const uint32_t *order, const uint32_t *blkCnt, uint32_t first, uint32_t n,
                                uint8_t *dst, int64_t *cnt)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t k = t >> 5, lane = t & 31;
	if (k >= n) return;
	const uint32_t p = order[first + k];
	reinterpret_cast<uint4*>(dst + (size_t)k * RB2_BLK)[lane] = reinterpret_cast<const uint4*>(pool + (size_t)p * RB2_BLK)[lane];
	if (cnt && lane < 6) cnt[(size_t)k * 6 + lane] = blkCnt[(size_t)p * 6 + lane];
}

// import: place n staged blocks at physical ids base.. and append them to the order array
__global__ void k_scatter_blocks(uint8_t *pool, uint32_t *blkCnt, uint32_t base, uint32_t n, const uint8_t *src, const int64_t *cnt)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t k = t >> 5, lane = t & 31;
	if (k >= n) return;
	reinterpret_cast<uint4*>(pool + (size_t)(base + k) * RB2_BLK)[lane] = reinterpret_cast<const uint4*>(src + (size_t)k * RB2_BLK)[lane];
	if (lane < 6) blkCnt[(size_t)(base + k) * 6 + lane] = (uint32_t)cnt[(size_t)k * 6 + lane];
}

__global__ void k_fill_u32(uint32_t *p, uint32_t n, uint32_t v0, uint32_t step)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v0 + i * step;
}

#include "rb2_flat.cuh"

// =====================================================================================
// Host side
// =====================================================================================

template <typename T> struct DevBuf {
	T *p = 0; size_t cap = 0;
	void need(size_t n) {
		if (n <= cap) return;
		if (p) RB2_CUDA(cudaFree(p));
		cap = n + n / 8 + 64;
		RB2_CUDA(cudaMalloc(&p, cap * sizeof(T)));
	}
	void release() { if (p) RB2_CUDA(cudaFree(p)); p = 0; cap = 0; }
};

enum { PH_H2D, PH_TRANSPOSE, PH_MEMBERS, PH_GROUPS, PH_MERGE, PH_DIR, PH_DIR2, PH_MEMBERS2, PH_MERGE2, PH_EXCH, PH_CONVERT, PH_N };

// dense regime (rb2_flat.cuh): the BWT as a flat array of symbols for the duration of one batch
struct FlatState {
	bool on = false; int cur = 0; uint64_t n = 0; uint32_t pending = 0; // pending: phases whose events wait for the next host sync
	bool valid = false, blocksStale = false; // the array holds the current index (resident between dense batches) / the leaf blocks do not
	DevBuf<uint8_t> s[2]; DevBuf<int64_t> dir[2]; DevBuf<uint32_t> tileCnt; DevBuf<TileDesc> desc; DevBuf<uint8_t> sliceBkt;
	DevBuf<uint8_t> chunkBytes; DevBuf<uint64_t> chunkPre, scanU64, midU64;
	void release() { for (int k = 0; k < 2; ++k) { s[k].release(); dir[k].release(); } tileCnt.release(); desc.release(); sliceBkt.release(); chunkBytes.release(); chunkPre.release(); scanU64.release(); midU64.release(); }
};

struct rb2_engine {
	int dev, so, nSM;
	cudaStream_t st;
	// leaf block pool + directory
	uint8_t *pool; uint32_t *blkCnt; uint32_t poolCap;
	Dir dir[2]; int cur;
	uint32_t nlog; uint32_t blkBkt[NBA];
	int nb;             // 6 (bucket = following symbol) or 36 (sharded: following two symbols)
	// sharded build (rb2_shard.inl): my rank, the exchange layer, the owner of every sub-bucket, the
	// whole-index symbol totals of all sub-buckets (identical on every rank), directory offsets
	int rank, nranks; Comm *comm; int owner[NBMAX];
	int64_t gtot[NBMAX][6];
	int64_t *dDirOff, *hDirOff;       // offsets of the directory (post-column while a column runs)
	int64_t *dDirOffPre, *hDirOffPre; // the same in front of the column (dense regime: record positions are pre-column)
	uint32_t *hPlan; DevBuf<uint32_t> plan;
	cudaStream_t st2; cudaEvent_t evEarly, evMerge; // second stream: the part of the exchange that overlaps the merge
	DevBuf<int64_t> gLrx[2]; DevBuf<uint32_t> sidrx[2]; // sharded: interval starts / string ids of the current and the next column (alternating)
	PeerRoute *dRoute;       // direct delivery: the merge epilogue's routing table of the column
	int64_t *peerGL[2][RB2_MAX_RANKS]; uint32_t *peerSid[2][RB2_MAX_RANKS]; bool p2pMapped; // every rank's gLrx[k] / sidrx[k] in my address space (kept across batches)
	// RB2_GPUS > 1 (rb2_cluster.inl): this engine is a proxy in front of nChild sharded engines
	int nChild; rb2_engine *child[RB2_MAX_RANKS]; rb2_group *grp;
	FlatState flat; DevBuf<uint32_t> recPre;
	int64_t tot[6][6]; int64_t bktLen[6];
	Ctl *dctl, *hctl;   // device control block and its pinned host mirror
	int64_t *dRankOut, *hRankOut;
	// batch scratch
	DevBuf<uint8_t> sbuf, T, asym, asym2, stage;
	DevBuf<int64_t> strEnd, gL[2], gSize[2], sizes6, recP, stageCnt;
	DevBuf<uint32_t> gOff[2], sid[2], tileA, tileB, grpCta, recSC, recDst, recHi, itemOff, itemPieces, itemFirst, itemRest, todo, todoA, scanCta;
	DevBuf<ItemMeta> itemMeta;
	DevBuf<int64_t> scanCta64, midTmp64;
	DevBuf<uint32_t> midTmp;
	unsigned long long *dMaxLen;
	// stats
	rb2_stats_t stats;
	int64_t lastP; int lastBkt; // single-string batches: where the sentinel went
	cudaEvent_t ev[PH_N][2], evTot[2];
	// asynchronous pipeline of the host entry point (rb2_async.inl): rb2_insert_multi returns when the batch is on the
	// device; a worker thread runs the insertions in order while the caller prepares / copies the next batch
	struct AsyncState *as;
};

static void dir_alloc(Dir &d, size_t cap)
{
	d.cap = cap;
	RB2_CUDA(cudaMalloc(&d.order, cap * sizeof(uint32_t)));
	RB2_CUDA(cudaMalloc(&d.cumLen, (cap + 1) * sizeof(int64_t)));
	RB2_CUDA(cudaMalloc(&d.cumCnt, (cap + 1) * 6 * sizeof(int64_t)));
}
static void dir_free(Dir &d)
{
	if (d.order) { RB2_CUDA(cudaFree(d.order)); RB2_CUDA(cudaFree(d.cumLen)); RB2_CUDA(cudaFree(d.cumCnt)); }
	d.order = 0; d.cumLen = 0; d.cumCnt = 0; d.cap = 0;
}

#include <chrono>
static int rb2_trace_on(void) { static int v = -1; if (v < 0) { const char *s = getenv("RB2_TRACE"); v = s && *s && *s != '0'; } return v; }
struct TraceT {
	rb2_engine *e; const char *name; std::chrono::steady_clock::time_point t0;
	TraceT(rb2_engine *e_, const char *n);
	~TraceT();
};

static inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

#define LAUNCH(e, kernel, grid, block, smem, ...) do { \
	RB2_KERNEL_LAUNCH(kernel, (grid), (block), (smem), (e)->st, __VA_ARGS__); ++(e)->stats.n_launches; \
	cudaError_t le_ = cudaGetLastError(); if (le_ != cudaSuccess) RB2_FATAL("launch of %s failed: %s", #kernel, cudaGetErrorString(le_)); } while (0)

TraceT::TraceT(rb2_engine *e_, const char *n) : e(e_), name(n) { if (rb2_trace_on()) { cudaStreamSynchronize(e->st); t0 = std::chrono::steady_clock::now(); } }
TraceT::~TraceT() { if (rb2_trace_on()) { cudaStreamSynchronize(e->st); double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); fprintf(stderr, "[trace] %-14s %9.1f us\n", name, us); } }

// exclusive scan of rows[n][K] in place + grand totals; hierarchical above 4096 rows
template <int K, typename T>
static void run_mid(rb2_engine *e, T *rows, uint64_t n, T *grandDev, DevBuf<T> &tmp)
{
	if (n <= 4096) { LAUNCH(e, (scan_mid<K, T>), 1, 1024, 0, rows, n, grandDev); return; }
	const uint32_t nChunk = cdiv(n, MID_ROWS);
	tmp.need((size_t)nChunk * K + K);
	LAUNCH(e, (mid_reduce<K, T>), nChunk, 256, 0, rows, n, tmp.p);
	LAUNCH(e, (scan_mid<K, T>), 1, 1024, 0, tmp.p, (uint64_t)nChunk, grandDev);
	LAUNCH(e, (mid_apply<K, T>), nChunk, 256, 0, rows, n, tmp.p);
}

// three-phase scan driver over n elements
template <int K, typename T, class F>
static void run_scan(rb2_engine *e, F f, uint64_t n, DevBuf<T> &cta, T *grandDev, DevBuf<T> &tmp)
{
	if (n == 0) return;
	uint32_t nCta = cdiv(n, SCAN_NT);
	cta.need((size_t)nCta * K + K);
	LAUNCH(e, (scan_reduce<K, T, F>), nCta, SCAN_NT, 0, f, n, cta.p);
	run_mid<K, T>(e, cta.p, nCta, grandDev ? grandDev : cta.p + (size_t)nCta * K, tmp);
	LAUNCH(e, (scan_apply<K, T, F>), nCta, SCAN_NT, 0, f, n, cta.p);
}

static void ctl_pull(rb2_engine *e)
{
	RB2_CUDA(cudaMemcpyAsync(e->hctl, e->dctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	if (e->hctl->err) RB2_FATAL("device error flags 0x%x (1=block pool exhausted, 2=8-byte run in pool, 4=staging overflow, 8=scan mismatch, 16=too many pieces)", e->hctl->err);
}
// host copy of the control block -> device.  The block travels as a KERNEL ARGUMENT, not as a host-to-device copy:
// a DMA copy would queue behind the 10 GB batch that the caller's thread may be uploading at the same time
// (rb2_async.inl) and stall the column loop for the rest of that transfer.
static_assert(sizeof(Ctl) <= 4000, "Ctl must fit the kernel parameter space");
__global__ void __launch_bounds__(256) k_ctl_store(Ctl *dst, const Ctl v)
{
	const uint32_t *src = reinterpret_cast<const uint32_t*>(&v);
	uint32_t *d = reinterpret_cast<uint32_t*>(dst);
	for (uint32_t i = threadIdx.x; i < sizeof(Ctl) / 4; i += 256) d[i] = src[i];
}
static void ctl_push(rb2_engine *e)
{
	LAUNCH(e, k_ctl_store, 1, 256, 0, e->dctl, *e->hctl);
}

// rebuild cumLen/cumCnt of the current directory from blkCnt + order
static void rebuild_directory(rb2_engine *e, bool afterMerge)
{
	Dir &d = e->dir[e->cur];
	// a sharded engine adds the symbols other ranks hold in front of each of its sub-buckets; right
	// behind a merge the device control block holds the new block ranges in blkBktNew
	DirScan f = { d.order, e->blkCnt, e->nlog, d.cumLen, d.cumCnt, e->comm ? e->dDirOff : (const int64_t*)0,
	              afterMerge ? e->dctl->blkBktNew : e->dctl->blkBkt, (uint32_t)e->nb };
	run_scan<7, int64_t, DirScan>(e, f, e->nlog, e->scanCta64, (int64_t*)0, e->midTmp64);
}

// refresh host mirrors of per-bucket totals from the directory
static void pull_totals(rb2_engine *e)
{
	Dir &d = e->dir[e->cur];
	std::vector<int64_t> c(7 * 6);
	for (int b = 0; b <= 6; ++b)
		RB2_CUDA(cudaMemcpyAsync(&c[b * 6], d.cumCnt + (size_t)e->blkBkt[b] * 6, 48, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	for (int b = 0; b < 6; ++b) {
		e->bktLen[b] = 0;
		for (int a = 0; a < 6; ++a) { e->tot[b][a] = c[(b + 1) * 6 + a] - c[b * 6 + a]; e->bktLen[b] += e->tot[b][a]; }
	}
}

// Grow everything that is sized by the number of leaf blocks: the pool, the per-block counts,
// both directory buffers and the per-block planning arrays.  Allocation is slow (tens of ms),
// so callers reserve generously once per batch; the merge kernel tolerates running out (retry).
static void reserve_blocks(rb2_engine *e, uint64_t blocks)
{
	if (blocks <= e->poolCap) return;
	if (blocks > 0xfffffff0ull) RB2_FATAL("block pool would exceed 2^32 blocks");
	const uint64_t perBlock = RB2_BLK + 24 + 2 * 60 + 8;
	size_t freeB = 0, totB = 0;
	RB2_CUDA(cudaStreamSynchronize(e->st));
	RB2_CUDA(cudaMemGetInfo(&freeB, &totB));
	if (blocks * perBlock > freeB)
		RB2_FATAL("out of HBM: need %llu leaf blocks (%.1f GB), %.1f GB free", (unsigned long long)blocks, blocks * perBlock * 1e-9, freeB * 1e-9);
	const uint64_t cap = blocks;
	uint8_t *np; uint32_t *nc;
	RB2_CUDA(cudaMalloc(&np, cap * RB2_BLK));
	RB2_CUDA(cudaMalloc(&nc, cap * 6 * sizeof(uint32_t)));
	const uint32_t used = e->hctl->poolUsed;
	if (e->pool) {
		RB2_CUDA(cudaMemcpyAsync(np, e->pool, (size_t)used * RB2_BLK, cudaMemcpyDeviceToDevice, e->st));
		RB2_CUDA(cudaMemcpyAsync(nc, e->blkCnt, (size_t)used * 24, cudaMemcpyDeviceToDevice, e->st));
		RB2_CUDA(cudaStreamSynchronize(e->st));
		RB2_CUDA(cudaFree(e->pool)); RB2_CUDA(cudaFree(e->blkCnt));
	}
	e->pool = np; e->blkCnt = nc; e->poolCap = (uint32_t)cap;
	for (int k = 0; k < 2; ++k) {
		Dir &d = e->dir[k];
		Dir nd; dir_alloc(nd, cap);
		if (k == e->cur && e->nlog && d.order) {
			RB2_CUDA(cudaMemcpyAsync(nd.order, d.order, (size_t)e->nlog * 4, cudaMemcpyDeviceToDevice, e->st));
			RB2_CUDA(cudaMemcpyAsync(nd.cumLen, d.cumLen, ((size_t)e->nlog + 1) * 8, cudaMemcpyDeviceToDevice, e->st));
			RB2_CUDA(cudaMemcpyAsync(nd.cumCnt, d.cumCnt, ((size_t)e->nlog + 1) * 48, cudaMemcpyDeviceToDevice, e->st));
			RB2_CUDA(cudaStreamSynchronize(e->st));
		}
		dir_free(d);
		d = nd;
	}
	e->recHi.need(cap); e->itemOff.need(cap + 1);
	e->hctl->poolCap = e->poolCap;
}

static void reserve_items(rb2_engine *e, uint64_t n)
{
	e->itemMeta.need(n); e->itemPieces.need(n); e->itemFirst.need(n); e->itemRest.need(n); e->todo.need(n); e->todoA.need(n);
}

extern "C" int rb2_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

static bool async_enabled(const rb2_engine *e); // rb2_async.inl: the host entry point as a pipeline
static void async_drain(rb2_engine *e);
static void async_init(rb2_engine *e);
static void async_destroy(rb2_engine *e);
static void async_reset(rb2_engine *e);
static void ensure_blocks(rb2_engine *e);  // rb2_flat_host.inl: rebuild the leaf blocks from the resident flat array if they are stale
static void blocks_edited(rb2_engine *e);  // ... and drop the array when the blocks change underneath it
static void cluster_attach(rb2_engine *e, int device, int sorting_order);
static void cluster_destroy(rb2_engine *e);
static void cluster_reset(rb2_engine *e);
static void cluster_insert_multi(rb2_engine *e, int64_t len, const uint8_t *s);
static int64_t cluster_num_blocks(rb2_engine *e, int bucket);
static int64_t cluster_fetch_blocks(rb2_engine *e, int bucket, int64_t first, int64_t n, uint8_t *dst, int64_t *cnt);
static void cluster_rank1(rb2_engine *e, int64_t x, int64_t c[6]);
#define RB2_NO_CLUSTER(e, what) do { if ((e)->nChild) RB2_FATAL(what " is not available with RB2_GPUS > 1"); } while (0)

static rb2_engine *engine_create(int device, int sorting_order, bool multi);
extern "C" rb2_engine_t *rb2_create(int device, int sorting_order) { return engine_create(device, sorting_order, false); }
// what mr_init uses: with RB2_GPUS=P (P > 1) the engine is a proxy over P sharded engines (rb2_cluster.inl)
extern "C" rb2_engine_t *rb2_create_auto(int device, int sorting_order) { return engine_create(device, sorting_order, true); }
static rb2_engine *engine_create(int device, int sorting_order, bool multi)
{
	if (sorting_order < 0 || sorting_order > 2) RB2_FATAL("sorting order must be 0, 1 or 2 (mrope.c:18)");
	int n = rb2_device_count();
	if (n <= 0) RB2_FATAL("no CUDA device visible: this library has no CPU path");
	if (device < 0 || device >= n) RB2_FATAL("device %d out of range (0..%d)", device, n - 1);
	RB2_CUDA(cudaSetDevice(device));
	rb2_engine *e = new rb2_engine();
	memset(&e->stats, 0, sizeof(e->stats));
	e->dev = device; e->so = sorting_order;
	e->nChild = 0; e->grp = 0;
	e->rank = 0; e->nranks = 1; e->comm = 0; e->dDirOff = 0; e->hDirOff = 0; e->dDirOffPre = 0; e->hDirOffPre = 0; e->hPlan = 0; e->dRoute = 0; e->p2pMapped = false; memset(e->peerGL, 0, sizeof(e->peerGL)); memset(e->peerSid, 0, sizeof(e->peerSid));
	RB2_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
	RB2_CUDA(cudaMalloc(&e->dctl, sizeof(Ctl)));
	RB2_CUDA(cudaMallocHost(&e->hctl, sizeof(Ctl)));
	RB2_CUDA(cudaMalloc(&e->dRankOut, 12 * sizeof(int64_t)));
	RB2_CUDA(cudaMallocHost(&e->hRankOut, 12 * sizeof(int64_t)));
	RB2_CUDA(cudaMalloc(&e->dMaxLen, sizeof(unsigned long long)));
	memset(e->hctl, 0, sizeof(Ctl));
	for (int p = 0; p < PH_N; ++p) for (int k = 0; k < 2; ++k) RB2_CUDA(cudaEventCreate(&e->ev[p][k]));
	for (int k = 0; k < 2; ++k) RB2_CUDA(cudaEventCreate(&e->evTot[k]));
	e->pool = 0; e->blkCnt = 0; e->poolCap = 0;
	memset(e->dir, 0, sizeof(e->dir)); e->cur = 0;
	RB2_CUDA(cudaFuncSetAttribute(k_merge_half, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MERGE_WARPS * 2 * sizeof(HalfSmem))));
	RB2_CUDA(cudaFuncSetAttribute(k_merge_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MERGE_WARPS * sizeof(FastSmem))));
	RB2_CUDA(cudaFuncSetAttribute(k_merge_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MERGE_WARPS * sizeof(GenSmem))));
#define RB2_MERGE_SMEM(NS, GEN, SH) RB2_CUDA(cudaFuncSetAttribute(NS::k_flat_merge<GEN, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FS_WARPS * sizeof(NS::SliceWarpSmem))))
	RB2_MERGE_SMEM(fs2, true, false); RB2_MERGE_SMEM(fs2, false, false); RB2_MERGE_SMEM(fs4, true, false); RB2_MERGE_SMEM(fs4, false, false);
	RB2_MERGE_SMEM(fs2, true, true); RB2_MERGE_SMEM(fs2, false, true); RB2_MERGE_SMEM(fs4, true, true); RB2_MERGE_SMEM(fs4, false, true);
#undef RB2_MERGE_SMEM
	{ cudaDeviceProp pr; RB2_CUDA(cudaGetDeviceProperties(&pr, device)); e->nSM = pr.multiProcessorCount; }
	// six empty buckets, one empty leaf block each (rope_init, rope.c:55-69)
	reserve_blocks(e, 4096);
	RB2_CUDA(cudaMemsetAsync(e->pool, 0, 6 * RB2_BLK, e->st));
	RB2_CUDA(cudaMemsetAsync(e->blkCnt, 0, 6 * 24, e->st));
	LAUNCH(e, k_fill_u32, 1, 32, 0, e->dir[0].order, 6u, 0u, 1u);
	e->nlog = 6;
	for (int b = 0; b < NBA; ++b) e->blkBkt[b] = b < 6 ? b : 6;
	e->nb = 6; e->hctl->nb = 6; e->hctl->tables = 0;
	e->hctl->poolUsed = 6; e->hctl->poolCap = e->poolCap;
	ctl_push(e);
	rebuild_directory(e, false);
	pull_totals(e);
	e->as = 0;
	if (multi) cluster_attach(e, device, sorting_order);
	if (async_enabled(e)) async_init(e);
	return e;
}

// Empty the index but keep every allocation (bench steps and tests reuse one engine).
static void shard_reset_index(rb2_engine *e);
static void reset_index(rb2_engine *e);
extern "C" void rb2_reset(rb2_engine_t *e)
{
	if (e->nChild) { cluster_reset(e); return; }
	RB2_CUDA(cudaSetDevice(e->dev));
	if (async_enabled(e)) { async_reset(e); return; } // queued behind the batches in flight
	reset_index(e);
}
static void reset_index(rb2_engine *e)
{
	e->flat.valid = false; e->flat.blocksStale = false;
	if (e->comm) { shard_reset_index(e); return; }
	RB2_CUDA(cudaMemsetAsync(e->pool, 0, 6 * RB2_BLK, e->st));
	RB2_CUDA(cudaMemsetAsync(e->blkCnt, 0, 6 * 24, e->st));
	LAUNCH(e, k_fill_u32, 1, 32, 0, e->dir[e->cur].order, 6u, 0u, 1u);
	e->nlog = 6;
	for (int b = 0; b < NBA; ++b) e->blkBkt[b] = b < 6 ? b : 6;
	e->hctl->poolUsed = 6; e->hctl->poolCap = e->poolCap; e->hctl->err = 0;
	ctl_push(e);
	rebuild_directory(e, false);
	pull_totals(e);
	e->stats.pool_blocks = 6;
}

extern "C" void *rb2_host_alloc(int64_t bytes)
{
	void *p = 0;
	RB2_CUDA(cudaMallocHost(&p, (size_t)bytes));
	return p;
}
extern "C" void rb2_host_free(void *p) { RB2_CUDA(cudaFreeHost(p)); }

extern "C" void rb2_destroy(rb2_engine_t *e)
{
	if (!e) return;
	if (e->nChild) cluster_destroy(e);
	RB2_CUDA(cudaSetDevice(e->dev));
	async_drain(e); async_destroy(e);
	RB2_CUDA(cudaStreamSynchronize(e->st));
	if (e->pool) { RB2_CUDA(cudaFree(e->pool)); RB2_CUDA(cudaFree(e->blkCnt)); }
	dir_free(e->dir[0]); dir_free(e->dir[1]);
	e->sbuf.release(); e->T.release(); e->asym.release(); e->asym2.release(); e->stage.release();
	e->strEnd.release(); e->sizes6.release(); e->recP.release(); e->stageCnt.release();
	for (int k = 0; k < 2; ++k) { e->gL[k].release(); e->gSize[k].release(); e->gOff[k].release(); e->sid[k].release(); }
	e->tileA.release(); e->tileB.release(); e->grpCta.release(); e->recSC.release(); e->recDst.release(); e->recHi.release();
	e->itemOff.release(); e->itemMeta.release(); e->itemPieces.release(); e->itemFirst.release(); e->itemRest.release(); e->todo.release(); e->todoA.release();
	e->scanCta.release(); e->scanCta64.release(); e->midTmp.release(); e->midTmp64.release();
	RB2_CUDA(cudaFree(e->dctl)); RB2_CUDA(cudaFreeHost(e->hctl));
	RB2_CUDA(cudaFree(e->dRankOut)); RB2_CUDA(cudaFreeHost(e->hRankOut));
	RB2_CUDA(cudaFree(e->dMaxLen));
	if (e->comm && e->p2pMapped) { // not collective: close what I imported; my own buffers are freed below (rb2_sharded_quiesce is the collective, ordered way)
		for (int k = 0; k < 2; ++k) { e->comm->p2p_unmap((void**)e->peerGL[k]); e->comm->p2p_unmap((void**)e->peerSid[k]); }
		e->p2pMapped = false;
	}
	if (e->comm) { delete e->comm; cudaStreamDestroy(e->st2); cudaEventDestroy(e->evEarly); cudaEventDestroy(e->evMerge); RB2_CUDA(cudaFree(e->dDirOff)); RB2_CUDA(cudaFreeHost(e->hDirOff)); RB2_CUDA(cudaFree(e->dDirOffPre)); RB2_CUDA(cudaFreeHost(e->hDirOffPre)); RB2_CUDA(cudaFreeHost(e->hPlan)); e->plan.release(); e->gLrx[0].release(); e->gLrx[1].release(); e->sidrx[0].release(); e->sidrx[1].release(); if (e->dRoute) RB2_CUDA(cudaFree(e->dRoute)); }
	e->flat.release(); e->recPre.release();
	for (int p = 0; p < PH_N; ++p) for (int k = 0; k < 2; ++k) cudaEventDestroy(e->ev[p][k]);
	for (int k = 0; k < 2; ++k) cudaEventDestroy(e->evTot[k]);
	RB2_CUDA(cudaStreamDestroy(e->st));
	delete e;
}

extern "C" int rb2_sorting_order(const rb2_engine_t *e) { return e->so; }

static inline void ph_begin(rb2_engine *e, int p) { RB2_CUDA(cudaEventRecord(e->ev[p][0], e->st)); }
static inline void ph_end(rb2_engine *e, int p) { RB2_CUDA(cudaEventRecord(e->ev[p][1], e->st)); }
static void ph_collect(rb2_engine *e, uint32_t mask)
{
	double *acc[PH_N] = { &e->stats.ms_h2d, &e->stats.ms_transpose, &e->stats.ms_members, &e->stats.ms_groups, &e->stats.ms_merge,
	                      &e->stats.ms_directory, &e->stats.ms_directory, &e->stats.ms_members, &e->stats.ms_merge_general, &e->stats.ms_exchange, &e->stats.ms_convert };
	for (int p = 0; p < PH_N; ++p) if (mask >> p & 1) {
		float ms = 0;
		RB2_CUDA(cudaEventElapsedTime(&ms, e->ev[p][0], e->ev[p][1]));
		*acc[p] += ms;
	}
}

// Merge `nrec` insertion records (sorted by position; per-bucket ranges in dctl->recBkt) into
// the leaf blocks and rebuild the directory.  The device control block must already hold
// blkBkt / recBkt / cpost for this step; the host mirror e->hctl must be current.
// gLNext[recDst[r]] receives cpost[sym] + occ(sym, position) for every record with a target.
static void apply_records(rb2_engine *e, uint32_t nrec, int64_t *gLNext)
{
	Ctl *h = e->hctl;
	// ---- plan work items -----------------------------------------------------------
	ph_begin(e, PH_DIR);
	const uint64_t maxItems = (uint64_t)std::min<uint64_t>(e->nlog, nrec) + nrec / RMAX + 1;
	reserve_items(e, maxItems);                    // no-op when the batch-level reservation holds
	if ((uint64_t)h->poolUsed + 1024 > e->poolCap) { // keep a little slack so most columns never retry
		reserve_blocks(e, (uint64_t)e->poolCap + e->poolCap / 2 + 4096);
		ctl_push(e);
	}
	{
		Dir &dc = e->dir[e->cur];
		LAUNCH(e, k_rec_hi, cdiv(e->nlog, 256), 256, 0, dc, e->nlog, e->dctl, e->recP.p, e->recHi.p);
		ItemScan is = { e->dctl, e->recHi.p, e->nlog, e->itemOff.p, e->itemMeta.p, e->dctl, dc.order, dc.cumLen };
		run_scan<1, uint32_t, ItemScan>(e, is, e->nlog, e->scanCta, (uint32_t*)0, e->midTmp);
		RB2_CUDA(cudaMemsetAsync(e->itemPieces.p, 0, maxItems * 4, e->st)); // 0 = not merged yet
	}
	ph_end(e, PH_DIR);
	const uint32_t usedBefore = h->poolUsed;
	uint32_t nItems = 0;
	for (int attempt = 0;; ++attempt) {
		Dir &dc = e->dir[e->cur], &dnx = e->dir[e->cur ^ 1];
		// ---- merge -----------------------------------------------------------------
		if (attempt == 0) ph_begin(e, PH_MERGE);
		MergeArgs ma = { e->pool, e->blkCnt, dc, e->nlog, e->recHi.p, e->itemOff.p, e->itemMeta.p,
		                 e->recP.p, e->recSC.p, e->recDst.p, gLNext,
		                 e->itemPieces.p, e->itemFirst.p, e->itemRest.p, e->todoA.p, e->todo.p, e->dctl };
		LAUNCH(e, k_merge_half, cdiv(maxItems, MERGE_WARPS * 2), MERGE_WARPS * 32, MERGE_WARPS * 2 * sizeof(HalfSmem), ma);
		if (attempt == 0) { ph_end(e, PH_MERGE); ph_begin(e, PH_MERGE2); }
		LAUNCH(e, k_merge_fast, e->nSM * MERGE_MINCTA, MERGE_WARPS * 32, MERGE_WARPS * sizeof(FastSmem), ma);
		LAUNCH(e, k_merge_general, e->nSM * 10, MERGE_WARPS * 32, MERGE_WARPS * sizeof(GenSmem), ma);
		if (attempt == 0) ph_end(e, PH_MERGE2);
		++e->stats.n_merge_launches;
		// ---- new logical order ----------------------------------------------------------
		if (attempt == 0) ph_begin(e, PH_DIR2);
		RebuildScan rs = { e->dctl, e->dctl, e->nlog, dc.order, e->itemOff.p, e->itemPieces.p, e->itemFirst.p, e->itemRest.p, dnx.order };
		run_scan<1, uint32_t, RebuildScan>(e, rs, e->nlog, e->scanCta, (uint32_t*)0, e->midTmp);
		ctl_pull(e);
		nItems = h->nItems;
		if (!h->overflow) break;
		// pool ran out: the items that did not fit are untouched.  Grow and run them again.
		if (attempt > 8) RB2_FATAL("block pool growth did not converge");
		h->poolUsed = h->failBase; h->overflow = 0; h->failBase = NONE32; h->nTodo = 0; h->todoNext = 0; h->nTodoA = 0; h->todoANext = 0;
		reserve_blocks(e, (uint64_t)e->poolCap + e->poolCap / 2 + maxItems / 4 + 4096);
		ctl_push(e);
	}
	e->nlog = h->nlogNew;
	for (int b = 0; b < e->nb + 2; ++b) e->blkBkt[b] = h->blkBktNew[b];
	e->cur ^= 1;
	rebuild_directory(e, true);
	ph_end(e, PH_DIR2);
	RB2_CUDA(cudaStreamSynchronize(e->st));
	ph_collect(e, (1u << PH_MERGE) | (1u << PH_MERGE2) | (1u << PH_DIR) | (1u << PH_DIR2));
	e->stats.general_items += h->nTodoA;
	e->stats.merge_blocks += nItems;
	// every item reads one leaf block and writes it back, plus the freshly allocated pieces
	e->stats.merge_bytes_rw += ((int64_t)nItems * 2 + (int64_t)(h->poolUsed - usedBefore)) * RB2_BLK;
}

#include "rb2_flat_host.inl"

static FILE *column_log(void)
{
	static FILE *f = 0; static int init = 0;
	if (!init) { init = 1; const char *p = getenv("RB2_COLLOG"); if (p && *p) f = fopen(p, "w"); }
	return f;
}

static void insert_string_range(rb2_engine *e, const uint8_t *s, uint32_t kBase, uint32_t m);
// padding a batch's symbol matrix may carry before the batch is cut (RB2_SPLIT_SLACK bytes: the tests lower it)
static size_t split_slack(void) { const char *v = getenv("RB2_SPLIT_SLACK"); return v && *v ? (size_t)atoll(v) : (size_t)64 << 20; }

// One sub-batch whose strings already sit in device memory at `s` (len bytes, ends with NUL).
static void insert_device_batch(rb2_engine *e, int64_t len, const uint8_t *s)
{
	// ---- split into strings, transpose -------------------------------------------------
	ph_begin(e, PH_TRANSPOSE);
	const uint32_t nT = cdiv(len, 4096);
	e->tileA.need((size_t)nT + 2);
	LAUNCH(e, k_count_nul, nT, 256, 0, s, len, e->tileA.p);
	uint32_t *dTot = e->tileA.p + nT; // grand total lands behind the tile array
	run_mid<1, uint32_t>(e, e->tileA.p, (uint64_t)nT, dTot, e->midTmp);
	uint32_t m = 0;
	RB2_CUDA(cudaMemcpyAsync(&m, dTot, 4, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	if (m == 0) RB2_FATAL("batch holds no terminated string");
	e->strEnd.need(m);
	LAUNCH(e, k_string_ends, nT, 256, 0, s, len, e->tileA.p, e->strEnd.p);
	ph_end(e, PH_TRANSPOSE);
	RB2_CUDA(cudaStreamSynchronize(e->st));
	ph_collect(e, 1u << PH_TRANSPOSE);
	insert_string_range(e, s, 0, m);
	e->stats.n_symbols += len;
}

// Strings kBase .. kBase+m-1 of the batch at `s` (string ends in e->strEnd).  The per-batch state is a dense
// rectangle of (longest string + 1) columns x m strings: when one long string among many short ones (or a
// long-read batch) would make it far larger than the strings themselves, or larger than the free memory, the
// range is cut in two and the halves are inserted one after the other -- the BWT does not depend on how a
// batch is cut, and the reference accepts any mix of lengths.
static void insert_string_range(rb2_engine *e, const uint8_t *s, uint32_t kBase, uint32_t m)
{
	const int sorted = e->so != RB2_SO_IO;
	ph_begin(e, PH_TRANSPOSE);
	RB2_CUDA(cudaMemsetAsync(e->dMaxLen, 0, 8, e->st));
	LAUNCH(e, k_maxlen, cdiv(m, 256), 256, 0, e->strEnd.p, kBase, m, e->dMaxLen);
	unsigned long long maxlen = 0;
	int64_t ends[2] = { -1, 0 };
	RB2_CUDA(cudaMemcpyAsync(&maxlen, e->dMaxLen, 8, cudaMemcpyDeviceToHost, e->st));
	if (kBase) RB2_CUDA(cudaMemcpyAsync(&ends[0], e->strEnd.p + kBase - 1, 8, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaMemcpyAsync(&ends[1], e->strEnd.p + kBase + m - 1, 8, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	const int64_t ncol = (int64_t)maxlen + 1, len = ends[1] - ends[0];
	{
		size_t freeB = 0, totB = 0;
		RB2_CUDA(cudaMemGetInfo(&freeB, &totB));
		const size_t need = (size_t)ncol * t_stride(m);
		const bool tooBig = need > e->T.cap && need > (freeB + e->T.cap) / 2;
		const bool wasteful = need > (size_t)8 * (size_t)len + split_slack(); // the matrix holds 4 bits per symbol: > 16x padding
		if ((tooBig || wasteful) && m > 1) {
			ph_end(e, PH_TRANSPOSE);
			insert_string_range(e, s, kBase, m / 2);
			insert_string_range(e, s, kBase + m / 2, m - m / 2);
			return;
		}
		if (tooBig) RB2_FATAL("one string of %lld symbols does not fit in HBM as a column-major symbol matrix", (long long)maxlen);
	}
	e->T.need((size_t)ncol * t_stride(m) + 16);
	LAUNCH(e, k_transpose, cdiv(m, TR_S), 256, 0, s, e->strEnd.p, kBase, m, ncol, e->T.p);
	ph_end(e, PH_TRANSPOSE);

	// ---- state for column 0 (mrope.c:279-285) -------------------------------------------
	for (int k = 0; k < 2; ++k) { e->gL[k].need((size_t)m + 64); e->gSize[k].need((size_t)m + 64); e->gOff[k].need((size_t)m + 64); e->sid[k].need((size_t)m + 64); } // (+64: the merge kernel's TMA slices are widened to 16-byte boundaries)
	e->asym.need((size_t)m + 64);
	const size_t recCap = (size_t)m + m / RB2_MAXRUN + 64;
	e->recP.need(recCap); e->recSC.need(recCap); e->recDst.need(recCap);
	// dense regime: the whole batch runs on a flat symbol array (rb2_flat.cuh), re-encoded at the end
	const bool flat = flat_choose(e, m, (uint64_t)len);
	if (flat) { e->recPre.need(recCap + 1); flat_begin(e, (uint64_t)len); }
	else {
		ensure_blocks(e); blocks_edited(e); // a sparse batch edits the leaf blocks
		// Reserve leaf blocks for the whole batch up front (2 bytes of pool per new symbol covers random
		// data at B+-tree fill plus blocks retired by multi-item merges); more is added on demand.
		reserve_blocks(e, (uint64_t)e->hctl->poolUsed + (uint64_t)len * 2 / RB2_FILL + 4096);
		reserve_items(e, std::min<uint64_t>(e->poolCap, recCap) + recCap / RMAX + 2);
	}
	const int64_t n0 = e->bktLen[0];
	const bool useSizes = sorted && n0 > 0;
	if (useSizes) e->sizes6.need((size_t)m * 6);
	int cs = 0; // current state buffer
	// all-singleton columns of a dense batch run as one kernel (k_column_fused): the next symbols travel with the strings
	static int fusedPref = -1;
	if (fusedPref < 0) { const char *fp = getenv("RB2_FUSED"); fusedPref = !(fp && *fp == '0'); }
	bool asymReady = false; // asymCur already holds this column's symbols (fetched by the previous column's kernel)
	uint8_t *asymCur = 0, *asymNxt = 0;
	LAUNCH(e, k_init_state, cdiv(m, 256), 256, 0, sorted, m, n0, e->gL[0].p, e->gSize[0].p, e->gOff[0].p, e->sid[0].p);
	uint32_t G = sorted ? 1 : m, M = m;
	uint32_t gBkt[8], mBkt[8];
	for (int b = 0; b < 8; ++b) { gBkt[b] = b == 0 ? 0 : G; mBkt[b] = b == 0 ? 0 : M; }
	RB2_CUDA(cudaStreamSynchronize(e->st));
	ph_collect(e, 1u << PH_TRANSPOSE);

	for (int64_t col = 0; M > 0; ++col) {
		if (col >= ncol) RB2_FATAL("internal: live strings beyond the last column");
		Dir &d = e->dir[e->cur];
		Ctl *h = e->hctl;
		// control block for this column
		for (int b = 0; b < 8; ++b) { h->gBkt[b] = gBkt[b]; h->mBkt[b] = mBkt[b]; h->blkBkt[b] = e->blkBkt[b]; }
		{ // bucket starts after this column: every member of bucket b inserts one symbol into it
			int64_t acc = 0;
			for (int b = 0; b < 6; ++b) { h->cpost[b] = acc; acc += e->bktLen[b] + (mBkt[b + 1] - mBkt[b]); }
			h->cpost[6] = h->cpost[7] = acc;
		}
		h->poolCap = e->poolCap; h->nItems = 0; h->err = 0; h->overflow = 0; h->failBase = NONE32; h->nTodo = 0; h->todoNext = 0; h->nTodoA = 0; h->todoANext = 0;
		ctl_push(e);

		// ---- members: next symbol + tile histograms ---------------------------------
		ph_begin(e, PH_MEMBERS);
		const uint32_t nTile = cdiv(M, MEM_TILE);
		const bool fused = fusedPref && flat && G == M && m > 1;
		if (!asymCur) { e->asym2.need((size_t)m + 64); asymCur = e->asym.p; asymNxt = e->asym2.p; }
		e->tileB.need(((size_t)nTile + 1) * 6 + 8);
		RB2_CUDA(cudaMemsetAsync(e->tileB.p + (size_t)nTile * 6, 0, 24, e->st)); // terminal entry -> totals
		if (!asymReady) {
			TView tv; memset(&tv, 0, sizeof(tv));
			tv.n = 1; tv.off[1] = m; tv.col[0] = e->T.p + (size_t)col * t_stride(m);
			LAUNCH(e, k_member_fetch, nTile, 256, 0, tv, e->sid[cs].p, M, asymCur, e->tileB.p);
		} else LAUNCH(e, k_tile_hist, nTile, 256, 0, asymCur, M, e->tileB.p); // the symbols came with the strings
		run_mid<6, uint32_t>(e, e->tileB.p, (uint64_t)nTile + 1, e->dctl->memTot, e->midTmp);
		ph_end(e, PH_MEMBERS);

		// ---- groups: interval sizes, histograms, records ------------------------------
		ph_begin(e, PH_GROUPS);
		if (useSizes && !fused) {
			if (flat) LAUNCH(e, k_flat_rank_groups, cdiv(G, 128), 128, 0, e->flat.s[e->flat.cur].p, e->flat.dir[e->flat.cur].p, G, e->gL[cs].p, e->gSize[cs].p,
			                 e->sizes6.p, e->dctl, (const int64_t*)0, e->nb);
			else LAUNCH(e, k_rank_groups, cdiv(G, 128), 128, 0, e->pool, d, e->nlog, G, e->gL[cs].p, e->gSize[cs].p, e->sizes6.p, e->dctl);
		}
		const bool lean = flat && !useSizes && G == M && m > 1;
		if (fused) {
			// every group is a singleton, dense regime: partition + insertion points + next-symbol fetch in one kernel
			FusedArgs fa = { e->sid[cs].p, asymCur, M, e->tileB.p, col + 1 < ncol ? e->T.p + (size_t)(col + 1) * t_stride(m) : (const uint8_t*)0,
			                 e->gL[cs].p, e->gSize[cs].p, e->flat.s[e->flat.cur].p, e->flat.dir[e->flat.cur].p, e->dctl,
			                 e->sid[cs ^ 1].p, asymNxt, e->recDst.p, e->recP.p, e->gSize[cs ^ 1].p };
			if (useSizes) { if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_column_fused<true, true>), nTile, 256, 0, fa); else LAUNCH(e, (k_column_fused<true, false>), nTile, 256, 0, fa); }
			else LAUNCH(e, (k_column_fused<false, false>), nTile, 256, 0, fa);
			ph_end(e, PH_GROUPS);
			ph_begin(e, PH_MEMBERS2); ph_end(e, PH_MEMBERS2);
		} else if (G == M) {
			// every group is a singleton: records, next groups and the partition in one kernel
			LAUNCH(e, k_col_bases_single, 1, 1, 0, e->dctl, e->gOff[cs ^ 1].p, M, flat ? e->recPre.p : (uint32_t*)0);
			SingleArgs sa = { e->sid[cs].p, asymCur, M, e->tileB.p, e->gL[cs].p, e->gSize[cs].p, useSizes ? e->sizes6.p : 0, e->dctl,
			                  e->sid[cs ^ 1].p, e->gSize[cs ^ 1].p, e->gOff[cs ^ 1].p, e->recP.p, e->recSC.p, e->recDst.p, flat ? e->recPre.p : (uint32_t*)0, lean ? 1 : 0 };
			if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_column_singletons<true>), nTile, 256, 0, sa);
			else LAUNCH(e, (k_column_singletons<false>), nTile, 256, 0, sa);
			ph_end(e, PH_GROUPS);
			ph_begin(e, PH_MEMBERS2); ph_end(e, PH_MEMBERS2);
		} else {
			const uint32_t nGC = cdiv(G, 256);
			e->grpCta.need((size_t)nGC * NGC + NGC);
			GroupArgs ga = { e->gOff[cs].p, e->gL[cs].p, e->gSize[cs].p, useSizes ? e->sizes6.p : 0, asymCur, e->tileB.p, G,
			                 e->grpCta.p, e->gSize[cs ^ 1].p, e->gOff[cs ^ 1].p, e->recP.p, e->recSC.p, e->recDst.p, e->dctl, flat ? e->recPre.p : (uint32_t*)0 };
			if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_group_pass<0, true>), nGC, 256, 0, ga);
			else LAUNCH(e, (k_group_pass<0, false>), nGC, 256, 0, ga);
			run_mid<NGC, uint32_t>(e, e->grpCta.p, (uint64_t)nGC, e->dctl->grpTot, e->midTmp);
			LAUNCH(e, k_col_bases, 1, 1, 0, e->dctl, e->gOff[cs ^ 1].p, flat ? e->recPre.p : (uint32_t*)0, M);
			if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_group_pass<1, true>), nGC, 256, 0, ga);
			else LAUNCH(e, (k_group_pass<1, false>), nGC, 256, 0, ga);
			ph_end(e, PH_GROUPS);

			ph_begin(e, PH_MEMBERS2);
			LAUNCH(e, k_partition, nTile, 256, 0, e->sid[cs].p, asymCur, M, e->tileB.p, e->dctl, e->sid[cs ^ 1].p);
			ph_end(e, PH_MEMBERS2);
		}
		ctl_pull(e);
		ph_collect(e, (1u << PH_MEMBERS) | (1u << PH_GROUPS) | (1u << PH_MEMBERS2) | e->flat.pending);
		e->flat.pending = 0;
		if (fused) { // the next column's ranges follow from this column's symbol totals (k_col_bases_single on the host)
			uint32_t ms = 0;
			h->gSymBase[0] = h->mSymBase[0] = 0;
			for (int a = 1; a <= 6; ++a) { h->gSymBase[a] = h->mSymBase[a] = ms; if (a < 6) ms += h->memTot[a]; }
			h->gSymBase[7] = h->mSymBase[7] = ms;
			h->Gnext = h->Mnext = ms; h->nrec = M;
		}
		const uint32_t nrec = h->nrec;

		if (m == 1) { // remember where the string's sentinel goes (mr_insert1's return value)
			RB2_CUDA(cudaMemcpyAsync(&e->lastP, e->recP.p, 8, cudaMemcpyDeviceToHost, e->st));
			RB2_CUDA(cudaStreamSynchronize(e->st));
			e->lastBkt = 0;
			for (int b = 0; b < 6; ++b) if (gBkt[b + 1] > gBkt[b]) e->lastBkt = b;
		}
		const rb2_stats_t before = e->stats;
		if (flat) {
			flat_apply_records(e, nrec, M, e->gL[cs ^ 1].p, fused ? (useSizes ? e->recP.p : e->gL[cs].p) : (lean ? e->gL[cs].p : (const int64_t*)0), asymCur);
			if (column_log()) { RB2_CUDA(cudaStreamSynchronize(e->st)); ph_collect(e, e->flat.pending); e->flat.pending = 0; } // per-column times
		} else apply_records(e, nrec, e->gL[cs ^ 1].p);
		e->stats.n_records += nrec;
		++e->stats.n_columns;
		if (FILE *cl = column_log()) // developer aid: RB2_COLLOG=<file> gets one line per column
			fprintf(cl, "col %lld M %u G %u nrec %u nlog %u items %lld deferred %lld ms_half %.3f ms_fast_gen %.3f ms_dir %.3f\n", (long long)col, M, G, nrec, e->nlog,
			        (long long)(e->stats.merge_blocks - before.merge_blocks), (long long)(e->stats.general_items - before.general_items),
			        e->stats.ms_merge - before.ms_merge, e->stats.ms_merge_general - before.ms_merge_general, e->stats.ms_directory - before.ms_directory);

		// ---- advance to the next column ---------------------------------------------
		for (int b = 0; b < 6; ++b) e->bktLen[b] += mBkt[b + 1] - mBkt[b];
		for (int b = 0; b < 8; ++b) { gBkt[b] = h->gSymBase[b]; mBkt[b] = h->mSymBase[b]; }
		G = h->Gnext; M = h->Mnext;
		cs ^= 1;
		asymReady = fused;
		if (fused) std::swap(asymCur, asymNxt);
	}
	if (flat) { // the array stays resident; leaf blocks are rebuilt when something asks for them (ensure_blocks)
		flat_finish(e);
		++e->stats.flat_batches;
	} else pull_totals(e);
	e->stats.n_strings += m;
	e->stats.pool_blocks = e->hctl->poolUsed;
	e->stats.pool_capacity = e->poolCap;
}

// sub-batching: the text output is independent of how a batch is cut, so
// huge inputs are processed as several device batches to bound the per-batch working set.
#define RB2_MAX_BATCH_BYTES (24ll << 30)

#include "rb2_async.inl"

extern "C" void rb2_insert_multi_dev(rb2_engine_t *e, int64_t len, const uint8_t *s_dev)
{
	if (len <= 0) RB2_FATAL("mr_insert_multi: empty batch (mrope.c:268)");
	RB2_NO_CLUSTER(e, "rb2_insert_multi_dev");
	RB2_CUDA(cudaSetDevice(e->dev));
	async_drain(e);
	if (((uintptr_t)s_dev & 15) != 0) RB2_FATAL("device batch must be 16-byte aligned");
	if (len > RB2_MAX_BATCH_BYTES) RB2_FATAL("device-resident batches are limited to %lld bytes; use the host entry point", (long long)RB2_MAX_BATCH_BYTES);
	RB2_CUDA(cudaEventRecord(e->evTot[0], e->st));
	insert_device_batch(e, len, s_dev);
	RB2_CUDA(cudaEventRecord(e->evTot[1], e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, e->evTot[0], e->evTot[1]));
	e->stats.ms_total += ms;
	if (e->as) memcpy(e->as->pub, e->tot, sizeof(e->tot));
}

extern "C" void rb2_insert_multi(rb2_engine_t *e, int64_t len, const uint8_t *s)
{
	if (len <= 0 || s[len - 1] != 0) RB2_FATAL("mr_insert_multi: batch must be non-empty and end with NUL (mrope.c:268)");
	if (e->nChild) { cluster_insert_multi(e, len, s); return; }
	RB2_CUDA(cudaSetDevice(e->dev));
	if (async_enabled(e) && len <= RB2_MAX_BATCH_BYTES) { async_insert_multi(e, len, s); return; }
	async_drain(e);
	RB2_CUDA(cudaEventRecord(e->evTot[0], e->st));
	int64_t off = 0;
	while (off < len) {
		int64_t n = len - off;
		if (n > RB2_MAX_BATCH_BYTES) { // cut behind the last NUL inside the window
			n = RB2_MAX_BATCH_BYTES;
			while (n > 0 && s[off + n - 1] != 0) --n;
			if (n == 0) RB2_FATAL("a single string longer than %lld bytes is not supported", (long long)RB2_MAX_BATCH_BYTES);
		}
		e->sbuf.need((size_t)n + 16);
		ph_begin(e, PH_H2D);
		RB2_CUDA(cudaMemcpyAsync(e->sbuf.p, s + off, (size_t)n, cudaMemcpyHostToDevice, e->st));
		ph_end(e, PH_H2D);
		RB2_CUDA(cudaStreamSynchronize(e->st));
		ph_collect(e, 1u << PH_H2D);
		insert_device_batch(e, n, e->sbuf.p);
		off += n;
	}
	RB2_CUDA(cudaEventRecord(e->evTot[1], e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, e->evTot[0], e->evTot[1]));
	e->stats.ms_total += ms;
	if (e->as) memcpy(e->as->pub, e->tot, sizeof(e->tot));
}

// marginal counts c[b][a] (mr_get_c).  With batches still in flight these are the counts of everything SUBMITTED:
// they follow from the batches alone (k_pair_hist), so the call does not wait for the insertions.
extern "C" void rb2_counts(rb2_engine_t *e, int64_t c[36])
{
	if (e->as && e->as->started) { for (int b = 0; b < 6; ++b) for (int a = 0; a < 6; ++a) c[b * 6 + a] = e->as->pub[b][a]; return; }
	for (int b = 0; b < 6; ++b) for (int a = 0; a < 6; ++a) c[b * 6 + a] = e->tot[b][a];
}

// wait for every queued batch (what any other call does implicitly)
extern "C" void rb2_sync(rb2_engine_t *e) { RB2_CUDA(cudaSetDevice(e->dev)); async_drain(e); }

// what each of the last batches (oldest first, at most `max`) added to the statistics; ms_h2d = its copy.  Lets a
// caller that streams batches back to back see per-batch numbers without waiting between the calls.
extern "C" int rb2_job_history(rb2_engine_t *e, rb2_stats_t *out, int max)
{
	async_drain(e);
	if (!e->as) return 0;
	AsyncState &A = *e->as;
	const int n = (int)std::min<size_t>(A.hist.size(), (size_t)max);
	for (int i = 0; i < n; ++i) {
		const size_t k = A.hist.size() - n + i;
		out[i] = A.hist[k];
		if (k < A.histH2d.size()) { out[i].ms_h2d = A.histH2d[k]; out[i].ms_total += A.histH2d[k]; }
	}
	return n;
}

// device time of a stream of calls: rb2_span_begin() ... calls ... rb2_span_ms() = milliseconds between the first copy
// and the end of the last insertion, measured with CUDA events (copies of later batches overlap earlier insertions)
extern "C" void rb2_span_begin(rb2_engine_t *e)
{
	RB2_CUDA(cudaSetDevice(e->dev));
	async_drain(e);
	cudaEvent_t ev = e->as ? e->as->evSpan[0] : e->evTot[0];
	RB2_CUDA(cudaEventRecord(ev, e->as ? e->as->copySt : e->st));
}
extern "C" double rb2_span_ms(rb2_engine_t *e)
{
	RB2_CUDA(cudaSetDevice(e->dev));
	async_drain(e);
	if (!e->as) return 0.0;
	RB2_CUDA(cudaEventRecord(e->as->evSpan[1], e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, e->as->evSpan[0], e->as->evSpan[1]));
	return ms;
}

extern "C" void rb2_rank2a(rb2_engine_t *e, int64_t x, int64_t y, int64_t cx[6], int64_t cy[6])
{
	if (e->nChild) {
		cluster_rank1(e, x, cx);
		if (cy && y >= 0) cluster_rank1(e, y, cy);
		RB2_CUDA(cudaSetDevice(e->dev));
		return;
	}
	RB2_CUDA(cudaSetDevice(e->dev));
	async_drain(e);
	int64_t total = 0;
	for (int b = 0; b < 6; ++b) total += e->bktLen[b];
	if (x < 0 || x > total || y > total) RB2_FATAL("rank position out of range");
	if (!cy) y = -1;
	ensure_blocks(e);
	LAUNCH(e, k_rank_query, 1, 32, 0, e->pool, e->dir[e->cur], e->nlog, x, y, e->dRankOut, e->dctl);
	RB2_CUDA(cudaMemcpyAsync(e->hRankOut, e->dRankOut, 12 * 8, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	for (int a = 0; a < 6; ++a) { cx[a] = e->hRankOut[a]; if (y >= 0) cy[a] = e->hRankOut[6 + a]; }
}

// n rank queries in one call: out[i*6+a] = #a in BWT[0, x[i]) (the batched form of mr_rank1a)
extern "C" void rb2_rank_batch(rb2_engine_t *e, int64_t n, const int64_t *x, int64_t *out)
{
	RB2_CUDA(cudaSetDevice(e->dev));
	RB2_NO_CLUSTER(e, "rb2_rank_batch");
	if (e->comm) RB2_FATAL("rb2_rank_batch: not available on a sharded engine yet");
	async_drain(e);
	const bool onFlat = e->flat.valid; // after dense batches the resident array answers directly (no leaf blocks needed)
	if (!onFlat) ensure_blocks(e);
	int64_t total = 0;
	for (int b = 0; b < 6; ++b) total += e->bktLen[b];
	const int64_t CH = 1 << 20;
	DevBuf<int64_t> dx, dout;
	dx.need(CH < n ? CH : n); dout.need((size_t)(CH < n ? CH : n) * 6);
	for (int64_t o = 0; o < n; o += CH) {
		const int64_t m = n - o < CH ? n - o : CH;
		for (int64_t i = 0; i < m; ++i) if (x[o + i] < 0 || x[o + i] > total) RB2_FATAL("rank position out of range");
		RB2_CUDA(cudaMemcpyAsync(dx.p, x + o, (size_t)m * 8, cudaMemcpyHostToDevice, e->st));
		if (onFlat) LAUNCH(e, k_flat_rank_batch, std::min<uint32_t>(cdiv(m, 4), (uint32_t)e->nSM * 16), 128, 0, e->flat.s[e->flat.cur].p, e->flat.dir[e->flat.cur].p, (uint32_t)m, dx.p, dout.p);
		else LAUNCH(e, k_rank_batch, std::min<uint32_t>(cdiv(m, 4), (uint32_t)e->nSM * 16), 128, 0, e->pool, e->dir[e->cur], e->nlog, (uint32_t)m, dx.p, dout.p, e->dctl);
		RB2_CUDA(cudaMemcpyAsync(out + o * 6, dout.p, (size_t)m * 48, cudaMemcpyDeviceToHost, e->st));
		RB2_CUDA(cudaStreamSynchronize(e->st));
	}
	dx.release(); dout.release();
}

extern "C" int64_t rb2_num_blocks(rb2_engine_t *e, int bucket)
{
	if (bucket < 0 || bucket >= e->nb) RB2_FATAL("bucket out of range");
	if (e->nChild) return cluster_num_blocks(e, bucket);
	RB2_CUDA(cudaSetDevice(e->dev));
	async_drain(e);
	ensure_blocks(e);
	return (int64_t)e->blkBkt[bucket + 1] - e->blkBkt[bucket];
}

extern "C" int64_t rb2_fetch_blocks(rb2_engine_t *e, int bucket, int64_t first, int64_t n, uint8_t *dst, int64_t *cnt)
{
	if (e->nChild) return cluster_fetch_blocks(e, bucket, first, n, dst, cnt);
	RB2_CUDA(cudaSetDevice(e->dev));
	int64_t nb = rb2_num_blocks(e, bucket);
	if (first < 0 || first > nb) RB2_FATAL("block index out of range");
	if (n > nb - first) n = nb - first;
	if (n <= 0) return 0;
	e->stage.need((size_t)n * RB2_BLK);
	if (cnt) e->stageCnt.need((size_t)n * 6);
	LAUNCH(e, k_gather_blocks, cdiv((uint64_t)n * 32, 256), 256, 0, e->pool, e->dir[e->cur].order, e->blkCnt,
	       (uint32_t)(e->blkBkt[bucket] + first), (uint32_t)n, e->stage.p, cnt ? e->stageCnt.p : (int64_t*)0);
	RB2_CUDA(cudaMemcpyAsync(dst, e->stage.p, (size_t)n * RB2_BLK, cudaMemcpyDeviceToHost, e->st));
	if (cnt) RB2_CUDA(cudaMemcpyAsync(cnt, e->stageCnt.p, (size_t)n * 48, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	return n;
}

extern "C" void rb2_load_blocks(rb2_engine_t *e, int bucket, int64_t n, const uint8_t *src, const int64_t *cnt)
{
	RB2_NO_CLUSTER(e, "restoring an index (rb2_load_blocks)");
	RB2_CUDA(cudaSetDevice(e->dev));
	if (bucket < 0 || bucket > 5) RB2_FATAL("bucket out of range");
	if (n <= 0) return;
	async_drain(e);
	ensure_blocks(e); blocks_edited(e);
	// the bucket's blocks are [blkBkt[b], blkBkt[b+1]); new blocks are inserted at its right end.
	// An initially empty bucket consists of one empty block, which is replaced.
	const uint32_t used = e->hctl->poolUsed;
	reserve_blocks(e, (uint64_t)used + n + 4096);
	e->stage.need((size_t)n * RB2_BLK);
	e->stageCnt.need((size_t)n * 6);
	RB2_CUDA(cudaMemcpyAsync(e->stage.p, src, (size_t)n * RB2_BLK, cudaMemcpyHostToDevice, e->st));
	RB2_CUDA(cudaMemcpyAsync(e->stageCnt.p, cnt, (size_t)n * 48, cudaMemcpyHostToDevice, e->st));
	LAUNCH(e, k_scatter_blocks, cdiv((uint64_t)n * 32, 256), 256, 0, e->pool, e->blkCnt, used, (uint32_t)n, e->stage.p, e->stageCnt.p);
	// splice into the logical order (host side: restore is not on the hot path)
	std::vector<uint32_t> ord(e->nlog);
	RB2_CUDA(cudaMemcpyAsync(ord.data(), e->dir[e->cur].order, (size_t)e->nlog * 4, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	const bool replaceEmpty = e->bktLen[bucket] == 0 && e->blkBkt[bucket + 1] - e->blkBkt[bucket] == 1;
	std::vector<uint32_t> nw;
	nw.reserve(e->nlog + n);
	const uint32_t cutAt = e->blkBkt[bucket + 1];
	for (uint32_t i = 0; i < cutAt - (replaceEmpty ? 1 : 0); ++i) nw.push_back(ord[i]);
	for (int64_t k = 0; k < n; ++k) nw.push_back(used + (uint32_t)k);
	for (uint32_t i = cutAt; i < e->nlog; ++i) nw.push_back(ord[i]);
	const uint32_t delta = (uint32_t)n - (replaceEmpty ? 1 : 0);
	for (int b = bucket + 1; b < 8; ++b) e->blkBkt[b] += delta;
	e->nlog = (uint32_t)nw.size();
	RB2_CUDA(cudaMemcpyAsync(e->dir[e->cur].order, nw.data(), nw.size() * 4, cudaMemcpyHostToDevice, e->st));
	e->hctl->poolUsed = used + (uint32_t)n;
	e->hctl->poolCap = e->poolCap;
	ctl_push(e);
	rebuild_directory(e, false);
	pull_totals(e);
	e->stats.pool_blocks = e->hctl->poolUsed;
	e->stats.pool_capacity = e->poolCap;
	if (e->as) memcpy(e->as->pub, e->tot, sizeof(e->tot));
}

extern "C" void rb2_get_stats(rb2_engine_t *e, rb2_stats_t *st)
{
	async_drain(e);
	*st = e->stats;
	if (e->as) { st->ms_h2d += e->as->h2dMs; st->ms_total += e->as->h2dMs; } // the copies run on the caller's thread
}
extern "C" void rb2_reset_stats(rb2_engine_t *e)
{
	async_drain(e);
	if (e->as) { e->as->h2dMs = 0; e->as->hist.clear(); e->as->histH2d.clear(); }
	for (int r = 0; r < e->nChild; ++r) rb2_reset_stats(e->child[r]);
	int64_t pb = e->stats.pool_blocks, pc = e->stats.pool_capacity;
	memset(&e->stats, 0, sizeof(e->stats));
	e->stats.pool_blocks = pb; e->stats.pool_capacity = pc;
}
extern "C" void *rb2_stream(rb2_engine_t *e) { return (void*)e->st; }
extern "C" void *rb2_dev_alloc(rb2_engine_t *e, int64_t bytes)
{
	void *p = 0;
	RB2_CUDA(cudaSetDevice(e->dev));
	RB2_CUDA(cudaMalloc(&p, (size_t)bytes + 16));
	return p;
}
extern "C" void rb2_dev_free(rb2_engine_t *e, void *p) { RB2_CUDA(cudaSetDevice(e->dev)); RB2_CUDA(cudaFree(p)); }
extern "C" void rb2_dev_upload(rb2_engine_t *e, void *dst, const void *src, int64_t bytes)
{
	RB2_CUDA(cudaSetDevice(e->dev));
	RB2_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
}

#include "rb2_shard.inl"
#include "rb2_cluster.inl"

// ---- single-run and bucket-local entry points behind rope.h ---------------------------------

// global position of the start of bucket b
static int64_t bucket_start(const rb2_engine *e, int b)
{
	int64_t s = 0;
	for (int x = 0; x < b; ++x) s += e->bktLen[x];
	return s;
}

// rope_insert_run (rope.c:114-148): insert rl copies of a behind the first x symbols of
// `bucket`; returns the bucket-local rank(a, x) before the insertion.
extern "C" int64_t rb2_insert_run(rb2_engine_t *e, int bucket, int64_t x, int a, int64_t rl)
{
	RB2_NO_CLUSTER(e, "rope_insert_run");
	RB2_CUDA(cudaSetDevice(e->dev));
	if (bucket < 0 || bucket > 5 || a < 0 || a > 5 || rl <= 0) RB2_FATAL("rb2_insert_run: bad argument");
	async_drain(e);
	if (x < 0 || x > e->bktLen[bucket]) RB2_FATAL("rb2_insert_run: position out of range");
	ensure_blocks(e); blocks_edited(e);
	const uint32_t k = (uint32_t)((rl + RB2_MAXRUN - 1) / RB2_MAXRUN);
	e->recP.need(k); e->recSC.need(k); e->recDst.need(k);
	std::vector<int64_t> P(k, bucket_start(e, bucket) + x);
	std::vector<uint32_t> C(k, (RB2_MAXRUN << 3) | (uint32_t)a), D(k, NONE32);
	C[k - 1] = ((uint32_t)(rl - (int64_t)(k - 1) * RB2_MAXRUN) << 3) | (uint32_t)a;
	D[0] = 0;
	RB2_CUDA(cudaMemcpyAsync(e->recP.p, P.data(), k * 8, cudaMemcpyHostToDevice, e->st));
	RB2_CUDA(cudaMemcpyAsync(e->recSC.p, C.data(), k * 4, cudaMemcpyHostToDevice, e->st));
	RB2_CUDA(cudaMemcpyAsync(e->recDst.p, D.data(), k * 4, cudaMemcpyHostToDevice, e->st));
	Ctl *h = e->hctl;
	for (int b = 0; b < 8; ++b) { h->blkBkt[b] = e->blkBkt[b]; h->recBkt[b] = b <= bucket ? 0 : k; h->cpost[b] = 0; }
	h->poolCap = e->poolCap; h->nItems = 0; h->err = 0; h->overflow = 0; h->failBase = NONE32; h->nTodo = 0; h->todoNext = 0; h->nTodoA = 0; h->todoANext = 0;
	ctl_push(e);
	reserve_items(e, (uint64_t)k + 2);
	apply_records(e, k, e->dRankOut);
	RB2_CUDA(cudaMemcpyAsync(e->hRankOut, e->dRankOut, 8, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	int64_t z = e->hRankOut[0];
	for (int b = 0; b < bucket; ++b) z -= e->tot[b][a];
	pull_totals(e);
	e->stats.pool_blocks = e->hctl->poolUsed; e->stats.pool_capacity = e->poolCap;
	if (e->as) memcpy(e->as->pub, e->tot, sizeof(e->tot));
	return z;
}

// rope_rank2a (rope.c:179-194) on one bucket: counts inside the bucket only
extern "C" void rb2_bucket_rank2a(rb2_engine_t *e, int bucket, int64_t x, int64_t y, int64_t cx[6], int64_t cy[6])
{
	if (bucket < 0 || bucket > 5) RB2_FATAL("bucket out of range");
	async_drain(e);
	if (x < 0 || x > e->bktLen[bucket] || y > e->bktLen[bucket]) RB2_FATAL("rope_rank2a: position out of range");
	const int64_t s = bucket_start(e, bucket);
	int64_t tmp[6];
	rb2_rank2a(e, s + x, y >= 0 && cy ? s + y : -1, cx, cy ? cy : tmp);
	for (int a = 0; a < 6; ++a) {
		int64_t pre = 0;
		for (int b = 0; b < bucket; ++b) pre += e->tot[b][a];
		cx[a] -= pre;
		if (y >= 0 && cy) cy[a] -= pre;
	}
}

// mr_insert1's return value (mrope.c:67): the rank of the sentinel of the string that was
// inserted last as a single-string batch, local to the bucket it went into.
extern "C" int64_t rb2_last_sentinel_rank(rb2_engine_t *e)
{
	RB2_NO_CLUSTER(e, "mr_insert1's return value");
	async_drain(e);
	int64_t cx[6];
	rb2_rank2a(e, e->lastP, -1, cx, 0);
	int64_t z = cx[0];
	for (int b = 0; b < e->lastBkt; ++b) z -= e->tot[b][0];
	return z;
}
