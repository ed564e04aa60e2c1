// rb2_codec.cuh -- the reference's "43+3" run-length byte codec (rle.h:39-75) on the GPU.
//
// A leaf block in HBM is 512 bytes: [uint16 nbytes][run bytes ...], byte-identical to the
// reference's leaf layout (rle.h:36), so blocks can be handed to rle_dec1-based consumers
// (main.c:292-313) and written into .fmr dumps without transcoding.
//
// Run encodings (rle.h:53-75), c = 3-bit symbol, l = length:
//   1 byte   0lllllccc                         l < 16
//   2 bytes  110llccc 10llllll                 l < 256
//   4 bytes  1110lccc 10llllll x3              l < 2^19
//   8 bytes  1111lccc 10llllll x7              l < 2^43   (never produced on the device)
// Device invariant: every run in the pool has l <= RB2_MAXRUN (4-byte form), so a block
// holds < 2^26 symbols and all in-block arithmetic is 32-bit.  Longer logical runs are
// simply stored as adjacent runs of the same symbol, which every consumer of the format
// merges (rld_enc rld0.c:153-161, crlf_write crlf.h:103-114) or prints identically.
//
// One warp decodes one block: lane i owns bytes [16i, 16i+16) (one 128-bit load) and
// decodes the runs that START in its bytes; continuation bytes are 10xxxxxx, so run starts
// are recognisable from any byte (the codec is self-synchronising like UTF-8).
#pragma once
#include "rb2_common.cuh"

#define RB2_BLK 512
#define RB2_FILL 494
#define RB2_MAXRUN ((1u << 19) - 1)
#define RB2_RUNS_STRIDE 17  // per-lane stride (words) of the decoded-run scratch: odd => conflict-free
// decoded run word: byte offset inside the lane's 16 bytes (4 bits) | length (19 bits) | symbol (3 bits)
#define RUN_SYM(r) ((r) & 7u)
#define RUN_LEN(r) (((r) >> 3) & 0x7ffffu)
#define RUN_OFF(r) ((r) >> 22)

struct LaneDec {
	uint32_t nr;     // runs starting in this lane's 16 bytes
	uint32_t len;    // symbols in those runs
	uint32_t c[6];   // per-symbol counts of those runs
};

// byte i (0..19) of the 20-byte window {own 16 bytes, first word of the next lane}
__device__ __forceinline__ uint32_t win_byte(const uint32_t (&W)[5], int i)
{
	return (W[i >> 2] >> ((i & 3) * 8)) & 0xffu;
}

// Decode the runs starting in this lane's bytes into runs[0..nr) as (byteoff << 22 | len << 3 | sym).
// `runs` points at this lane's private slice (stride RB2_RUNS_STRIDE words); `lcs` is a lane-private
// scratch of 6 shared-memory words used to accumulate the per-symbol counts (indexing registers by
// a run-time symbol would cost a 6-way select per run).
__device__ __forceinline__ void decode_lane(const uint4 &own, uint32_t next0, int lane, uint32_t nbytes,
                                            uint32_t *runs, uint32_t *lcs, LaneDec &d, uint32_t &err)
{
	const uint32_t W[5] = { own.x, own.y, own.z, own.w, next0 };
	uint32_t nr = 0, tot = 0;
	const int lim = (int)nbytes + 2 - lane * 16; // byte i of this lane is a run byte iff i < lim (and i >= 2 in lane 0)
#pragma unroll
	for (int a = 0; a < 6; ++a) lcs[a] = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		const uint32_t b = win_byte(W, i);
		const bool start = (i >= 2 || lane > 0) && i < lim && (b & 0xC0u) != 0x80u;
		if (start) {
			uint32_t l;
			if (b < 0x80u) l = b >> 3;
			else if (b < 0xE0u) l = ((b & 0x18u) << 3) | (win_byte(W, i + 1) & 0x3fu);
			else {
				if (b >= 0xF0u) err |= RB2_ERR_RUN8;
				l = ((b & 0x08u) << 15) | ((win_byte(W, i + 1) & 0x3fu) << 12)
				  | ((win_byte(W, i + 2) & 0x3fu) << 6) | (win_byte(W, i + 3) & 0x3fu);
			}
			const uint32_t s = b & 7u;
			runs[nr++] = ((uint32_t)i << 22) | (l << 3) | s;
			tot += l;
			lcs[s] += l;
		}
	}
	d.nr = nr; d.len = tot;
#pragma unroll
	for (int a = 0; a < 6; ++a) d.c[a] = lcs[a];
}

// Whole-warp decode of one block.  On return, for this lane:
//   d        runs starting in the lane
//   basePos  symbols in front of the lane's first run (exclusive scan of d.len)
//   baseCnt  per-symbol counts in front of the lane (exclusive scan of d.c)
//   blkLen   symbols in the block, blkCnt[6] per-symbol totals (same in all lanes)
// cntScratch: 32 x 7 shared-memory words (lane-private slices).
__device__ __forceinline__ void warp_decode_block(const uint8_t *blk, int lane, uint32_t *runsWarp, uint32_t *cntScratch,
                                                  LaneDec &d, uint32_t &basePos, uint32_t (&baseCnt)[6],
                                                  uint32_t &blkLen, uint32_t (&blkCnt)[6], uint32_t &nbytes, uint32_t &err, uint4 &own)
{
	// plain (coherent) load: k_merge_blocks rewrites the same block in place later on
	own = *(reinterpret_cast<const uint4*>(blk) + lane);
	nbytes = __shfl_sync(FULLMASK, own.x, 0) & 0xffffu;
	uint32_t next0 = __shfl_down_sync(FULLMASK, own.x, 1);
	if (lane == 31) next0 = 0;
	decode_lane(own, next0, lane, nbytes, runsWarp + lane * RB2_RUNS_STRIDE, cntScratch + lane * 7, d, err);
	uint32_t incl = warp_incl_scan(d.len, lane);
	basePos = incl - d.len;
	blkLen = __shfl_sync(FULLMASK, incl, 31);
	if (blkLen < 65536u) { // every count fits 16 bits: scan two symbols per word
#pragma unroll
		for (int a = 0; a < 6; a += 2) {
			const uint32_t v = d.c[a] | (d.c[a + 1] << 16);
			const uint32_t x = warp_incl_scan(v, lane);
			const uint32_t ex = x - v, t = __shfl_sync(FULLMASK, x, 31);
			baseCnt[a] = ex & 0xffffu; baseCnt[a + 1] = ex >> 16;
			blkCnt[a] = t & 0xffffu; blkCnt[a + 1] = t >> 16;
		}
	} else {
#pragma unroll
		for (int a = 0; a < 6; ++a) {
			uint32_t x = warp_incl_scan(d.c[a], lane);
			baseCnt[a] = x - d.c[a];
			blkCnt[a] = __shfl_sync(FULLMASK, x, 31);
		}
	}
	__syncwarp();
}

__device__ __forceinline__ int run_nbytes(uint32_t l) { return l < 16u ? 1 : (l < 256u ? 2 : 4); }

// encode one run (l <= RB2_MAXRUN) at p; returns bytes written (rle_enc1, rle.h:53-75)
__device__ __forceinline__ int enc_run(uint8_t *p, uint32_t s, uint32_t l)
{
	if (l < 16u) { p[0] = (uint8_t)(l << 3 | s); return 1; }
	if (l < 256u) { p[0] = (uint8_t)(0xC0u | (l >> 6) << 3 | s); p[1] = (uint8_t)(0x80u | (l & 0x3fu)); return 2; }
	p[0] = (uint8_t)(0xE0u | (l >> 18) << 3 | s);
	p[1] = (uint8_t)(0x80u | ((l >> 12) & 0x3fu));
	p[2] = (uint8_t)(0x80u | ((l >> 6) & 0x3fu));
	p[3] = (uint8_t)(0x80u | (l & 0x3fu));
	return 4;
}
