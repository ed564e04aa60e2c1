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
// One warp decodes one block.  Lane i owns bytes [16i, 16i+16) (one 128-bit load) and accounts for
// the runs that START in its bytes.  The block image is kept in shared memory (`img`, 512 + 16
// bytes, the 16 extra bytes zero) so that individual runs can be parsed on demand.
#define RB2_IMG_BYTES (RB2_BLK + 16)

struct LaneDec {
	uint32_t nr;     // runs starting in this lane's 16 bytes
	uint32_t len;    // symbols in those runs
	uint32_t fb;     // byte offset (0..15) of the first run start in the lane, 16 if none
	uint32_t c[6];   // per-symbol counts of those runs
};

// parse the run whose first byte is img[bp]: symbol, length, encoded size
__device__ __forceinline__ void parse_run(const uint8_t *img, uint32_t bp, uint32_t &sym, uint32_t &len, uint32_t &nb)
{
	const uint32_t b = img[bp];
	sym = b & 7u;
	if (b < 0x80u) { len = b >> 3; nb = 1; }
	else if (b < 0xE0u) { len = ((b & 0x18u) << 3) | (img[bp + 1] & 0x3fu); nb = 2; }
	else {
		len = ((b & 0x08u) << 15) | ((img[bp + 1] & 0x3fu) << 12) | ((img[bp + 2] & 0x3fu) << 6) | (img[bp + 3] & 0x3fu);
		nb = 4;
	}
}

// Run-serial accounting of one lane (any mix of 1/2/4-byte runs), parsed from the shared-memory
// image.  Deliberately a rolled loop: it is the cold path for short-run data and must not bloat the
// instruction footprint of the kernels that inline it.  `lcs` = 6 lane-private shared-memory words
// (indexing registers by a run-time symbol would cost a 6-way select per run).
// Accounts for the runs starting in image bytes [base+lo, base+hi); `none` = value of fb when no run starts there.
__device__ __noinline__ uint64_t decode_span_serial(const uint8_t *img, uint32_t base, uint32_t lo, uint32_t hi, uint32_t none, uint32_t *lcs)
{
	// returns nr | fb << 8 | err << 16 | (uint64)len << 32; per-symbol counts are left in lcs[0..5]
	uint32_t i = lo, nr = 0, tot = 0, err = 0;
#pragma unroll
	for (int a = 0; a < 6; ++a) lcs[a] = 0;
	while (i < hi && (img[base + i] & 0xC0u) == 0x80u) ++i; // tail of a run that started in the previous lane
	const uint32_t fb = i < hi ? i : none;
#pragma unroll 1
	while (i < hi) {
		uint32_t s, l, nb;
		parse_run(img, base + i, s, l, nb);
		if (img[base + i] >= 0xF0u) err = RB2_ERR_RUN8;
		++nr; tot += l; lcs[s] += l;
		i += nb;
	}
	return (uint64_t)nr | (uint64_t)fb << 8 | (uint64_t)err << 16 | (uint64_t)tot << 32;
}

__device__ __forceinline__ uint64_t decode_lane_serial(const uint8_t *img, int lane, uint32_t nbytes, uint32_t *lcs)
{
	const int lim = (int)nbytes + 2 - lane * 16; // byte i of this lane is a run byte iff i < lim (and i >= 2 in lane 0)
	const uint32_t hi = lim > 16 ? 16u : (lim < 0 ? 0u : (uint32_t)lim);
	return decode_span_serial(img, lane * 16, lane == 0 ? 2u : 0u, hi, 16u, lcs);
}

// exclusive scans of six per-lane counts that do not fit 16 bits (blocks with very long runs): cold.
// lcs (lane-private shared-memory slice): in = the lane's counts, out = their exclusive prefixes.
__device__ __noinline__ void wide_count_scans(int lane, uint32_t *lcs)
{
#pragma unroll 1
	for (int a = 0; a < 6; ++a) {
		const uint32_t c = lcs[a];
		lcs[a] = warp_incl_scan(c, lane) - c;
	}
	__syncwarp();
}

// SIMD accounting of a lane whose run bytes are all 1-byte runs (every byte < 0x80; bytes behind
// the block's last run are zero by invariant and count as nothing).  Per 4 bytes: lengths =
// (w >> 3) & 0x0f.., the four symbols become a PRMT selector, and for each symbol x a one-hot byte
// table looked up through PRMT gives 0/1 weights for a 4-way dot product (DP4A) with the lengths.
__device__ __forceinline__ void decode_lane_simd(const uint4 &own, int lane, uint32_t nbytes, LaneDec &d)
{
	uint32_t w[4] = { own.x, own.y, own.z, own.w };
	if (lane == 0) w[0] &= 0xffff0000u; // the 2-byte header is not run data
	const int lim = (int)nbytes + 2 - lane * 16;
	const int hi = lim < 0 ? 0 : (lim > 16 ? 16 : lim), lo = lane == 0 ? 2 : 0;
	if (hi < 16) { // the one lane holding the end of the data: ignore whatever sits behind it
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int keep = hi - 4 * j;
			w[j] = keep >= 4 ? w[j] : (keep <= 0 ? 0u : w[j] & ((1u << (8 * keep)) - 1u));
		}
	}
	d.nr = hi > lo ? hi - lo : 0;
	d.fb = hi > lo ? lo : 16;
	uint32_t tot = 0, c[6] = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const uint32_t lens = (w[j] >> 3) & 0x0f0f0f0fu;
		const uint32_t sy = w[j] & 0x07070707u;
		const uint32_t t = sy | (sy >> 4);
		const uint32_t sel = (t & 0xffu) | ((t >> 8) & 0xff00u);
		tot = __dp4a(lens, 0x01010101u, tot);
		c[0] = __dp4a(lens, __byte_perm(0x00000001u, 0u, sel), c[0]);
		c[1] = __dp4a(lens, __byte_perm(0x00000100u, 0u, sel), c[1]);
		c[2] = __dp4a(lens, __byte_perm(0x00010000u, 0u, sel), c[2]);
		c[3] = __dp4a(lens, __byte_perm(0x01000000u, 0u, sel), c[3]);
		c[4] = __dp4a(lens, __byte_perm(0u, 0x00000001u, sel), c[4]);
		c[5] = __dp4a(lens, __byte_perm(0u, 0x00000100u, sel), c[5]);
	}
	d.len = tot;
#pragma unroll
	for (int a = 0; a < 6; ++a) d.c[a] = c[a];
}

// Whole-warp decode of one block.  On return, for this lane:
//   d        accounting of the runs starting in the lane
//   basePos  symbols in front of the lane's first run (exclusive scan of d.len)
//   baseCnt  per-symbol counts in front of the lane (exclusive scan of d.c)
//   blkLen   symbols in the block, blkCnt[6] per-symbol totals (same in all lanes)
// img: RB2_IMG_BYTES of shared memory receiving the block image; cntScratch: 32 x 7 words.
__device__ __forceinline__ void warp_decode_block(const uint8_t *blk, int lane, uint8_t *img, uint32_t *cntScratch,
                                                  LaneDec &d, uint32_t &basePos, uint32_t (&baseCnt)[6],
                                                  uint32_t &blkLen, uint32_t (&blkCnt)[6], uint32_t &nbytes, uint32_t &err, uint4 &own,
                                                  uint32_t *pureMask = 0)
{
	// plain (coherent) load: the merge kernels rewrite the same block in place later on
	own = *(reinterpret_cast<const uint4*>(blk) + lane);
	reinterpret_cast<uint4*>(img)[lane] = own;
	if (lane == 0) reinterpret_cast<uint4*>(img)[32] = make_uint4(0, 0, 0, 0);
	nbytes = __shfl_sync(FULLMASK, own.x, 0) & 0xffffu;
	__syncwarp(); // the image is complete: serial lanes read runs that reach into the next lane
	// a lane is "pure" if none of its run bytes has the top bit set (no multi-byte run touches it)
	uint32_t any = own.x | own.y | own.z | own.w;
	if (lane == 0) any = (own.x & 0xffff0000u) | own.y | own.z | own.w;
	const int lim = (int)nbytes + 2 - lane * 16;
	const bool pure = (any & 0x80808080u) == 0 || lim <= 0;
	if (pureMask) *pureMask = __ballot_sync(FULLMASK, pure);
	if (pure) decode_lane_simd(own, lane, nbytes, d);
	else {
		uint32_t *lcs = cntScratch + lane * 7;
		const uint64_t r = decode_lane_serial(img, lane, nbytes, lcs);
		d.nr = (uint32_t)r & 0xffu; d.fb = ((uint32_t)r >> 8) & 0xffu; err |= ((uint32_t)r >> 16) & 0xffu; d.len = (uint32_t)(r >> 32);
#pragma unroll
		for (int a = 0; a < 6; ++a) d.c[a] = lcs[a];
	}
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
		uint32_t *lcs = cntScratch + lane * 7;
		__syncwarp();
#pragma unroll
		for (int a = 0; a < 6; ++a) lcs[a] = d.c[a];
		wide_count_scans(lane, lcs);
#pragma unroll
		for (int a = 0; a < 6; ++a) { baseCnt[a] = lcs[a]; }
		__syncwarp();
#pragma unroll
		for (int a = 0; a < 6; ++a) blkCnt[a] = cntScratch[31 * 7 + a] + __shfl_sync(FULLMASK, d.c[a], 31);
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
