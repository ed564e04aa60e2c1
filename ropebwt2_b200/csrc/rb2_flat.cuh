// rb2_flat.cuh -- the DENSE regime of the BCR column step: stream-rewrite of a fixed-width BWT.
//
// When a batch inserts into (nearly) every leaf block in every column -- short reads, the headline
// workload: 100 M records per column into <= 20 M blocks -- updating run-length coded blocks record
// by record is instruction bound (profiles/README.md).  For such a batch the engine keeps the BWT,
// for the duration of the batch only, as ONE flat array of nt6 codes (one byte per symbol, all six
// buckets concatenated) plus a directory of per-symbol counts in front of every FT_DIR-th symbol,
// and one column becomes a single streaming pass (k_flat_merge):
//
//   new[P_r + pre_r .. + c_r) = symbol of record r     (pre_r = members in front of record r, i.e. the
//   old symbol i moves to i + #record symbols with P<=i  symbols this column inserts in front of it)
//
// which is exactly the stable merge rope_insert_run performs one run at a time (rope.c:114-148;
// new symbols go in FRONT of the old symbol at the same position, mrope.c:206-218), and
// rank(a, P_r) -- the return value of rope_insert_run / rle_insert_cached (rle.c:10-89) -- is
// directory[P_r / FT_DIR][a] + a count over < FT_DIR + 4096 symbols held in shared memory.
// Output-stationary: CTA t produces new[t*4096, (t+1)*4096) with aligned 128-bit stores; the old
// symbols it needs are one contiguous range.  At the end of the batch the array is re-encoded into
// leaf blocks of the reference's format (k_flat_encode), so everything outside the batch (iterator,
// dump, rank queries, sparse batches) sees the usual block pool.
#pragma once
#include "rb2_codec.cuh"

#define FT_OUT   4096  // output symbols per CTA of k_flat_merge: 256 threads x 16 bytes
#define FT_DIR   1024  // directory granularity (48 bytes of counts per 1024 symbols)
#define FT_SUB   (FT_OUT / FT_DIR)
#define FT_OLDMAX (FT_OUT + FT_DIR) // old symbols one CTA can need: from the directory tile of its first one
#define FT_PAD   (FT_OLDMAX + 64)   // readable slack behind a flat array
#define FE_CHUNK 64    // flat -> blocks: symbols encoded by one thread (<= 64 bytes of runs)
#define FE_T     (RB2_FILL - FE_CHUNK + 1) // block k of a bucket takes the chunks that start in bytes [k*FE_T, (k+1)*FE_T)

// ---- per-symbol counting without per-symbol compares ---------------------------------------------
// nt6 codes: $=000 A=001 C=010 G=011 T=100 N=101.  For four codes in a word, the bytes of
//   m0 = bit0, m1 = bit1, m2 = bit2, m01 = bit0&bit1 (G), m02 = bit0&bit2 (N)
// are 0/1 and can be summed byte-wise over up to 255 words; five horizontal sums then give
//   N = s02, G = s01, A = s0 - s01 - s02, C = s1 - s01, T = s2 - s02, $ = n - (A+C+G+T+N).
// "Raw" counts (s0, s1, s2, s01, s02, n) are linear, so prefix sums are taken on them and converted
// only where a symbol count is needed.
struct Raw6 { uint32_t s0, s1, s2, s01, s02, n; };

__device__ __forceinline__ void raw_add_words(const uint32_t *w, int nw, uint32_t (&acc)[5])
{
#pragma unroll
	for (int j = 0; j < nw; ++j) {
		const uint32_t x = w[j];
		const uint32_t m0 = x & 0x01010101u, m1 = (x >> 1) & 0x01010101u, m2 = (x >> 2) & 0x01010101u;
		acc[0] += m0; acc[1] += m1; acc[2] += m2; acc[3] += m0 & m1; acc[4] += m0 & m2;
	}
}
__device__ __forceinline__ uint32_t hsum4(uint32_t x) { return __dp4a(x, 0x01010101u, 0u); }

__device__ __forceinline__ uint32_t raw_symbol(const Raw6 &r, uint32_t a)
{
	const uint32_t A = r.s0 - r.s01 - r.s02, Cc = r.s1 - r.s01, G = r.s01, T = r.s2 - r.s02, N = r.s02;
	switch (a) {
	case 0: return r.n - (A + Cc + G + T + N);
	case 1: return A;
	case 2: return Cc;
	case 3: return G;
	case 4: return T;
	default: return N;
	}
}

// two 64-bit words hold the six raw counts in 21-bit fields (CTA-level sums stay below 2^21)
__device__ __forceinline__ void raw_pack(const Raw6 &r, uint64_t (&p)[2])
{
	p[0] = (uint64_t)r.s0 | (uint64_t)r.s1 << 21 | (uint64_t)r.s2 << 42;
	p[1] = (uint64_t)r.s01 | (uint64_t)r.s02 << 21 | (uint64_t)r.n << 42;
}
__device__ __forceinline__ Raw6 raw_unpack(const uint64_t (&p)[2])
{
	Raw6 r;
	r.s0 = (uint32_t)p[0] & 0x1fffffu; r.s1 = (uint32_t)(p[0] >> 21) & 0x1fffffu; r.s2 = (uint32_t)(p[0] >> 42) & 0x1fffffu;
	r.s01 = (uint32_t)p[1] & 0x1fffffu; r.s02 = (uint32_t)(p[1] >> 21) & 0x1fffffu; r.n = (uint32_t)(p[1] >> 42) & 0x1fffffu;
	return r;
}

// raw counts of the first nb (0..16) bytes of four words
__device__ __forceinline__ Raw6 raw_count16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t nb)
{
	uint32_t w[4] = { w0, w1, w2, w3 };
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const uint32_t kb = nb > (uint32_t)j * 4 ? nb - j * 4 : 0;
		w[j] = kb >= 4 ? w[j] : (kb ? w[j] & ((1u << (kb * 8)) - 1u) : 0u);
	}
	uint32_t acc[5] = { 0, 0, 0, 0, 0 };
	raw_add_words(w, 4, acc);
	Raw6 r;
	r.s0 = hsum4(acc[0]); r.s1 = hsum4(acc[1]); r.s2 = hsum4(acc[2]); r.s01 = hsum4(acc[3]); r.s02 = hsum4(acc[4]);
	r.n = nb;
	return r;
}

// ---- tile -> record ranges ---------------------------------------------------------------------------
// tileR0[t] = first record whose output run starts at or behind t*FT_OUT (key_r = P_r + pre_r); one
// thread per record (plus a virtual one behind the last) fills the tiles between its predecessor and itself.
__global__ void __launch_bounds__(256) k_flat_splits(const int64_t *recP, const uint32_t *recPre, uint32_t R, uint64_t nTiles, uint32_t *tileR0)
{
	const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (r > R) return;
	const int64_t tPrev = r == 0 ? -1 : (int64_t)((uint64_t)(recP[r - 1] + recPre[r - 1]) / FT_OUT);
	const int64_t t = r == R ? (int64_t)nTiles : (int64_t)((uint64_t)(recP[r] + recPre[r]) / FT_OUT);
	for (int64_t tt = tPrev + 1; tt <= t; ++tt) tileR0[tt] = (uint32_t)r;
}

// Per-tile geometry, one thread per tile boundary t (output position min(t*FT_OUT, nNew)): the first old
// symbol that lands at or behind it, the first record that starts there, and the record run that reaches
// across it from the left.  Computed ahead of k_flat_merge so that its CTAs start with two independent loads.
struct alignas(16) TileDesc { uint64_t i0; uint32_t r0, carry; }; // carry = (symbols of the crossing run behind the boundary, capped at FT_OUT) << 3 | symbol

__global__ void __launch_bounds__(256) k_flat_geo(const int64_t *recP, const uint32_t *recPre, const uint32_t *recSC, const uint32_t *tileR0,
                                                  uint64_t nTiles, uint64_t nNew, TileDesc *desc)
{
	const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (t > nTiles) return;
	const uint64_t o0 = t * FT_OUT < nNew ? t * FT_OUT : nNew;
	const uint32_t r0 = tileR0[t];
	uint64_t before = 0; uint32_t carry = 0;
	if (r0 > 0) {
		const uint64_t pre = recPre[r0 - 1], key = (uint64_t)recP[r0 - 1] + pre; const uint32_t sc = recSC[r0 - 1];
		const uint64_t end = key + (sc >> 3);
		if (end > o0) { before = pre + (o0 - key); const uint64_t rem = end - o0; carry = (uint32_t)(rem < FT_OUT ? rem : FT_OUT) << 3 | (sc & 7u); }
		else before = pre + (sc >> 3);
	}
	TileDesc d; d.i0 = o0 - before; d.r0 = r0; d.carry = carry;
	desc[t] = d;
}

struct FlatArgs {
	const uint8_t *oldS; const int64_t *oldDir;   // old array, counts in front of every FT_DIR-th old symbol
	uint8_t *newS; uint64_t nNew; uint32_t *newTileCnt; // new array and its raw per-FT_DIR-tile symbol counts
	const int64_t *recP; const uint32_t *recPre, *recSC, *recDst; uint32_t R;
	const TileDesc *desc;
	int64_t *gLNext; const Ctl *ctl;
	// sharded engines: records carry whole-index positions; bucket b of this rank sits recOff[b*7+6]
	// symbols (recOff[b*7+a] symbols a) further right in the whole index than in the local array
	const int64_t *recOff; int nb;
};

// raw counts of 16 bytes, no masking (callers only use prefixes that end inside valid data)
__device__ __forceinline__ void raw_acc16(const uint4 &x, uint32_t (&acc)[5])
{
	const uint32_t w[4] = { x.x, x.y, x.z, x.w };
	raw_add_words(w, 4, acc);
}
// six raw counts as three words of two 16-bit fields (sums stay below 2^16 inside one tile)
__device__ __forceinline__ void raw_pack16(const uint32_t (&acc)[5], uint32_t n, uint32_t (&p)[3])
{
	p[0] = hsum4(acc[0]) | hsum4(acc[1]) << 16; p[1] = hsum4(acc[2]) | hsum4(acc[3]) << 16; p[2] = hsum4(acc[4]) | n << 16;
}
__device__ __forceinline__ Raw6 raw_unpack16(uint32_t p0, uint32_t p1, uint32_t p2)
{
	Raw6 r = { p0 & 0xffffu, p0 >> 16, p1 & 0xffffu, p1 >> 16, p2 & 0xffffu, p2 >> 16 };
	return r;
}

#define FT_NCH (FT_OLDMAX / 16 + 2)   // 16-byte chunks of old symbols one tile can hold
struct FlatSmem {
	uint4    old4[FT_NCH];               // the old symbols this tile needs, from a directory tile boundary
	uint32_t chunkPre[FT_NCH][3];        // raw counts in front of every 16-byte chunk of old4 (16-bit fields)
	uint16_t sKey[FT_OUT + 1], sPre[FT_OUT + 1]; // staged records: run start inside the tile, record symbols in front of it
	uint16_t sLS[FT_OUT + 1];            // (run length inside the tile) << 3 | symbol
	uint16_t sFirst[FT_OUT / 16 + 2];    // first staged record that starts at or behind each 16-symbol output chunk
	uint32_t warpTot[8][3];
	uint32_t recCnt[FT_SUB][6];          // symbols the records put into each FT_DIR sub-tile
	uint32_t subX[FT_SUB + 1];           // old symbols (local index) in front of each sub-tile
};

__global__ void __launch_bounds__(256) k_flat_merge(FlatArgs A)
{
	extern __shared__ __align__(16) uint8_t smraw[];
	FlatSmem &S = *reinterpret_cast<FlatSmem*>(smraw);
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint64_t o0 = (uint64_t)blockIdx.x * FT_OUT;
	const uint32_t tileLen = A.nNew - o0 < FT_OUT ? (uint32_t)(A.nNew - o0) : FT_OUT;
	const TileDesc d0 = A.desc[blockIdx.x], d1 = A.desc[blockIdx.x + 1];
	const uint32_t r0 = d0.r0, r1 = d1.r0;
	const uint32_t carrySym = d0.carry & 7u, carryLen = (d0.carry >> 3) < tileLen ? (d0.carry >> 3) : tileLen;
	const uint64_t i0 = d0.i0;                     // old symbols [i0, i1) land in this tile
	const uint64_t a0 = i0 & ~(uint64_t)(FT_DIR - 1);
	const uint32_t loadLen = (uint32_t)(d1.i0 - a0), skip = (uint32_t)(i0 - a0);
	const uint32_t recIn = tileLen - (loadLen - skip); // record symbols inside the tile
	const uint64_t before = o0 - i0;               // record symbols in front of the tile
	const uint32_t nLoad = (loadLen >> 4) + 2;     // chunks that are read (<= FT_NCH)
	const uint32_t nCarry = carryLen ? 1u : 0u, nS = nCarry + (r1 - r0);
	constexpr int NCNT = 160;                      // threads (5 warps) that load + count; the other 3 warps stage records

	// ---- phase A: old symbols -> shared memory + raw counts per chunk pair | records -> shared memory -------
	uint32_t p[3] = { 0, 0, 0 }, q[3] = { 0, 0, 0 }, inc[3] = { 0, 0, 0 };
	if (tid < NCNT) {
		// thread j owns chunks 2j, 2j+1 (bytes behind loadLen are whatever follows in the array: the
		// prefixes that include them are never used)
		const uint4 *src = reinterpret_cast<const uint4*>(A.oldS + a0);
		if ((uint32_t)tid * 2 < nLoad) {
			const uint4 x = src[tid * 2], y = src[tid * 2 + 1];
			S.old4[tid * 2] = x; S.old4[tid * 2 + 1] = y;
			uint32_t ax[5] = { 0, 0, 0, 0, 0 }, ay[5] = { 0, 0, 0, 0, 0 };
			raw_acc16(x, ax); raw_acc16(y, ay);
			raw_pack16(ax, 16, q);
			raw_pack16(ay, 16, p);
#pragma unroll
			for (int k = 0; k < 3; ++k) p[k] += q[k];
		}
#pragma unroll
		for (int k = 0; k < 3; ++k) inc[k] = p[k];
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
			for (int k = 0; k < 3; ++k) { const uint32_t y = __shfl_up_sync(FULLMASK, inc[k], o); if (lane >= o) inc[k] += y; }
		}
		if (lane == 31) { S.warpTot[wid][0] = inc[0]; S.warpTot[wid][1] = inc[1]; S.warpTot[wid][2] = inc[2]; }
	} else {
		if (tid == NCNT && nCarry) { S.sKey[0] = 0; S.sPre[0] = 0; S.sLS[0] = (uint16_t)((carryLen << 3) | carrySym); }
		if (tid - NCNT < FT_SUB * 6) (&S.recCnt[0][0])[tid - NCNT] = 0;
		for (uint32_t k = tid - NCNT; k < r1 - r0; k += 256 - NCNT) {
			const uint32_t r = r0 + k;
			const uint32_t pre = A.recPre[r], sc = A.recSC[r];
			const uint32_t key = (uint32_t)((uint64_t)A.recP[r] + pre - o0);
			uint32_t len = sc >> 3;
			if (len > FT_OUT - key) len = FT_OUT - key;
			S.sKey[nCarry + k] = (uint16_t)key; S.sPre[nCarry + k] = (uint16_t)(pre - before); S.sLS[nCarry + k] = (uint16_t)((len << 3) | (sc & 7u));
		}
	}
	__syncthreads();
	// ---- phase B: prefix in front of every chunk | first record per output chunk, record symbols per sub-tile ----
	if (tid < NCNT) {
		if ((uint32_t)tid * 2 < nLoad) {
			uint32_t base[3] = { 0, 0, 0 };
			for (int w = 0; w < wid; ++w) { base[0] += S.warpTot[w][0]; base[1] += S.warpTot[w][1]; base[2] += S.warpTot[w][2]; }
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const uint32_t ex = base[k] + inc[k] - p[k];
				S.chunkPre[tid * 2][k] = ex; S.chunkPre[tid * 2 + 1][k] = ex + q[k];
			}
		}
	} else {
		for (uint32_t k = tid - NCNT; k <= nS; k += 256 - NCNT) {
			const uint32_t cLo = k == 0 ? 0u : ((uint32_t)S.sKey[k - 1] >> 4) + 1, cHi = k == nS ? FT_OUT / 16 : (uint32_t)S.sKey[k] >> 4;
			for (uint32_t c = cLo; c <= cHi; ++c) S.sFirst[c] = (uint16_t)k;
			if (k < nS) { // symbols the record puts into each sub-tile
				uint32_t key = S.sKey[k], len = (uint32_t)S.sLS[k] >> 3; const uint32_t a = S.sLS[k] & 7u;
				while (len) {
					const uint32_t sb = key / FT_DIR, room = (sb + 1) * FT_DIR - key, n = len < room ? len : room;
					atomicAdd(&S.recCnt[sb][a], n);
					key += n; len -= n;
				}
			}
		}
	}
	__syncthreads();
	// raw counts in front of local old index x
	auto prefix_at = [&](uint32_t x) -> Raw6 {
		const uint32_t c = x >> 4;
		Raw6 rr = raw_unpack16(S.chunkPre[c][0], S.chunkPre[c][1], S.chunkPre[c][2]);
		if (x & 15u) {
			const uint4 v = S.old4[c];
			const Raw6 part = raw_count16(v.x, v.y, v.z, v.w, x & 15u);
			rr.s0 += part.s0; rr.s1 += part.s1; rr.s2 += part.s2; rr.s01 += part.s01; rr.s02 += part.s02; rr.n += part.n;
		}
		return rr;
	};
	// ---- phase C: assemble 16 output symbols per thread ------------------------------------------------
	const uint32_t rel = tid * 16;
	const uint32_t k0 = S.sFirst[tid];  // first entry with sKey >= rel
	uint32_t runRem = 0, runSym = 0, oldIdx;
	if (k0 > 0 && (uint32_t)S.sKey[k0 - 1] + ((uint32_t)S.sLS[k0 - 1] >> 3) > rel) {
		runRem = (uint32_t)S.sKey[k0 - 1] + ((uint32_t)S.sLS[k0 - 1] >> 3) - rel; runSym = S.sLS[k0 - 1] & 7u;
		oldIdx = (uint32_t)S.sKey[k0 - 1] - S.sPre[k0 - 1] + skip;
	} else {
		const uint32_t preAt = k0 < nS ? S.sPre[k0] : recIn;
		oldIdx = rel - preAt + skip;
	}
	if ((tid & (FT_DIR / 16 - 1)) == 0) S.subX[tid / (FT_DIR / 16)] = oldIdx < loadLen ? oldIdx : loadLen;
	if (tid == 0) S.subX[FT_SUB] = loadLen;
	uint32_t ow[4];
	{
		// 16 old symbols from oldIdx on (unaligned)
		const uint32_t *wp = reinterpret_cast<const uint32_t*>(S.old4) + (oldIdx >> 2);
		const uint32_t sh = (oldIdx & 3) * 8;
		const uint32_t x0 = wp[0], x1 = wp[1], x2 = wp[2], x3 = wp[3], x4 = wp[4];
		ow[0] = __funnelshift_r(x0, x1, sh); ow[1] = __funnelshift_r(x1, x2, sh); ow[2] = __funnelshift_r(x2, x3, sh); ow[3] = __funnelshift_r(x3, x4, sh);
	}
	const uint32_t nextKey = k0 < nS ? S.sKey[k0] : 0xffffffffu;
	const uint32_t nextKey2 = k0 + 1 < nS ? S.sKey[k0 + 1] : 0xffffffffu;
	if (runRem >= 16) {                             // inside one long run
		ow[0] = ow[1] = ow[2] = ow[3] = runSym * 0x01010101u;
	} else if (runRem == 0 && nextKey < rel + 16 && nextKey2 >= rel + 16 && ((uint32_t)S.sLS[k0] >> 3) == 1) {
		// the common case late in a batch: exactly one single-symbol record among these 16 symbols.
		// out[j] = V0[j] (j < p), symbol (j == p), V0[j-1] (j > p)
		const uint32_t pp = nextKey - rel, pw = pp >> 2, pb = (pp & 3) * 8, sy = S.sLS[k0] & 7u;
		const uint32_t w1[4] = { ow[0] << 8, __funnelshift_l(ow[0], ow[1], 8), __funnelshift_l(ow[1], ow[2], 8), __funnelshift_l(ow[2], ow[3], 8) };
		const uint32_t lowMask = (1u << pb) - 1u;     // bytes in front of p inside its word
#pragma unroll
		for (int w = 0; w < 4; ++w) {
			const uint32_t mid = (ow[w] & lowMask) | (sy << pb) | (w1[w] & ~((lowMask << 8) | 0xffu));
			ow[w] = (uint32_t)w < pw ? ow[w] : ((uint32_t)w > pw ? w1[w] : mid);
		}
	} else if (runRem || nextKey < rel + 16) {      // records start (or a run ends) inside these 16 symbols
		typedef unsigned __int128 u128;
		auto fill = [](uint32_t sy) -> u128 { const uint64_t f = 0x0101010101010101ull * sy; return ((u128)f << 64) | f; };
		u128 O = ((u128)(((uint64_t)ow[3] << 32) | ow[2]) << 64) | (((uint64_t)ow[1] << 32) | ow[0]); // byte 0 = next old symbol
		u128 R = 0;
		uint32_t pos = 0, k = k0;
		if (runRem) { R = fill(runSym) & ((((u128)1) << (8 * runRem)) - 1); pos = runRem; }
		while (pos < 16) {
			uint32_t nk = k < nS ? (uint32_t)S.sKey[k] - rel : 16u;
			if (nk > 16) nk = 16;
			const uint32_t cnt = nk - pos;            // old symbols in front of the next record (< 16 here)
			if (cnt) { R |= (O & ((((u128)1) << (8 * cnt)) - 1)) << (8 * pos); O >>= 8 * cnt; pos = nk; }
			if (pos >= 16) break;
			uint32_t len = (uint32_t)S.sLS[k] >> 3; const uint32_t sy = S.sLS[k] & 7u;
			if (len > 16 - pos) len = 16 - pos;
			if (len == 16) { R = fill(sy); break; }
			R |= (fill(sy) & ((((u128)1) << (8 * len)) - 1)) << (8 * pos);
			pos += len; ++k;
		}
		ow[0] = (uint32_t)R; ow[1] = (uint32_t)(R >> 32); ow[2] = (uint32_t)(R >> 64); ow[3] = (uint32_t)(R >> 96);
	}
	// symbols behind the end of the array (last tile) are zero
	if (rel + 16 > tileLen) {
		const uint32_t nValid = rel >= tileLen ? 0u : tileLen - rel;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t kb = nValid > (uint32_t)j * 4 ? nValid - j * 4 : 0;
			ow[j] = kb >= 4 ? ow[j] : (kb ? ow[j] & ((1u << (kb * 8)) - 1u) : 0u);
		}
	}
	reinterpret_cast<uint4*>(A.newS + o0)[tid] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
	// ---- rank(a, P) for the records that start in this tile (the last warps take them: they had the least to do) ----
	{
		const int64_t *dirRow = A.oldDir + (a0 / FT_DIR) * 6;
		for (uint32_t k = 255 - tid; k < r1 - r0; k += 256) {
			const uint32_t r = r0 + k, dst = A.recDst[r];
			if (dst == NONE32) continue;
			const uint32_t a = S.sLS[nCarry + k] & 7u;
			const uint32_t x = (uint32_t)S.sKey[nCarry + k] - S.sPre[nCarry + k] + skip; // = P - a0
			const Raw6 rr = prefix_at(x);
			int64_t g = A.ctl->cpost[a] + dirRow[a] + raw_symbol(rr, a);
			if (A.recOff) // sharded: which of my buckets the record belongs to -> whole-index coordinates
				g += A.recOff[bucket_of(A.ctl->recBkt, (uint32_t)A.nb, r) * 7 + a];
			A.gLNext[dst] = g;
		}
	}
	// ---- symbol counts of the four FT_DIR sub-tiles = old symbols in them + record symbols -----------------
	__syncthreads();
	if (tid < FT_SUB * 6) {
		const uint32_t sb = (uint32_t)tid / 6u, f = (uint32_t)tid % 6u;
		const uint64_t tile = (uint64_t)blockIdx.x * FT_SUB + sb;
		if (tile * FT_DIR < A.nNew || tile == 0) {
			const Raw6 lo = prefix_at(S.subX[sb]), hi = prefix_at(S.subX[sb + 1]);
			A.newTileCnt[tile * 6 + f] = raw_symbol(hi, f) - raw_symbol(lo, f) + S.recCnt[sb][f];
		}
	}
}

struct FlatDirScan { // K=6 (int64): per-tile symbol counts -> counts in front of every tile
	const uint32_t *tileCnt; uint64_t nTile; int64_t *dir;
	__device__ void load(uint64_t i, int64_t (&v)[6]) const {
#pragma unroll
		for (int a = 0; a < 6; ++a) v[a] = tileCnt[i * 6 + a];
	}
	__device__ void store(uint64_t i, const int64_t (&own)[6], const int64_t (&pre)[6]) const {
#pragma unroll
		for (int a = 0; a < 6; ++a) dir[i * 6 + a] = pre[a];
		if (i + 1 == nTile) {
#pragma unroll
			for (int a = 0; a < 6; ++a) dir[(i + 1) * 6 + a] = pre[a] + own[a];
		}
	}
};

// per-tile symbol counts of a flat array (after blocks -> flat)
__global__ void __launch_bounds__(256) k_flat_count_tiles(const uint8_t *flat, uint64_t n, uint32_t *tileCnt)
{
	// one warp per FT_DIR tile: 32 lanes x 32 bytes
	const int lane = threadIdx.x & 31;
	const uint64_t tile = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
	if (tile * FT_DIR >= n && tile != 0) return;
	const uint64_t base = tile * FT_DIR + lane * 32;
	const uint32_t rem = base >= n ? 0u : (n - base < 32 ? (uint32_t)(n - base) : 32u);
	const uint4 *src = reinterpret_cast<const uint4*>(flat + base);
	const uint4 x = src[0], y = src[1];
	const Raw6 c0 = raw_count16(x.x, x.y, x.z, x.w, rem < 16 ? rem : 16), c1 = raw_count16(y.x, y.y, y.z, y.w, rem > 16 ? rem - 16 : 0);
	Raw6 s = { c0.s0 + c1.s0, c0.s1 + c1.s1, c0.s2 + c1.s2, c0.s01 + c1.s01, c0.s02 + c1.s02, c0.n + c1.n };
	s.s0 = warp_sum(s.s0); s.s1 = warp_sum(s.s1); s.s2 = warp_sum(s.s2); s.s01 = warp_sum(s.s01); s.s02 = warp_sum(s.s02); s.n = warp_sum(s.n);
	if (lane < 6) tileCnt[tile * 6 + lane] = raw_symbol(s, (uint32_t)lane);
}

// occ(a, x) on the flat array, all six symbols, one warp; result in every lane
__device__ __forceinline__ void flat_rank6(const uint8_t *flat, const int64_t *dir, int64_t x, int lane, int64_t (&out)[6])
{
	const uint64_t t = (uint64_t)x / FT_DIR;
	const uint32_t part = (uint32_t)((uint64_t)x - t * FT_DIR);
	const uint32_t rem = part > (uint32_t)lane * 32 ? (part - lane * 32 < 32 ? part - lane * 32 : 32u) : 0u;
	Raw6 s = { 0, 0, 0, 0, 0, 0 };
	if (rem) {
		const uint4 *src = reinterpret_cast<const uint4*>(flat + t * FT_DIR + lane * 32);
		const uint4 xx = src[0], y = src[1];
		const Raw6 c0 = raw_count16(xx.x, xx.y, xx.z, xx.w, rem < 16 ? rem : 16), c1 = raw_count16(y.x, y.y, y.z, y.w, rem > 16 ? rem - 16 : 0);
		s.s0 = c0.s0 + c1.s0; s.s1 = c0.s1 + c1.s1; s.s2 = c0.s2 + c1.s2; s.s01 = c0.s01 + c1.s01; s.s02 = c0.s02 + c1.s02; s.n = c0.n + c1.n;
	}
	s.s0 = warp_sum(s.s0); s.s1 = warp_sum(s.s1); s.s2 = warp_sum(s.s2); s.s01 = warp_sum(s.s01); s.s02 = warp_sum(s.s02); s.n = warp_sum(s.n);
#pragma unroll
	for (int a = 0; a < 6; ++a) out[a] = dir[t * 6 + a] + raw_symbol(s, (uint32_t)a);
}

// sizes6[g][a] = #a in [gL, gL+gSize) for every group with a non-empty interval (rope_rank2a, mrope.c:202).
// posOff: sharded engines pass whole-index positions; bucket b's local position = position - posOff[b*7+6].
__global__ void __launch_bounds__(128) k_flat_rank_groups(const uint8_t *flat, const int64_t *dir, uint32_t G, const int64_t *gL, const int64_t *gSize,
                                                          int64_t *sizes6, const Ctl *ctl, const int64_t *posOff, int nb)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t g0 = (blockIdx.x * 4 + wid) * 32;
	if (g0 >= G) return;
	const uint32_t g = g0 + lane;
	int64_t myL = g < G ? gL[g] : 0; const int64_t mySz = g < G ? gSize[g] : 0;
	if (posOff && g < G && mySz > 0) myL -= posOff[bucket_of(ctl->gBkt, (uint32_t)nb, g) * 7 + 6];
	uint32_t todo = __ballot_sync(FULLMASK, mySz > 0);
	while (todo) {
		const int src = __ffs(todo) - 1; todo &= todo - 1;
		const int64_t L = __shfl_sync(FULLMASK, myL, src), sz = __shfl_sync(FULLMASK, mySz, src);
		int64_t cl[6], cu[6];
		flat_rank6(flat, dir, L, lane, cl);
		flat_rank6(flat, dir, L + sz, lane, cu);
		if (lane < 6) {
			int64_t v = 0;
#pragma unroll
			for (int a = 0; a < 6; ++a) if (lane == a) v = cu[a] - cl[a];
			sizes6[(size_t)(g0 + src) * 6 + lane] = v;
		}
	}
}

// ---- leaf blocks -> flat --------------------------------------------------------------------------------
// one warp per logical block: every lane expands the runs that start in its 16 bytes
// (sharded engines: off[b*7+6] = symbols of the whole index in front of bucket b that other ranks hold)
__global__ void __launch_bounds__(128) k_blocks_to_flat(const uint8_t *pool, const uint32_t *order, const int64_t *cumLen, uint32_t nlog,
                                                        const int64_t *off, const uint32_t *bkt, int nb, uint8_t *flat, Ctl *ctl)
{
	__shared__ __align__(16) uint8_t sImg[4][RB2_IMG_BYTES];
	__shared__ uint32_t sCnt[4][32 * 7];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t i = blockIdx.x * 4 + wid;
	if (i >= nlog) return;
	LaneDec d; uint32_t basePos, baseCnt[6], blkLen, blkCnt[6], nbytes, err = 0; uint4 own;
	warp_decode_block(pool + (size_t)order[i] * RB2_BLK, lane, sImg[wid], sCnt[wid], d, basePos, baseCnt, blkLen, blkCnt, nbytes, err, own);
	const int64_t posBase = off ? off[bucket_of(bkt, (uint32_t)nb, i) * 7 + 6] : 0;
	uint8_t *dst = flat + (cumLen[i] - posBase) + basePos;
	uint32_t bp = lane * 16 + d.fb;
	for (uint32_t q = 0; q < d.nr; ++q) {
		uint32_t l, s, nb;
		parse_run(sImg[wid], bp, s, l, nb);
		bp += nb;
		for (uint32_t j = 0; j < l; ++j) dst[j] = (uint8_t)s;
		dst += l;
	}
	if (err && lane == 0) atomicOr(&ctl->err, err);
}

// ---- flat -> leaf blocks --------------------------------------------------------------------------------
// Buckets are encoded independently (a block never spans two buckets).  Chunk j of bucket b covers its
// symbols [64j, 64j+64); a chunk is encoded on its own (a run never crosses a chunk boundary -- equal
// neighbours are legal and merged by every consumer, rld0.c:153-161), so its size is known locally.
struct EncTab { int nb; uint64_t symStart[NBMAX + 1]; uint64_t chunkStart[NBMAX + 1]; uint64_t byteStart[NBMAX + 1]; uint32_t blkStart[NBMAX + 1]; };

__device__ __forceinline__ int enc_bucket_of_chunk(const EncTab &T, uint64_t c)
{
	int b = 0;
	for (int x = 1; x < T.nb; ++x) b += c >= T.chunkStart[x];
	return b;
}

// bytes of the encoded chunk
__global__ void __launch_bounds__(256) k_flat_chunk_bytes(const uint8_t *flat, EncTab T, uint64_t nChunk, uint8_t *chunkBytes)
{
	const uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (c >= nChunk) return;
	const int b = enc_bucket_of_chunk(T, c);
	const uint64_t s0 = T.symStart[b] + (c - T.chunkStart[b]) * FE_CHUNK;
	const uint64_t s1 = s0 + FE_CHUNK < T.symStart[b + 1] ? s0 + FE_CHUNK : T.symStart[b + 1];
	uint32_t bytes = 0, prev = 8, len = 0;
	for (uint64_t i = s0; i < s1; ++i) {
		const uint32_t s = flat[i];
		if (s != prev) { if (len) bytes += len < 16 ? 1 : 2; prev = s; len = 0; }
		++len;
	}
	if (len) bytes += len < 16 ? 1 : 2;
	chunkBytes[c] = (uint8_t)bytes;
}

struct ChunkScan { // K=1 (uint64): exclusive prefix of the chunk sizes
	const uint8_t *chunkBytes; uint64_t n; uint64_t *pre;
	__device__ void load(uint64_t i, uint64_t (&v)[1]) const { v[0] = chunkBytes[i]; }
	__device__ void store(uint64_t i, const uint64_t (&own)[1], const uint64_t (&p)[1]) const {
		pre[i] = p[0];
		if (i + 1 == n) pre[n] = p[0] + own[0];
	}
};

// one warp per output block: find its chunks, encode them into a shared-memory image, store it
__global__ void __launch_bounds__(128) k_flat_encode(const uint8_t *flat, EncTab T, const uint64_t *chunkPre, uint32_t nBlocks, uint8_t *pool, uint32_t *blkCnt)
{
	__shared__ __align__(16) uint8_t sImg[4][RB2_BLK + 64];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t k = blockIdx.x * 4 + wid;
	if (k >= nBlocks) return;
	int b = 0;
	for (int x = 1; x < T.nb; ++x) b += k >= T.blkStart[x];
	const uint32_t kk = k - T.blkStart[b];
	const uint64_t cLo = T.chunkStart[b], cHi = T.chunkStart[b + 1], base = T.byteStart[b];
	// first chunk whose start byte (relative to the bucket) is >= kk*FE_T, resp. >= (kk+1)*FE_T
	auto lower = [&](uint64_t key) { uint64_t lo = cLo, hi = cHi; while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (chunkPre[mid] - base >= key) hi = mid; else lo = mid + 1; } return lo; };
	const uint64_t c0 = lower((uint64_t)kk * FE_T), c1 = lower((uint64_t)(kk + 1) * FE_T);
	const uint64_t b0 = c0 < cHi ? chunkPre[c0] : chunkPre[cHi];
	const uint32_t nbytes = (uint32_t)((c1 < cHi ? chunkPre[c1] : chunkPre[cHi]) - b0);
	uint8_t *img = sImg[wid];
	for (int j = lane; j < (RB2_BLK + 64) / 4; j += 32) reinterpret_cast<uint32_t*>(img)[j] = 0;
	__syncwarp();
	uint32_t cnt[6] = { 0, 0, 0, 0, 0, 0 };
	for (uint64_t c = c0 + lane; c < c1; c += 32) {
		const uint64_t s0 = T.symStart[b] + (c - cLo) * FE_CHUNK;
		const uint64_t s1 = s0 + FE_CHUNK < T.symStart[b + 1] ? s0 + FE_CHUNK : T.symStart[b + 1];
		uint8_t *o = img + 2 + (uint32_t)(chunkPre[c] - b0);
		uint32_t prev = 8, len = 0;
		for (uint64_t i = s0; i < s1; ++i) {
			const uint32_t s = flat[i];
			if (s != prev) { if (len) o += enc_run(o, prev, len); prev = s; len = 0; }
			++len;
#pragma unroll
			for (int a = 0; a < 6; ++a) cnt[a] += s == (uint32_t)a;
		}
		if (len) o += enc_run(o, prev, len);
	}
	__syncwarp();
	if (lane == 0) { img[0] = (uint8_t)(nbytes & 0xff); img[1] = (uint8_t)(nbytes >> 8); }
	__syncwarp();
	reinterpret_cast<uint4*>(pool + (size_t)k * RB2_BLK)[lane] = reinterpret_cast<const uint4*>(img)[lane];
#pragma unroll
	for (int a = 0; a < 6; ++a) cnt[a] = warp_sum(cnt[a]);
	if (lane < 6) {
		uint32_t v = 0;
#pragma unroll
		for (int a = 0; a < 6; ++a) if (lane == a) v = cnt[a];
		blkCnt[(size_t)k * 6 + lane] = v;
	}
}
