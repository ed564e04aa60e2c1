// rb2_flat.cuh -- the DENSE regime of the BCR column step: stream-rewrite of a fixed-width BWT.
//
// When a batch inserts into (nearly) every leaf block in every column -- short reads, the headline
// workload: 100 M records per column into <= 20 M blocks -- updating run-length coded blocks record
// by record is instruction bound (profiles/README.md).  For such a batch the engine keeps the BWT,
// for the duration of the batch only, as ONE flat array of nt6 codes (4 bits per symbol, two per
// byte, low nibble first; all six buckets concatenated) plus a directory of per-symbol counts in
// front of every FT_DIR-th symbol, and one column becomes a single streaming pass (k_flat_merge):
//
//   new[P_r + pre_r .. + c_r) = symbol of record r     (pre_r = members in front of record r, i.e. the
//   old symbol i moves to i + #record symbols with P<=i  symbols this column inserts in front of it)
//
// which is exactly the stable merge rope_insert_run performs one run at a time (rope.c:114-148;
// new symbols go in FRONT of the old symbol at the same position, mrope.c:206-218), and
// rank(a, P_r) -- the return value of rope_insert_run / rle_insert_cached (rle.c:10-89) -- is
// directory[P_r / FT_DIR][a] + a count over < FT_DIR + FT_OUT symbols held in shared memory.
// Output-stationary: CTA t produces new[t*FT_OUT, (t+1)*FT_OUT) with aligned 128-bit stores; the old
// symbols it needs are one contiguous range.  At the end of the batch the array is re-encoded into
// leaf blocks of the reference's format (k_flat_encode), so everything outside the batch (iterator,
// dump, rank queries, sparse batches) sees the usual block pool.
#pragma once
#include "rb2_codec.cuh"

#define FT_OUT   8192  // output symbols per CTA of k_flat_merge: 256 threads x 32 symbols (16 bytes)
#define FT_DIR   2048  // directory granularity (48 bytes of counts per 1024 bytes of symbols)
#define FT_SUB   (FT_OUT / FT_DIR)
#define FT_OLDMAX (FT_OUT + FT_DIR) // old symbols one CTA can need: from the directory tile of its first one
#define FT_PAD   (FT_OLDMAX + 128)  // readable slack (symbols) behind a flat array
#define FT_CH    32    // symbols per 16-byte chunk
#define FE_CHUNK 64    // flat -> blocks: symbols encoded by one thread (<= 64 bytes of runs)
#define FE_T     (RB2_FILL - FE_CHUNK + 1) // block k of a bucket takes the chunks that start in bytes [k*FE_T, (k+1)*FE_T)

__host__ __device__ __forceinline__ uint64_t flat_bytes(uint64_t symbols) { return (symbols + 1) >> 1; }
__device__ __forceinline__ uint32_t flat_get(const uint8_t *flat, uint64_t i) { return (flat[i >> 1] >> ((i & 1) * 4)) & 15u; }

// ---- per-symbol counting without per-symbol compares ---------------------------------------------
// nt6 codes: $=000 A=001 C=010 G=011 T=100 N=101 (bit 3 of a nibble is always 0).  For the eight
// codes of a word, popcounts of
//   m0 = bit0, m1 = bit1, m2 = bit2, m0&m1 (G), m0&m2 (N)
// give N = s02, G = s01, A = s0 - s01 - s02, C = s1 - s01, T = s2 - s02, $ = n - (A+C+G+T+N).
// "Raw" counts (s0, s1, s2, s01, s02, n) are linear, so prefix sums are taken on them and converted
// only where a symbol count is needed.
struct Raw6 { uint32_t s0, s1, s2, s01, s02, n; };

__device__ __forceinline__ void raw_add_word(uint32_t x, uint32_t (&acc)[5])
{
	const uint32_t m0 = x & 0x11111111u, m1 = (x >> 1) & 0x11111111u, m2 = (x >> 2) & 0x11111111u;
	acc[0] += __popc(m0); acc[1] += __popc(m1); acc[2] += __popc(m2); acc[3] += __popc(m0 & m1); acc[4] += __popc(m0 & m2);
}

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

// keep the first ns (0..32) symbols of a 16-byte chunk, zero the rest
__device__ __forceinline__ void chunk_mask(uint32_t (&w)[4], uint32_t ns)
{
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const uint32_t k = ns > (uint32_t)j * 8 ? ns - j * 8 : 0;
		w[j] = k >= 8 ? w[j] : (k ? w[j] & ((1u << (k * 4)) - 1u) : 0u);
	}
}

// raw counts of the first ns (0..32) symbols of a 16-byte chunk
__device__ __forceinline__ Raw6 raw_count_chunk(const uint4 &v, uint32_t ns)
{
	uint32_t w[4] = { v.x, v.y, v.z, v.w };
	if (ns < FT_CH) chunk_mask(w, ns);
	uint32_t acc[5] = { 0, 0, 0, 0, 0 };
#pragma unroll
	for (int j = 0; j < 4; ++j) raw_add_word(w[j], acc);
	Raw6 r = { acc[0], acc[1], acc[2], acc[3], acc[4], ns };
	return r;
}
__device__ __forceinline__ void raw_addto(Raw6 &a, const Raw6 &b) { a.s0 += b.s0; a.s1 += b.s1; a.s2 += b.s2; a.s01 += b.s01; a.s02 += b.s02; a.n += b.n; }

// The records of a column.  In the all-singleton regime without interval sizes (the bulk of a short-read
// batch) a record is fully described by the state arrays themselves -- position = the group's interval
// start, symbol = the member's next symbol, count 1, members in front = its own index -- so the column
// kernel writes none of recP / recSC / recPre and the pointers below are the state arrays / null.
struct RecView {
	const int64_t *P; const uint32_t *pre, *sc; const uint8_t *asym;
	__device__ __forceinline__ uint32_t Pre(uint32_t r) const { return pre ? pre[r] : r; }
	__device__ __forceinline__ uint32_t SC(uint32_t r) const { return sc ? sc[r] : (8u | asym[r]); }
	__device__ __forceinline__ uint64_t Key(uint32_t r) const { return (uint64_t)P[r] + Pre(r); }
};

// ---- tile -> record ranges ---------------------------------------------------------------------------
// tileR0[t] = first record whose output run starts at or behind t*FT_OUT (key_r = P_r + pre_r); one
// thread per record (plus a virtual one behind the last) fills the tiles between its predecessor and itself.
__global__ void __launch_bounds__(256) k_flat_splits(const RecView V, uint32_t R, uint64_t nTiles, uint32_t *tileR0)
{
	const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (r > R) return;
	const int64_t tPrev = r == 0 ? -1 : (int64_t)(V.Key((uint32_t)r - 1) / FT_OUT);
	const int64_t t = r == R ? (int64_t)nTiles : (int64_t)(V.Key((uint32_t)r) / FT_OUT);
	for (int64_t tt = tPrev + 1; tt <= t; ++tt) tileR0[tt] = (uint32_t)r;
}

// Per-tile geometry, one thread per tile boundary t (output position min(t*FT_OUT, nNew)): the first old
// symbol that lands at or behind it, the first record that starts there, and the record run that reaches
// across it from the left.  Computed ahead of k_flat_merge so that its CTAs start with two independent loads.
struct alignas(16) TileDesc { uint64_t i0; uint32_t r0, carry; }; // carry = (symbols of the crossing run behind the boundary, capped at FT_OUT) << 3 | symbol

__global__ void __launch_bounds__(256) k_flat_geo(const RecView V, const uint32_t *tileR0, uint64_t nTiles, uint64_t nNew, TileDesc *desc)
{
	const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (t > nTiles) return;
	const uint64_t o0 = t * FT_OUT < nNew ? t * FT_OUT : nNew;
	const uint32_t r0 = tileR0[t];
	uint64_t before = 0; uint32_t carry = 0;
	if (r0 > 0) {
		const uint64_t pre = V.Pre(r0 - 1), key = (uint64_t)V.P[r0 - 1] + pre; const uint32_t sc = V.SC(r0 - 1);
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
	RecView V; const uint32_t *recDst; uint32_t R;
	const TileDesc *desc;
	uint32_t *ovf;           // [0] tiles left to k_flat_merge_dense, [1] its work counter, [2..] the tiles
	int64_t *gLNext; const Ctl *ctl;
	// sharded engines: records carry whole-index positions; bucket b of this rank sits recOff[b*7+6]
	// symbols (recOff[b*7+a] symbols a) further right in the whole index than in the local array
	const int64_t *recOff; int nb;
};

// six raw counts as three words of two 16-bit fields (sums stay below 2^16 inside one tile)
__device__ __forceinline__ void raw_pack16(const uint32_t (&acc)[5], uint32_t n, uint32_t (&p)[3])
{
	p[0] = acc[0] | acc[1] << 16; p[1] = acc[2] | acc[3] << 16; p[2] = acc[4] | n << 16;
}
__device__ __forceinline__ Raw6 raw_unpack16(uint32_t p0, uint32_t p1, uint32_t p2)
{
	Raw6 r = { p0 & 0xffffu, p0 >> 16, p1 & 0xffffu, p1 >> 16, p2 & 0xffffu, p2 >> 16 };
	return r;
}

#define FT_NCH (FT_OLDMAX / FT_CH + 2)   // 16-byte chunks of old symbols one tile can hold
#define FT_NOC (FT_OUT / FT_CH)          // 32-symbol output chunks per tile = threads per CTA
#define FT_CAP_SMALL 2047                // records per tile the main kernel stages (more: overflow kernel, same code)
template <int CAP> struct FlatSmemT {
	uint4    old4[FT_NCH];               // the old symbols this tile needs, from a directory tile boundary
	uint32_t chunkPre[FT_NCH][3];        // raw counts in front of every 16-byte chunk of old4 (16-bit fields)
	uint16_t sKey[CAP + 1];              // staged records: run start inside the tile
	uint16_t sLS[CAP + 1];               // (run length inside the tile - 1) << 3 | symbol
	uint32_t cntC[FT_NOC + 1];           // staged records that start in each 32-symbol output chunk
	uint16_t k0C[FT_NOC + 2];            // ... and their exclusive prefix = the first record at or behind each chunk
	uint32_t anyLong;                    // some staged record is longer than one symbol
	uint32_t warpTot[8][3];
	uint32_t recCnt[FT_SUB][6];          // symbols the records put into each FT_DIR sub-tile
	uint32_t subX[FT_SUB + 1];           // old symbols (local index) in front of each sub-tile
	uint32_t tile;                       // overflow kernel: the tile this CTA works on
};

// insert one symbol at nibble position pp (0..31) of a 32-symbol vector; the last symbol falls out
__device__ __forceinline__ void insert_symbol(uint32_t (&ow)[4], uint32_t pp, uint32_t sy)
{
	const uint32_t pw = pp >> 3, pb = (pp & 7) * 4;
	const uint32_t w1[4] = { ow[0] << 4, __funnelshift_l(ow[0], ow[1], 4), __funnelshift_l(ow[1], ow[2], 4), __funnelshift_l(ow[2], ow[3], 4) };
	const uint32_t lowMask = (1u << pb) - 1u;     // symbols in front of pp inside its word
#pragma unroll
	for (int w = 0; w < 4; ++w) {
		const uint32_t mid = (ow[w] & lowMask) | (sy << pb) | (w1[w] & ~((lowMask << 4) | 0xfu));
		ow[w] = (uint32_t)w < pw ? ow[w] : ((uint32_t)w > pw ? w1[w] : mid);
	}
}

// One output tile.  Two barriers: (A) old symbols + their counts | records -> shared memory, (B) every
// thread assembles its 32 output symbols, (C) ranks of the tile's records | symbol counts of its sub-tiles.
template <int CAP>
__device__ __forceinline__ void flat_merge_tile(const FlatArgs &A, FlatSmemT<CAP> &S, const uint32_t tile, const TileDesc d0, const TileDesc d1)
{
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint64_t o0 = (uint64_t)tile * FT_OUT;
	const uint32_t tileLen = A.nNew - o0 < FT_OUT ? (uint32_t)(A.nNew - o0) : FT_OUT;
	const uint32_t r0 = d0.r0, r1 = d1.r0;
	const uint32_t carrySym = d0.carry & 7u, carryLen = (d0.carry >> 3) < tileLen ? (d0.carry >> 3) : tileLen;
	const uint64_t i0 = d0.i0;                     // old symbols [i0, i1) land in this tile
	const uint64_t a0 = i0 & ~(uint64_t)(FT_DIR - 1);
	const uint32_t loadLen = (uint32_t)(d1.i0 - a0), skip = (uint32_t)(i0 - a0);
	const uint32_t recIn = tileLen - (loadLen - skip); // record symbols inside the tile
	const uint64_t before = o0 - i0;               // record symbols in front of the tile
	const uint32_t nLoad = loadLen / FT_CH + 2;    // chunks that are read (<= FT_NCH)
	const uint32_t nCarry = carryLen ? 1u : 0u, nS = nCarry + (r1 - r0);
	constexpr int NCNT = 160;                      // threads (5 warps) that load + count; the other 3 warps stage records

	// ---- phase A: old symbols -> shared memory + raw counts per chunk pair | records -> shared memory -------
	uint32_t p[3] = { 0, 0, 0 }, q[3] = { 0, 0, 0 }, inc[3] = { 0, 0, 0 };
	if (tid < NCNT) {
		// thread j owns chunks 2j, 2j+1 (symbols behind loadLen are whatever follows in the array: the
		// prefixes that include them are never used)
		const uint4 *src = reinterpret_cast<const uint4*>(A.oldS + (a0 >> 1));
		if ((uint32_t)tid * 2 < nLoad) {
			const uint4 x = src[tid * 2], y = src[tid * 2 + 1];
			S.old4[tid * 2] = x; S.old4[tid * 2 + 1] = y;
			uint32_t acc[5] = { 0, 0, 0, 0, 0 };
			raw_add_word(x.x, acc); raw_add_word(x.y, acc); raw_add_word(x.z, acc); raw_add_word(x.w, acc);
			raw_pack16(acc, FT_CH, q);
			raw_add_word(y.x, acc); raw_add_word(y.y, acc); raw_add_word(y.z, acc); raw_add_word(y.w, acc);
			raw_pack16(acc, 2 * FT_CH, p);
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
		// the three staging warps also index the records by output chunk (they would otherwise wait for the
		// counting warps): histogram of chunk ids, then one warp scans it.  Barrier 1 is private to them.
		const int st = tid - NCNT;
		for (int c = st; c <= FT_NOC; c += 256 - NCNT) S.cntC[c] = 0;
		if (st == 0) S.anyLong = 0;
		if (st < FT_SUB * 6) (&S.recCnt[0][0])[st] = 0;
		RB2_NAMED_BAR(1, 96);
		if (st == 0 && nCarry) {
			S.sKey[0] = 0; S.sLS[0] = (uint16_t)(((carryLen - 1) << 3) | carrySym);
			atomicAdd(&S.cntC[0], 1u);
			if (carryLen > 1) S.anyLong = 1;
		}
		for (uint32_t k = st; k < r1 - r0; k += 256 - NCNT) {
			const uint32_t r = r0 + k;
			const uint32_t pre = A.V.Pre(r), sc = A.V.SC(r);
			const uint32_t key = (uint32_t)((uint64_t)A.V.P[r] + pre - o0);
			uint32_t len = sc >> 3;
			if (len > FT_OUT - key) len = FT_OUT - key;
			S.sKey[nCarry + k] = (uint16_t)key; S.sLS[nCarry + k] = (uint16_t)(((len - 1) << 3) | (sc & 7u));
			atomicAdd(&S.cntC[key / FT_CH], 1u);
			if (len > 1) S.anyLong = 1;
		}
		RB2_NAMED_BAR(1, 96);
		if (wid == NCNT / 32) { // eight chunks per lane
			uint32_t v[8], sum = 0;
#pragma unroll
			for (int i = 0; i < 8; ++i) { v[i] = S.cntC[lane * 8 + i]; sum += v[i]; }
			uint32_t ex = warp_incl_scan(sum, lane) - sum;
#pragma unroll
			for (int i = 0; i < 8; ++i) { S.k0C[lane * 8 + i] = (uint16_t)ex; ex += v[i]; }
			if (lane == 31) S.k0C[FT_NOC] = (uint16_t)ex;
		}
	}
	__syncthreads();
	// ---- phase B: prefix in front of every chunk (needed behind the next barrier) ------------------------
	if (tid < NCNT && (uint32_t)tid * 2 < nLoad) {
		uint32_t base[3] = { 0, 0, 0 };
		for (int w = 0; w < wid; ++w) { base[0] += S.warpTot[w][0]; base[1] += S.warpTot[w][1]; base[2] += S.warpTot[w][2]; }
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const uint32_t ex = base[k] + inc[k] - p[k];
			S.chunkPre[tid * 2][k] = ex; S.chunkPre[tid * 2 + 1][k] = ex + q[k];
		}
	}
	// ---- phase B: assemble 32 output symbols per thread ------------------------------------------------
	// record symbols of this tile in front of staged entry k (the carried run counts from the tile start)
	auto pre_rel = [&](uint32_t k) -> uint32_t { return k < nCarry ? 0u : (k < nS ? (uint32_t)(A.V.Pre(r0 + k - nCarry) - before) : recIn); };
	auto run_len = [&](uint32_t k) -> uint32_t { return ((uint32_t)S.sLS[k] >> 3) + 1; };
	const uint32_t rel = tid * FT_CH;
	uint32_t k0;                        // first entry with sKey >= rel
	k0 = S.k0C[tid];                    // (a bisection over sKey costs 70 instructions per warp; a proportional guess + walk twice that)
	const uint32_t kEnd = S.k0C[tid + 1]; // one past the last entry that starts in this chunk
	uint32_t runRem = 0, runSym = 0, oldIdx;
	if (k0 > 0 && (uint32_t)S.sKey[k0 - 1] + run_len(k0 - 1) > rel) { // inside the run of entry k0-1
		runRem = (uint32_t)S.sKey[k0 - 1] + run_len(k0 - 1) - rel; runSym = S.sLS[k0 - 1] & 7u;
		oldIdx = (uint32_t)S.sKey[k0 - 1] - pre_rel(k0 - 1) + skip;
	} else oldIdx = rel - pre_rel(k0) + skip;
	if ((tid & (FT_DIR / FT_CH - 1)) == 0) S.subX[tid / (FT_DIR / FT_CH)] = oldIdx < loadLen ? oldIdx : loadLen;
	if (tid == 0) S.subX[FT_SUB] = loadLen;
	const uint32_t sbMine = rel / FT_DIR;          // the sub-tile this chunk lies in
	uint32_t ow[4];
	{
		// 32 old symbols from oldIdx on (unaligned)
		const uint32_t *wp = reinterpret_cast<const uint32_t*>(S.old4) + (oldIdx >> 3);
		const uint32_t sh = (oldIdx & 7) * 4;
		const uint32_t x0 = wp[0], x1 = wp[1], x2 = wp[2], x3 = wp[3], x4 = wp[4];
		ow[0] = __funnelshift_r(x0, x1, sh); ow[1] = __funnelshift_r(x1, x2, sh); ow[2] = __funnelshift_r(x2, x3, sh); ow[3] = __funnelshift_r(x3, x4, sh);
	}
	if (runRem >= FT_CH) {                          // inside one long run
		ow[0] = ow[1] = ow[2] = ow[3] = runSym * 0x11111111u;
		atomicAdd(&S.recCnt[sbMine][runSym], FT_CH);
	} else if (runRem || kEnd > k0) {               // records start (or a run ends) inside these 32 symbols
		// are all of them single symbols?  (the rule late in a batch: then no record of the tile is longer)
		uint32_t k = kEnd; bool single = runRem == 0;
		if (single && S.anyLong) { k = k0; while (single && k < kEnd) { single = ((uint32_t)S.sLS[k] >> 3) == 0; ++k; } }
		if (single) {
			for (uint32_t j = k0; j < k; ++j) {
				const uint32_t sy = S.sLS[j] & 7u;
				insert_symbol(ow, (uint32_t)S.sKey[j] - rel, sy);
				atomicAdd(&S.recCnt[sbMine][sy], 1u);
			}
		} else {
			typedef unsigned __int128 u128;
			auto fill = [](uint32_t sy) -> u128 { const uint64_t f = 0x1111111111111111ull * sy; return ((u128)f << 64) | f; };
			u128 O = ((u128)(((uint64_t)ow[3] << 32) | ow[2]) << 64) | (((uint64_t)ow[1] << 32) | ow[0]); // nibble 0 = next old symbol
			u128 R = 0;
			uint32_t pos = 0;
			k = k0;
			if (runRem) { R = fill(runSym) & ((((u128)1) << (4 * runRem)) - 1); pos = runRem; atomicAdd(&S.recCnt[sbMine][runSym], runRem); }
			while (pos < FT_CH) {
				uint32_t nk = k < nS ? (uint32_t)S.sKey[k] - rel : (uint32_t)FT_CH;
				if (nk > FT_CH) nk = FT_CH;
				const uint32_t cnt = nk - pos;            // old symbols in front of the next record (< 32 here)
				if (cnt) { R |= (O & ((((u128)1) << (4 * cnt)) - 1)) << (4 * pos); O >>= 4 * cnt; pos = nk; }
				if (pos >= FT_CH) break;
				uint32_t len = run_len(k); const uint32_t sy = S.sLS[k] & 7u;
				if (len > FT_CH - pos) len = FT_CH - pos;
				atomicAdd(&S.recCnt[sbMine][sy], len);
				if (len == FT_CH) { R = fill(sy); break; }
				R |= (fill(sy) & ((((u128)1) << (4 * len)) - 1)) << (4 * pos);
				pos += len; ++k;
			}
			ow[0] = (uint32_t)R; ow[1] = (uint32_t)(R >> 32); ow[2] = (uint32_t)(R >> 64); ow[3] = (uint32_t)(R >> 96);
		}
	}
	// symbols behind the end of the array (last tile) are zero
	if (rel + FT_CH > tileLen) chunk_mask(ow, rel >= tileLen ? 0u : tileLen - rel);
	reinterpret_cast<uint4*>(A.newS + (o0 >> 1))[tid] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
	__syncthreads();
	// raw counts in front of local old index x
	auto prefix_at = [&](uint32_t x) -> Raw6 {
		const uint32_t c = x / FT_CH;
		Raw6 rr = raw_unpack16(S.chunkPre[c][0], S.chunkPre[c][1], S.chunkPre[c][2]);
		if (x & (FT_CH - 1)) raw_addto(rr, raw_count_chunk(S.old4[c], x & (FT_CH - 1)));
		return rr;
	};
	// ---- phase C: symbol counts of the four FT_DIR sub-tiles = old symbols in them + record symbols (warp 0) ----
	if (tid < FT_SUB * 6) {
		const uint32_t sb = (uint32_t)tid / 6u, f = (uint32_t)tid % 6u;
		const uint64_t dt = (uint64_t)tile * FT_SUB + sb;
		if (dt * FT_DIR < A.nNew || dt == 0) {
			const Raw6 lo = prefix_at(S.subX[sb]), hi = prefix_at(S.subX[sb + 1]);
			A.newTileCnt[dt * 6 + f] = raw_symbol(hi, f) - raw_symbol(lo, f) + S.recCnt[sb][f];
		}
	}
	// ---- phase C: rank(a, P) for the records that start in this tile (taken from the last warp downwards) ----
	{
		const int64_t *dirRow = A.oldDir + (a0 / FT_DIR) * 6;
		for (uint32_t k = 255 - tid; k < r1 - r0; k += 256) {
			const uint32_t r = r0 + k, dst = A.recDst[r];
			if (dst == NONE32) continue;
			const uint32_t a = S.sLS[nCarry + k] & 7u;
			const Raw6 rr = prefix_at((uint32_t)((uint64_t)A.V.P[r] - a0));
			int64_t g = A.ctl->cpost[a] + dirRow[a] + raw_symbol(rr, a);
			if (A.recOff) // sharded: which of my buckets the record belongs to -> whole-index coordinates
				g += A.recOff[bucket_of(A.ctl->recBkt, (uint32_t)A.nb, r) * 7 + a];
			A.gLNext[dst] = g;
		}
	}
}

#ifndef FT_MINCTA
#define FT_MINCTA 8
#endif
// main kernel: one CTA per tile; tiles with more records than it stages go to the overflow list
__global__ void __launch_bounds__(256, FT_MINCTA) k_flat_merge(FlatArgs A)
{
	RB2_DYN_SMEM(smraw);
	FlatSmemT<FT_CAP_SMALL> &S = *reinterpret_cast<FlatSmemT<FT_CAP_SMALL>*>(smraw);
	const TileDesc d0 = A.desc[blockIdx.x], d1 = A.desc[blockIdx.x + 1];
	if (d1.r0 - d0.r0 + 1 > FT_CAP_SMALL) { // (+1: a run carried in from the left)
		if (threadIdx.x == 0) A.ovf[2 + atomicAdd(&A.ovf[0], 1u)] = blockIdx.x;
		return;
	}
	flat_merge_tile<FT_CAP_SMALL>(A, S, blockIdx.x, d0, d1);
}

// overflow kernel (persistent): tiles where records are dense -- small indexes, first columns of an input-order batch
__global__ void __launch_bounds__(256) k_flat_merge_dense(FlatArgs A)
{
	RB2_DYN_SMEM(smraw);
	FlatSmemT<FT_OUT> &S = *reinterpret_cast<FlatSmemT<FT_OUT>*>(smraw);
	const uint32_t n = A.ovf[0];
	for (;;) {
		if (threadIdx.x == 0) S.tile = atomicAdd(&A.ovf[1], 1u);
		__syncthreads();
		const uint32_t q = S.tile;
		if (q >= n) break;
		const uint32_t tile = A.ovf[2 + q];
		flat_merge_tile<FT_OUT>(A, S, tile, A.desc[tile], A.desc[tile + 1]);
		__syncthreads();
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
	// one warp per FT_DIR tile: 32 lanes x 2 chunks
	const int lane = threadIdx.x & 31;
	const uint64_t tile = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
	if (tile * FT_DIR >= n && tile != 0) return;
	const uint64_t base = tile * FT_DIR + lane * 2 * FT_CH;
	const uint32_t rem = base >= n ? 0u : (n - base < 2 * FT_CH ? (uint32_t)(n - base) : 2u * FT_CH);
	const uint4 *src = reinterpret_cast<const uint4*>(flat + (base >> 1));
	Raw6 s = raw_count_chunk(src[0], rem < FT_CH ? rem : FT_CH);
	raw_addto(s, raw_count_chunk(src[1], rem > FT_CH ? rem - FT_CH : 0));
	s.s0 = warp_sum(s.s0); s.s1 = warp_sum(s.s1); s.s2 = warp_sum(s.s2); s.s01 = warp_sum(s.s01); s.s02 = warp_sum(s.s02); s.n = warp_sum(s.n);
	if (lane < 6) tileCnt[tile * 6 + lane] = raw_symbol(s, (uint32_t)lane);
}

// occ(a, x) on the flat array, all six symbols, one warp; result in every lane
__device__ __forceinline__ void flat_rank6(const uint8_t *flat, const int64_t *dir, int64_t x, int lane, int64_t (&out)[6])
{
	const uint64_t t = (uint64_t)x / FT_DIR;
	const uint32_t part = (uint32_t)((uint64_t)x - t * FT_DIR);
	const uint32_t rem = part > (uint32_t)lane * 2 * FT_CH ? (part - lane * 2 * FT_CH < 2 * FT_CH ? part - lane * 2 * FT_CH : 2u * FT_CH) : 0u;
	Raw6 s = { 0, 0, 0, 0, 0, 0 };
	if (rem) {
		const uint4 *src = reinterpret_cast<const uint4*>(flat + ((t * FT_DIR + lane * 2 * FT_CH) >> 1));
		s = raw_count_chunk(src[0], rem < FT_CH ? rem : FT_CH);
		raw_addto(s, raw_count_chunk(src[1], rem > FT_CH ? rem - FT_CH : 0));
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
// two steps: expand the runs to one byte per symbol (one warp per logical block: every lane expands the
// runs that start in its 16 bytes), then pack two symbols per byte
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

__global__ void __launch_bounds__(256) k_pack_nibbles(const uint8_t *bytes, uint64_t n, uint8_t *flat)
{
	const uint64_t i = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 8; // 8 symbols -> one word
	if (i >= n) return;
	uint32_t w = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) if (i + j < n) w |= (uint32_t)(bytes[i + j] & 15u) << (4 * j);
	reinterpret_cast<uint32_t*>(flat)[i >> 3] = w;
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

// the FE_CHUNK symbols that start at symbol s0, as eight words (low nibble first)
__device__ __forceinline__ void load_chunk64(const uint8_t *flat, uint64_t s0, uint32_t (&w)[8])
{
	const uint32_t *wp = reinterpret_cast<const uint32_t*>(flat) + (s0 >> 3);
	const uint32_t sh = (uint32_t)(s0 & 7) * 4;
	uint32_t x[9];
#pragma unroll
	for (int j = 0; j < 9; ++j) x[j] = wp[j];
#pragma unroll
	for (int j = 0; j < 8; ++j) w[j] = __funnelshift_r(x[j], x[j + 1], sh);
}

// bytes of the encoded chunk
__global__ void __launch_bounds__(256) k_flat_chunk_bytes(const uint8_t *flat, EncTab T, uint64_t nChunk, uint8_t *chunkBytes)
{
	const uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (c >= nChunk) return;
	const int b = enc_bucket_of_chunk(T, c);
	const uint64_t s0 = T.symStart[b] + (c - T.chunkStart[b]) * FE_CHUNK;
	const uint32_t n = (uint32_t)(s0 + FE_CHUNK < T.symStart[b + 1] ? FE_CHUNK : T.symStart[b + 1] - s0);
	uint32_t w[8];
	load_chunk64(flat, s0, w);
	uint32_t bytes = 0, prev = 8, len = 0;
#pragma unroll
	for (int i = 0; i < FE_CHUNK; ++i) if ((uint32_t)i < n) {
		const uint32_t sy = (w[i >> 3] >> ((i & 7) * 4)) & 15u;
		if (sy != prev) { bytes += len == 0 ? 0 : (len < 16 ? 1 : 2); prev = sy; len = 0; }
		++len;
	}
	bytes += len == 0 ? 0 : (len < 16 ? 1 : 2);
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

// eight lanes per output block (four blocks per warp; a block holds ~9 chunks of random reads): find its
// chunks, encode them into a shared-memory image, store it
__global__ void __launch_bounds__(128) k_flat_encode(const uint8_t *flat, EncTab T, const uint64_t *chunkPre, uint32_t nBlocks, uint8_t *pool, uint32_t *blkCnt)
{
	__shared__ __align__(16) uint8_t sImg[16][RB2_BLK + 64];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, grp = lane >> 3, gl = lane & 7;
	const uint32_t gmask = 0xffu << (grp * 8);
	const uint32_t k = (blockIdx.x * 4 + wid) * 4 + grp;
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
	uint8_t *img = sImg[wid * 4 + grp];
	for (int j = gl; j < (RB2_BLK + 64) / 4; j += 8) reinterpret_cast<uint32_t*>(img)[j] = 0;
	__syncwarp(gmask);
	uint32_t acc[5] = { 0, 0, 0, 0, 0 }, nsym = 0;
	for (uint64_t c = c0 + gl; c < c1; c += 8) {
		const uint64_t s0 = T.symStart[b] + (c - cLo) * FE_CHUNK;
		const uint32_t n = (uint32_t)(s0 + FE_CHUNK < T.symStart[b + 1] ? FE_CHUNK : T.symStart[b + 1] - s0);
		uint32_t w[8];
		load_chunk64(flat, s0, w);
		uint8_t *o = img + 2 + (uint32_t)(chunkPre[c] - b0);
		uint32_t prev = 8, len = 0;
#pragma unroll
		for (int i = 0; i < FE_CHUNK; ++i) if ((uint32_t)i < n) {
			const uint32_t sy = (w[i >> 3] >> ((i & 7) * 4)) & 15u;
			if (sy != prev) { if (len) o += enc_run(o, prev, len); prev = sy; len = 0; }
			++len;
		}
		if (len) o += enc_run(o, prev, len);
		// symbol counts of the chunk (symbols behind n masked away)
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t kv = n > (uint32_t)j * 8 ? n - j * 8 : 0;
			raw_add_word(kv >= 8 ? w[j] : (kv ? w[j] & ((1u << (kv * 4)) - 1u) : 0u), acc);
		}
		nsym += n;
	}
	// sums over the eight lanes of the group
#pragma unroll
	for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
		for (int q = 0; q < 5; ++q) acc[q] += __shfl_xor_sync(gmask, acc[q], o);
		nsym += __shfl_xor_sync(gmask, nsym, o);
	}
	if (gl == 0) { img[0] = (uint8_t)(nbytes & 0xff); img[1] = (uint8_t)(nbytes >> 8); }
	__syncwarp(gmask);
	uint4 *dst = reinterpret_cast<uint4*>(pool + (size_t)k * RB2_BLK);
	const uint4 *src = reinterpret_cast<const uint4*>(img);
#pragma unroll
	for (int j = 0; j < 4; ++j) dst[j * 8 + gl] = src[j * 8 + gl];
	if (gl < 6) {
		const Raw6 r = { acc[0], acc[1], acc[2], acc[3], acc[4], nsym };
		blkCnt[(size_t)k * 6 + gl] = raw_symbol(r, (uint32_t)gl);
	}
}
