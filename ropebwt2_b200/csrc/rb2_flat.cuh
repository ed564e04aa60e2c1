// rb2_flat.cuh -- the DENSE regime of the BCR column step: stream-rewrite of a fixed-width BWT.
//
// When a batch inserts into (nearly) every leaf block in every column -- short reads, the headline
// workload: 100 M records per column into <= 20 M blocks -- updating run-length coded blocks record
// by record is instruction bound (profiles/README.md).  For such batches the engine keeps the BWT as
// ONE flat array of nt6 codes, 3 bits per symbol in BIT PLANES: the array is a sequence of cells of 32
// symbols, a cell is three 32-bit words (bit 0, bit 1, bit 2 of the 32 codes; all six buckets
// concatenated), plus a directory of per-symbol counts in front of every FT_DIR-th symbol.  One column
// is a single streaming pass (k_flat_merge):
//
//   new[P_r + pre_r .. + c_r) = symbol of record r     (pre_r = members in front of record r, i.e. the
//   old symbol i moves to i + #record symbols with P<=i  symbols this column inserts in front of it)
//
// which is exactly the stable merge rope_insert_run performs one run at a time (rope.c:114-148;
// new symbols go in FRONT of the old symbol at the same position, mrope.c:206-218), and
// rank(a, P_r) -- the return value of rope_insert_run / rle_insert_cached (rle.c:10-89) -- is
// directory[P_r / FT_DIR][a] + a count over < FT_DIR + one slice of symbols held in shared memory.
// Why planes: counting a symbol over 32 codes is one AND/ANDN pair + one POPC (nibbles needed ~50
// instructions), inserting a symbol at a bit position is shift/mask work on three words, and the array
// is 25 % smaller than at 4 bits per symbol.
// Output-stationary: one warp produces one slice new[t*slice, (t+1)*slice) (2048 or 4096 symbols): the old
// symbols it needs are one contiguous range, fetched into shared memory by one TMA bulk copy
// (cp.async.bulk + mbarrier, a two-stage ring per warp), and the finished slice leaves through one TMA bulk
// store (rb2_flat_merge.inl).  The array stays resident between batches; it is re-encoded into leaf blocks of
// the reference's format (k_flat_encode) only when something needs blocks (iterator, dump, a sparse batch).
#pragma once
#include "rb2_codec.cuh"

#define FT_DIR   2048  // directory granularity (48 bytes of counts per 768 bytes of symbols)
#define FT_PAD   (8192 + FT_DIR + 256)  // readable / writable slack (symbols) behind a flat array (> one slice + one directory tile)
#define FT_CH    32    // symbols per cell (three words)
#define FE_CHUNK 64    // flat -> blocks: symbols encoded by one thread (<= 64 bytes of runs)
#define FE_T     (RB2_FILL - FE_CHUNK + 1) // block k of a bucket takes the chunks that start in bytes [k*FE_T, (k+1)*FE_T)

__host__ __device__ __forceinline__ uint64_t flat_bytes(uint64_t symbols) { return ((symbols + FT_CH - 1) / FT_CH) * 12; }
__host__ __device__ __forceinline__ uint32_t flat_get(const uint8_t *flat, uint64_t i)
{
	const uint32_t *w = reinterpret_cast<const uint32_t*>(flat) + (i >> 5) * 3; const uint32_t b = (uint32_t)(i & 31);
	return ((w[0] >> b) & 1u) | ((w[1] >> b) & 1u) << 1 | ((w[2] >> b) & 1u) << 2;
}

// ---- per-symbol counting ----------------------------------------------------------------------------
// nt6 codes: $=000 A=001 C=010 G=011 T=100 N=101 (110 and 111 never occur).  With the three planes
// b0, b1, b2 of a cell the positions of a symbol are one or two bitwise operations (cell_match), and the
// five popcounts s0 = |b0|, s1 = |b1|, s2 = |b2|, s01 = |b0&b1| (G), s02 = |b0&b2| (N) give all six
// counts: A = s0 - s01 - s02, C = s1 - s01, T = s2 - s02, $ = n - (A+C+G+T+N).  These "raw" counts are
// linear, so prefix sums are taken on them and converted only where a symbol count is needed.
struct Raw6 { uint32_t s0, s1, s2, s01, s02, n; };
struct Cell { uint32_t b0, b1, b2; };

__device__ __forceinline__ Cell cell_load(const uint32_t *w) { Cell c = { w[0], w[1], w[2] }; return c; }
__device__ __forceinline__ uint32_t cell_match(const Cell &c, uint32_t a)
{
	// positions whose code equals a: every plane agrees with the corresponding bit of a (branch-free: a differs from lane to lane)
	const uint32_t m0 = 0u - (a & 1u), m1 = 0u - ((a >> 1) & 1u), m2 = 0u - ((a >> 2) & 1u);
	return ~((c.b0 ^ m0) | (c.b1 ^ m1) | (c.b2 ^ m2));
}
// raw counts of the symbols selected by mask m (n = number of selected positions)
__device__ __forceinline__ Raw6 raw_of_cell(const Cell &c, uint32_t m, uint32_t n)
{
	const uint32_t x0 = c.b0 & m, x1 = c.b1 & m, x2 = c.b2 & m;
	Raw6 r = { (uint32_t)__popc(x0), (uint32_t)__popc(x1), (uint32_t)__popc(x2), (uint32_t)__popc(x0 & x1), (uint32_t)__popc(x0 & x2), n };
	return r;
}
__device__ __forceinline__ uint32_t low_mask(uint32_t n) { return n >= 32 ? 0xffffffffu : (1u << n) - 1u; } // the n lowest bits, 0 <= n <= 32

__device__ __forceinline__ uint32_t raw_symbol(const Raw6 &r, uint32_t a)
{
	const uint32_t A = r.s0 - r.s01 - r.s02, Cc = r.s1 - r.s01, G = r.s01, T = r.s2 - r.s02, N = r.s02;
	const uint32_t D = r.n - (A + Cc + G + T + N);
	uint32_t v = D; // selects, not branches: a differs from lane to lane
	v = a == 1 ? A : v; v = a == 2 ? Cc : v; v = a == 3 ? G : v; v = a == 4 ? T : v; v = a == 5 ? N : v;
	return v;
}
__device__ __forceinline__ void raw_addto(Raw6 &a, const Raw6 &b) { a.s0 += b.s0; a.s1 += b.s1; a.s2 += b.s2; a.s01 += b.s01; a.s02 += b.s02; a.n += b.n; }
// six raw counts as three words of two 16-bit fields (sums stay below 2^16 inside one tile)
__device__ __forceinline__ void raw_pack16(const Raw6 &r, uint32_t (&p)[3]) { p[0] = r.s0 | r.s1 << 16; p[1] = r.s2 | r.s01 << 16; p[2] = r.s02 | r.n << 16; }
__device__ __forceinline__ Raw6 raw_unpack16(uint32_t p0, uint32_t p1, uint32_t p2)
{
	Raw6 r = { p0 & 0xffffu, p0 >> 16, p1 & 0xffffu, p1 >> 16, p2 & 0xffffu, p2 >> 16 };
	return r;
}

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

// ---- slice geometry ----------------------------------------------------------------------------------------
// The unit of work of the merge is a SLICE: `slice` = 2048 or 4096 output symbols (2 or 4 output cells per lane),
// a whole number of directory tiles, produced by one warp.  One thread per slice boundary t (output position
// min(t*slice, nNew)): the first record whose output run starts at or behind it (bisection over the strictly
// increasing keys key_r = P_r + pre_r), the first old symbol that lands at or behind it, and the record run that
// reaches across it from the left.
struct alignas(16) TileDesc { uint64_t i0; uint32_t r0, carry; }; // carry = (symbols of the crossing run behind the boundary, capped at the slice size) << 3 | symbol

// sliceBkt (sharded engines, else null): the bucket of the slice's first record -- nearly every slice lies inside one
// bucket, so the merge looks the bucket up once per slice instead of bisecting the bucket table for every record
__global__ void __launch_bounds__(256) k_flat_geo(const RecView V, uint32_t R, uint64_t nSlices, uint64_t nNew, uint32_t slice, TileDesc *desc,
                                                  const Ctl *ctl, int nb, uint8_t *sliceBkt)
{
	const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (t > nSlices) return;
	const uint64_t o0 = t * slice < nNew ? t * slice : nNew;
	uint32_t lo = 0, hi = R; // first r with Key(r) >= t*slice (R if none)
	if (t == nSlices) lo = R;
	while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (V.Key(mid) >= t * slice) hi = mid; else lo = mid + 1; }
	const uint32_t r0 = lo;
	uint64_t before = 0; uint32_t carry = 0;
	if (r0 > 0) {
		const uint64_t pre = V.Pre(r0 - 1), key = (uint64_t)V.P[r0 - 1] + pre; const uint32_t sc = V.SC(r0 - 1);
		const uint64_t end = key + (sc >> 3);
		if (end > o0) { before = pre + (o0 - key); const uint64_t rem = end - o0; carry = (uint32_t)(rem < slice ? rem : slice) << 3 | (sc & 7u); }
		else before = pre + (sc >> 3);
	}
	TileDesc d; d.i0 = o0 - before; d.r0 = r0; d.carry = carry;
	desc[t] = d;
	if (sliceBkt) sliceBkt[t] = (uint8_t)bucket_of(ctl->recBkt, (uint32_t)nb, r0 < R ? r0 : (R ? R - 1 : 0));
}

// Sharded build, direct delivery: entry dst of this rank's column output (the rank of a string after this column)
// belongs in the state array of the rank that owns the string's next sub-bucket.  The output order is cut into at
// most 5 x 36 pieces (symbol x source sub-bucket), each contiguous on both sides; base[k] is the address in THIS
// rank's address space (a peer mapping over NVLink, rb2_comm.h p2p_map) of where entry 0 WOULD go if piece k
// started there, so that the store is base[k][dst].
// A record's piece follows from its symbol a and its (source) sub-bucket b: pieceOf[a*36+b].
#define ROUTE_MAXPC 192
// In an all-singleton column the string ids travel the same way (base32: the owner's id array), so such a column
// needs no send/recv at all.
struct PeerRoute { int64_t *base[ROUTE_MAXPC]; uint32_t *base32[ROUTE_MAXPC]; uint8_t pieceOf[6 * 36]; };

__global__ void k_route_store(PeerRoute *dst, const PeerRoute v)
{
	for (uint32_t k = threadIdx.x; k < ROUTE_MAXPC; k += blockDim.x) { dst->base[k] = v.base[k]; dst->base32[k] = v.base32[k]; }
	for (uint32_t k = threadIdx.x; k < 6 * 36; k += blockDim.x) dst->pieceOf[k] = v.pieceOf[k];
}

struct FlatArgs {
	const uint8_t *oldS; const int64_t *oldDir;   // old array, counts in front of every FT_DIR-th old symbol
	uint8_t *newS; uint64_t nNew; uint32_t *newTileCnt; // new array and its per-FT_DIR-tile symbol counts
	RecView V; const uint32_t *recDst; uint32_t R;
	const TileDesc *desc; uint32_t nSlices;
	int64_t *gLNext; const Ctl *ctl;
	// sharded engines: records carry whole-index positions; bucket b of this rank sits recOff[b*7+6]
	// symbols (recOff[b*7+a] symbols a) further right in the whole index than in the local array
	const int64_t *recOff; int nb;
	const PeerRoute *route; // sharded, direct delivery: where gLNext[dst] really lives (null: gLNext is a local array)
	const uint8_t *sliceBkt; // sharded: bucket of every slice's first record (k_flat_geo)
	const uint32_t *sidCur;  // direct delivery, all-singleton column: string id of record r, delivered through route->base32
};

// ---- the merge: one warp per slice, no block-wide synchronisation ---------------------------------------------
// A slice needs: the old symbols that land in it, from the directory boundary a0 in front of its first one
// (<= FT_DIR + slice symbols, one TMA bulk copy), and its records (read from the record arrays).  The warp
//   (1) counts the old symbols cell by cell (raw prefix counts from a0: what rank() needs),
//   (2) scatters its records into per-output-cell bit masks (which output positions are new, and their planes),
//   (3) assembles its output cells, two or four per lane: 32 old symbols from the right offset (funnel shift),
//       one zero bit pushed in per new position, the new symbols' planes OR-ed on top,
//   (4) sends the slice off with one TMA bulk store, writes the raw symbol counts of its directory tiles,
//   (5) returns rank(a, P) = directory row + prefix count + partial cell count for each of its records.
// The kernel is persistent: every warp owns two shared-memory stages and fetches the old symbols of its slice
// i+2 while it merges slice i (its lane 0 is the producer); the slice's records are read straight from the
// record arrays (consecutive records in consecutive lanes).  No block-wide barrier anywhere.
#define FS_STAGES 2
#define FS_WARPS 4                                  // warps per CTA of the merge kernel

// where the records of one slice are, indexed from 0 (shared memory or global memory)
struct SliceIn {
	const uint32_t *old; const int64_t *P; const uint32_t *pre, *sc, *dst; const uint8_t *asym; uint32_t r0;
	__device__ __forceinline__ uint32_t Pre(uint32_t k) const { return pre ? pre[k] : r0 + k; }
	__device__ __forceinline__ uint32_t SC(uint32_t k) const { return sc ? sc[k] : (8u | asym[k]); }
};

// record k = lane of a slice, fetched one slice ahead (the loads' latency hides under the previous slice's merge)
struct RecRegs { int64_t P; uint32_t pre, sc, dst, sid; bool have; };

// one output cell: 32 old symbols from local index oldIdx on, a gap pushed in at every bit of m, the new planes on top
__device__ __forceinline__ Cell slice_cell(const uint32_t *old, uint32_t oldIdx, const uint32_t (&mk)[4])
{
	Cell x;
	const uint32_t *wp = old + (oldIdx >> 5) * 3; const uint32_t sh = oldIdx & 31;
	x.b0 = __funnelshift_r(wp[0], wp[3], sh); x.b1 = __funnelshift_r(wp[1], wp[4], sh); x.b2 = __funnelshift_r(wp[2], wp[5], sh);
	uint32_t m = mk[0];
	if (m == 0xffffffffu) { x.b0 = 0; x.b1 = 0; x.b2 = 0; }
	else while (m) { // increasing positions: each gap shifts what is behind it by one
		const uint32_t low = (m & (0u - m)) - 1u; // the bits below the lowest set bit of m
		x.b0 = (x.b0 & low) | ((x.b0 & ~low) << 1); x.b1 = (x.b1 & low) | ((x.b1 & ~low) << 1); x.b2 = (x.b2 & low) | ((x.b2 & ~low) << 1);
		m &= m - 1;
	}
	x.b0 = (x.b0 & ~mk[0]) | mk[1]; x.b1 = (x.b1 & ~mk[0]) | mk[2]; x.b2 = (x.b2 & ~mk[0]) | mk[3];
	return x;
}

#define FS_CPL 2
namespace fs2 {
#include "rb2_flat_merge.inl"
}
#define FS_CPL 4
namespace fs4 {
#include "rb2_flat_merge.inl"
}

struct FlatDirScan { // K=6 (int64): per-tile raw counts (three packed words) -> symbol counts in front of every tile
	const uint32_t *tileCnt; uint64_t nTile; int64_t *dir;
	__device__ void load(uint64_t i, int64_t (&v)[6]) const {
		const Raw6 r = raw_unpack16(tileCnt[i * 3], tileCnt[i * 3 + 1], tileCnt[i * 3 + 2]);
#pragma unroll
		for (int a = 0; a < 6; ++a) v[a] = raw_symbol(r, (uint32_t)a);
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
	// one warp per FT_DIR tile: 32 lanes x 2 cells
	const int lane = threadIdx.x & 31;
	const uint64_t tile = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
	if (tile * FT_DIR >= n && tile != 0) return;
	const uint64_t base = tile * FT_DIR + lane * 2 * FT_CH;
	const uint32_t rem = base >= n ? 0u : (n - base < 2 * FT_CH ? (uint32_t)(n - base) : 2u * FT_CH);
	const uint32_t *src = reinterpret_cast<const uint32_t*>(flat) + (base / FT_CH) * 3;
	const uint32_t n0 = rem < FT_CH ? rem : FT_CH, n1 = rem > FT_CH ? rem - FT_CH : 0;
	Raw6 s = { 0, 0, 0, 0, 0, 0 };
	if (n0) s = raw_of_cell(cell_load(src), low_mask(n0), n0);
	if (n1) raw_addto(s, raw_of_cell(cell_load(src + 3), low_mask(n1), n1));
	s.s0 = warp_sum(s.s0); s.s1 = warp_sum(s.s1); s.s2 = warp_sum(s.s2); s.s01 = warp_sum(s.s01); s.s02 = warp_sum(s.s02); s.n = warp_sum(s.n);
	if (lane < 3) { uint32_t pk[3]; raw_pack16(s, pk); tileCnt[tile * 3 + lane] = lane == 0 ? pk[0] : (lane == 1 ? pk[1] : pk[2]); }
}

// occ(a, x) on the flat array, all six symbols, one warp; result in every lane
__device__ __forceinline__ void flat_rank6(const uint8_t *flat, const int64_t *dir, int64_t x, int lane, int64_t (&out)[6])
{
	const uint64_t t = (uint64_t)x / FT_DIR;
	const uint32_t part = (uint32_t)((uint64_t)x - t * FT_DIR);
	const uint32_t rem = part > (uint32_t)lane * 2 * FT_CH ? (part - lane * 2 * FT_CH < 2 * FT_CH ? part - lane * 2 * FT_CH : 2u * FT_CH) : 0u;
	Raw6 s = { 0, 0, 0, 0, 0, 0 };
	if (rem) {
		const uint32_t *src = reinterpret_cast<const uint32_t*>(flat) + ((t * FT_DIR + lane * 2 * FT_CH) / FT_CH) * 3;
		const uint32_t n0 = rem < FT_CH ? rem : FT_CH, n1 = rem > FT_CH ? rem - FT_CH : 0;
		s = raw_of_cell(cell_load(src), low_mask(n0), n0);
		if (n1) raw_addto(s, raw_of_cell(cell_load(src + 3), low_mask(n1), n1));
	}
	s.s0 = warp_sum(s.s0); s.s1 = warp_sum(s.s1); s.s2 = warp_sum(s.s2); s.s01 = warp_sum(s.s01); s.s02 = warp_sum(s.s02); s.n = warp_sum(s.n);
#pragma unroll
	for (int a = 0; a < 6; ++a) out[a] = dir[t * 6 + a] + raw_symbol(s, (uint32_t)a);
}

// sizes6[g][a] = #a in [gL, gL+gSize) for every group with a non-empty interval (rope_rank2a, mrope.c:202).
// Short intervals (the rule: an interval holds the suffixes that share the string's suffix read so far,
// at most the coverage of the data after the first columns) are counted directly by the owning thread;
// long ones by the whole warp from the directory.
// posOff: sharded engines pass whole-index positions; bucket b's local position = position - posOff[b*7+6].
#define FT_SHORT_IV 128
__global__ void __launch_bounds__(128) k_flat_rank_groups(const uint8_t *flat, const int64_t *dir, uint32_t G, const int64_t *gL, const int64_t *gSize,
                                                          int64_t *sizes6, const Ctl *ctl, const int64_t *posOff, int nb)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t g0 = (blockIdx.x * 4 + wid) * 32;
	if (g0 >= G) return;
	const uint32_t g = g0 + lane;
	int64_t myL = g < G ? gL[g] : 0; const int64_t mySz = g < G ? gSize[g] : 0;
	if (posOff && g < G && mySz > 0) myL -= posOff[bucket_of(ctl->gBkt, (uint32_t)nb, g) * 7 + 6];
	if (mySz > 0 && mySz <= FT_SHORT_IV) {
		uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
		uint64_t p = (uint64_t)myL; const uint64_t end = p + (uint64_t)mySz;
		while (p < end) {
			const uint32_t lo = (uint32_t)(p & (FT_CH - 1));
			const uint32_t n = end - p < FT_CH - lo ? (uint32_t)(end - p) : FT_CH - lo;
			const Raw6 r = raw_of_cell(cell_load(reinterpret_cast<const uint32_t*>(flat) + (p / FT_CH) * 3), low_mask(n) << lo, n);
#pragma unroll
			for (int a = 0; a < 6; ++a) c[a] += raw_symbol(r, (uint32_t)a);
			p += n;
		}
#pragma unroll
		for (int a = 0; a < 6; ++a) sizes6[(size_t)g * 6 + a] = c[a];
	}
	uint32_t todo = __ballot_sync(FULLMASK, mySz > FT_SHORT_IV);
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

// ---- one kernel per all-singleton column -------------------------------------------------------------------
// Once every group is a single string (always in input order; in the sorted modes as soon as all suffixes read
// so far are distinct) a column is: partition the strings by their next symbol (mrope.c:303-310), and -- with
// non-empty intervals, i.e. a sorted mode on a non-empty index -- find each string's insertion point inside
// its interval (rope_rank2a, mrope.c:199-218).  k_column_fused does all of it in ONE pass over the strings:
//   * the symbol a string inserts in this column was fetched by the PREVIOUS column's kernel while it moved
//     the string (asym is stored in string order, so it is read sequentially here); k_tile_hist + one small scan
//     give the per-tile prefixes and the grand totals per symbol, from which every destination index follows;
//   * the string's NEXT symbol is gathered from the column-major symbol matrix and travels with it.
// SIZED: the interval [gL, gL+gSize) of the flat array is counted by the string's thread (short intervals) or
// by its warp from the directory (long ones); the record position goes to recP, the next interval size along.
#define CF_TILE 1024   // strings per CTA (256 threads x 4) = MEM_TILE

// per-tile histogram of the symbols of this column: tileTot[tile][6]
__global__ void __launch_bounds__(256) k_tile_hist(const uint8_t *asym, uint32_t M, uint32_t *tileTot)
{
	__shared__ uint32_t sm[6 * 8];
	const uint32_t k = blockIdx.x * CF_TILE + threadIdx.x * 4;
	uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
	if (k < M) {
		const uint32_t a4 = *reinterpret_cast<const uint32_t*>(asym + k); // asym is padded to a multiple of 4
#pragma unroll
		for (int i = 0; i < 4; ++i) if (k + i < M) {
			const uint32_t a = (a4 >> (8 * i)) & 0xff;
#pragma unroll
			for (int x = 0; x < 6; ++x) c[x] += a == (uint32_t)x;
		}
	}
#pragma unroll
	for (int x = 0; x < 6; ++x) {
		const uint32_t t = warp_redux_add(c[x]);
		if ((threadIdx.x & 31) == 0) sm[x * 8 + (threadIdx.x >> 5)] = t;
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		uint32_t t = 0;
		for (int w = 0; w < 8; ++w) t += sm[threadIdx.x * 8 + w];
		tileTot[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
	}
}

struct FusedArgs {
	const uint32_t *sid; const uint8_t *asym; uint32_t M;
	const uint32_t *tilePre;      // exclusive prefix of k_tile_hist's rows; ctl->memTot = the grand totals
	const uint8_t *Tnext;         // the next column of the symbol matrix (4 bits per symbol), null behind the last column
	const int64_t *gL, *gSize; const uint8_t *flat; const int64_t *dir; // SIZED
	const Ctl *ctl;
	uint32_t *sidNext; uint8_t *asymNext; uint32_t *recDst; int64_t *recP, *gSizeNext;
};

template <bool SIZED, bool COMP>
__global__ void __launch_bounds__(256) k_column_fused(FusedArgs A)
{
	__shared__ uint64_t sm[2 * 8];
	const int lane = threadIdx.x & 31;
	const uint32_t k = blockIdx.x * CF_TILE + threadIdx.x * 4;
	uint32_t a4 = 0, n4 = 0, id[4] = { 0, 0, 0, 0 };
	uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
	if (k < A.M) {
		a4 = *reinterpret_cast<const uint32_t*>(A.asym + k); // asym is padded to a multiple of 4
		if (k + 4 <= A.M) { const uint4 v = *reinterpret_cast<const uint4*>(A.sid + k); id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w; }
		else for (int i = 0; i < 4; ++i) id[i] = k + i < A.M ? A.sid[k + i] : 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) if (k + i < A.M) {
			const uint32_t a = (a4 >> (8 * i)) & 0xff;
			if (a && A.Tnext) n4 |= ((A.Tnext[id[i] >> 1] >> ((id[i] & 1) * 4)) & 15u) << (8 * i);
#pragma unroll
			for (int x = 0; x < 6; ++x) c[x] += a == (uint32_t)x;
		}
	}
	// exclusive prefix inside the CTA; a CTA holds 1024 strings, so counters share 64-bit words (16 bits each)
	uint64_t pk[2] = { (uint64_t)c[1] | (uint64_t)c[2] << 16 | (uint64_t)c[3] << 32 | (uint64_t)c[4] << 48, (uint64_t)c[5] }, pt[2];
	cta_excl_scan<2, 256, uint64_t>(pk, pt, sm);
	if (k >= A.M) return;
	uint32_t base[6];
	{
		uint32_t m = 0;
		base[0] = 0;
#pragma unroll
		for (int x = 1; x < 6; ++x) {
			const uint32_t ex = x < 5 ? (uint32_t)(pk[0] >> (16 * (x - 1))) & 0xffffu : (uint32_t)pk[1];
			base[x] = m + A.tilePre[(size_t)blockIdx.x * 6 + x] + ex;
			m += A.ctl->memTot[x];
		}
	}
	constexpr int ord[6] = { 0, COMP ? 4 : 1, COMP ? 3 : 2, COMP ? 2 : 3, COMP ? 1 : 4, 5 }; // mrope.c:209-210
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const bool valid = k + i < A.M;
		const uint32_t g = k + i, a = (a4 >> (8 * i)) & 0xff;
		uint32_t d = NONE32;
		if (valid) {
#pragma unroll
			for (int x = 1; x < 6; ++x) if (a == (uint32_t)x) d = base[x]++;
		}
		if (SIZED) {
			int64_t P = valid ? A.gL[g] : 0, sza = 0;
			const int64_t sz = valid ? A.gSize[g] : 0;
			if (sz > 0 && sz <= FT_SHORT_IV) {
				Raw6 acc = { 0, 0, 0, 0, 0, 0 };
				uint64_t p = (uint64_t)P; const uint64_t end = p + (uint64_t)sz;
				while (p < end) {
					const uint32_t lo = (uint32_t)(p & (FT_CH - 1));
					const uint32_t n = end - p < FT_CH - lo ? (uint32_t)(end - p) : FT_CH - lo;
					raw_addto(acc, raw_of_cell(cell_load(reinterpret_cast<const uint32_t*>(A.flat) + (p / FT_CH) * 3), low_mask(n) << lo, n));
					p += n;
				}
#pragma unroll
				for (int slot = 0; slot < 6; ++slot) { // insertion point: behind the old symbols of the earlier slots
					const uint32_t z = raw_symbol(acc, (uint32_t)ord[slot]);
					if ((uint32_t)ord[slot] == a) { sza = z; break; }
					P += z;
				}
			}
			uint32_t todo = __ballot_sync(FULLMASK, sz > FT_SHORT_IV);
			while (todo) { // long intervals: the whole warp, from the directory
				const int src = __ffs(todo) - 1; todo &= todo - 1;
				const int64_t L = __shfl_sync(FULLMASK, P, src), szv = __shfl_sync(FULLMASK, sz, src);
				int64_t cl[6], cu[6];
				flat_rank6(A.flat, A.dir, L, lane, cl);
				flat_rank6(A.flat, A.dir, L + szv, lane, cu);
				if (lane == src) {
#pragma unroll
					for (int slot = 0; slot < 6; ++slot) {
						const int64_t z = cu[ord[slot]] - cl[ord[slot]];
						if ((uint32_t)ord[slot] == a) { sza = z; break; }
						P += z;
					}
				}
			}
			if (valid) {
				A.recP[g] = P;
				if (a) A.gSizeNext[d] = sza;
			}
		}
		if (valid) {
			if (a) { A.sidNext[d] = id[i]; A.asymNext[d] = (uint8_t)((n4 >> (8 * i)) & 0xff); }
			A.recDst[g] = d;
		}
	}
}

// batched rank queries on the resident array (rb2_rank_batch): one warp per position, grid-stride
__global__ void __launch_bounds__(128) k_flat_rank_batch(const uint8_t *flat, const int64_t *dir, uint32_t n, const int64_t *x, int64_t *out)
{
	const int lane = threadIdx.x & 31;
	for (uint32_t q = blockIdx.x * 4 + (threadIdx.x >> 5); q < n; q += gridDim.x * 4) {
		int64_t c[6];
		flat_rank6(flat, dir, x[q], lane, c);
		if (lane < 6) {
			int64_t v = 0;
#pragma unroll
			for (int a = 0; a < 6; ++a) if (lane == a) v = c[a];
			out[(size_t)q * 6 + lane] = v;
		}
	}
}

// per-bucket symbol totals of the array: out[k][a] = occ(a, pos[k]) for up to 64 positions (one warp each)
struct RankAtPos { int64_t pos[8]; };
__global__ void __launch_bounds__(32) k_flat_rank_at(const uint8_t *flat, const int64_t *dir, const RankAtPos P, int64_t *out)
{
	int64_t c[6];
	flat_rank6(flat, dir, P.pos[blockIdx.x], threadIdx.x, c);
	if (threadIdx.x < 6) {
		int64_t v = 0;
#pragma unroll
		for (int a = 0; a < 6; ++a) if ((int)threadIdx.x == a) v = c[a];
		out[blockIdx.x * 6 + threadIdx.x] = v;
	}
}

// ---- leaf blocks -> flat --------------------------------------------------------------------------------
// One warp per logical block: every lane expands the runs that start in its 16 bytes straight into the
// (zeroed) planes -- a run sets a bit range in the planes its symbol has a 1 in; the two end words of a
// range may be shared with neighbouring runs (atomicOr), the words in between are owned by the run.
// (sharded engines: off[b*7+6] = symbols of the whole index in front of bucket b that other ranks hold)
__device__ __forceinline__ void plane_set_range(uint32_t *flatW, int plane, uint64_t s, uint32_t l)
{
	const uint64_t e = s + l; // bits [s, e)
	uint64_t c = s / FT_CH; const uint64_t cl = (e - 1) / FT_CH;
	const uint32_t lo = (uint32_t)(s & (FT_CH - 1));
	if (c == cl) { atomicOr(flatW + c * 3 + plane, low_mask(l) << lo); return; }
	atomicOr(flatW + c * 3 + plane, 0xffffffffu << lo);
	for (++c; c < cl; ++c) flatW[c * 3 + plane] = 0xffffffffu;
	atomicOr(flatW + cl * 3 + plane, low_mask((uint32_t)(e - cl * FT_CH)));
}

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
	uint64_t dst = (uint64_t)(cumLen[i] - posBase) + basePos;
	uint32_t *flatW = reinterpret_cast<uint32_t*>(flat);
	uint32_t bp = lane * 16 + d.fb;
	for (uint32_t q = 0; q < d.nr; ++q) {
		uint32_t l, s, nbr;
		parse_run(sImg[wid], bp, s, l, nbr);
		bp += nbr;
		if (l) {
			if (s & 1u) plane_set_range(flatW, 0, dst, l);
			if (s & 2u) plane_set_range(flatW, 1, dst, l);
			if (s & 4u) plane_set_range(flatW, 2, dst, l);
		}
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

// the FE_CHUNK symbols that start at symbol s0, as three 64-bit planes; `starts` = the positions (< n)
// where a run starts (position 0 and wherever a symbol differs from its predecessor)
struct Chunk64 { uint64_t b0, b1, b2, starts; };
__device__ __forceinline__ Chunk64 load_chunk64(const uint8_t *flat, uint64_t s0, uint32_t n)
{
	const uint32_t *wp = reinterpret_cast<const uint32_t*>(flat) + (s0 / FT_CH) * 3;
	const uint32_t sh = (uint32_t)(s0 & (FT_CH - 1));
	uint64_t b[3];
#pragma unroll
	for (int p = 0; p < 3; ++p) {
		const uint32_t x0 = wp[p], x1 = wp[3 + p], x2 = wp[6 + p];
		b[p] = (uint64_t)__funnelshift_r(x1, x2, sh) << 32 | __funnelshift_r(x0, x1, sh);
	}
	const uint64_t vm = n >= 64 ? ~0ull : (1ull << n) - 1ull;
	Chunk64 c;
	c.b0 = b[0] & vm; c.b1 = b[1] & vm; c.b2 = b[2] & vm;
	c.starts = (((c.b0 ^ (c.b0 << 1)) | (c.b1 ^ (c.b1 << 1)) | (c.b2 ^ (c.b2 << 1))) | 1ull) & vm;
	return c;
}

// bytes of the encoded chunk
__global__ void __launch_bounds__(256) k_flat_chunk_bytes(const uint8_t *flat, EncTab T, uint64_t nChunk, uint8_t *chunkBytes)
{
	const uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (c >= nChunk) return;
	const int b = enc_bucket_of_chunk(T, c);
	const uint64_t s0 = T.symStart[b] + (c - T.chunkStart[b]) * FE_CHUNK;
	const uint32_t n = (uint32_t)(s0 + FE_CHUNK < T.symStart[b + 1] ? FE_CHUNK : T.symStart[b + 1] - s0);
	const Chunk64 ch = load_chunk64(flat, s0, n);
	// a run of 16 or more symbols takes two bytes: those are the starts followed by 15 non-starts
	uint64_t ns = ~ch.starts, long15 = ns >> 1;
#pragma unroll
	for (int k = 2; k <= 15; ++k) long15 &= ns >> k;   // bit i: positions i+1 .. i+15 are no starts
	const uint64_t in15 = n >= 16 ? ((n - 15 >= 64 ? ~0ull : (1ull << (n - 15)) - 1ull)) : 0ull; // start i with i + 15 < n
	chunkBytes[c] = (uint8_t)(__popcll(ch.starts) + __popcll(ch.starts & long15 & in15));
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
	Raw6 acc = { 0, 0, 0, 0, 0, 0 };
	for (uint64_t c = c0 + gl; c < c1; c += 8) {
		const uint64_t s0 = T.symStart[b] + (c - cLo) * FE_CHUNK;
		const uint32_t n = (uint32_t)(s0 + FE_CHUNK < T.symStart[b + 1] ? FE_CHUNK : T.symStart[b + 1] - s0);
		const Chunk64 ch = load_chunk64(flat, s0, n);
		uint8_t *o = img + 2 + (uint32_t)(chunkPre[c] - b0);
		uint64_t st = ch.starts;
		while (st) { // run by run
			const uint32_t i = (uint32_t)__ffsll((long long)st) - 1;
			st &= st - 1;
			const uint32_t nx = st ? (uint32_t)__ffsll((long long)st) - 1 : n;
			const uint32_t sy = (uint32_t)((ch.b0 >> i) & 1ull) | (uint32_t)((ch.b1 >> i) & 1ull) << 1 | (uint32_t)((ch.b2 >> i) & 1ull) << 2;
			o += enc_run(o, sy, nx - i);
		}
		// symbol counts of the chunk
		acc.s0 += __popcll(ch.b0); acc.s1 += __popcll(ch.b1); acc.s2 += __popcll(ch.b2);
		acc.s01 += __popcll(ch.b0 & ch.b1); acc.s02 += __popcll(ch.b0 & ch.b2); acc.n += n;
	}
	// sums over the eight lanes of the group
#pragma unroll
	for (int o = 4; o > 0; o >>= 1) {
		acc.s0 += __shfl_xor_sync(gmask, acc.s0, o); acc.s1 += __shfl_xor_sync(gmask, acc.s1, o); acc.s2 += __shfl_xor_sync(gmask, acc.s2, o);
		acc.s01 += __shfl_xor_sync(gmask, acc.s01, o); acc.s02 += __shfl_xor_sync(gmask, acc.s02, o); acc.n += __shfl_xor_sync(gmask, acc.n, o);
	}
	if (gl == 0) { img[0] = (uint8_t)(nbytes & 0xff); img[1] = (uint8_t)(nbytes >> 8); }
	__syncwarp(gmask);
	uint4 *dst = reinterpret_cast<uint4*>(pool + (size_t)k * RB2_BLK);
	const uint4 *src = reinterpret_cast<const uint4*>(img);
#pragma unroll
	for (int j = 0; j < 4; ++j) dst[j * 8 + gl] = src[j * 8 + gl];
	if (gl < 6) blkCnt[(size_t)k * 6 + gl] = raw_symbol(acc, (uint32_t)gl);
}
