// rb2_flat_merge.inl -- the slice merge of the dense regime for ONE slice size; included by rb2_flat.cuh once per
// size (FS_CPL = output cells per lane = 2 or 4, i.e. 2048 or 4096 symbols per warp), each time inside its own namespace.
// Small slices keep more warps resident (32 per SM) and win when a column inserts many symbols per slice; large ones
// spread the fixed cost of a slice (geometry, prefix scan, waits, the directory slack in front of the slice) over
// twice the symbols and win once the index is large against the batch (measured: profiles/README.md).
#define FS_SLICE (FS_CPL * 1024)
#define FS_CELLS (FS_SLICE / FT_CH)                 // output cells per slice, FS_CPL per lane
#define FS_TILES (FS_SLICE / FT_DIR)                // directory tiles per slice
#define FS_LPT   (32 / FS_TILES)                    // lanes per directory tile
#define FS_PCL   (FS_CPL + FT_DIR / 1024)           // old cells per lane in the prefix pass
#define FS_OLDC  (FS_PCL * 32 + 2)                  // cells of old symbols a slice can need (+ funnel-shift partner)
#define FS_MINCTA (FS_CPL == 2 ? 8 : (FS_CPL == 4 ? 5 : 3)) // CTAs per SM the shared memory allows
constexpr uint32_t kSlice = FS_SLICE, kMinCta = FS_MINCTA;
#define FS_OLDW  ((FS_OLDC * 3 + 3) & ~3)           // ... as words, a multiple of 16 bytes
#define FS_OUTW  (FS_CELLS * 3)                     // words of one output slice

struct SliceStage { alignas(16) uint32_t old[FS_OLDW]; TileDesc d0, d1; uint32_t slice, pad[3]; };
struct SliceWork {
	alignas(16) uint32_t out[FS_OUTW];              // the finished slice (TMA bulk store)
	alignas(16) uint32_t mask[FS_CELLS][4];         // per output cell: new positions, and planes 0..2 of the new symbols
	uint32_t pre[FS_OLDC][3];                       // raw counts in front of every old cell, from a0 (16-bit fields)
};
struct SliceWarpSmem { SliceStage st[FS_STAGES]; SliceWork W; alignas(8) uint64_t full[FS_STAGES]; };

// set the new-symbol masks of output positions [key, key+len) of the slice
__device__ __forceinline__ void slice_mark(SliceWork &W, uint32_t key, uint32_t len, uint32_t sy)
{
	while (len) {
		const uint32_t c = key / FT_CH, q = key & (FT_CH - 1), n = len < FT_CH - q ? len : FT_CH - q;
		const uint32_t bits = low_mask(n) << q;
		atomicOr(&W.mask[c][0], bits);
		if (sy & 1u) atomicOr(&W.mask[c][1], bits);
		if (sy & 2u) atomicOr(&W.mask[c][2], bits);
		if (sy & 4u) atomicOr(&W.mask[c][3], bits);
		key += n; len -= n;
	}
}

// SHARDED: the engine is one rank of a sharded build (whole-index coordinates, direct delivery); compiled out otherwise
template <bool GENERAL, bool SHARDED>
__device__ __forceinline__ void flat_merge_slice(const FlatArgs &A, SliceWork &W, const SliceIn &in, const uint32_t slice, const TileDesc d0, const TileDesc d1, const int lane, const RecRegs &pf, const int64_t cpostLane)
{
	const uint64_t o0 = (uint64_t)slice * FS_SLICE;
	const uint32_t sliceLen = A.nNew - o0 < FS_SLICE ? (uint32_t)(A.nNew - o0) : FS_SLICE;
	const uint32_t r0 = d0.r0, nr = d1.r0 - d0.r0;
	const uint32_t carrySym = d0.carry & 7u, carryLen = (d0.carry >> 3) < sliceLen ? (d0.carry >> 3) : sliceLen;
	const uint64_t a0 = d0.i0 & ~(uint64_t)(FT_DIR - 1);
	const uint32_t skip = (uint32_t)(d0.i0 - a0);      // old symbols of the window in front of the slice's first one
	// lane a < 6: start of bucket a behind this column + #a in front of the directory boundary (fetched now, used by the
	// rank epilogue: the load's latency hides under the merge)
	int64_t baseLane = lane < 6 ? cpostLane + A.oldDir[(a0 / FT_DIR) * 6 + lane] : 0;
	// sharded engines: the bucket of the slice's records gives the shift into whole-index coordinates and (direct delivery)
	// the peer array each symbol's ranks go to; one bucket for the whole slice except where a bucket boundary crosses it
	bool oneBkt = true;
	int64_t *routeLane = 0; uint32_t *route32Lane = 0;
	if (SHARDED) {
		const uint32_t b0 = A.sliceBkt[slice];
		oneBkt = r0 + nr <= A.ctl->recBkt[b0 + 1];
		if (lane < 6) {
			if (oneBkt) baseLane += A.recOff[b0 * 7 + lane];
			if (A.route) {
				const uint32_t pc = A.route->pieceOf[lane * 36 + b0];
				routeLane = A.route->base[pc];
				if (!GENERAL && A.sidCur) route32Lane = A.route->base32[pc];
			}
		}
	}
	// ---- (1) masks := 0; raw prefix counts of the old cells (lane l: cells l*FS_PCL .. +FS_PCL-1) -----------
#pragma unroll
	for (int j = 0; j < FS_CPL; ++j) reinterpret_cast<uint4*>(&W.mask[0][0])[lane + 32 * j] = make_uint4(0, 0, 0, 0);
	{
		uint32_t pk[FS_PCL][3], inc[3];
		Raw6 r = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
		for (int j = 0; j < FS_PCL; ++j) {
			raw_pack16(r, pk[j]); // exclusive inside the lane
			raw_addto(r, raw_of_cell(cell_load(in.old + (lane * FS_PCL + j) * 3), 0xffffffffu, FT_CH));
		}
		raw_pack16(r, inc);
		const uint32_t own[3] = { inc[0], inc[1], inc[2] };
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
			for (int k = 0; k < 3; ++k) { const uint32_t y = __shfl_up_sync(FULLMASK, inc[k], o); if (lane >= o) inc[k] += y; }
		}
#pragma unroll
		for (int j = 0; j < FS_PCL; ++j) {
#pragma unroll
			for (int k = 0; k < 3; ++k) W.pre[lane * FS_PCL + j][k] = inc[k] - own[k] + pk[j][k];
		}
	}
	__syncwarp();
	// ---- (2) records -> masks ----------------------------------------------------------------------------
	if (lane == 0 && carryLen) slice_mark(W, 0, carryLen, carrySym);
	for (uint32_t k = lane; k < nr; k += 32) {
		const bool reg = pf.have && k < 32;
		const uint32_t key = (uint32_t)((uint64_t)(reg ? pf.P : in.P[k]) + (reg ? pf.pre : in.Pre(k)) - o0);
		if (GENERAL) {
			const uint32_t sc = reg ? pf.sc : in.SC(k);
			uint32_t len = sc >> 3;
			if (len > FS_SLICE - key) len = FS_SLICE - key;
			slice_mark(W, key, len, sc & 7u);
		} else {
			const uint32_t sy = reg ? pf.sc : (uint32_t)in.asym[k], c = key / FT_CH, bit = 1u << (key & (FT_CH - 1));
			atomicOr(&W.mask[c][0], bit);
			if (sy & 1u) atomicOr(&W.mask[c][1], bit);
			if (sy & 2u) atomicOr(&W.mask[c][2], bit);
			if (sy & 4u) atomicOr(&W.mask[c][3], bit);
		}
	}
	if (lane == 0) bulk_wait_read(); // the previous slice's store has read W.out
	__syncwarp();
	// ---- (3) assemble: lane l makes output cells l*FS_CPL .. +FS_CPL-1 ---------------------------------------
	{
		uint4 mk4[FS_CPL];
		uint32_t cnt = 0;
#pragma unroll
		for (int j = 0; j < FS_CPL; ++j) { mk4[j] = reinterpret_cast<const uint4*>(&W.mask[0][0])[lane * FS_CPL + j]; cnt += __popc(mk4[j].x); }
		uint32_t ex = warp_incl_scan(cnt, lane) - cnt;   // new symbols of the slice in front of this lane's first cell
		uint32_t ow[FS_CPL * 3];
		Raw6 r = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
		for (int j = 0; j < FS_CPL; ++j) {
			const uint32_t mk[4] = { mk4[j].x, mk4[j].y, mk4[j].z, mk4[j].w };
			const uint32_t rel = (uint32_t)(lane * FS_CPL + j) * FT_CH;
			Cell x = slice_cell(in.old, skip + rel - ex, mk);
			ex += __popc(mk[0]);
			uint32_t nv = FT_CH, vm = 0xffffffffu;
			if (sliceLen < FS_SLICE) { // (warp-uniform: the last slice only) symbols behind the end of the array are zero
				nv = rel >= sliceLen ? 0u : (sliceLen - rel < FT_CH ? sliceLen - rel : FT_CH);
				vm = low_mask(nv);
				x.b0 &= vm; x.b1 &= vm; x.b2 &= vm;
			}
			ow[3 * j] = x.b0; ow[3 * j + 1] = x.b1; ow[3 * j + 2] = x.b2;
			raw_addto(r, raw_of_cell(x, vm, nv));
		}
		uint32_t *o = W.out + lane * (FS_CPL * 3);
		if (FS_CPL % 4 == 0) {
#pragma unroll
			for (int i = 0; i < FS_CPL * 3; i += 4) *reinterpret_cast<uint4*>(o + i) = make_uint4(ow[i], ow[i + 1], ow[i + 2], ow[i + 3]);
		} else {
#pragma unroll
			for (int i = 0; i < FS_CPL * 3; i += 2) *reinterpret_cast<uint2*>(o + i) = make_uint2(ow[i], ow[i + 1]);
		}
		fence_proxy_async();
		// ---- (4) symbol counts of the slice's directory tiles (FS_LPT lanes each): raw counts, converted by FlatDirScan ----
		uint32_t pk[3];
		raw_pack16(r, pk);
		const uint32_t gm = FS_LPT == 32 ? FULLMASK : (((1u << (FS_LPT & 31)) - 1u) << ((lane / FS_LPT) * FS_LPT));
#pragma unroll
		for (int k = 0; k < 3; ++k) pk[k] = __reduce_add_sync(gm, pk[k]);
		const uint64_t dt = (uint64_t)slice * FS_TILES + lane / FS_LPT;
		const uint32_t gl = lane % FS_LPT;
		if (gl < 3 && (dt * FT_DIR < A.nNew || dt == 0)) A.newTileCnt[dt * 3 + gl] = gl == 0 ? pk[0] : (gl == 1 ? pk[1] : pk[2]);
	}
	__syncwarp();
	if (lane == 0) { bulk_s2g(A.newS + (uint64_t)slice * (FS_OUTW * 4), W.out, FS_OUTW * 4); bulk_commit(); }
	// ---- (5) rank(a, P) for the records that start in this slice ------------------------------------------------
	{
		for (uint32_t k0 = 0; k0 < nr; k0 += 32) { // (uniform trip count: the shuffle below needs every lane)
			const uint32_t k = k0 + lane;
			const bool valid = k < nr, reg = pf.have && k < 32;
			const uint32_t dst = !valid ? NONE32 : (reg ? pf.dst : in.dst[k]);
			const uint32_t a = !valid ? 0u : (reg ? (pf.sc & 7u) : (GENERAL ? (in.SC(k) & 7u) : (uint32_t)in.asym[k]));
			const int64_t base = __shfl_sync(FULLMASK, baseLane, (int)a);
			int64_t *rp = SHARDED && A.route ? (int64_t*)__shfl_sync(FULLMASK, (long long)routeLane, (int)a) : A.gLNext;
			const bool ids = SHARDED && !GENERAL && A.sidCur != 0; // (uniform)
			uint32_t *rp32 = ids ? (uint32_t*)__shfl_sync(FULLMASK, (long long)route32Lane, (int)a) : (uint32_t*)0;
			if (dst == NONE32) continue;
			const uint32_t xo = (uint32_t)((uint64_t)(reg ? pf.P : in.P[k]) - a0); // old symbols of the window in front of the record
			const uint32_t c = xo / FT_CH;
			const Raw6 rr = raw_unpack16(W.pre[c][0], W.pre[c][1], W.pre[c][2]);
			const uint32_t part = __popc(cell_match(cell_load(in.old + c * 3), a) & low_mask(xo & (FT_CH - 1)));
			int64_t g = base + raw_symbol(rr, a) + part;
			if (SHARDED && !oneBkt) { // (a bucket boundary inside the slice: look the record's own bucket up)
				const uint32_t b = (uint32_t)bucket_of(A.ctl->recBkt, (uint32_t)A.nb, r0 + k);
				g += A.recOff[b * 7 + a];
				if (A.route) { const uint32_t pc = A.route->pieceOf[a * 36 + b]; rp = A.route->base[pc]; if (ids) rp32 = A.route->base32[pc]; }
			}
			if (ids) rp32[dst] = reg ? pf.sid : A.sidCur[r0 + k]; // the string itself moves to the owner of its next sub-bucket
			rp[dst] = g; // (direct delivery: straight into the next owner's state array, a peer store over NVLink)
		}
	}
	__syncwarp(); // the inputs and W.mask / W.pre may be reused
}

// main kernel: persistent warps, each its own producer for the old symbols (the bulk of the bytes); the slice's
// records are read straight from the record arrays (coalesced: consecutive records, consecutive lanes)
template <bool GENERAL, bool SHARDED>
__global__ void __launch_bounds__(FS_WARPS * 32, FS_MINCTA) k_flat_merge(FlatArgs A)
{
	RB2_DYN_SMEM(smraw);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	SliceWarpSmem &S = reinterpret_cast<SliceWarpSmem*>(smraw)[wid];
	const uint32_t nWarps = gridDim.x * FS_WARPS;
	if (lane == 0) { for (int s = 0; s < FS_STAGES; ++s) mbar_init(&S.full[s], 1); }
	__syncwarp();
	// fetch the old symbols of slice `sl` into stage s (lane 0); a slice index behind the last one ends the ring
	auto issue = [&](uint32_t s, uint32_t sl, const TileDesc d0, const TileDesc d1) {
		if (lane != 0) return;
		SliceStage &st = S.st[s];
		if (sl >= A.nSlices) { st.slice = NONE32; mbar_arrive(&S.full[s]); return; }
		const uint64_t a0 = d0.i0 & ~(uint64_t)(FT_DIR - 1);
		const uint32_t bytes = (((uint32_t)(d1.i0 - a0) / FT_CH + 2) * 12u + 15u) & ~15u;
		st.d0 = d0; st.d1 = d1; st.slice = sl;
		mbar_expect_tx(&S.full[s], bytes);
		bulk_g2s(st.old, A.oldS + (a0 / FT_CH) * 12, bytes, &S.full[s]);
	};
	const uint32_t first = blockIdx.x * FS_WARPS + wid;
	const TileDesc none = { 0, 0, 0 };
	for (uint32_t s = 0; s < FS_STAGES; ++s) {
		const uint32_t sl = first + s * nWarps;
		const bool ok = lane == 0 && sl < A.nSlices;
		issue(s, sl, ok ? A.desc[sl] : none, ok ? A.desc[sl + 1] : none);
	}
	__syncwarp();
	RecRegs pf = { 0, 0, 0, 0, 0, false }, pfNext = { 0, 0, 0, 0, 0, false };
	const int64_t cpostLane = lane < 8 ? A.ctl->cpost[lane] : 0; // start of bucket `lane` behind this column
	for (uint32_t n = 0; ; ++n) {
		const uint32_t s = n % FS_STAGES, ph = (n / FS_STAGES) & 1u;
		mbar_wait(&S.full[s], ph);
		const SliceStage &st = S.st[s];
		const uint32_t slice = st.slice;
		if (slice == NONE32) break;
		{ // records k = lane of the slice in the other stage (merged next; its geometry was stored when it was issued)
			const SliceStage &sn = S.st[s ^ 1];
			pfNext.have = false;
			if (FS_STAGES == 2 && sn.slice != NONE32) {
				const uint32_t q0 = sn.d0.r0, qn = sn.d1.r0 - q0;
				pfNext.have = true;
				if ((uint32_t)lane < qn) {
					const uint32_t r = q0 + lane;
					pfNext.P = A.V.P[r]; pfNext.dst = A.recDst[r];
					pfNext.pre = GENERAL ? A.V.pre[r] : r;
					pfNext.sc = GENERAL ? A.V.sc[r] : (uint32_t)A.V.asym[r];
					if (SHARDED && !GENERAL && A.sidCur) pfNext.sid = A.sidCur[r];
				}
			}
		}
		// geometry of the slice this stage gets next: loaded now, used behind the merge (the latency hides under it)
		const uint32_t nextSl = slice + FS_STAGES * nWarps;
		const bool ok = lane == 0 && nextSl < A.nSlices;
		const TileDesc nd0 = ok ? A.desc[nextSl] : none, nd1 = ok ? A.desc[nextSl + 1] : none;
		const uint32_t r0 = st.d0.r0;
		SliceIn in = { st.old, A.V.P + r0, GENERAL ? A.V.pre + r0 : (const uint32_t*)0, GENERAL ? A.V.sc + r0 : (const uint32_t*)0, A.recDst + r0,
		               GENERAL ? (const uint8_t*)0 : A.V.asym + r0, r0 };
		flat_merge_slice<GENERAL, SHARDED>(A, S.W, in, slice, st.d0, st.d1, lane, pf, cpostLane);
		pf = pfNext;
		issue(s, nextSl, nd0, nd1); // (behind the slice's closing __syncwarp: every lane is done with the stage)
		__syncwarp();               // the stage's new geometry is visible to every lane (they prefetch its records next time round)
	}
	if (lane == 0) bulk_wait_read();
}


#undef FS_SLICE
#undef FS_CELLS
#undef FS_TILES
#undef FS_LPT
#undef FS_PCL
#undef FS_OLDC
#undef FS_OLDW
#undef FS_OUTW
#undef FS_MINCTA
#undef FS_CPL
