// rb2_flat_host.inl -- host side of the dense regime (rb2_flat.cuh); included by rb2_engine.cu.

// RB2_FLAT=0 never / 1 always / unset: by cost estimate
static int flat_pref(void)
{
	const char *s = getenv("RB2_FLAT");
	if (!s || !*s) return -1;
	return *s != '0';
}

// local symbol count of the index this engine holds, and (sharded) which buckets it owns
static bool bucket_mine(const rb2_engine *e, int b) { return !e->comm || e->owner[b] == e->rank; }
static int64_t bucket_len(const rb2_engine *e, int b)
{
	if (!e->comm) return e->bktLen[b];
	int64_t t = 0;
	for (int a = 0; a < 6; ++a) t += e->gtot[b][a];
	return t;
}
static uint64_t local_symbols(const rb2_engine *e)
{
	uint64_t n = 0;
	for (int b = 0; b < e->nb; ++b) if (bucket_mine(e, b)) n += (uint64_t)bucket_len(e, b);
	return n;
}

// Dense or sparse?  One dense column streams the whole local array (3 bits per symbol, read + written: ~0.34 ps per
// symbol at the measured merge rate); one sparse column costs ~1 ns per record plus a directory rebuild that is
// itself proportional to the index (~0.12 ps per symbol: the flat directory is re-scanned every column).  Sparse
// therefore only wins when the index holds more than ~5000 symbols per record of the column (measured on
// configs 2 and 4, profiles/README.md).  `strings` = records per column at the start of the batch, `addLocal` =
// symbols this engine expects to receive.  RB2_FLAT_RATIO overrides the crossover.
static bool flat_choose(rb2_engine *e, uint64_t strings, uint64_t addLocal)
{
	const int pref = flat_pref();
	if (pref == 0) return false;
	const uint64_t n0 = local_symbols(e), cap = n0 + addLocal + FT_PAD;
	size_t freeB = 0, totB = 0;
	RB2_CUDA(cudaMemGetInfo(&freeB, &totB));
	const uint64_t have = e->flat.s[0].cap + e->flat.s[1].cap + (e->flat.dir[0].cap + e->flat.dir[1].cap) * 8;
	const uint64_t need = 2 * flat_bytes(cap) + cap / 16 + (64ull << 20); // two arrays, their directories and tile tables
	if (need > freeB + have) return false; // does not fit: block-wise updates need far less
	if (pref == 1) return true;
	static double ratio = -1;
	if (ratio < 0) { const char *s = getenv("RB2_FLAT_RATIO"); ratio = s && *s ? atof(s) : 6000.0; }
	return strings >= 65536 && (double)n0 + 0.5 * (double)addLocal < ratio * (double)strings;
}

static void flat_scan_dir(rb2_engine *e, int which, uint64_t n)
{
	FlatState &f = e->flat;
	const uint64_t nd = n ? (n + FT_DIR - 1) / FT_DIR : 1;
	FlatDirScan fs = { f.tileCnt.p, nd, f.dir[which].p };
	run_scan<6, int64_t, FlatDirScan>(e, fs, nd, e->scanCta64, (int64_t*)0, e->midTmp64);
}

// grow a buffer that holds live data (the current array / directory): allocate, copy, free
template <typename T> static void grow_keep(rb2_engine *e, DevBuf<T> &b, size_t n, size_t live)
{
	if (n <= b.cap) return;
	T *np; const size_t cap = n + n / 16 + 64;
	RB2_CUDA(cudaMalloc(&np, cap * sizeof(T)));
	if (b.p && live) RB2_CUDA(cudaMemcpyAsync(np, b.p, live * sizeof(T), cudaMemcpyDeviceToDevice, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	if (b.p) RB2_CUDA(cudaFree(b.p));
	b.p = np; b.cap = cap;
}

// Start of a dense batch: the flat array either is resident from the previous dense batch (f.valid) and only
// has to make room, or is built from the leaf blocks.
static void flat_begin(rb2_engine *e, uint64_t addLocal)
{
	FlatState &f = e->flat;
	ph_begin(e, PH_CONVERT);
	const uint64_t n0 = local_symbols(e);
	uint64_t cap = n0 + addLocal + FT_PAD;
	// Growing a multi-GB array costs an allocation and a copy: grow in steps of 1.5x while memory is plentiful, or
	// straight to RB2_RESERVE symbols when the caller knows the final size of the index (the driver knows its input)
	if (flat_bytes(cap) > f.s[f.cur].cap || flat_bytes(cap) > f.s[f.cur ^ 1].cap) {
		static uint64_t reserve = ~0ull;
		if (reserve == ~0ull) { const char *rs = getenv("RB2_RESERVE"); reserve = rs && *rs ? strtoull(rs, 0, 10) : 0; }
		size_t freeB = 0, totB = 0;
		RB2_CUDA(cudaMemGetInfo(&freeB, &totB));
		const uint64_t have = f.s[0].cap + f.s[1].cap;
		uint64_t want = cap + cap / 2;
		if (reserve + FT_PAD > cap) want = reserve + FT_PAD;
		if (2 * flat_bytes(want) + want / 16 + ((uint64_t)24 << 30) < freeB + have) cap = want;
	}
	if (f.valid) {
		if (f.n != n0) RB2_FATAL("internal: resident flat array holds %llu symbols, the index %llu", (unsigned long long)f.n, (unsigned long long)n0);
		// the other buffer holds nothing live: release it first so that the peak is old + new current array
		if (flat_bytes(cap) > f.s[f.cur ^ 1].cap) f.s[f.cur ^ 1].release();
		grow_keep(e, f.s[f.cur], flat_bytes(cap), flat_bytes(n0 + FT_PAD) < f.s[f.cur].cap ? flat_bytes(n0 + FT_PAD) : f.s[f.cur].cap);
		grow_keep(e, f.dir[f.cur], (cap / FT_DIR + 3) * 6, (n0 / FT_DIR + 2) * 6);
	} else {
		f.cur = 0;
		f.s[0].need(flat_bytes(cap)); f.dir[0].need((cap / FT_DIR + 3) * 6);
	}
	f.s[f.cur ^ 1].need(flat_bytes(cap)); f.dir[f.cur ^ 1].need((cap / FT_DIR + 3) * 6);
	f.tileCnt.need((cap / FT_DIR + 3) * 3);
	f.desc.need(cap / fs2::kSlice + 4);
	if (!f.valid) {
		f.n = n0;
		if (n0 > 0) {
			Dir &d = e->dir[e->cur];
			RB2_CUDA(cudaMemsetAsync(f.s[0].p, 0, flat_bytes(n0 + FT_CH), e->st));
			LAUNCH(e, k_blocks_to_flat, cdiv(e->nlog, 4), 128, 0, e->pool, d.order, d.cumLen, e->nlog, e->comm ? e->dDirOff : (const int64_t*)0,
			       e->dctl->blkBkt, e->nb, f.s[0].p, e->dctl);
			LAUNCH(e, k_flat_count_tiles, cdiv((n0 + FT_DIR - 1) / FT_DIR, 8), 256, 0, f.s[0].p, n0, f.tileCnt.p);
		} else RB2_CUDA(cudaMemsetAsync(f.tileCnt.p, 0, 12, e->st));
		flat_scan_dir(e, 0, n0);
	}
	ph_end(e, PH_CONVERT);
	f.pending |= 1u << PH_CONVERT;
	f.on = true; f.valid = true; f.blocksStale = true;
}

// one column: merge nrec records (inserting `inserted` symbols) into the flat array
static void flat_apply_records(rb2_engine *e, uint32_t nrec, uint64_t inserted, int64_t *gLNext, const int64_t *leanP = 0, const uint8_t *asym = 0,
                               const PeerRoute *route = 0, const uint32_t *sidCur = 0)
{
	FlatState &f = e->flat;
	const uint64_t nNew = f.n + inserted;
	// slice size: 2048 symbols per warp while the column inserts many symbols per slice, 4096 once the array is large
	// against the batch (RB2_WIDE_RATIO symbols per record; measured crossover, profiles/README.md)
	static double wideRatio = -1;
	if (wideRatio < 0) { const char *ws = getenv("RB2_WIDE_RATIO"); wideRatio = ws && *ws ? atof(ws) : 192.0; }
	const bool wide = (double)nNew > wideRatio * (double)(nrec ? nrec : 1);
	const uint32_t slice = wide ? fs4::kSlice : fs2::kSlice;
	const uint64_t nTiles = (nNew + slice - 1) / slice; // slices: one warp each
	// the target buffers hold nothing live: grow them if this rank receives more than was estimated
	f.s[f.cur ^ 1].need(flat_bytes(nNew + FT_PAD)); f.dir[f.cur ^ 1].need((nNew / FT_DIR + 3) * 6);
	f.tileCnt.need((nNew / FT_DIR + 3) * 3); f.desc.need(nNew / fs2::kSlice + 4);
	ph_begin(e, PH_MERGE);
	// leanP: all-singleton column -- the records are the state arrays themselves (position = leanP[r], symbol = asym[r], count 1, r symbols in front)
	const RecView V = leanP ? RecView{ leanP, 0, 0, asym ? asym : e->asym.p } : RecView{ e->recP.p, e->recPre.p, e->recSC.p, 0 };
	if (e->comm) f.sliceBkt.need(nNew / fs2::kSlice + 4);
	LAUNCH(e, k_flat_geo, cdiv(nTiles + 1, 256), 256, 0, V, nrec, nTiles, nNew, slice, f.desc.p, e->dctl, e->nb, e->comm ? f.sliceBkt.p : (uint8_t*)0);
	if (nTiles >= 0xfffffff0ull) RB2_FATAL("flat array of %llu symbols: more than 2^32 slices", (unsigned long long)nNew);
	FlatArgs fa = { f.s[f.cur].p, f.dir[f.cur].p, f.s[f.cur ^ 1].p, nNew, f.tileCnt.p, V, e->recDst.p, nrec,
	                f.desc.p, (uint32_t)nTiles, gLNext, e->dctl, e->comm ? e->dDirOffPre : (const int64_t*)0, e->nb, route, e->comm ? f.sliceBkt.p : (const uint8_t*)0, sidCur };
	const uint32_t grid = (uint32_t)std::min<uint64_t>(cdiv(nTiles, FS_WARPS), (uint64_t)e->nSM * (wide ? fs4::kMinCta : fs2::kMinCta));
#define RB2_MERGE_LAUNCH(NS, GEN, SH) LAUNCH(e, (NS::k_flat_merge<GEN, SH>), grid, FS_WARPS * 32, FS_WARPS * sizeof(NS::SliceWarpSmem), fa)
	if (e->comm) { // (a rank of a sharded build: whole-index coordinates, direct delivery)
		if (wide) { if (V.sc) RB2_MERGE_LAUNCH(fs4, true, true); else RB2_MERGE_LAUNCH(fs4, false, true); }
		else { if (V.sc) RB2_MERGE_LAUNCH(fs2, true, true); else RB2_MERGE_LAUNCH(fs2, false, true); }
	} else {
		if (wide) { if (V.sc) RB2_MERGE_LAUNCH(fs4, true, false); else RB2_MERGE_LAUNCH(fs4, false, false); }
		else { if (V.sc) RB2_MERGE_LAUNCH(fs2, true, false); else RB2_MERGE_LAUNCH(fs2, false, false); }
	}
#undef RB2_MERGE_LAUNCH
	ph_end(e, PH_MERGE);
	ph_begin(e, PH_DIR);
	flat_scan_dir(e, f.cur ^ 1, nNew);
	ph_end(e, PH_DIR);
	++e->stats.n_merge_launches;
	e->stats.merge_blocks += nTiles;
	e->stats.merge_bytes_rw += (int64_t)(flat_bytes(f.n) + flat_bytes(nNew)) + (int64_t)nrec * ((leanP ? 21 : 28) + (sidCur ? 8 : 0)); // old array read, new written (3 bits per symbol), records (13 / 20 B) read, ranks (8 B) written (+ ids read and delivered)
	f.cur ^= 1; f.n = nNew;
	f.pending |= (1u << PH_MERGE) | (1u << PH_DIR);
	if (getenv("RB2_FLAT_DEBUG") && nNew <= 4096 && !leanP && gLNext) { // developer aid: dump tiny arrays column by column
		RB2_CUDA(cudaStreamSynchronize(e->st));
		std::vector<uint8_t> hs(nNew); std::vector<int64_t> hp(nrec), hd(12); std::vector<uint32_t> hpre(nrec + 1), hsc(nrec), hdst(nrec);
		{ std::vector<uint8_t> pk(flat_bytes(nNew)); RB2_CUDA(cudaMemcpy(pk.data(), f.s[f.cur].p, pk.size(), cudaMemcpyDeviceToHost));
		  for (uint64_t i = 0; i < nNew; ++i) hs[i] = (uint8_t)flat_get(pk.data(), i); }
		RB2_CUDA(cudaMemcpy(hp.data(), e->recP.p, nrec * 8, cudaMemcpyDeviceToHost));
		RB2_CUDA(cudaMemcpy(hpre.data(), e->recPre.p, (nrec + 1) * 4, cudaMemcpyDeviceToHost));
		RB2_CUDA(cudaMemcpy(hsc.data(), e->recSC.p, nrec * 4, cudaMemcpyDeviceToHost));
		RB2_CUDA(cudaMemcpy(hdst.data(), e->recDst.p, nrec * 4, cudaMemcpyDeviceToHost));
		RB2_CUDA(cudaMemcpy(hd.data(), f.dir[f.cur].p, 96, cudaMemcpyDeviceToHost));
		fprintf(stderr, "[flat] n=%llu records:", (unsigned long long)nNew);
		for (uint32_t r = 0; r < nrec && r < 40; ++r) {
			int64_t g = -1;
			if (hdst[r] != NONE32) RB2_CUDA(cudaMemcpy(&g, gLNext + hdst[r], 8, cudaMemcpyDeviceToHost));
			fprintf(stderr, " (P%lld pre%u %c x%u ->%lld)", (long long)hp[r], hpre[r], "$ACGTN"[hsc[r] & 7], hsc[r] >> 3, (long long)g);
		}
		fprintf(stderr, " pre[R]=%u\n[flat] array: ", hpre[nrec]);
		for (uint64_t i = 0; i < nNew && i < 200; ++i) fputc("$ACGTN??"[hs[i] & 7], stderr);
		fprintf(stderr, "\n[flat] dir row1:");
		for (int a = 0; a < 6; ++a) fprintf(stderr, " %lld", (long long)hd[6 + a]);
		fprintf(stderr, "\n");
	}
}

// End of a dense batch: the array stays resident (the leaf blocks are now stale); refresh the host
// mirrors of the per-bucket symbol totals from the array's directory.
static void flat_finish(rb2_engine *e)
{
	FlatState &f = e->flat;
	if (f.pending) { RB2_CUDA(cudaStreamSynchronize(e->st)); ph_collect(e, f.pending); f.pending = 0; }
	f.on = false;
	if (e->comm) return; // sharded engines keep their totals in gtot
	int64_t hpos[8], hout[8 * 6];
	int64_t acc = 0;
	for (int b = 0; b <= 6; ++b) { hpos[b] = acc; if (b < 6) acc += e->bktLen[b]; }
	if ((uint64_t)acc != f.n) RB2_FATAL("internal: flat array holds %llu symbols, the buckets %lld", (unsigned long long)f.n, (long long)acc);
	e->stageCnt.need(8 + 8 * 6);
	RankAtPos rp; for (int b = 0; b < 8; ++b) rp.pos[b] = b <= 6 ? hpos[b] : 0;
	LAUNCH(e, k_flat_rank_at, 7, 32, 0, f.s[f.cur].p, f.dir[f.cur].p, rp, e->stageCnt.p + 8); // (positions as a kernel argument: no host-to-device copy on this path)
	RB2_CUDA(cudaMemcpyAsync(hout, e->stageCnt.p + 8, 7 * 6 * 8, cudaMemcpyDeviceToHost, e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	for (int b = 0; b < 6; ++b) for (int a = 0; a < 6; ++a) e->tot[b][a] = hout[(b + 1) * 6 + a] - hout[b * 6 + a];
}

// batch scratch that can be rebuilt at any time (released when the conversion below is short of memory)
static void release_batch_scratch(rb2_engine *e)
{
	RB2_CUDA(cudaStreamSynchronize(e->st));
	e->sbuf.release(); e->T.release(); e->asym.release(); e->sizes6.release(); e->recP.release(); e->recSC.release(); e->recDst.release(); e->recPre.release();
	for (int k = 0; k < 2; ++k) { e->gL[k].release(); e->gSize[k].release(); e->gOff[k].release(); e->sid[k].release(); }
	if (e->comm && !e->p2pMapped) { e->gLrx[0].release(); e->gLrx[1].release(); e->sidrx[0].release(); e->sidrx[1].release(); } // (not while peers map them)
	e->strEnd.release(); e->tileA.release(); e->tileB.release(); e->grpCta.release();
	FlatState &f = e->flat;
	f.s[f.cur ^ 1].release(); f.dir[f.cur ^ 1].release(); f.desc.release();
}

// flat array -> leaf blocks: buckets are encoded independently
static void flat_to_blocks(rb2_engine *e)
{
	FlatState &f = e->flat;
	if (f.pending) { RB2_CUDA(cudaStreamSynchronize(e->st)); ph_collect(e, f.pending); f.pending = 0; }
	ph_begin(e, PH_CONVERT);
	EncTab T; memset(&T, 0, sizeof(T));
	T.nb = e->nb;
	uint64_t sym = 0, ch = 0;
	for (int b = 0; b < e->nb; ++b) {
		T.symStart[b] = sym; T.chunkStart[b] = ch;
		if (bucket_mine(e, b)) { const uint64_t l = (uint64_t)bucket_len(e, b); sym += l; ch += (l + FE_CHUNK - 1) / FE_CHUNK; }
	}
	T.symStart[e->nb] = sym; T.chunkStart[e->nb] = ch;
	if (sym != f.n) RB2_FATAL("internal: flat array holds %llu symbols, the buckets %llu", (unsigned long long)f.n, (unsigned long long)sym);
	const uint64_t nChunk = ch;
	{ // the conversion needs ~9 bytes per chunk plus the pool: make room if the batch scratch is in the way
		size_t freeB = 0, totB = 0;
		RB2_CUDA(cudaMemGetInfo(&freeB, &totB));
		if ((nChunk + 2) * 9 + f.n + (1ull << 30) > freeB) release_batch_scratch(e);
	}
	f.chunkBytes.need(nChunk + 1); f.chunkPre.need(nChunk + 2);
	std::vector<uint64_t> edge(2 * (size_t)e->nb, 0);
	if (nChunk) {
		LAUNCH(e, k_flat_chunk_bytes, cdiv(nChunk, 256), 256, 0, f.s[f.cur].p, T, nChunk, f.chunkBytes.p);
		ChunkScan cs = { f.chunkBytes.p, nChunk, f.chunkPre.p };
		run_scan<1, uint64_t, ChunkScan>(e, cs, nChunk, f.scanU64, (uint64_t*)0, f.midU64);
		for (int b = 0; b < e->nb; ++b) if (T.chunkStart[b + 1] > T.chunkStart[b]) {
			RB2_CUDA(cudaMemcpyAsync(&edge[2 * b], f.chunkPre.p + T.chunkStart[b], 8, cudaMemcpyDeviceToHost, e->st));
			RB2_CUDA(cudaMemcpyAsync(&edge[2 * b + 1], f.chunkPre.p + T.chunkStart[b + 1] - 1, 8, cudaMemcpyDeviceToHost, e->st));
		}
		RB2_CUDA(cudaStreamSynchronize(e->st));
	}
	f.chunkBytes.release();
	uint32_t nBlocks = 0;
	for (int b = 0; b < e->nb; ++b) {
		T.blkStart[b] = nBlocks; T.byteStart[b] = edge[2 * b];
		if (T.chunkStart[b + 1] > T.chunkStart[b]) nBlocks += (uint32_t)((edge[2 * b + 1] - edge[2 * b]) / FE_T) + 1;
		else if (bucket_mine(e, b)) nBlocks += 1; // an empty bucket keeps one empty block
	}
	T.blkStart[e->nb] = nBlocks;
	e->hctl->poolUsed = 0; // the pool is rewritten from scratch: nothing to carry over when it grows
	reserve_blocks(e, (uint64_t)nBlocks + nBlocks / 16 + 4096);
	LAUNCH(e, k_flat_encode, cdiv(nBlocks, 16), 128, 0, f.s[f.cur].p, T, f.chunkPre.p, nBlocks, e->pool, e->blkCnt);
	LAUNCH(e, k_fill_u32, cdiv(nBlocks, 256), 256, 0, e->dir[e->cur].order, nBlocks, 0u, 1u);
	e->nlog = nBlocks;
	for (int b = 0; b < NBA; ++b) e->blkBkt[b] = b <= e->nb ? T.blkStart[b] : nBlocks;
	Ctl *h = e->hctl;
	h->poolUsed = nBlocks; h->poolCap = e->poolCap; h->nb = (uint32_t)e->nb;
	for (int b = 0; b < NBA; ++b) h->blkBkt[b] = e->blkBkt[b];
	ctl_push(e);
	rebuild_directory(e, false);
	ph_end(e, PH_CONVERT);
	RB2_CUDA(cudaStreamSynchronize(e->st));
	ph_collect(e, 1u << PH_CONVERT);
	f.chunkPre.release();
	f.blocksStale = false;
	e->stats.pool_blocks = e->hctl->poolUsed;
	e->stats.pool_capacity = e->poolCap;
}

// Everything that reads or edits leaf blocks (iterator, dump, rank queries, rope_insert_run, a sparse
// batch, loading blocks) calls this first: after dense batches the blocks are rebuilt from the array.
static void ensure_blocks(rb2_engine *e)
{
	if (e->flat.valid && e->flat.blocksStale) flat_to_blocks(e);
}
// ... and whatever edits the blocks invalidates the resident array
static void blocks_edited(rb2_engine *e) { e->flat.valid = false; e->flat.blocksStale = false; }
