// rb2_shard.inl -- the sharded (multi-GPU) BCR build; included by rb2_engine.cu.
//
// Partition.  The reference runs the six buckets of one column on five threads (mrope.c:312-325):
// bucket b holds the BWT symbols whose suffix starts with symbol b.  Here every bucket is cut once
// more by the SECOND symbol of the suffix: sub-bucket (x,y) = the BWT symbols whose suffix starts
// with "xy" -- a contiguous rank interval of bucket x (its "interior rank interval").  The 36
// sub-buckets are separate block sequences, statically assigned to ranks as contiguous ranges in
// (x,y) order.  A string that sits in (x,y) and inserts symbol a continues in (a,x) -- known from
// the string's own last two symbols, so no position lookup is needed to route it, and every rank
// derives the complete transfer plan of a column from one small all-gather of counts.
//
// One column on rank r:
//   local   : next symbols, groups -> records, 6-way partition (same kernels as the one-GPU engine,
//             36 buckets instead of 6)
//   gather  : grpPre / memPre tables of all ranks                        (Comm::allgather_host)
//             -> post-column symbol totals of all 36 sub-buckets = the cross-bucket offsets of
//                mrope.c:332-340, -> the directory offsets of my blocks, -> the transfer plan
//   local   : merge the records into my leaf blocks; ranks come out in whole-index coordinates
//   exchange: interval start / size, member ranges, member ids move to the owner of the next
//             sub-bucket                                                  (Comm::exchange)

// owner of each sub-bucket: contiguous ranges of the 16 sub-buckets that hold ACGT x ACGT (the ones
// that carry the data); the small ones ($ and N related) go with their neighbours
static void shard_owner_map(int nranks, int owner[NBMAX])
{
	int k = 0; // main sub-buckets in front of s
	for (int s = 0; s < NBMAX; ++s) {
		const int x = s / 6, y = s % 6;
		const bool main_ = x >= 1 && x <= 4 && y >= 1 && y <= 4;
		int o = (int)((int64_t)k * nranks / 16);
		if (o > nranks - 1) o = nranks - 1;
		owner[s] = o;
		if (main_) ++k;
	}
}

extern "C" int rb2_shard_owner(int nranks, int subbucket)
{
	if (nranks < 1 || nranks > RB2_MAX_RANKS || subbucket < 0 || subbucket >= NBMAX) return -1;
	int owner[NBMAX];
	shard_owner_map(nranks, owner);
	return owner[subbucket];
}

extern "C" rb2_group_t *rb2_group_create(int nranks)
{
	if (nranks < 1 || nranks > RB2_MAX_RANKS) RB2_FATAL("rb2_group_create: 1..%d ranks", RB2_MAX_RANKS);
	rb2_group *g = new rb2_group();
	g->n = nranks; g->arrived.store(0); g->phase.store(0); g->joined.store(0);
	// peer access lets the pulls of LocalComm run over NVLink; same-device "virtual ranks" need nothing
	int nd = rb2_device_count();
	for (int a = 0; a < nd; ++a) for (int b = 0; b < nd; ++b) if (a != b) {
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) { cudaSetDevice(a); cudaDeviceEnablePeerAccess(b, 0); }
	}
	cudaGetLastError();
	return g;
}
extern "C" void rb2_group_destroy(rb2_group_t *g) { delete g; }

extern "C" void rb2_nccl_unique_id(uint8_t out[128])
{
	ncclUniqueId id;
	RB2_NCCL(nccl_api()->GetUniqueId(&id));
	memcpy(out, &id, sizeof(id) < 128 ? sizeof(id) : 128);
}

// symbols / per-symbol counts of the index that other ranks hold in front of each of my sub-buckets
static void shard_dir_offsets(rb2_engine *e, const int64_t tot[NBMAX][6], bool pre = false)
{
	int64_t *hOff = pre ? e->hDirOffPre : e->hDirOff, *dOff = pre ? e->dDirOffPre : e->dDirOff;
	int64_t acc[7] = { 0, 0, 0, 0, 0, 0, 0 };
	for (int s = 0; s < NBMAX; ++s) {
		for (int a = 0; a < 7; ++a) hOff[s * 7 + a] = acc[a];
		if (e->owner[s] != e->rank)
			for (int a = 0; a < 6; ++a) { acc[a] += tot[s][a]; acc[6] += tot[s][a]; }
	}
	RB2_CUDA(cudaMemcpyAsync(dOff, hOff, NBMAX * 7 * sizeof(int64_t), cudaMemcpyHostToDevice, e->st));
}

// dense regime: record positions are whole-index coordinates; my flat array starts at my first symbol
__global__ void k_flat_localize(int64_t *recP, uint32_t R, const Ctl *ctl, const int64_t *off, int nb)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < R) recP[r] -= off[bucket_of(ctl->recBkt, (uint32_t)nb, r) * 7 + 6];
}

static void shard_publish_totals(rb2_engine *e)
{
	for (int x = 0; x < 6; ++x) {
		e->bktLen[x] = 0;
		for (int a = 0; a < 6; ++a) {
			e->tot[x][a] = 0;
			for (int y = 0; y < 6; ++y) e->tot[x][a] += e->gtot[x * 6 + y][a];
			e->bktLen[x] += e->tot[x][a];
		}
	}
}

// empty sharded index: every sub-bucket I own consists of one empty leaf block
static void shard_reset_index(rb2_engine *e)
{
	std::vector<uint32_t> ord;
	for (int s = 0; s <= NBMAX; ++s) {
		e->blkBkt[s] = (uint32_t)ord.size();
		if (s < NBMAX && e->owner[s] == e->rank) ord.push_back((uint32_t)ord.size());
	}
	for (int s = NBMAX + 1; s < NBA; ++s) e->blkBkt[s] = (uint32_t)ord.size();
	e->nlog = (uint32_t)ord.size();
	if (e->nlog == 0) RB2_FATAL("rank %d owns no sub-bucket (more ranks than the partition supports)", e->rank);
	RB2_CUDA(cudaMemsetAsync(e->pool, 0, (size_t)e->nlog * RB2_BLK, e->st));
	RB2_CUDA(cudaMemsetAsync(e->blkCnt, 0, (size_t)e->nlog * 24, e->st));
	LAUNCH(e, k_fill_u32, 1, 64, 0, e->dir[e->cur].order, e->nlog, 0u, 1u);
	memset(e->gtot, 0, sizeof(e->gtot));
	Ctl *h = e->hctl;
	h->nb = NBMAX; h->tables = 1; h->poolUsed = e->nlog; h->poolCap = e->poolCap; h->err = 0;
	for (int b = 0; b < NBA; ++b) h->blkBkt[b] = e->blkBkt[b];
	ctl_push(e);
	shard_dir_offsets(e, e->gtot);
	rebuild_directory(e, false);
	RB2_CUDA(cudaStreamSynchronize(e->st));
	shard_publish_totals(e);
	e->stats.pool_blocks = e->nlog;
}

extern "C" rb2_engine_t *rb2_create_sharded(int device, int sorting_order, int rank, int nranks, rb2_group_t *group, const uint8_t *nccl_uid)
{
	if (nranks < 1 || nranks > RB2_MAX_RANKS || rank < 0 || rank >= nranks) RB2_FATAL("rb2_create_sharded: rank %d of %d", rank, nranks);
	if ((group != 0) == (nccl_uid != 0)) RB2_FATAL("rb2_create_sharded: pass either an in-process group or an NCCL unique id");
	rb2_engine *e = rb2_create(device, sorting_order);
	e->nb = NBMAX; e->rank = rank; e->nranks = nranks;
	shard_owner_map(nranks, e->owner);
	RB2_CUDA(cudaMalloc(&e->dDirOff, NBMAX * 7 * sizeof(int64_t)));
	RB2_CUDA(cudaMallocHost(&e->hDirOff, NBMAX * 7 * sizeof(int64_t)));
	{ // the overlapped part of the exchange must get SM / copy slots while the merge grid is still draining
		int lo = 0, hi = 0;
		RB2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		RB2_CUDA(cudaStreamCreateWithPriority(&e->st2, cudaStreamNonBlocking, hi));
	}
	RB2_CUDA(cudaEventCreateWithFlags(&e->evEarly, cudaEventDisableTiming));
	RB2_CUDA(cudaEventCreateWithFlags(&e->evMerge, cudaEventDisableTiming));
	RB2_CUDA(cudaMalloc(&e->dDirOffPre, NBMAX * 7 * sizeof(int64_t)));
	RB2_CUDA(cudaMallocHost(&e->hDirOffPre, NBMAX * 7 * sizeof(int64_t)));
	RB2_CUDA(cudaMallocHost(&e->hPlan, 2 * 2 * (NBMAX * 6 + 8) * sizeof(uint32_t))); // two columns in flight
	if (group) { if (group->n != nranks) RB2_FATAL("group size mismatch"); e->comm = new LocalComm(group, rank); }
	else e->comm = new NcclComm(rank, nranks, nccl_uid);
	shard_reset_index(e);
	return e;
}

extern "C" int rb2_num_buckets(const rb2_engine_t *e) { return e->nb; }

__global__ void k_rebase_goff(uint32_t *gOff, uint32_t G, uint32_t M, const uint32_t *pieceStart, const uint32_t *pieceDelta, int np)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g == 0) gOff[G] = M;
	if (g >= G) return;
	int lo = 0, hi = np - 1; // last piece that starts at or in front of g
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pieceStart[mid] <= g) lo = mid; else hi = mid - 1; }
	gOff[g] += pieceDelta[lo];
}

struct ShardTab { uint32_t grp[(NBMAX + 1) * 6], mem[(NBMAX + 1) * 6]; };

// Collective (every rank holds mappings or none does): close my mappings of the peers' state buffers.  Behind the
// barrier nobody maps anybody's buffers: they may be freed or grown.
static void shard_unmap_peers(rb2_engine *e)
{
	if (!e->p2pMapped) return;
	RB2_CUDA(cudaStreamSynchronize(e->st));
	for (int k = 0; k < 2; ++k) { e->comm->p2p_unmap((void**)e->peerGL[k]); e->comm->p2p_unmap((void**)e->peerSid[k]); }
	e->comm->barrier(e->st);
	e->p2pMapped = false;
}

extern "C" void rb2_sharded_quiesce(rb2_engine_t *e)
{
	if (!e->comm) RB2_FATAL("rb2_sharded_quiesce: engine was not created with rb2_create_sharded");
	RB2_CUDA(cudaSetDevice(e->dev));
	shard_unmap_peers(e);
}

static void insert_sharded_range(rb2_engine *e, const uint8_t *s, uint32_t kBase, uint32_t m);

// One batch: every rank passes ITS strings (device resident, NUL-terminated, reversed; len may be 0).
// Collective: all ranks of the communicator must call it.
static void insert_sharded_batch(rb2_engine *e, int64_t len, const uint8_t *s)
{
	uint32_t m = 0;
	if (len > 0) { // my strings: where each one ends
		const uint32_t nT = cdiv(len, 4096);
		e->tileA.need((size_t)nT + 2);
		LAUNCH(e, k_count_nul, nT, 256, 0, s, len, e->tileA.p);
		uint32_t *dTot = e->tileA.p + nT;
		run_mid<1, uint32_t>(e, e->tileA.p, (uint64_t)nT, dTot, e->midTmp);
		RB2_CUDA(cudaMemcpyAsync(&m, dTot, 4, cudaMemcpyDeviceToHost, e->st));
		RB2_CUDA(cudaStreamSynchronize(e->st));
		if (m == 0) RB2_FATAL("batch holds no terminated string");
		e->strEnd.need(m);
		LAUNCH(e, k_string_ends, nT, 256, 0, s, len, e->tileA.p, e->strEnd.p);
	}
	insert_sharded_range(e, s, 0, m);
	e->stats.n_strings += m;
	e->stats.n_symbols += len;
}

// Strings kBase .. kBase+m-1 of my share (string ends in e->strEnd; m may be 0).  Like the one-GPU engine
// (insert_string_range) the ranks cut a batch in two when its symbol matrices -- dense rectangles of (longest
// string + 1) columns, replicated on every rank -- would not fit or would be mostly padding (one long string among
// short ones): the cut is made in the GLOBAL string order (rank, position), so input order is preserved, and every
// rank takes the same decision from the all-gathered shapes.
static void insert_sharded_range(rb2_engine *e, const uint8_t *s, uint32_t kBase, uint32_t m)
{
	Comm *cm = e->comm;
	const int P = cm->n, me = cm->rank;
	const int sorted = e->so != RB2_SO_IO;
	Ctl *h = e->hctl;
	// RB2_TRACE=1: host wall-clock of the batch's stages on stderr (developer aid)
	static const bool trace = getenv("RB2_TRACE") && atoi(getenv("RB2_TRACE"));
	double tr[8]; int ntr = 0;
	auto mark = [&]() { if (trace && ntr < 8) { RB2_CUDA(cudaStreamSynchronize(e->st)); timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); tr[ntr++] = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; } };
	mark();
	double trCtl = 0, trGather = 0; // host time blocked in the column loop: waiting for the column's counts / in the table all-gather
	auto now_ms = [&]() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };

	// ---- my strings: split, lengths ----------------------------------------------------------
	ph_begin(e, PH_TRANSPOSE);
	unsigned long long maxlen = 0;
	int64_t len = 0;
	if (m) {
		int64_t ends[2] = { -1, 0 };
		RB2_CUDA(cudaMemsetAsync(e->dMaxLen, 0, 8, e->st));
		LAUNCH(e, k_maxlen, cdiv(m, 256), 256, 0, e->strEnd.p, kBase, m, e->dMaxLen);
		RB2_CUDA(cudaMemcpyAsync(&maxlen, e->dMaxLen, 8, cudaMemcpyDeviceToHost, e->st));
		if (kBase) RB2_CUDA(cudaMemcpyAsync(&ends[0], e->strEnd.p + kBase - 1, 8, cudaMemcpyDeviceToHost, e->st));
		RB2_CUDA(cudaMemcpyAsync(&ends[1], e->strEnd.p + kBase + m - 1, 8, cudaMemcpyDeviceToHost, e->st));
		RB2_CUDA(cudaStreamSynchronize(e->st));
		len = ends[1] - ends[0];
	}
	// ---- shapes of all ranks; replicate the column-major symbol matrices --------------------------
	struct Shape { uint64_t m, ncol, len, freeB, tcap; } mine = { m, m ? maxlen + 1 : 0, (uint64_t)len, 0, e->T.cap }, all[RB2_MAX_RANKS];
	{ size_t freeB = 0, totB = 0; RB2_CUDA(cudaMemGetInfo(&freeB, &totB)); mine.freeB = freeB; }
	cm->allgather_host(&mine, sizeof(mine), all, e->st);
	uint64_t mAll = 0, lenAll = 0, tBytes = 0, ncolAll = 0;
	uint64_t strOff[RB2_MAX_RANKS + 1], tOff[RB2_MAX_RANKS + 1];
	for (int r = 0; r < P; ++r) {
		strOff[r] = mAll; tOff[r] = tBytes;
		mAll += all[r].m; lenAll += all[r].len;
		tBytes += t_stride(all[r].m) * all[r].ncol;
		ncolAll = std::max(ncolAll, all[r].ncol);
	}
	strOff[P] = mAll; tOff[P] = tBytes;
	if (mAll == 0) RB2_FATAL("mr_insert_multi: empty batch (mrope.c:268)");
	if (mAll >= 0xfffffff0ull) RB2_FATAL("a sharded batch is limited to 2^32 strings");
	{
		bool tooBig = false;
		for (int r = 0; r < P; ++r) tooBig = tooBig || (tBytes > all[r].tcap && tBytes > (all[r].freeB + all[r].tcap) / 2);
		const bool wasteful = tBytes > 8 * lenAll + split_slack(); // the matrices hold 4 bits per symbol: > 16x padding
		if ((tooBig || wasteful) && mAll > 1) {
			ph_end(e, PH_TRANSPOSE);
			const uint64_t cut = mAll / 2, lo = strOff[me];
			const uint32_t k1 = (uint32_t)(cut <= lo ? 0 : std::min<uint64_t>(cut - lo, m));
			insert_sharded_range(e, s, kBase, k1);
			insert_sharded_range(e, s, kBase + k1, m - k1);
			return;
		}
		if (tooBig) RB2_FATAL("one string of %llu symbols does not fit in HBM as a column-major symbol matrix on every rank", (unsigned long long)ncolAll);
	}
	e->T.need(tBytes + 16);
	if (m) LAUNCH(e, k_transpose, cdiv(m, TR_S), 256, 0, s, e->strEnd.p, kBase, m, (int64_t)all[me].ncol, e->T.p + tOff[me]);
	{
		uint8_t *dst[RB2_MAX_RANKS]; size_t bytes[RB2_MAX_RANKS];
		for (int r = 0; r < P; ++r) { dst[r] = e->T.p + tOff[r]; bytes[r] = t_stride(all[r].m) * all[r].ncol; }
		cm->gather_blocks(dst, bytes, e->st);
	}
	ph_end(e, PH_TRANSPOSE);

	// ---- state for column 0 (mrope.c:279-285): everything starts in sub-bucket ($,$) -------------
	const int64_t n0 = e->bktLen[0];
	const bool useSizes = sorted && n0 > 0;
	// dense or sparse regime: every rank must take the same path, so the ranks vote
	const uint64_t addLocal = lenAll / P + lenAll / (4 * P) + 4096;
	uint32_t myVote = flat_choose(e, mAll / P + 1, addLocal) ? 1u : 0u, votes[RB2_MAX_RANKS];
	cm->allgather_host(&myVote, 4, votes, e->st);
	bool flat = true;
	for (int r = 0; r < P; ++r) flat = flat && votes[r] != 0;
	if (flat) {
		if (!e->flat.valid) { // the array is built from the leaf blocks: their bucket table and my offsets must be current on the device
			for (int b = 0; b < NBA; ++b) h->blkBkt[b] = e->blkBkt[b];
			h->nb = NBMAX; h->tables = 1;
			ctl_push(e);
			shard_dir_offsets(e, e->gtot);
		}
		flat_begin(e, addLocal);
	} else { ensure_blocks(e); blocks_edited(e); } // a sparse batch edits the leaf blocks
	if (!flat) reserve_blocks(e, (uint64_t)h->poolUsed + (lenAll * 2 / RB2_FILL) / P * 5 / 4 + 4096);
	mark();
	// ---- direct delivery of the interval starts (dense regime): every rank maps every rank's two state buffers ----
	// The buffers are sized for the worst case (no rank ever holds more than every string of the batch), so the
	// mappings stay valid for the whole batch -- and for the batches after it, as long as no rank needs larger
	// buffers (mapping costs 30-90 ms, measured): the ranks vote, and re-map together when one of them must grow.
	bool direct = false;
	if (P > 1) {
		const char *ws = getenv("RB2_P2P");
		const int want = ws && *ws ? atoi(ws) : 1;
		const size_t capG = (size_t)mAll + 64;
		const bool fits = e->gLrx[0].cap >= capG && e->gLrx[1].cap >= capG && e->sidrx[0].cap >= capG && e->sidrx[1].cap >= capG;
		size_t freeB = 0, totB = 0;
		RB2_CUDA(cudaMemGetInfo(&freeB, &totB));
		const size_t extra = (e->gLrx[0].cap < capG ? capG * 9 : 0) + (e->gLrx[1].cap < capG ? capG * 9 : 0) +
		                     (e->sidrx[0].cap < capG ? capG * 5 : 0) + (e->sidrx[1].cap < capG ? capG * 5 : 0);
		uint32_t mine = (want && flat && extra + ((size_t)2 << 30) < freeB ? 1u : 0u) | (e->p2pMapped && fits ? 2u : 0u), got[RB2_MAX_RANKS];
		cm->allgather_host(&mine, 4, got, e->st);
		bool keep = true;
		direct = true;
		for (int r = 0; r < P; ++r) { direct = direct && (got[r] & 1u); keep = keep && (got[r] & 2u); }
		if (!direct || !keep) shard_unmap_peers(e); // (a send/recv batch grows these buffers on demand: they must not stay mapped)
		if (direct && !e->p2pMapped) {
			e->gLrx[0].need(capG); e->gLrx[1].need(capG); e->sidrx[0].need(capG); e->sidrx[1].need(capG);
			if (!e->dRoute) RB2_CUDA(cudaMalloc(&e->dRoute, sizeof(PeerRoute)));
			void *mineBuf[4] = { e->gLrx[0].p, e->gLrx[1].p, e->sidrx[0].p, e->sidrx[1].p };
			void **peerTab[4] = { (void**)e->peerGL[0], (void**)e->peerGL[1], (void**)e->peerSid[0], (void**)e->peerSid[1] };
			for (int k = 0; k < 4 && direct; ++k)
				if (!cm->p2p_map(mineBuf[k], peerTab[k], e->st)) { for (int j = 0; j < k; ++j) cm->p2p_unmap(peerTab[j]); direct = false; }
			e->p2pMapped = direct;
		}
	}
	int64_t *(*const peerGL)[RB2_MAX_RANKS] = e->peerGL;
	uint32_t G = 0, M = 0;
	uint32_t gBkt[NBA], mBkt[NBA];
	uint64_t mglob[NBMAX];
	memset(mglob, 0, sizeof(mglob)); mglob[0] = mAll;
	uint64_t Gglob = sorted ? 1 : mAll, Mglob = mAll;
	mark();
	const int cs = 0; // current state lives in buffer 0; buffer 1 receives a column's output in source order
	int gcur = 0;     // (the interval starts alternate between gLrx[0] and gLrx[1]: peers may write the next while I read the current)
	if (e->owner[0] == me) {
		e->gLrx[0].need(Gglob + 64); e->gSize[0].need(Gglob + 64); e->gOff[0].need(Gglob + 64); e->sidrx[0].need(mAll + 64);
		LAUNCH(e, k_init_state, cdiv(mAll, 256), 256, 0, sorted, (uint32_t)mAll, n0, e->gLrx[0].p, e->gSize[0].p, e->gOff[0].p, e->sidrx[0].p);
		G = (uint32_t)Gglob; M = (uint32_t)mAll;
	}
	for (int b = 0; b < NBA; ++b) { gBkt[b] = b == 0 ? 0 : G; mBkt[b] = b == 0 ? 0 : M; }
	RB2_CUDA(cudaStreamSynchronize(e->st));
	ph_collect(e, 1u << PH_TRANSPOSE);

	std::vector<ShardTab> tabs(P);
	std::vector<Piece> pcG, pcM;
	bool latePending = false;
	for (int64_t col = 0; Mglob > 0; ++col) {
		if ((uint64_t)col >= ncolAll) RB2_FATAL("internal: live strings beyond the last column");
		Dir &d = e->dir[e->cur];
		int64_t *const gLc = e->gLrx[gcur].p; // interval starts of this column (null on a rank without groups)
		uint32_t *const sidc = e->sidrx[gcur].p; // string ids of this column
		// all-singleton column without interval sizes, on every rank: with direct delivery the merge kernel hands
		// the ids on as well and the column needs no send/recv at all
		const bool colSingle = flat && !useSizes && Gglob == Mglob;
		// ---- per-column capacity (contents of these buffers are dead here) ---------------------
		// (a group yields at most one next group and one record per symbol, plus records for counts above the run limit)
		const size_t gnMax = std::min<uint64_t>(M, 6ull * G);
		e->gL[1].need(gnMax + 64); e->gSize[1].need(useSizes ? gnMax + 64 : 0); e->gOff[1].need(gnMax + 64); e->sid[1].need((size_t)M + 64);
		e->asym.need((size_t)M + 64);
		const size_t recCap = gnMax + M / RB2_MAXRUN + 64;
		e->recP.need(recCap); e->recSC.need(recCap); e->recDst.need(recCap);
		if (flat) { e->recPre.need(recCap + 1); shard_dir_offsets(e, e->gtot, true); }
		else reserve_items(e, std::min<uint64_t>(e->poolCap, recCap) + recCap / RMAX + 2);
		if (useSizes) e->sizes6.need((size_t)G * 6);
		// ---- control block -------------------------------------------------------------------
		for (int b = 0; b < NBA; ++b) { h->gBkt[b] = gBkt[b]; h->mBkt[b] = mBkt[b]; h->blkBkt[b] = e->blkBkt[b]; }
		{ // start of bucket a after this column, whole index (every member inserts one symbol into its sub-bucket)
			int64_t acc = 0;
			for (int s = 0; s < NBMAX; ++s) {
				if (s % 6 == 0) h->cpost[s / 6] = acc;
				for (int a = 0; a < 6; ++a) acc += e->gtot[s][a];
				acc += (int64_t)mglob[s];
			}
			h->cpost[6] = h->cpost[7] = acc;
		}
		h->nb = NBMAX; h->tables = 1;
		h->poolCap = e->poolCap; h->nItems = 0; h->err = 0; h->overflow = 0; h->failBase = NONE32; h->nTodo = 0; h->todoNext = 0; h->nTodoA = 0; h->todoANext = 0;
		h->nrec = 0; h->Gnext = 0; h->Mnext = 0;
		memset(h->grpPre, 0, sizeof(h->grpPre)); memset(h->memPre, 0, sizeof(h->memPre));
		memset(h->gSymBase, 0, sizeof(h->gSymBase)); memset(h->mSymBase, 0, sizeof(h->mSymBase));
		ctl_push(e);

		uint32_t nrec = 0;
		const bool lean = flat && !useSizes && M > 0 && G == M && Gglob == Mglob; // records = the state arrays (RecView)
		// the interval starts of this column are still arriving (second stream) while its first kernels run
		auto wait_late = [&]() {
			if (!latePending) return;
			RB2_CUDA(cudaEventSynchronize(e->ev[PH_EXCH][1]));
			ph_collect(e, 1u << PH_EXCH);
			latePending = false;
		};
		if (M > 0) {
			// ---- members: next symbol + tile histograms -------------------------------------
			ph_begin(e, PH_MEMBERS);
			const uint32_t nTile = cdiv(M, MEM_TILE);
			e->tileB.need(((size_t)nTile + 1) * 6 + 8);
			RB2_CUDA(cudaMemsetAsync(e->tileB.p + (size_t)nTile * 6, 0, 24, e->st));
			TView tv; memset(&tv, 0, sizeof(tv));
			tv.n = P;
			for (int r = 0; r <= P; ++r) tv.off[r] = (uint32_t)strOff[r];
			for (int r = 0; r < P; ++r) tv.col[r] = (uint64_t)col < all[r].ncol ? e->T.p + tOff[r] + (size_t)col * t_stride(all[r].m) : (const uint8_t*)0;
			LAUNCH(e, k_member_fetch, nTile, 256, 0, tv, sidc, M, e->asym.p, e->tileB.p);
			run_mid<6, uint32_t>(e, e->tileB.p, (uint64_t)nTile + 1, e->dctl->memTot, e->midTmp);
			ph_end(e, PH_MEMBERS);
			// ---- groups ---------------------------------------------------------------------
			if (!lean) wait_late(); // the group kernels read the interval starts
			ph_begin(e, PH_GROUPS);
			if (useSizes) {
				if (flat) LAUNCH(e, k_flat_rank_groups, cdiv(G, 128), 128, 0, e->flat.s[e->flat.cur].p, e->flat.dir[e->flat.cur].p, G, gLc, e->gSize[cs].p,
				                 e->sizes6.p, e->dctl, e->dDirOffPre, e->nb);
				else LAUNCH(e, k_rank_groups, cdiv(G, 128), 128, 0, e->pool, d, e->nlog, G, gLc, e->gSize[cs].p, e->sizes6.p, e->dctl);
			}
			if (G == M) {
				LAUNCH(e, k_col_bases_single, 1, 1, 0, e->dctl, e->gOff[1].p, M, flat ? e->recPre.p : (uint32_t*)0);
				SingleArgs sa = { sidc, e->asym.p, M, e->tileB.p, gLc, e->gSize[cs].p, useSizes ? e->sizes6.p : 0, e->dctl,
				                  direct && colSingle ? (uint32_t*)0 : e->sid[1].p, e->gSize[1].p, e->gOff[1].p, e->recP.p, e->recSC.p, e->recDst.p, flat ? e->recPre.p : (uint32_t*)0, lean ? 1 : 0 };
				if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_column_singletons<true>), nTile, 256, 0, sa);
				else LAUNCH(e, (k_column_singletons<false>), nTile, 256, 0, sa);
				ph_end(e, PH_GROUPS);
				ph_begin(e, PH_MEMBERS2); ph_end(e, PH_MEMBERS2);
			} else {
				const uint32_t nGC = cdiv(G, 256);
				e->grpCta.need((size_t)nGC * NGC + NGC);
				GroupArgs ga = { e->gOff[cs].p, gLc, e->gSize[cs].p, useSizes ? e->sizes6.p : 0, e->asym.p, e->tileB.p, G,
				                 e->grpCta.p, e->gSize[1].p, e->gOff[1].p, e->recP.p, e->recSC.p, e->recDst.p, e->dctl, flat ? e->recPre.p : (uint32_t*)0 };
				if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_group_pass<0, true>), nGC, 256, 0, ga);
				else LAUNCH(e, (k_group_pass<0, false>), nGC, 256, 0, ga);
				run_mid<NGC, uint32_t>(e, e->grpCta.p, (uint64_t)nGC, e->dctl->grpTot, e->midTmp);
				LAUNCH(e, k_col_bases, 1, 1, 0, e->dctl, e->gOff[1].p, flat ? e->recPre.p : (uint32_t*)0, M);
				if (e->so == RB2_SO_RCLO) LAUNCH(e, (k_group_pass<1, true>), nGC, 256, 0, ga);
				else LAUNCH(e, (k_group_pass<1, false>), nGC, 256, 0, ga);
				ph_end(e, PH_GROUPS);
				ph_begin(e, PH_MEMBERS2);
				LAUNCH(e, k_partition, nTile, 256, 0, sidc, e->asym.p, M, e->tileB.p, e->dctl, e->sid[1].p);
				ph_end(e, PH_MEMBERS2);
			}
			{ const double t0 = trace ? now_ms() : 0; ctl_pull(e); if (trace) trCtl += now_ms() - t0; }
			ph_collect(e, (1u << PH_MEMBERS) | (1u << PH_GROUPS) | (1u << PH_MEMBERS2) | e->flat.pending);
			e->flat.pending = 0;
			nrec = h->nrec;
			wait_late();
			if (flat && nrec) LAUNCH(e, k_flat_localize, cdiv(nrec, 256), 256, 0, lean ? gLc : e->recP.p, nrec, e->dctl, e->dDirOffPre, e->nb);
		}
		// ---- gather every rank's tables -----------------------------------------------------------
		ShardTab mineT;
		memcpy(mineT.mem, h->memPre, sizeof(mineT.mem));
		memcpy(mineT.grp, (M > 0 && G == M) ? h->memPre : h->grpPre, sizeof(mineT.grp));
		{ const double t0 = trace ? now_ms() : 0; cm->allgather_host(&mineT, sizeof(ShardTab), tabs.data(), e->st); if (trace) trGather += now_ms() - t0; }
		// post-column totals; symbol bases of every rank's output arrays
		uint32_t gSym[RB2_MAX_RANKS][8], mSym[RB2_MAX_RANKS][8];
		for (int r = 0; r < P; ++r) {
			uint32_t g = 0, mm = 0;
			gSym[r][0] = mSym[r][0] = 0;
			for (int a = 1; a < 6; ++a) { gSym[r][a] = g; mSym[r][a] = mm; g += tabs[r].grp[NBMAX * 6 + a]; mm += tabs[r].mem[NBMAX * 6 + a]; }
			for (int s = 0; s < NBMAX; ++s) {
				uint64_t inBkt = 0;
				for (int a = 0; a < 6; ++a) {
					const uint32_t nm = tabs[r].mem[(s + 1) * 6 + a] - tabs[r].mem[s * 6 + a];
					if (nm && e->owner[s] != r) RB2_FATAL("internal: rank %d reports members in sub-bucket %d it does not own", r, s);
					e->gtot[s][a] += nm; inBkt += nm;
				}
				if (e->owner[s] == r && inBkt != mglob[s]) RB2_FATAL("internal: member count of sub-bucket %d diverged (%llu vs %llu)", s, (unsigned long long)inBkt, (unsigned long long)mglob[s]);
			}
		}
		shard_dir_offsets(e, e->gtot); // the directory rebuilt behind the merge is in post-column coordinates
		// ---- transfer plan: piece (a,x,y) goes from owner(x,y) to owner(a,x), pieces of one target in y order ----
		pcG.clear(); pcM.clear();
		uint64_t curG[RB2_MAX_RANKS], curM[RB2_MAX_RANKS], mglobNext[NBMAX];
		uint32_t gBktN[NBA], mBktN[NBA];
		memset(curG, 0, sizeof(curG)); memset(curM, 0, sizeof(curM)); memset(mglobNext, 0, sizeof(mglobNext));
		int nMyPieces = 0, nMyOut = 0;
		uint8_t pieceOfMine[6 * NBMAX]; // (symbol, source sub-bucket of mine) -> index among the pieces I send
		memset(pieceOfMine, 0, sizeof(pieceOfMine));
		uint32_t *hPlan = e->hPlan + (col & 1) * 2 * (NBMAX * 6 + 8); // the copy of the previous column may still be queued
		for (int t = 0; t < NBMAX; ++t) {
			const int a = t / 6, x = t % 6, dst = e->owner[t];
			gBktN[t] = (uint32_t)curG[me]; mBktN[t] = (uint32_t)curM[me];
			if (a == 0) continue; // nothing continues into bucket $
			for (int y = 0; y < 6; ++y) {
				const int sb = x * 6 + y, src = e->owner[sb];
				const uint32_t ng = tabs[src].grp[(sb + 1) * 6 + a] - tabs[src].grp[sb * 6 + a];
				const uint32_t nm = tabs[src].mem[(sb + 1) * 6 + a] - tabs[src].mem[sb * 6 + a];
				if (nm == 0) continue;
				Piece pg = { src, dst, (uint64_t)gSym[src][a] + tabs[src].grp[sb * 6 + a], curG[dst], ng };
				Piece pm = { src, dst, (uint64_t)mSym[src][a] + tabs[src].mem[sb * 6 + a], curM[dst], nm };
				// neighbouring pieces of one (source, target) pair that are contiguous on both sides travel as one
				// message (NCCL's point-to-point bandwidth collapses with many small messages per peer)
				if (!pcG.empty() && pcG.back().src == src && pcG.back().dst == dst && pcG.back().so + pcG.back().n == pg.so &&
				    pcG.back().dof + pcG.back().n == pg.dof && pcM.back().so + pcM.back().n == pm.so && pcM.back().dof + pcM.back().n == pm.dof) {
					pcG.back().n += ng; pcM.back().n += nm;
				} else {
					pcG.push_back(pg); pcM.push_back(pm);
					if (src == me) ++nMyOut;
					if (dst == me) { // rebase table of the member ranges I receive
						hPlan[nMyPieces] = (uint32_t)pg.dof;
						hPlan[NBMAX * 6 + 8 + nMyPieces] = (uint32_t)pm.dof - (uint32_t)pm.so;
						++nMyPieces;
					}
				}
				if (src == me) pieceOfMine[a * NBMAX + sb] = (uint8_t)(nMyOut - 1);
				curG[dst] += ng; curM[dst] += nm; mglobNext[t] += nm;
			}
		}
		for (int t = NBMAX; t < NBA; ++t) { gBktN[t] = (uint32_t)curG[me]; mBktN[t] = (uint32_t)curM[me]; }
		uint64_t GglobN = 0, MglobN = 0;
		for (int r = 0; r < P; ++r) { GglobN += curG[r]; MglobN += curM[r]; }
		for (int r = 0; r < P; ++r) if (curM[r] >= 0xfffffff0ull) RB2_FATAL("too many strings on one rank");

		// ---- merge my records | move the string state to the owners of the next sub-buckets --------------
		// Member ids, member ranges and interval sizes are final before the merge: they travel on a second
		// stream while the merge runs; the interval starts (ranks) follow behind the merge.
		const uint32_t Gn = (uint32_t)curG[me], Mn = (uint32_t)curM[me];
		const bool singles = GglobN == MglobN;
		// (RB2_EXCH_SERIAL=1, an experiment switch: on the main stream in front of the merge instead of beside it)
		static const bool exchSerial = getenv("RB2_EXCH_SERIAL") && atoi(getenv("RB2_EXCH_SERIAL"));
		auto exchange_early = [&]() {
			cudaStream_t xs = exchSerial ? e->st : e->st2;
			cm->group_begin();
			if (useSizes) cm->exchange(e->gSize[1].p, e->gSize[cs].p, 8, pcG.data(), (int)pcG.size(), xs);
			if (!singles) cm->exchange(e->gOff[1].p, e->gOff[cs].p, 4, pcG.data(), (int)pcG.size(), xs);
			cm->exchange(e->sid[1].p, e->sidrx[gcur ^ 1].p, 4, pcM.data(), (int)pcM.size(), xs);
			cm->group_end(xs);
			RB2_CUDA(cudaEventRecord(e->evEarly, xs));
		};
		// direct delivery: the pieces of MY output order and where each lands (the next column's buffer of its target rank)
		const bool deliver = direct && MglobN > 0;
		const bool deliverIds = deliver && colSingle; // (then nothing is left for exchange_early)
		if (deliver && nrec > 0) {
			PeerRoute rt; memset(&rt, 0, sizeof(rt));
			int np = 0;
			for (size_t k = 0; k < pcG.size(); ++k) if (pcG[k].src == me) {
				if (np >= ROUTE_MAXPC) RB2_FATAL("internal: more than %d output pieces", ROUTE_MAXPC);
				rt.base[np] = peerGL[gcur ^ 1][pcG[k].dst] + (int64_t)pcG[k].dof - (int64_t)pcG[k].so;
				// (all-singleton column: member pieces = group pieces)
				rt.base32[np] = e->peerSid[gcur ^ 1][pcM[k].dst] + (int64_t)pcM[k].dof - (int64_t)pcM[k].so;
				if (deliverIds && (pcM[k].so != pcG[k].so || pcM[k].n != pcG[k].n)) RB2_FATAL("internal: group and member pieces of an all-singleton column differ");
				++np;
			}
			if (np != nMyOut) RB2_FATAL("internal: output pieces miscounted");
			// (no piece at all: every record of mine ends its string -- none has a target)
			memcpy(rt.pieceOf, pieceOfMine, sizeof(rt.pieceOf));
			LAUNCH(e, k_route_store, 1, 128, 0, e->dRoute, rt);
		}
		auto merge = [&]() {
			if (flat) { // (no records: my array does not change)
				if (deliverIds && nrec > 0 && !lean) RB2_FATAL("internal: all-singleton column with full records");
				if (nrec > 0) flat_apply_records(e, nrec, M, deliver ? (int64_t*)0 : e->gL[1].p, lean ? gLc : (const int64_t*)0, 0, deliver ? e->dRoute : (const PeerRoute*)0,
				                                 deliverIds ? sidc : (const uint32_t*)0);
			}
			else if (nrec > 0) apply_records(e, nrec, e->gL[1].p);
			else rebuild_directory(e, false);
		};
		// the dense merge is fully asynchronous, so it is queued first; the block merge synchronises with the host
		wait_late(); // (ranks without members get here with the previous column's transfer possibly still in flight)
		if (MglobN > 0) { // the next column's arrays (their old contents are dead: this column's kernels are through)
			e->gSize[cs].need(useSizes ? (size_t)Gn + 64 : 0); e->gOff[cs].need((size_t)Gn + 64);
			if (!direct) e->sidrx[gcur ^ 1].need((size_t)Mn + 64);
		}
		if (MglobN > 0 && (!flat || exchSerial) && !deliverIds) exchange_early();
		merge();
		if (MglobN > 0 && flat && !exchSerial && !deliverIds) exchange_early();
		e->stats.n_records += nrec;
		++e->stats.n_columns;
		if (MglobN > 0) {
			if (deliver) {
				// interval starts: the merge kernels have stored them where they belong; once every rank's merge
				// is through, every rank's next-column buffer is complete
				cm->barrier_stream(e->st);
			} else {
				// interval starts: behind the merge, on the second stream -- the next column only needs them when
				// its records are merged (all-singleton columns) or its groups are scanned
				RB2_CUDA(cudaEventRecord(e->evMerge, e->st));
				e->gLrx[gcur ^ 1].need((size_t)Gn + 64);
				RB2_CUDA(cudaStreamWaitEvent(e->st2, e->evMerge, 0));
				RB2_CUDA(cudaEventRecord(e->ev[PH_EXCH][0], e->st2));
				cm->group_begin();
				cm->exchange(e->gL[1].p, e->gLrx[gcur ^ 1].p, 8, pcG.data(), (int)pcG.size(), e->st2);
				cm->group_end(e->st2);
				RB2_CUDA(cudaEventRecord(e->ev[PH_EXCH][1], e->st2));
				latePending = true;
			}
			gcur ^= 1;
			// member ids / ranges arrived on the second stream: finish the member ranges on the main one
			if (!deliverIds) RB2_CUDA(cudaStreamWaitEvent(e->st, e->evEarly, 0));
			if (singles) { if (Gn + 1 > 0) LAUNCH(e, k_fill_u32, cdiv((uint64_t)Gn + 1, 256), 256, 0, e->gOff[cs].p, Gn + 1, 0u, 1u); }
			else if (Gn > 0) {
				e->plan.need(2 * (NBMAX * 6 + 8));
				RB2_CUDA(cudaMemcpyAsync(e->plan.p, hPlan, 2 * (NBMAX * 6 + 8) * sizeof(uint32_t), cudaMemcpyHostToDevice, e->st));
				LAUNCH(e, k_rebase_goff, cdiv(Gn, 256), 256, 0, e->gOff[cs].p, Gn, Mn, e->plan.p, e->plan.p + NBMAX * 6 + 8, nMyPieces);
			}
			if (!flat) { // (the dense regime keeps the host running ahead: its events are collected at the next control-block read)
				RB2_CUDA(cudaStreamSynchronize(e->st));
				ph_collect(e, e->flat.pending);
				e->flat.pending = 0;
			}
			e->stats.exch_bytes += ((int64_t)Gn * (8 + (useSizes ? 8 : 0) + (singles ? 0 : 4)) + (int64_t)Mn * 4);
		}
		// ---- advance -------------------------------------------------------------------------------
		G = Gn; M = Mn;
		for (int b = 0; b < NBA; ++b) { gBkt[b] = gBktN[b]; mBkt[b] = mBktN[b]; }
		memcpy(mglob, mglobNext, sizeof(mglob));
		Gglob = GglobN; Mglob = MglobN;
	}
	mark();
	if (direct) ++e->stats.p2p_batches; // (the peer mappings stay for the next batch)
	if (flat) {
		RB2_CUDA(cudaStreamSynchronize(e->st));
		ph_collect(e, e->flat.pending); e->flat.pending = 0;
		flat_finish(e); // the array stays resident; the leaf blocks of my sub-buckets are rebuilt when somebody fetches them
		++e->stats.flat_batches;
	}
	shard_publish_totals(e);
	mark();
	if (trace) fprintf(stderr, "[rb2 trace] rank %d: split+replicate+regime %.1f ms, peer mappings %.1f ms, columns %.1f ms, unmap+finish %.1f ms (direct=%d); "
	                   "in the column loop the host waited %.1f ms for the columns' counts and %.1f ms in the table all-gathers\n",
	                   me, tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], (int)direct, trCtl, trGather);
	e->stats.pool_blocks = e->hctl->poolUsed;
	e->stats.pool_capacity = e->poolCap;
}

extern "C" void rb2_insert_multi_sharded_dev(rb2_engine_t *e, int64_t len, const uint8_t *s_dev)
{
	if (!e->comm) RB2_FATAL("rb2_insert_multi_sharded: engine was not created with rb2_create_sharded");
	RB2_CUDA(cudaSetDevice(e->dev));
	if (((uintptr_t)s_dev & 15) != 0) RB2_FATAL("device batch must be 16-byte aligned");
	RB2_CUDA(cudaEventRecord(e->evTot[0], e->st));
	insert_sharded_batch(e, len, s_dev);
	RB2_CUDA(cudaEventRecord(e->evTot[1], e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, e->evTot[0], e->evTot[1]));
	e->stats.ms_total += ms;
}

extern "C" void rb2_insert_multi_sharded(rb2_engine_t *e, int64_t len, const uint8_t *s)
{
	if (!e->comm) RB2_FATAL("rb2_insert_multi_sharded: engine was not created with rb2_create_sharded");
	if (len < 0 || (len > 0 && s[len - 1] != 0)) RB2_FATAL("mr_insert_multi: a non-empty batch must end with NUL (mrope.c:268)");
	RB2_CUDA(cudaSetDevice(e->dev));
	RB2_CUDA(cudaEventRecord(e->evTot[0], e->st));
	e->sbuf.need((size_t)len + 16);
	if (len > 0) {
		ph_begin(e, PH_H2D);
		RB2_CUDA(cudaMemcpyAsync(e->sbuf.p, s, (size_t)len, cudaMemcpyHostToDevice, e->st));
		ph_end(e, PH_H2D);
		RB2_CUDA(cudaStreamSynchronize(e->st));
		ph_collect(e, 1u << PH_H2D);
	}
	insert_sharded_batch(e, len, e->sbuf.p);
	RB2_CUDA(cudaEventRecord(e->evTot[1], e->st));
	RB2_CUDA(cudaStreamSynchronize(e->st));
	float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, e->evTot[0], e->evTot[1]));
	e->stats.ms_total += ms;
}
