// rb2_async.inl -- the host entry point as a two-stage pipeline; included by rb2_engine.cu.
//
// mr_insert_multi's contract is: the buffer is borrowed and the caller reuses it as soon
// as the call returns (main.c:243) -- so the COPY must be complete on return, the insertion need not be.  The
// reference driver spends the time between two calls parsing the next 10 GB of reads on one host thread; a
// production caller streams batches back to back.  Either way the GPU work of batch k can run while batch k+1
// is parsed / copied:
//
//   caller thread   rb2_insert_multi: wait for a free staging buffer -> H2D copy on the copy stream ->
//                   k_pair_hist on the copied bytes (the marginal counts mr_get_c reports are known from the
//                   batch alone) -> queue the batch -> return
//   worker thread   takes batches (and resets) in order and runs insert_device_batch on the engine's stream
//
// Everything else in the C-ABI first waits for the queue to drain (async_drain), so nothing observes a half
// inserted batch; after a drain the marginals computed from the batches must equal the index's own totals --
// checked every time, a free end-to-end self test.  Sharded engines, RB2_GPUS clusters and RB2_ASYNC=0 keep the
// synchronous path.
struct AsyncJob { int kind; int64_t len; int stage; }; // kind 0 = insert the batch in staging buffer `stage`, 1 = reset the index

struct AsyncState {
	std::thread th; std::mutex mu; std::condition_variable cv;
	std::deque<AsyncJob> q; bool busy = false, stop = false, started = false;
	DevBuf<uint8_t> stage[2]; bool stageBusy[2] = { false, false }; int next = 0;
	cudaStream_t copySt = 0; cudaEvent_t evH[2], evSpan[2];
	unsigned long long *dHist = 0, *hHist = 0;
	int64_t pub[6][6];   // marginal counts of everything submitted so far (what rb2_counts answers without waiting)
	std::vector<rb2_stats_t> hist; std::vector<double> histH2d; // per batch: what the insertion added to the statistics, and its copy time
	double h2dMs = 0;    // copy time of the caller's thread (folded into the statistics when they are read)
};

static bool async_enabled(const rb2_engine *e)
{
	static int pref = -1;
	if (pref < 0) { const char *s = getenv("RB2_ASYNC"); pref = !(s && *s == '0'); }
	return pref && !e->comm && !e->nChild;
}

// b - a, field by field (every field of rb2_stats_t is 8 bytes wide: int64_t or double)
static rb2_stats_t stats_delta(const rb2_stats_t &a, const rb2_stats_t &b)
{
	static_assert(sizeof(rb2_stats_t) == 24 * 8, "rb2_stats_t: 24 fields of 8 bytes");
	const uint32_t isDouble = 0xffu << 10 | 1u << 19 | 1u << 21; // ms_total .. ms_merge_general, ms_exchange, ms_convert
	rb2_stats_t d;
	for (int i = 0; i < 24; ++i) {
		if (isDouble >> i & 1) ((double*)&d)[i] = ((const double*)&b)[i] - ((const double*)&a)[i];
		else ((int64_t*)&d)[i] = ((const int64_t*)&b)[i] - ((const int64_t*)&a)[i];
	}
	d.pool_blocks = b.pool_blocks; d.pool_capacity = b.pool_capacity;
	return d;
}

static void async_worker(rb2_engine *e)
{
	AsyncState &A = *e->as;
	RB2_CUDA(cudaSetDevice(e->dev));
	for (;;) {
		AsyncJob j;
		{
			std::unique_lock<std::mutex> lk(A.mu);
			A.cv.wait(lk, [&] { return A.stop || !A.q.empty(); });
			if (A.q.empty()) return; // stop
			j = A.q.front(); A.q.pop_front(); A.busy = true;
		}
		const rb2_stats_t before = e->stats;
		if (j.kind == 0) {
			RB2_CUDA(cudaEventRecord(e->evTot[0], e->st));
			insert_device_batch(e, j.len, A.stage[j.stage].p);
			RB2_CUDA(cudaEventRecord(e->evTot[1], e->st));
			RB2_CUDA(cudaStreamSynchronize(e->st));
			float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, e->evTot[0], e->evTot[1]));
			e->stats.ms_total += ms;
		} else reset_index(e);
		{
			std::lock_guard<std::mutex> lk(A.mu);
			if (j.kind == 0) { A.stageBusy[j.stage] = false; A.hist.push_back(stats_delta(before, e->stats)); }
			A.busy = false;
		}
		A.cv.notify_all();
	}
}

static void async_init(rb2_engine *e)
{
	e->as = new AsyncState();
	AsyncState &A = *e->as;
	memset(A.pub, 0, sizeof(A.pub));
	RB2_CUDA(cudaStreamCreateWithFlags(&A.copySt, cudaStreamNonBlocking));
	for (int k = 0; k < 2; ++k) { RB2_CUDA(cudaEventCreate(&A.evH[k])); RB2_CUDA(cudaEventCreate(&A.evSpan[k])); }
	RB2_CUDA(cudaMalloc(&A.dHist, 36 * sizeof(unsigned long long)));
	RB2_CUDA(cudaMallocHost(&A.hHist, 36 * sizeof(unsigned long long)));
}

// wait until the worker has nothing left to do; then the index's own totals must equal the submitted marginals
static void async_drain(rb2_engine *e)
{
	if (!e->as || !e->as->started) return;
	AsyncState &A = *e->as;
	{
		std::unique_lock<std::mutex> lk(A.mu);
		A.cv.wait(lk, [&] { return A.q.empty() && !A.busy; });
	}
	for (int b = 0; b < 6; ++b) for (int a = 0; a < 6; ++a)
		if (A.pub[b][a] != e->tot[b][a])
			RB2_FATAL("internal: bucket %d holds %lld symbols %d, the batches submitted so far %lld", b, (long long)e->tot[b][a], a, (long long)A.pub[b][a]);
}

static void async_destroy(rb2_engine *e)
{
	if (!e->as) return;
	AsyncState &A = *e->as;
	if (A.started) {
		{ std::lock_guard<std::mutex> lk(A.mu); A.stop = true; }
		A.cv.notify_all();
		A.th.join();
	}
	A.stage[0].release(); A.stage[1].release();
	cudaStreamDestroy(A.copySt);
	for (int k = 0; k < 2; ++k) { cudaEventDestroy(A.evH[k]); cudaEventDestroy(A.evSpan[k]); }
	cudaFree(A.dHist); cudaFreeHost(A.hHist);
	delete e->as; e->as = 0;
}

static void async_submit(rb2_engine *e, AsyncJob j)
{
	AsyncState &A = *e->as;
	if (!A.started) { A.started = true; A.th = std::thread(async_worker, e); }
	{ std::lock_guard<std::mutex> lk(A.mu); A.q.push_back(j); }
	A.cv.notify_all();
}

// the host entry point, asynchronous flavour: copy, count, queue
static void async_insert_multi(rb2_engine *e, int64_t len, const uint8_t *s)
{
	AsyncState &A = *e->as;
	const int b = A.next; A.next ^= 1;
	{
		std::unique_lock<std::mutex> lk(A.mu);
		A.cv.wait(lk, [&] { return !A.stageBusy[b]; });
		A.stageBusy[b] = true;
	}
	A.stage[b].need((size_t)len + 64);
	RB2_CUDA(cudaMemsetAsync(A.dHist, 0, 36 * sizeof(unsigned long long), A.copySt));
	RB2_CUDA(cudaEventRecord(A.evH[0], A.copySt));
	RB2_CUDA(cudaMemcpyAsync(A.stage[b].p, s, (size_t)len, cudaMemcpyHostToDevice, A.copySt));
	RB2_CUDA(cudaEventRecord(A.evH[1], A.copySt));
	RB2_KERNEL_LAUNCH(k_pair_hist, (unsigned)std::min<int64_t>((len + 4095) / 4096, (int64_t)e->nSM * 4), 256, 0, A.copySt, A.stage[b].p, len, A.dHist);
	RB2_CUDA(cudaMemcpyAsync(A.hHist, A.dHist, 36 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, A.copySt));
	RB2_CUDA(cudaStreamSynchronize(A.copySt)); // the caller may reuse its buffer now
	float ms = 0; RB2_CUDA(cudaEventElapsedTime(&ms, A.evH[0], A.evH[1]));
	int64_t tot = 0;
	for (int k = 0; k < 36; ++k) { A.pub[k / 6][k % 6] += (int64_t)A.hHist[k]; tot += (int64_t)A.hHist[k]; }
	if (tot != len) RB2_FATAL("mr_insert_multi: the batch holds bytes outside the nt6 alphabet 0..5 (%lld of %lld are valid)", (long long)tot, (long long)len);
	A.h2dMs += ms;
	{ std::lock_guard<std::mutex> lk(A.mu); A.histH2d.push_back(ms); }
	AsyncJob j = { 0, len, b };
	async_submit(e, j);
}

static void async_reset(rb2_engine *e)
{
	AsyncState &A = *e->as;
	memset(A.pub, 0, sizeof(A.pub));
	if (!A.started) { reset_index(e); return; }
	AsyncJob j = { 1, 0, 0 };
	async_submit(e, j);
}
