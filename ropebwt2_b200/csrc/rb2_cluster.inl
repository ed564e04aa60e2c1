// rb2_cluster.inl -- several GPUs behind the ONE-engine C-ABI; included by rb2_engine.cu.
//
// With RB2_GPUS=P (P > 1) in the environment rb2_create_auto() (what mr_init calls) returns a proxy engine that owns P sharded
// engines (ranks = threads of this process, LocalComm over peer copies; one GPU each, or all on one
// device with RB2_GPUS_SAME_DEVICE=1).  The reference-facing API on top (mrope.h: mr_insert_multi, the
// block iterator, mr_dump, mr_rank2a, the count mirrors) works unchanged: a bucket of the reference is
// the concatenation of its six sub-buckets, each fetched from its owner.  That is how the unmodified
// reference driver (main.c) uses every GPU of a node:  RB2_GPUS=8 ropebwt2_b200 -LRs reads.txt
#include <thread>


static void cluster_attach(rb2_engine *e, int device, int sorting_order)
{
	const char *s = getenv("RB2_GPUS");
	const int P = s && *s ? atoi(s) : 0;
	if (P <= 1) return;
	if (P > RB2_MAX_RANKS) RB2_FATAL("RB2_GPUS=%d: at most %d", P, RB2_MAX_RANKS);
	const char *same = getenv("RB2_GPUS_SAME_DEVICE");
	const int nd = rb2_device_count();
	if (!(same && *same && *same != '0') && P > nd) RB2_FATAL("RB2_GPUS=%d but only %d CUDA devices are visible", P, nd);
	e->grp = rb2_group_create(P);
	for (int r = 0; r < P; ++r)
		e->child[r] = rb2_create_sharded(same && *same && *same != '0' ? device : (device + r) % nd, sorting_order, r, P, e->grp, 0);
	e->nChild = P;
	RB2_CUDA(cudaSetDevice(e->dev));
}

static void cluster_publish(rb2_engine *e) // whole-index marginals are identical on every rank
{
	memcpy(e->tot, e->child[0]->tot, sizeof(e->tot));
	memcpy(e->bktLen, e->child[0]->bktLen, sizeof(e->bktLen));
}

static void cluster_destroy(rb2_engine *e)
{
	for (int r = 0; r < e->nChild; ++r) rb2_destroy(e->child[r]);
	rb2_group_destroy(e->grp);
	e->nChild = 0;
}

static void cluster_reset(rb2_engine *e)
{
	for (int r = 0; r < e->nChild; ++r) rb2_reset(e->child[r]);
	cluster_publish(e);
}

// the batch is cut into P contiguous shares behind string terminators; rank order = input order
static void cluster_insert_multi(rb2_engine *e, int64_t len, const uint8_t *s)
{
	const int P = e->nChild;
	int64_t cut[RB2_MAX_RANKS + 1];
	cut[0] = 0; cut[P] = len;
	for (int r = 1; r < P; ++r) {
		int64_t c = len * r / P;
		if (c < cut[r - 1]) c = cut[r - 1];
		if (c > 0 && c < len) { // move behind the next NUL
			const void *z = memchr(s + c - 1, 0, (size_t)(len - c + 1));
			c = z ? (const uint8_t*)z - s + 1 : len;
		}
		cut[r] = c;
	}
	std::vector<std::thread> th;
	for (int r = 0; r < P; ++r)
		th.emplace_back([=]() { rb2_insert_multi_sharded(e->child[r], cut[r + 1] - cut[r], s + cut[r]); });
	for (auto &t : th) t.join();
	cluster_publish(e);
	rb2_stats_t s0 = e->child[0]->stats; // timings of rank 0, volumes of all ranks
	s0.n_strings = s0.n_symbols = 0;
	for (int r = 0; r < P; ++r) { s0.n_strings += e->child[r]->stats.n_strings; s0.n_symbols += e->child[r]->stats.n_symbols; }
	e->stats = s0;
}

static int64_t cluster_num_blocks(rb2_engine *e, int bucket)
{
	int64_t n = 0;
	for (int y = 0; y < 6; ++y) { const int sb = bucket * 6 + y; n += rb2_num_blocks(e->child[e->child[0]->owner[sb]], sb); }
	return n;
}

static int64_t cluster_fetch_blocks(rb2_engine *e, int bucket, int64_t first, int64_t n, uint8_t *dst, int64_t *cnt)
{
	int64_t done = 0;
	for (int y = 0; y < 6 && n > 0; ++y) {
		const int sb = bucket * 6 + y;
		rb2_engine *c = e->child[e->child[0]->owner[sb]];
		const int64_t nb = rb2_num_blocks(c, sb);
		if (first >= nb) { first -= nb; continue; }
		const int64_t k = rb2_fetch_blocks(c, sb, first, n, dst + done * RB2_BLK, cnt ? cnt + done * 6 : 0);
		done += k; n -= k; first = 0;
	}
	RB2_CUDA(cudaSetDevice(e->dev));
	return done;
}

// occ(., x): the rank that owns the sub-bucket x falls into answers in whole-index coordinates
static void cluster_rank1(rb2_engine *e, int64_t x, int64_t c[6])
{
	for (int a = 0; a < 6; ++a) c[a] = 0;
	if (x <= 0) return;
	rb2_engine *c0 = e->child[0];
	int64_t start = 0;
	for (int sb = 0; sb < NBMAX; ++sb) {
		int64_t l = 0;
		for (int a = 0; a < 6; ++a) l += c0->gtot[sb][a];
		if (l > 0 && x > start && x <= start + l) { rb2_rank2a(e->child[c0->owner[sb]], x, -1, c, 0); return; }
		start += l;
	}
	RB2_FATAL("rank position out of range");
}
