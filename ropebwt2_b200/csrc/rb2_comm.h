// rb2_comm.h -- the exchange layer of the sharded (multi-GPU) build.
//
// One BCR column on P ranks needs three cross-rank steps:
//   allgather_host : a few KB per rank (the per-(sub-bucket, symbol) group / member counts of the
//                    column -- the reference's cross-bucket offsets, mrope.c:332-340, generalised
//                    from six buckets on one host to 36 sub-buckets on P GPUs)
//   exchange       : the string state (interval start, interval size, member ids) moves to the
//                    rank that owns the sub-bucket a string inserts into next -- contiguous pieces,
//                    source offset / destination offset / length known to every rank from the
//                    gathered counts
//   gather_blocks  : once per batch, every rank's column-major symbol matrix is replicated
//   p2p_map / barrier_stream : direct delivery -- every rank maps the other ranks' receive buffers
//                    (CUDA IPC between processes, plain pointers between threads), so that the merge
//                    kernel's epilogue stores each rank it computes straight into the memory of the
//                    rank that needs it next (NVLink stores issued under the merge), and a
//                    stream-ordered barrier replaces the send/recv of those 8 bytes per string
//
// Two back ends behind one interface:
//   NcclComm  : one process per GPU (torchrun); ncclSend/ncclRecv/ncclBroadcast/ncclAllGather over
//               NVLink.  libnccl is dlopen()ed so that the library loads on machines without it and
//               shares the copy torch.distributed has already loaded.
//   LocalComm : P engines driven by P host threads of ONE process (the unmodified reference driver
//               with RB2_GPUS=P, or P "virtual ranks" on a single GPU in the tests): pointers are
//               published through shared host memory and pieces are pulled with peer copies.
#pragma once
#include <atomic>
#include <dlfcn.h>
#include <sched.h>
#include <nccl.h>
#include "rb2_common.cuh"

#define RB2_MAX_RANKS 8

struct Piece { int src, dst; uint64_t so, dof, n; }; // n elements from src's buffer at so to dst's buffer at dof

struct Comm {
	int rank, n;
	virtual ~Comm() {}
	virtual void allgather_host(const void *send, size_t bytes, void *recv, cudaStream_t st) = 0;
	virtual void group_begin() = 0;
	virtual void exchange(const void *sendBuf, void *recvBuf, size_t esz, const Piece *pc, int npc, cudaStream_t st) = 0;
	virtual void group_end(cudaStream_t st) = 0;
	// dst[r] (in MY memory) receives rank r's block of bytes[r] bytes; dst[rank] already holds mine
	virtual void gather_blocks(uint8_t *const *dst, const size_t *bytes, cudaStream_t st) = 0;
	virtual void barrier(cudaStream_t st) = 0;
	// Collective.  peers[r] = an address in MY address space of rank r's buffer `mine` (peers[rank] = mine), writable
	// by my kernels.  false (on every rank alike) when some rank cannot map some buffer: nothing stays mapped.
	virtual bool p2p_map(void *mine, void **peers, cudaStream_t st) = 0;
	virtual void p2p_unmap(void **peers) = 0;
	// Collective, stream ordered: what any rank queued on its stream before this call has completed (peer stores
	// included) before anything a rank queues behind it starts.
	virtual void barrier_stream(cudaStream_t st) = 0;
};

// ------------------------------------------------------------------------------------------
// LocalComm: ranks are threads of one process
// ------------------------------------------------------------------------------------------
struct rb2_group {
	int n;
	std::atomic<int> arrived, phase, joined;
	const void *slot[RB2_MAX_RANKS];
	uint8_t small[RB2_MAX_RANKS][4096];
};

struct LocalComm : Comm {
	rb2_group *g;
	int myPhase;
	LocalComm(rb2_group *g_, int rank_) : g(g_), myPhase(0) { rank = rank_; n = g_->n; }
	void sync_threads() {
		const int next = myPhase + 1;
		if (g->arrived.fetch_add(1) + 1 == g->n) { g->arrived.store(0); g->phase.store(next); }
		else while (g->phase.load() != next) sched_yield();
		myPhase = next;
	}
	void barrier(cudaStream_t st) override { RB2_CUDA(cudaStreamSynchronize(st)); sync_threads(); }
	void allgather_host(const void *send, size_t bytes, void *recv, cudaStream_t) override {
		if (bytes > sizeof(g->small[0])) RB2_FATAL("LocalComm: small all-gather of %zu bytes", bytes);
		memcpy(g->small[rank], send, bytes);
		sync_threads();
		for (int r = 0; r < n; ++r) memcpy((uint8_t*)recv + (size_t)r * bytes, g->small[r], bytes);
		sync_threads();
	}
	void group_begin() override {}
	void group_end(cudaStream_t) override {}
	void exchange(const void *sendBuf, void *recvBuf, size_t esz, const Piece *pc, int npc, cudaStream_t st) override {
		RB2_CUDA(cudaStreamSynchronize(st)); // my outgoing data is complete
		g->slot[rank] = sendBuf;
		sync_threads();
		for (int k = 0; k < npc; ++k) if (pc[k].dst == rank && pc[k].n)
			RB2_CUDA(cudaMemcpyAsync((uint8_t*)recvBuf + pc[k].dof * esz, (const uint8_t*)g->slot[pc[k].src] + pc[k].so * esz,
			                         pc[k].n * esz, cudaMemcpyDefault, st));
		RB2_CUDA(cudaStreamSynchronize(st));
		sync_threads(); // everybody has pulled: send buffers may be reused
	}
	bool p2p_map(void *mine, void **peers, cudaStream_t st) override {
		struct Ad { void *p; int dev; } me_, all[RB2_MAX_RANKS];
		me_.p = mine; RB2_CUDA(cudaGetDevice(&me_.dev));
		allgather_host(&me_, sizeof(me_), all, st);
		int ok = 1, oks[RB2_MAX_RANKS];
		for (int r = 0; r < n; ++r) {
			peers[r] = all[r].p;
			if (all[r].dev != me_.dev) { int can = 0; if (cudaDeviceCanAccessPeer(&can, me_.dev, all[r].dev) != cudaSuccess || !can) ok = 0; } // (enabled by rb2_group_create)
		}
		cudaGetLastError();
		allgather_host(&ok, sizeof(ok), oks, st);
		for (int r = 0; r < n; ++r) ok = ok && oks[r];
		return ok != 0;
	}
	void p2p_unmap(void **) override {}
	void barrier_stream(cudaStream_t st) override { barrier(st); }
	void gather_blocks(uint8_t *const *dst, const size_t *bytes, cudaStream_t st) override {
		RB2_CUDA(cudaStreamSynchronize(st));
		g->slot[rank] = dst[rank];
		sync_threads();
		for (int r = 0; r < n; ++r) if (r != rank && bytes[r])
			RB2_CUDA(cudaMemcpyAsync(dst[r], g->slot[r], bytes[r], cudaMemcpyDefault, st));
		RB2_CUDA(cudaStreamSynchronize(st));
		sync_threads();
	}
};

// ------------------------------------------------------------------------------------------
// NcclComm: ranks are processes
// ------------------------------------------------------------------------------------------
struct NcclApi {
	void *h;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi *nccl_api(void)
{
	static NcclApi api; static int ready = 0;
	if (ready) return &api;
	const char *names[] = { getenv("RB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
	for (int i = 0; i < 3 && !api.h; ++i) if (names[i] && *names[i]) api.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!api.h) RB2_FATAL("cannot load libnccl.so.2 (%s); set RB2_NCCL_LIB", dlerror());
#define RB2_NCCL_SYM(field, name) do { *(void**)(&api.field) = dlsym(api.h, name); if (!api.field) RB2_FATAL("libnccl lacks %s", name); } while (0)
	RB2_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); RB2_NCCL_SYM(CommInitRank, "ncclCommInitRank"); RB2_NCCL_SYM(CommDestroy, "ncclCommDestroy");
	RB2_NCCL_SYM(Send, "ncclSend"); RB2_NCCL_SYM(Recv, "ncclRecv"); RB2_NCCL_SYM(GroupStart, "ncclGroupStart"); RB2_NCCL_SYM(GroupEnd, "ncclGroupEnd");
	RB2_NCCL_SYM(AllGather, "ncclAllGather"); RB2_NCCL_SYM(Broadcast, "ncclBroadcast"); RB2_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef RB2_NCCL_SYM
	ready = 1;
	return &api;
}

#define RB2_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
	RB2_FATAL("NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl_api()->GetErrorString(r_)); } while (0)

struct NcclComm : Comm {
	NcclApi *api; ncclComm_t comm;
	uint8_t *hSend, *hRecv, *dSend, *dRecv; // staging of the small all-gather
	uint8_t *dBar; // send + receive words of barrier_stream
	enum { SMALL = 4096 };
	NcclComm(int rank_, int n_, const void *uid) {
		rank = rank_; n = n_;
		// the exchange is a handful of large point-to-point transfers per column: let NCCL spread each over
		// many channels (its default for send/recv is tuned for many small peers); the caller's setting wins
		setenv("NCCL_MIN_P2P_NCHANNELS", "16", 0);
		setenv("NCCL_MAX_P2P_NCHANNELS", "32", 0);
		api = nccl_api();
		ncclUniqueId id; memcpy(&id, uid, sizeof(id));
		RB2_NCCL(api->CommInitRank(&comm, n, id, rank));
		RB2_CUDA(cudaMallocHost(&hSend, SMALL)); RB2_CUDA(cudaMallocHost(&hRecv, SMALL * RB2_MAX_RANKS));
		RB2_CUDA(cudaMalloc(&dSend, SMALL)); RB2_CUDA(cudaMalloc(&dRecv, SMALL * RB2_MAX_RANKS));
		RB2_CUDA(cudaMalloc(&dBar, 16 * (RB2_MAX_RANKS + 1))); RB2_CUDA(cudaMemset(dBar, 0, 16 * (RB2_MAX_RANKS + 1)));
	}
	~NcclComm() override {
		api->CommDestroy(comm);
		cudaFreeHost(hSend); cudaFreeHost(hRecv); cudaFree(dSend); cudaFree(dRecv); cudaFree(dBar);
	}
	void allgather_host(const void *send, size_t bytes, void *recv, cudaStream_t st) override {
		if (bytes > SMALL) RB2_FATAL("NcclComm: small all-gather of %zu bytes", bytes);
		memcpy(hSend, send, bytes);
		RB2_CUDA(cudaMemcpyAsync(dSend, hSend, bytes, cudaMemcpyHostToDevice, st));
		RB2_NCCL(api->AllGather(dSend, dRecv, bytes, ncclUint8, comm, st));
		RB2_CUDA(cudaMemcpyAsync(hRecv, dRecv, bytes * n, cudaMemcpyDeviceToHost, st));
		RB2_CUDA(cudaStreamSynchronize(st));
		memcpy(recv, hRecv, bytes * n);
	}
	void group_begin() override { RB2_NCCL(api->GroupStart()); }
	void group_end(cudaStream_t) override { RB2_NCCL(api->GroupEnd()); }
	void exchange(const void *sendBuf, void *recvBuf, size_t esz, const Piece *pc, int npc, cudaStream_t st) override {
		// both ends of a pair walk the piece list in the same order, so sends and receives match up
		for (int k = 0; k < npc; ++k) {
			const Piece &p = pc[k];
			if (!p.n) continue;
			const uint8_t *s = (const uint8_t*)sendBuf + p.so * esz; uint8_t *d = (uint8_t*)recvBuf + p.dof * esz;
			if (p.src == rank && p.dst == rank) RB2_CUDA(cudaMemcpyAsync(d, s, p.n * esz, cudaMemcpyDeviceToDevice, st));
			else if (p.src == rank) RB2_NCCL(api->Send(s, p.n * esz, ncclUint8, p.dst, comm, st));
			else if (p.dst == rank) RB2_NCCL(api->Recv(d, p.n * esz, ncclUint8, p.src, comm, st));
		}
	}
	void gather_blocks(uint8_t *const *dst, const size_t *bytes, cudaStream_t st) override {
		RB2_NCCL(api->GroupStart());
		for (int r = 0; r < n; ++r) if (bytes[r]) RB2_NCCL(api->Broadcast(dst[r], dst[r], bytes[r], ncclUint8, r, comm, st));
		RB2_NCCL(api->GroupEnd());
	}
	void barrier(cudaStream_t st) override { uint32_t x = 0, all[RB2_MAX_RANKS]; allgather_host(&x, 4, all, st); }
	// CUDA IPC: the buffer must be a whole cudaMalloc allocation.  The handles travel through the small all-gather.
	bool p2p_map(void *mine, void **peers, cudaStream_t st) override {
		struct Msg { cudaIpcMemHandle_t h; int ok; } me_, all[RB2_MAX_RANKS];
		memset(&me_, 0, sizeof(me_));
		me_.ok = cudaIpcGetMemHandle(&me_.h, mine) == cudaSuccess;
		cudaGetLastError();
		allgather_host(&me_, sizeof(me_), all, st);
		int ok = 1, oks[RB2_MAX_RANKS];
		for (int r = 0; r < n; ++r) { peers[r] = 0; ok = ok && all[r].ok; }
		for (int r = 0; r < n && ok; ++r) {
			if (r == rank) { peers[r] = mine; continue; }
			if (cudaIpcOpenMemHandle(&peers[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { peers[r] = 0; ok = 0; }
		}
		cudaGetLastError();
		allgather_host(&ok, sizeof(ok), oks, st);
		for (int r = 0; r < n; ++r) ok = ok && oks[r];
		if (!ok) p2p_unmap(peers);
		return ok != 0;
	}
	void p2p_unmap(void **peers) override {
		for (int r = 0; r < n; ++r) { if (r != rank && peers[r]) cudaIpcCloseMemHandle(peers[r]); peers[r] = 0; }
		cudaGetLastError();
	}
	// a 16-byte all-gather: its kernel on one rank cannot finish before every other rank's has started, i.e. before
	// everything they queued in front of it (their merge kernels and the peer stores those issued) has completed
	void barrier_stream(cudaStream_t st) override { RB2_NCCL(api->AllGather(dBar, dBar + 16, 16, ncclUint8, comm, st)); }
};
