// cuda_emu.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A small CPU emulator of the CUDA execution model, just large enough to run the kernels of
// ropebwt2_b200/csrc/*.cu UNCHANGED on a machine without a GPU (this container): the `-m "not gpu"`
// tests compile the engine with `g++ -x c++ -DRB2_EMU -include cuda_emu.h` into
// tests/emu/_build/libropebwt2_b200_emu.so and check its kernels' LOGIC against the oracle.  The
// product library (ropebwt2_b200/_build/libropebwt2_b200.so, nvcc, sm_100a) never sees this file,
// and nothing under ropebwt2_b200/ loads the emulated library: it is reachable only through the
// explicit path tests/ passes to the binding.  No performance claim is ever made from it.
//
// Model: one CTA at a time per host thread; every CUDA thread of the CTA is a fibre (own stack, a
// ~20-instruction context switch); __syncthreads / named barriers / warp collectives (__shfl*_sync,
// __ballot_sync, __syncwarp, __reduce_add_sync) and mbarrier waits yield to a round-robin scheduler.
// Threads that have exited count as arrived, as on the hardware.  A round in which no fibre makes
// progress aborts with "deadlock" (a barrier that not every thread reaches).  Device memory is host
// memory; streams and events are no-ops (everything is synchronous); TMA bulk copies are memcpy.
// Data races are NOT detected (fibres never run concurrently) -- compute-sanitizer on the GPU box does that.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <vector>
#include <algorithm>
#include <deque>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <atomic>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__
#define __constant__ static

struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = { x, y, z, w }; return v; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v = { x, y }; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = { x, y, z, w }; return v; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 v = { x, y }; return v; }
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

namespace rb2emu {

struct Bar { unsigned gen = 0, count = 0; };
struct Warp { // one barrier per distinct participation mask (sub-warp groups synchronise independently)
	Bar bar[8]; unsigned mask[8]; unsigned nmask = 0; uint64_t slot[32];
	Bar &get(unsigned m) {
		for (unsigned i = 0; i < nmask; ++i) if (mask[i] == m) return bar[i];
		if (nmask == 8) { fprintf(stderr, "[cuda_emu] more than 8 distinct warp masks in one CTA\n"); abort(); }
		mask[nmask] = m; bar[nmask] = Bar(); return bar[nmask++];
	}
};

struct Fibre {
	void *sp = 0; char *stack = 0; bool done = false, started = false;
	unsigned tid = 0;
};

struct Cta {
	unsigned nthreads = 0, live = 0;
	std::vector<Fibre> fib;
	std::vector<Warp> warp;
	Bar bars[16];
	unsigned progress = 0;   // bumped whenever a barrier releases or a fibre finishes
	const std::function<void()> *body = 0;
	dim3 bidx, bdim, gdim;
	void *sched_sp = 0;
	Fibre *cur = 0;
	unsigned char *dynsmem = 0;
};

extern thread_local Cta *cta;            // the CTA running on this host thread
extern thread_local dim3 cur_threadIdx;

void yield();                             // fibre -> scheduler
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body, const char *name = "?");
unsigned live_in_mask(unsigned warp, unsigned mask);

inline void bar_wait(Bar &b, unsigned expected_fn(void*), void *arg)
{
	const unsigned my = b.gen;
	++b.count;
	for (;;) {
		if (b.gen != my) return;
		if (b.count >= expected_fn(arg)) { b.count = 0; ++b.gen; ++cta->progress; return; }
		yield();
	}
}

inline unsigned exp_cta(void*) { return cta->live; }
struct NamedArg { unsigned n; };
inline unsigned exp_named(void *a) { return ((NamedArg*)a)->n; }
struct WarpArg { unsigned w, mask; };
inline unsigned exp_warp(void *a) { WarpArg *x = (WarpArg*)a; return live_in_mask(x->w, x->mask); }

inline void syncthreads() { bar_wait(cta->bars[0], exp_cta, 0); }
inline void named_barrier(int id, unsigned nthreads) { NamedArg a = { nthreads }; bar_wait(cta->bars[id], exp_named, &a); }
inline void syncwarp(unsigned mask) { WarpArg a = { cur_threadIdx.x >> 5, mask }; bar_wait(cta->warp[a.w].get(mask), exp_warp, &a); }

template <typename T> inline T exchange(unsigned mask, T v, unsigned src_lane)
{
	static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
	Warp &w = cta->warp[cur_threadIdx.x >> 5];
	uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
	w.slot[cur_threadIdx.x & 31] = raw;
	syncwarp(mask);
	raw = w.slot[src_lane & 31];
	syncwarp(mask);
	T r; memcpy(&r, &raw, sizeof(T));
	return r;
}

} // namespace rb2emu

#define threadIdx (rb2emu::cur_threadIdx)
#define blockIdx  (rb2emu::cta->bidx)
#define blockDim  (rb2emu::cta->bdim)
#define gridDim   (rb2emu::cta->gdim)
#define warpSize  32

static inline void __syncthreads() { rb2emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { rb2emu::syncwarp(mask); }
static inline int __syncthreads_or(int p)
{
	static thread_local int acc;
	rb2emu::syncthreads(); // (readers of the previous call are done)
	acc = 0;
	rb2emu::syncthreads();
	if (p) acc = 1;
	rb2emu::syncthreads();
	const int r = acc;
	rb2emu::syncthreads();
	return r;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	const unsigned lane = threadIdx.x & 31;
	const unsigned s = width == 32 ? (unsigned)src & 31 : (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (width - 1));
	return rb2emu::exchange(mask, v, s);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	const unsigned lane = threadIdx.x & 31, base = lane & ~(unsigned)(width - 1);
	const unsigned s = lane >= base + delta ? lane - delta : lane;
	return rb2emu::exchange(mask, v, s);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	const unsigned lane = threadIdx.x & 31, base = lane & ~(unsigned)(width - 1);
	const unsigned s = lane + delta < base + width ? lane + delta : lane;
	return rb2emu::exchange(mask, v, s);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
	const unsigned lane = threadIdx.x & 31;
	(void)width;
	return rb2emu::exchange(mask, v, lane ^ (unsigned)x);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	rb2emu::Warp &w = rb2emu::cta->warp[threadIdx.x >> 5];
	const unsigned lane = threadIdx.x & 31;
	w.slot[lane] = pred ? 1 : 0;
	rb2emu::syncwarp(mask);
	unsigned r = 0;
	for (unsigned l = 0; l < 32; ++l) if ((mask >> l & 1) && (threadIdx.x - lane + l) < rb2emu::cta->nthreads && !rb2emu::cta->fib[threadIdx.x - lane + l].done && w.slot[l]) r |= 1u << l;
	rb2emu::syncwarp(mask);
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
	rb2emu::Warp &w = rb2emu::cta->warp[threadIdx.x >> 5];
	const unsigned lane = threadIdx.x & 31;
	w.slot[lane] = v;
	rb2emu::syncwarp(mask);
	unsigned r = 0;
	for (unsigned l = 0; l < 32; ++l) if ((mask >> l & 1) && (threadIdx.x - lane + l) < rb2emu::cta->nthreads && !rb2emu::cta->fib[threadIdx.x - lane + l].done) r += (unsigned)w.slot[l];
	rb2emu::syncwarp(mask);
	return r;
}

// ---- integer intrinsics --------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (x >> i & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (sh & 31)) >> 32); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) { sh = sh > 32 ? 32 : sh; return sh == 32 ? hi : (unsigned)((((uint64_t)hi << 32) | lo) >> sh); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh) { sh = sh > 32 ? 32 : sh; return sh == 32 ? lo : (unsigned)(((((uint64_t)hi << 32) | lo) << sh) >> 32); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
	const uint64_t v = ((uint64_t)y << 32) | x;
	unsigned r = 0;
	for (int i = 0; i < 4; ++i) {
		const unsigned sel = (s >> (4 * i)) & 0xf;
		unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
		if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;
		r |= b << (8 * i);
	}
	return r;
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b)
{
	unsigned r = 0;
	for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
	return r;
}
static inline unsigned __vcmpgtu4(unsigned a, unsigned b)
{
	unsigned r = 0;
	for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xff) > ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
	return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
	for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
	return c;
}
static inline int __dp4a(int a, int b, int c)
{
	for (int i = 0; i < 4; ++i) c += (int)(int8_t)((unsigned)a >> (8 * i)) * (int)(int8_t)((unsigned)b >> (8 * i));
	return c;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename A, typename B> static inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <typename A, typename B> static inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }

// ---- atomics (fibres of one host thread never run concurrently; several host threads = several
// "devices", which only share memory through explicit peer copies) -------------------------------
template <typename T, typename U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U> static inline T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <typename T, typename U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U> static inline T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename U, typename V> static inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { rb2emu::yield(); }

// ---- runtime API subset ------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct rb2emu_stream *cudaStream_t;
typedef struct rb2emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; char name[64]; };

static inline const char *cudaGetErrorName(cudaError_t) { return "cudaErrorEmulated"; }
static inline const char *cudaGetErrorString(cudaError_t) { return "error in the CUDA emulator"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline int rb2emu_env(const char *k, int d) { const char *s = getenv(k); return s && *s ? atoi(s) : d; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = rb2emu_env("RB2_EMU_DEVICES", 8); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); p->multiProcessorCount = rb2emu_env("RB2_EMU_SMS", 2); p->totalGlobalMem = (size_t)64 << 30; strcpy(p->name, "CUDA emulator (CPU)"); return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)rb2emu_env("RB2_EMU_FREE_MB", 16384) << 20; *t = (size_t)64 << 30; return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { void *q = 0; if (posix_memalign(&q, 256, n ? n : 1)) return cudaErrorMemoryAllocation; if (rb2emu_env("RB2_EMU_POISON", 1)) memset(q, 0xA5, n < ((size_t)64 << 20) ? n : ((size_t)64 << 20)); *p = (T*)q; return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { void *q = 0; if (posix_memalign(&q, 256, n ? n : 1)) return cudaErrorMemoryAllocation; *p = (T*)q; return cudaSuccess; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
// (CUDA IPC is what NcclComm maps peer buffers with; ranks of the emulator are threads, which use LocalComm)
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorMemoryAllocation; }
static inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorMemoryAllocation; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }

// dynamic shared memory of the running CTA
#define RB2_EMU_DYN_SMEM (rb2emu::cta->dynsmem)
