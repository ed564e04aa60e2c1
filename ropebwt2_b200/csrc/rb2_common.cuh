// rb2_common.cuh -- shared device helpers: error handling, warp/CTA scans, and the
// three-phase (reduce / mid / apply) multi-counter prefix sum every column kernel uses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define RB2_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "[ropebwt2_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
	abort(); } } while (0)

#define RB2_FATAL(...) do { fprintf(stderr, "[ropebwt2_b200] fatal: " __VA_ARGS__); fputc('\n', stderr); abort(); } while (0)

#define FULLMASK 0xffffffffu

// Three constructs have no plain-C++ spelling.  The product build (nvcc, sm_100a) uses the CUDA forms;
// RB2_EMU is defined only by the CPU kernel-logic emulator of the test suite (tests/emu/cuda_emu.h).
#ifdef RB2_EMU
#define RB2_DYN_SMEM(name) uint8_t *name = RB2_EMU_DYN_SMEM
#define RB2_NAMED_BAR(id, nthreads) rb2emu::named_barrier((id), (nthreads))
#define RB2_KERNEL_LAUNCH(kernel, grid, block, smem, stream, ...) rb2emu::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); }, #kernel)
#else
#define RB2_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#define RB2_NAMED_BAR(id, nthreads) asm volatile("bar.sync %0, %1;" :: "n"(id), "n"(nthreads) : "memory")
#define RB2_KERNEL_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// ---- TMA (1-D bulk copies, cp.async.bulk) + mbarrier ------------------------------------------------
// One elected thread arms the mbarrier with the byte count and issues the copy; whoever needs the
// data waits on the barrier's phase parity.  Addresses and sizes are multiples of 16 bytes.
#ifdef RB2_EMU
// emulated mbarrier: count:8 | pending arrivals:8 | phase:16 | pending transaction bytes:32 (signed)
static inline void rb2_emu_check16(const void *a, const void *b, uint32_t n) { if ((((uintptr_t)a | (uintptr_t)b | n) & 15) != 0) { fprintf(stderr, "[cuda_emu] bulk copy not 16-byte aligned\n"); abort(); } }
static inline void rb2_emu_mbar_update(uint64_t *bar, int arrive, int64_t tx)
{
	uint64_t v = *bar;
	uint32_t count = (uint32_t)(v & 0xff), pend = (uint32_t)((v >> 8) & 0xff), phase = (uint32_t)((v >> 16) & 0xffff);
	int64_t t = (int32_t)(uint32_t)(v >> 32);
	if (arrive) { if (pend == 0) { fprintf(stderr, "[cuda_emu] mbarrier: more arrivals than its count\n"); abort(); } --pend; }
	t += tx;
	if (pend == 0 && t == 0) { ++phase; pend = count; }
	*bar = (uint64_t)count | (uint64_t)pend << 8 | (uint64_t)(phase & 0xffff) << 16 | (uint64_t)(uint32_t)(int32_t)t << 32;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { *bar = (uint64_t)count | (uint64_t)count << 8; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { rb2_emu_mbar_update(bar, 1, bytes); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { rb2_emu_mbar_update(bar, 1, 0); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { rb2_emu_check16(dst, src, bytes); memcpy(dst, src, bytes); rb2_emu_mbar_update(bar, 0, -(int64_t)bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while ((((uint32_t)(*bar >> 16)) & 1u) == parity) rb2emu::yield(); }
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) { rb2_emu_check16(dst, src, bytes); memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_read() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ uint32_t warp_redux_add(uint32_t v) { return __reduce_add_sync(FULLMASK, v); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// global -> shared; completes `bytes` transaction bytes on the mbarrier when the data has landed
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while (!ok);
}
// shared -> global (bulk group); the source must stay untouched until bulk_wait_read()
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (the bulk store that follows)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t warp_redux_add(uint32_t v) { return __reduce_add_sync(FULLMASK, v); }
#endif

// device-side error codes (Ctl::err)
enum { RB2_ERR_NONE = 0, RB2_ERR_POOL = 1, RB2_ERR_RUN8 = 2, RB2_ERR_STAGE = 4, RB2_ERR_ORDER = 8, RB2_ERR_PIECES = 16 };

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane)
{
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		T y = __shfl_up_sync(FULLMASK, v, o);
		if (lane >= o) v += y;
	}
	return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
	return v;
}

// Exclusive scan of K counters across a CTA of NT threads (NT multiple of 32, <= 1024).
// v[] in: this thread's values; out: exclusive prefix inside the CTA.  tot[]: CTA totals
// (valid in every thread).  smem must hold K*(NT/32) elements of T.
template <int K, int NT, typename T>
__device__ __forceinline__ void cta_excl_scan(T (&v)[K], T (&tot)[K], T *smem)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	constexpr int NW = NT / 32;
	T incl[K];
#pragma unroll
	for (int k = 0; k < K; ++k) {
		incl[k] = warp_incl_scan(v[k], lane);
		if (lane == 31) smem[k * NW + wid] = incl[k];
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < K; ++k) {
		T base = 0, t = 0;
#pragma unroll
		for (int w = 0; w < NW; ++w) {
			T x = smem[k * NW + w];
			if (w < wid) base += x;
			t += x;
		}
		v[k] = base + incl[k] - v[k];
		tot[k] = t;
	}
	__syncthreads();
}

// ---------------------------------------------------------------------------------
// Three-phase prefix sum over n elements, K counters of type T per element.
//   F::load(i, v[K])                 produce element i's counters (may recompute)
//   F::store(i, v[K], pre[K])        consume element i's counters + exclusive prefix
// scan_reduce -> per-CTA totals; scan_mid -> exclusive scan of the CTA totals (one CTA)
// and the grand totals; scan_apply -> per-element exclusive prefix.
// ---------------------------------------------------------------------------------
#define SCAN_NT 256

template <int K, typename T, class F>
__global__ void __launch_bounds__(SCAN_NT) scan_reduce(F f, uint64_t n, T *ctaTot)
{
	__shared__ T sm[K * (SCAN_NT / 32)];
	uint64_t i = (uint64_t)blockIdx.x * SCAN_NT + threadIdx.x;
	T v[K], tot[K];
#pragma unroll
	for (int k = 0; k < K; ++k) v[k] = 0;
	if (i < n) f.load(i, v);
	cta_excl_scan<K, SCAN_NT, T>(v, tot, sm);
	if (threadIdx.x == 0) {
#pragma unroll
		for (int k = 0; k < K; ++k) ctaTot[(uint64_t)blockIdx.x * K + k] = tot[k];
	}
}

// one CTA of 1024 threads: in-place exclusive scan of ctaTot[nCta][K]; grand[K] = totals
template <int K, typename T>
__global__ void __launch_bounds__(1024) scan_mid(T *ctaTot, uint64_t nCta, T *grand)
{
	__shared__ T sm[K * 32];
	const uint64_t per = (nCta + 1023) / 1024;
	const uint64_t lo = (uint64_t)threadIdx.x * per;
	const uint64_t hi = lo + per < nCta ? lo + per : nCta;
	T v[K], tot[K];
#pragma unroll
	for (int k = 0; k < K; ++k) v[k] = 0;
	for (uint64_t i = lo; i < hi; ++i)
#pragma unroll
		for (int k = 0; k < K; ++k) v[k] += ctaTot[i * K + k];
	cta_excl_scan<K, 1024, T>(v, tot, sm);
	for (uint64_t i = lo; i < hi; ++i)
#pragma unroll
		for (int k = 0; k < K; ++k) {
			T x = ctaTot[i * K + k];
			ctaTot[i * K + k] = v[k];
			v[k] += x;
		}
	if (threadIdx.x == 0) {
#pragma unroll
		for (int k = 0; k < K; ++k) grand[k] = tot[k];
	}
}

// Two-level variant of scan_mid for long arrays (one CTA would serialise ~n/1024 rows per thread):
// mid_reduce sums chunks of MID_ROWS rows, scan_mid scans the chunk sums, mid_apply scans inside
// each chunk.  Rows are K consecutive counters.
#define MID_ROWS 1024
template <int K, typename T>
__global__ void __launch_bounds__(256) mid_reduce(const T *rows, uint64_t n, T *chunkTot)
{
	__shared__ T sm[K * 8];
	const uint64_t r0 = (uint64_t)blockIdx.x * MID_ROWS + threadIdx.x * 4;
	T v[K], tot[K];
#pragma unroll
	for (int k = 0; k < K; ++k) v[k] = 0;
	for (int j = 0; j < 4; ++j) if (r0 + j < n) {
#pragma unroll
		for (int k = 0; k < K; ++k) v[k] += rows[(r0 + j) * K + k];
	}
	cta_excl_scan<K, 256, T>(v, tot, sm);
	if (threadIdx.x == 0) {
#pragma unroll
		for (int k = 0; k < K; ++k) chunkTot[(uint64_t)blockIdx.x * K + k] = tot[k];
	}
}

template <int K, typename T>
__global__ void __launch_bounds__(256) mid_apply(T *rows, uint64_t n, const T *chunkPre)
{
	__shared__ T sm[K * 8];
	const uint64_t r0 = (uint64_t)blockIdx.x * MID_ROWS + threadIdx.x * 4;
	T v[K], tot[K], own[4][K];
#pragma unroll
	for (int k = 0; k < K; ++k) v[k] = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
#pragma unroll
		for (int k = 0; k < K; ++k) { own[j][k] = r0 + j < n ? rows[(r0 + j) * K + k] : 0; v[k] += own[j][k]; }
	}
	cta_excl_scan<K, 256, T>(v, tot, sm);
#pragma unroll
	for (int k = 0; k < K; ++k) v[k] += chunkPre[(uint64_t)blockIdx.x * K + k];
#pragma unroll
	for (int j = 0; j < 4; ++j) if (r0 + j < n) {
#pragma unroll
		for (int k = 0; k < K; ++k) { rows[(r0 + j) * K + k] = v[k]; v[k] += own[j][k]; }
	}
}

template <int K, typename T, class F>
__global__ void __launch_bounds__(SCAN_NT) scan_apply(F f, uint64_t n, const T *ctaPre)
{
	__shared__ T sm[K * (SCAN_NT / 32)];
	uint64_t i = (uint64_t)blockIdx.x * SCAN_NT + threadIdx.x;
	T v[K], own[K], tot[K];
#pragma unroll
	for (int k = 0; k < K; ++k) v[k] = 0;
	if (i < n) f.load(i, v);
#pragma unroll
	for (int k = 0; k < K; ++k) own[k] = v[k];
	cta_excl_scan<K, SCAN_NT, T>(v, tot, sm);
	if (i < n) {
#pragma unroll
		for (int k = 0; k < K; ++k) v[k] += ctaPre[(uint64_t)blockIdx.x * K + k];
		f.store(i, own, v);
	}
}
