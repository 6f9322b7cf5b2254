// =================================================================================================
// TEST INFRASTRUCTURE ONLY -- a small SIMT emulator so that the CUDA kernels of this repository can
// be unit-tested in the CPU-only build container (there is no GPU there; gpurun time is scarce).
//
// The product library (libsccg_b200.so) is compiled by nvcc for sm_100a and never sees this file.
// tests/emu/build_emu.sh compiles the SAME kernel sources with g++ -DSCCG_EMU into
// tests/emu/libsccg_b200_emu.so, where every CUDA thread is a fiber, warp collectives
// (__shfl_*_sync, __ballot_sync, __syncwarp, __reduce_*_sync) rendezvous between the fibers of a warp,
// __syncthreads() between the fibers of a block, and blocks run on a pool of OS threads.  It checks
// kernel LOGIC (indexing, parse rules, scans); it is not a performance model and not a fallback.
// =================================================================================================
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <algorithm>

// ------------------------------------------------------------------------------------------------
// CUDA keywords
// ------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static const
#define __align__(n) alignas(n)
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ------------------------------------------------------------------------------------------------
// emulator core (simt_emu.cpp)
// ------------------------------------------------------------------------------------------------
namespace emu {
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
    uint3 tid{0, 0, 0};
    int linear = 0;   // linear thread id in block
};
struct Slot {          // one in-flight warp collective
    uint32_t mask = 0, arrived = 0, read_pending = 0;
    bool complete = false;
    uint64_t vals[32];
};
struct BlockCtx {
    dim3 grid, block;
    uint3 bidx{0, 0, 0};
    int nthreads = 0, alive = 0;
    Fiber* fibers = nullptr;
    Fiber* cur = nullptr;
    void* sched_sp = nullptr;
    int bar_count = 0;
    unsigned bar_gen = 0;
    int bar_red[2] = {0, 0};      // __syncthreads_count / _or / _and accumulators, indexed by barrier generation parity
    Slot slots[32][4];
    unsigned long long progress = 0;
    unsigned char* dyn_smem = nullptr;
    const std::function<void()>* body = nullptr;
};
extern thread_local BlockCtx* g_blk;

void yield();
extern const char* g_kernel_name;      // diagnostics: the kernel of the current launch
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
Slot* collective_arrive(uint32_t mask, uint64_t val);   // blocks until every lane of mask arrived
void collective_release(Slot* s);
void block_barrier();
int block_barrier_count(int pred);
[[noreturn]] void fail(const char* msg);

inline int lane_id() { return g_blk->cur->linear & 31; }
inline int warp_id() { return g_blk->cur->linear >> 5; }
inline unsigned char* dyn_smem() { return g_blk->dyn_smem; }

template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace emu

#define threadIdx (emu::g_blk->cur->tid)
#define blockIdx (emu::g_blk->bidx)
#define blockDim (emu::g_blk->block)
#define gridDim (emu::g_blk->grid)
#define warpSize 32

// ------------------------------------------------------------------------------------------------
// synchronisation + warp collectives
// ------------------------------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(); }
inline int __syncthreads_count(int pred) { return emu::block_barrier_count(pred); }
inline int __syncthreads_or(int pred) { return emu::block_barrier_count(pred) != 0; }
inline int __syncthreads_and(int pred) { return emu::block_barrier_count(!pred) == 0; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::Slot* s = emu::collective_arrive(mask, 0); emu::collective_release(s); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) {}
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    int lane = emu::lane_id();
    int base = lane & ~(width - 1);
    int from = base + (src & (width - 1));
    T r = ((mask >> from) & 1u) ? emu::from_bits<T>(s->vals[from]) : v;
    emu::collective_release(s);
    return r;
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    int lane = emu::lane_id();
    int base = lane & ~(width - 1);
    int from = lane - (int)delta;
    T r = (from >= base && ((mask >> from) & 1u)) ? emu::from_bits<T>(s->vals[from]) : v;
    emu::collective_release(s);
    return r;
}
template <typename T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    int lane = emu::lane_id();
    int base = lane & ~(width - 1);
    int from = lane + (int)delta;
    T r = (from < base + width && ((mask >> from) & 1u)) ? emu::from_bits<T>(s->vals[from]) : v;
    emu::collective_release(s);
    return r;
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    int lane = emu::lane_id();
    int from = lane ^ lanemask;
    (void)width;
    T r = (from < 32 && ((mask >> from) & 1u)) ? emu::from_bits<T>(s->vals[from]) : v;
    emu::collective_release(s);
    return r;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::Slot* s = emu::collective_arrive(mask, pred ? 1 : 0);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((mask >> i) & 1u) && s->vals[i]) r |= 1u << i;
    emu::collective_release(s);
    return r;
}
template <typename T> inline unsigned __match_any_sync(unsigned mask, T v) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    uint64_t mine = emu::to_bits(v);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((mask >> i) & 1u) && s->vals[i] == mine) r |= 1u << i;
    emu::collective_release(s);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline unsigned __activemask() { return 0xffffffffu; }

#define EMU_REDUCE(NAME, T, INIT, OP)                                              \
    inline T NAME(unsigned mask, T v) {                                            \
        emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));              \
        T r = INIT;                                                                \
        for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) { T x = emu::from_bits<T>(s->vals[i]); r = OP; } \
        emu::collective_release(s);                                                \
        return r;                                                                  \
    }
EMU_REDUCE(__reduce_add_sync, unsigned, 0u, r + x)
EMU_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (x < r ? x : r))
EMU_REDUCE(__reduce_max_sync, unsigned, 0u, (x > r ? x : r))
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, r | x)
EMU_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, r & x)
inline int __reduce_add_sync(unsigned mask, int v) { return (int)__reduce_add_sync(mask, (unsigned)v); }
inline int __reduce_min_sync(unsigned mask, int v) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    int r = 0x7fffffff;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r = std::min(r, emu::from_bits<int>(s->vals[i]));
    emu::collective_release(s);
    return r;
}
inline int __reduce_max_sync(unsigned mask, int v) {
    emu::Slot* s = emu::collective_arrive(mask, emu::to_bits(v));
    int r = (int)0x80000000;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r = std::max(r, emu::from_bits<int>(s->vals[i]));
    emu::collective_release(s);
    return r;
}

// ------------------------------------------------------------------------------------------------
// atomics (blocks run concurrently on OS threads -> real atomics)
// ------------------------------------------------------------------------------------------------
template <typename T> inline T atomicAdd(T* a, T v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicSub(T* a, T v) { return __atomic_fetch_sub(a, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicOr(T* a, T v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicAnd(T* a, T v) { return __atomic_fetch_and(a, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicExch(T* a, T v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicCAS(T* a, T cmp, T v) { __atomic_compare_exchange_n(a, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template <typename T> inline T atomicMax(T* a, T v) { T o = __atomic_load_n(a, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <typename T> inline T atomicMin(T* a, T v) { T o = __atomic_load_n(a, __ATOMIC_SEQ_CST); while (o > v && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }

// ------------------------------------------------------------------------------------------------
// intrinsics
// ------------------------------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (sh & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (sh & 31)) >> 32); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) { unsigned s = (sel >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * s)) & 0xff) << (8 * i); }
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
using std::min;
using std::max;

// ------------------------------------------------------------------------------------------------
// host runtime subset
// ------------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef struct emu_stream_t* cudaStream_t;
typedef struct emu_event_t* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaHostAllocDefault = 0 };
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t n);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaDeviceGetAttribute(int* v, int attr, int dev);
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
