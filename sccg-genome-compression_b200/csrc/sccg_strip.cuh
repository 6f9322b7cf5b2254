// N-stripping by stream compaction + upper-casing.
//   compress side (global): toupper, then erase every 'N'   (compression.cpp:523-524, :556-557)  -> drops 'n' and 'N'
//   decompress side       : erase every 'N', then toupper   (decompression.cpp:105-110)          -> keeps 'n' (becomes 'N')
// HBM-bound: 1 B read per pass (two passes) + 1 B written per kept symbol (packed in shared memory, stored 16 B at a time).
#pragma once
#include "sccg_scan.cuh"

namespace sccg {

static const int STRIP_T = 256;
static const int STRIP_TILE = STRIP_T * 16;

// bit j set <=> byte j of the 16 bytes at src+i is kept
template <int UPPER_FIRST> __device__ __forceinline__ u32 strip_keep_mask(const u8* __restrict__ src, i64 n, i64 i, ulonglong2* v) {
    if (i >= n) return 0u;
    *v = *reinterpret_cast<const ulonglong2*>(src + i);
    u64 nx = eq_flags8(v->x, 'N'), ny = eq_flags8(v->y, 'N');
    if (UPPER_FIRST) { nx |= eq_flags8(v->x, 'n'); ny |= eq_flags8(v->y, 'n'); }
    u32 drop = movemask8(nx) | (movemask8(ny) << 8);
    u32 m = ~drop & 0xffffu;
    i64 left = n - i;
    if (left < 16) m &= (1u << (int)left) - 1u;
    return m;
}

template <int UPPER_FIRST>
__global__ void __launch_bounds__(STRIP_T) strip_count_k(const u8* __restrict__ src, i64 n, u32* __restrict__ cnt) {
    __shared__ u32 sm[40];
    i64 i = (i64)blockIdx.x * STRIP_TILE + (i64)threadIdx.x * 16;
    ulonglong2 v;
    u32 m = strip_keep_mask<UPPER_FIRST>(src, n, i, &v);
    u32 tot;
    block_scan_excl((u32)__popc(m), sm, &tot);
    if (threadIdx.x == 0) cnt[blockIdx.x] = tot;
}

template <int UPPER_FIRST>
__global__ void __launch_bounds__(STRIP_T) strip_write_k(const u8* __restrict__ src, i64 n, const u32* __restrict__ tile_off, u8* __restrict__ dst) {
    __shared__ u32 sm[40];
    __align__(16) __shared__ u8 stage[STRIP_TILE + 32];
    i64 i = (i64)blockIdx.x * STRIP_TILE + (i64)threadIdx.x * 16;
    ulonglong2 v;
    v.x = 0; v.y = 0;
    u32 m = strip_keep_mask<UPPER_FIRST>(src, n, i, &v);
    u32 tot;
    u32 excl = block_scan_excl((u32)__popc(m), sm, &tot);
    block_compact_store(stage, dst + tile_off[blockIdx.x], excl, m, upper8(v.x), upper8(v.y), tot);
}

// out[i] = toupper(src[i]), 16 bytes per thread
__global__ void __launch_bounds__(256) upper_k(const u8* __restrict__ src, i64 n, u8* __restrict__ dst) {
    i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= n) return;
    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + i);
    v.x = upper8(v.x); v.y = upper8(v.y);
    *reinterpret_cast<ulonglong2*>(dst + i) = v;                     // buffers are padded to 16 B
}

// dst <- N-stripped, upper-cased src on the current lane; *d_count (device) receives the kept length.  No host round trip.
template <int UPPER_FIRST>
static int strip_n_enqueue(sccg_ctx* c, const u8* d_src, i64 n, u8* d_dst, int slot_cnt, u32* d_count) {
    if (n <= 0) { SCCG_CK(cudaMemsetAsync(d_count, 0, sizeof(u32), c->stream)); return SCCG_OK; }
    unsigned ntiles = div_up(n, STRIP_TILE);
    u32* cnt = nullptr;
    SCCG_TRY(buf(c, slot_cnt, (size_t)ntiles + 1, &cnt));
    LAUNCH(c, strip_count_k<UPPER_FIRST>, dim3(ntiles), dim3(STRIP_T), 0, d_src, n, cnt);
    SCCG_TRY(scan_exclusive_u32(c, cnt, cnt, (i64)ntiles, d_count));
    LAUNCH(c, strip_write_k<UPPER_FIRST>, dim3(ntiles), dim3(STRIP_T), 0, d_src, n, (const u32*)cnt, d_dst);
    return SCCG_OK;
}

// same; *h_count receives the kept length as well (one host round trip)
template <int UPPER_FIRST>
static int strip_n(sccg_ctx* c, const u8* d_src, i64 n, u8* d_dst, int slot_cnt, u32* d_count, i64* h_count) {
    if (n <= 0) { *h_count = 0; return SCCG_OK; }
    SCCG_TRY(strip_n_enqueue<UPPER_FIRST>(c, d_src, n, d_dst, slot_cnt, d_count));
    SCCG_CK(cudaMemcpyAsync(c->h_pinned, d_count, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    SCCG_CK(cudaStreamSynchronize(c->stream));
    *h_count = (i64)*(u32*)c->h_pinned;
    return SCCG_OK;
}

}  // namespace sccg
