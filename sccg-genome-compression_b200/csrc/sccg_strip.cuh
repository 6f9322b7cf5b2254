// N-stripping by stream compaction + upper-casing.
//   compress side (global): toupper, then erase every 'N'   (compression.cpp:523-524, :556-557)  -> drops 'n' and 'N'
//   decompress side       : erase every 'N', then toupper   (decompression.cpp:105-110)          -> keeps 'n' (becomes 'N')
// HBM-bound: 1 B read + 1 B written per kept symbol (packed in shared memory, stored 16 B at a time), one pass.
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

// out[i] = toupper(src[i]), 16 bytes per thread
__global__ void __launch_bounds__(256) upper_k(const u8* __restrict__ src, i64 n, u8* __restrict__ dst) {
    i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= n) return;
    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + i);
    v.x = upper8(v.x); v.y = upper8(v.y);
    *reinterpret_cast<ulonglong2*>(dst + i) = v;                     // buffers are padded to 16 B
}

// Count, offsets and compaction in ONE pass over the source: a CTA owns STRIP_SUB consecutive 4 KiB blocks, counts what it
// keeps, learns where its output begins by decoupled look-back over the tiles before it (the descriptors and the tile
// counter of sccg_scan.cuh: tiles are numbered in start order, a tile only waits for tiles that are running) and stores.
// 1 B read + 1 B written per symbol instead of 2 B read + 1 B written and a scan launch in between.
static const int STRIP_SUB = 4;
template <int UPPER_FIRST>
__global__ void __launch_bounds__(STRIP_T) strip_onepass_k(const u8* __restrict__ src, i64 n, u8* __restrict__ dst, u64* desc, u32* counter, u32 counter_base,
                                                           u32 epoch, u32* __restrict__ total_out) {
    __shared__ u32 sm[40];
    __shared__ u32 s_tile, s_prefix;
    __align__(16) __shared__ u8 stage[STRIP_TILE + 32];
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u) - counter_base;
    __syncthreads();
    const u32 tile = s_tile;
    const i64 base = (i64)tile * (STRIP_TILE * STRIP_SUB) + (i64)threadIdx.x * 16;
    ulonglong2 v[STRIP_SUB];
    u32 m[STRIP_SUB], excl[STRIP_SUB], sub_tot[STRIP_SUB];
#pragma unroll
    for (int q = 0; q < STRIP_SUB; ++q) { v[q].x = 0; v[q].y = 0; m[q] = strip_keep_mask<UPPER_FIRST>(src, n, base + (i64)q * STRIP_TILE, &v[q]); }
    u32 tot = 0;
#pragma unroll
    for (int q = 0; q < STRIP_SUB; ++q) { excl[q] = block_scan_excl((u32)__popc(m[q]), sm, &sub_tot[q]); tot += sub_tot[q]; }
    if (threadIdx.x < 32) {
        const int lane = lane_of();
        u32 prefix = 0;
        if (tile == 0) {
            if (lane == 0) SCCG_ST_RELAXED_U64(desc, scan_desc(epoch, 1u, tot));
        } else {
            if (lane == 0) SCCG_ST_RELAXED_U64(desc + tile, scan_desc(epoch, 0u, tot));
            for (i64 top = (i64)tile - 1; top >= 0; top -= 32) {
                const i64 idx = top - lane;
                u64 d = scan_desc(epoch, 1u, 0u);
                if (idx >= 0) { do { d = SCCG_LD_RELAXED_U64(desc + idx); } while ((u32)(d >> 33) != epoch); }
                const u32 is_prefix = (u32)(d >> 32) & 1u;
                const u32 bal = __ballot_sync(SCCG_FULL_MASK, is_prefix != 0u);
                const int stop = bal ? __ffs((int)bal) - 1 : 31;
                prefix += __reduce_add_sync(SCCG_FULL_MASK, lane <= stop ? (u32)d : 0u);
                if (bal) break;
            }
            if (lane == 0) SCCG_ST_RELAXED_U64(desc + tile, scan_desc(epoch, 1u, prefix + tot));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if ((i64)(tile + 1) * (STRIP_TILE * STRIP_SUB) >= n) *total_out = prefix + tot;
        }
    }
    __syncthreads();
    u32 off = s_prefix;
#pragma unroll
    for (int q = 0; q < STRIP_SUB; ++q) {
        block_compact_store(stage, dst + off, excl[q], m[q], upper8(v[q].x), upper8(v[q].y), sub_tot[q]);
        off += sub_tot[q];
        __syncthreads();                                              // the staging area is refilled by the next block
    }
}

// dst <- N-stripped, upper-cased src on the current lane; *d_count (device) receives the kept length.  No host round trip.
template <int UPPER_FIRST>
static int strip_n_enqueue(sccg_ctx* c, const u8* d_src, i64 n, u8* d_dst, int slot_cnt, u32* d_count) {
    (void)slot_cnt;
    if (n <= 0) { SCCG_CK(cudaMemsetAsync(d_count, 0, sizeof(u32), c->stream)); return SCCG_OK; }
    const unsigned ntiles = div_up(n, STRIP_TILE * STRIP_SUB);
    u64* desc = nullptr; u32* counter = nullptr; int ln = 0;
    SCCG_TRY(scan_state(c, (size_t)ntiles, &desc, &counter, &ln));
    LAUNCH(c, strip_onepass_k<UPPER_FIRST>, dim3(ntiles), dim3(STRIP_T), 0, d_src, n, d_dst, desc, counter, c->scan_counter_base[ln], c->scan_epoch[ln], d_count);
    c->scan_counter_base[ln] += ntiles;                          // modulo 2^32, like the device counter
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
