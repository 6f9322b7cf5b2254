// Device-wide exclusive prefix sum over u32 in ONE launch: tiles of 4096 elements, chained through
// tile descriptors with decoupled look-back (every tile publishes its aggregate, then its inclusive
// prefix; a tile sums the aggregates of its predecessors until it meets a published prefix).
// Used for every "count -> offset" step of the record pipeline; HBM-bound, 8 B per element.
//   * tile ids are handed out by an atomic counter in START order, so a tile only ever waits for
//     tiles that are already running: no deadlock for any grid size;
//   * descriptors and the counter are never reset: a descriptor is valid only if it carries the
//     epoch of the current launch, and the host knows the counter value every launch starts from.
#pragma once
#include "sccg_ctx.cuh"

namespace sccg {

static const int SCAN_T = 256;
static const int SCAN_I = 16;
static const int SCAN_TILE = SCAN_T * SCAN_I;

// descriptor: value in the low word, (epoch << 1) | is_prefix in the high word
__device__ __forceinline__ u64 scan_desc(u32 epoch, u32 is_prefix, u32 value) { return ((u64)((epoch << 1) | is_prefix) << 32) | value; }

__global__ void __launch_bounds__(SCAN_T) scan_onepass_k(const u32* in, u32* out, i64 n, u64* desc, u32* counter, u32 counter_base, u32 epoch, u32* total_out) {
    __shared__ u32 sm[40];
    __shared__ u32 s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u) - counter_base;
    __syncthreads();
    const u32 tile = s_tile;
    const i64 base = (i64)tile * SCAN_TILE + (i64)threadIdx.x * SCAN_I;
    u32 v[SCAN_I];
    if (base + SCAN_I <= n && (((uintptr_t)in) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < SCAN_I / 4; ++q) {
            uint4 x = *reinterpret_cast<const uint4*>(in + base + 4 * q);
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_I; ++j) v[j] = (base + j < n) ? in[base + j] : 0u;
    }
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) s += v[j];
    u32 tot;
    const u32 excl = block_scan_excl(s, sm, &tot);
    if (threadIdx.x < 32) {
        const int lane = lane_of();
        u32 prefix = 0;
        if (tile == 0) {
            if (lane == 0) SCCG_ST_RELAXED_U64(desc, scan_desc(epoch, 1u, tot));
        } else {
            if (lane == 0) SCCG_ST_RELAXED_U64(desc + tile, scan_desc(epoch, 0u, tot));
            // look-back: 32 predecessors per round, nearest first
            for (i64 top = (i64)tile - 1; top >= 0; top -= 32) {
                const i64 idx = top - lane;
                u64 d = scan_desc(epoch, 1u, 0u);                          // before tile 0: a prefix of 0
                if (idx >= 0) { do { d = SCCG_LD_RELAXED_U64(desc + idx); } while ((u32)(d >> 33) != epoch); }
                const u32 is_prefix = (u32)(d >> 32) & 1u;
                const u32 bal = __ballot_sync(SCCG_FULL_MASK, is_prefix != 0u);
                const int stop = bal ? __ffs((int)bal) - 1 : 31;           // nearest predecessor that already knows its prefix
                prefix += __reduce_add_sync(SCCG_FULL_MASK, lane <= stop ? (u32)d : 0u);
                if (bal) break;
            }
            if (lane == 0) SCCG_ST_RELAXED_U64(desc + tile, scan_desc(epoch, 1u, prefix + tot));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (total_out && (i64)(tile + 1) * SCAN_TILE >= n) *total_out = prefix + tot;
        }
    }
    __syncthreads();
    u32 off = s_prefix + excl;
    if (base + SCAN_I <= n && (((uintptr_t)out) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < SCAN_I / 4; ++q) {
            uint4 x;
            x.x = off; off += v[4 * q];
            x.y = off; off += v[4 * q + 1];
            x.z = off; off += v[4 * q + 2];
            x.w = off; off += v[4 * q + 3];
            *reinterpret_cast<uint4*>(out + base + 4 * q) = x;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_I; ++j) { if (base + j < n) out[base + j] = off; off += v[j]; }
    }
}

// Three exclusive scans of equally long arrays in ONE launch (the tokenizer's per-chunk counters: symbols, segments,
// tokens): arrays at arr, arr + stride, arr + 2 * stride, scanned in place.  Tiles of 2048 elements per array; warp ch of
// the CTA runs the look-back of channel ch (descriptor 3 * tile + ch), so the three chains advance side by side.
static const int SCAN3_I = 8;
static const int SCAN3_TILE = SCAN_T * SCAN3_I;

__global__ void __launch_bounds__(SCAN_T) scan_onepass3_k(u32* arr, size_t stride, i64 n, u64* desc, u32* counter, u32 counter_base, u32 epoch,
                                                          u32* total0, u32* total1, u32* total2) {
    __shared__ u32 sm[40];
    __shared__ u32 s_tile, s_prefix[3], s_tot[3];
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u) - counter_base;
    __syncthreads();
    const u32 tile = s_tile;
    const i64 base = (i64)tile * SCAN3_TILE + (i64)threadIdx.x * SCAN3_I;
    const bool fast = base + SCAN3_I <= n && (((uintptr_t)arr) & 15) == 0 && (stride & 3) == 0;
    u32 v[3][SCAN3_I], excl[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const u32* in = arr + (size_t)ch * stride;
        if (fast) {
#pragma unroll
            for (int q = 0; q < SCAN3_I / 4; ++q) {
                uint4 x = *reinterpret_cast<const uint4*>(in + base + 4 * q);
                v[ch][4 * q] = x.x; v[ch][4 * q + 1] = x.y; v[ch][4 * q + 2] = x.z; v[ch][4 * q + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SCAN3_I; ++j) v[ch][j] = (base + j < n) ? in[base + j] : 0u;
        }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        u32 s = 0;
#pragma unroll
        for (int j = 0; j < SCAN3_I; ++j) s += v[ch][j];
        u32 tot;
        excl[ch] = block_scan_excl(s, sm, &tot);
        if (threadIdx.x == 0) s_tot[ch] = tot;
    }
    __syncthreads();
    if (threadIdx.x < 96) {
        const int lane = lane_of(), ch = (int)(threadIdx.x >> 5);
        const u32 tot = s_tot[ch];
        u64* dch = desc + ch;                                          // descriptor of (tile t, channel ch): dch[3 * t]
        u32 prefix = 0;
        if (tile == 0) {
            if (lane == 0) SCCG_ST_RELAXED_U64(dch, scan_desc(epoch, 1u, tot));
        } else {
            if (lane == 0) SCCG_ST_RELAXED_U64(dch + 3 * (size_t)tile, scan_desc(epoch, 0u, tot));
            for (i64 top = (i64)tile - 1; top >= 0; top -= 32) {
                const i64 idx = top - lane;
                u64 d = scan_desc(epoch, 1u, 0u);
                if (idx >= 0) { do { d = SCCG_LD_RELAXED_U64(dch + 3 * idx); } while ((u32)(d >> 33) != epoch); }
                const u32 is_prefix = (u32)(d >> 32) & 1u;
                const u32 bal = __ballot_sync(SCCG_FULL_MASK, is_prefix != 0u);
                const int stop = bal ? __ffs((int)bal) - 1 : 31;
                prefix += __reduce_add_sync(SCCG_FULL_MASK, lane <= stop ? (u32)d : 0u);
                if (bal) break;
            }
            if (lane == 0) SCCG_ST_RELAXED_U64(dch + 3 * (size_t)tile, scan_desc(epoch, 1u, prefix + tot));
        }
        if (lane == 0) {
            s_prefix[ch] = prefix;
            u32* total_out = ch == 0 ? total0 : (ch == 1 ? total1 : total2);
            if (total_out && (i64)(tile + 1) * SCAN3_TILE >= n) *total_out = prefix + tot;
        }
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        u32* out = arr + (size_t)ch * stride;
        u32 off = s_prefix[ch] + excl[ch];
        if (fast) {
#pragma unroll
            for (int q = 0; q < SCAN3_I / 4; ++q) {
                uint4 x;
                x.x = off; off += v[ch][4 * q];
                x.y = off; off += v[ch][4 * q + 1];
                x.z = off; off += v[ch][4 * q + 2];
                x.w = off; off += v[ch][4 * q + 3];
                *reinterpret_cast<uint4*>(out + base + 4 * q) = x;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SCAN3_I; ++j) { if (base + j < n) out[base + j] = off; off += v[ch][j]; }
        }
    }
}

__global__ void scan_zero_total_k(u32* total_out) { *total_out = 0; }

// descriptor array + tile counter of the calling lane (scans on the main and on the side stream may run concurrently):
// zeroed when (re)allocated, then only ever advanced (epoch / counter base)
static int scan_state(sccg_ctx* c, size_t ndesc, u64** desc, u32** counter, int* lane_out) {
    const int ln = lane_of_stream(c);
    const int slot_desc = ln ? B_SCAN2 : B_SCAN0, slot_cnt = ln ? B_SCAN3 : B_SCAN1;
    const size_t before = c->bufs[slot_desc].cap;
    SCCG_TRY(buf(c, slot_desc, ndesc + 2, desc));
    if (c->bufs[slot_desc].cap != before) {
        SCCG_CK(cudaMemsetAsync(*desc, 0, c->bufs[slot_desc].cap, c->stream));
        c->scan_epoch[ln] = 0;
    }
    SCCG_TRY(buf(c, slot_cnt, 64, counter));
    if (!c->scan_counter_ready[ln]) {
        SCCG_CK(cudaMemsetAsync(*counter, 0, 256, c->stream));
        c->scan_counter_ready[ln] = 1; c->scan_counter_base[ln] = 0;
    }
    if (++c->scan_epoch[ln] >= 0x7fffffffu) {                    // epoch space exhausted: start over with clean descriptors
        SCCG_CK(cudaMemsetAsync(*desc, 0, c->bufs[slot_desc].cap, c->stream));
        c->scan_epoch[ln] = 1;
    }
    *lane_out = ln;
    return SCCG_OK;
}

// in may alias out.  d_total: optional device scalar receiving the grand total.
static int scan_exclusive_u32(sccg_ctx* c, const u32* in, u32* out, i64 n, u32* d_total) {
    if (n <= 0) {
        if (d_total) LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, d_total);
        return SCCG_OK;
    }
    const unsigned ntiles = div_up(n, SCAN_TILE);
    u64* desc = nullptr; u32* counter = nullptr; int ln = 0;
    SCCG_TRY(scan_state(c, (size_t)ntiles, &desc, &counter, &ln));
    LAUNCH(c, scan_onepass_k, dim3(ntiles), dim3(SCAN_T), 0, in, out, n, desc, counter, c->scan_counter_base[ln], c->scan_epoch[ln], d_total);
    c->scan_counter_base[ln] += ntiles;                          // modulo 2^32, like the device counter
    return SCCG_OK;
}

// arr, arr + stride, arr + 2 * stride: three arrays of n elements, scanned in place by one launch; totals[ch]: device scalars
static int scan_exclusive_u32x3(sccg_ctx* c, u32* arr, size_t stride, i64 n, u32* total0, u32* total1, u32* total2) {
    if (n <= 0) {
        LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, total0);
        LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, total1);
        LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, total2);
        return SCCG_OK;
    }
    const unsigned ntiles = div_up(n, SCAN3_TILE);
    u64* desc = nullptr; u32* counter = nullptr; int ln = 0;
    SCCG_TRY(scan_state(c, (size_t)ntiles * 3, &desc, &counter, &ln));
    LAUNCH(c, scan_onepass3_k, dim3(ntiles), dim3(SCAN_T), 0, arr, stride, n, desc, counter, c->scan_counter_base[ln], c->scan_epoch[ln], total0, total1, total2);
    c->scan_counter_base[ln] += ntiles;
    return SCCG_OK;
}

}  // namespace sccg
