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

__global__ void scan_zero_total_k(u32* total_out) { *total_out = 0; }

// in may alias out.  d_total: optional device scalar receiving the grand total.
static int scan_exclusive_u32(sccg_ctx* c, const u32* in, u32* out, i64 n, u32* d_total) {
    if (n <= 0) {
        if (d_total) LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, d_total);
        return SCCG_OK;
    }
    const unsigned ntiles = div_up(n, SCAN_TILE);
    // descriptor array + counter (one set per lane: scans on the main and on the side stream may run concurrently):
    // zeroed when (re)allocated, then only ever advanced (epoch / counter base)
    const int ln = lane_of_stream(c);
    const int slot_desc = ln ? B_SCAN2 : B_SCAN0, slot_cnt = ln ? B_SCAN3 : B_SCAN1;
    const size_t before = c->bufs[slot_desc].cap;
    u64* desc = nullptr;
    SCCG_TRY(buf(c, slot_desc, (size_t)ntiles + 2, &desc));
    if (c->bufs[slot_desc].cap != before) {
        SCCG_CK(cudaMemsetAsync(desc, 0, c->bufs[slot_desc].cap, c->stream));
        c->scan_epoch[ln] = 0;
    }
    u32* counter = nullptr;
    SCCG_TRY(buf(c, slot_cnt, 64, &counter));
    if (!c->scan_counter_ready[ln]) {
        SCCG_CK(cudaMemsetAsync(counter, 0, 256, c->stream));
        c->scan_counter_ready[ln] = 1; c->scan_counter_base[ln] = 0;
    }
    if (++c->scan_epoch[ln] >= 0x7fffffffu) {                    // epoch space exhausted: start over with clean descriptors
        SCCG_CK(cudaMemsetAsync(desc, 0, c->bufs[slot_desc].cap, c->stream));
        c->scan_epoch[ln] = 1;
    }
    LAUNCH(c, scan_onepass_k, dim3(ntiles), dim3(SCAN_T), 0, in, out, n, desc, counter, c->scan_counter_base[ln], c->scan_epoch[ln], d_total);
    c->scan_counter_base[ln] += ntiles;                          // modulo 2^32, like the device counter
    return SCCG_OK;
}

}  // namespace sccg
