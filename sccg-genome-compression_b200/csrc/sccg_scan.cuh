// Device-wide exclusive prefix sum over u32 (reduce-then-scan, three launches per level).
// Used for every "count -> offset" step of the record pipeline; HBM-bound, 12 B per element.
#pragma once
#include "sccg_ctx.cuh"

namespace sccg {

static const int SCAN_T = 256;
static const int SCAN_I = 8;
static const int SCAN_TILE = SCAN_T * SCAN_I;

__global__ void __launch_bounds__(SCAN_T) scan_reduce_k(const u32* __restrict__ in, u32* __restrict__ sums, i64 n) {
    __shared__ u32 sm[40];
    i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_I;
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) if (base + j < n) s += in[base + j];
    u32 tot;
    block_scan_excl(s, sm, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// out[i] = tile_off[tile] + exclusive prefix inside the tile; total_out (optional) = sum of everything
__global__ void __launch_bounds__(SCAN_T) scan_apply_k(const u32* in, u32* out, const u32* __restrict__ tile_off, i64 n, u32* total_out) {
    __shared__ u32 sm[40];
    i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_I;
    u32 v[SCAN_I];
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) { v[j] = (base + j < n) ? in[base + j] : 0u; s += v[j]; }
    u32 tot;
    u32 excl = block_scan_excl(s, sm, &tot);
    u32 off = (tile_off ? tile_off[blockIdx.x] : 0u) + excl;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) { if (base + j < n) out[base + j] = off; off += v[j]; }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = (tile_off ? tile_off[blockIdx.x] : 0u) + tot;
}

__global__ void scan_zero_total_k(u32* total_out) { *total_out = 0; }

// in may alias out.  d_total: optional device scalar receiving the grand total.
static int scan_exclusive_u32(sccg_ctx* c, const u32* in, u32* out, i64 n, u32* d_total, int depth = 0) {
    if (n <= 0) {
        if (d_total) LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, d_total);
        return SCCG_OK;
    }
    unsigned ntiles = div_up(n, SCAN_TILE);
    if (ntiles == 1) {
        LAUNCH(c, scan_apply_k, dim3(1), dim3(SCAN_T), 0, in, out, (const u32*)nullptr, n, d_total);
        return SCCG_OK;
    }
    if (depth > 2) return set_error(SCCG_E_ARG, "scan: input too large");
    u32* sums = nullptr;
    SCCG_TRY(buf(c, B_SCAN0 + depth, (size_t)ntiles, &sums));
    LAUNCH(c, scan_reduce_k, dim3(ntiles), dim3(SCAN_T), 0, in, sums, n);
    SCCG_TRY(scan_exclusive_u32(c, sums, sums, (i64)ntiles, nullptr, depth + 1));
    LAUNCH(c, scan_apply_k, dim3(ntiles), dim3(SCAN_T), 0, in, out, (const u32*)sums, n, d_total);
    return SCCG_OK;
}

}  // namespace sccg
