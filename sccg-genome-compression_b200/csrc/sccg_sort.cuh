// Stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass: the reference k-mer index
// build of global mode (compression.cpp:41-47 over the whole N-stripped reference).
//   key   = 32-bit hash of the k-mer at position p,  value = p
// Stable + ascending input order  =>  inside every key the positions stay ascending, which is the
// reference's per-bucket order (vector<int>::push_back in ascending i).
// Per pass: histogram (read 4 B/elem), scan of 256 x nblocks counters, scatter (read 8 B, write 8 B).
#pragma once
#include "sccg_scan.cuh"

namespace sccg {

static const int RS_WARPS = 8;
static const int RS_T = RS_WARPS * 32;
static const int RS_CHUNKS = 16;                    // 32-element chunks per warp
static const int RS_TILE = RS_T * RS_CHUNKS;        // 4096 elements per block

__global__ void __launch_bounds__(RS_T) rs_hist_k(const u32* __restrict__ keys, i64 n, int shift, u32* __restrict__ hist, unsigned nblocks) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    i64 base = (i64)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_CHUNKS; ++r) {
        i64 i = base + (i64)r * RS_T + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];     // digit-major for the scan
}

// Stable scatter of one 4096-element tile.  The tile is first ordered by digit in shared memory (per-warp ranks from
// __match_any_sync, warps and digits combined by small scans), then written out: consecutive shared-memory entries of
// one digit go to consecutive global addresses, so the stores are coalesced instead of 32 scattered sectors per warp.
__global__ void __launch_bounds__(RS_T) rs_scatter_k(const u32* __restrict__ keys_in, const u32* __restrict__ vals_in, i64 n, int shift,
                                                    const u32* __restrict__ offs, unsigned nblocks, u32* __restrict__ keys_out, u32* __restrict__ vals_out) {
    __shared__ u32 cnt[RS_WARPS][256];              // per-warp digit counts, then running tile-local positions
    __shared__ u32 lstart[256];                     // tile-local start of every digit
    __shared__ u32 gbase[256];                      // global start of (digit, this tile)
    __shared__ u32 sk[RS_TILE], sv[RS_TILE];
    __shared__ u32 wsum[RS_WARPS];
    const int lane = lane_of(), w = (int)(threadIdx.x >> 5);
    for (int x = (int)threadIdx.x; x < RS_WARPS * 256; x += RS_T) (&cnt[0][0])[x] = 0u;
    __syncthreads();
    // each warp owns a contiguous sub-tile and walks it in order: chunk c = elements [c*32, c*32+32)
    const i64 tbase = (i64)blockIdx.x * RS_TILE;
    const i64 wbase = tbase + (i64)w * (RS_CHUNKS * 32);
    u32 k[RS_CHUNKS], v[RS_CHUNKS];
#pragma unroll
    for (int c = 0; c < RS_CHUNKS; ++c) {
        i64 i = wbase + c * 32 + lane;
        bool valid = i < n;
        k[c] = valid ? keys_in[i] : 0xffffffffu;
        v[c] = valid ? vals_in[i] : 0u;
        u32 d = (k[c] >> shift) & 255u;
        u32 peers = __match_any_sync(SCCG_FULL_MASK, valid ? d : 256u);
        if (valid && lane == __ffs((int)peers) - 1) cnt[w][d] += (u32)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: tile total, exclusive prefix over the warps, exclusive scan over the digits
        const u32 d = threadIdx.x;
        u32 tot = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { u32 t = cnt[ww][d]; cnt[ww][d] = tot; tot += t; }
        u32 incl = warp_scan_incl(tot);
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        u32 before = 0;
        for (int ww = 0; ww < w; ++ww) before += wsum[ww];
        u32 ls = before + incl - tot;
        lstart[d] = ls;
        gbase[d] = offs[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) cnt[ww][d] += ls;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < RS_CHUNKS; ++c) {
        i64 i = wbase + c * 32 + lane;
        bool valid = i < n;
        u32 d = (k[c] >> shift) & 255u;
        u32 peers = __match_any_sync(SCCG_FULL_MASK, valid ? d : 256u);
        u32 below = peers & ((1u << lane) - 1u);
        if (valid) {
            u32 pos = cnt[w][d] + (u32)__popc(below);
            sk[pos] = k[c];
            sv[pos] = v[c];
        }
        __syncwarp();
        if (valid && lane == __ffs((int)peers) - 1) cnt[w][d] += (u32)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    const int tile_n = (int)((n - tbase) < RS_TILE ? (n - tbase) : RS_TILE);
    for (int i = (int)threadIdx.x; i < tile_n; i += RS_T) {
        u32 key = sk[i];
        u32 d = (key >> shift) & 255u;
        u32 pos = gbase[d] + ((u32)i - lstart[d]);
        keys_out[pos] = key;
        vals_out[pos] = sv[i];
    }
}

// sorts (keys, vals) of length n by the low 8 * passes bits of the key; (keys2, vals2) is scratch of the same size.
// *out_keys / *out_vals: where the result is (the input arrays for an even number of passes, the scratch arrays otherwise)
static int radix_sort_pairs(sccg_ctx* c, u32* keys, u32* vals, u32* keys2, u32* vals2, i64 n, int slot_hist, int passes, u32** out_keys, u32** out_vals) {
    *out_keys = keys; *out_vals = vals;
    if (n <= 1) return SCCG_OK;
    unsigned nblocks = div_up(n, RS_TILE);
    u32* hist = nullptr;
    SCCG_TRY(buf(c, slot_hist, (size_t)nblocks * 256 + 1, &hist));
    u32 *ki = keys, *vi = vals, *ko = keys2, *vo = vals2;
    for (int pass = 0; pass < passes; ++pass) {
        int shift = pass * 8;
        LAUNCH(c, rs_hist_k, dim3(nblocks), dim3(RS_T), 0, (const u32*)ki, n, shift, hist, nblocks);
        SCCG_TRY(scan_exclusive_u32(c, hist, hist, (i64)nblocks * 256, nullptr));
        LAUNCH(c, rs_scatter_k, dim3(nblocks), dim3(RS_T), 0, (const u32*)ki, (const u32*)vi, n, shift, (const u32*)hist, nblocks, ko, vo);
        u32* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    *out_keys = ki; *out_vals = vi;
    return SCCG_OK;
}

}  // namespace sccg
