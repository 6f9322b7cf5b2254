// Stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass: the reference k-mer index
// build of global mode (compression.cpp:41-47 over the whole N-stripped reference).
//   key   = hash of the k-mer at position p (24 significant bits -> 3 passes),  value = p
// Stable + ascending input order  =>  inside every key the positions stay ascending, which is the
// reference's per-bucket order (vector<int>::push_back in ascending i).
// Per pass: histogram (read 4 B/elem), scan of 256 x nblocks counters, scatter (read 8 B, write 8 B).  The scatter ranks the
// keys of a 4096-element tile by digit with ballots, orders the tile in shared memory and stores runs of equal digits.
// The first pass takes its pairs from a SOURCE functor (key(i), val(i)): the k-mer index computes hash and position on
// the fly from the sequence instead of materialising 8 B per k-mer first.
#pragma once
#include "sccg_scan.cuh"

namespace sccg {

#ifndef SCCG_RS_MINB
#define SCCG_RS_MINB 4
#endif
#ifndef SCCG_RS_WARPS
#define SCCG_RS_WARPS 8
#endif
static const int RS_WARPS = SCCG_RS_WARPS;          // >= 8: the first 256 threads also own one digit each
static const int RS_T = RS_WARPS * 32;
static const int RS_CHUNKS = 128 / RS_WARPS;        // 32-element chunks per warp
static const int RS_TILE = RS_T * RS_CHUNKS;        // 4096 elements per block

struct RsPairSource {                               // pairs that already sit in memory
    const u32* keys; const u32* vals;
    static const bool kRun16 = false;               // true: the source has keys16(i0, out) for 16 consecutive indices
    __device__ __forceinline__ void keys16(i64, u32*) const {}
    __device__ __forceinline__ u32 key(i64 i) const { return keys[i]; }
    __device__ __forceinline__ u32 val(i64 i) const { return vals[i]; }
};

template <class Src>
__global__ void __launch_bounds__(RS_T) rs_hist_k(const Src src, i64 n, int shift, u32* __restrict__ hist, unsigned nblocks) {
    __shared__ u32 h[256];
    if (threadIdx.x < 256u) h[threadIdx.x] = 0u;
    __syncthreads();
    const i64 base = (i64)blockIdx.x * RS_TILE;
    u32 kk[RS_CHUNKS];                                  // all loads first: the kernel was latency-bound with load-use pairs
    if (Src::kRun16 && RS_CHUNKS == 16) {               // the order inside a tile does not matter here: 16 consecutive elements per thread
        const i64 i0 = base + (i64)threadIdx.x * 16;
        if (i0 < n) src.keys16(i0, kk);
#pragma unroll
        for (int r = 0; r < RS_CHUNKS; ++r)
            if (i0 + r < n) atomicAdd(&h[(kk[r] >> shift) & 255u], 1u);
    } else {
#pragma unroll
        for (int r = 0; r < RS_CHUNKS; ++r) {
            const i64 i = base + (i64)r * RS_T + threadIdx.x;
            kk[r] = i < n ? src.key(i) : 0u;
        }
#pragma unroll
        for (int r = 0; r < RS_CHUNKS; ++r) {
            const i64 i = base + (i64)r * RS_T + threadIdx.x;
            if (i < n) atomicAdd(&h[(kk[r] >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 256u) hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];     // digit-major for the scan
}

// lanes of the warp whose 8-bit digit equals this lane's: eight ballots, one per digit bit.  (match.any is a single
// instruction but its latency is in the hundreds of cycles; ncu showed rs_scatter_k 75 % stalled on it.)
__device__ __forceinline__ u32 digit_peers(u32 d) {
    u32 peers = SCCG_FULL_MASK;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const u32 bal = __ballot_sync(SCCG_FULL_MASK, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// Stable scatter of one 4096-element tile.  The tile is first ordered by digit in shared memory (per-warp ranks from
// ballots, warps and digits combined by small scans), then written out: consecutive shared-memory entries of one digit
// go to consecutive global addresses, so the stores are coalesced instead of 32 scattered sectors per warp.
// Elements past n take the key 0xffffffff: digit 255 in every pass and last in tile order, so they rank after every
// real element of the tile and are simply not written.
template <class Src>
__global__ void __launch_bounds__(RS_T, SCCG_RS_MINB) rs_scatter_k(const Src src, i64 n, int shift, const u32* __restrict__ offs, unsigned nblocks,
                                                                 u32* __restrict__ keys_out, u32* __restrict__ vals_out) {
    __shared__ u16 cnt[RS_WARPS][256];              // per-warp digit counts, then the tile-local start of (warp, digit); all <= RS_TILE = 4096
    __shared__ u32 gdelta[256];                     // global start of (digit, this tile) minus its tile-local start
    __shared__ u32 sk[RS_TILE], sv[RS_TILE];
    __shared__ u32 wsum[8];
    const int lane = lane_of(), w = (int)(threadIdx.x >> 5);
    for (int x = (int)threadIdx.x; x < RS_WARPS * 128; x += RS_T) reinterpret_cast<u32*>(&cnt[0][0])[x] = 0u;
    // each warp owns a contiguous sub-tile and walks it in order: chunk c = elements [c*32, c*32+32)
    const i64 tbase = (i64)blockIdx.x * RS_TILE;
    const i64 wbase = tbase + (i64)w * (RS_CHUNKS * 32);
    u32 k[RS_CHUNKS], rk[RS_CHUNKS];
    if (Src::kRun16 && RS_CHUNKS == 16) {
        // every lane computes the keys of 16 consecutive elements of the warp's sub-tile; the part of sk that the warp will
        // later fill transposes them into the chunk layout (chunk c, lane l = element c * 32 + l)
        u32* my = sk + w * (RS_CHUNKS * 32);
        const i64 i0 = wbase + (i64)lane * 16;
        if (i0 < n) src.keys16(i0, rk);
#pragma unroll
        for (int r = 0; r < 16; ++r) my[lane * 16 + ((r + lane) & 15)] = rk[r];     // rotated: conflict-free columns
        __syncwarp();
#pragma unroll
        for (int c = 0; c < RS_CHUNKS; ++c) {
            const int e = c * 32 + lane;                                            // element index inside the sub-tile
            const u32 kv = my[(e & ~15) + (((e & 15) + (e >> 4)) & 15)];
            k[c] = wbase + e < n ? kv : 0xffffffffu;
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int c = 0; c < RS_CHUNKS; ++c) {
            const i64 i = wbase + c * 32 + lane;
            k[c] = i < n ? src.key(i) : 0xffffffffu;
        }
    }
#pragma unroll
    for (int c = 0; c < RS_CHUNKS; ++c) rk[c] = digit_peers((k[c] >> shift) & 255u);
    __syncthreads();
    // rank inside the warp's sub-tile: elements of the same digit in earlier chunks + same-digit lanes below
#pragma unroll
    for (int c = 0; c < RS_CHUNKS; ++c) {
        const u32 d = (k[c] >> shift) & 255u;
        const u32 peers = rk[c];
        const int leader = __ffs((int)peers) - 1;
        u32 old = 0;
        if (lane == leader) { old = cnt[w][d]; cnt[w][d] = (u16)(old + (u32)__popc(peers)); }
        old = __shfl_sync(SCCG_FULL_MASK, old, leader);
        rk[c] = old + (u32)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: tile total, exclusive prefix over the warps, exclusive scan over the digits
        const u32 d = threadIdx.x;
        u32 tot = 0, incl = 0;
        if (d < 256u) {
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) { u32 t = cnt[ww][d]; cnt[ww][d] = (u16)tot; tot += t; }
            incl = warp_scan_incl(tot);
            if (lane == 31) wsum[w] = incl;
        }
        __syncthreads();
        if (d < 256u) {
            u32 before = 0;
            for (int ww = 0; ww < w; ++ww) before += wsum[ww];
            const u32 ls = before + incl - tot;
            gdelta[d] = offs[(size_t)d * nblocks + blockIdx.x] - ls;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) cnt[ww][d] = (u16)(cnt[ww][d] + ls);
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < RS_CHUNKS; ++c) {
        const i64 i = wbase + c * 32 + lane;
        const u32 pos = cnt[w][(k[c] >> shift) & 255u] + rk[c];
        sk[pos] = k[c];
        sv[pos] = i < n ? src.val(i) : 0u;
    }
    __syncthreads();
    const int tile_n = (int)((n - tbase) < RS_TILE ? (n - tbase) : RS_TILE);
#pragma unroll 4
    for (int i = (int)threadIdx.x; i < tile_n; i += RS_T) {
        const u32 key = sk[i];
        const u32 pos = gdelta[(key >> shift) & 255u] + (u32)i;
        keys_out[pos] = key;
        vals_out[pos] = sv[i];
    }
}

// sorts the n pairs of `first` by the low 8 * passes bits of the key.  (keys, vals) and (keys2, vals2) are buffers of n
// entries each; *out_keys / *out_vals: the pair that holds the result (the first pass writes into (keys, vals)).
template <class Src>
static int radix_sort_pairs(sccg_ctx* c, const Src& first, u32* keys, u32* vals, u32* keys2, u32* vals2, i64 n, int slot_hist, int passes, u32** out_keys, u32** out_vals) {
    *out_keys = keys; *out_vals = vals;
    if (n <= 0 || passes <= 0) return SCCG_OK;
    unsigned nblocks = div_up(n, RS_TILE);
    u32* hist = nullptr;
    SCCG_TRY(buf(c, slot_hist, (size_t)nblocks * 256 + 1, &hist));
    u32 *ki = keys2, *vi = vals2, *ko = keys, *vo = vals;
    for (int pass = 0; pass < passes; ++pass) {
        int shift = pass * 8;
        if (pass == 0) LAUNCH(c, rs_hist_k<Src>, dim3(nblocks), dim3(RS_T), 0, first, n, shift, hist, nblocks);
        else { RsPairSource ps{ki, vi}; LAUNCH(c, rs_hist_k<RsPairSource>, dim3(nblocks), dim3(RS_T), 0, ps, n, shift, hist, nblocks); }
        SCCG_TRY(scan_exclusive_u32(c, hist, hist, (i64)nblocks * 256, nullptr));
        if (pass == 0) LAUNCH(c, rs_scatter_k<Src>, dim3(nblocks), dim3(RS_T), 0, first, n, shift, (const u32*)hist, nblocks, ko, vo);
        else { RsPairSource ps{ki, vi}; LAUNCH(c, rs_scatter_k<RsPairSource>, dim3(nblocks), dim3(RS_T), 0, ps, n, shift, (const u32*)hist, nblocks, ko, vo); }
        u32* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    *out_keys = ki; *out_vals = vi;
    return SCCG_OK;
}

}  // namespace sccg
