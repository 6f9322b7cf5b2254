// Lowercase-run and N-run extraction by parallel RLE, and the run-list text writer.
//   reference: compression.cpp:341-367 (lowercase, raw target), :527-554 (N, upper-cased target).
// HBM-bound: the target is streamed once per pass as 16-byte vectors (1 B per base per pass).
#pragma once
#include "sccg_scan.cuh"

namespace sccg {

static const int RLE_T = 256;
static const int RLE_PER_THREAD = 64;                  // bytes per thread: one block scan per 16 KB tile
static const int RLE_TILE = RLE_T * RLE_PER_THREAD;

// MODE 0: islower(raw byte)            (compression.cpp:345)
// MODE 1: toupper(raw byte) == 'N'     (compression.cpp:523, :531)
template <int MODE> __device__ __forceinline__ int rle_pred1(u8 c) {
    return MODE == 0 ? (c >= 'a' && c <= 'z') : (c == 'N' || c == 'n');
}
template <int MODE> __device__ __forceinline__ u32 rle_pred8(u64 w) {
    if (MODE == 0) return movemask8(lower_flags8(w));
    return movemask8(eq_flags8(w, 'N') | eq_flags8(w, 'n'));
}

// start / end masks of the 64 positions [i, i+64) owned by this thread (bit b <=> position i + b)
// *paren (optional): non-zero iff one of the owned positions MAY hold '(' (a literal '(' changes what the reference's
// text-level delta_encode does, compression.cpp:262-292; the compressor then takes its text-level delta pass, which is exact
// whether or not a '(' is really there).  Cheap test: '(' = 0x28 has bit 6 clear, every letter has it set, so two ANDs per
// 16 bytes rule out the whole chunk; only chunks with a non-letter byte (digits, IUPAC is fine, '>' headers) are looked at.
template <int MODE> __device__ __forceinline__ void rle_masks(const u8* __restrict__ src, i64 n, i64 i, u64* starts, u64* ends, u64* pred, u64* paren = nullptr) {
    *starts = 0; *ends = 0; *pred = 0;
    if (paren) *paren = 0;
    if (i >= n) return;
    u64 m = 0, pm = 0, all6 = ~0ull;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        if (i + 16 * v < n) {
            ulonglong2 w = *reinterpret_cast<const ulonglong2*>(src + i + 16 * v);   // buffers carry >= 64 B of slack
            m |= (u64)(rle_pred8<MODE>(w.x) | (rle_pred8<MODE>(w.y) << 8)) << (16 * v);
            if (paren) all6 &= w.x & w.y;
        }
    }
    i64 left = n - i;
    if (left < 64) m &= (1ull << (int)left) - 1ull;
    if (paren && (all6 & 0x4040404040404040ULL) != 0x4040404040404040ULL) {
        for (int b = 0; b < 64 && b < left; ++b) pm |= (u64)(src[i + b] == '(') << b;       // rare: exact look at this chunk
    }
    if (paren) *paren = pm;
    *pred = m;
    u64 prev = (i > 0) ? (u64)rle_pred1<MODE>(src[i - 1]) : 0ull;
    u64 next = (i + 64 < n) ? (u64)rle_pred1<MODE>(src[i + 64]) : 0ull;
    *starts = m & ~((m << 1) | prev);
    *ends = m & ~((m >> 1) | (next << 63));
}

template <int MODE>
__global__ void __launch_bounds__(RLE_T) rle_count_k(const u8* __restrict__ src, i64 n, u32* __restrict__ cnt_s, u32* __restrict__ cnt_e, u64* __restrict__ pred_mask,
                                                     u32* paren_flag) {
    __shared__ u32 sm[40];
    i64 i = (i64)blockIdx.x * RLE_TILE + (i64)threadIdx.x * RLE_PER_THREAD;
    u64 s, e, m;
    if (MODE == 0 && paren_flag) {
        u64 pm;
        rle_masks<MODE>(src, n, i, &s, &e, &m, &pm);
        if (pm) atomicOr(paren_flag, 1u);
    } else {
        rle_masks<MODE>(src, n, i, &s, &e, &m);
    }
    if (i < n) pred_mask[i >> 6] = m;                               // 1 bit per symbol: the write pass never reads the symbols again
    u32 packed = (u32)__popcll(s) | ((u32)__popcll(e) << 16);         // <= 8192 each per tile: no overflow
    u32 tot;
    block_scan_excl(packed, sm, &tot);
    if (threadIdx.x == 0) { cnt_s[blockIdx.x] = tot & 0xffffu; cnt_e[blockIdx.x] = tot >> 16; }
}

// run k: [run_start[k], run_end[k]) ; the k-th start pairs with the k-th end.  Works on the predicate bit masks left by
// rle_count_k (n / 8 bytes instead of n).
__global__ void __launch_bounds__(RLE_T) rle_write_k(const u64* __restrict__ pred_mask, i64 n, const u32* __restrict__ off_s, const u32* __restrict__ off_e,
                                                     int* __restrict__ run_start, int* __restrict__ run_end) {
    __shared__ u32 sm[40];
    const i64 t = (i64)blockIdx.x * RLE_T + threadIdx.x;
    const i64 i = t * RLE_PER_THREAD;
    u64 s = 0, e = 0;
    if (i < n) {
        const u64 m = pred_mask[t];
        const u64 prev = t > 0 ? pred_mask[t - 1] >> 63 : 0ull;
        const u64 next = i + 64 < n ? pred_mask[t + 1] & 1ull : 0ull;
        s = m & ~((m << 1) | prev);
        e = m & ~((m >> 1) | (next << 63));
    }
    u32 packed = (u32)__popcll(s) | ((u32)__popcll(e) << 16);
    u32 tot;
    u32 excl = block_scan_excl(packed, sm, &tot);
    u32 ks = off_s[blockIdx.x] + (excl & 0xffffu);
    u32 ke = off_e[blockIdx.x] + (excl >> 16);
    while (s) { int b = __ffsll((long long)s) - 1; s &= s - 1; run_start[ks++] = (int)(i + b); }
    while (e) { int b = __ffsll((long long)e) - 1; e &= e - 1; run_end[ke++] = (int)(i + b + 1); }
}

// text length of run-list item k  (compression.cpp:351-366)
__device__ __forceinline__ int run_item_bytes(int delta, int len, bool last_at_end) {
    if (len == 1) return dec_len_i32(delta) + (last_at_end ? 0 : 1);        // "d,"  or  "d" at the very end
    return 3 + dec_len_i32(delta) + dec_len_i32(len);                       // "(d,len)"
}

// A run list that is produced in pieces (one target slice per GPU, sccg_shard_*): positions are shifted by pos_off, the
// first `skip_first` runs of the slice continue a run of the previous slice and are not emitted, the last run is
// lengthened by what the following slices add to it, the first emitted delta refers to prev_start.  The unsharded
// paths pass the neutral element.
struct RunCarry { i64 pos_off; i64 prev_start; i64 extra_last; int skip_first; int reaches_end; };   // reaches_end: -1 = "run ends at n"
static RunCarry run_carry_none() { RunCarry rc; rc.pos_off = 0; rc.prev_start = 0; rc.extra_last = 0; rc.skip_first = 0; rc.reaches_end = -1; return rc; }

__device__ __forceinline__ bool run_item(const int* __restrict__ run_start, const int* __restrict__ run_end, u32 K, u32 k, i64 n, const RunCarry& rc,
                                         int* delta, int* len, bool* last_at_end) {
    if (k < (u32)rc.skip_first) return false;
    const i64 st = (i64)run_start[k] + rc.pos_off;
    i64 ln = (i64)run_end[k] - (i64)run_start[k];
    if (k == K - 1) ln += rc.extra_last;
    const i64 prev = k > (u32)rc.skip_first ? (i64)run_start[k - 1] + rc.pos_off : rc.prev_start;
    *delta = (int)(st - prev);
    *len = (int)ln;
    *last_at_end = k == K - 1 && (rc.reaches_end < 0 ? (i64)run_end[k] == n : rc.reaches_end != 0);
    return true;
}

__global__ void runs_bytes_k(const int* __restrict__ run_start, const int* __restrict__ run_end, const u32* __restrict__ d_count, i64 n, u32* __restrict__ bytes,
                             RunCarry rc) {
    u32 K = *d_count;
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int delta, len; bool at_end;
    bytes[k] = run_item(run_start, run_end, K, k, n, rc, &delta, &len, &at_end) ? (u32)run_item_bytes(delta, len, at_end) : 0u;
}

__global__ void runs_write_k(const int* __restrict__ run_start, const int* __restrict__ run_end, const u32* __restrict__ d_count, i64 n,
                             const u32* __restrict__ offs, u8* __restrict__ dst, RunCarry rc) {
    u32 K = *d_count;
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int delta, len; bool at_end;
    if (!run_item(run_start, run_end, K, k, n, rc, &delta, &len, &at_end)) return;
    u8* o = dst + offs[k];
    if (len == 1) {
        int w = write_dec_i32(o, delta);
        if (!at_end) o[w] = ',';
    } else {
        int w = 0;
        o[w++] = '(';
        w += write_dec_i32(o + w, delta);
        o[w++] = ',';
        w += write_dec_i32(o + w, len);
        o[w] = ')';
    }
}

// Phase 1: count runs.  Leaves the per-tile exclusive offsets in cnt_s / cnt_e and the run count in d_count.
template <int MODE>
static int rle_count(sccg_ctx* c, const u8* d_src, i64 n, int slot_cnt, int slot_mask, u32** cnt_s, u32** cnt_e, u64** pred_mask, u32* d_count, u32* d_count_e,
                     u32* d_paren_flag = nullptr) {
    unsigned ntiles = div_up(n > 0 ? n : 1, RLE_TILE);
    u32* cnt = nullptr;
    SCCG_TRY(buf(c, slot_cnt, (size_t)ntiles * 2, &cnt));
    SCCG_TRY(buf(c, slot_mask, (size_t)ntiles * RLE_T + 2, pred_mask));
    *cnt_s = cnt; *cnt_e = cnt + ntiles;
    LAUNCH(c, rle_count_k<MODE>, dim3(ntiles), dim3(RLE_T), 0, d_src, n, *cnt_s, *cnt_e, *pred_mask, d_paren_flag);
    SCCG_TRY(scan_exclusive_u32(c, *cnt_s, *cnt_s, (i64)ntiles, d_count));
    SCCG_TRY(scan_exclusive_u32(c, *cnt_e, *cnt_e, (i64)ntiles, d_count_e));
    return SCCG_OK;
}

// Phase 2 (K known on the host): materialise the runs and their text.  d_text_len receives the text length;
// the text is written at dst (capacity >= 24 * K).
template <int MODE>
static int rle_emit(sccg_ctx* c, const u64* pred_mask, i64 n, u32 K, const u32* cnt_s, const u32* cnt_e, const u32* d_count,
                    int slot_start, int slot_end, int slot_bytes, int** run_start, int** run_end, u8* dst, u32* d_text_len,
                    bool materialise_runs = true, RunCarry rc = run_carry_none()) {
    unsigned ntiles = div_up(n > 0 ? n : 1, RLE_TILE);
    SCCG_TRY(buf(c, slot_start, (size_t)K + 1, run_start));
    SCCG_TRY(buf(c, slot_end, (size_t)K + 1, run_end));
    u32* bytes = nullptr;
    SCCG_TRY(buf(c, slot_bytes, (size_t)K + 1, &bytes));
    if (K == 0) { LAUNCH(c, scan_zero_total_k, dim3(1), dim3(1), 0, d_text_len); return SCCG_OK; }
    if (materialise_runs) LAUNCH(c, rle_write_k, dim3(ntiles), dim3(RLE_T), 0, pred_mask, n, cnt_s, cnt_e, *run_start, *run_end);
    unsigned g = div_up(K, 256);
    LAUNCH(c, runs_bytes_k, dim3(g), dim3(256), 0, (const int*)*run_start, (const int*)*run_end, d_count, n, bytes, rc);
    SCCG_TRY(scan_exclusive_u32(c, bytes, bytes, (i64)K, d_text_len));
    if (dst) LAUNCH(c, runs_write_k, dim3(g), dim3(256), 0, (const int*)*run_start, (const int*)*run_end, d_count, n, (const u32*)bytes, dst, rc);
    return SCCG_OK;
}

}  // namespace sccg
