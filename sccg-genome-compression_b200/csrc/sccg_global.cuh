// Global mode (compression.cpp:484-574): N-run extraction, N-stripping, reference k-mer index and
// the stateful banded greedy parse  match_sequences(ref, tgt, k = 14, m = 100, global = true)  (:561).
//
//   index   : 24-bit hash of every reference k-mer + stable radix sort (3 passes)  ->  (key, position) sorted by key,
//             positions ascending inside a key (the reference's bucket order, :41-47)
//   parse   : a CTA walks the state (index, prev_match_end) exactly like :64-161 (gp_step); the target is parsed in
//             speculative chunks by many CTAs and stitched by an exact front (gp_spec_k / gp_front_k below).
//             * until the first match (prev_match_end == -1, every candidate is "in range") and in the
//               `pn2 == 0` fall-through (:134) candidates come from the sorted index (binary search);
//             * otherwise only candidates with |p - prev_match_end| <= m can be used (:83-96, :116), i.e.
//               the k-mers of a 2m+1 window of the reference: the CTA stages that window in shared
//               memory, tries the next four positions by brute force, else hashes the window into a small
//               filter and scans the target forward 1 position per thread until a position hits the window ("no k-mer" and "no candidate in range" are
//               the same literal step, :77-81 vs :92-96).
//             * candidate selection is the order-independent form of the ascending-p fold (:114-130).
//   writer  : tokens "(dp,l)" with the delta chain of delta_encode (:258-292) + literal gaps.
#pragma once
#include "sccg_compress.cuh"
#include "sccg_strip.cuh"
#include "sccg_sort.cuh"

namespace sccg {

#ifndef SCCG_GP_T
#define SCCG_GP_T 256
#endif
static const int GP_T = SCCG_GP_T;           // threads of the parse CTA
static const int GP_MAX_M = 120;             // window = 2m+1 <= 241 positions
static_assert(GP_T >= 64 && GP_T % 32 == 0, "64 diagonal probes, one per thread");
static const int GP_WIN_BYTES = 288;
static const int GP_FILTER = 512;
static const int GP_BF = 4;                  // positions tried by brute force before the window filter is built
static const int GP_SHORT = 64;              // per-thread extension before the block-wide one takes over

// 24-bit hash of the k-mer s[0..k), 8 <= k <= 16; the words may come from global or shared memory.  24 bits = three radix
// passes for the index; k-mers that share a hash are told apart by comparing the symbols (every consumer does).
static const int GP_HASH_BITS = 24;
static const unsigned GP_FIRST_CAP = 4096;     // sampled index mode: capacity of the occurrence list of the first matching target k-mer
static const int GP_FIRST_W = 256;             //   target positions searched for it
enum { GP_FIRST_N0 = 0, GP_FIRST_J = 1, GP_FIRST_N1 = 2 };
// The offset table over the sorted keys is addressed by the top `bits` bits of the hash: 20 bits (4 MB, L2-resident, a few
// dozen keys per bucket) for chromosome-sized indexes, fewer for small ones so that filling the table stays proportional
// to the index.  (All 24 bits -- one bucket per hash value, no search through the keys -- measured no faster.)
static const int GP_BUCKET_BITS_MAX = 20;
__host__ __device__ inline int gp_bucket_bits(i64 nk) {
    int b = 12;
    while (b < GP_BUCKET_BITS_MAX && ((i64)1 << (b + 1)) <= nk) ++b;
    return b;
}
__device__ __forceinline__ u32 kmer_hash_words(u64 w0, u64 w1, int k) {
    if (k < 16) w1 &= (k == 8) ? 0ull : (~0ull >> (8 * (16 - k)));
    u64 x = (w0 * 0x9E3779B97F4A7C15ULL) ^ ((w1 + 0x632BE59BD9B4E019ULL) * 0xD6E8FEB86659FD93ULL);
    x ^= x >> 29;
    x *= 0x94D049BB133111EBULL;
    return (u32)(x >> (64 - GP_HASH_BITS));
}
__device__ __forceinline__ bool kmer_equal_words(u64 a0, u64 a1, u64 b0, u64 b1, int k) {
    u64 m1 = (k == 8) ? 0ull : (k < 16 ? (~0ull >> (8 * (16 - k))) : ~0ull);
    return a0 == b0 && ((a1 ^ b1) & m1) == 0ull;
}

// the (hash, position) pairs of the reference k-mers, computed where the sort reads them (no 8 B per k-mer staging pass)
struct KmerPairSource {
    const u8* R; int k;
    static const bool kRun16 = true;
    __device__ __forceinline__ u32 key(i64 p) const { return kmer_hash_words(ld_unaligned64(R + p), ld_unaligned64(R + p + 8), k); }
    __device__ __forceinline__ u32 val(i64 p) const { return (u32)p; }
    // keys of the 16 consecutive positions p0 .. p0 + 15: 32 bytes loaded once, every 16-byte window cut out with
    // constant shifts (key() costs 3-4 loads and two variable funnel shifts per position)
    __device__ __forceinline__ void keys16(i64 p0, u32* out) const {
        const uintptr_t a = (uintptr_t)(R + p0);
        const u64* q = reinterpret_cast<const u64*>(a & ~(uintptr_t)7);
        const u32 sh = (u32)(a & 7) * 8u;
        u64 W[4];
        if (sh == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) W[j] = q[j];
        } else {
            u64 x = q[0];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const u64 y = q[j + 1]; W[j] = (x >> sh) | (y << (64u - sh)); x = y; }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = i >> 3, b = (i & 7) * 8;
            const u64 w0 = b ? (W[j] >> b) | (W[j + 1] << (64 - b)) : W[j];
            const u64 w1 = b ? (W[j + 1] >> b) | (W[j + 2] << (64 - b)) : W[j + 1];
            out[i] = kmer_hash_words(w0, w1, k);
        }
    }
};

// SAMPLED index: only the reference positions 0, stride, 2 * stride, ... (element i = position i * stride).  Enough for the
// diagonal guesses of gp_spec_k (64 consecutive probe positions meet 64 / stride sampled positions of the true diagonal);
// lookups that must see EVERY occurrence are served otherwise (gp_first_k) or make the host build the full index.
struct KmerSampledSource {
    const u8* R; int k; int stride;
    static const bool kRun16 = false;
    __device__ __forceinline__ void keys16(i64, u32*) const {}
    __device__ __forceinline__ u32 key(i64 i) const { const u8* q = R + i * stride; return kmer_hash_words(ld_unaligned64(q), ld_unaligned64(q + 8), k); }
    __device__ __forceinline__ u32 val(i64 i) const { return (u32)(i * stride); }
};

// bucket[b] for every b in [0, 2^bits]: index i owns the buckets that begin between keys[i-1] and keys[i]
// (i = nk: the buckets past the last key).  8 indices per thread (two 16-byte loads; one index per thread was latency-bound).
static const int KB_SPAN = 8;
__global__ void __launch_bounds__(256) kmer_buckets_k(const u32* __restrict__ keys, i64 nk, u32* __restrict__ bucket, int bits) {
    const int shift = GP_HASH_BITS - bits;
    const i64 i0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * KB_SPAN;
    if (i0 > nk) return;
    u32 kk[KB_SPAN];
    if (i0 + KB_SPAN <= nk) {                                  // keys is a 16-byte aligned allocation
        const uint4 a = *reinterpret_cast<const uint4*>(keys + i0), b = *reinterpret_cast<const uint4*>(keys + i0 + 4);
        kk[0] = a.x; kk[1] = a.y; kk[2] = a.z; kk[3] = a.w; kk[4] = b.x; kk[5] = b.y; kk[6] = b.z; kk[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < KB_SPAN; ++j) kk[j] = i0 + j < nk ? keys[i0 + j] : 0u;
    }
    u32 prev = i0 == 0 ? 0u : (keys[i0 - 1] >> shift) + 1u;                // first bucket not yet assigned
#pragma unroll
    for (int j = 0; j < KB_SPAN; ++j) {
        const i64 i = i0 + j;
        if (i > nk) break;
        const u32 hi = i == nk ? (1u << bits) : (kk[j] >> shift);
        for (u32 bkt = prev; bkt <= hi; ++bkt) bucket[bkt] = (u32)i;
        prev = hi + 1u;
    }
}

// Sampled mode: every reference position whose k-mer equals the FIRST k-mer of the target, T[0..k) -- the candidates of
// the parse's first step (prev_match_end == -1: all of them are in range, :87).  One pass over the reference with all SMs,
// 16 positions per thread: the 4-byte window at every position (one funnel shift) against the first 4 symbols of the k-mer,
// the rare survivors symbol by symbol.  The list is unordered (the candidate fold does not depend on the order).
// R must be 16-byte aligned (device buffers are); 24 bytes are read per thread (buffers carry >= 64 bytes of slack).
__device__ __forceinline__ void first_windows(const u8* __restrict__ R, i64 p0, u32* x) {
    const uint4 v = *reinterpret_cast<const uint4*>(R + p0);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    const uint2 w = *reinterpret_cast<const uint2*>(R + p0 + 16);
    x[4] = w.x; x[5] = w.y;
}
__device__ __forceinline__ void first_occurrences16(const u8* __restrict__ R, i64 nk, i64 p0, const u8* __restrict__ tj, int k, u32* __restrict__ list, u32 cap, u32* __restrict__ d_n) {
    const u64 t0 = ld_unaligned64(tj), t1 = ld_unaligned64(tj + 8);
    const u32 t32 = (u32)t0;
    u32 x[6];
    first_windows(R, p0, x);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 w = (i & 3) ? __funnelshift_r(x[i >> 2], x[(i >> 2) + 1], 8 * (i & 3)) : x[i >> 2];
        if (w == t32 && p0 + i < nk) {
            const u8* q = R + p0 + i;
            if (kmer_equal_words(ld_unaligned64(q), ld_unaligned64(q + 8), t0, t1, k)) {
                const u32 at = atomicAdd(d_n, 1u);
                if (at < cap) list[at] = (u32)(p0 + i);
            }
        }
    }
}
__global__ void __launch_bounds__(256) gp_first_k(const u8* __restrict__ R, i64 nk, const u8* __restrict__ T, int k, u32* __restrict__ list, u32 cap, u32* __restrict__ first) {
    const i64 p0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 >= nk) return;
    first_occurrences16(R, nk, p0, T, k, list, cap, first + GP_FIRST_N0);
}
// ... of T[j..j+k) where j is the position gp_first_scan_k found; does nothing when T[0..k) had occurrences or the scan found none
__global__ void __launch_bounds__(256) gp_first_collect_k(const u8* __restrict__ R, i64 nk, const u8* __restrict__ T, int k, u32* __restrict__ list, u32 cap, u32* __restrict__ first) {
    if (first[GP_FIRST_N0] != 0u || first[GP_FIRST_J] == 0xffffffffu) return;        // (nobody writes these two words any more)
    const i64 p0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 >= nk) return;
    first_occurrences16(R, nk, p0, T + first[GP_FIRST_J], k, list, cap, first + GP_FIRST_N1);
}

// ------------------------------------------------------------------------------------------------
// the persistent parse CTA
// ------------------------------------------------------------------------------------------------
struct GpShared {
    u8 win[GP_WIN_BYTES];
    u32 f_hash[GP_FILTER];
    u8 f_off[GP_FILTER];
    int i_scratch[8];
    int bf_hit[4];
    unsigned long long key_scratch;
    int long_list[GP_T > 128 ? GP_T : 128];       // also the 64 diagonal votes (i64) of gp_spec_k
    int long_n;
    // accumulators of the candidate fold
    int best_l, cnt, zero_in;
    unsigned long long best_key;
};

struct GpArgs;
__device__ __forceinline__ i64 index_lower_bound(const GpArgs& a, u32 h);

struct GpArgs {
    const u8* R; i64 nr;          // N-stripped, upper-cased reference
    const u8* T; i64 nt;          // N-stripped, upper-cased target
    const u32* keys; const u32* vals; i64 nk;
    const u32* bucket;            // bucket[b] = first index whose key >> bucket_shift is >= b (2^bits + 1 entries)
    int bucket_shift;             // GP_HASH_BITS - bits
    int k, m;
    int full;                     // 1: (keys, vals) index every reference k-mer; 0: a sampled index (guesses only)
    const u32* first_list;        // sampled mode: the reference positions of the first target k-mer that occurs in the reference at all, unordered
    const u32* first;             //   first[GP_FIRST_N0]: occurrences of T[0..k) (gp_first_k); if 0: first[GP_FIRST_J] = first position < GP_FIRST_W
                                  //   whose k-mer occurs (gp_first_scan_k, ~0u = none), first[GP_FIRST_N1] = its occurrences (gp_first_collect_k)
    u32* need_full;               //   raised by a lookup that the sampled mode cannot serve: the host repeats the parse with the full index
    int* m_tpos; int* m_p; int* m_l;   // out: matches
    u32* d_count;                      // out: number of matches
};

// first index whose key is >= h (sorted index): the offset table narrows the search to one bucket (a few dozen entries)
__device__ __forceinline__ i64 index_lower_bound(const GpArgs& a, u32 h) {
    const u32 b = h >> a.bucket_shift;
    if (b >= (1u << (GP_HASH_BITS - a.bucket_shift))) return a.nk;      // h == 2^24: one past the last key
    i64 lo = a.bucket[b];
    i64 hi = a.bucket[b + 1];
    while (lo < hi) { i64 mid = (lo + hi) >> 1; if (a.keys[mid] < h) lo = mid + 1; else hi = mid; }
    return lo;
}
// vals[lo .. hi): the reference positions whose k-mer has the 24-bit hash h, ascending.  Every consumer compares the symbols.
__device__ __forceinline__ void index_range(const GpArgs& a, u32 h, i64* lo, i64* hi) {
    *lo = index_lower_bound(a, h);
    *hi = index_lower_bound(a, h + 1u);
}

// length of the common prefix of R[p..] and T[j..], capped at cap (cap <= remaining lengths)
__device__ __forceinline__ int serial_lcp(const u8* __restrict__ R, i64 p, const u8* __restrict__ T, i64 j, int cap) {
    int l = 0;
    while (l < cap) {
        u64 d = ld_unaligned64(R + p + l) ^ ld_unaligned64(T + j + l);
        if (d) { l += (__ffsll((long long)d) - 1) >> 3; break; }
        l += 8;
    }
    return l < cap ? l : cap;
}

// block-wide: min(maxl, lcp(R[p..], T[j..])) ; all threads, uniform arguments
__device__ __forceinline__ i64 block_lcp(GpShared& S, const u8* __restrict__ R, i64 p, const u8* __restrict__ T, i64 j, i64 maxl) {
    for (i64 base = 0; base < maxl; base += (i64)GP_T * 8) {
        i64 o = base + (i64)threadIdx.x * 8;
        int mis = -1;                                            // offset of the first mismatch inside my 8 bytes
        if (o >= maxl) mis = 0;
        else {
            u64 d = ld_unaligned64(R + p + o) ^ ld_unaligned64(T + j + o);
            if (d) mis = (__ffsll((long long)d) - 1) >> 3;
        }
        if (threadIdx.x == 0) S.i_scratch[0] = 0x7fffffff;
        if (__syncthreads_or(mis >= 0)) {
            if (mis >= 0) atomicMin(&S.i_scratch[0], (int)threadIdx.x * 8 + mis);
            __syncthreads();
            i64 l = base + S.i_scratch[0];
            __syncthreads();
            return l < maxl ? l : maxl;
        }
    }
    return maxl;
}

__device__ __forceinline__ void fold_reset(GpShared& S) {
    if (threadIdx.x == 0) { S.best_l = 0; S.cnt = 0; S.zero_in = 0; S.best_key = ~0ull; }
    __syncthreads();
}

// merges one chunk of candidates (one per thread, p < 0 = none) into the fold accumulators: the order-independent
// form of compression.cpp:114-130 (max length; ties: nearest to prev_match_end, then smaller p; p == 0 is "unset")
__device__ __forceinline__ void fold_chunk(GpShared& S, const GpArgs& a, i64 p, i64 j, int e) {
    int my_l = 0;
    bool is_long = false;
    if (p >= 0) {
        i64 maxl = (a.nr - p) < (a.nt - j) ? (a.nr - p) : (a.nt - j);
        int cap = maxl < GP_SHORT ? (int)maxl : GP_SHORT;
        int l = serial_lcp(a.R, p, a.T, j, cap);                 // extend_alignment :27-34 (from offset 0: verifies the k-mer too)
        if (l >= a.k) { my_l = l; is_long = (l == GP_SHORT && maxl > GP_SHORT); }
    }
    if (threadIdx.x == 0) S.long_n = 0;
    int n_long = __syncthreads_count(is_long);
    if (n_long) {
        if (is_long) S.long_list[atomicAdd(&S.long_n, 1)] = (int)threadIdx.x;
        __syncthreads();
        for (int i = 0; i < n_long; ++i) {
            int owner = S.long_list[i];
            if ((int)threadIdx.x == owner) { S.i_scratch[2] = (int)(p & 0xffffffff); S.i_scratch[3] = (int)(p >> 32); }
            __syncthreads();
            i64 pp = ((i64)S.i_scratch[3] << 32) | (u32)S.i_scratch[2];
            i64 maxl = (a.nr - pp) < (a.nt - j) ? (a.nr - pp) : (a.nt - j);
            i64 l = GP_SHORT + block_lcp(S, a.R, pp + GP_SHORT, a.T, j + GP_SHORT, maxl - GP_SHORT);
            if ((int)threadIdx.x == owner) my_l = (int)l;
        }
    }
    // chunk maximum
    if (threadIdx.x == 0) S.i_scratch[1] = 0;
    __syncthreads();
    if (my_l > 0) atomicMax(&S.i_scratch[1], my_l);
    __syncthreads();
    int cm = S.i_scratch[1];
    if (cm == 0) return;
    if (threadIdx.x == 0 && cm > S.best_l) { S.best_l = cm; S.cnt = 0; S.zero_in = 0; S.best_key = ~0ull; }
    __syncthreads();
    if (cm == S.best_l) {
        bool is = my_l == cm;
        if (is) {
            atomicAdd(&S.cnt, 1);
            if (p == 0) S.zero_in = 1;
            else {
                i64 d = p - (i64)e; if (d < 0) d = -d;
                atomicMin(&S.best_key, ((unsigned long long)d << 32) | (unsigned long long)(u32)p);
            }
        }
    }
    __syncthreads();
}

// every candidate of the k-mer T[j..j+k) from the sorted index folded with prev_match_end = e (pn1 / ln1 of :124-129)
__device__ __forceinline__ void fold_index_candidates(GpShared& S, const GpArgs& a, i64 j, int e) {
    fold_reset(S);
    u32 h = kmer_hash_words(ld_unaligned64(a.T + j), ld_unaligned64(a.T + j + 8), a.k);
    i64 lo, hi;
    index_range(a, h, &lo, &hi);                                 // uniform, every thread computes it
    for (i64 base = lo; base < hi; base += GP_T) {
        const i64 idx = base + threadIdx.x;
        fold_chunk(S, a, idx < hi ? (i64)a.vals[idx] : -1, j, e);  // (fold_chunk extends from offset 0: entries of other k-mers drop out there)
    }
}

// the same fold over the occurrence list of the FIRST target k-mer (sampled mode; the fold is order-independent)
__device__ __forceinline__ void fold_first_candidates(GpShared& S, const GpArgs& a, u32 n, i64 j, int e) {
    fold_reset(S);
    for (u32 base = 0; base < n; base += GP_T) {
        const u32 idx = base + threadIdx.x;
        fold_chunk(S, a, idx < n ? (i64)a.first_list[idx] : -1, j, e);
    }
}

__device__ __forceinline__ int fold_result_p(const GpShared& S) {
    return (S.cnt == 1 && S.zero_in) ? 0 : (int)(S.best_key & 0xffffffffULL);
}

// One step of the parse loop (:64-161) from state (j, e): scans [j, scan_end) for the first position that yields a match
// (everything before it is literal, :77-96).  Found: j = that position, (sel_p, sel_l) = the chosen candidate, returns
// true.  Not found: j = scan_end (state e unchanged), returns false.  All threads, uniform arguments.
__device__ __forceinline__ bool gp_step(GpShared& S, const GpArgs& a, i64& j, int e, i64 scan_end, int& sel_p, int& sel_l) {
    const int tid = (int)threadIdx.x;
    const int k = a.k;
    i64 found = -1;
    if (e == -1 && !a.full) {
        // ---- sampled mode, no match yet: the first position whose k-mer occurs anywhere in the reference (:77) and its
        //      occurrences were found by brute force (position 0 for the usual pair, else one of the next GP_FIRST_W - 1);
        //      a target that begins with more unmatched symbols than that needs the full index
        u32 n = a.first[GP_FIRST_N0];
        i64 fj = 0;
        if (n == 0u && a.first[GP_FIRST_J] != 0xffffffffu) { fj = (i64)a.first[GP_FIRST_J]; n = a.first[GP_FIRST_N1]; }
        if (j > fj || n == 0u || n > GP_FIRST_CAP) {
            if (tid == 0) *reinterpret_cast<volatile u32*>(a.need_full) = 1u;
            j = scan_end;
            return false;
        }
        if (fj >= scan_end) { j = scan_end; return false; }      // literal steps, the state stays (.., -1)
        j = fj;
        fold_first_candidates(S, a, n, j, e);
        sel_p = fold_result_p(S); sel_l = S.best_l;
        __syncthreads();
        return true;
    }
    if (e == -1) {
        // ---- no match yet: first position whose k-mer occurs anywhere in the reference (:77, all candidates in range :87)
        for (; j < scan_end; j += GP_T) {
            i64 pos = j + tid;
            bool hit = false;
            if (pos < scan_end) {
                u64 w0 = ld_unaligned64(a.T + pos), w1 = ld_unaligned64(a.T + pos + 8);
                u32 h = kmer_hash_words(w0, w1, k);
                i64 lo, hi;
                index_range(a, h, &lo, &hi);
                for (; lo < hi && !hit; ++lo) {
                    const u8* rp = a.R + a.vals[lo];
                    hit = kmer_equal_words(ld_unaligned64(rp), ld_unaligned64(rp + 8), w0, w1, k);
                }
            }
            if (tid == 0) S.i_scratch[4] = 0x7fffffff;
            if (__syncthreads_or(hit)) {
                if (hit) atomicMin(&S.i_scratch[4], tid);
                __syncthreads();
                found = j + S.i_scratch[4];
                __syncthreads();
                break;
            }
        }
        if (found < 0) { j = scan_end; return false; }
        j = found;
        fold_index_candidates(S, a, j, e);
        sel_p = fold_result_p(S); sel_l = S.best_l;
        __syncthreads();
        return true;
    }
    // ---- banded state: candidates must satisfy |p - e| <= m (:87, :116)
    i64 wlo = (i64)e - a.m; if (wlo < 0) wlo = 0;
    i64 whi = (i64)e + a.m; if (whi > a.nr - k) whi = a.nr - k;
    if (whi < wlo) { j = scan_end; return false; }            // no reference k-mer can be in range: literals only
    const int wlen = (int)(whi - wlo + 1);
    const int wbytes = wlen + k - 1;
    // (the target words of the brute-force step are requested before the window is staged: one memory latency instead of two)
    const u64 bt0 = ld_unaligned64(a.T + j), bt1 = ld_unaligned64(a.T + j + 8), bt2 = ld_unaligned64(a.T + j + 16);
    for (int x = tid; x < GP_FILTER; x += GP_T) S.f_hash[x] = 0u;
    for (int x = tid; x < GP_WIN_BYTES; x += GP_T) S.win[x] = x < wbytes ? a.R[wlo + x] : (u8)0;
    if (tid < GP_BF) S.bf_hit[tid] = 0;
    __syncthreads();
    // The first GP_BF positions by brute force: one thread per window position compares its k-mer with the k-mers at
    // j .. j + GP_BF - 1.  After a substitution the parse resumes one or two positions later on the same diagonal, and
    // then neither the filter nor the scan round is needed (the threads that hit ARE the candidate list).
    bool have_cand = false;
    i64 cand = -1;
    if (GP_T >= 2 * GP_MAX_M + 1) {
        const int np = scan_end - j < GP_BF ? (int)(scan_end - j) : GP_BF;
        u32 mymask = 0u;
        if (tid < wlen) {
            const u64 t0 = bt0, t1 = bt1, t2 = bt2;
            const u64 w0 = ld_unaligned64(S.win + tid), w1 = ld_unaligned64(S.win + tid + 8);
#pragma unroll
            for (int q = 0; q < GP_BF; ++q) {
                const u64 a0 = q ? (t0 >> (8 * q)) | (t1 << (64 - 8 * q)) : t0;
                const u64 a1 = q ? (t1 >> (8 * q)) | (t2 << (64 - 8 * q)) : t1;
                if (q < np && kmer_equal_words(w0, w1, a0, a1, k)) { mymask |= 1u << q; S.bf_hit[q] = 1; }
            }
        }
        __syncthreads();
        int first = -1;
#pragma unroll
        for (int q = GP_BF - 1; q >= 0; --q) if (q < np && S.bf_hit[q]) first = q;
        if (first >= 0) {
            found = j + first;
            have_cand = true;
            if ((mymask >> first) & 1u) cand = wlo + tid;
        } else {
            j += np;                                           // no candidate in range at these positions: literals
            if (j >= scan_end) { j = scan_end; __syncthreads(); return false; }       // (bf_hit is rewritten by the next call)
        }
    }
    if (!have_cand) {
        for (int x = tid; x < wlen; x += GP_T) {
            u32 h = kmer_hash_words(ld_unaligned64(S.win + x), ld_unaligned64(S.win + x + 8), k) | 1u;
            u32 slot = (h >> 1) & (GP_FILTER - 1);
            while (atomicCAS(&S.f_hash[slot], 0u, h) != 0u) slot = (slot + 1) & (GP_FILTER - 1);
            S.f_off[slot] = (u8)x;
        }
        __syncthreads();
        for (; j < scan_end; j += GP_T) {
            i64 pos = j + tid;
            bool hit = false;
            if (pos < scan_end) {
                u64 w0 = ld_unaligned64(a.T + pos), w1 = ld_unaligned64(a.T + pos + 8);
                u32 h = kmer_hash_words(w0, w1, k) | 1u;
                u32 slot = (h >> 1) & (GP_FILTER - 1);
                for (u32 fh; (fh = S.f_hash[slot]) != 0u && !hit; slot = (slot + 1) & (GP_FILTER - 1)) {
                    if (fh == h) {
                        const u8* wp = S.win + S.f_off[slot];
                        hit = kmer_equal_words(ld_unaligned64(wp), ld_unaligned64(wp + 8), w0, w1, k);
                    }
                }
            }
            if (tid == 0) S.i_scratch[4] = 0x7fffffff;
            if (__syncthreads_or(hit)) {
                if (hit) atomicMin(&S.i_scratch[4], tid);
                __syncthreads();
                found = j + S.i_scratch[4];
                __syncthreads();
                break;
            }
        }
        if (found < 0) { j = scan_end; return false; }
    }
    j = found;
    // in-range candidates: the window positions whose k-mer equals T[j..j+k)  (pn2 / ln2, :116-123)
    if (have_cand && __syncthreads_count(cand >= 0) == 1) {
        // the usual step (the parse resumes on its diagonal behind a substitution): ONE candidate, whose k-mer the brute-force
        // compare has verified -- it is chosen whatever its length (:118-122 with a single p), so the fold and its barriers
        // are skipped and the whole CTA extends it at once
        if (cand >= 0) { S.i_scratch[2] = (int)(cand & 0xffffffff); S.i_scratch[3] = (int)(cand >> 32); }
        __syncthreads();
        const i64 pp = ((i64)S.i_scratch[3] << 32) | (u32)S.i_scratch[2];
        const i64 maxl = (a.nr - pp) < (a.nt - j) ? (a.nr - pp) : (a.nt - j);
        sel_p = (int)pp;
        sel_l = (int)block_lcp(S, a.R, pp, a.T, j, maxl);
        if (sel_p == 0 && !a.full) {
            if (tid == 0) *reinterpret_cast<volatile u32*>(a.need_full) = 1u;
        } else if (sel_p == 0) {
            fold_index_candidates(S, a, j, e);
            sel_p = fold_result_p(S); sel_l = S.best_l;
            __syncthreads();
        }
        return true;
    }
    fold_reset(S);
    if (have_cand) {
        fold_chunk(S, a, cand, j, e);
    } else {
        const u64 w0 = ld_unaligned64(a.T + j), w1 = ld_unaligned64(a.T + j + 8);
        for (int x0 = 0; x0 < wlen; x0 += GP_T) {                // the fold is order-independent: one thread-wide slice at a time
            const int x = x0 + tid;
            i64 cd = -1;
            if (x < wlen && kmer_equal_words(ld_unaligned64(S.win + x), ld_unaligned64(S.win + x + 8), w0, w1, k)) cd = wlo + x;
            fold_chunk(S, a, cd, j, e);
        }
    }
    sel_p = fold_result_p(S); sel_l = S.best_l;
    __syncthreads();
    if (sel_p == 0 && !a.full) {                              // (sampled mode cannot serve this: the result of this parse will be discarded)
        if (tid == 0) *reinterpret_cast<volatile u32*>(a.need_full) = 1u;
    } else if (sel_p == 0) {                                  // `pn2 != 0` fails (:134): unrestricted best over ALL candidates
        fold_index_candidates(S, a, j, e);
        sel_p = fold_result_p(S); sel_l = S.best_l;
        __syncthreads();
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Chunk-speculative execution of the sequential parse.
//
// The parse is a deterministic function of its state (index j, prev_match_end e).  The target is cut into chunks of
// GP_CHUNK positions; gp_spec_k parses every chunk in parallel (one CTA each) from a GUESSED entry state until the
// index leaves the chunk, recording its matches and exit state.  gp_front_k then walks the true state from (0, -1):
// whenever the true state equals a chunk's entry guess or the state after one of its recorded matches, the rest of
// that chunk's speculative output is provably what the sequential parse would produce and is accepted wholesale;
// otherwise the front executes exact steps itself until the states meet again.  Two guesses per chunk:
//   slot 0 (DIAG): the parse is inside a match on the diagonal d found by looking one k-mer near the chunk start up in
//                  the index; the entry state is then (first mismatch on that diagonal after the boundary, its e);
//   slot 1 (LOST): the parse reaches the boundary with a given e and no usable candidate ("lost" after a rearrangement);
//                  re-issued by the host every time the front gets lost with a new e.
// ------------------------------------------------------------------------------------------------
static const int GP_CHUNK_DEFAULT = 8192;       // positions per speculative chunk (measured: 21.6 ms vs 29.6 ms at 32768 on the divergent chr21-shaped pair) (SCCG_GP_CHUNK overrides it: tests use tiny chunks)
static const int GP_DIAG_PROBES = 64;           // must stay 64 (vote encoding)
static const int GP_WARM = 128;                 // positions before its boundary at which a chunk's diagonal guess is anchored

struct GpChunkInfo { i64 entry_j; i64 exit_j; int entry_e; int exit_e; u32 count; int valid; };

struct GpSpecArgs {
    GpArgs a;
    GpChunkInfo* info;          // [2][nchunks]
    int* c_tpos; int* c_p; int* c_l;   // [2][nchunks][cap_c]
    u32 nchunks, cap_c;
    int chunk;                  // positions per chunk
    int slot;                   // 0 DIAG, 1 LOST
    u32 first_chunk;            // chunks >= first_chunk are (re)computed
    int lost_e;                 // slot 1: the e every chunk is entered with
    u32* ctl;                   // ctl[GP_CTL_CANCEL]: the front got lost, speculation is pointless
};
enum { GP_CTL_CANCEL = 0, GP_CTL_NEED_FULL = 1 };
// GpChunkInfo::valid: 0 = not computed yet (the front may be running concurrently with the second batch of chunks), 1 = usable,
// 2 = computed but unusable (no diagonal guess, or cancelled).  Written LAST, behind a fence; the front reads chunk records
// with L2 loads (ld_info) and only trusts the other fields once it has seen valid != 0.
enum { GP_CHUNK_PENDING = 0, GP_CHUNK_OK = 1, GP_CHUNK_UNUSABLE = 2 };
__device__ __forceinline__ void gp_publish(GpChunkInfo* info, int valid) {
    __threadfence();
    *reinterpret_cast<volatile int*>(&info->valid) = valid;
}
__device__ __forceinline__ GpChunkInfo ld_info(const GpChunkInfo* p) {
    static_assert(sizeof(GpChunkInfo) == 32, "two 16-byte halves");
    GpChunkInfo ci;
    const int4 b = __ldcg(reinterpret_cast<const int4*>(p) + 1);          // entry_e, exit_e, count, valid
    ci.entry_e = b.x; ci.exit_e = b.y; ci.count = (u32)b.z; ci.valid = b.w;
    ci.entry_j = 0; ci.exit_j = 0;
    if (b.w == GP_CHUNK_OK) {
        __threadfence();
        const int4 a = __ldcg(reinterpret_cast<const int4*>(p));
        ci.entry_j = (i64)(((u64)(u32)a.y << 32) | (u32)a.x); ci.exit_j = (i64)(((u64)(u32)a.w << 32) | (u32)a.z);
    }
    return ci;
}

#ifndef SCCG_GP_MINB
#define SCCG_GP_MINB 6            // 256-thread CTAs: 48 warps/SM at <= 40 registers (measured 4 / 5 / 6 CTAs: 2.60 / 2.51 / 2.47 ms on the gap pair)
#endif
__global__ void __launch_bounds__(GP_T, SCCG_GP_MINB) gp_spec_k(GpSpecArgs s) {
    __shared__ GpShared S;
    const GpArgs& a = s.a;
    const int tid = (int)threadIdx.x;
    const u32 c = s.first_chunk + blockIdx.x;
    if (c >= s.nchunks) return;
    const i64 B = (i64)c * s.chunk, last_j = a.nt - a.k;
    const i64 j_stop = B + s.chunk;
    GpChunkInfo* info = s.info + (size_t)s.slot * s.nchunks + c;
    const size_t base = ((size_t)s.slot * s.nchunks + c) * s.cap_c;
    i64 j = B;
    int e = s.lost_e;
    bool valid = true;
    // the front (running concurrently with the second batch) got lost: everything speculated from here on would be thrown
    // away (a "lost" parse only comes back by a chance hit, with another e).  The chunk stays invalid (info is zeroed).
    if (__syncthreads_or(tid == 0 && *reinterpret_cast<volatile u32*>(s.ctl + GP_CTL_CANCEL) != 0u)) {
        if (tid == 0) gp_publish(info, GP_CHUNK_UNUSABLE);
        return;
    }
    if (s.slot == 0 && c == 0) {
        e = -1;                                                   // the true initial state (0, -1): not a guess
    } else if (s.slot == 0) {
        // diagonal guess: every probe position whose k-mer occurs exactly once in the reference votes for its diagonal; the
        // most frequent diagonal wins if probes at least 16 positions apart agree on it (a k-mer hit by a mutation -- or a k-mer
        // of an inserted stretch -- may be unique somewhere else: one vote, or a few adjacent ones, prove nothing).  Probes come in rounds of 64 positions from the
        // boundary on, so a chunk that begins inside an insertion still finds the diagonal of its homologous part.
        i64* votes = reinterpret_cast<i64*>(S.long_list);         // GP_DIAG_PROBES diagonals (scratch reuse)
        i64 Q = B;                                                // first position of the winning round
        i64 d = 0;
        valid = false;
        for (; Q < j_stop && Q <= last_j; Q += GP_DIAG_PROBES) {
            i64 my_d = 0;
            bool have = false;
            if (tid < GP_DIAG_PROBES && Q + tid <= last_j) {
                i64 pos = Q + tid;
                u64 w0 = ld_unaligned64(a.T + pos), w1 = ld_unaligned64(a.T + pos + 8);
                u32 h = kmer_hash_words(w0, w1, a.k);
                i64 lo = index_lower_bound(a, h);                     // entries of one key are contiguous: no second search for the end
                const i64 hi = (i64)a.bucket[(h >> a.bucket_shift) + 1u];
                int hits = 0;
                for (; lo < hi && hits < 2 && a.keys[lo] == h; ++lo) {
                    const u8* rp = a.R + a.vals[lo];
                    if (kmer_equal_words(ld_unaligned64(rp), ld_unaligned64(rp + 8), w0, w1, a.k)) { ++hits; my_d = (i64)a.vals[lo] - pos; }
                }
                have = hits == 1;
            }
            if (tid < GP_DIAG_PROBES) votes[tid] = have ? my_d : (i64)0x7fffffffffffffffLL;
            if (tid == 0) S.i_scratch[5] = -1;
            __syncthreads();
            int score = -1, first = GP_DIAG_PROBES, last = -1;
            if (have) { score = 0; for (int x = 0; x < GP_DIAG_PROBES; ++x) if (votes[x] == my_d) { ++score; if (x < first) first = x; last = x; } }
            // agreeing probes at least 16 positions apart: a chance match of k+1 or k+2 symbols cannot fake that
            if (score > 1 && last - first >= 16) atomicMax(&S.i_scratch[5], score * 64 + (63 - tid));   // most votes, then the earliest probe
            __syncthreads();
            const int best = S.i_scratch[5];
            if (best >= 0) { d = votes[63 - (best & 63)]; valid = true; }
            __syncthreads();
            if (valid) break;
        }
        if (valid) {
            // The guess is anchored GP_WARM positions BEFORE the boundary ("the parse is inside a match on diagonal d there") and
            // replayed up to it: two parses on one diagonal meet at the end of the next common match, so by the boundary the
            // replay has normally become the true parse even where the boundary falls between two matches (divergent pairs: a
            // third of the chunks otherwise fail to splice and cost the sequential front exact steps).  A diagonal that was only
            // found further inside the chunk is anchored where it was found.
            i64 Bw = Q > B ? Q : B - GP_WARM;
            if (Bw < 0) Bw = 0;
            if (Bw + d < 0) Bw = -d;
            i64 p0 = Bw + d;
            if (Q + d < 0 || p0 >= a.nr || Bw > Q) valid = false;
            else {
                i64 maxl = (a.nr - p0) < (a.nt - Bw) ? (a.nr - p0) : (a.nt - Bw);
                i64 l = block_lcp(S, a.R, p0, a.T, Bw, maxl);      // the match covering the anchor ends at its first mismatch
                j = Bw + l;
                i64 ee = j + d - 1;
                if (ee < 0 || ee > 0x7fffffff) valid = false; else e = (int)ee;
            }
        }
        __syncthreads();
        // replay up to the boundary; nothing is recorded (those matches belong to the previous chunk)
        while (valid && j < B && j <= last_j) {
            int sel_p = 0, sel_l = 0;
            const i64 scan_end = B < last_j + 1 ? B : last_j + 1;
            if (!gp_step(S, a, j, e, scan_end, sel_p, sel_l)) break;     // no candidate before the boundary: j == scan_end
            e = sel_p + sel_l - 1;
            j += sel_l;
        }
    }
    u32 n = 0;
    const i64 entry_j = j;
    const int entry_e = e;
    if (valid) {
        while (j < j_stop && j <= last_j) {
            int sel_p = 0, sel_l = 0;
            i64 scan_end = j_stop < last_j + 1 ? j_stop : last_j + 1;
            if (!gp_step(S, a, j, e, scan_end, sel_p, sel_l)) break;
            if (tid == 0) { s.c_tpos[base + n] = (int)j; s.c_p[base + n] = sel_p; s.c_l[base + n] = sel_l; }
            ++n;
            e = sel_p + sel_l - 1;
            j += sel_l;
            if ((n & 15u) == 0u && __syncthreads_or(tid == 0 && *reinterpret_cast<volatile u32*>(s.ctl + GP_CTL_CANCEL) != 0u)) { valid = false; break; }
        }
    }
    if (tid == 0) {
        info->entry_j = entry_j; info->entry_e = entry_e; info->exit_j = j; info->exit_e = e; info->count = n;
        gp_publish(info, valid ? GP_CHUNK_OK : GP_CHUNK_UNUSABLE);          // (the matches were written by this thread too: ordered by the fence)
    }
}

// pieces of the final match list, in order: a run of the front's own matches or a suffix of a chunk's speculative list
struct GpPiece { u32 src; u32 first; u32 count; };          // src: 0xffffffff = front buffer, else slot * nchunks + chunk
struct GpFrontState { i64 j; int e; int status; u32 npieces; u32 nfront; u32 rounds; int lost_e; u32 steps; u32 spliced; };     // rounds: front launches that did something
enum { GP_RUNNING = 0, GP_DONE = 1, GP_LOST = 2, GP_FULL = 3 };
static const int GP_LOST_STREAK = 3;            // chunk remainders without any usable candidate before the front asks for a scan of the rest

__device__ __forceinline__ unsigned long long gp_now_ns() {
#if defined(__CUDA_ARCH__)
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
#else
    return 0ull;
#endif
}
// development aid (SCCG_GP_TRACE=1): trace[0] = entries used, then (kind, begin ns, end ns, info) per front / scan launch
static const int GP_TRACE_MAX = 120;
__device__ __forceinline__ void gp_trace(unsigned long long* trace, unsigned long long kind, unsigned long long t0, unsigned long long info) {
    if (!trace) return;
    const unsigned long long i = atomicAdd(trace, 1ull);
    if (i < (unsigned long long)GP_TRACE_MAX) { trace[1 + 4 * i] = kind; trace[2 + 4 * i] = t0; trace[3 + 4 * i] = gp_now_ns(); trace[4 + 4 * i] = info; }
}

struct GpFrontArgs {
    GpSpecArgs s;
    GpFrontState* st;
    GpPiece* pieces; u32 cap_pieces;
    int* f_tpos; int* f_p; int* f_l;       // the front's own matches
    unsigned long long* d_hit;             // result of gp_lost_scan_k for the state the previous launch got lost in (the front resets it when it gets lost)
    unsigned long long* trace;             // development aid, usually NULL
};

__global__ void __launch_bounds__(GP_T) gp_front_k(GpFrontArgs f) {
    __shared__ GpShared S;
    const GpArgs& a = f.s.a;
    const int tid = (int)threadIdx.x;
    const i64 last_j = a.nt - a.k;
    i64 j = f.st->j;
    int e = f.st->e;
    // launches are queued ahead of the host's knowledge of the state: a front after the end of the parse returns at once
    // (uniform: every thread reads the same word, nobody has written yet)
    const int status_in = f.st->status;
    if (status_in == GP_DONE || status_in == GP_FULL) return;
    const unsigned long long t_begin = f.trace ? gp_now_ns() : 0ull;
    u32 npieces = f.st->npieces, nfront = f.st->nfront, steps = f.st->steps, spliced = f.st->spliced;
    const u32 rounds = f.st->rounds + 1u;
    int lost_streak = 0;
    int status = GP_RUNNING;
    const bool was_lost = status_in == GP_LOST;                // chance-hit regime: declare "lost" again quickly
    if (was_lost) {
        // the previous launch ended "lost" in state (j, e) and gp_lost_scan_k searched the rest of the target for the next
        // position whose k-mer occurs in the window of e: everything before it is a literal step (:83-96), state unchanged
        const unsigned long long hit = *f.d_hit;
        j = hit == ~0ull ? last_j + 1 : (i64)hit;
    }
    __syncthreads();
    const GpChunkInfo* inf = f.s.info;
    while (true) {
        if (j > last_j) { status = GP_DONE; break; }
        const u32 c = (u32)(j / f.s.chunk);
        // The chunk the front stands in may still be in the making (second batch, other stream): while the parse is
        // synchronised it is worth waiting for -- its CTA is running or about to -- unless speculation has been cancelled.
        if (tid == 0 && lost_streak == 0) {
            while (__ldcg(&inf[c].valid) == GP_CHUNK_PENDING && *reinterpret_cast<volatile u32*>(f.s.ctl + GP_CTL_CANCEL) == 0u) __nanosleep(200);
        }
        __syncthreads();
        // ---- bulk splice: thread t checks chunk c + t; a run of chunks whose entry state equals the exit state of its
        //      predecessor (the first one: the true state) is accepted in one go
        bool took = false;
        {
            const u32 cc = c + (u32)tid;
            bool ok = false;
            u32 my_count = 0;
            if (cc < f.s.nchunks) {
                const GpChunkInfo ci = ld_info(inf + cc);
                if (ci.valid == GP_CHUNK_OK) {
                    my_count = ci.count;
                    if (tid == 0) ok = ci.entry_j == j && ci.entry_e == e;
                    else {
                        const GpChunkInfo pv = ld_info(inf + cc - 1);
                        ok = pv.valid == GP_CHUNK_OK && ci.entry_j == pv.exit_j && ci.entry_e == pv.exit_e && (u32)(pv.exit_j / f.s.chunk) == cc;
                    }
                }
            }
            if (tid == 0) S.i_scratch[5] = GP_T;
            __syncthreads();
            if (!ok) atomicMin(&S.i_scratch[5], tid);
            __syncthreads();
            const int run = S.i_scratch[5];                        // chunks c .. c + run - 1 splice
            __syncthreads();
            if (run > 0) {
                if (npieces + (u32)run + 2 >= f.cap_pieces) { status = GP_FULL; break; }
                if (tid < run) { GpPiece pc; pc.src = cc; pc.first = 0; pc.count = my_count; f.pieces[npieces + tid] = pc; }
                npieces += (u32)run;
                const GpChunkInfo lastc = ld_info(inf + c + run - 1);
                j = lastc.exit_j; e = lastc.exit_e;
                took = true; spliced += (u32)run;
                lost_streak = 0;
                __syncthreads();
            }
        }
        if (took) continue;
        // ---- splice into the middle of a speculative run: does chunk c pass through the true state (j, e) after one of its matches?
        {
            const GpChunkInfo ci = ld_info(inf + c);
            if (ci.valid == GP_CHUNK_OK) {
                const size_t base = (size_t)c * f.s.cap_c;
                int first = -1;
                if (ci.count) {
                    // the state after match i is (tpos + l, p + l - 1); tpos + l is increasing: binary search
                    int lo = 0, hi = (int)ci.count;
                    while (lo < hi) { int mid = (lo + hi) >> 1; if ((i64)__ldcg(f.s.c_tpos + base + mid) + __ldcg(f.s.c_l + base + mid) < j) lo = mid + 1; else hi = mid; }
                    if (lo < (int)ci.count) {
                        const int tp = __ldcg(f.s.c_tpos + base + lo), ll = __ldcg(f.s.c_l + base + lo), pp = __ldcg(f.s.c_p + base + lo);
                        if ((i64)tp + ll == j && pp + ll - 1 == e) first = lo + 1;
                    }
                }
                if (first >= 0) {
                    if ((u32)first < ci.count) {
                        if (tid == 0) { GpPiece pc; pc.src = c; pc.first = (u32)first; pc.count = ci.count - (u32)first; f.pieces[npieces] = pc; }
                        ++npieces;
                    }
                    j = ci.exit_j; e = ci.exit_e;
                    took = true; ++spliced;
                    lost_streak = 0;
                }
            }
        }
        if (took) { if (npieces + 2 >= f.cap_pieces) { status = GP_FULL; break; } continue; }
#ifdef SCCG_EMU_TRACE
        if (tid == 0) { const GpChunkInfo ci = ld_info(inf + c); const GpChunkInfo cn = c + 1 < f.s.nchunks ? ld_info(inf + c + 1) : ci;
            printf("front: no splice at j=%lld e=%d chunk %u: valid=%d entry=(%lld,%d) exit=(%lld,%d) count=%u | next entry=(%lld,%d) valid=%d\n", (long long)j, e, c, ci.valid, (long long)ci.entry_j, ci.entry_e, (long long)ci.exit_j, ci.exit_e, ci.count, (long long)cn.entry_j, cn.entry_e, cn.valid); }
#endif
        // ---- exact step of the sequential parse, at most to the end of this chunk
        int sel_p = 0, sel_l = 0;
        i64 scan_end = (i64)(c + 1) * f.s.chunk;
        if (was_lost && scan_end > j + GP_T) scan_end = j + GP_T;    // after a chance hit: look one round ahead, then hand the search back to the scan kernel
        if (scan_end > last_j + 1) scan_end = last_j + 1;
        const i64 j_before = j;
        ++steps;
        if (gp_step(S, a, j, e, scan_end, sel_p, sel_l)) {
            if (tid == 0) {
                f.f_tpos[nfront] = (int)j; f.f_p[nfront] = sel_p; f.f_l[nfront] = sel_l;
                if (npieces && f.pieces[npieces - 1].src == 0xffffffffu && f.pieces[npieces - 1].first + f.pieces[npieces - 1].count == nfront)
                    f.pieces[npieces - 1].count++;
            }
            __syncthreads();
            bool extend = npieces && f.pieces[npieces - 1].src == 0xffffffffu && f.pieces[npieces - 1].first + f.pieces[npieces - 1].count == nfront + 1;
            if (!extend) {
                if (tid == 0) { GpPiece pc; pc.src = 0xffffffffu; pc.first = nfront; pc.count = 1; f.pieces[npieces] = pc; }
                ++npieces;
            }
            __syncthreads();
            ++nfront;
            e = sel_p + sel_l - 1;
            j += sel_l;
            lost_streak = 0;
        } else if (e == -1 && !a.full) {
            // sampled mode: literal steps up to the first matching position -- or that position is out of reach: need_full is
            // raised, the host repeats the parse with the full index
            __syncthreads();
            if (*reinterpret_cast<volatile u32*>(a.need_full) != 0u) {
                if (tid == 0) *reinterpret_cast<volatile u32*>(f.s.ctl + GP_CTL_CANCEL) = 1u;
                status = GP_DONE; break;
            }
        } else if (e != -1) {
            // no usable candidate up to the end of the chunk.  A few chunks are walked exactly (an insertion of a few kb in the
            // target); then the parse counts as "lost" with this e and the rest of the target is searched by many CTAs at once
            (void)j_before;
            if (++lost_streak >= (was_lost ? 1 : GP_LOST_STREAK) && j <= last_j) {
                if (tid == 0) *reinterpret_cast<volatile u32*>(f.s.ctl + GP_CTL_CANCEL) = 1u;     // whatever is still being speculated is useless now
                status = GP_LOST; break;
            }
        }
        if (npieces + 2 >= f.cap_pieces) { status = GP_FULL; break; }
    }
    if (tid == 0) {
        f.st->j = j; f.st->e = e; f.st->status = status; f.st->npieces = npieces; f.st->nfront = nfront; f.st->steps = steps; f.st->spliced = spliced;
        f.st->rounds = rounds; f.st->lost_e = e;
        if (status == GP_LOST) *f.d_hit = ~0ull;                   // gp_lost_scan_k (next in the stream) lowers it to the first hit
        gp_trace(f.trace, 1ull, t_begin, ((unsigned long long)steps << 32) | (unsigned long long)(u32)status);
    }
}

// ------------------------------------------------------------------------------------------------
// "Lost" scan.  From a state (j0, e) without a usable candidate nearby the next event of the parse is the first position
// j >= j0 whose k-mer occurs in the window R'[e-m .. e+m+k) (:83-96: every position before it is a literal step and the
// state does not change).  On a divergent pair that position is a chance hit a million symbols ahead, so instead of one CTA
// walking there, many CTAs search the rest of the target in interleaved tiles, nearest tiles first; the smallest hit wins
// (atomicMin) and tiles beyond it are not looked at.  The front resumes exactly there.
// ------------------------------------------------------------------------------------------------
struct GpScanArgs { GpArgs a; const GpFrontState* st; unsigned long long* d_hit; unsigned long long* trace; };
#ifndef SCCG_GP_SCAN_ROUNDS
#define SCCG_GP_SCAN_ROUNDS 4            // positions per thread and tile (measured on the divergent pair, same positions per wave: 16 rounds x 4 CTAs/SM 2.22 ms of parse, 4 x 16: 2.11)
#endif
static const int GP_SCAN_ROUNDS = SCCG_GP_SCAN_ROUNDS;
static const int GP_SCAN_TILE = GP_T * GP_SCAN_ROUNDS;

__global__ void __launch_bounds__(GP_T) gp_lost_scan_k(GpScanArgs s) {
    __shared__ GpShared S;
    const GpArgs& a = s.a;
    const int tid = (int)threadIdx.x, k = a.k;
    if (s.st->status != GP_LOST) return;                      // queued behind a front that did not get lost
    const unsigned long long t_begin = s.trace ? gp_now_ns() : 0ull;
    const i64 j0 = s.st->j, last_j = a.nt - k;
    const int e = s.st->e;
    i64 wlo = (i64)e - a.m; if (wlo < 0) wlo = 0;
    i64 whi = (i64)e + a.m; if (whi > a.nr - k) whi = a.nr - k;
    if (whi < wlo) return;                                   // no reference k-mer can be in range: literals to the end
    const int wlen = (int)(whi - wlo + 1), wbytes = wlen + k - 1;
    for (int x = tid; x < GP_FILTER; x += GP_T) S.f_hash[x] = 0u;
    for (int x = tid; x < GP_WIN_BYTES; x += GP_T) S.win[x] = x < wbytes ? a.R[wlo + x] : (u8)0;
    __syncthreads();
    for (int x = tid; x < wlen; x += GP_T) {
        u32 h = kmer_hash_words(ld_unaligned64(S.win + x), ld_unaligned64(S.win + x + 8), k) | 1u;
        u32 slot = (h >> 1) & (GP_FILTER - 1);
        while (atomicCAS(&S.f_hash[slot], 0u, h) != 0u) slot = (slot + 1) & (GP_FILTER - 1);
        S.f_off[slot] = (u8)x;
    }
    __syncthreads();
    for (i64 tile = blockIdx.x;; tile += gridDim.x) {
        const i64 t0 = j0 + tile * GP_SCAN_TILE;
        if (t0 > last_j || (unsigned long long)t0 >= *reinterpret_cast<volatile unsigned long long*>(s.d_hit)) break;
#pragma unroll 4
        for (int r = 0; r < GP_SCAN_ROUNDS; ++r) {
            const i64 pos = t0 + (i64)r * GP_T + tid;
            if (pos > last_j) break;
            const u64 w0 = ld_unaligned64(a.T + pos), w1 = ld_unaligned64(a.T + pos + 8);
            const u32 h = kmer_hash_words(w0, w1, k) | 1u;
            u32 slot = (h >> 1) & (GP_FILTER - 1);
            for (u32 fh; (fh = S.f_hash[slot]) != 0u; slot = (slot + 1) & (GP_FILTER - 1)) {
                if (fh == h) {
                    const u8* wp = S.win + S.f_off[slot];
                    if (kmer_equal_words(ld_unaligned64(wp), ld_unaligned64(wp + 8), w0, w1, k)) { atomicMin(s.d_hit, (unsigned long long)pos); break; }
                }
            }
        }
    }
    if (s.trace && blockIdx.x == 0 && tid == 0) gp_trace(s.trace, 2ull, t_begin, (unsigned long long)j0);
}

// Sampled index mode, T[0..k) does not occur in the reference: the smallest j in [1, GP_FIRST_W) whose k-mer does (:77-81:
// the positions before it are literal steps).  Roles swapped against gp_lost_scan_k: the filter holds the target k-mers
// (keyed by their first 8 symbols: two 4-byte windows and two multiplications per reference position instead of a k-mer
// hash), all SMs run over the reference, 16 positions per thread.  Does nothing when gp_first_k found occurrences.
__device__ __forceinline__ u32 first_key(u32 wa, u32 wb) { return (wa * 0x9E3779B1u + wb * 0x85EBCA77u) | 1u; }
__global__ void __launch_bounds__(GP_T) gp_first_scan_k(GpArgs a, u32* first) {
    __shared__ GpShared S;
    if (first[GP_FIRST_N0] != 0u) return;
    const int tid = (int)threadIdx.x, k = a.k;
    const i64 last_j = a.nt - k, nk = a.nr - k + 1;
    const int w = last_j + 1 < GP_FIRST_W ? (int)(last_j + 1) : GP_FIRST_W;
    for (int x = tid; x < GP_FILTER; x += GP_T) S.f_hash[x] = 0u;
    __syncthreads();
    for (int x = 1 + tid; x < w; x += GP_T) {
        const u64 t0 = ld_unaligned64(a.T + x);
        const u32 h = first_key((u32)t0, (u32)(t0 >> 32));
        u32 slot = (h >> 23) & (GP_FILTER - 1);
        while (atomicCAS(&S.f_hash[slot], 0u, h) != 0u) slot = (slot + 1) & (GP_FILTER - 1);
        S.f_off[slot] = (u8)x;
    }
    __syncthreads();
    for (i64 p0 = ((i64)blockIdx.x * GP_T + tid) * 16; p0 < nk; p0 += (i64)gridDim.x * GP_T * 16) {
        u32 x[6];
        first_windows(a.R, p0, x);
        u32 wv[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) wv[i] = (i & 3) ? __funnelshift_r(x[i >> 2], x[(i >> 2) + 1], 8 * (i & 3)) : x[i >> 2];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const u32 h = first_key(wv[i], wv[i + 4]);
            u32 slot = (h >> 23) & (GP_FILTER - 1);
            for (u32 fh; (fh = S.f_hash[slot]) != 0u; slot = (slot + 1) & (GP_FILTER - 1)) {      // every entry: equal k-mers sit at several offsets
                if (fh == h && p0 + i < nk) {
                    const u8* tp = a.T + S.f_off[slot];
                    const u8* q = a.R + p0 + i;
                    if (kmer_equal_words(ld_unaligned64(tp), ld_unaligned64(tp + 8), ld_unaligned64(q), ld_unaligned64(q + 8), k)) atomicMin(first + GP_FIRST_J, (u32)S.f_off[slot]);
                }
            }
        }
    }
}

// final match list = concatenation of the pieces
__global__ void gp_piece_counts_k(const GpPiece* __restrict__ pieces, u32 n, u32* __restrict__ counts) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = pieces[i].count;
}
__global__ void __launch_bounds__(128) gp_concat_k(GpFrontArgs f, const u32* __restrict__ offs, u32 npieces, int* __restrict__ o_tpos, int* __restrict__ o_p, int* __restrict__ o_l) {
    for (u32 i = blockIdx.x; i < npieces; i += gridDim.x) {
        const GpPiece pc = f.pieces[i];
        const int *st, *sp, *sl;
        if (pc.src == 0xffffffffu) { st = f.f_tpos; sp = f.f_p; sl = f.f_l; }
        else { size_t base = (size_t)pc.src * f.s.cap_c; st = f.s.c_tpos + base; sp = f.s.c_p + base; sl = f.s.c_l + base; }
        const u32 o = offs[i];
        for (u32 x = threadIdx.x; x < pc.count; x += blockDim.x) {
            o_tpos[o + x] = st[pc.first + x]; o_p[o + x] = sp[pc.first + x]; o_l[o + x] = sl[pc.first + x];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// record writer for one long parse (compression.cpp:564-573 + delta_encode :258-292)
// ------------------------------------------------------------------------------------------------
// bytes[i] = literal gap before match i + its token
__global__ void g_match_bytes_k(const int* __restrict__ tpos, const int* __restrict__ mp, const int* __restrict__ ml, u32 M, u32* __restrict__ bytes, int absolute) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int prev_end = i ? tpos[i - 1] + ml[i - 1] : 0;
    int prev_p = (i && !absolute) ? mp[i - 1] : 0;
    bytes[i] = (u32)(tpos[i] - prev_end) + 3u + (u32)dec_len_i32(mp[i] - prev_p) + (u32)dec_len_u32((u32)ml[i]);
}
__global__ void g_write_tokens_k(const int* __restrict__ tpos, const int* __restrict__ mp, const int* __restrict__ ml, u32 M, const u32* __restrict__ offs,
                                 u8* __restrict__ out, const u32* __restrict__ d_body_base, int absolute) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int prev_end = i ? tpos[i - 1] + ml[i - 1] : 0;
    int prev_p = (i && !absolute) ? mp[i - 1] : 0;
    write_token(out + *d_body_base + offs[i] + (u32)(tpos[i] - prev_end), mp[i] - prev_p, ml[i]);
}
// literal symbols: every target position not covered by a match.  GL_SPAN positions per thread: one binary search, then
// the thread walks the GAPS that intersect its span (a span inside one long match costs the search and nothing else).
// Long spans suit match-dominated bodies (few searches); literal-dominated bodies want short ones (coalescing).
template <int GL_SPAN>
__global__ void __launch_bounds__(256) g_write_literals_k(const u8* __restrict__ T, i64 nt, const int* __restrict__ tpos, const int* __restrict__ ml, u32 M,
                                                         const u32* __restrict__ offs, const u32* __restrict__ d_tok_total, u8* __restrict__ out,
                                                         const u32* __restrict__ d_body_base) {
    const i64 x0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * GL_SPAN;
    if (x0 >= nt) return;
    const i64 x1 = x0 + GL_SPAN < nt ? x0 + GL_SPAN : nt;
    u8* body = out + *d_body_base;
    // i = number of matches starting at or before x0
    int lo = 0, hi = (int)M;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((i64)tpos[mid] <= x0) lo = mid + 1; else hi = mid; }
    int i = lo;
    i64 x = x0;
    while (x < x1) {
        // invariant: matches 0 .. i-1 start at or before x, match i (if any) starts after x
        const i64 gap_start = i ? (i64)tpos[i - 1] + ml[i - 1] : 0;       // the gap before match i begins where match i-1 ends
        const i64 next_start = i < (int)M ? (i64)tpos[i] : nt;
        if (x < gap_start) x = gap_start;                                  // inside match i-1
        const i64 lim = next_start < x1 ? next_start : x1;
        if (x < lim) {
            const u32 o = (i < (int)M ? offs[i] : *d_tok_total) + (u32)(x - gap_start);
            for (i64 y = x; y < lim; ++y) body[o + (u32)(y - x)] = T[y];
        }
        if (next_start >= x1) break;
        x = next_start; ++i;                                               // x sits on the first symbol of match i (now "i-1")
    }
}

// match_sequences(ref, tgt, k, m, true) on device-resident, already prepared sequences.  Leaves the matches in
// (m_tpos, m_p, m_l) and their number in *h_count.
struct GlobalMatches { int* tpos; int* p; int* l; u32 count; };

// Index modes.  FULL: every reference k-mer (what the reference builds, :41-47).  SAMPLED (references of a million k-mers and
// more): every GP_STRIDE-th position only -- an eighth of the sort.  The parse itself needs the whole-reference index in
// three places: the diagonal guesses (heuristic: a sampled index serves them), the first step
// (prev_match_end == -1: the occurrences of T[0..k) come from one brute-force pass, gp_first_k) and two rare events --
// no occurrence of the first k-mer at all, or the `pn2 == 0` fall-through of :134 -- which raise need_full: the parse is
// repeated with the full index, so the result never depends on the mode.
static const int GP_STRIDE = 8;
static int global_match_device(sccg_ctx* c, const u8* R, i64 nr, const u8* T, i64 nt, int k, int m, u32* sc, GlobalMatches* out, bool force_full = false) {
    if (k < 8 || k > 16) return set_error(SCCG_E_ARG, "global match_sequences supports 8 <= k <= 16");
    if (m < 0 || m > GP_MAX_M) return set_error(SCCG_E_ARG, "global match_sequences supports 0 <= m <= 120");
    // ---- reference k-mer index (:41-47)
    SCCG_CK(cudaEventRecord(c->ev_x[0], c->stream));
    const i64 nk = nr - k + 1 > 0 ? nr - k + 1 : 0;
    int stride = nk >= ((i64)1 << 20) ? GP_STRIDE : 1;
    if (const char* env = getenv("SCCG_GP_STRIDE")) { int v = atoi(env); if (v >= 1 && v <= 64) stride = v; }          // tests; 1 = always the full index
    if (force_full || nt < k || ((uintptr_t)R & 15u)) stride = 1;         // (the brute-force passes of the sampled mode read aligned 16-byte vectors)
    const i64 nidx = stride > 1 ? (nk + stride - 1) / stride : nk;
    // 24-bit hash keys + stable radix sort: positions ascending inside every key, the reference's per-bucket order.
    // (Measured alternative: a counting sort by hash bucket -- count, scan, fill with atomics -- needs 3.9 ms for the 59 M
    // k-mers of the chr19-shaped pair where the three radix passes need 1.4 ms: scattered 4-byte writes and 2 x 59 M global
    // atomics lose against coalesced tile-ordered stores.)
    u32 *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr, *bucket = nullptr;
    SCCG_TRY(buf(c, B_GKEYS, (size_t)nidx + 1, &keys));
    SCCG_TRY(buf(c, B_GVALS, (size_t)nidx + 1, &vals));
    SCCG_TRY(buf(c, B_GKEYS2, (size_t)nidx + 1, &keys2));
    SCCG_TRY(buf(c, B_GVALS2, (size_t)nidx + 1, &vals2));
    if (nidx > 0) {
        u32 *sk = nullptr, *sv = nullptr;
        if (stride > 1) {
            KmerSampledSource src{R, k, stride};
            SCCG_TRY(radix_sort_pairs(c, src, keys, vals, keys2, vals2, nidx, B_GHIST, GP_HASH_BITS / 8, &sk, &sv));
        } else {
            KmerPairSource src{R, k};
            SCCG_TRY(radix_sort_pairs(c, src, keys, vals, keys2, vals2, nidx, B_GHIST, GP_HASH_BITS / 8, &sk, &sv));
        }
        keys = sk; vals = sv;
    }
    int bucket_bits = gp_bucket_bits(nidx);
    if (const char* env = getenv("SCCG_GP_BUCKET_BITS")) { int v = atoi(env); if (v >= 4 && v <= GP_HASH_BITS) bucket_bits = v; }    // tests: small inputs through the 24-bit path
    SCCG_TRY(buf(c, B_GBUCKET, ((size_t)1 << bucket_bits) + 2, &bucket));
    LAUNCH(c, kmer_buckets_k, dim3(div_up(nidx + 1, 256 * KB_SPAN)), dim3(256), 0, (const u32*)keys, nidx, bucket, bucket_bits);
    // ---- chunk-speculative parse (:64-161)
    int chunk = GP_CHUNK_DEFAULT;
    if (const char* env = getenv("SCCG_GP_CHUNK")) { int v = atoi(env); if (v >= 64 && v <= (1 << 24)) chunk = v; }
    const u32 nchunks = nt > 0 ? div_up(nt, chunk) : 1;
    const u32 cap_c = (u32)(chunk / k) + 2;
    const size_t cap_all = (size_t)(nt / k) + 2;
    GpChunkInfo* info = nullptr; int* cbuf = nullptr; int* fbuf = nullptr; int* obuf = nullptr; GpPiece* pieces = nullptr; GpFrontState* st = nullptr; u32* pcounts = nullptr;
    const u32 cap_pieces = (u32)(cap_all < 0x7ffffff0u ? cap_all : 0x7ffffff0u) + 2 * nchunks + 8;
    SCCG_TRY(buf(c, B_GTMP1, (size_t)2 * nchunks + 1, &info));
    SCCG_TRY(buf(c, B_GLIT, (size_t)2 * nchunks * cap_c * 3 + 1, &cbuf));
    SCCG_TRY(buf(c, B_GTMP2, cap_all * 3 + 1, &fbuf));
    SCCG_TRY(buf(c, B_GREC, cap_all * 3 + 1, &obuf));
    SCCG_TRY(buf(c, B_GTMP3, (size_t)cap_pieces + 1, &pieces));
    SCCG_TRY(buf(c, B_GOFFS, 16, &st));                     // + the scan result 256 bytes behind the state
    SCCG_CK(cudaMemsetAsync(info, 0, sizeof(GpChunkInfo) * 2 * nchunks, c->stream));
    GpFrontState h_st; memset(&h_st, 0, sizeof h_st);
    h_st.j = 0; h_st.e = -1; h_st.status = GP_RUNNING;
    SCCG_CK(cudaMemcpyAsync(st, &h_st, sizeof h_st, cudaMemcpyHostToDevice, c->stream));
    GpFrontArgs f;
    f.s.a.R = R; f.s.a.nr = nr; f.s.a.T = T; f.s.a.nt = nt; f.s.a.keys = keys; f.s.a.vals = vals; f.s.a.nk = nidx; f.s.a.bucket = bucket; f.s.a.bucket_shift = GP_HASH_BITS - bucket_bits; f.s.a.k = k; f.s.a.m = m;
    f.s.a.m_tpos = nullptr; f.s.a.m_p = nullptr; f.s.a.m_l = nullptr; f.s.a.d_count = nullptr;
    f.s.info = info; f.s.c_tpos = cbuf; f.s.c_p = cbuf + (size_t)2 * nchunks * cap_c; f.s.c_l = cbuf + (size_t)4 * nchunks * cap_c;
    f.s.nchunks = nchunks; f.s.cap_c = cap_c; f.s.chunk = chunk; f.s.slot = 0; f.s.first_chunk = 0; f.s.lost_e = 0;
    f.st = st; f.pieces = pieces; f.cap_pieces = cap_pieces; f.f_tpos = fbuf; f.f_p = fbuf + cap_all; f.f_l = fbuf + 2 * cap_all;
    unsigned long long* d_hit = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(st) + 256);
    u32* ctl = reinterpret_cast<u32*>(reinterpret_cast<char*>(st) + 384);
    SCCG_CK(cudaMemsetAsync(ctl, 0, 8 * sizeof(u32), c->stream));
    u32* first = ctl + 4;
    SCCG_CK(cudaMemsetAsync(first + GP_FIRST_J, 0xff, sizeof(u32), c->stream));
    f.d_hit = d_hit; f.s.ctl = ctl;
    f.s.a.full = stride > 1 ? 0 : 1; f.s.a.first_list = nullptr; f.s.a.first = first; f.s.a.need_full = ctl + GP_CTL_NEED_FULL;
    if (stride > 1) {
        // the first step of the parse (state (0, -1): every occurrence is a candidate) by brute force over the reference; the
        // second and third kernel return at once unless T[0..k) has no occurrence (a target that begins with a mutation)
        u32* first_list = nullptr;
        SCCG_TRY(buf(c, B_GFIRST, (size_t)GP_FIRST_CAP + 1, &first_list));
        f.s.a.first_list = first_list;
        LAUNCH(c, gp_first_k, dim3(div_up(nk, 256 * 16)), dim3(256), 0, R, nk, T, k, first_list, GP_FIRST_CAP, first);
        LAUNCH(c, gp_first_scan_k, dim3((unsigned)c->sm_count * 8u), dim3(GP_T), 0, f.s.a, first);
        LAUNCH(c, gp_first_collect_k, dim3(div_up(nk, 256 * 16)), dim3(256), 0, R, nk, T, k, first_list, GP_FIRST_CAP, first);
    }
    SCCG_CK(cudaEventRecord(c->ev_x[1], c->stream));
    // Speculation in two batches: a small one on this stream, the rest on the side stream UNDERNEATH the front, which follows the
    // chunks as they are published (per-chunk ready flags).  A synchronised parse (the usual pair) splices through at the pace
    // of the speculation; a parse that gets lost (divergent pair) cancels what has not started yet: those chunks return at once.
    // (The second batch is enqueued before the front: tools that serialise kernels run it first and the front never waits.)
    u32 batch0 = (u32)c->sm_count * 2u;
    if (const char* env = getenv("SCCG_GP_BATCH0")) { int v = atoi(env); if (v >= 1) batch0 = (u32)v; }              // tests
    if (batch0 > nchunks) batch0 = nchunks;
    LAUNCH(c, gp_spec_k, dim3(batch0), dim3(GP_T), 0, f.s);                              // slot 0: chunk 0 exact, diagonal guesses for the others
    bool side_busy = false;
    if (nchunks > batch0) {
        SCCG_CK(cudaEventRecord(c->ev_side[0], c->stream));
        SCCG_CK(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
        GpSpecArgs rest = f.s; rest.first_chunk = batch0;
        SideLane side(c);
        LAUNCH(c, gp_spec_k, dim3(nchunks - batch0), dim3(GP_T), 0, rest);
        SCCG_CK(cudaEventRecord(c->ev_side[1], c->stream));
        side_busy = true;
    }
    // The host does not sit between the rounds: a group of [front, scan] pairs is queued at once and every kernel decides from the
    // state in device memory whether it has anything to do (front: not after the end; scan: only behind a front that got lost).
    int chain = 4;
    if (const char* env = getenv("SCCG_GP_CHAIN")) { int v = atoi(env); if (v >= 1 && v <= 64) chain = v; }              // tests
    unsigned scan_grid = (unsigned)c->sm_count * 16u;
    if (const char* env = getenv("SCCG_GP_SCAN_GRID")) { int v = atoi(env); if (v >= 1) scan_grid = (unsigned)v; }       // tests: tiny grids
    unsigned long long* d_trace = nullptr;
    const bool tracing = getenv("SCCG_GP_TRACE") != nullptr;
    if (tracing) { SCCG_TRY(buf(c, B_GHIST, (size_t)(2 * (1 + 4 * GP_TRACE_MAX)), (u32**)&d_trace)); SCCG_CK(cudaMemsetAsync(d_trace, 0, 8 * (1 + 4 * GP_TRACE_MAX), c->stream)); }
    f.trace = d_trace;
    GpScanArgs sa; sa.a = f.s.a; sa.st = st; sa.d_hit = d_hit; sa.trace = d_trace;
    for (int group = 0;; ++group) {
        for (int r = 0; r < (group == 0 ? 1 : chain); ++r) {         // the usual pair is done after the first front
            LAUNCH(c, gp_front_k, dim3(1), dim3(GP_T), 0, f);
            LAUNCH(c, gp_lost_scan_k, dim3(scan_grid), dim3(GP_T), 0, sa);
        }
        SCCG_CK(cudaMemcpyAsync(c->h_pinned, st, 416, cudaMemcpyDeviceToHost, c->stream));     // state, scan result, control words
        SCCG_CK(cudaStreamSynchronize(c->stream));
        memcpy(&h_st, c->h_pinned, sizeof h_st);
        c->prof.spec_rounds = (int32_t)h_st.rounds;
        if (reinterpret_cast<const u32*>(static_cast<const char*>(c->h_pinned) + 384)[GP_CTL_NEED_FULL] != 0u && stride > 1) {
            // the sampled mode met a lookup it cannot serve: same parse again with the index of every k-mer
            if (side_busy) SCCG_CK(cudaStreamWaitEvent(c->stream, c->ev_side[1], 0));
            return global_match_device(c, R, nr, T, nt, k, m, sc, out, true);
        }
        if (h_st.status == GP_DONE) break;
        if (h_st.status == GP_FULL || h_st.npieces + 2 >= cap_pieces) return set_error(SCCG_E_NOMEM, "internal: piece list overflow in the global parse");
        if (h_st.status != GP_LOST || group > 1000000) return set_error(SCCG_E_CUDA, "internal: global parse did not terminate");
    }
    if (tracing) {
        static unsigned long long h_tr[1 + 4 * GP_TRACE_MAX];
        cudaMemcpy(h_tr, d_trace, sizeof h_tr, cudaMemcpyDeviceToHost);
        const unsigned long long n = h_tr[0] < (unsigned long long)GP_TRACE_MAX ? h_tr[0] : GP_TRACE_MAX;
        for (unsigned long long i = 0; i < n; ++i)
            fprintf(stderr, "gp_trace %s begin %+9.1f us  dur %8.1f us  info %llx\n", h_tr[1 + 4 * i] == 1 ? "front" : "scan ", (double)(long long)(h_tr[2 + 4 * i] - h_tr[2]) / 1e3,
                    (double)(h_tr[3 + 4 * i] - h_tr[2 + 4 * i]) / 1e3, h_tr[4 + 4 * i]);
    }
    if (side_busy) SCCG_CK(cudaStreamWaitEvent(c->stream, c->ev_side[1], 0));               // (cancelled or not, the second batch must be off the buffers)
    // ---- concatenate the accepted pieces
    const u32 np = h_st.npieces;
    SCCG_TRY(buf(c, B_GTMP0, (size_t)np + 1, &pcounts));
    if (np) LAUNCH(c, gp_piece_counts_k, dim3(div_up(np, 256)), dim3(256), 0, (const GpPiece*)pieces, np, pcounts);
    SCCG_TRY(scan_exclusive_u32(c, pcounts, pcounts, (i64)np, sc + S_G1));
    if (np) {
        unsigned g = np < (unsigned)c->sm_count * 16u ? np : (unsigned)c->sm_count * 16u;
        LAUNCH(c, gp_concat_k, dim3(g), dim3(128), 0, f, (const u32*)pcounts, np, obuf, obuf + cap_all, obuf + 2 * cap_all);
    }
    SCCG_CK(cudaEventRecord(c->ev_x[2], c->stream));
    u32 h[S_COUNT];
    SCCG_TRY(read_scalars(c, sc, h, S_COUNT));
    cudaEventElapsedTime(&c->prof.index_ms, c->ev_x[0], c->ev_x[1]);
    cudaEventElapsedTime(&c->prof.parse_ms, c->ev_x[1], c->ev_x[2]);
    out->tpos = obuf; out->p = obuf + cap_all; out->l = obuf + 2 * cap_all; out->count = h[S_G1];
    c->prof.front_steps = (int32_t)h_st.steps;
    c->prof.index_stride = stride;
    return SCCG_OK;
}

struct GlobalPrep { u8* R2; u8* T2; u32* ncnt_s; u32* ncnt_e; u64* n_mask; };
// N runs of the upper-cased target, original coordinates (:527-554): count; toupper + erase every 'N' from both sequences
// (:523-524, :556-557).  Two lanes: the target's passes go to the side stream, the reference's stay on this one (either alone
// leaves half of the HBM bandwidth idle at chromosome size); no host round trip: sc[S_G2] / sc[S_G3] receive the stripped
// lengths, sc[S_N_K] / sc[S_N_KE] the run count.
static int global_prepare_enqueue(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, u32* sc, GlobalPrep* gp) {
    SCCG_TRY(buf(c, B_GREF, (size_t)nr + 64, &gp->R2));
    SCCG_TRY(buf(c, B_GTGT, (size_t)nt + 64, &gp->T2));
    SCCG_CK(cudaEventRecord(c->ev_side[0], c->stream));
    SCCG_CK(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
    {
        SideLane side(c);
        SCCG_TRY(strip_n_enqueue<1>(c, d_tgt, nt, gp->T2, B_TILE2, sc + S_G3));
        SCCG_TRY(rle_count<1>(c, d_tgt, nt, B_NRUN_CNT, B_NRUN_MASK, &gp->ncnt_s, &gp->ncnt_e, &gp->n_mask, sc + S_N_K, sc + S_N_KE));
        SCCG_CK(cudaEventRecord(c->ev_side[1], c->stream));
    }
    SCCG_TRY(strip_n_enqueue<1>(c, d_ref, nr, gp->R2, B_TILE3, sc + S_G2));
    SCCG_CK(cudaStreamWaitEvent(c->stream, c->ev_side[1], 0));
    return SCCG_OK;
}

static int compress_global_device(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, const char* header, i64 nh,
                                  u32 low_k, const u8* d_low_text, int text_delta, CompressResult* res) {
    u32* sc = nullptr;
    SCCG_TRY(buf(c, B_SCALARS, (size_t)S_COUNT, &sc));
    const bool step_trace = getenv("SCCG_STEP_TRACE") != nullptr;             // development aid: where the call's time goes
    if (step_trace) SCCG_CK(cudaEventRecord(c->ev_side[2], c->stream));
    GlobalPrep gp; memset(&gp, 0, sizeof gp);
    u32 h[S_COUNT];
    SCCG_TRY(global_prepare_enqueue(c, d_ref, nr, d_tgt, nt, sc, &gp));
    SCCG_TRY(read_scalars(c, sc, h, S_COUNT));                                // one host round trip for all the counts
    u8 *R2 = gp.R2, *T2 = gp.T2;
    u32 *ncnt_s = gp.ncnt_s, *ncnt_e = gp.ncnt_e;
    u64* n_mask = gp.n_mask;
    const i64 nr2 = (i64)h[S_G2], nt2 = (i64)h[S_G3];
    SCCG_CK(cudaMemsetAsync(R2 + nr2, 0, 64, c->stream));          // unaligned word loads run up to 15 B past the end
    SCCG_CK(cudaMemsetAsync(T2 + nt2, 0, 64, c->stream));
    if (h[S_N_K] != h[S_N_KE]) return set_error(SCCG_E_CUDA, "internal: N-run start/end counts differ");
    const u32 n_k = h[S_N_K];

    // ---- the global parse (:561)
    SCCG_CK(cudaEventRecord(c->ev[1], c->stream));
    GlobalMatches gm;
    SCCG_TRY(global_match_device(c, R2, nr2, T2, nt2, K1, GLOBAL_M, sc, &gm));
    SCCG_CK(cudaEventRecord(c->ev[2], c->stream));
    const u32 M = gm.count;

    // ---- sizes
    u32* mbytes = nullptr;
    SCCG_TRY(buf(c, B_GTMP0, (size_t)M + 1, &mbytes));
    if (M) LAUNCH(c, g_match_bytes_k, dim3(div_up(M, 256)), dim3(256), 0, (const int*)gm.tpos, (const int*)gm.p, (const int*)gm.l, M, mbytes, text_delta);
    SCCG_TRY(scan_exclusive_u32(c, mbytes, mbytes, (i64)M, sc + S_G4));
    u32 last[3] = {0, 0, 0};                                      // end of the last match in the target
    if (M) {
        SCCG_CK(cudaMemcpyAsync(&last[0], gm.tpos + (M - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaMemcpyAsync(&last[1], gm.l + (M - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    // the N-run text (staged: its place in the image depends on the length of the lowercase text)
    int *nrun_s = nullptr, *nrun_e = nullptr;
    u8* ntext = nullptr;
    SCCG_TRY(buf(c, B_NRUN_TEXT, 24ull * n_k + 16, &ntext));
    SCCG_TRY(rle_emit<1>(c, n_mask, nt, n_k, ncnt_s, ncnt_e, sc + S_N_K, B_NRUN_START, B_NRUN_END, B_NRUN_BYTES, &nrun_s, &nrun_e,
                         ntext, sc + S_N_TEXT));
    SCCG_TRY(read_scalars(c, sc, h, S_COUNT));                    // body size and the lengths of both run-list texts, one round trip
    const i64 trailing = nt2 - ((i64)last[0] + (i64)last[1]);
    const u64 body_bytes = (u64)h[S_G4] + (u64)trailing;
    const u32 low_text = h[S_LOW_TEXT], n_text = h[S_N_TEXT];

    // ---- assemble "<header>\n<lowercase runs>\n<N runs>\n<body>"
    const size_t hdr_bytes = nh > 0 ? (size_t)nh + 1 : 0;
    const size_t cap = hdr_bytes + (size_t)low_text + 1 + (size_t)n_text + 1 + body_bytes;
    if (cap >= 0xffffffffull) return set_error(SCCG_E_ARG, "encoded output would exceed 4 GiB");
    u8* out = nullptr;
    SCCG_TRY(buf(c, B_OUT, cap + 16, &out));
    SCCG_TRY(write_header(c, out, header, nh));
    // the lowercase-run text was produced on the side lane of compress_device (sc[S_LOW_TEXT] holds its length)
    if (low_k) LAUNCH(c, copy_text_k, dim3(low_k < 4096 ? 8 : (unsigned)c->sm_count * 16u), dim3(256), 0, out + hdr_bytes, d_low_text, (const u32*)(sc + S_LOW_TEXT));
    LAUNCH(c, put_separators_k, dim3(1), dim3(1), 0, out, (u32)hdr_bytes, sc, 1, 0u, 0u);
    if (n_text) SCCG_CK(cudaMemcpyAsync(out + hdr_bytes + low_text + 1, ntext, n_text, cudaMemcpyDeviceToDevice, c->stream));
    if (M) LAUNCH(c, g_write_tokens_k, dim3(div_up(M, 256)), dim3(256), 0, (const int*)gm.tpos, (const int*)gm.p, (const int*)gm.l, M,
                  (const u32*)mbytes, out, (const u32*)(sc + S_BODY_BASE), text_delta);
    if (nt2 > 0) {
        if ((i64)body_bytes * 4 > nt2)                            // literal-heavy body
            LAUNCH(c, g_write_literals_k<16>, dim3(div_up(nt2, 256 * 16)), dim3(256), 0, (const u8*)T2, nt2, (const int*)gm.tpos, (const int*)gm.l, M,
                   (const u32*)mbytes, (const u32*)(sc + S_G4), out, (const u32*)(sc + S_BODY_BASE));
        else
            LAUNCH(c, g_write_literals_k<64>, dim3(div_up(nt2, 256 * 64)), dim3(256), 0, (const u8*)T2, nt2, (const int*)gm.tpos, (const int*)gm.l, M,
                   (const u32*)mbytes, (const u32*)(sc + S_G4), out, (const u32*)(sc + S_BODY_BASE));
    }
    SCCG_CK(cudaEventRecord(c->ev[3], c->stream));
    SCCG_CK(cudaStreamSynchronize(c->stream));
    res->d_out = out;
    res->out_len = (i64)hdr_bytes + low_text + 1 + n_text + 1 + (i64)body_bytes;
    res->mode = 1;
    res->stoi_failed = 0;
    if (text_delta) {
        SCCG_TRY(finish_text_delta(c, sc, (u32)(hdr_bytes + low_text + 1 + n_text + 1), res));
        SCCG_CK(cudaEventRecord(c->ev[3], c->stream));
        SCCG_CK(cudaStreamSynchronize(c->stream));
    }
    float ms = 0.f;
    if (step_trace) {
        float a = 0, b = 0, d = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev_side[2]); cudaEventElapsedTime(&b, c->ev_side[2], c->ev[1]); cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
        fprintf(stderr, "step_trace (global): local attempt %.1f us, N runs + strip %.1f, index %.1f, parse %.1f, sizes + writers %.1f\n", a * 1e3, b * 1e3,
                c->prof.index_ms * 1e3, c->prof.parse_ms * 1e3, d * 1e3);
    }
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); c->prof.kernels_ms = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->prof.match_ms = ms;
    c->prof.serialize_ms = c->prof.kernels_ms - c->prof.match_ms;
    c->prof.mode = 1;
    return SCCG_OK;
}

// function-level match_sequences(Sr, St, k, m, global = true, offset): sequences are used as given (:561 passes
// upper-cased, N-free strings); records go back to the host as struct-of-arrays
static int match_sequences_global(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, const char* h_tgt, int k, int m, int offset, sccg_records* out) {
    u32* sc = nullptr;
    SCCG_TRY(buf(c, B_SCALARS, (size_t)S_COUNT, &sc));
    SCCG_CK(cudaMemsetAsync(sc, 0, sizeof(u32) * S_COUNT, c->stream));
    SCCG_CK(cudaMemsetAsync((void*)(d_ref + nr), 0, 64, c->stream));
    SCCG_CK(cudaMemsetAsync((void*)(d_tgt + nt), 0, 64, c->stream));
    GlobalMatches gm;
    SCCG_TRY(global_match_device(c, d_ref, nr, d_tgt, nt, k, m, sc, &gm));
    const u32 M = gm.count;
    int* hm = (int*)malloc(sizeof(int) * 3 * (size_t)(M + 1));
    if (!hm) return set_error(SCCG_E_NOMEM, "out of host memory");
    if (M) {
        cudaMemcpyAsync(hm, gm.tpos, sizeof(int) * M, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(hm + M, gm.p, sizeof(int) * M, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(hm + 2 * (size_t)M, gm.l, sizeof(int) * M, cudaMemcpyDeviceToHost, c->stream);
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { free(hm); return set_error(SCCG_E_CUDA, "download of the matches failed"); }
    int64_t cap = 2 * (int64_t)M + 2;
    out->p = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    out->l = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    out->lit_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap + 1));
    out->lits = (char*)malloc((size_t)nt + 1);
    if (!out->p || !out->l || !out->lit_off || !out->lits) { free(hm); sccg_records_free(out); return set_error(SCCG_E_NOMEM, "out of host memory"); }
    int64_t n = 0, lit = 0, pos = 0;
    for (u32 i = 0; i <= M; ++i) {
        int64_t tpos = i < M ? hm[i] : nt;
        if (tpos > pos) {
            out->p[n] = -1; out->l[n] = 0; out->lit_off[n] = lit;
            memcpy(out->lits + lit, h_tgt + pos, (size_t)(tpos - pos));
            lit += tpos - pos; ++n;
        }
        if (i < M) {
            out->p[n] = hm[M + i] + offset; out->l[n] = hm[2 * (size_t)M + i]; out->lit_off[n] = lit; ++n;
            pos = tpos + hm[2 * (size_t)M + i];
        }
    }
    out->lit_off[n] = lit;
    out->n = n;
    free(hm);
    return SCCG_OK;
}

}  // namespace sccg
