// FASTA ingest on the device: read_genomes_from_files (compression.cpp:181-220) and the reference reader
// of decompress_genome (decompression.cpp:47-58) on the raw file image, so that the host never touches the symbols.
//
//   reference file: every line that starts with '>' is skipped (:194), the rest is concatenated and every isspace() byte
//                   is removed (:200)                       [empty lines hold nothing, skipping them is a no-op]
//   target file   : only the FIRST line that starts with '>' is the header (kept verbatim, :211-215); later '>' lines stay
//                   in the sequence; isspace() bytes removed (:219)
//
//   1. fasta_hdr_count_k / fasta_hdr_fill_k : header lines = '>' at a line start, up to the next '\n'; they are rare, so
//      they become a sorted list of byte ranges (count, scan, fill);
//   2. fasta_count_k / fasta_write_k        : stream compaction of the file image with the keep mask
//      !isspace(byte) && !(byte inside a header range)      (16 bytes per thread, byte-SWAR masks, tile counts + scan).
// HBM-bound: the file image is read three times (3 B per file byte) and the symbols written once.
#pragma once
#include "sccg_scan.cuh"

namespace sccg {

static const int FA_T = 256;
static const int FA_TILE = FA_T * 16;

// bit 7 of every byte that is C-locale isspace(): ' ' or 0x09..0x0D
__device__ __forceinline__ u64 space_flags8(u64 w) {
    u64 x = w & SCCG_B7F;
    u64 ge9 = x + (u64)(0x80 - 9) * SCCG_B01;            // bit7 <=> (b & 0x7f) >= 9
    u64 ge14 = x + (u64)(0x80 - 14) * SCCG_B01;          // bit7 <=> (b & 0x7f) >= 14
    return ((ge9 & ~ge14 & ~w) | eq_flags8(w, ' ')) & SCCG_B80;
}
__device__ __forceinline__ u32 mask16(ulonglong2 v, u8 ch) { return movemask8(eq_flags8(v.x, ch)) | (movemask8(eq_flags8(v.y, ch)) << 8); }

// bit b <=> F[i0 + b] is a '>' at the start of a line
__device__ __forceinline__ u32 fasta_hdr_starts(const u8* __restrict__ F, i64 n, i64 i0, ulonglong2* v) {
    if (i0 >= n) return 0u;
    *v = *reinterpret_cast<const ulonglong2*>(F + i0);
    const i64 left = n - i0;
    const u32 V = left >= 16 ? 0xffffu : ((1u << (int)left) - 1u);
    const u32 gt = mask16(*v, '>') & V;
    if (!gt) return 0u;
    const u32 nl = mask16(*v, '\n');
    const u32 prev_nl = (i0 == 0 || F[i0 - 1] == '\n') ? 1u : 0u;
    return gt & ((nl << 1) | prev_nl);
}

__global__ void __launch_bounds__(FA_T) fasta_hdr_count_k(const u8* __restrict__ F, i64 n, u32* __restrict__ th_cnt) {
    const i64 t = (i64)blockIdx.x * FA_T + threadIdx.x;
    const i64 i0 = t * 16;
    if (i0 >= n) return;
    ulonglong2 v;
    th_cnt[t] = (u32)__popc(fasta_hdr_starts(F, n, i0, &v));
}

// ranges[2k], ranges[2k+1] = [start, end) of header line k (end = position of its '\n', or n)
__global__ void __launch_bounds__(FA_T) fasta_hdr_fill_k(const u8* __restrict__ F, i64 n, const u32* __restrict__ th_off, i64* __restrict__ ranges) {
    const i64 t = (i64)blockIdx.x * FA_T + threadIdx.x;
    const i64 i0 = t * 16;
    if (i0 >= n) return;
    ulonglong2 v;
    u32 k = th_off[t];
    for (u32 m = fasta_hdr_starts(F, n, i0, &v); m; m &= m - 1) {
        const i64 st = i0 + __ffs((int)m) - 1;
        i64 en = st + 1;
        while (en < n && F[en] != '\n') ++en;                 // header lines are short
        ranges[2 * k] = st; ranges[2 * k + 1] = en;
        ++k;
    }
}

// bit b <=> F[i0 + b] is kept
__device__ __forceinline__ u32 fasta_keep_mask(const u8* __restrict__ F, i64 n, i64 i0, const i64* __restrict__ ranges, int nr, ulonglong2* v) {
    if (i0 >= n) return 0u;
    *v = *reinterpret_cast<const ulonglong2*>(F + i0);
    const i64 left = n - i0;
    u32 keep = left >= 16 ? 0xffffu : ((1u << (int)left) - 1u);
    keep &= ~(movemask8(space_flags8(v->x)) | (movemask8(space_flags8(v->y)) << 8));
    if (nr > 0 && keep) {
        // header ranges that overlap [i0, i0 + 16): ranges are sorted and disjoint
        int lo = 0, hi = nr;                                   // first range with end > i0
        while (lo < hi) { int mid = (lo + hi) >> 1; if (ranges[2 * mid + 1] <= i0) lo = mid + 1; else hi = mid; }
        for (int r = lo; r < nr && ranges[2 * r] < i0 + 16; ++r) {
            const i64 a = ranges[2 * r] > i0 ? ranges[2 * r] - i0 : 0;
            const i64 b = ranges[2 * r + 1] < i0 + 16 ? ranges[2 * r + 1] - i0 : 16;
            if (b > a) keep &= ~(((1u << (int)b) - 1u) & ~((1u << (int)a) - 1u));
        }
    }
    return keep;
}

__global__ void __launch_bounds__(FA_T) fasta_count_k(const u8* __restrict__ F, i64 n, const i64* __restrict__ ranges, int nr, u32* __restrict__ cnt) {
    __shared__ u32 sm[40];
    const i64 i0 = (i64)blockIdx.x * FA_TILE + (i64)threadIdx.x * 16;
    ulonglong2 v;
    const u32 m = fasta_keep_mask(F, n, i0, ranges, nr, &v);
    u32 tot;
    block_scan_excl((u32)__popc(m), sm, &tot);
    if (threadIdx.x == 0) cnt[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(FA_T) fasta_write_k(const u8* __restrict__ F, i64 n, const i64* __restrict__ ranges, int nr, const u32* __restrict__ tile_off,
                                                     u8* __restrict__ dst) {
    __shared__ u32 sm[40];
    __align__(16) __shared__ u8 stage[FA_TILE + 32];
    const i64 i0 = (i64)blockIdx.x * FA_TILE + (i64)threadIdx.x * 16;
    ulonglong2 v; v.x = 0; v.y = 0;
    u32 m = fasta_keep_mask(F, n, i0, ranges, nr, &v);
    u32 tot;
    const u32 excl = block_scan_excl((u32)__popc(m), sm, &tot);
    block_compact_store(stage, dst + tile_off[blockIdx.x], excl, m, v.x, v.y, tot);
}

struct FastaSeq { u8* d_seq; i64 len; i64 hdr_start, hdr_end; };   // hdr_*: the target header line inside the file image (-1: none)

// d_file: the raw file image in device memory (16-byte aligned, >= 16 readable bytes past the end).
// slot_seq receives the symbols; slot_tmp / slot_rng are scratch.
static int fasta_ingest(sccg_ctx* c, const u8* d_file, i64 n, bool is_target, int slot_seq, int slot_tmp, int slot_rng, u32* d_scalar, FastaSeq* out) {
    out->d_seq = nullptr; out->len = 0; out->hdr_start = out->hdr_end = -1;
    SCCG_TRY(buf(c, slot_seq, (size_t)(n > 0 ? n : 1) + 128, &out->d_seq));
    if (n <= 0) return SCCG_OK;
    const i64 nth = (n + 15) / 16;
    const unsigned blocks = div_up(nth, FA_T);
    // ---- 1. header lines
    u32* th_cnt = nullptr;
    SCCG_TRY(buf(c, slot_tmp, (size_t)nth + 4, &th_cnt));
    LAUNCH(c, fasta_hdr_count_k, dim3(blocks), dim3(FA_T), 0, d_file, n, th_cnt);
    SCCG_TRY(scan_exclusive_u32(c, th_cnt, th_cnt, nth, d_scalar));
    SCCG_CK(cudaMemcpyAsync(c->h_pinned, d_scalar, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    SCCG_CK(cudaStreamSynchronize(c->stream));
    const u32 nhdr = *(u32*)c->h_pinned;
    i64* ranges = nullptr;
    SCCG_TRY(buf(c, slot_rng, (size_t)nhdr * 2 + 2, &ranges));
    int nr = 0;
    if (nhdr) {
        LAUNCH(c, fasta_hdr_fill_k, dim3(blocks), dim3(FA_T), 0, d_file, n, (const u32*)th_cnt, ranges);
        nr = is_target ? 1 : (int)nhdr;                          // target: only the first '>' line is a header (:211)
        if (is_target) {
            SCCG_CK(cudaMemcpyAsync(c->h_pinned, ranges, sizeof(i64) * 2, cudaMemcpyDeviceToHost, c->stream));
            SCCG_CK(cudaStreamSynchronize(c->stream));
            out->hdr_start = ((i64*)c->h_pinned)[0]; out->hdr_end = ((i64*)c->h_pinned)[1];
        }
    }
    // ---- 2. compaction
    u32* cnt = th_cnt;                                           // th_cnt is dead once the ranges exist; tiles <= threads
    LAUNCH(c, fasta_count_k, dim3(blocks), dim3(FA_T), 0, d_file, n, (const i64*)ranges, nr, cnt);
    SCCG_TRY(scan_exclusive_u32(c, cnt, cnt, (i64)blocks, d_scalar));
    LAUNCH(c, fasta_write_k, dim3(blocks), dim3(FA_T), 0, d_file, n, (const i64*)ranges, nr, (const u32*)cnt, out->d_seq);
    SCCG_CK(cudaMemcpyAsync(c->h_pinned, d_scalar, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    SCCG_CK(cudaStreamSynchronize(c->stream));
    out->len = (i64)*(u32*)c->h_pinned;
    SCCG_CK(cudaMemsetAsync(out->d_seq + out->len, 0, 64, c->stream));
    return SCCG_OK;
}

}  // namespace sccg
