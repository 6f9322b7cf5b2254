// One chromosome over several GPUs: every GPU gets a contiguous range of segment pairs -- the matching
// slices of reference and target, nothing else -- and the local segment-matching path (compression.cpp:381-481) runs on
// it in two phases:
//   sccg_shard_match : lowercase-run masks, segment matcher, driver bookkeeping of the slice; returns the few values that
//                      cross shard borders (T2 statuses of the border segments, last match, first / last lowercase run)
//   sccg_shard_write : given the carries computed from all shards' values, writes this shard's part of the lowercase-run
//                      line and of the body, byte-identical to the corresponding part of the unsharded file.
// The exchange between the phases is a few dozen bytes per GPU (one all_gather, sharding.py); the parts are concatenated on
// rank 0.  No data-path collective.  Aborts to global mode and targets that contain '(' are reported to the caller, which
// takes the unsharded path for that pair.
#pragma once
#include "sccg_compress.cuh"

namespace sccg {

// index of the last segment with at least one match (-1: none)
__global__ void shard_last_match_k(const u32* __restrict__ seginfo, int n_iter, int* __restrict__ last) {
    int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    int v = (i < n_iter && SEGINFO_NMATCH(seginfo[i])) ? i : -1;
    v = __reduce_max_sync(SCCG_FULL_MASK, v);
    if (lane_of() == 0 && v >= 0) atomicMax(last, v);
}

// what a shard reports across its borders, as it travels in the all-gather (same layout on the device and on the host)
struct ShardBorder {
    i64 n_segments, n_runs, first_run_start, first_run_len, last_run_start, last_run_len;
    int abort_inside, has_paren, has_match, last_p;
    int head_status[4], tail_status[4];
    i64 low_text_len, body_len;           // filled in by the write phase (second all-gather of the multi-GPU layer)
    i64 pad[2];
};
static_assert(sizeof(ShardBorder) == 128, "one 128-byte record per rank");

// bit 0 = the segment increments the T2 counter, bit 1 = it can end an abort window (:417-424, :454-462)
__device__ __forceinline__ int shard_seg_status(u32 x) {
    if (x == 0xffffffffu) return 0;                                // not processed (early abort): the shard reports abort_inside anyway
    const int inc = !SEGINFO_ALLN(x) && (SEGINFO_NMATCH(x) == 0 || SEGINFO_BAD(x));
    const int end = !SEGINFO_ALLN(x) && SEGINFO_NMATCH(x) == 0;
    return inc | (end << 1);
}

// one thread: collects everything that crosses the shard's borders into one record (a single copy / all-gather instead of
// a dozen small device-to-host reads)
__global__ void shard_border_k(const u32* __restrict__ seginfo, const u32* __restrict__ matches, int n_iter, const int* __restrict__ d_last,
                               const int* __restrict__ run_s, const int* __restrict__ run_e, const u32* __restrict__ sc, i64 seg_base, ShardBorder* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    ShardBorder b;
    memset(&b, 0, sizeof b);
    b.n_segments = n_iter;
    b.abort_inside = sc[S_ABORT] ? 1 : 0;
    b.has_paren = sc[S_PAREN] ? 1 : 0;
    const u32 low_k = sc[S_LOW_K];
    b.n_runs = low_k;
    const i64 tgt_off = seg_base * SEG;
    if (low_k) {
        b.first_run_start = tgt_off + run_s[0]; b.first_run_len = run_e[0] - run_s[0];
        b.last_run_start = tgt_off + run_s[low_k - 1]; b.last_run_len = run_e[low_k - 1] - run_s[low_k - 1];
    }
    const int nb = n_iter < 4 ? n_iter : 4;
    for (int i = 0; i < nb; ++i) { b.head_status[i] = shard_seg_status(seginfo[i]); b.tail_status[i] = shard_seg_status(seginfo[n_iter - nb + i]); }
    const int last_seg = n_iter > 0 ? *d_last : -1;
    if (last_seg >= 0 && !b.abort_inside) {
        const u32 info_w = seginfo[last_seg];
        const u32 pk = matches[(i64)last_seg * LM_SLOT + (SEGINFO_NMATCH(info_w) - 1)];
        b.has_match = 1;
        b.last_p = (int)((last_seg + seg_base) * SEG + (i64)((pk >> 10) & 0x3ffu));
    }
    *out = b;
}

// arr != NULL: the slices are still arriving on the copy stream (enqueue_pair_upload); the matcher runs chunk by chunk.
// Leaves the border record in device memory (*d_border_out, owned by the context).
static int shard_match(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, i64 seg_base, int is_last, const ChunkArrival* arr, ShardBorder** d_border_out) {
    ShardState& st = c->shard;
    memset(&st, 0, sizeof st);
    u32* sc = nullptr;
    SCCG_TRY(buf(c, B_SCALARS, (size_t)S_COUNT, &sc));
    SCCG_CK(cudaMemsetAsync(sc, 0, sizeof(u32) * S_COUNT, c->stream));
    SCCG_CK(cudaEventRecord(c->ev[0], c->stream));
    // lowercase runs of the slice (:341-367) on the side lane, underneath the matcher
    u32 *cnt_s = nullptr, *cnt_e = nullptr; u64* low_mask = nullptr;
    SCCG_CK(cudaEventRecord(c->ev_side[0], c->stream));
    SCCG_CK(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
    if (arr) {
        SCCG_CK(cudaStreamWaitEvent(c->side_stream, arr->ev_tgt, 0));
        SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_tgt, 0));
    }
    {
        SideLane side(c);
        SCCG_TRY(rle_count<0>(c, d_tgt, nt, B_RUN_CNT, B_RUN_MASK, &cnt_s, &cnt_e, &low_mask, sc + S_LOW_K, sc + S_LOW_KE, sc + S_PAREN));
    }
    // segment matcher + driver bookkeeping (:381-474)
    const i64 n_rseg = (nr + SEG - 1) / SEG, n_tseg = (nt + SEG - 1) / SEG;
    const int n_iter = (int)(n_rseg < n_tseg ? n_rseg : n_tseg);
    u32 *seginfo = nullptr, *matches = nullptr, *seg_bytes = nullptr; int* seg_prev = nullptr;
    SCCG_TRY(buf(c, B_SEGINFO, (size_t)n_iter + 1, &seginfo));
    SCCG_TRY(buf(c, B_MATCH, (size_t)n_iter * LM_SLOT + 1, &matches));
    SCCG_TRY(buf(c, B_SEGBYTES, (size_t)n_iter + 1, &seg_bytes));
    SCCG_TRY(buf(c, B_SEGPREV, (size_t)n_iter + 1, &seg_prev));
    int* d_last = (int*)(sc + S_G6);
    SCCG_CK(cudaEventRecord(c->ev[1], c->stream));
    if (n_iter > 0) {
        const size_t smem = sizeof(LmWarpSmem) * LM_WARPS;
        SCCG_SET_MAX_SMEM(seg_match_k<SCCG_LM_CLAIM>, smem);
        const unsigned cap = (unsigned)c->sm_count * (unsigned)LM_CTAS_PER_SM;
        SCCG_CK(cudaMemsetAsync(seginfo, 0xff, sizeof(u32) * (size_t)n_iter, c->stream));
        SCCG_CK(cudaMemsetAsync(d_last, 0xff, sizeof(int), c->stream));
        const int n_launch = arr ? arr->n : 1;
        int seg_lo = 0;
        for (int i = 0; i < n_launch && seg_lo < n_iter; ++i) {                // one launch per arrived reference chunk
            int seg_hi = n_iter;
            if (arr && i + 1 < n_launch) {
                const i64 resident = (i64)(i + 1) * arr->chunk;
                seg_hi = (int)(resident / SEG < (i64)n_iter ? resident / SEG : (i64)n_iter);
            }
            if (arr) SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_ref[i + 1 < n_launch ? i : arr->n - 1], 0));
            if (seg_hi <= seg_lo) continue;
            const unsigned w = div_up(seg_hi - seg_lo, LM_WARPS);
            LAUNCH(c, seg_match_k<SCCG_LM_CLAIM>, dim3(w < cap ? w : cap), dim3(LM_WARPS * 32), smem, d_ref, nr, d_tgt, nt, seg_lo, seg_hi, n_iter, K1, K2, seginfo, matches,
                   sc + S_WORK + (i & 31), sc + S_ABORT, c->use_diag);
            seg_lo = seg_hi;
        }
        SCCG_CK(cudaEventRecord(c->ev[2], c->stream));
        LAUNCH(c, seg_bytes_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, (const u32*)matches, n_iter, seg_bytes, seg_prev, sc + S_ABORT, 0, 0, 0);
        LAUNCH(c, shard_last_match_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, n_iter, d_last);
    } else {
        SCCG_CK(cudaEventRecord(c->ev[2], c->stream));
    }
    // side lane: run count -> the runs themselves (needed for the border values now, for the text in shard_write)
    u32 h[S_COUNT];
    int *run_s = nullptr, *run_e = nullptr;
    u32 low_k = 0;
    {
        SideLane side(c);
        SCCG_TRY(read_scalars(c, sc, h, S_COUNT));                            // synchronises the side stream only
        if (h[S_LOW_K] != h[S_LOW_KE]) return set_error(SCCG_E_CUDA, "internal: run start/end counts differ");
        low_k = h[S_LOW_K];
        SCCG_TRY(buf(c, B_RUN_START, (size_t)low_k + 1, &run_s));
        SCCG_TRY(buf(c, B_RUN_END, (size_t)low_k + 1, &run_e));
        if (low_k) LAUNCH(c, rle_write_k, dim3(div_up(nt > 0 ? nt : 1, RLE_TILE)), dim3(RLE_T), 0, (const u64*)low_mask, nt, (const u32*)cnt_s, (const u32*)cnt_e, run_s, run_e);
        SCCG_CK(cudaEventRecord(c->ev_side[1], c->stream));
    }
    SCCG_CK(cudaStreamWaitEvent(c->stream, c->ev_side[1], 0));
    ShardBorder* d_border = nullptr;
    SCCG_TRY(buf(c, B_SHARD, 1, &d_border));
    LAUNCH(c, shard_border_k, dim3(1), dim3(32), 0, (const u32*)seginfo, (const u32*)matches, n_iter, (const int*)d_last, (const int*)run_s, (const int*)run_e, (const u32*)sc,
           seg_base, d_border);
    *d_border_out = d_border;
    st.valid = 1; st.d_tgt = d_tgt; st.nt = nt; st.n_iter = n_iter; st.seg_base = seg_base; st.is_last = is_last; st.low_k = low_k;
    st.leftover = n_tseg > n_iter ? nt - (i64)n_iter * SEG : 0;
    return SCCG_OK;
}

static void shard_info_from_border(const ShardBorder& b, sccg_shard_info* info) {
    memset(info, 0, sizeof *info);
    info->n_segments = b.n_segments; info->abort_inside = b.abort_inside; info->has_paren = b.has_paren;
    info->has_match = b.has_match; info->last_p = b.last_p; info->n_runs = b.n_runs;
    info->first_run_start = b.first_run_start; info->first_run_len = b.first_run_len;
    info->last_run_start = b.last_run_start; info->last_run_len = b.last_run_len;
    for (int i = 0; i < 4; ++i) { info->head_status[i] = b.head_status[i]; info->tail_status[i] = b.tail_status[i]; }
}

// Write phase, first half: everything whose result is a SIZE (text length of the run-list part in sc[S_LOW_TEXT], of the body
// part without leftover segments in sc[S_BODY_MAIN]).  The run-list text itself is already written (B_RUN_TEXT).
static int shard_write_sizes(sccg_ctx* c, const sccg_shard_carry* carry, u8** d_low) {
    ShardState& st = c->shard;
    if (!st.valid) return set_error(SCCG_E_ARG, "sccg_shard_write without a preceding sccg_shard_match on this context");
    u32* sc = (u32*)c->bufs[B_SCALARS].p;
    u32* seginfo = (u32*)c->bufs[B_SEGINFO].p; u32* matches = (u32*)c->bufs[B_MATCH].p;
    u32* seg_bytes = (u32*)c->bufs[B_SEGBYTES].p; int* seg_prev = (int*)c->bufs[B_SEGPREV].p;
    int* run_s = (int*)c->bufs[B_RUN_START].p; int* run_e = (int*)c->bufs[B_RUN_END].p;
    const int n_iter = st.n_iter;
    // ---- body part: delta chain continued from the shards before this one
    if (n_iter > 0)
        LAUNCH(c, seg_bytes_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, (const u32*)matches, n_iter, seg_bytes, seg_prev, sc + S_ABORT, 0,
               (int)st.seg_base, (int)carry->prev_p);
    SCCG_TRY(scan_exclusive_u32(c, seg_bytes, seg_bytes, (i64)n_iter, sc + S_BODY_MAIN));
    // ---- lowercase-run part
    RunCarry rc;
    rc.pos_off = st.seg_base * SEG; rc.prev_start = carry->prev_run_start; rc.extra_last = carry->extra_last_len;
    rc.skip_first = carry->skip_first_run ? 1 : 0; rc.reaches_end = carry->last_run_reaches_end ? 1 : 0;
    u8* low_text = nullptr;
    SCCG_TRY(buf(c, B_RUN_TEXT, 24ull * st.low_k + 16, &low_text));
    int *rs = run_s, *re = run_e;
    SCCG_TRY(rle_emit<0>(c, (const u64*)nullptr, st.nt, st.low_k, (const u32*)nullptr, (const u32*)nullptr, sc + S_LOW_K, B_RUN_START, B_RUN_END, B_RUN_BYTES,
                         &rs, &re, low_text, sc + S_LOW_TEXT, false, rc));
    *d_low = low_text;
    return SCCG_OK;
}

// Write phase, second half (body_main = sc[S_BODY_MAIN] as read by the caller): the body part
static int shard_write_body(sccg_ctx* c, u32 body_main, u8** d_body, i64* body_len) {
    ShardState& st = c->shard;
    if (!st.valid) return set_error(SCCG_E_ARG, "sccg_shard_write without a preceding sccg_shard_match on this context");
    u32* sc = (u32*)c->bufs[B_SCALARS].p;
    u32* seginfo = (u32*)c->bufs[B_SEGINFO].p; u32* matches = (u32*)c->bufs[B_MATCH].p;
    u32* seg_bytes = (u32*)c->bufs[B_SEGBYTES].p; int* seg_prev = (int*)c->bufs[B_SEGPREV].p;
    const int n_iter = st.n_iter;
    const size_t cap = (size_t)body_main + (size_t)st.leftover;
    if (cap >= 0xffffffffull) return set_error(SCCG_E_ARG, "encoded output would exceed 4 GiB");
    u8* out = nullptr;
    SCCG_TRY(buf(c, B_OUT, cap + 16, &out));
    SCCG_CK(cudaMemsetAsync(sc + S_BODY_BASE, 0, sizeof(u32), c->stream));
    if (n_iter > 0) {
        unsigned want = div_up(n_iter, 8 * 32), capg = (unsigned)c->sm_count * 8u;
        LAUNCH(c, seg_write_k, dim3(want < capg ? want : capg), dim3(256), 0, st.d_tgt, st.nt, (const u32*)seginfo, (const u32*)matches,
               (const u32*)seg_bytes, (const int*)seg_prev, n_iter, out, (const u32*)(sc + S_BODY_BASE), 0, (int)st.seg_base, (const u32*)(sc + S_BODY_MAIN));
    }
    if (st.leftover > 0 && st.is_last) {                                      // :476-481 (only the last shard can have leftover target segments)
        unsigned g = div_up(st.leftover, 256 * 16), capg = (unsigned)c->sm_count * 8u;
        LAUNCH(c, upper_copy_k, dim3(g < capg ? g : capg), dim3(256), 0, st.d_tgt + (i64)n_iter * SEG, st.leftover, out, (const u32*)(sc + S_BODY_BASE), body_main, (const u32*)nullptr);
    }
    *d_body = out; *body_len = (i64)body_main + (st.is_last ? st.leftover : 0);
    st.valid = 0;
    return SCCG_OK;
}

static int shard_write(sccg_ctx* c, const sccg_shard_carry* carry, u8** d_low, i64* low_len, u8** d_body, i64* body_len) {
    SCCG_TRY(shard_write_sizes(c, carry, d_low));
    u32 h[S_COUNT];
    SCCG_TRY(read_scalars(c, (u32*)c->bufs[B_SCALARS].p, h, S_COUNT));
    *low_len = h[S_LOW_TEXT];
    return shard_write_body(c, h[S_BODY_MAIN], d_body, body_len);
}

}  // namespace sccg
