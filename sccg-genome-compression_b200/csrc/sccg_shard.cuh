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

static int shard_match(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, i64 seg_base, int is_last, sccg_shard_info* info) {
    memset(info, 0, sizeof *info);
    ShardState& st = c->shard;
    memset(&st, 0, sizeof st);
    u32* sc = nullptr;
    SCCG_TRY(buf(c, B_SCALARS, (size_t)S_COUNT, &sc));
    SCCG_CK(cudaMemsetAsync(sc, 0, sizeof(u32) * S_COUNT, c->stream));
    // lowercase runs of the slice (:341-367)
    u32 *cnt_s = nullptr, *cnt_e = nullptr; u64* low_mask = nullptr;
    SCCG_TRY(rle_count<0>(c, d_tgt, nt, B_RUN_CNT, B_RUN_MASK, &cnt_s, &cnt_e, &low_mask, sc + S_LOW_K, sc + S_LOW_KE, sc + S_PAREN));
    // segment matcher + driver bookkeeping (:381-474)
    const i64 n_rseg = (nr + SEG - 1) / SEG, n_tseg = (nt + SEG - 1) / SEG;
    const int n_iter = (int)(n_rseg < n_tseg ? n_rseg : n_tseg);
    u32 *seginfo = nullptr, *matches = nullptr, *seg_bytes = nullptr; int* seg_prev = nullptr;
    SCCG_TRY(buf(c, B_SEGINFO, (size_t)n_iter + 1, &seginfo));
    SCCG_TRY(buf(c, B_MATCH, (size_t)n_iter * LM_SLOT + 1, &matches));
    SCCG_TRY(buf(c, B_SEGBYTES, (size_t)n_iter + 1, &seg_bytes));
    SCCG_TRY(buf(c, B_SEGPREV, (size_t)n_iter + 1, &seg_prev));
    int* d_last = (int*)(sc + S_G6);
    if (n_iter > 0) {
        const size_t smem = sizeof(LmWarpSmem) * LM_WARPS;
        SCCG_SET_MAX_SMEM(seg_match_k<SCCG_LM_CLAIM>, smem);
        const unsigned cap = (unsigned)c->sm_count * (unsigned)LM_CTAS_PER_SM, w = div_up(n_iter, LM_WARPS);
        SCCG_CK(cudaMemsetAsync(seginfo, 0xff, sizeof(u32) * (size_t)n_iter, c->stream));
        SCCG_CK(cudaMemsetAsync(d_last, 0xff, sizeof(int), c->stream));
        LAUNCH(c, seg_match_k<SCCG_LM_CLAIM>, dim3(w < cap ? w : cap), dim3(LM_WARPS * 32), smem, d_ref, nr, d_tgt, nt, 0, n_iter, n_iter, K1, K2, seginfo, matches,
               sc + S_WORK, sc + S_ABORT, c->use_diag);
        LAUNCH(c, seg_bytes_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, (const u32*)matches, n_iter, seg_bytes, seg_prev, sc + S_ABORT, 0, 0, 0);
        LAUNCH(c, shard_last_match_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, n_iter, d_last);
    }
    u32 h[S_COUNT];
    SCCG_TRY(read_scalars(c, sc, h, S_COUNT));
    if (h[S_LOW_K] != h[S_LOW_KE]) return set_error(SCCG_E_CUDA, "internal: run start/end counts differ");
    info->n_segments = n_iter;
    info->abort_inside = h[S_ABORT] ? 1 : 0;
    info->has_paren = h[S_PAREN] ? 1 : 0;
    const u32 low_k = h[S_LOW_K];
    info->n_runs = low_k;
    // the runs themselves (needed for the border values now, for the text in sccg_shard_write)
    int *run_s = nullptr, *run_e = nullptr;
    SCCG_TRY(buf(c, B_RUN_START, (size_t)low_k + 1, &run_s));
    SCCG_TRY(buf(c, B_RUN_END, (size_t)low_k + 1, &run_e));
    const i64 tgt_off = seg_base * SEG;
    int border[4] = {0, 0, 0, 0};
    if (low_k) {
        LAUNCH(c, rle_write_k, dim3(div_up(nt > 0 ? nt : 1, RLE_TILE)), dim3(RLE_T), 0, (const u64*)low_mask, nt, (const u32*)cnt_s, (const u32*)cnt_e, run_s, run_e);
        SCCG_CK(cudaMemcpyAsync(&border[0], run_s, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaMemcpyAsync(&border[1], run_e, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaMemcpyAsync(&border[2], run_s + (low_k - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaMemcpyAsync(&border[3], run_e + (low_k - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    // border segments and the last match
    u32 head[4] = {0, 0, 0, 0}, tail[4] = {0, 0, 0, 0};
    const int nb = n_iter < 4 ? n_iter : 4;
    if (nb) {
        SCCG_CK(cudaMemcpyAsync(head, seginfo, sizeof(u32) * nb, cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaMemcpyAsync(tail, seginfo + (n_iter - nb), sizeof(u32) * nb, cudaMemcpyDeviceToHost, c->stream));
    }
    SCCG_CK(cudaStreamSynchronize(c->stream));
    if (low_k) {
        info->first_run_start = tgt_off + border[0]; info->first_run_len = border[1] - border[0];
        info->last_run_start = tgt_off + border[2]; info->last_run_len = border[3] - border[2];
    }
    auto status = [](u32 x) -> int {
        if (x == 0xffffffffu) return 0;                            // not processed (early abort): the shard reports abort_inside anyway
        const int inc = !SEGINFO_ALLN(x) && (SEGINFO_NMATCH(x) == 0 || SEGINFO_BAD(x));
        const int end = !SEGINFO_ALLN(x) && SEGINFO_NMATCH(x) == 0;
        return inc | (end << 1);
    };
    for (int i = 0; i < 4; ++i) { info->head_status[i] = i < nb ? status(head[i]) : 0; info->tail_status[i] = i < nb ? status(tail[i]) : 0; }
    const int last_seg = (int)h[S_G6];
    if (n_iter > 0 && last_seg >= 0) {
        u32 info_w = 0, pk = 0;
        SCCG_CK(cudaMemcpyAsync(&info_w, seginfo + last_seg, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaStreamSynchronize(c->stream));
        SCCG_CK(cudaMemcpyAsync(&pk, matches + (i64)last_seg * LM_SLOT + (SEGINFO_NMATCH(info_w) - 1), sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaStreamSynchronize(c->stream));
        info->has_match = 1;
        info->last_p = (int32_t)((last_seg + seg_base) * SEG + (i64)((pk >> 10) & 0x3ffu));
    }
    st.valid = 1; st.d_tgt = d_tgt; st.nt = nt; st.n_iter = n_iter; st.seg_base = seg_base; st.is_last = is_last; st.low_k = low_k;
    st.leftover = n_tseg > n_iter ? nt - (i64)n_iter * SEG : 0;
    return SCCG_OK;
}

static int shard_write(sccg_ctx* c, const sccg_shard_carry* carry, u8** d_low, i64* low_len, u8** d_body, i64* body_len) {
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
    u32 h[S_COUNT];
    SCCG_TRY(read_scalars(c, sc, h, S_COUNT));
    const size_t cap = (size_t)h[S_BODY_MAIN] + (size_t)st.leftover;
    if (cap >= 0xffffffffull) return set_error(SCCG_E_ARG, "encoded output would exceed 4 GiB");
    u8* out = nullptr;
    SCCG_TRY(buf(c, B_OUT, cap + 16, &out));
    SCCG_CK(cudaMemsetAsync(sc + S_BODY_BASE, 0, sizeof(u32), c->stream));
    if (n_iter > 0) {
        unsigned want = div_up(n_iter, 8 * 32), capg = (unsigned)c->sm_count * 8u;
        LAUNCH(c, seg_write_k, dim3(want < capg ? want : capg), dim3(256), 0, st.d_tgt, st.nt, (const u32*)seginfo, (const u32*)matches,
               (const u32*)seg_bytes, (const int*)seg_prev, n_iter, out, (const u32*)(sc + S_BODY_BASE), 0, (int)st.seg_base);
    }
    if (st.leftover > 0 && st.is_last) {                                      // :476-481 (only the last shard can have leftover target segments)
        unsigned g = div_up(st.leftover, 256 * 16), capg = (unsigned)c->sm_count * 8u;
        LAUNCH(c, upper_copy_k, dim3(g < capg ? g : capg), dim3(256), 0, st.d_tgt + (i64)n_iter * SEG, st.leftover, out, (const u32*)(sc + S_BODY_BASE), h[S_BODY_MAIN]);
    }
    *d_low = low_text; *low_len = h[S_LOW_TEXT];
    *d_body = out; *body_len = (i64)h[S_BODY_MAIN] + (st.is_last ? st.leftover : 0);
    st.valid = 0;
    return SCCG_OK;
}

}  // namespace sccg
