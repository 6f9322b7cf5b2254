// compress_genome on the device (compression.cpp:320-579 minus file I/O and 7z): orchestration of
// the run-list, segment-match, driver and writer kernels.  Global fallback lives in sccg_global.cuh.
#pragma once
#include "sccg_rle.cuh"
#include "sccg_local.cuh"
#include "sccg_delta.cuh"

namespace sccg {

// device scalars (u32 each)
enum Scalar {
    S_LOW_K = 0, S_LOW_KE, S_LOW_TEXT, S_ABORT, S_BODY_MAIN, S_BODY_BASE, S_N_K, S_N_KE, S_N_TEXT,
    S_G0, S_G1, S_G2, S_G3, S_G4, S_G5, S_G6, S_G7, S_PAREN, S_DT = 24 /* 4 slots */, S_QUEUE = 32 /* own 128-byte line: number of queued segments */, S_WORK = 64, S_COUNT = 128   // S_WORK: own 128-byte line (hot atomic)
};

// writes the separator after the lowercase line and publishes where the body starts
//   local : "<low>\n,\n<body>"   (compression.cpp:368)      global: "<low>\n<nruns>\n<body>" (:522, :555)
// Local mode is assembled BEFORE the host knows the outcome of the matcher: out_cap is the capacity the host reserved on a
// guess.  If the matcher aborted (-> global mode) or the body does not fit, the body base becomes BODY_BASE_NONE and the
// writer kernels behind this one return at once; the host reads the scalars once, at the end, and repeats the assembly
// in a larger buffer in the rare second case.
static const u32 BODY_BASE_NONE = 0xffffffffu;
__global__ void put_separators_k(u8* out, u32 hdr_bytes, u32* scalars, int global_mode, u32 out_cap, u32 leftover) {
    u32 low = scalars[S_LOW_TEXT];
    u8* o = out + hdr_bytes + low;
    if (!global_mode) {
        const unsigned long long need = (unsigned long long)hdr_bytes + low + 3ull + scalars[S_BODY_MAIN] + leftover;
        if (scalars[S_ABORT] || need > out_cap) { scalars[S_BODY_BASE] = BODY_BASE_NONE; return; }
        o[0] = '\n'; o[1] = ','; o[2] = '\n';
        scalars[S_BODY_BASE] = hdr_bytes + low + 3u;
    } else {
        u32 nt = scalars[S_N_TEXT];
        o[0] = '\n';                       // the N-run text is written at o + 1 by runs_write_k
        o[1 + nt] = '\n';
        scalars[S_BODY_BASE] = hdr_bytes + low + 1u + nt + 1u;
    }
}

struct CompressResult {
    u8* d_out;
    i64 out_len;
    int mode;
    int stoi_failed;        // delta_encode's stoi would have thrown (:279): d_out is the un-rewritten, pre-delta image
};

// the reference's delta_encode replayed at text level on a pre-delta image (sccg_delta.cuh); updates res
static int finish_text_delta(sccg_ctx* c, u32* sc, u32 body_base, CompressResult* res) {
    u8* fin = nullptr; i64 fin_len = 0; bool failed = false;
    SCCG_TRY(delta_text_pass(c, res->d_out, res->out_len, body_base, sc + S_DT, &fin, &fin_len, &failed));
    if (failed) res->stoi_failed = 1;
    else { res->d_out = fin; res->out_len = fin_len; }
    return SCCG_OK;
}

static int compress_global_device(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, const char* header, i64 nh,
                                  u32 low_k, const u8* d_low_text, int text_delta, CompressResult* res);

static int read_scalars(sccg_ctx* c, const u32* d_scalars, u32* host, int count) {
    SCCG_CK(cudaMemcpyAsync(c->h_pinned, d_scalars, sizeof(u32) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    SCCG_CK(cudaStreamSynchronize(c->stream));
    memcpy(host, c->h_pinned, sizeof(u32) * (size_t)count);
    return SCCG_OK;
}

// dst[0 .. *d_len) = src[0 .. *d_len): places a text whose length only the device knows yet
__global__ void copy_text_k(u8* __restrict__ dst, const u8* __restrict__ src, const u32* __restrict__ d_len) {
    const u32 n = *d_len;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

static int write_header(sccg_ctx* c, u8* d_out, const char* header, i64 nh) {
    if (nh <= 0) return SCCG_OK;                                              // compression.cpp:337-339
    if ((size_t)nh + 1 + 16384 > c->h_pinned_cap) return set_error(SCCG_E_ARG, "header line too long");
    char* stage = (char*)c->h_pinned + 16384;                                 // own part of the staging area: no sync needed here, every call ends with one
    memcpy(stage, header, (size_t)nh);
    stage[nh] = '\n';
    SCCG_CK(cudaMemcpyAsync(d_out, stage, (size_t)nh + 1, cudaMemcpyHostToDevice, c->stream));
    return SCCG_OK;
}

// host entry points: the inputs are still arriving on the copy stream when compress_device starts
struct ChunkArrival {
    i64 chunk;               // bytes per chunk of the sequence that arrives in chunks
    int n;                   // number of chunks
    cudaEvent_t* ev_ref;     // ev_ref[i]: bytes [0, (i + 1) * chunk) of the chunked sequence are resident
    cudaEvent_t ev_tgt;      // the whole target is resident
    int tgt_chunked;         // 0: the target arrives first and whole, the reference in chunks (sccg_compress*)
                             // 1: the reference is already resident, the TARGET arrives in chunks (sccg_compress_resident*)
};

// Host pair -> device, pipelined: the target goes up first (the run-list pipeline only needs it), then the reference in
// chunks on the copy stream; the matcher starts on every chunk as it lands, so the kernels hide under the PCIe transfer.
// Records ev[4] (start) and ev[5] (all copies done) for the profile.  The caller synchronises c->s_h2d before it returns.
static int enqueue_pair_upload(sccg_ctx* c, const char* ref, i64 ref_len, const char* tgt, i64 tgt_len, u8** d_ref_out, u8** d_tgt_out, ChunkArrival* arr) {
    u8 *d_ref = nullptr, *d_tgt = nullptr;
    SCCG_TRY(pipe_streams(c));
    SCCG_TRY(buf(c, B_REF, (size_t)ref_len + 128, &d_ref));
    SCCG_TRY(buf(c, B_TGT, (size_t)tgt_len + 128, &d_tgt));
    if (getenv("SCCG_PIPE_POISON")) {                                         // tests: a chunk read before it arrived shows up
        SCCG_CK(cudaMemsetAsync(d_tgt, 0xEE, (size_t)tgt_len, c->stream));
        SCCG_CK(cudaMemsetAsync(d_ref, 0xEE, (size_t)ref_len, c->stream));
    }
    SCCG_CK(cudaEventRecord(c->ev[4], c->stream));
    SCCG_CK(cudaStreamWaitEvent(c->s_h2d, c->ev[4], 0));
    arr->chunk = pipe_chunk_bytes(ref_len);
    arr->n = ref_len > 0 ? (int)((ref_len + arr->chunk - 1) / arr->chunk) : 0;
    arr->ev_ref = c->ev_h2d; arr->ev_tgt = c->ev_pipe[1]; arr->tgt_chunked = 0;
    if (tgt_len > 0) SCCG_CK(cudaMemcpyAsync(d_tgt, tgt, (size_t)tgt_len, cudaMemcpyHostToDevice, c->s_h2d));
    SCCG_CK(cudaMemsetAsync(d_tgt + tgt_len, 0, 64, c->s_h2d));
    SCCG_CK(cudaEventRecord(arr->ev_tgt, c->s_h2d));
    for (int i = 0; i < arr->n; ++i) {
        const i64 off = (i64)i * arr->chunk, len = (ref_len - off) < arr->chunk ? (ref_len - off) : arr->chunk;
        SCCG_CK(cudaMemcpyAsync(d_ref + off, ref + off, (size_t)len, cudaMemcpyHostToDevice, c->s_h2d));
        SCCG_CK(cudaEventRecord(arr->ev_ref[i], c->s_h2d));
    }
    SCCG_CK(cudaEventRecord(c->ev[5], c->s_h2d));
    *d_ref_out = d_ref; *d_tgt_out = d_tgt;
    return SCCG_OK;
}

static const int LM_PROBE_SEGS = 40;             // segments of the abort probe (see compress_device)
// two-phase matcher launches (sccg_local.cuh, LM_DEFER / LM_QUEUE) for pairs of this many segments and more
static inline int lm_two_phase_min() { const char* e = getenv("SCCG_LM_TWO_PHASE_MIN"); return e ? atoi(e) : 65536; }    // segments; tests force 0 / a huge value
static inline int lm_queue_ctas() { const char* e = getenv("SCCG_LM_QUEUE_CTAS"); int v = e ? atoi(e) : 0; return v >= 1 && v <= LM_CTAS_PER_SM ? v : LM_CTAS_PER_SM; }

static int compress_device(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, const char* header, i64 nh, CompressResult* res,
                           const ChunkArrival* arr) {
    u32* sc = nullptr;
    SCCG_TRY(buf(c, B_SCALARS, (size_t)S_COUNT, &sc));
    SCCG_CK(cudaMemsetAsync(sc, 0, sizeof(u32) * S_COUNT, c->stream));
    SCCG_CK(cudaEventRecord(c->ev[0], c->stream));

    // ---- lowercase runs of the raw target (:341-367): the whole run-list pipeline is HBM-bound and independent of the
    //      matcher, which is ALU-bound -> it runs on the side stream underneath seg_match_k
    u32 *cnt_s = nullptr, *cnt_e = nullptr;
    u64* low_mask = nullptr;
    SCCG_CK(cudaEventRecord(c->ev_side[0], c->stream));
    SCCG_CK(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
    if (arr) {                                                                // the run-list lane needs the whole target
        SCCG_CK(cudaStreamWaitEvent(c->side_stream, arr->ev_tgt, 0));
        if (!arr->tgt_chunked) SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_tgt, 0));     // chunked target: the matcher waits chunk by chunk
    }
    // ---- local segment matching (:381-474)
    const i64 n_rseg = (nr + SEG - 1) / SEG, n_tseg = (nt + SEG - 1) / SEG;
    const int n_iter = (int)(n_rseg < n_tseg ? n_rseg : n_tseg);             // :392
    // (Measured and dropped: preparing the inputs of global mode -- N-strip of both sequences -- on the side lane underneath the
    // probe for pairs of different length: the local attempt gets as much slower as the preparation takes, 1.19 ms either way.)
    const bool probe = !arr && n_iter > 4 * LM_PROBE_SEGS && (nr > nt ? nr - nt : nt - nr) >= SEG;
    // (measured and dropped: starting the run-list lane behind the bulk launch of a two-phase matcher -- the bulk launch got no
    // faster alone, the step went from 0.339 to 0.414 ms)
    const bool two_phase = !arr && c->use_diag && n_iter > 0 && n_iter >= lm_two_phase_min();
    {
        SideLane side(c);
        SCCG_TRY(rle_count<0>(c, d_tgt, nt, B_RUN_CNT, B_RUN_MASK, &cnt_s, &cnt_e, &low_mask, sc + S_LOW_K, sc + S_LOW_KE, sc + S_PAREN));
    }
    u32 *seginfo = nullptr, *matches = nullptr, *seg_bytes = nullptr;
    int* seg_prev = nullptr;
    SCCG_TRY(buf(c, B_SEGINFO, (size_t)n_iter + 1, &seginfo));
    SCCG_TRY(buf(c, B_MATCH, (size_t)n_iter * LM_SLOT + 1, &matches));
    SCCG_TRY(buf(c, B_SEGBYTES, (size_t)n_iter + 1, &seg_bytes));
    SCCG_TRY(buf(c, B_SEGPREV, (size_t)n_iter + 1, &seg_prev));
    SCCG_CK(cudaEventRecord(c->ev[1], c->stream));
    if (n_iter > 0) {
#ifndef SCCG_LM_EXTRA_SMEM
#define SCCG_LM_EXTRA_SMEM 0                  // development aid: occupancy sensitivity experiments
#endif
        const size_t smem = sizeof(LmWarpSmem) * LM_WARPS + SCCG_LM_EXTRA_SMEM;
        SCCG_SET_MAX_SMEM(seg_match_k<SCCG_LM_CLAIM>, smem);
        SCCG_SET_MAX_SMEM(seg_match_k<1>, smem);
        const unsigned cap = (unsigned)c->sm_count * (unsigned)LM_CTAS_PER_SM;   // resident CTAs: they claim segments dynamically
        SCCG_CK(cudaMemsetAsync(seginfo, 0xff, sizeof(u32) * (size_t)n_iter, c->stream));    // "not done" markers for the early T2 abort
        // Sequences of different length are usually shifted against each other from some indel / N-gap on, and then the tail
        // of the common range fails segment after segment.  The T2 abort condition (:454-473) is a property of 5 consecutive
        // segments, wherever they are, so a small probe launch over the last segments can raise the abort flag before the
        // main launch starts: every warp of the main launch then leaves after its first flag poll instead of burning one
        // failing (= most expensive) segment per resident warp.  No abort in the probe range: the main launch runs as usual.
        if (probe) {
            LAUNCH(c, seg_match_k<1>, dim3(div_up(LM_PROBE_SEGS, LM_WARPS)), dim3(LM_WARPS * 32), smem, d_ref, nr, d_tgt, nt, n_iter - LM_PROBE_SEGS, n_iter, n_iter, K1, K2,
                   seginfo, matches, sc + S_WORK + 31, sc + S_ABORT, c->use_diag);
        }
        // one launch per arrived reference chunk (device-resident inputs: a single launch)
        const int n_launch = arr ? arr->n : 1;
        int seg_lo = 0;
        for (int i = 0; i < n_launch && seg_lo < n_iter; ++i) {
            int seg_hi = n_iter;
            if (arr && i + 1 < n_launch) {
                const i64 resident = (i64)(i + 1) * arr->chunk;
                seg_hi = (int)(resident / SEG < (i64)n_iter ? resident / SEG : (i64)n_iter);
            }
            if (arr) SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_ref[i + 1 < n_launch ? i : arr->n - 1], 0));
            if (seg_hi <= seg_lo) continue;
            const unsigned w = div_up(seg_hi - seg_lo, LM_WARPS);
            if (two_phase) {
                // device-resident pair: the bulk launch queues the segments that need the generic path (the queue lives in
                // seg_bytes until seg_bytes_k overwrites it), a second launch with fewer warps per SM works them off
                SCCG_SET_MAX_SMEM(seg_match_defer_k, smem);
                SCCG_SET_MAX_SMEM(seg_match_queue_k, smem);
                LAUNCH(c, seg_match_defer_k, dim3(w < cap ? w : cap), dim3(LM_WARPS * 32), smem, d_ref, nr, d_tgt, nt, seg_lo, seg_hi, n_iter, K1, K2, seginfo, matches,
                       sc + S_WORK + (i & 31), sc + S_ABORT, c->use_diag, seg_bytes, sc + S_QUEUE);
                const unsigned qg = (unsigned)c->sm_count * (unsigned)lm_queue_ctas();
                LAUNCH(c, seg_match_queue_k, dim3(w < qg ? w : qg), dim3(LM_WARPS * 32), smem, d_ref, nr, d_tgt, nt, n_iter, K1, K2, seginfo, matches,
                       sc + S_WORK + 30, sc + S_ABORT, seg_bytes, sc + S_QUEUE);
            } else
            LAUNCH(c, seg_match_k<SCCG_LM_CLAIM>, dim3(w < cap ? w : cap), dim3(LM_WARPS * 32), smem, d_ref, nr, d_tgt, nt, seg_lo, seg_hi, n_iter, K1, K2, seginfo, matches,
                   sc + S_WORK + (i & 31), sc + S_ABORT, c->use_diag);
            seg_lo = seg_hi;
        }
        if (arr && arr->tgt_chunked) SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_tgt, 0));     // leftover target segments (:476-481) lie past the last launch
        SCCG_CK(cudaEventRecord(c->ev[2], c->stream));
        LAUNCH(c, seg_bytes_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, (const u32*)matches, n_iter, seg_bytes, seg_prev, sc + S_ABORT, 0, 0, 0);
    } else {
        if (arr && arr->tgt_chunked) SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_tgt, 0));
        SCCG_CK(cudaEventRecord(c->ev[2], c->stream));
    }
    SCCG_TRY(scan_exclusive_u32(c, seg_bytes, seg_bytes, (i64)n_iter, sc + S_BODY_MAIN));

    // side lane, while the matcher runs: run count -> runs -> "<lowercase runs>" text, staged (the global path places it
    // itself) and copied behind the header of the local-mode image.  The image buffer is reserved HERE, before the matcher
    // has finished: header + run text are known, the body is a guess (what the buffer already holds from earlier pairs, at
    // least 1/16 of the target); put_separators_k checks the guess on the device.
    const i64 leftover = n_tseg > n_iter ? nt - (i64)n_iter * SEG : 0;        // :476-481
    const size_t hdr_bytes = nh > 0 ? (size_t)nh + 1 : 0;
    u32 h[S_COUNT];
    u8 *low_text = nullptr, *out = nullptr;
    size_t out_cap = 0;
    auto reserve_out = [&](u32 low_k, size_t body) -> int {
        const size_t want = hdr_bytes + 24ull * low_k + 3 + body + (size_t)leftover;
        if (want >= 0xffffffffull) return set_error(SCCG_E_ARG, "encoded output would exceed 4 GiB");
        SCCG_TRY(buf(c, B_OUT, want + 16, &out));
        out_cap = c->bufs[B_OUT].cap - 80;                                    // buf() keeps 64 bytes of slack, we asked for 16 more
        if (out_cap > 0xfffffff0ull) out_cap = 0xfffffff0ull;
        return SCCG_OK;
    };
    {
        SideLane side(c);
        SCCG_TRY(read_scalars(c, sc, h, S_COUNT));                            // synchronises the side stream only
        if (h[S_LOW_K] != h[S_LOW_KE]) return set_error(SCCG_E_CUDA, "internal: run start/end counts differ");
        const u32 low_k = h[S_LOW_K];
        int *run_s = nullptr, *run_e = nullptr;
        SCCG_TRY(buf(c, B_RUN_TEXT, 24ull * low_k + 16, &low_text));
        SCCG_TRY(rle_emit<0>(c, low_mask, nt, low_k, cnt_s, cnt_e, sc + S_LOW_K, B_RUN_START, B_RUN_END, B_RUN_BYTES, &run_s, &run_e, low_text, sc + S_LOW_TEXT));
        size_t guess = (size_t)(nt / 16) + 65536;
        const char* env = getenv("SCCG_OUT_GUESS");                           // tests: a body guess in bytes, taken literally (forces the second assembly)
        if (env && atoll(env) >= 0) guess = (size_t)atoll(env);
        SCCG_TRY(reserve_out(low_k, guess));
        if (env) out_cap = hdr_bytes + 24ull * low_k + 3 + guess + (size_t)leftover;      // ... without the slack of the buffer
        if (low_k) LAUNCH(c, copy_text_k, dim3(low_k < 4096 ? 8 : (unsigned)c->sm_count * 16u), dim3(256), 0, out + hdr_bytes, (const u8*)low_text, (const u32*)(sc + S_LOW_TEXT));
        SCCG_CK(cudaEventRecord(c->ev_side[1], c->stream));
    }
    const u32 low_k = h[S_LOW_K];
    const int text_delta = h[S_PAREN] != 0;                                   // set by rle_count_k on the side lane (finished: the host waited for it)
    // a '(' somewhere in the target: tokens are written with absolute p and delta_encode is replayed at text level
    if (text_delta && n_iter > 0) {
        LAUNCH(c, seg_bytes_k, dim3(div_up(n_iter, 256)), dim3(256), 0, (const u32*)seginfo, (const u32*)matches, n_iter, seg_bytes, seg_prev, sc + S_ABORT, 1, 0, 0);
        SCCG_TRY(scan_exclusive_u32(c, seg_bytes, seg_bytes, (i64)n_iter, sc + S_BODY_MAIN));
    }

    // ---- assemble "<header>\n<lowercase runs>\n,\n<body>" (no host round trip between the matcher and the writers)
    auto assemble = [&](bool place_text) -> int {
        SCCG_TRY(write_header(c, out, header, nh));
        if (place_text && low_k) LAUNCH(c, copy_text_k, dim3(low_k < 4096 ? 8 : (unsigned)c->sm_count * 16u), dim3(256), 0, out + hdr_bytes, (const u8*)low_text, (const u32*)(sc + S_LOW_TEXT));
        LAUNCH(c, put_separators_k, dim3(1), dim3(1), 0, out, (u32)hdr_bytes, sc, 0, (u32)out_cap, (u32)leftover);
        if (n_iter > 0) {
            unsigned want = div_up(n_iter, 8 * 32);                              // 8 warps per CTA, 32 segments per warp
            unsigned capg = (unsigned)c->sm_count * 8u;
            LAUNCH(c, seg_write_k, dim3(want < capg ? want : capg), dim3(256), 0, d_tgt, nt, (const u32*)seginfo, (const u32*)matches,
                   (const u32*)seg_bytes, (const int*)seg_prev, n_iter, out, (const u32*)(sc + S_BODY_BASE), text_delta, 0, (const u32*)(sc + S_BODY_MAIN));
        }
        if (leftover > 0) {
            unsigned g = div_up(leftover, 256 * 16);
            unsigned capg = (unsigned)c->sm_count * 8u;
            LAUNCH(c, upper_copy_k, dim3(g < capg ? g : capg), dim3(256), 0, d_tgt + (i64)n_iter * SEG, leftover, out,
                   (const u32*)(sc + S_BODY_BASE), 0u, (const u32*)(sc + S_BODY_MAIN));
        }
        SCCG_CK(cudaEventRecord(c->ev[3], c->stream));
        return read_scalars(c, sc, h, S_COUNT);
    };
    const bool step_trace = getenv("SCCG_STEP_TRACE") != nullptr;             // development aid: where the step's time goes
    if (step_trace) SCCG_CK(cudaEventRecord(c->ev_x[0], c->stream));
    SCCG_CK(cudaStreamWaitEvent(c->stream, c->ev_side[1], 0));                // the run-list text is in place
    if (step_trace) SCCG_CK(cudaEventRecord(c->ev_x[1], c->stream));
    SCCG_TRY(assemble(false));
    if (step_trace) {
        float a = 0, b = 0, d = 0, e = 0, f = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]); cudaEventElapsedTime(&b, c->ev[1], c->ev[2]); cudaEventElapsedTime(&d, c->ev[2], c->ev_x[0]);
        cudaEventElapsedTime(&e, c->ev_x[0], c->ev_x[1]); cudaEventElapsedTime(&f, c->ev_x[1], c->ev[3]);
        fprintf(stderr, "step_trace: pre %.1f us, matcher %.1f, sizes %.1f, wait for the run-list lane %.1f, writers %.1f\n", a * 1e3, b * 1e3, d * 1e3, e * 1e3, f * 1e3);
    }
    if (h[S_ABORT]) {                                                         // :462-473 -> global (:484-574)
        if (arr && arr->n > 0) SCCG_CK(cudaStreamWaitEvent(c->stream, arr->ev_ref[arr->n - 1], 0));     // the global parse reads all of the reference
        return compress_global_device(c, d_ref, nr, d_tgt, nt, header, nh, low_k, low_text, text_delta, res);
    }
    if (h[S_BODY_BASE] == BODY_BASE_NONE) {                                   // the guess was too small: now the size is known
        SCCG_TRY(reserve_out(low_k, (size_t)h[S_BODY_MAIN]));
        SCCG_TRY(assemble(true));
        if (h[S_BODY_BASE] == BODY_BASE_NONE) return set_error(SCCG_E_CUDA, "internal: the encoded image does not fit its buffer");
    }
    res->d_out = out;
    res->out_len = (i64)hdr_bytes + h[S_LOW_TEXT] + 3 + h[S_BODY_MAIN] + leftover;
    res->mode = 0;
    res->stoi_failed = 0;
    if (text_delta) {
        SCCG_TRY(finish_text_delta(c, sc, (u32)(hdr_bytes + h[S_LOW_TEXT] + 3), res));
        SCCG_CK(cudaEventRecord(c->ev[3], c->stream));
        SCCG_CK(cudaStreamSynchronize(c->stream));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); c->prof.kernels_ms = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->prof.match_ms = ms;
    c->prof.serialize_ms = c->prof.kernels_ms - c->prof.match_ms;
    c->prof.mode = 0;
    return SCCG_OK;
}

}  // namespace sccg
