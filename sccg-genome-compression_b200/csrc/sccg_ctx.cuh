// Context, grow-only device buffers, error plumbing, launch accounting.
#pragma once
#include "sccg_common.cuh"
#include "../../include/sccg.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

namespace sccg {

static thread_local std::string g_last_error;

static int set_error(int code, const char* fmt, const char* a = "", const char* b = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    g_last_error = buf;
    return code;
}

#define SCCG_CK(call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return sccg::set_error(SCCG_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// grow-only device buffer slots (one cudaMalloc per slot after warm-up)
enum Slot {
    B_REF = 0, B_TGT, B_OUT, B_OUT2, B_SEGINFO, B_MATCH, B_SEGBYTES, B_SEGPREV, B_SCAN0, B_SCAN1, B_SCAN2, B_SCALARS,
    B_RUN_CNT, B_RUN_START, B_RUN_END, B_RUN_BYTES, B_RUN_TEXT, B_NRUN_CNT, B_NRUN_START, B_NRUN_END, B_NRUN_BYTES, B_NRUN_TEXT,
    B_ENC, B_NIDX, B_LOW, B_TOK_FLAG, B_TOK_POS, B_ITEM_OFF, B_ITEM_SRC, B_NUM0, B_NUM1, B_NUM2, B_NUM3, B_NUM4, B_NUM5,
    B_LRUN_S, B_LRUN_E, B_NRUNS_S, B_NRUNS_E, B_NRUNS_CUM, B_TILE0, B_TILE1, B_TILE2, B_TILE3,
    B_GREF, B_GTGT, B_GKEYS, B_GVALS, B_GKEYS2, B_GVALS2, B_GHIST, B_GOFFS, B_GREC, B_GLIT, B_GTMP0, B_GTMP1, B_GTMP2, B_GTMP3,
    B_NSLOTS
};

struct DevBuf { void* p; size_t cap; };

}  // namespace sccg

struct sccg_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    sccg::DevBuf bufs[sccg::B_NSLOTS];
    void* h_pinned;            // small pinned staging area for scalars
    size_t h_pinned_cap;
    cudaEvent_t ev[8];
    sccg_profile prof;
    unsigned scan_epoch;           // single-pass scan: epoch of the last launch, tiles handed out so far
    unsigned scan_counter_base;
    int scan_counter_ready;
    int use_diag;                  // seg_match_k: try the diagonal-hypothesis parse first (SCCG_NO_DIAG=1 disables it)
};

namespace sccg {

static int buf_reserve(sccg_ctx* c, int slot, size_t bytes, void** out) {
    DevBuf& b = c->bufs[slot];
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes < 256) bytes = 256;
    if (b.cap < bytes) {
        if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
        size_t want = bytes + bytes / 8;                       // a little headroom against regrowth
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&b.p, want = bytes); }
        if (e != cudaSuccess) { b.p = nullptr; return set_error(SCCG_E_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
        b.cap = want;
    }
    *out = b.p;
    return SCCG_OK;
}
template <typename T> static int buf(sccg_ctx* c, int slot, size_t count, T** out) {
    void* p = nullptr;
    int rc = buf_reserve(c, slot, count * sizeof(T) + 64, &p);   // +64: vector loads may run past the end
    *out = (T*)p;
    return rc;
}

#define SCCG_TRY(expr) do { int rc_ = (expr); if (rc_ != SCCG_OK) return rc_; } while (0)

// every kernel launch goes through here so that gpu_launches can be reported
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                      \
    do {                                                                 \
        SCCG_LAUNCH(kernel, grid, block, smem, (ctx)->stream, __VA_ARGS__); \
        (ctx)->prof.launches++;                                          \
    } while (0)

static inline unsigned div_up(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// copies a host buffer into a grow-only device slot (pageable source: cudaMemcpyAsync stages it)
static int upload(sccg_ctx* c, int slot, const void* h, i64 n, u8** d) {
    SCCG_TRY(buf(c, slot, (size_t)n + 64, d));
    if (n > 0) SCCG_CK(cudaMemcpyAsync(*d, h, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return SCCG_OK;
}

static int download(sccg_ctx* c, const u8* d, i64 n, char** out) {
    char* h = (char*)malloc((size_t)n + 1);
    if (!h) return set_error(SCCG_E_NOMEM, "malloc of the result failed");
    if (n > 0) {
        cudaError_t e = cudaMemcpyAsync(h, d, (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { free(h); return set_error(SCCG_E_CUDA, "result download failed: %s", cudaGetErrorString(e)); }
    }
    h[n] = 0;
    *out = h;
    return SCCG_OK;
}

// result delivery: either into a buffer of the caller (dst != NULL; pin it for full PCIe speed) or into a fresh malloc
static int deliver(sccg_ctx* c, const u8* d, i64 n, char* dst, i64 dst_cap, char** out_alloc, int64_t* out_len) {
    *out_len = n;
    if (!dst) return download(c, d, n, out_alloc);
    if (dst_cap < n) return set_error(SCCG_E_ARG, "output buffer too small (required size returned in *out_len)");
    if (n > 0) {
        SCCG_CK(cudaMemcpyAsync(dst, d, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        SCCG_CK(cudaStreamSynchronize(c->stream));
    }
    return SCCG_OK;
}



}  // namespace sccg
