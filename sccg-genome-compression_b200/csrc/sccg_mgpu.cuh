// Multi-GPU layer in C++: one rank per GPU (process or host thread), NCCL's C API underneath.
//   (1) by chromosome : LPT packing of the pairs onto the ranks (the reference's unit of work is one FASTA pair,
//                       compression.cpp:584-610); every rank compresses its own pairs, the encoded record streams -- and only
//                       those, about 2 % of the input -- are gathered to rank 0: sizes by ncclAllGather, payload by grouped
//                       ncclSend / ncclRecv.  No collective on a kernel's critical path.
//   (2) by segment range inside one chromosome (local path, compression.cpp:381-481): every rank uploads and matches its
//       slice (chunked upload, matcher per arrived chunk), the values that cross shard borders travel in ONE ncclAllGather
//       of a 128-byte record per rank, every rank derives the same carries (plan_carries), writes its part of the two
//       lines, and the parts go to rank 0 with grouped ncclSend / ncclRecv.
//   (3) decompression by output range: sccg_decompress_part per rank, nothing is gathered.
// NCCL is bound at run time (dlopen of libnccl.so.2): the single-GPU entry points do not depend on it.  The emulator build
// (tests) replaces the transport by an in-process hub between host threads so that the host logic runs without a GPU.
#pragma once
#include "sccg_shard.cuh"

#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <map>
#include <mutex>
#include <vector>

#ifndef SCCG_EMU
#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only; the symbols are resolved with dlsym
#endif

namespace sccg {

struct MgP2P { void* ptr; size_t bytes; int peer; };

struct MgTransport {
    virtual ~MgTransport() {}
    // every rank contributes `bytes` from d_send; d_recv receives world * bytes in rank order.  Device pointers, stream-ordered.
    virtual int all_gather(const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) = 0;
    // grouped point-to-point transfers: sends / receives between the same two ranks match in order
    virtual int p2p(const MgP2P* sends, int ns, const MgP2P* recvs, int nr, cudaStream_t s) = 0;
};

#ifndef SCCG_EMU
// ---- NCCL, resolved at run time ---------------------------------------------------------------------------------------
struct NcclApi {
    void* lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    static bool ok = false;
    std::call_once(once, [] {
        const char* names[] = {getenv("SCCG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { if (n && *n && (api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!api.lib) return;
#define SCCG_NCCL_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name); if (!api.field) return
        SCCG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        SCCG_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        SCCG_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        SCCG_NCCL_SYM(AllGather, "ncclAllGather");
        SCCG_NCCL_SYM(Send, "ncclSend");
        SCCG_NCCL_SYM(Recv, "ncclRecv");
        SCCG_NCCL_SYM(GroupStart, "ncclGroupStart");
        SCCG_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        SCCG_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef SCCG_NCCL_SYM
        ok = true;
    });
    return ok ? &api : nullptr;
}
#define SCCG_NCCL(call)                                                                                          \
    do {                                                                                                         \
        ncclResult_t r_ = (call);                                                                                \
        if (r_ != ncclSuccess) return sccg::set_error(SCCG_E_CUDA, "%s failed: %s", #call, nccl_api()->GetErrorString(r_)); \
    } while (0)

struct NcclTransport : MgTransport {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    ~NcclTransport() override { if (comm && nccl_api()) nccl_api()->CommDestroy(comm); }
    int all_gather(const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) override {
        SCCG_NCCL(nccl_api()->AllGather(d_send, d_recv, bytes, ncclUint8, comm, s));
        return SCCG_OK;
    }
    int p2p(const MgP2P* sends, int ns, const MgP2P* recvs, int nr, cudaStream_t s) override {
        if (ns + nr == 0) return SCCG_OK;
        SCCG_NCCL(nccl_api()->GroupStart());
        for (int i = 0; i < ns; ++i) SCCG_NCCL(nccl_api()->Send(sends[i].ptr, sends[i].bytes, ncclUint8, sends[i].peer, comm, s));
        for (int i = 0; i < nr; ++i) SCCG_NCCL(nccl_api()->Recv(recvs[i].ptr, recvs[i].bytes, ncclUint8, recvs[i].peer, comm, s));
        SCCG_NCCL(nccl_api()->GroupEnd());
        return SCCG_OK;
    }
};
#endif  // !SCCG_EMU

// ---- in-process hub: ranks are host threads of one process ("device" memory is plain memory in the emulator build; in the
//      product build it is only reachable when SCCG_MGPU_HUB=1 asks for it, for debugging on one GPU) -----------------------
struct MgHub {
    std::mutex m;
    std::condition_variable cv;
    int world = 0, arrived = 0, joined = 0, left = 0;
    unsigned gen = 0;
    const void* ptr[64];
    std::vector<MgP2P> sends[64];
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
static std::mutex g_hub_mutex;
static std::map<std::string, MgHub*> g_hubs;

struct HubTransport : MgTransport {
    MgHub* hub = nullptr;
    std::string key;
    int rank = 0, world = 1;
    ~HubTransport() override {
        std::lock_guard<std::mutex> lk(g_hub_mutex);
        if (hub && ++hub->left == hub->world) { g_hubs.erase(key); delete hub; }
    }
    int all_gather(const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) override {
        if (cudaStreamSynchronize(s) != cudaSuccess) return set_error(SCCG_E_CUDA, "hub transport: stream synchronisation failed");
        hub->ptr[rank] = d_send;
        hub->barrier();
        for (int r = 0; r < world; ++r)
            if (cudaMemcpy((u8*)d_recv + (size_t)r * bytes, hub->ptr[r], bytes, cudaMemcpyDefault) != cudaSuccess) return set_error(SCCG_E_CUDA, "hub transport: copy failed");
        hub->barrier();
        return SCCG_OK;
    }
    int p2p(const MgP2P* sends, int ns, const MgP2P* recvs, int nr, cudaStream_t s) override {
        if (cudaStreamSynchronize(s) != cudaSuccess) return set_error(SCCG_E_CUDA, "hub transport: stream synchronisation failed");
        hub->sends[rank].assign(sends, sends + ns);
        hub->barrier();
        int taken[64] = {0};
        int rc = SCCG_OK;
        for (int i = 0; i < nr && rc == SCCG_OK; ++i) {
            const int peer = recvs[i].peer;
            // the taken[peer]-th send of `peer` that is addressed to this rank
            int seen = 0; const MgP2P* src = nullptr;
            for (const MgP2P& sd : hub->sends[peer]) if (sd.peer == rank && seen++ == taken[peer]) { src = &sd; break; }
            ++taken[peer];
            if (!src || src->bytes != recvs[i].bytes) { rc = set_error(SCCG_E_CUDA, "hub transport: unmatched receive"); break; }
            if (src->bytes && cudaMemcpy(recvs[i].ptr, src->ptr, src->bytes, cudaMemcpyDefault) != cudaSuccess) rc = set_error(SCCG_E_CUDA, "hub transport: copy failed");
        }
        hub->barrier();
        return rc;
    }
};

static const int MG_MAX_ITEMS = 64;       // pairs per rank in one gather (a human genome has 24)

}  // namespace sccg

struct sccg_mgpu {
    sccg_ctx* ctx;
    int rank, world;
    sccg::MgTransport* tr;
    // outgoing encoded streams of this rank (device), in stash order
    unsigned char* d_stash; size_t stash_cap, stash_len;
    int n_items; int32_t item_id[sccg::MG_MAX_ITEMS]; long long item_len[sccg::MG_MAX_ITEMS];
    // exchange buffers
    long long *d_meta, *d_meta_all, *h_meta;      // per rank: count, total, (id, len) x MG_MAX_ITEMS
    unsigned char* d_recv; size_t recv_cap;
    sccg::ShardBorder *d_border_all, *h_border;   // world records each
    int last_sharded;
};

namespace sccg {

static const size_t MG_META_I64 = 2 + 2 * (size_t)MG_MAX_ITEMS;

static int mg_grow(unsigned char** p, size_t* cap, size_t need, size_t keep, cudaStream_t s) {
    if (need <= *cap) return SCCG_OK;
    size_t want = need + need / 4 + 4096;
    unsigned char* q = nullptr;
    if (cudaMalloc((void**)&q, want) != cudaSuccess) { cudaGetLastError(); return set_error(SCCG_E_NOMEM, "cudaMalloc of a multi-GPU exchange buffer failed"); }
    if (*p) {
        if (keep) SCCG_CK(cudaMemcpyAsync(q, *p, keep, cudaMemcpyDeviceToDevice, s));
        SCCG_CK(cudaStreamSynchronize(s));
        cudaFree(*p);
    }
    *p = q; *cap = want;
    return SCCG_OK;
}

// Longest-processing-time packing of the items onto `world` ranks; deterministic (ties: lower index, lower rank)
static void mg_assign(const int64_t* lengths, int n, int world, int32_t* owner) {
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lengths[a] > lengths[b]; });
    std::vector<long long> load(world, 0);
    for (int idx : order) {
        int best = 0;
        for (int r = 1; r < world; ++r) if (load[r] < load[best]) best = r;
        owner[idx] = best;
        load[best] += lengths[idx];
    }
}

// contiguous ranges [a, b) of segment-pair indices, one per rank (compression.cpp:385-392: n = min(#r, #t) pairs); false if the
// pair is too small to shard (every shard needs >= 8 pairs so that its border windows do not overlap)
static bool mg_segment_ranges(i64 ref_len, i64 tgt_len, int world, std::vector<i64>* a, std::vector<i64>* b) {
    const i64 n_rseg = (ref_len + SEG - 1) / SEG, n_tseg = (tgt_len + SEG - 1) / SEG;
    const i64 n_iter = n_rseg < n_tseg ? n_rseg : n_tseg;
    if (world < 2 || n_iter < 8 * (i64)world) return false;
    const i64 base = n_iter / world, rem = n_iter % world;
    i64 cur = 0;
    for (int r = 0; r < world; ++r) { a->push_back(cur); cur += base + (r < rem ? 1 : 0); b->push_back(cur); }
    return true;
}

// From the border reports of all shards: the carry of every shard; false when the pair has to take the unsharded path (T2
// abort -> global mode, compression.cpp:462-473; '(' in the target -> text-level delta_encode).  Deterministic: every rank
// computes the same plan from the all-gathered reports.
static bool mg_plan_carries(const ShardBorder* in, const std::vector<i64>& ra, const std::vector<i64>& rb, i64 tgt_len, std::vector<sccg_shard_carry>* out) {
    const int world = (int)ra.size();
    for (int r = 0; r < world; ++r) if (in[r].abort_inside || in[r].has_paren) return false;
    // T2 windows that cross a shard border: the counter exceeds 4 at a failed, non-all-N segment whose 4 predecessors all
    // incremented it (:417-424, :454-462).  Every shard has >= 8 segments, so a window touches at most two shards.
    for (int r = 1; r < world; ++r) {
        int seq[8];
        for (int i = 0; i < 4; ++i) { seq[i] = in[r - 1].tail_status[i]; seq[4 + i] = in[r].head_status[i]; }
        for (int e = 4; e < 8; ++e) {
            bool all = (seq[e] & 2) != 0;
            for (int x = e - 4; x <= e && all; ++x) all = (seq[x] & 1) != 0;
            if (all) return false;
        }
    }
    std::vector<i64> starts(world), ends(world);
    std::vector<char> t_start(world), t_end(world), whole(world), skip(world);
    for (int r = 0; r < world; ++r) {
        starts[r] = ra[r] * SEG;
        ends[r] = r + 1 < world ? rb[r] * SEG : tgt_len;
        t_start[r] = in[r].n_runs > 0 && in[r].first_run_start == starts[r];
        t_end[r] = in[r].n_runs > 0 && in[r].last_run_start + in[r].last_run_len == ends[r];
        whole[r] = in[r].n_runs == 1 && t_start[r] && t_end[r];
    }
    for (int r = 0; r < world; ++r) skip[r] = r > 0 && t_start[r] && t_end[r - 1];
    auto cont_reaches_end = [&](int q) {        // a run that enters shard q at its first symbol runs to the end of the target
        for (;; ++q) {
            if (!(t_start[q] && whole[q])) return false;
            if (q == world - 1) return true;
        }
    };
    out->assign(world, sccg_shard_carry());
    for (int r = 0; r < world; ++r) {
        sccg_shard_carry& cy = (*out)[r];
        memset(&cy, 0, sizeof cy);
        for (int q = r - 1; q >= 0; --q) if (in[q].has_match) { cy.prev_p = in[q].last_p; break; }
        if (t_end[r]) {
            for (int q = r + 1; q < world && t_start[q]; ++q) { cy.extra_last_len += in[q].first_run_len; if (!whole[q]) break; }
        }
        for (int q = r - 1; q >= 0; --q) if (in[q].n_runs - (skip[q] ? 1 : 0) > 0) { cy.prev_run_start = in[q].last_run_start; break; }   // shard q starts a run of its own
        cy.skip_first_run = skip[r] ? 1 : 0;
        cy.last_run_reaches_end = (t_end[r] && (r == world - 1 || cont_reaches_end(r + 1))) ? 1 : 0;
    }
    return true;
}

}  // namespace sccg

// ------------------------------------------------------------------------------------------------
// communicator set-up
// ------------------------------------------------------------------------------------------------
namespace sccg {

static bool mg_use_hub() {
#ifdef SCCG_EMU
    return true;
#else
    const char* e = getenv("SCCG_MGPU_HUB");
    return e && atoi(e) != 0;
#endif
}

static int mg_unique_id(char* id128) {
    memset(id128, 0, SCCG_MGPU_ID_BYTES);
    if (mg_use_hub()) {
        static std::mutex m; static unsigned long long counter = 0;
        std::lock_guard<std::mutex> lk(m);
        snprintf(id128, SCCG_MGPU_ID_BYTES, "sccg-hub-%d-%llu", (int)getpid(), ++counter);
        return SCCG_OK;
    }
#ifndef SCCG_EMU
    NcclApi* api = nccl_api();
    if (!api) return set_error(SCCG_E_CUDA, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == SCCG_MGPU_ID_BYTES, "ncclUniqueId is 128 bytes");
    SCCG_NCCL(api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
#endif
    return SCCG_OK;
}

static int mg_init(sccg_ctx* c, const char* id128, int rank, int world, sccg_mgpu** out) {
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return set_error(SCCG_E_ARG, "rank / world out of range (1 <= world <= 64)");
    SCCG_CK(cudaSetDevice(c->device));
    sccg_mgpu* g = new (std::nothrow) sccg_mgpu();
    if (!g) return set_error(SCCG_E_NOMEM, "out of host memory");
    memset(g, 0, sizeof *g);
    g->ctx = c; g->rank = rank; g->world = world;
    if (mg_use_hub()) {
        HubTransport* t = new HubTransport();
        t->rank = rank; t->world = world; t->key.assign(id128, strnlen(id128, SCCG_MGPU_ID_BYTES));
        {
            std::lock_guard<std::mutex> lk(g_hub_mutex);
            MgHub*& h = g_hubs[t->key];
            if (!h) { h = new MgHub(); h->world = world; }
            t->hub = h;
        }
        g->tr = t;
    } else {
#ifndef SCCG_EMU
        NcclApi* api = nccl_api();
        if (!api) { delete g; return set_error(SCCG_E_CUDA, "NCCL (libnccl.so.2) could not be loaded"); }
        NcclTransport* t = new NcclTransport();
        t->rank = rank; t->world = world;
        ncclUniqueId id; memcpy(&id, id128, sizeof id);
        ncclResult_t r = api->CommInitRank(&t->comm, world, id, rank);
        if (r != ncclSuccess) { delete t; delete g; return set_error(SCCG_E_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
        g->tr = t;
#endif
    }
    const size_t meta_bytes = sizeof(long long) * MG_META_I64;
    bool ok = cudaMalloc((void**)&g->d_meta, meta_bytes) == cudaSuccess && cudaMalloc((void**)&g->d_meta_all, meta_bytes * (size_t)world) == cudaSuccess &&
              cudaMallocHost((void**)&g->h_meta, meta_bytes * (size_t)(world + 1)) == cudaSuccess &&
              cudaMalloc((void**)&g->d_border_all, sizeof(ShardBorder) * (size_t)world) == cudaSuccess &&
              cudaMallocHost((void**)&g->h_border, sizeof(ShardBorder) * (size_t)world) == cudaSuccess;
    if (!ok) { cudaGetLastError(); delete g->tr; delete g; return set_error(SCCG_E_NOMEM, "allocation of the multi-GPU exchange buffers failed"); }
    *out = g;
    return SCCG_OK;
}

static void mg_destroy(sccg_mgpu* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->main_stream);
    delete g->tr;
    if (g->d_stash) cudaFree(g->d_stash);
    if (g->d_recv) cudaFree(g->d_recv);
    if (g->d_meta) cudaFree(g->d_meta);
    if (g->d_meta_all) cudaFree(g->d_meta_all);
    if (g->h_meta) cudaFreeHost(g->h_meta);
    if (g->d_border_all) cudaFree(g->d_border_all);
    if (g->h_border) cudaFreeHost(g->h_border);
    delete g;
}

// ------------------------------------------------------------------------------------------------
// (1) by chromosome: stash the encoded stream of every pair of this rank, gather once
// ------------------------------------------------------------------------------------------------
static int mg_stash(sccg_mgpu* g, int32_t item, const void* d_data, i64 len) {
    sccg_ctx* c = g->ctx;
    if (g->n_items >= MG_MAX_ITEMS) return set_error(SCCG_E_ARG, "too many stashed items on this rank (64 per gather)");
    SCCG_TRY(mg_grow(&g->d_stash, &g->stash_cap, g->stash_len + (size_t)len + 64, g->stash_len, c->main_stream));
    if (len > 0) SCCG_CK(cudaMemcpyAsync(g->d_stash + g->stash_len, d_data, (size_t)len, cudaMemcpyDeviceToDevice, c->main_stream));
    g->item_id[g->n_items] = item; g->item_len[g->n_items] = len; ++g->n_items;
    g->stash_len += (size_t)len;
    return SCCG_OK;
}

// collective.  Rank 0: out receives the streams back to back, (ids, offs, lens) describe them in (rank, stash order).
// out == NULL with d_out != NULL: the gathered streams stay in rank 0's device memory (*d_out, owned by the communicator, valid until its next gather)
static int mg_gather(sccg_mgpu* g, char* out, i64 out_cap, void** d_out, int32_t* ids, int64_t* offs, int64_t* lens, int32_t cap_items, int32_t* n_items, int64_t* total) {
    sccg_ctx* c = g->ctx;
    cudaStream_t s = c->main_stream;
    const int W = g->world;
    SCCG_CK(cudaEventRecord(c->ev_x[0], s));
    long long* mine = g->h_meta;                       // [0]: own record, [1 .. W]: everybody's
    memset(mine, 0, sizeof(long long) * MG_META_I64);
    mine[0] = g->n_items; mine[1] = (long long)g->stash_len;
    for (int i = 0; i < g->n_items; ++i) { mine[2 + 2 * i] = g->item_id[i]; mine[3 + 2 * i] = g->item_len[i]; }
    const size_t mb = sizeof(long long) * MG_META_I64;
    SCCG_CK(cudaMemcpyAsync(g->d_meta, mine, mb, cudaMemcpyHostToDevice, s));
    SCCG_TRY(g->tr->all_gather(g->d_meta, g->d_meta_all, mb, s));
    long long* all = g->h_meta + MG_META_I64;
    SCCG_CK(cudaMemcpyAsync(all, g->d_meta_all, mb * (size_t)W, cudaMemcpyDeviceToHost, s));
    SCCG_CK(cudaStreamSynchronize(s));
    int rc = SCCG_OK;
    if (g->rank == 0) {
        i64 sum = 0; int cnt = 0;
        std::vector<i64> roff(W);
        for (int r = 0; r < W; ++r) { roff[r] = sum; sum += all[r * MG_META_I64 + 1]; cnt += (int)all[r * MG_META_I64]; }
        if (total) *total = sum;
        if (n_items) *n_items = cnt;
        if (cnt > cap_items || (out && sum > out_cap)) rc = set_error(SCCG_E_ARG, "gather: output arrays too small (required sizes returned)");
        // the receives are posted even when the caller's buffers are too small: the other ranks are already sending
        SCCG_TRY(mg_grow(&g->d_recv, &g->recv_cap, (size_t)sum + 64, 0, s));
        std::vector<MgP2P> recvs;
        for (int r = 1; r < W; ++r) if (all[r * MG_META_I64 + 1] > 0) recvs.push_back(MgP2P{g->d_recv + roff[r], (size_t)all[r * MG_META_I64 + 1], r});
        if (g->stash_len) SCCG_CK(cudaMemcpyAsync(g->d_recv, g->d_stash, g->stash_len, cudaMemcpyDeviceToDevice, s));
        SCCG_TRY(g->tr->p2p(nullptr, 0, recvs.data(), (int)recvs.size(), s));
        if (rc == SCCG_OK) {
            if (sum > 0 && out) SCCG_CK(cudaMemcpyAsync(out, g->d_recv, (size_t)sum, cudaMemcpyDeviceToHost, s));
            if (d_out) *d_out = g->d_recv;
            int k = 0;
            for (int r = 0; r < W; ++r) {
                i64 o = roff[r];
                for (int i = 0; i < (int)all[r * MG_META_I64]; ++i, ++k) {
                    ids[k] = (int32_t)all[r * MG_META_I64 + 2 + 2 * i]; offs[k] = o; lens[k] = all[r * MG_META_I64 + 3 + 2 * i];
                    o += lens[k];
                }
            }
        }
    } else {
        MgP2P sd{g->d_stash, g->stash_len, 0};
        if (g->stash_len) SCCG_TRY(g->tr->p2p(&sd, 1, nullptr, 0, s));
        if (total) *total = 0;
        if (n_items) *n_items = 0;
    }
    SCCG_CK(cudaEventRecord(c->ev_x[1], s));
    SCCG_CK(cudaStreamSynchronize(s));
    memset(&c->prof, 0, sizeof c->prof);
    cudaEventElapsedTime(&c->prof.exchange_ms, c->ev_x[0], c->ev_x[1]);
    c->prof.kernels_ms = c->prof.exchange_ms;
    g->n_items = 0; g->stash_len = 0;
    return rc;
}

}  // namespace sccg
