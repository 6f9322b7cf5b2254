// TEST INFRASTRUCTURE ONLY -- see simt_emu.h.
#include "simt_emu.h"

#include <sys/mman.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

// ---- context switch (x86-64 System V): saves callee-saved registers on the current stack ----
extern "C" void emu_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_ctx_switch
.type emu_ctx_switch,@function
emu_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_ctx_switch,.-emu_ctx_switch
)");

namespace emu {
thread_local BlockCtx* g_blk = nullptr;
const char* g_kernel_name = "?";

static const size_t kStackBytes = 64 * 1024;

[[noreturn]] void fail(const char* msg) {
    fprintf(stderr, "[simt_emu] FATAL: %s\n", msg);
    abort();
}

void yield() {
    BlockCtx* b = g_blk;
    emu_ctx_switch(&b->cur->sp, b->sched_sp);
}

static void fiber_entry() {
    BlockCtx* b = g_blk;
    (*b->body)();
    b = g_blk;
    b->cur->done = true;
    b->alive--;
    b->progress++;
    if (b->alive > 0 && b->bar_count == b->alive) { b->bar_count = 0; b->bar_gen++; }   // exited threads release a barrier
    emu_ctx_switch(&b->cur->sp, b->sched_sp);
    fail("resumed a finished fiber");
}

static void prepare_fiber(Fiber& f) {
    uintptr_t top = ((uintptr_t)(f.stack + kStackBytes)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                  // fake return address of fiber_entry
    *--sp = (void*)&fiber_entry;      // popped by `ret`
    for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = (void*)sp;
    f.done = false;
}

Slot* collective_arrive(uint32_t mask, uint64_t val) {
    BlockCtx* b = g_blk;
    const int lane = lane_id();
    const uint32_t bit = 1u << lane;
    if (!(mask & bit)) fail("warp collective: calling lane is not in its own mask");
    Slot* slots = b->slots[warp_id()];
    Slot* s = nullptr;
    for (;;) {
        for (int i = 0; i < 4 && !s; ++i)     // join an open rendezvous with the same mask
            if (!slots[i].complete && slots[i].arrived != 0 && slots[i].mask == mask && !(slots[i].arrived & bit)) s = &slots[i];
        for (int i = 0; i < 4 && !s; ++i)     // or open a new one
            if (!slots[i].complete && slots[i].arrived == 0 && slots[i].read_pending == 0) { s = &slots[i]; s->mask = mask; }
        if (s) break;
        yield();
    }
    s->arrived |= bit;
    s->vals[lane] = val;
    b->progress++;
    if (s->arrived == s->mask) { s->complete = true; s->read_pending = s->mask; }
    while (!s->complete) yield();
    return s;
}

void collective_release(Slot* s) {
    s->read_pending &= ~(1u << lane_id());
    if (s->read_pending == 0) { s->complete = false; s->arrived = 0; s->mask = 0; }
}

void block_barrier() {
    BlockCtx* b = g_blk;
    unsigned gen = b->bar_gen;
    b->progress++;
    if (++b->bar_count == b->alive) { b->bar_count = 0; b->bar_gen++; return; }
    while (b->bar_gen == gen) yield();
}

int block_barrier_count(int pred) {
    BlockCtx* b = g_blk;
    unsigned gen = b->bar_gen;
    int slot = (int)(gen & 1u);
    if (b->bar_count == 0) b->bar_red[slot] = 0;          // first arriver of this generation
    if (pred) b->bar_red[slot]++;
    b->progress++;
    if (++b->bar_count == b->alive) { b->bar_count = 0; b->bar_gen++; return b->bar_red[slot]; }
    while (b->bar_gen == gen) yield();
    return b->bar_red[slot];
}

// ---- per-worker resources ----
struct Worker {
    BlockCtx ctx;
    std::vector<Fiber> fibers;
    char* stacks = nullptr;
    size_t nstacks = 0;
    std::vector<unsigned char> smem;
    ~Worker() { if (stacks) munmap(stacks, nstacks * kStackBytes); }
    void ensure(int nthreads, size_t smem_bytes) {
        if ((size_t)nthreads > nstacks) {
            if (stacks) munmap(stacks, nstacks * kStackBytes);
            nstacks = (size_t)nthreads;
            stacks = (char*)mmap(nullptr, nstacks * kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (stacks == (char*)MAP_FAILED) fail("mmap of fiber stacks failed");
        }
        fibers.resize((size_t)nthreads);
        for (int i = 0; i < nthreads; ++i) fibers[(size_t)i].stack = stacks + (size_t)i * kStackBytes;
        if (smem.size() < smem_bytes + 16) smem.resize(smem_bytes + 16);
    }
};

static int sweep_order() {          // 0 ascending, 1 descending, 2 mixed: shakes out missing __syncwarp / __syncthreads
    const char* e = getenv("SCCG_EMU_ORDER");
    return e ? atoi(e) : 0;
}

static void run_block(Worker& w, dim3 grid, dim3 block, uint3 bidx, size_t smem_bytes, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    w.ensure(nthreads, smem_bytes);
    BlockCtx& b = w.ctx;
    b.grid = grid; b.block = block; b.bidx = bidx;
    b.nthreads = nthreads; b.alive = nthreads;
    b.fibers = w.fibers.data();
    b.bar_count = 0; b.bar_gen = 0; b.progress = 0;
    for (auto& ws : b.slots) for (auto& s : ws) { s.mask = 0; s.arrived = 0; s.read_pending = 0; s.complete = false; }
    b.dyn_smem = (unsigned char*)(((uintptr_t)w.smem.data() + 15) & ~(uintptr_t)15);
    b.body = &body;
    for (int i = 0; i < nthreads; ++i) {
        Fiber& f = w.fibers[(size_t)i];
        f.linear = i;
        f.tid.x = (unsigned)i % block.x;
        f.tid.y = ((unsigned)i / block.x) % block.y;
        f.tid.z = (unsigned)i / (block.x * block.y);
        prepare_fiber(f);
    }
    g_blk = &b;
    const int order = sweep_order();
    unsigned rng = 12345u + bidx.x * 7919u;
    unsigned long long sweeps = 0;
    while (b.alive > 0) {
        unsigned long long before = b.progress;
        bool desc = (order == 1) || (order == 2 && ((rng = rng * 1664525u + 1013904223u) >> 16 & 1));
        for (int j = 0; j < nthreads; ++j) {
            int i = desc ? nthreads - 1 - j : j;
            Fiber& f = w.fibers[(size_t)i];
            if (f.done) continue;
            b.cur = &f;
            emu_ctx_switch(&b.sched_sp, f.sp);
        }
        if (b.progress == before && b.alive > 0) {
            fprintf(stderr, "[simt_emu] %s: deadlock in block (%u,%u,%u): %d threads alive, barrier count %d, sweep %llu\n", g_kernel_name, bidx.x, bidx.y, bidx.z, b.alive, b.bar_count, sweeps);
            fail("deadlock (divergent barrier / collective with exited or non-arriving lanes)");
        }
        ++sweeps;
    }
    g_blk = nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    if (nblocks == 0) return;
    static int nworkers = [] { const char* e = getenv("SCCG_EMU_THREADS"); int n = e ? atoi(e) : (int)std::thread::hardware_concurrency(); return n < 1 ? 1 : (n > 64 ? 64 : n); }();
    int use = (int)std::min<unsigned long long>((unsigned long long)nworkers, nblocks);
    std::atomic<unsigned long long> next{0};
    auto work = [&] {
        Worker w;
        for (;;) {
            unsigned long long i = next.fetch_add(1);
            if (i >= nblocks) break;
            uint3 bidx;
            bidx.x = (unsigned)(i % grid.x);
            bidx.y = (unsigned)((i / grid.x) % grid.y);
            bidx.z = (unsigned)(i / ((unsigned long long)grid.x * grid.y));
            run_block(w, grid, block, bidx, smem_bytes, body);
        }
    };
    if (use == 1) { work(); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < use; ++t) ts.emplace_back(work);
    for (auto& t : ts) t.join();
}
}  // namespace emu

// ---- host runtime subset ----
struct emu_stream_t { int unused; };
struct emu_event_t { std::chrono::steady_clock::time_point t; };
cudaError_t cudaMalloc(void** p, size_t n) { *p = n ? aligned_alloc(256, (n + 255) & ~(size_t)255) : nullptr; return 0; }
cudaError_t cudaFree(void* p) { free(p); return 0; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return 0; }
cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return 0; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return 0; }
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new emu_stream_t(); return 0; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return 0; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
cudaError_t cudaDeviceSynchronize() { return 0; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event_t(); return 0; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return 0; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0; }
cudaError_t cudaGetLastError() { return 0; }
cudaError_t cudaPeekAtLastError() { return 0; }
const char* cudaGetErrorString(cudaError_t) { return "emu: no error"; }
cudaError_t cudaSetDevice(int) { return 0; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("SCCG_EMU_DEVICES"); *n = e && atoi(e) > 0 ? atoi(e) : 1; return 0; }      // tests of the multi-GPU host logic
cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 8; return 0; }
