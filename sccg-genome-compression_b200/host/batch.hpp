// Batch / whole-genome driver shared by the two command-line programs: one process, many pairs, N GPUs.
//   * the pairs are packed onto the GPUs by size (sccg_mgpu_assign: longest processing time first);
//   * every GPU is driven by TWO worker threads with one sccg_ctx each: while one worker's pair is on the GPU, the other one
//     reads its next files from disk straight into page-locked memory (sccg_pinned_alloc) or writes its last result, so file
//     I/O, PCIe and kernels overlap without any pipeline logic in the workers;
//   * a worker's buffers are reused from pair to pair.
// The reference needs one process -- one CUDA-free but single-threaded run -- per pair (compression.cpp:584-610).
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "sccg.h"

namespace sccg_host {

struct PinnedBuf {
    char* p = nullptr; size_t cap = 0;
    ~PinnedBuf() { sccg_pinned_free(p); }
    bool ensure(size_t n) {
        if (n <= cap) return true;
        sccg_pinned_free(p);
        cap = n + n / 8 + 4096;
        p = (char*)sccg_pinned_alloc((int64_t)cap);
        if (!p) cap = 0;
        return p != nullptr;
    }
};

// whole file -> page-locked buffer (the copy to the GPU then runs at full PCIe speed); *n = bytes read, false: cannot open
inline bool read_file_pinned(const std::string& path, PinnedBuf& b, int64_t* n) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    size_t size = fstat(fd, &st) == 0 && st.st_size > 0 ? (size_t)st.st_size : 0;
    if (!b.ensure(size + 64)) { close(fd); return false; }
    size_t got = 0;
    for (;;) {
        if (got == b.cap - 64 && !b.ensure(b.cap * 2)) { close(fd); return false; }      // (a file that grew, or no size from fstat)
        ssize_t r = read(fd, b.p + got, b.cap - 64 - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    *n = (int64_t)got;
    return true;
}

inline bool write_all(int fd, const char* p, size_t n) {
    while (n) { ssize_t w = write(fd, p, n); if (w <= 0) return false; p += w; n -= (size_t)w; }
    return true;
}

inline int64_t file_size(const std::string& path) { struct stat st; return stat(path.c_str(), &st) == 0 ? (int64_t)st.st_size : 0; }

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Job { std::string a, b, out; int64_t weight = 0; };
struct Worker { sccg_ctx* ctx = nullptr; int device = 0; PinnedBuf in0, in1, outb; };
struct Timing { double read_s = 0, gpu_s = 0, write_s = 0, ext_s = 0; };

inline bool read_job_list(const char* path, std::vector<Job>* jobs) {
    std::ifstream list(path);
    if (!list.is_open()) return false;
    std::string line;
    while (std::getline(list, line)) {
        std::istringstream is(line);
        Job j;
        if (is >> j.a >> j.b >> j.out) jobs->push_back(j);                    // blank lines are skipped
    }
    return true;
}

// fn(worker, job, timing) -> 0 ok / 1 what the reference exits with.  Returns 0 iff every job succeeded.
template <class Fn>
int run_batch(std::vector<Job>& jobs, int n_gpus, Fn fn, std::ostream& log) {
    const int n = (int)jobs.size();
    if (n_gpus < 1) n_gpus = 1;
    std::vector<int64_t> w(n);
    for (int i = 0; i < n; ++i) w[i] = jobs[i].weight > 0 ? jobs[i].weight : 1;
    std::vector<int32_t> owner(n, 0);
    if (n && sccg_mgpu_assign(w.data(), n, n_gpus, owner.data()) != SCCG_OK) return 1;
    // per GPU: its jobs, largest first (the LPT order), handed out through one atomic cursor shared by its two workers
    std::vector<std::vector<int>> queue(n_gpus);
    for (int i = 0; i < n; ++i) queue[owner[i]].push_back(i);
    for (auto& q : queue) std::stable_sort(q.begin(), q.end(), [&](int x, int y) { return w[x] > w[y]; });
    std::vector<std::atomic<int>> cursor(n_gpus);
    for (auto& c : cursor) c = 0;
    std::atomic<int> failed{0};
    std::mutex log_m;
    const double t0 = now_s();
    auto work = [&](int dev, int slot) {
        Worker wk; wk.device = dev;
        if (queue[dev].empty()) return;
        wk.ctx = sccg_create(dev);
        if (!wk.ctx) { std::lock_guard<std::mutex> lk(log_m); log << "Error: " << sccg_last_error() << "\n"; failed++; return; }
        for (;;) {
            const int k = cursor[dev].fetch_add(1);
            if (k >= (int)queue[dev].size()) break;
            const int i = queue[dev][k];
            Timing tm;
            const double a = now_s();
            const int rc = fn(wk, jobs[i], tm);
            const double b = now_s();
            if (rc != 0) failed++;
            std::lock_guard<std::mutex> lk(log_m);
            log << "pair " << i << " -> " << jobs[i].out << ": " << (rc == 0 ? "ok" : "FAILED") << ", GPU " << dev << "." << slot << ", " << (b - a) << " s (read " << tm.read_s
                << ", gpu " << tm.gpu_s << ", write " << tm.write_s << ", 7z " << tm.ext_s << "), done at " << (b - t0) << " s\n";
        }
        sccg_destroy(wk.ctx);
    };
    std::vector<std::thread> th;
    for (int d = 0; d < n_gpus; ++d) for (int s = 0; s < 2; ++s) th.emplace_back(work, d, s);
    for (auto& t : th) t.join();
    log << "batch: " << n << " pairs on " << n_gpus << " GPU(s) in " << (now_s() - t0) << " s" << (failed ? " -- with failures" : "") << "\n";
    return failed ? 1 : 0;
}

// trailing "--gpus N" of a batch command line (default 1)
inline int parse_gpus(int argc, char** argv, int from) {
    for (int i = from; i + 1 < argc; ++i) if (std::string(argv[i]) == "--gpus") return atoi(argv[i + 1]);
    return 1;
}

}  // namespace sccg_host
