// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Harness around the UNMODIFIED reference translation unit compression.cpp.  The reference
// source is not copied into this repository: it is #included from where it lies (the path is
// passed by oracle/Makefile as -DSCCG_REF_COMPRESSION_CPP="\"/root/reference/compression.cpp\"")
// with its main() renamed, so that tests can call the reference's own free functions
//   match_sequences   (compression.cpp:36)
//   delta_encode      (compression.cpp:222)
//   compress_genome   (compression.cpp:320)
// directly.  Output of the build goes to oracle/_ref/ only (git-ignored, travels with gpurun).
#define main sccg_ref_compress_main
#include SCCG_REF_COMPRESSION_CPP
#undef main

#include <fcntl.h>
#include <unistd.h>
#include <cstring>

namespace {
// The reference prints O(tokens) DEBUG text on stdout (compression.cpp:286, :396, ...); silence
// fd 1 for the duration of a call.
struct StdoutSilencer {
    int saved = -1;
    StdoutSilencer() {
        std::cout.flush();
        fflush(stdout);
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        if (devnull >= 0) { dup2(devnull, 1); close(devnull); }
    }
    ~StdoutSilencer() {
        std::cout.flush();
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};
}  // namespace

extern "C" {

// Calls the reference match_sequences and flattens vector<Position> into caller-freed arrays.
// Record i is a match iff lit_off[i+1] == lit_off[i]; p/l are start_reference/length.
int sccg_ref_match_sequences(const char* Sr, long nr, const char* St, long nt, int k, int m,
                             int global, int offset, long* n_out, int** p_out, int** l_out,
                             long** lit_off_out, char** lits_out) {
    try {
        StdoutSilencer quiet;
        std::string sr(Sr, (size_t)nr), st(St, (size_t)nt);
        std::vector<Position> res = match_sequences(sr, st, k, m, global != 0, offset);
        long n = (long)res.size();
        int* p = (int*)malloc(sizeof(int) * (n + 1));
        int* l = (int*)malloc(sizeof(int) * (n + 1));
        long* off = (long*)malloc(sizeof(long) * (n + 2));
        long total = 0;
        for (long i = 0; i < n; ++i) total += (long)res[i].mismatch.size();
        char* lits = (char*)malloc((size_t)total + 1);
        long cur = 0;
        for (long i = 0; i < n; ++i) {
            p[i] = res[i].start_reference;
            l[i] = res[i].length;
            off[i] = cur;
            memcpy(lits + cur, res[i].mismatch.data(), res[i].mismatch.size());
            cur += (long)res[i].mismatch.size();
        }
        off[n] = cur;
        *n_out = n; *p_out = p; *l_out = l; *lit_off_out = off; *lits_out = lits;
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Runs the reference delta_encode on a file in place.
int sccg_ref_delta_encode(const char* path) {
    try {
        StdoutSilencer quiet;
        delta_encode(std::string(path));
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Runs the reference compress_genome (FASTA read, match, write, delta_encode, `7z a`).  The
// caller must have a `7z` on PATH (oracle/7z_shim.sh installed as oracle/_ref/bin/7z), otherwise
// the reference calls exit(1) after writing compressed_genome.txt (compression.cpp:311-313).
// seconds_out receives the wall time of the call, measured like compression.cpp:597-601.
int sccg_ref_compress_genome(const char* ref_path, const char* tgt_path, const char* out_dir,
                             double* seconds_out) {
    try {
        StdoutSilencer quiet;
        auto t0 = std::chrono::high_resolution_clock::now();
        compress_genome(std::string(ref_path), std::string(tgt_path), std::string(out_dir));
        auto t1 = std::chrono::high_resolution_clock::now();
        if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// FASTA reader of the compress side (compression.cpp:181-220); header_out may be NULL.
int sccg_ref_read_genomes(const char* ref_path, const char* tgt_path, char** ref_out, long* nr,
                          char** tgt_out, long* nt, char** header_out, long* nh) {
    try {
        StdoutSilencer quiet;
        std::string r, t, h;
        read_genomes_from_files(std::string(ref_path), std::string(tgt_path), r, t, h);
        *ref_out = (char*)malloc(r.size() + 1); memcpy(*ref_out, r.data(), r.size()); *nr = (long)r.size();
        *tgt_out = (char*)malloc(t.size() + 1); memcpy(*tgt_out, t.data(), t.size()); *nt = (long)t.size();
        if (header_out) {
            *header_out = (char*)malloc(h.size() + 1); memcpy(*header_out, h.data(), h.size()); *nh = (long)h.size();
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

void sccg_ref_free(void* p) { free(p); }

}  // extern "C"
