// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Harness around the UNMODIFIED reference translation unit decompression.cpp, #included from
// where it lies (-DSCCG_REF_DECOMPRESSION_CPP="\"/root/reference/decompression.cpp\"") with
// its main() renamed, so tests can call reconstruct_genome (decompression.cpp:117) directly.
#define main sccg_ref_decompress_main
#include SCCG_REF_DECOMPRESSION_CPP
#undef main

#include <fcntl.h>
#include <unistd.h>
#include <cstring>

namespace {
struct StdoutSilencerD {
    int saved = -1;
    StdoutSilencerD() {
        std::cout.flush();
        fflush(stdout);
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        if (devnull >= 0) { dup2(devnull, 1); close(devnull); }
    }
    ~StdoutSilencerD() {
        std::cout.flush();
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};
}  // namespace

extern "C" {

// reconstruct_genome(reference, encoded, n_indices, lowercase_indices) -> wrapped text.
// `ref` must already be prepared the way decompress_genome does it (decompression.cpp:105-110:
// N-stripped unless the N line is ",", then upper-cased).  Returns 1 if the reference threw
// (e.g. stoi).  NOTE: the bounds error path calls exit(1) (decompression.cpp:223-229) and
// cannot be intercepted here -- use the executable for that case.
// seconds_out: wall time of the reconstruct_genome call alone.
int sccg_ref_reconstruct_genome(const char* ref, long nr, const char* enc, long ne,
                                const char* n_idx, long nn, const char* low_idx, long nl,
                                char** out, long* out_len, double* seconds_out) {
    try {
        StdoutSilencerD quiet;
        std::string r(ref, (size_t)nr), e(enc, (size_t)ne), n(n_idx, (size_t)nn), l(low_idx, (size_t)nl);
        auto t0 = std::chrono::high_resolution_clock::now();
        std::string res = reconstruct_genome(r, e, n, l);
        auto t1 = std::chrono::high_resolution_clock::now();
        if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
        *out = (char*)malloc(res.size() + 1);
        memcpy(*out, res.data(), res.size());
        *out_len = (long)res.size();
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

}  // extern "C"
