// compress <reference_file> <target_file> <output_folder>
// Drop-in for the reference's compression.cpp main (:584-610) with the hot path (match-and-encode) on a B200 through
// libsccg_b200.so.  Same argv, same output files (<out>/compressed_genome.txt, then the external `7z a -mx=9`), same
// exit codes.  There is no CPU matcher: without a usable GPU the program fails.
//
// Additive: `compress --batch <list>` runs many pairs in one process -- one line per pair,
// "<reference_file> <target_file> <output_folder>" -- so the CUDA context (about a second per process) is paid once.
#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>

#include "fasta_io.hpp"
#include "sccg.h"

// one pair: compress_genome (:320-582) with the reference's files and messages.  0 = ok, 1 = what the reference exits with.
static int compress_pair(sccg_ctx* ctx, const std::string& ref_path, const std::string& tgt_path, const std::string& out_dir) {
    if (!std::filesystem::exists(out_dir)) std::filesystem::create_directory(out_dir);
    auto t0 = std::chrono::high_resolution_clock::now();
    // the raw FASTA images go to the GPU as they are: read_genomes_from_files (:181-220) runs there
    std::string ref_file, tgt_file;
    if (!sccg_host::read_file(ref_path, ref_file)) { std::cerr << "Error opening reference file: " << ref_path << "\n"; return 1; }
    if (!sccg_host::read_file(tgt_path, tgt_file)) { std::cerr << "Error opening target file: " << tgt_path << "\n"; return 1; }
    char* out = nullptr; int64_t out_len = 0; int mode = 0;
    int rc = sccg_compress_fasta(ctx, ref_file.data(), (int64_t)ref_file.size(), tgt_file.data(), (int64_t)tgt_file.size(), &out, &out_len, &mode);
    if (rc != SCCG_OK && rc != SCCG_E_STOI) { std::cerr << "Error: " << sccg_last_error() << "\n"; return 1; }
    sccg_profile prof; sccg_get_profile(ctx, &prof);

    std::filesystem::create_directories(out_dir);                            // :334
    const std::string txt = out_dir + "/compressed_genome.txt";
    FILE* f = fopen(txt.c_str(), "wb");
    if (!f || fwrite(out, 1, (size_t)out_len, f) != (size_t)out_len) { std::cerr << "Greska pri otvaranju datoteke: " << txt << "\n"; if (f) fclose(f); sccg_free(out); return 1; }
    fclose(f);
    sccg_free(out);
    if (rc == SCCG_E_STOI) { std::cerr << "Error: stoi\n"; return 1; }         // delta_encode threw (:279 -> :604-607): file left un-rewritten, no 7z
    std::cout << "mode: " << (mode ? "global" : "local") << ", GPU kernels " << prof.kernels_ms << " ms, H2D " << prof.h2d_ms << " ms, D2H "
              << prof.d2h_ms << " ms\n";

    const std::string cmd = "7z a -mx=9 \"" + txt + ".7z\" \"" + txt + "\"";   // :308, the external stage stays as it is
    if (system(cmd.c_str()) != 0) { std::cerr << "Greska prilikom komprimiranja datoteke 7-zipom !\n"; return 1; }
    std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
    std::cout << "Time taken to compress: " << dt.count() << " s\n";           // :602
    return 0;
}

int main(int argc, char* argv[]) {
    const bool batch = argc == 3 && std::string(argv[1]) == "--batch";
    if (argc != 4 && !batch) {                                              // compression.cpp:587-590
        std::cerr << "Usage: " << argv[0] << " <reference_file> <target_file> <output_folder>\n"
                  << "       " << argv[0] << " --batch <list of such triples, one per line>\n";
        return 1;
    }
    try {
        const char* dev = getenv("SCCG_DEVICE");
        sccg_ctx* ctx = sccg_create(dev ? atoi(dev) : 0);
        if (!ctx) { std::cerr << "Error: " << sccg_last_error() << "\n"; return 1; }
        int status = 0;
        if (!batch) {
            status = compress_pair(ctx, argv[1], argv[2], argv[3]);
        } else {
            std::ifstream list(argv[2]);
            if (!list.is_open()) { std::cerr << "Error opening list file: " << argv[2] << "\n"; sccg_destroy(ctx); return 1; }
            std::string line;
            while (std::getline(list, line)) {
                std::istringstream is(line);
                std::string r, t, o;
                if (!(is >> r >> t >> o)) continue;                          // blank line
                if (compress_pair(ctx, r, t, o) != 0) status = 1;            // keep going: the pairs are independent
            }
        }
        sccg_destroy(ctx);
        return status;
    } catch (const std::exception& ex) {
        std::cerr << "Error: " << ex.what() << "\n";
        return 1;
    }
}
