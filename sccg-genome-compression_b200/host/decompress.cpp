// decompress <compressed_file> <reference_file> <output_folder>
// Drop-in for the reference's decompression.cpp main (:281-330): `7z e`, reference FASTA read, then the record
// decode (decompress_genome's in-memory part + reconstruct_genome) on a B200 through libsccg_b200.so.
//
// Additive: `decompress --batch <list>` runs many archives in one process -- one line per archive,
// "<compressed_file> <reference_file> <output_folder>" -- so the CUDA context is paid once.
#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>

#include "fasta_io.hpp"
#include "sccg.h"

namespace fs = std::filesystem;

static int decompress_one(sccg_ctx* ctx, const std::string& arc, const std::string& ref_path, const std::string& out_dir) {
    auto t0 = std::chrono::high_resolution_clock::now();
    if (!fs::exists(out_dir)) fs::create_directory(out_dir);
    const std::string cmd = "7z e \"" + arc + "\" -o\"" + out_dir + "\" -y";   // :34
    if (system(cmd.c_str()) != 0) { std::cerr << "Greska pri dekompresiji: " << arc << "\n"; return 1; }
    const std::string inter_path = out_dir + "/" + fs::path(arc).stem().string();   // :43-44

    // the raw FASTA image goes to the GPU as it is: header lines and whitespace are removed there (decompression.cpp:47-58)
    std::string file, inter;
    if (!sccg_host::read_file(ref_path, file)) { std::cerr << "Greska pri otvaranju reference: " << ref_path << "\n"; return 1; }
    if (!sccg_host::read_file(inter_path, inter)) { std::cerr << "Greska pri otvaranju datoteke: " << inter_path << "\n"; return 1; }

    char* out = nullptr; int64_t out_len = 0;
    int rc = sccg_decompress_fasta(ctx, file.data(), (int64_t)file.size(), inter.data(), (int64_t)inter.size(), &out, &out_len);
    if (rc != SCCG_OK) {
        // SCCG_E_BOUNDS: the reference prints the same ERROR and exit(1)s (:223-229); SCCG_E_FORMAT: it dies in stoi (:309-312)
        std::cerr << (rc == SCCG_E_BOUNDS ? "" : "Error during reconstruction: ") << sccg_last_error() << "\n";
        return 1;
    }
    auto t1 = std::chrono::high_resolution_clock::now();                       // the reference stops its clock before the write (:313)
    const std::string out_path = out_dir + "/reconstructed_genome.fa";
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f || fwrite(out, 1, (size_t)out_len, f) != (size_t)out_len) { std::cerr << "Error opening output file: " << out_dir << "\n"; if (f) fclose(f); sccg_free(out); return 1; }
    fclose(f);
    sccg_free(out);
    std::chrono::duration<double> dt = t1 - t0;
    std::cout << "Time taken to decompress: " << dt.count() << " s\n";         // :327
    return 0;
}

int main(int argc, char* argv[]) {
    const bool batch = argc == 3 && std::string(argv[1]) == "--batch";
    if (argc != 4 && !batch) {                                              // decompression.cpp:283-286
        std::cerr << "Usage: " << argv[0] << " <compressed_file> <reference_file> <output_folder>\n"
                  << "       " << argv[0] << " --batch <list of such triples, one per line>\n";
        return 1;
    }
    try {
        const char* dev = getenv("SCCG_DEVICE");
        sccg_ctx* ctx = sccg_create(dev ? atoi(dev) : 0);
        if (!ctx) { std::cerr << "Error: " << sccg_last_error() << "\n"; return 1; }
        int status = 0;
        if (!batch) {
            status = decompress_one(ctx, argv[1], argv[2], argv[3]);
        } else {
            std::ifstream list(argv[2]);
            if (!list.is_open()) { std::cerr << "Error opening list file: " << argv[2] << "\n"; sccg_destroy(ctx); return 1; }
            std::string line;
            while (std::getline(list, line)) {
                std::istringstream is(line);
                std::string a, r, o;
                if (!(is >> a >> r >> o)) continue;
                if (decompress_one(ctx, a, r, o) != 0) status = 1;
            }
        }
        sccg_destroy(ctx);
        return status;
    } catch (const std::exception& ex) {
        std::cerr << "Error: " << ex.what() << "\n";
        return 1;
    }
}
