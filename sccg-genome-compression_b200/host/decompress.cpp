// decompress <compressed_file> <reference_file> <output_folder>
// Drop-in for the reference's decompression.cpp main (:281-330): `7z e`, reference FASTA read, then the record
// decode (decompress_genome's in-memory part + reconstruct_genome) on a B200 through libsccg_b200.so.
#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <iostream>

#include "fasta_io.hpp"
#include "sccg.h"

namespace fs = std::filesystem;

int main(int argc, char* argv[]) {
    if (argc != 4) {                                                        // decompression.cpp:283-286
        std::cerr << "Usage: " << argv[0] << " <compressed_file> <reference_file> <output_folder>\n";
        return 1;
    }
    auto t0 = std::chrono::high_resolution_clock::now();
    try {
        const std::string arc = argv[1], ref_path = argv[2], out_dir = argv[3];
        if (!fs::exists(out_dir)) fs::create_directory(out_dir);
        const std::string cmd = "7z e \"" + arc + "\" -o\"" + out_dir + "\" -y";   // :34
        if (system(cmd.c_str()) != 0) { std::cerr << "Greska pri dekompresiji: " << arc << "\n"; exit(1); }
        const std::string inter_path = out_dir + "/" + fs::path(arc).stem().string();   // :43-44

        // the raw FASTA image goes to the GPU as it is: header lines and whitespace are removed there (decompression.cpp:47-58)
        std::string file, inter;
        if (!sccg_host::read_file(ref_path, file)) { std::cerr << "Greska pri otvaranju reference: " << ref_path << "\n"; exit(1); }
        if (!sccg_host::read_file(inter_path, inter)) { std::cerr << "Greska pri otvaranju datoteke: " << inter_path << "\n"; exit(1); }

        const char* dev = getenv("SCCG_DEVICE");
        sccg_ctx* ctx = sccg_create(dev ? atoi(dev) : 0);
        if (!ctx) { std::cerr << "Error: " << sccg_last_error() << "\n"; return 1; }
        char* out = nullptr; int64_t out_len = 0;
        int rc = sccg_decompress_fasta(ctx, file.data(), (int64_t)file.size(), inter.data(), (int64_t)inter.size(), &out, &out_len);
        if (rc != SCCG_OK) {
            // SCCG_E_BOUNDS: the reference prints the same ERROR and exit(1)s (:223-229); SCCG_E_FORMAT: it dies in stoi (:309-312)
            std::cerr << (rc == SCCG_E_BOUNDS ? "" : "Error during reconstruction: ") << sccg_last_error() << "\n";
            sccg_destroy(ctx);
            return 1;
        }
        auto t1 = std::chrono::high_resolution_clock::now();                   // the reference stops its clock before the write (:313)
        const std::string out_path = out_dir + "/reconstructed_genome.fa";
        FILE* f = fopen(out_path.c_str(), "wb");
        if (!f || fwrite(out, 1, (size_t)out_len, f) != (size_t)out_len) { std::cerr << "Error opening output file: " << out_dir << "\n"; return 1; }
        fclose(f);
        sccg_free(out);
        sccg_destroy(ctx);
        std::chrono::duration<double> dt = t1 - t0;
        std::cout << "Time taken to decompress: " << dt.count() << " s\n";     // :327
    } catch (const std::exception& ex) {
        std::cerr << "Error: " << ex.what() << "\n";
        return 1;
    }
    return 0;
}
