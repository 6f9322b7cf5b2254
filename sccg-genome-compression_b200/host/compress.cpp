// compress <reference_file> <target_file> <output_folder>
// Drop-in for the reference's compression.cpp main (:584-610) with the hot path (match-and-encode) on a B200 through
// libsccg_b200.so.  Same argv, same output files (<out>/compressed_genome.txt, then the external `7z a -mx=9`), same
// exit codes.  There is no CPU matcher: without a usable GPU the program fails.
//
// Additive: `compress --batch <list> [--gpus N]` runs many pairs in one process -- one line per pair,
// "<reference_file> <target_file> <output_folder>" -- over N GPUs with two workers per GPU (host/batch.hpp): the CUDA
// context is paid once and file I/O, PCIe and kernels of neighbouring pairs overlap.
#include <cstdlib>
#include <filesystem>
#include <iostream>

#include "batch.hpp"
#include "sccg.h"

using sccg_host::Job; using sccg_host::Timing; using sccg_host::Worker;

// one pair: compress_genome (:320-582) with the reference's files and messages.  0 = ok, 1 = what the reference exits with.
static int compress_pair(Worker& w, const Job& job, Timing& tm, bool verbose) {
    const std::string &ref_path = job.a, &tgt_path = job.b, &out_dir = job.out;
    std::error_code ec;
    if (!std::filesystem::exists(out_dir)) std::filesystem::create_directory(out_dir, ec);
    const double t0 = sccg_host::now_s();
    // the raw FASTA images go to the GPU as they are (page-locked buffers): read_genomes_from_files (:181-220) runs there
    int64_t nr = 0, nt = 0;
    if (!sccg_host::read_file_pinned(ref_path, w.in0, &nr)) { std::cerr << "Error opening reference file: " << ref_path << "\n"; return 1; }
    if (!sccg_host::read_file_pinned(tgt_path, w.in1, &nt)) { std::cerr << "Error opening target file: " << tgt_path << "\n"; return 1; }
    const double t1 = sccg_host::now_s();
    int64_t out_len = 0; int mode = 0; int rc = SCCG_OK;
    if (!w.outb.ensure((size_t)nt / 4 + (1u << 20))) { std::cerr << "Error: out of page-locked memory\n"; return 1; }
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = sccg_compress_fasta_into(w.ctx, w.in0.p, nr, w.in1.p, nt, w.outb.p, (int64_t)w.outb.cap, &out_len, &mode);
        if (rc == SCCG_E_ARG && out_len > (int64_t)w.outb.cap && w.outb.ensure((size_t)out_len + 64)) continue;       // a literal-heavy body: grow once
        break;
    }
    if (rc != SCCG_OK && rc != SCCG_E_STOI) { std::cerr << "Error: " << sccg_last_error() << "\n"; return 1; }
    sccg_profile prof; sccg_get_profile(w.ctx, &prof);
    const double t2 = sccg_host::now_s();

    std::filesystem::create_directories(out_dir, ec);                        // :334
    const std::string txt = out_dir + "/compressed_genome.txt";
    int fd = open(txt.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0 || !sccg_host::write_all(fd, w.outb.p, (size_t)out_len)) { std::cerr << "Greska pri otvaranju datoteke: " << txt << "\n"; if (fd >= 0) close(fd); return 1; }
    close(fd);
    const double t3 = sccg_host::now_s();
    if (rc == SCCG_E_STOI) { std::cerr << "Error: stoi\n"; return 1; }         // delta_encode threw (:279 -> :604-607): file left un-rewritten, no 7z
    if (verbose)
        std::cout << "mode: " << (mode ? "global" : "local") << ", GPU kernels " << prof.kernels_ms << " ms, H2D " << prof.h2d_ms << " ms, D2H " << prof.d2h_ms << " ms\n";
    const std::string cmd = "7z a -mx=9 \"" + txt + ".7z\" \"" + txt + "\"" + (verbose ? "" : " > /dev/null");   // :308, the external stage stays as it is
    if (system(cmd.c_str()) != 0) { std::cerr << "Greska prilikom komprimiranja datoteke 7-zipom !\n"; return 1; }
    const double t4 = sccg_host::now_s();
    tm.read_s = t1 - t0; tm.gpu_s = t2 - t1; tm.write_s = t3 - t2; tm.ext_s = t4 - t3;
    if (verbose) {
        if (getenv("SCCG_TIMING")) std::cerr << "timing: read " << tm.read_s << " s, gpu call " << tm.gpu_s << " s, write " << tm.write_s << " s, 7z " << tm.ext_s << " s\n";
        std::cout << "Time taken to compress: " << (t4 - t0) << " s\n";        // :602
    }
    return 0;
}

int main(int argc, char* argv[]) {
    const bool batch = argc >= 3 && std::string(argv[1]) == "--batch";
    if (argc != 4 && !batch) {                                              // compression.cpp:587-590
        std::cerr << "Usage: " << argv[0] << " <reference_file> <target_file> <output_folder>\n"
                  << "       " << argv[0] << " --batch <list of such triples, one per line> [--gpus N]\n";
        return 1;
    }
    try {
        if (!batch) {
            const double t0 = sccg_host::now_s();
            const char* dev = getenv("SCCG_DEVICE");
            Worker w; w.device = dev ? atoi(dev) : 0;
            w.ctx = sccg_create(w.device);
            if (!w.ctx) { std::cerr << "Error: " << sccg_last_error() << "\n"; return 1; }
            if (getenv("SCCG_TIMING")) std::cerr << "timing: context " << (sccg_host::now_s() - t0) << " s\n";
            Job job; job.a = argv[1]; job.b = argv[2]; job.out = argv[3];
            Timing tm;
            const int status = compress_pair(w, job, tm, true);
            sccg_destroy(w.ctx);
            return status;
        }
        std::vector<Job> jobs;
        if (!sccg_host::read_job_list(argv[2], &jobs)) { std::cerr << "Error opening list file: " << argv[2] << "\n"; return 1; }
        for (Job& j : jobs) j.weight = sccg_host::file_size(j.b);
        return sccg_host::run_batch(jobs, sccg_host::parse_gpus(argc, argv, 3), [](Worker& w, const Job& j, Timing& tm) { return compress_pair(w, j, tm, false); }, std::cout);
    } catch (const std::exception& ex) {
        std::cerr << "Error: " << ex.what() << "\n";
        return 1;
    }
}
