// decompress <compressed_file> <reference_file> <output_folder>
// Drop-in for the reference's decompression.cpp main (:281-330): `7z e`, reference FASTA read, then the record
// decode (decompress_genome's in-memory part + reconstruct_genome) on a B200 through libsccg_b200.so.
//
// Additive: `decompress --batch <list> [--gpus N]` runs many archives in one process -- one line per archive,
// "<compressed_file> <reference_file> <output_folder>" -- over N GPUs with two workers per GPU (host/batch.hpp).
// The reconstructed FASTA is STREAMED: the library hands it over piece by piece from a page-locked double buffer
// (sccg_decompress_fasta_stream) and every piece is written to the output file while the next one crosses PCIe.
#include <cstdlib>
#include <filesystem>
#include <iostream>

#include "batch.hpp"
#include "sccg.h"

namespace fs = std::filesystem;
using sccg_host::Job; using sccg_host::Timing; using sccg_host::Worker;

struct FileSink { int fd; bool failed; double write_s; };
static int file_sink(void* user, int64_t /*offset*/, const char* data, int64_t len) {      // pieces arrive in order: a plain sequential write
    FileSink* s = (FileSink*)user;
    const double t0 = sccg_host::now_s();
    if (!sccg_host::write_all(s->fd, data, (size_t)len)) { s->failed = true; return 1; }
    s->write_s += sccg_host::now_s() - t0;
    return 0;
}

static int decompress_one(Worker& w, const Job& job, Timing& tm, bool verbose) {
    const std::string &arc = job.a, &ref_path = job.b, &out_dir = job.out;
    const double t0 = sccg_host::now_s();
    std::error_code ec;
    if (!fs::exists(out_dir)) fs::create_directory(out_dir, ec);
    const std::string cmd = "7z e \"" + arc + "\" -o\"" + out_dir + "\" -y" + (verbose ? "" : " > /dev/null");   // :34
    if (system(cmd.c_str()) != 0) { std::cerr << "Greska pri dekompresiji: " << arc << "\n"; return 1; }
    const std::string inter_path = out_dir + "/" + fs::path(arc).stem().string();   // :43-44
    const double t1 = sccg_host::now_s();

    // the raw FASTA image goes to the GPU as it is: header lines and whitespace are removed there (decompression.cpp:47-58)
    int64_t nf = 0, ni = 0;
    if (!sccg_host::read_file_pinned(ref_path, w.in0, &nf)) { std::cerr << "Greska pri otvaranju reference: " << ref_path << "\n"; return 1; }
    if (!sccg_host::read_file_pinned(inter_path, w.in1, &ni)) { std::cerr << "Greska pri otvaranju datoteke: " << inter_path << "\n"; return 1; }
    const double t2 = sccg_host::now_s();

    // the reference opens the output only after reconstruct_genome succeeded (:316); here the file is created up front and
    // removed again if the record stream turns out to be malformed (reported before the first piece is delivered)
    const std::string out_path = out_dir + "/reconstructed_genome.fa";
    const std::string tmp_path = out_path + ".part";
    FileSink sink{open(tmp_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644), false, 0.0};
    if (sink.fd < 0) { std::cerr << "Error opening output file: " << out_dir << "\n"; return 1; }
    int64_t total = 0;
    const int rc = sccg_decompress_fasta_stream(w.ctx, w.in0.p, nf, w.in1.p, ni, file_sink, &sink, &total);
    close(sink.fd);
    if (rc != SCCG_OK) {
        unlink(tmp_path.c_str());
        if (sink.failed) { std::cerr << "Error opening output file: " << out_dir << "\n"; return 1; }
        // SCCG_E_BOUNDS: the reference prints the same ERROR and exit(1)s (:223-229); SCCG_E_FORMAT: it dies in stoi (:309-312)
        std::cerr << (rc == SCCG_E_BOUNDS ? "" : "Error during reconstruction: ") << sccg_last_error() << "\n";
        return 1;
    }
    if (rename(tmp_path.c_str(), out_path.c_str()) != 0) { std::cerr << "Error opening output file: " << out_dir << "\n"; return 1; }
    const double t3 = sccg_host::now_s();
    tm.ext_s = t1 - t0; tm.read_s = t2 - t1; tm.write_s = sink.write_s; tm.gpu_s = (t3 - t2) - sink.write_s;
    if (verbose) {
        if (getenv("SCCG_TIMING")) std::cerr << "timing: 7z " << tm.ext_s << " s, read " << tm.read_s << " s, gpu call " << tm.gpu_s << " s, write (overlapped with PCIe) " << tm.write_s << " s\n";
        std::cout << "Time taken to decompress: " << (t3 - t0) << " s\n";      // :327 (here the write is inside: it overlaps the transfer)
    }
    return 0;
}

int main(int argc, char* argv[]) {
    const bool batch = argc >= 3 && std::string(argv[1]) == "--batch";
    if (argc != 4 && !batch) {                                              // decompression.cpp:283-286
        std::cerr << "Usage: " << argv[0] << " <compressed_file> <reference_file> <output_folder>\n"
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
            const int status = decompress_one(w, job, tm, true);
            sccg_destroy(w.ctx);
            return status;
        }
        std::vector<Job> jobs;
        if (!sccg_host::read_job_list(argv[2], &jobs)) { std::cerr << "Error opening list file: " << argv[2] << "\n"; return 1; }
        for (Job& j : jobs) j.weight = sccg_host::file_size(j.b);
        return sccg_host::run_batch(jobs, sccg_host::parse_gpus(argc, argv, 3), [](Worker& w, const Job& j, Timing& tm) { return decompress_one(w, j, tm, false); }, std::cout);
    } catch (const std::exception& ex) {
        std::cerr << "Error: " << ex.what() << "\n";
        return 1;
    }
}
