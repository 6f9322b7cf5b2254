// test helper: runs the host FASTA reader of the CLIs on a file and dumps "<header>\n<sequence>"
#include <iostream>
#include "../sccg-genome-compression_b200/host/fasta_io.hpp"
int main(int argc, char** argv) {
    if (argc != 3) return 2;
    std::string file, seq, header;
    if (!sccg_host::read_file(argv[2], file)) return 1;
    sccg_host::parse_fasta(file, std::string(argv[1]) == "target", seq, &header);
    std::cout << header << "\n" << seq;
    return 0;
}
