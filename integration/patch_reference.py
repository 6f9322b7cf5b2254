#!/usr/bin/env python
"""Applies the reference-side change of INTEGRATION.md section 2 to a checkout of Jan-Celin/SCCG-genome-compression:
compress_genome's middle (compression.cpp:336-579: lowercase RLE, segment loop, global fallback, delta_encode) becomes one
sccg_compress call, reconstruct_genome's body (decompression.cpp:122-278) one sccg_reconstruct call.  Everything else of the
two programs -- main, argv, FASTA readers, the 7z stage, messages -- stays the reference's own code.

    python integration/patch_reference.py <reference checkout> <output directory>

The patched sources are written to the output directory (never into the checkout, never into this repository); build them with
    g++ -O3 -std=c++17 <out>/compression.cpp -Iinclude -Lsccg-genome-compression_b200 -lsccg_b200 -Wl,-rpath,<lib dir> -o compress
The edit is located by anchor lines, not by line numbers, and fails loudly if an anchor is missing."""
import re
import sys
from pathlib import Path

BINDING = 'extern "C" {\n#include "sccg.h"\n}\n'

COMPRESS_BODY = '''    // ---- libsccg_b200: compress_genome minus file I/O and 7z (replaces compression.cpp:336-579)
    static sccg_ctx* sccg_context = sccg_create(getenv("SCCG_DEVICE") ? atoi(getenv("SCCG_DEVICE")) : 0);
    if (!sccg_context) { cerr << "Error: " << sccg_last_error() << "\\n"; exit(1); }
    char* sccg_out = nullptr; int64_t sccg_out_len = 0; int sccg_mode = 0;
    int sccg_rc = sccg_compress(sccg_context, reference_genome.data(), (int64_t)reference_genome.size(),
                                target_genome.data(), (int64_t)target_genome.size(),
                                target_header.data(), (int64_t)target_header.size(), &sccg_out, &sccg_out_len, &sccg_mode);
    if (sccg_rc != SCCG_OK && sccg_rc != SCCG_E_STOI) { cerr << "Error: " << sccg_last_error() << "\\n"; exit(1); }
    {
        ofstream sccg_file(temp_file_path, ofstream::out | ofstream::trunc | ofstream::binary);
        sccg_file.write(sccg_out, sccg_out_len);                 // the final, delta-encoded compressed_genome.txt
    }
    sccg_free(sccg_out);
    if (sccg_rc == SCCG_E_STOI) throw invalid_argument("stoi");   // what delta_encode does on a literal '(' (:279): the un-rewritten file stays behind
'''

RECONSTRUCT_BODY = '''    // ---- libsccg_b200: reconstruct_genome (replaces decompression.cpp:122-278)
    static sccg_ctx* sccg_context = sccg_create(getenv("SCCG_DEVICE") ? atoi(getenv("SCCG_DEVICE")) : 0);
    if (!sccg_context) { cerr << "Error: " << sccg_last_error() << "\\n"; exit(1); }
    // decompress_genome hands the N line through as it is (:100): "," in local mode, which the original parser reads as "no runs"
    const string sccg_n = n_indices_str == "," ? string() : n_indices_str;
    char* sccg_out = nullptr; int64_t sccg_n_out = 0;
    int sccg_rc = sccg_reconstruct(sccg_context, reference_genome.data(), (int64_t)reference_genome.size(),
                                   encoded_genome.data(), (int64_t)encoded_genome.size(), sccg_n.data(), (int64_t)sccg_n.size(),
                                   lowercase_indices_str.data(), (int64_t)lowercase_indices_str.size(), &sccg_out, &sccg_n_out);
    if (sccg_rc == SCCG_E_BOUNDS) { cerr << sccg_last_error() << "\\n"; exit(1); }          // :223-229
    if (sccg_rc != SCCG_OK) throw runtime_error(sccg_last_error());                          // caught in main (:309-312)
    string result(sccg_out, (size_t)sccg_n_out);
    sccg_free(sccg_out);
    return result;
'''


def find(lines, pattern, start=0):
    rx = re.compile(pattern)
    for i in range(start, len(lines)):
        if rx.search(lines[i]):
            return i
    raise SystemExit(f"anchor not found: {pattern!r} (is this the expected reference checkout?)")


def patch_compression(text: str) -> str:
    lines = text.split("\n")
    u = find(lines, r"^using namespace std;")
    lines.insert(u + 1, BINDING.rstrip("\n"))
    f = find(lines, r"^void compress_genome\(")
    a = find(lines, r"ofstream temp_file\(temp_file_path", f)            # first line that goes: the file is written by the library's image
    b = find(lines, r"delta_encode\(temp_file_path\);", a)                # last line that goes
    z = find(lines, r"compress_genome_7z\(temp_file_path, output_file_path\);", b)
    assert b < z <= b + 3
    return "\n".join(lines[:a] + COMPRESS_BODY.rstrip("\n").split("\n") + lines[b + 1:])


def patch_decompression(text: str) -> str:
    lines = text.split("\n")
    u = find(lines, r"^using namespace std;")
    lines.insert(u + 1, BINDING.rstrip("\n"))
    f = find(lines, r"^string reconstruct_genome\(")
    a = find(lines, r"lowercase_indices_str\) \{", f)                       # end of the signature
    b = find(lines, r"^\s*return result;", a)                                # last statement of the body
    return "\n".join(lines[:a + 1] + RECONSTRUCT_BODY.rstrip("\n").split("\n") + lines[b + 1:])


def main():
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    ref, out = Path(sys.argv[1]), Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    (out / "compression.cpp").write_text(patch_compression((ref / "compression.cpp").read_text()))
    (out / "decompression.cpp").write_text(patch_decompression((ref / "decompression.cpp").read_text()))
    print(f"patched sources written to {out}")


if __name__ == "__main__":
    main()
