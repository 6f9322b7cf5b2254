// Host-side FASTA handling with the reference's exact semantics:
//   reference file: every empty line and every line starting with '>' is skipped, the rest is
//                   concatenated, then every isspace() byte is removed   (compression.cpp:193-200,
//                   decompression.cpp:53-58)
//   target file   : empty lines skipped; only the FIRST '>' line becomes the header (kept verbatim,
//                   including a trailing '\r'); later '>' lines stay in the sequence (compression.cpp:207-218)
#pragma once
#include <cstdio>
#include <cstring>
#include <string>

namespace sccg_host {

inline bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    out.resize(got);
    return true;
}

inline bool c_isspace(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }

// appends [b, e) without whitespace
inline void append_stripped(std::string& seq, const char* b, const char* e) {
    size_t old = seq.size();
    seq.resize(old + (size_t)(e - b));
    char* w = &seq[old];
    for (const char* p = b; p < e; ++p) if (!c_isspace((unsigned char)*p)) *w++ = *p;
    seq.resize((size_t)(w - seq.data()));
}

inline void parse_fasta(const std::string& file, bool is_target, std::string& seq, std::string* header) {
    seq.clear();
    seq.reserve(file.size());
    bool header_found = false;
    const char* p = file.data();
    const char* end = p + file.size();
    while (p < end) {                                         // std::getline semantics
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        if (le > p) {                                         // line.empty() -> skipped
            if (*p == '>' && (!is_target || !header_found)) {
                if (is_target) { header->assign(p, le); header_found = true; }
            } else {
                append_stripped(seq, p, le);
            }
        }
        p = nl ? nl + 1 : end;
    }
}

}  // namespace sccg_host
