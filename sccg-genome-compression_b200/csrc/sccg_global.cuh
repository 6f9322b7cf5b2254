#pragma once
#include "sccg_compress.cuh"
namespace sccg {
static int compress_global_device(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, const char* header, i64 nh,
                                  u32 low_k, const u32* cnt_s, const u32* cnt_e, CompressResult* res) {
    return set_error(SCCG_E_ARG, "global mode not implemented yet");
}
static int match_sequences_global(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_tgt, i64 nt, const char* h_tgt, int k, int m, int offset, sccg_records* out) {
    return set_error(SCCG_E_ARG, "global mode not implemented yet");
}
}
