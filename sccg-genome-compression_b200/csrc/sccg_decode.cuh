#pragma once
#include "sccg_compress.cuh"
namespace sccg {
static int reconstruct_device(sccg_ctx* c, const u8* d_ref, i64 nr, const u8* d_enc, i64 ne, const u8* d_n, i64 nn, const u8* d_low, i64 nl,
                              i64 header_reserve, u8** d_out, i64* out_len) {
    return set_error(SCCG_E_ARG, "decode not implemented yet");
}
static int decompress_host(sccg_ctx* c, const char* ref_raw, i64 ref_len, const char* inter, i64 inter_len, char** out, int64_t* out_len) {
    return set_error(SCCG_E_ARG, "decode not implemented yet");
}
}
