// Bulk asynchronous copies (the TMA engine's non-tensor form, sm_90+ / sm_100a): cp.async.bulk between global and shared
// memory, completion through an mbarrier (loads) or a bulk group (stores).  One elected thread issues a copy of any size
// (16-byte aligned on both sides, a multiple of 16 bytes); the copy engine moves the bytes, no thread spends issue slots on it.
// SASS: UBLKCP (copies), SYNCS (mbarrier).  The emulator build (tests) performs the copies at once.
#pragma once
#include "sccg_common.cuh"

namespace sccg {

#ifndef SCCG_EMU
__device__ __forceinline__ u32 smem_addr32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");          // visible to the async proxy before the first copy names it
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_addr32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_addr32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    u32 done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared, bytes % 16 == 0, both addresses 16-byte aligned; completes `bytes` of transaction count on bar
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr32(bar)) : "memory");
}
// shared -> global as one bulk group of the calling thread; the shared source must stay untouched until bulk_wait_read()
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst_gmem), "r"(smem_addr32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (a following bulk store)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ u32 smem_phase16(const void* p) { return smem_addr32(p) & 15u; }
#else
__device__ __forceinline__ void mbar_init(u64*, u32) {}
__device__ __forceinline__ void mbar_expect_tx(u64*, u32) {}
__device__ __forceinline__ void mbar_arrive(u64*) {}
__device__ __forceinline__ void mbar_wait(u64*, u32) {}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64*) { memcpy(dst_smem, src_gmem, bytes); }
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, u32 bytes) { memcpy(dst_gmem, src_smem, bytes); }
__device__ __forceinline__ void bulk_wait_read() {}
__device__ __forceinline__ void fence_smem_to_async() {}
__device__ __forceinline__ u32 smem_phase16(const void* p) { return (u32)((uintptr_t)p & 15u); }
#endif

}  // namespace sccg
