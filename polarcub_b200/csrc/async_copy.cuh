// async_copy.cuh -- bulk asynchronous copies (cp.async.bulk, the non-tensor form of TMA: SASS UBLKCP) completing on an
// mbarrier, for the HBM-streamed stages of the decoders.  One elected thread arms the barrier with the byte count and issues
// the copy; consumers wait on the barrier's phase parity and then read the staged rows from shared memory, so the data in
// flight costs no registers and the prefetch depth is a shared-memory ring, not an unroll factor.
#pragma once
#include "common.cuh"

namespace pc {

// barrier `id` (1..15) among `threads` threads of the CTA (whole warps)
#ifndef PC_EMU
__device__ __forceinline__ void pc_named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
#else
__device__ __forceinline__ void pc_named_barrier(int id, int threads) { emu::named_bar(id, threads); }
#endif

#ifndef PC_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` of transaction on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// waits for the completion of the phase with parity `parity` (the hardware suspends the thread between probes)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// ---- L2 eviction priorities (createpolicy + .L2::cache_hint): the list decoder streams its large levels (written once, read back a
// sub-tree later: they only pass through L2) and keeps re-using its small global levels; an evict_first policy on the former and
// evict_last on the latter keeps the small levels of all resident frames in L2.  A policy value 0 means "no hint".
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 16-byte global store / load with a policy (GLOBAL addresses only)
__device__ __forceinline__ void st_global_hint(double2 *p, const double2 v, const uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ld_global_hint(const double2 *p, const uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, const uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
#else
// CPU emulation (tests/emu): the copy happens at issue; the barrier word holds the parity of the phase in progress
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { *bar += (uint64_t)bytes << 8; }
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    memcpy(dst_smem, src_gmem, bytes);
    *bar -= (uint64_t)bytes << 8;
    if ((*bar >> 8) == 0) {
        *bar ^= 1;  // phase complete
        emu::S().cur->blk->progress = true;
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while ((*bar & 1) == parity) emu::yield();
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() { return 1; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { return 2; }
__device__ __forceinline__ void st_global_hint(double2 *p, const double2 v, const uint64_t) { *p = v; }
__device__ __forceinline__ double2 ld_global_hint(const double2 *p, const uint64_t) { return *p; }
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, const uint64_t) {
    bulk_g2s(dst_smem, src_gmem, bytes, bar);
}
#endif

}  // namespace pc
