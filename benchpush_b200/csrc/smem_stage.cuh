// Bulk staging between the env blocks in HBM and shared memory: cp.async.bulk (the TMA engine's linear copy, UBLKCP in SASS)
// completing on an mbarrier for loads, bulk-group stores for write-back.  Sizes are multiples of 16 bytes, both addresses
// 16-byte aligned.  The host build (BP_HOST_EMU, tests only) degrades to memcpy.
#pragma once
#include "warp_prims.cuh"
#include <string.h>

namespace bp {

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
BP_DEV unsigned bp_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
BP_DEV void bp_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bp_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arrive on the barrier and announce `bytes` of bulk-copy traffic issued by this thread
BP_DEV void bp_mbar_arrive_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bp_smem_addr(bar)), "r"(bytes) : "memory");
}
BP_DEV void bp_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "BP_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra BP_MBAR_DONE;\n"
        "bra BP_MBAR_WAIT;\n"
        "BP_MBAR_DONE:\n"
        "}\n" ::"r"(bp_smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted on `bar`
BP_DEV void bp_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bp_smem_addr(dst_smem)),
                 "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(bp_smem_addr(bar))
                 : "memory");
}
// shared -> global, part of the calling thread's current bulk group
BP_DEV void bp_bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst_gmem)), "r"(bp_smem_addr(src_smem)),
                 "r"(bytes)
                 : "memory");
}
BP_DEV void bp_bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// order generic-proxy writes (plain stores) before async-proxy accesses (bulk copies) of the same memory
BP_DEV void bp_fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
BP_DEV void bp_mbar_init(unsigned long long* bar, int count) { (void)bar; (void)count; }
BP_DEV void bp_mbar_arrive_tx(unsigned long long* bar, unsigned bytes) { (void)bar; (void)bytes; }
BP_DEV void bp_mbar_wait(unsigned long long* bar, unsigned parity) { (void)bar; (void)parity; }
BP_DEV void bp_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) { (void)bar; memcpy(dst, src, bytes); }
BP_DEV void bp_bulk_s2g(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
BP_DEV void bp_bulk_commit_wait() {}
BP_DEV void bp_fence_async() {}
#endif

}  // namespace bp
