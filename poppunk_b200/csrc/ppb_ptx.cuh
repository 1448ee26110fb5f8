// Thin wrappers over the sm_100a PTX this engine needs: mbarrier, 1-D TMA bulk copies, LOP3, named barriers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ppb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make barrier inits visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Waiting.  mbarrier.try_wait may SUSPEND the warp inside the instruction until the phase completes or a time
// limit passes; without an explicit limit the hardware's default is a few tens of cycles, so a waiting warp
// re-issues SYNCS/BRA continuously (measured at N=100k: 22 % of all issued instructions were the producer's and
// the epilogue warps' spin loops, competing with the LOP3 stream for issue slots).  PPB_WAIT_HINT_NS > 0 passes
// that suspend-time hint, so a waiting warp issues nothing until its barrier flips.
#ifndef PPB_WAIT_HINT_NS
#define PPB_WAIT_HINT_NS 1000000  // measured: 736.6 vs 744.6 ms at N=100k (profiles/r01_experiments.md)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#if PPB_WAIT_HINT_NS > 0 && !defined(PPB_WAIT_HINT_RELAXED_ONLY)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PPB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra PPB_DONE_%=;\n\t"
        "bra PPB_WAIT_%=;\n\t"
        "PPB_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((uint32_t)PPB_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PPB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PPB_DONE_%=;\n\t"
        "bra PPB_WAIT_%=;\n\t"
        "PPB_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#endif
}

// Same, for warps that are NOT on the critical path (TMA producer, epilogue).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, uint32_t sleep_ns) {
#if PPB_WAIT_HINT_NS > 0
    (void)sleep_ns;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PPB_WAITR_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra PPB_DONER_%=;\n\t"
        "bra PPB_WAITR_%=;\n\t"
        "PPB_DONER_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((uint32_t)PPB_WAIT_HINT_NS)
        : "memory");
#else
    uint32_t done;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(sleep_ns);
    }
#endif
}

// ---- TMA: 1-D bulk global -> shared copy, completion on an mbarrier (SASS: UBLKCP) -----------
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// TMA bulk copy with an L2 eviction-priority policy (createpolicy)
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- L2 eviction-priority policies ------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 0 normal, 1 evict_last (keep), 2 evict_first (stream)
    uint64_t pol;
    if (kind == 1)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg128_hint(const void *ptr, uint64_t policy) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(ptr), "l"(policy));
    return v;
}
__device__ __forceinline__ uint2 ldg64_hint(const void *ptr, uint64_t policy) {
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(ptr), "l"(policy));
    return v;
}

// ---- named barrier among a subset of the CTA's warps -----------------------------------------
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- the one ALU op of the hot loop: bits & ~(a ^ b)  (LOP3.LUT 0x90) -------------------------
__device__ __forceinline__ uint32_t and_xnor(uint32_t bits, uint32_t a, uint32_t b) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x90;" : "=r"(d) : "r"(bits), "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

}  // namespace ppb
