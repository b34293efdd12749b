// Device-side building blocks for the SOT kernels (sm_100a).
//
// Nothing here is a port: the reference (`losses.py:129-313`) is a chain of ~30 ATen launches
// (sort / gather / cumsum / cat+sort / searchsorted / take_along_dim / elementwise); these
// helpers implement the same mathematics as ONE pass per frame that never leaves the SM:
// TMA bulk copy -> registers -> packed-fp32 prefix sums with fp64 offsets -> merge-path partition ->
// sequential walk.  (PTX wrappers and the transport cost live here; the kernel is in sot_kernels.cuh.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SOT_DEVINL __device__ __forceinline__

namespace sot {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr float SAFE_EPS = 1e-7f;          // utils.py:135 (safe_divide eps)
constexpr float FLT_BIG = 3.402823466e+38f;  // FLT_MAX: data clamp so the +inf sentinel stays unique

// ------------------------------------------------------------------------------------------
// TMA bulk copy (1-D, `cp.async.bulk`, SASS: UBLKCP) + mbarrier plumbing
// ------------------------------------------------------------------------------------------
SOT_DEVINL uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
SOT_DEVINL void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SOT_DEVINL void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SOT_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
SOT_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on `bar`; all of dst, src, bytes 16-byte aligned
SOT_DEVINL void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global
SOT_DEVINL void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                 "r"(bytes)
                 : "memory");
}
SOT_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SOT_DEVINL void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async (TMA) proxy
SOT_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// |d|^p; PMODE 2 -> d*d compiled in (losses.py:313 with p=2); PMODE 3 -> the same, in a kernel that also has NO cutoff
// mask compiled into its merge walk (limit_quantile_range off: the BASELINE "NoCut" sweep); PMODE 0 -> any p at run time:
// d*d for 2, |d| for 1 (:311-312), powf(|d|, p) otherwise (:313)
template <int PMODE>
SOT_DEVINL float cost_of_gap(float d, float p);

template <int PMODE>
SOT_DEVINL float transport_cost(float pa, float pb, float p) {
    return cost_of_gap<PMODE>(pa - pb, p);
}

template <int PMODE>
SOT_DEVINL float cost_of_gap(float d, float p) {
    // __fmul_rn: never contracted into an FMA with a later subtraction, so |d|^2 is rounded
    // once on its own exactly like the reference's `diff.pow(2)` tensor (losses.py:313)
    if constexpr (PMODE == 2 || PMODE == 3) {
        return __fmul_rn(d, d);
    } else {
        if (p == 2.0f) return __fmul_rn(d, d);
        const float ad = fabsf(d);
        return (p == 1.0f) ? ad : powf(ad, p);
    }
}

}  // namespace sot
