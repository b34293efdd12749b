// Device-side building blocks for the SOT kernels (sm_100a).
//
// Nothing here is a port: the reference (`losses.py:129-313`) is a chain of ~30 ATen launches
// (sort / gather / cumsum / cat+sort / searchsorted / take_along_dim / elementwise); these
// helpers implement the same mathematics as ONE pass per frame that never leaves the SM:
// TMA bulk copy -> registers -> fp64 block scan -> merge-path partition -> sequential walk.
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

// ------------------------------------------------------------------------------------------
// frame-group synchronisation: a frame is processed by TPF threads (1, 2, 4 or 8 warps);
// FPC such groups share a CTA.  Named barrier ids 1..FPC (0 is __syncthreads).
// ------------------------------------------------------------------------------------------
template <int TPF>
SOT_DEVINL void group_sync(int g) {
    if constexpr (TPF == 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(TPF) : "memory");
    }
}

// Exclusive scan (prefix, or suffix when REVERSE) of two fp64 values across the TPF threads of
// a frame group, plus both group totals.  `slot` points at 2*(TPF/32) doubles of scratch that
// no other collective of this frame reuses (so no trailing barrier is needed).
template <int TPF, bool REVERSE>
SOT_DEVINL void group_scan2(double& a, double& b, double& total_a, double& total_b, double* slot, int tid,
                            int g) {
    constexpr int NW = TPF / 32;
    const int lane = tid & 31, w = tid >> 5;
    double ia = a, ib = b;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double ya = REVERSE ? __shfl_down_sync(FULL_MASK, ia, off) : __shfl_up_sync(FULL_MASK, ia, off);
        const double yb = REVERSE ? __shfl_down_sync(FULL_MASK, ib, off) : __shfl_up_sync(FULL_MASK, ib, off);
        const bool ok = REVERSE ? (lane + off < 32) : (lane >= off);
        if (ok) {
            ia += ya;
            ib += yb;
        }
    }
    // exclusive value = neighbour's inclusive value (exact, no subtraction)
    double ea = REVERSE ? __shfl_down_sync(FULL_MASK, ia, 1) : __shfl_up_sync(FULL_MASK, ia, 1);
    double eb = REVERSE ? __shfl_down_sync(FULL_MASK, ib, 1) : __shfl_up_sync(FULL_MASK, ib, 1);
    const int edge = REVERSE ? 31 : 0;
    if (lane == edge) {
        ea = 0.0;
        eb = 0.0;
    }
    const double wa = __shfl_sync(FULL_MASK, ia, REVERSE ? 0 : 31);  // warp totals
    const double wb = __shfl_sync(FULL_MASK, ib, REVERSE ? 0 : 31);
    if constexpr (NW == 1) {
        a = ea;
        b = eb;
        total_a = wa;
        total_b = wb;
    } else {
        if (lane == 0) {
            slot[2 * w] = wa;
            slot[2 * w + 1] = wb;
        }
        group_sync<TPF>(g);
        double pa = 0.0, pb = 0.0, ta = 0.0, tb = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) {  // left-to-right so prefix_w == prefix_{w-1} + total_{w-1}
            const int kk = REVERSE ? (NW - 1 - k) : k;
            const bool before = REVERSE ? (kk > w) : (kk < w);
            const double sa = slot[2 * kk], sb = slot[2 * kk + 1];
            if (before) {
                pa = ta + sa;
                pb = tb + sb;
            }
            ta += sa;
            tb += sb;
        }
        a = pa + ea;
        b = pb + eb;
        total_a = ta;
        total_b = tb;
    }
}

// Sum of two fp64 values over the frame group (all threads get the totals).
template <int TPF>
SOT_DEVINL void group_sum2(double& a, double& b, double* slot, int tid, int g) {
    constexpr int NW = TPF / 32;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(FULL_MASK, a, off);
        b += __shfl_xor_sync(FULL_MASK, b, off);
    }
    if constexpr (NW > 1) {
        const int lane = tid & 31, w = tid >> 5;
        if (lane == 0) {
            slot[2 * w] = a;
            slot[2 * w + 1] = b;
        }
        group_sync<TPF>(g);
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            ta += slot[2 * k];
            tb += slot[2 * k + 1];
        }
        a = ta;
        b = tb;
    }
}

// ------------------------------------------------------------------------------------------
// merge path: how many entries of A precede output slot k of the stable merge of the sorted
// rows A (n entries) and B (m entries), A first on equal values.  Replaces the reference's
// `sort(cat(cu, cv))` + 2x `searchsorted` (losses.py:295, 219).
// ------------------------------------------------------------------------------------------
SOT_DEVINL int merge_path(const float* __restrict__ A, const float* __restrict__ B, int n, int m, int k) {
    int lo = max(0, k - m), hi = min(k, n);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] <= B[k - 1 - mid])
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// |d|^p; PMODE 2 -> d*d compiled in (losses.py:313 with p=2); PMODE 0 -> any p at run time:
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
    if constexpr (PMODE == 2) {
        return __fmul_rn(d, d);
    } else {
        if (p == 2.0f) return __fmul_rn(d, d);
        const float ad = fabsf(d);
        return (p == 1.0f) ? ad : powf(ad, p);
    }
}

}  // namespace sot
