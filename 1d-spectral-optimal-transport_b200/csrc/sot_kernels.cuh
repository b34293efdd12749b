// SOT frame kernel (v2): per-frame 1-D Wasserstein (W_p^p) between two spectra, forward and
// fused forward+backward.  Persistent CTAs; a frame is processed by a group of TPF threads, FPC
// frames ("a quad") per CTA iteration.
//
// Reference semantics reproduced (file:line into /root/reference):
//   losses.py:172-184  square, mass, safe_divide (cut mode divides the prediction by the
//                      TARGET's mass), utils.py:135-142
//   losses.py:292-293  inclusive CDFs                     -> fp64 block scan, fp32 storage
//   losses.py:295      sort(cat(cu, cv))                  -> merge-path partition + sequential walk
//   losses.py:214-220  searchsorted-left + clamp + gather -> running co-ranks of the walk
//   losses.py:301-313  sum_k dq_k * |uq_k - vq_k|^p, strict `qs > 1` mask
//   autograd of all of the above          -> dL/dCDF written in place, suffix
//                                                           scans, normalisation chain rule, 2x
//
// Shared memory per CTA (n = bins of u, m = bins of v, S = n + m + 2):
//   LAND  : 4*FPC*(n+m) B   landing zone of the TMA bulk loads: FPC raw rows of u, then of v.
//                           Read once into registers; the NEXT quad's load is issued as soon as
//                           every thread has done so, i.e. it overlaps the whole computation.
//   PAIRS : 8*FPC*S B       per frame S float2 "pairs" (cdf value, support position):
//                           A[0..n] for u then B[0..m] for v, entries [n] / [m] = (+inf, last pos)
//                           (sentinel + the reference's index clamp, losses.py:220).  One LDS.64
//                           per merged slot.  At the end of an iteration of the gradient kernel the
//                           region is reused to stage the output rows.
//   (G)   : 4*FPC*S B       gradient kernel only, ALIASED onto LAND once the raw rows are in registers:
//                           dL/dCDF of every entry, same index space as PAIRS.  (Writing it into the .y
//                           of the consumed pair instead would race: a thread may load its end-of-range
//                           head pair after the neighbour that owns it has already consumed it.)  The
//                           price: the gradient kernel prefetches the next quad only after stage 4.
//   POS   : 4*S B           support positions (once per CTA when shared, else per frame)
//   small : scan scratch, first-slot mailbox + carry per thread, mbarrier
// Thread t of a group owns the E consecutive bins [t*E, t*E+E) of both rows ("blocked"); E is
// odd so every blocked access of a warp is bank-conflict free and no 16-byte alignment of the
// 4*n-byte rows is ever needed (n = 1025 / 257: rows are only 4-byte aligned).
#pragma once
#include "sot_device.cuh"

namespace sot {

enum : int {
    FLAG_SQUARE = 1,     // square_dist          losses.py:172-174
    FLAG_CUT_SCALE = 2,  // dont_normalize       losses.py:180-182
    FLAG_LIMIT = 4,      // limit_quantile_range losses.py:306-307
    FLAG_RAW = 8,        // rows are weights used as given (module-level wasserstein_1d, :223-313)
};

enum : int {
    MODE_SPECTRA = 0,  // inputs are magnitudes: full prologue
    MODE_CDF = 1,      // inputs are CDF rows (parity harness): skip the prologue; with grads the
                       // outputs are dL/dcu, dL/dcv (no suffix scan / chain rule)
};

enum : int {
    OUT_LOSS = 0,  // per-frame loss only (forward)
    OUT_GRAD = 1,  // loss + gradient rows (fused forward+backward)
    OUT_PLAN = 2,  // loss + the transport plan: qs, searchsorted indices, quantiles, CDFs
                   // (`return_quantiles=True`, losses.py:198-201, 299-300; also the parity harness)
};

struct FrameArgs {
    const float* u;  // [n_frames, n]  target spectra (or CDF rows)
    const float* v;  // [n_frames, m]  prediction spectra (or CDF rows)
    const float* pos_u;
    const float* pos_v;
    long long pos_u_stride;  // 0: one shared support row; else elements between per-frame rows
    long long pos_v_stride;
    const float* upstream;  // [n_frames] dL_total/dloss_n, or nullptr (= 1)
    float* loss;            // [n_frames] or nullptr
    float* grad_u;          // [n_frames, n] or nullptr
    float* grad_v;          // [n_frames, m] or nullptr
    // OUT_PLAN outputs, each nullable: [n_frames, n+m] merged grid, un-clamped lower-bound indices
    // and the quantile positions; [n_frames, n] / [n_frames, m] CDF rows
    float* plan_qs;
    int* plan_iu;
    int* plan_iv;
    float* plan_uq;
    float* plan_vq;
    float* plan_cu;
    float* plan_cv;
    long long n_frames;
    int n, m;
    float p;
    int flags;
};

SOT_DEVINL float f_inf() { return __int_as_float(0x7f800000); }
SOT_DEVINL float f_nan() { return __int_as_float(0x7fc00000); }

// Shared-memory carve-up, shared between host (size) and device (offsets), all in bytes.
struct SmemPlan {
    int land, pairs, pos, scratch, mailbox, carry, mbar, total;
};
__host__ __device__ inline SmemPlan smem_plan(int FPC, int n, int m, int tpf, bool pos_shared) {
    SmemPlan s;
    int off = 0;
    s.land = off;
    off += 4 * FPC * (n + m + 2);  // raw rows, later dL/dCDF (n + m + 2 entries per frame)
    off = (off + 15) & ~15;
    s.pairs = off;
    off += 8 * FPC * (n + m + 2);
    s.pos = off;
    off += 4 * (n + m + 2) * (pos_shared ? 1 : FPC);
    off = (off + 15) & ~15;
    s.scratch = off;
    off += 8 * FPC * 16 * (tpf / 32);  // per group: 8 collective slots x warps (+ spare)
    s.mailbox = off;
    off += 8 * FPC * (tpf + 1);  // (first q, first m*d) per thread + one end marker
    s.carry = off;
    off += 4 * FPC * tpf;
    off = (off + 15) & ~15;
    s.mbar = off;
    off += 16;
    s.total = off;
    return s;
}

// ---- raw shared-memory access by 32-bit shared address (exact instructions, no generic ptrs) ---
SOT_DEVINL float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
SOT_DEVINL void lds64(uint32_t a, float& x, float& y) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}
SOT_DEVINL void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
SOT_DEVINL void sts64(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
// One merge step's advance: the consumed side (v if b < a, else u: u first on equal values, the
// stable order of `cat(cu, cv)`) loads its next pair and moves its address; both halves are
// predicated -- no branch, no select.  `consumed` returns the address of the pair that was taken.
SOT_DEVINL void advance(float& a, float& pa, float& b, float& pb, uint32_t& adrA, uint32_t& adrB,
                        uint32_t& consumed) {
    asm volatile(
        "{\n"
        ".reg .pred tv;\n"
        "setp.lt.f32 tv, %2, %0;\n"
        "selp.u32 %6, %5, %4, tv;\n"
        "@tv  ld.shared.v2.f32 {%2, %3}, [%5+8];\n"
        "@!tv ld.shared.v2.f32 {%0, %1}, [%4+8];\n"
        "@tv  add.u32 %5, %5, 8;\n"
        "@!tv add.u32 %4, %4, 8;\n"
        "}"
        : "+f"(a), "+f"(pa), "+f"(b), "+f"(pb), "+r"(adrA), "+r"(adrB), "=r"(consumed));
}
SOT_DEVINL void advance_fwd(float& a, float& pa, float& b, float& pb, uint32_t& adrA, uint32_t& adrB) {
    asm volatile(
        "{\n"
        ".reg .pred tv;\n"
        "setp.lt.f32 tv, %2, %0;\n"
        "@tv  ld.shared.v2.f32 {%2, %3}, [%5+8];\n"
        "@!tv ld.shared.v2.f32 {%0, %1}, [%4+8];\n"
        "@tv  add.u32 %5, %5, 8;\n"
        "@!tv add.u32 %4, %4, 8;\n"
        "}"
        : "+f"(a), "+f"(pa), "+f"(b), "+f"(pb), "+r"(adrA), "+r"(adrB));
}

// fp64 reciprocal of a positive float >= 1e-7: hardware approximation + two Newton steps
SOT_DEVINL double recip_f64(float x) {
    const double xd = static_cast<double>(x);
    double r = static_cast<double>(__fdividef(1.0f, x));
    r = fma(r, fma(-xd, r, 1.0), r);
    r = fma(r, fma(-xd, r, 1.0), r);
    return r;
}

// Exclusive scan (prefix; suffix when REVERSE) of one fp64 value over the TPF threads of a group.
template <int TPF, bool REVERSE>
SOT_DEVINL void group_scan1(double& a, double& total, double* slot, int tid, int g) {
    constexpr int NW = TPF / 32;
    const int lane = tid & 31, w = tid >> 5;
    double ia = a;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double ya = REVERSE ? __shfl_down_sync(FULL_MASK, ia, off) : __shfl_up_sync(FULL_MASK, ia, off);
        const bool ok = REVERSE ? (lane + off < 32) : (lane >= off);
        if (ok) ia += ya;
    }
    double ea = REVERSE ? __shfl_down_sync(FULL_MASK, ia, 1) : __shfl_up_sync(FULL_MASK, ia, 1);
    if (lane == (REVERSE ? 31 : 0)) ea = 0.0;
    const double wa = __shfl_sync(FULL_MASK, ia, REVERSE ? 0 : 31);
    if constexpr (NW == 1) {
        a = ea;
        total = wa;
    } else {
        if (lane == 0) slot[w] = wa;
        group_sync<TPF>(g);
        double pa = 0.0, ta = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int kk = REVERSE ? (NW - 1 - k) : k;
            const double sa = slot[kk];
            if (REVERSE ? (kk > w) : (kk < w)) pa = ta + sa;
            ta += sa;
        }
        a = pa + ea;
        total = ta;
    }
}

// Register budget per thread -> resident CTAs per SM requested from ptxas.
constexpr int reg_budget(int e, int out) {
    return out == OUT_GRAD ? (e <= 9 ? 64 : (e <= 17 ? 96 : 128)) : (e <= 9 ? 64 : (e <= 17 ? 80 : 128));
}
constexpr int min_ctas(int threads, int e, int out) {
    const int c = 65536 / (threads * reg_budget(e, out));
    return c < 1 ? 1 : (c > 16 ? 16 : c);
}
__host__ __device__ constexpr int ilog2_ceil(int x) {
    int b = 0;
    while ((1 << b) < x) ++b;
    return b;
}

template <int TPF, int E, int FPC, int PMODE, int OUT, int MODE>
__global__ void __launch_bounds__(FPC* TPF, min_ctas(FPC* TPF, E, OUT)) sot_frames_kernel(const FrameArgs args) {
    constexpr bool WITH_GRAD = (OUT == OUT_GRAD);
    constexpr int NT = FPC * TPF;
    constexpr int NW = TPF / 32;
    constexpr int SEARCH_TOP = 1 << (ilog2_ceil(TPF * E + 1) - 1);  // largest power of two <= capacity
    constexpr uint32_t NO_FIX = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];

    const int n = args.n, m = args.m, K = n + m, S = n + m + 2;
    const bool square = args.flags & FLAG_SQUARE;
    const bool cut_scale = args.flags & FLAG_CUT_SCALE;
    const bool pos_shared = (args.pos_u_stride == 0) && (args.pos_v_stride == 0);
    // contributions with q above `thr` are dropped: the strict `qs > 1` mask (losses.py:307) when
    // limiting, otherwise nothing a real slot can reach
    const float thr = (args.flags & FLAG_LIMIT) ? 1.0f : FLT_BIG;
    const SmemPlan plan = smem_plan(FPC, n, m, TPF, pos_shared);
    const uint32_t sb = smem_u32(smem);

    const int g = threadIdx.x / TPF, tid = threadIdx.x % TPF;
    const int e0 = tid * E;
    float* const land = reinterpret_cast<float*>(smem + plan.land);
    float* const posf = reinterpret_cast<float*>(smem + plan.pos);
    float* const outf = reinterpret_cast<float*>(smem + plan.pairs);  // output staging (end of iteration)
    double* const scratch = reinterpret_cast<double*>(smem + plan.scratch) + g * 16 * NW;
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(smem + plan.mbar);
    const uint32_t landU = sb + plan.land + 4u * (g * n + e0);            // my first raw u bin
    const uint32_t landV = sb + plan.land + 4u * (FPC * n + g * m + e0);  // my first raw v bin
    const uint32_t PA0 = sb + plan.pairs + 8u * (g * S);                  // pair A[0] of my frame
    const uint32_t PB0 = PA0 + 8u * (n + 1);                              // pair B[0]
    const uint32_t posU = sb + plan.pos + 4u * ((pos_shared ? 0 : g * S) + e0);
    const uint32_t posV = posU + 4u * (n + 1);
    const uint32_t mbox = sb + plan.mailbox + 8u * (g * (TPF + 1));
    const uint32_t carry = sb + plan.carry + 4u * (g * TPF);
    const bool in_u = e0 + E <= n, in_v = e0 + E <= m;  // all of my E bins exist (no guards needed)

    // L consecutive merged slots per thread.  L is ODD on purpose: for a balanced merge thread t starts
    // near pair L*t/2 of each row, and an even L would put the 16 lanes of a half-warp on only two
    // distinct bank pairs (8-way conflicts on every LDS.64 of the walk, measured); odd L spreads them.
    const int L = ((K + TPF - 1) / TPF) | 1;
    const int k0 = min(tid * L, K);
    const int cnt = min(L, K - k0);
    const uint32_t GOFF = (sb + plan.land + 4u * (g * S)) - (PA0 >> 1);  // pair address -> dL/dCDF address
    (void)GOFF;

    const long long n_quads = (args.n_frames + FPC - 1) / FPC;
    const bool rows16 = ((FPC * n) % 4 == 0) && ((FPC * m) % 4 == 0);  // always true for FPC == 4
    const bool in_aligned = rows16 && ((reinterpret_cast<uintptr_t>(args.u) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(args.v) & 15) == 0);
    const bool out_aligned = rows16 && ((reinterpret_cast<uintptr_t>(args.grad_u) & 15) == 0) &&
                             ((reinterpret_cast<uintptr_t>(args.grad_v) & 15) == 0);

    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    if (pos_shared) {  // positions once per CTA; entry [n] / [m] repeats the last one (index clamp)
        for (int idx = threadIdx.x; idx < S; idx += NT)
            posf[idx] = idx <= n ? args.pos_u[min(idx, n - 1)] : args.pos_v[min(idx - n - 1, m - 1)];
    }
    if (tid == 0) sts64(mbox + 8u * TPF, f_inf(), 0.0f);  // virtual slot K: a new group with m*d = 0
    __syncthreads();

    auto issue_load = [&](long long quad) {  // one elected thread
        const long long f0 = quad * FPC;
        mbar_expect_tx(mbar, 4u * FPC * static_cast<uint32_t>(n + m));
        bulk_g2s(land, args.u + f0 * n, 4u * FPC * n, mbar);
        bulk_g2s(land + FPC * n, args.v + f0 * m, 4u * FPC * m, mbar);
    };
    auto quad_is_bulk = [&](long long quad) { return in_aligned && (quad + 1) * FPC <= args.n_frames; };

    long long quad = blockIdx.x;
    uint32_t parity = 0;
    if (quad < n_quads && quad_is_bulk(quad) && threadIdx.x == 0) issue_load(quad);

    for (; quad < n_quads; quad += gridDim.x) {
        const long long frame0 = quad * FPC;
        const long long left = args.n_frames - frame0;
        const int nfr = left < FPC ? static_cast<int>(left) : FPC;
        const bool active = g < nfr;
        const long long frame = frame0 + g;

        // ---- stage 0: the quad's raw rows are (or get) in LAND --------------------------------
        if (quad_is_bulk(quad)) {
            mbar_wait(mbar, parity);
            parity ^= 1;
        } else {  // ragged last quad or 4-byte aligned base pointers: plain coalesced loads
            const float* gu = args.u + frame0 * n;
            const float* gv = args.v + frame0 * m;
            for (int idx = threadIdx.x; idx < nfr * n; idx += NT) land[idx] = gu[idx];
            for (int idx = threadIdx.x; idx < nfr * m; idx += NT) land[FPC * n + idx] = gv[idx];
            __syncthreads();
        }
        if (!pos_shared) {  // per-frame supports: (re)load this quad's rows
            for (int idx = threadIdx.x; idx < nfr * S; idx += NT) {
                const int r = idx / S, c = idx - r * S;
                posf[idx] = c <= n ? args.pos_u[(frame0 + r) * args.pos_u_stride + min(c, n - 1)]
                                   : args.pos_v[(frame0 + r) * args.pos_v_stride + min(c - n - 1, m - 1)];
            }
        }

        // ---- stage 1: blocked read of my E bins of each row (conflict free: E is odd) ---------
        float xu[E], xv[E];
        if (in_u) {
#pragma unroll
            for (int c = 0; c < E; ++c) xu[c] = lds32(landU + 4 * c);
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) xu[c] = (e0 + c < n) ? lds32(landU + 4 * c) : 0.0f;
        }
        if (in_v) {
#pragma unroll
            for (int c = 0; c < E; ++c) xv[c] = lds32(landV + 4 * c);
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) xv[c] = (e0 + c < m) ? lds32(landV + 4 * c) : 0.0f;
        }
        if constexpr (WITH_GRAD) {
            // the previous iteration's output store must have finished READING the staging area
            // (PAIRS region) before stage 2 of this iteration overwrites it
            if (threadIdx.x == 0) bulk_wait_read_all();
        }
        __syncthreads();  // LAND is free again (and, per-frame supports, POS is complete)
        if constexpr (!WITH_GRAD) {  // (the gradient kernel reuses LAND for dL/dCDF first)
            const long long next = quad + gridDim.x;
            if (threadIdx.x == 0 && next < n_quads && quad_is_bulk(next)) issue_load(next);
        }

        float acc = 0.0f;    // my part of the frame's loss
        bool finite = true;  // masses (and the scaled totals) are finite numbers
        double inv_u = 1.0, inv_v = 1.0;
        bool u_live = false, v_live = false;  // mass above the safe_divide floor -> carries gradient
        float og_u[E], og_v[E];               // (WITH_GRAD) finished gradient values of my bins
        (void)og_u;
        (void)og_v;
        (void)inv_v;
        (void)u_live;
        (void)v_live;

        if (active) {
            // ---- stage 2: masses and CDFs (fp64 accumulation, one rounding to fp32 per entry),
            //      written as (cdf, position) pairs -------------------------------------------------
            if constexpr (MODE == MODE_SPECTRA) {
                double P[E];
                double t = 0.0;
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    t += static_cast<double>(square ? xu[c] * xu[c] : xu[c]);
                    P[c] = t;
                }
                double off = t, tot_u;
                group_scan1<TPF, false>(off, tot_u, scratch, tid, g);
                const float mass_u = static_cast<float>(tot_u);
                u_live = mass_u > SAFE_EPS;  // utils.py:137: den <= eps -> eps
                inv_u = recip_f64(u_live ? mass_u : SAFE_EPS);
                if (args.flags & FLAG_RAW) inv_u = 1.0;
                if (in_u) {
#pragma unroll
                    for (int c = 0; c < E; ++c)
                        sts64(PA0 + 8 * (e0 + c), static_cast<float>((off + P[c]) * inv_u), lds32(posU + 4 * c));
                } else {
#pragma unroll
                    for (int c = 0; c < E; ++c)
                        if (e0 + c < n)
                            sts64(PA0 + 8 * (e0 + c), static_cast<float>((off + P[c]) * inv_u), lds32(posU + 4 * c));
                }
                t = 0.0;
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    t += static_cast<double>(square ? xv[c] * xv[c] : xv[c]);
                    P[c] = t;
                }
                double tot_v;
                off = t;
                group_scan1<TPF, false>(off, tot_v, scratch + NW, tid, g);
                const float mass_v = static_cast<float>(tot_v);
                v_live = mass_v > SAFE_EPS;
                inv_v = cut_scale ? inv_u : recip_f64(v_live ? mass_v : SAFE_EPS);
                if (args.flags & FLAG_RAW) {  // no normalisation, hence no mass term in the gradient
                    inv_v = 1.0;
                    u_live = v_live = false;
                }
                if (in_v) {
#pragma unroll
                    for (int c = 0; c < E; ++c)
                        sts64(PB0 + 8 * (e0 + c), static_cast<float>((off + P[c]) * inv_v), lds32(posV + 4 * c));
                } else {
#pragma unroll
                    for (int c = 0; c < E; ++c)
                        if (e0 + c < m)
                            sts64(PB0 + 8 * (e0 + c), static_cast<float>((off + P[c]) * inv_v), lds32(posV + 4 * c));
                }
                // NaN / inf anywhere (or an overflowing cut-mode scale) poisons the frame: the walk
                // is skipped (its +inf sentinels must stay unique) and NaN is written instead
                finite = (fabs(tot_u * inv_u) <= static_cast<double>(FLT_BIG)) &&
                         (fabs(tot_v * inv_v) <= static_cast<double>(FLT_BIG));
            } else {
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    if (e0 + c < n) sts64(PA0 + 8 * (e0 + c), fminf(xu[c], FLT_BIG), lds32(posU + 4 * c));
                    if (e0 + c < m) sts64(PB0 + 8 * (e0 + c), fminf(xv[c], FLT_BIG), lds32(posV + 4 * c));
                }
            }
            if (tid == 0) {
                sts64(PA0 + 8 * n, f_inf(), lds32(posU + 4 * n));  // tid 0: e0 == 0
                sts64(PB0 + 8 * m, f_inf(), lds32(posV + 4 * m));
            }
            group_sync<TPF>(g);

            if (finite) {
                // ---- stage 3a: merge-path partition (fixed-trip, branch-free bit descent) --------
                // i0 = number of u entries among the first k0 merged slots (u first on equal values):
                // the largest i in [lo, hi] with A[i-1] <= B[k0-i]
                int i0;
                {
                    const int lo = max(0, k0 - m), hi = min(k0, n);
                    uint32_t cur = PA0 + 8u * lo;  // address of A[i] for the current i
                    const uint32_t hiA = PA0 + 8u * hi;
                    const uint32_t sumAB = PA0 + PB0 + 8u * k0;  // addr(A[i]) + addr(B[k0-i]) is constant
#pragma unroll
                    for (int step = SEARCH_TOP; step >= 1; step >>= 1) {
                        const uint32_t cand = cur + 8u * step;
                        if (cand <= hiA) {
                            const float av = lds32(cand - 8);      // A[i-1] for the candidate i
                            const float bv = lds32(sumAB - cand);  // B[k0-i]
                            if (av <= bv) cur = cand;
                        }
                    }
                    i0 = static_cast<int>((cur - PA0) >> 3);
                }
                uint32_t adrA = PA0 + 8u * i0, adrB = PB0 + 8u * (k0 - i0);
                float a, pa, b, pb;
                lds64(adrA, a, pa);
                lds64(adrB, b, pb);
                float qprev = 0.0f;  // the zero the reference pads in front of qs (losses.py:301)
                if (k0 > 0) {
                    const float al = i0 > 0 ? lds32(adrA - 8) : -f_inf();
                    const float bl = k0 - i0 > 0 ? lds32(adrB - 8) : -f_inf();
                    qprev = fmaxf(al, bl);
                }

                if constexpr (OUT == OUT_LOSS) {
                    // ---- stage 3b (forward): walk my slots ------------------------------------------
#pragma unroll 4
                    for (int s = 0; s < cnt; ++s) {
                        const float q = fminf(a, b);
                        const float D = transport_cost<PMODE>(pa, pb, args.p);
                        float dq = q - qprev;
                        dq = (q > thr) ? 0.0f : dq;
                        acc = fmaf(dq, D, acc);
                        qprev = q;
                        advance_fwd(a, pa, b, pb, adrA, adrB);
                    }
                } else if constexpr (OUT == OUT_GRAD) {
                    // ---- stage 3b (gradient): walk + dL/dCDF in place ---------------------------------
                    // m*d of a tie group is fixed at its first slot; dL/dCDF is nonzero only at a
                    // group's LAST slot: G = (m*d)_group - (m*d)_next group.  It is
                    // stored (into the mirrored dL/dCDF array) one step later, when the next slot is known.
                    float md_prev;
                    bool inherited;  // the open group started before my range: its m*d is not known yet
                    {
                        const float q = fminf(a, b);
                        const float D = transport_cost<PMODE>(pa, pb, args.p);
                        const float fm = (q > thr) ? 0.0f : D;
                        // what my left neighbour needs to close ITS last slot: my first value and the
                        // m*d my first slot has if it opens a group (idle thread: the end marker)
                        sts64(mbox + 8u * tid, cnt > 0 ? q : f_inf(), cnt > 0 ? fm : 0.0f);
                        inherited = (k0 > 0) && (q == qprev);
                        md_prev = inherited ? 0.0f : fm;
                        float dq = q - qprev;
                        dq = (q > thr) ? 0.0f : dq;
                        if (cnt > 0) {
                            acc = fmaf(dq, D, acc);
                            qprev = q;
                        }
                    }
                    group_sync<TPF>(g);  // mailbox complete before anyone reaches its final peek
                    uint32_t consumed = 0, fix = NO_FIX;
                    if (cnt > 0) advance(a, pa, b, pb, adrA, adrB, consumed);
                    auto step = [&]() {
                        const float q = fminf(a, b);
                        const float D = transport_cost<PMODE>(pa, pb, args.p);
                        const bool over = q > thr;
                        float dq = q - qprev;
                        dq = over ? 0.0f : dq;
                        acc = fmaf(dq, D, acc);
                        const bool same = (q == qprev);
                        const float md = same ? md_prev : (over ? 0.0f : D);
                        sts32((consumed >> 1) + GOFF, md_prev - md);  // dL/dCDF of the previous slot (0 inside a group)
                        fix = (inherited && !same) ? consumed : fix;
                        inherited = inherited && same;
                        md_prev = md;
                        qprev = q;
                        advance(a, pa, b, pb, adrA, adrB, consumed);
                    };
#pragma unroll 4
                    for (int s = 1; s < cnt; ++s) step();
                    float carry_out = -1.0f;  // -1 = "my whole range continues a group opened before me"
                    if (cnt > 0) {
                        // the slot after my range: first slot of the next thread, or the end marker
                        float qn, mdn;
                        lds64(mbox + 8u * (tid + 1), qn, mdn);
                        const bool same = (qn == qprev);
                        const float md = same ? md_prev : mdn;
                        sts32((consumed >> 1) + GOFF, md_prev - md);
                        fix = (inherited && !same) ? consumed : fix;
                        inherited = inherited && same;
                        carry_out = inherited ? -1.0f : md_prev;
                    }
                    // look-back: add the m*d of a tie group that was opened by an earlier thread
                    sts32(carry + 4u * tid, carry_out);
                    group_sync<TPF>(g);
                    if (fix != NO_FIX) {
                        uint32_t s = carry + 4u * (tid - 1);
                        float c = lds32(s);
                        while (c < 0.0f) {  // thread 0 never carries the marker
                            s -= 4;
                            c = lds32(s);
                        }
                        sts32((fix >> 1) + GOFF, lds32((fix >> 1) + GOFF) + c);
                    }
                    group_sync<TPF>(g);
                } else {
                    // ---- stage 3b (plan): emit qs, lower-bound indices, quantile positions, loss ------
                    // Same partition and tie-group rule; the per-group payload is the pair of lower-bound
                    // indices (#{cu < q}, #{cv < q}) -- what `searchsorted(cu, qs)` / `searchsorted(cv, qs)`
                    // return (losses.py:219).
                    int i = i0, j = k0 - i0, is = 0, js = 0, n_inherited = 0;
                    bool inherited = (k0 != 0);
                    for (int s = 0; s < cnt; ++s) {
                        const float q = fminf(a, b);
                        const bool take_v = b < a;
                        if (q != qprev) {
                            is = i;
                            js = j;
                            inherited = false;
                        }
                        if (inherited) ++n_inherited;
                        const long long o = frame * K + k0 + s;
                        if (args.plan_qs != nullptr) args.plan_qs[o] = q;
                        if (args.plan_iu != nullptr) args.plan_iu[o] = is;
                        if (args.plan_iv != nullptr) args.plan_iv[o] = js;
                        if (args.plan_uq != nullptr) args.plan_uq[o] = lds32(PA0 + 8u * is + 4);
                        if (args.plan_vq != nullptr) args.plan_vq[o] = lds32(PB0 + 8u * js + 4);
                        float dq = q - qprev;
                        dq = (q > thr) ? 0.0f : dq;
                        acc = fmaf(dq, transport_cost<PMODE>(pa, pb, args.p), acc);
                        qprev = q;
                        if (take_v) ++j; else ++i;
                        advance_fwd(a, pa, b, pb, adrA, adrB);
                    }
                    sts32(carry + 4u * tid, __int_as_float((cnt == 0 || inherited) ? -1 : ((is << 16) | js)));
                    group_sync<TPF>(g);
                    if (n_inherited > 0) {
                        uint32_t s = carry + 4u * (tid - 1);
                        int c = __float_as_int(lds32(s));
                        while (c < 0) {
                            s -= 4;
                            c = __float_as_int(lds32(s));
                        }
                        const int is2 = c >> 16, js2 = c & 0xffff;
                        for (int t = 0; t < n_inherited; ++t) {
                            const long long o = frame * K + k0 + t;
                            if (args.plan_iu != nullptr) args.plan_iu[o] = is2;
                            if (args.plan_iv != nullptr) args.plan_iv[o] = js2;
                            if (args.plan_uq != nullptr) args.plan_uq[o] = lds32(PA0 + 8u * is2 + 4);
                            if (args.plan_vq != nullptr) args.plan_vq[o] = lds32(PB0 + 8u * js2 + 4);
                        }
                    }
                    if (args.plan_cu != nullptr)
                        for (int e = tid; e < n; e += TPF) args.plan_cu[frame * n + e] = lds32(PA0 + 8u * e);
                    if (args.plan_cv != nullptr)
                        for (int e = tid; e < m; e += TPF) args.plan_cv[frame * m + e] = lds32(PB0 + 8u * e);
                    group_sync<TPF>(g);  // the pairs are rewritten by the next iteration's stage 2
                }
            }  // finite

            // ---- loss of the frame: fp32 partials, summed in fp64 -----------------------------------
            {
                double part = static_cast<double>(acc);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULL_MASK, part, off);
                if constexpr (NW > 1) {
                    if ((tid & 31) == 0) scratch[2 * NW + (tid >> 5)] = part;
                    group_sync<TPF>(g);
                    part = 0.0;
#pragma unroll
                    for (int k = 0; k < NW; ++k) part += scratch[2 * NW + k];
                }
                if (tid == 0 && args.loss != nullptr) args.loss[frame] = finite ? static_cast<float>(part) : f_nan();
            }

            if constexpr (WITH_GRAD) {
                if constexpr (MODE == MODE_SPECTRA) {
                    // ---- stage 4: cumsum transpose (suffix sums of dL/dCDF) and the normalisation chain
                    // rule.  sum_i gw_i w_i = sum_i (dL/dc_i) c_i (Abel summation), so the mass term needs
                    // only the CDF values and dL/dCDF that are already in shared memory.
                    const uint32_t GU0 = (PA0 >> 1) + GOFF + 4u * e0, GV0 = (PB0 >> 1) + GOFF + 4u * e0;
                    float lsu[E], lsv[E];
                    float su = 0.0f, sv = 0.0f, du = 0.0f, dv = 0.0f;
#pragma unroll
                    for (int c = E - 1; c >= 0; --c) {
                        if (e0 + c < n) {
                            const float gq = lds32(GU0 + 4 * c);
                            su += gq;
                            du = fmaf(gq, lds32(PA0 + 8 * (e0 + c)), du);
                        }
                        lsu[c] = su;
                        if (e0 + c < m) {
                            const float gq = lds32(GV0 + 4 * c);
                            sv += gq;
                            dv = fmaf(gq, lds32(PB0 + 8 * (e0 + c)), dv);
                        }
                        lsv[c] = sv;
                    }
                    double off_u = static_cast<double>(su), off_v = static_cast<double>(sv), tu, tv;
                    group_scan1<TPF, true>(off_u, tu, scratch + 4 * NW, tid, g);
                    group_scan1<TPF, true>(off_v, tv, scratch + 5 * NW, tid, g);
                    double dot_u = static_cast<double>(du), dot_v = static_cast<double>(dv);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        dot_u += __shfl_xor_sync(FULL_MASK, dot_u, off);
                        dot_v += __shfl_xor_sync(FULL_MASK, dot_v, off);
                    }
                    if constexpr (NW > 1) {
                        if ((tid & 31) == 0) {
                            scratch[6 * NW + (tid >> 5)] = dot_u;
                            scratch[7 * NW + (tid >> 5)] = dot_v;
                        }
                        group_sync<TPF>(g);
                        dot_u = 0.0;
                        dot_v = 0.0;
#pragma unroll
                        for (int k = 0; k < NW; ++k) {
                            dot_u += scratch[6 * NW + k];
                            dot_v += scratch[7 * NW + k];
                        }
                    }
                    // a clamped mass has no derivative (torch.where picks the constant branch)
                    const double corr_u = u_live ? (cut_scale ? dot_u + dot_v : dot_u) : 0.0;
                    const double corr_v = (!cut_scale && v_live) ? dot_v : 0.0;
                    // gw = offset + local suffix; (offset - corr) is formed in fp64 before rounding
                    const float bu = static_cast<float>(off_u - corr_u), bv = static_cast<float>(off_v - corr_v);
                    const float up = args.upstream != nullptr ? args.upstream[frame] : 1.0f;
                    const float ku = static_cast<float>(inv_u) * up * (square ? 2.0f : 1.0f);
                    const float kv = static_cast<float>(inv_v) * up * (square ? 2.0f : 1.0f);
#pragma unroll
                    for (int c = 0; c < E; ++c) {
                        float ga = (bu + lsu[c]) * ku;
                        float gb = (bv + lsv[c]) * kv;
                        if (square) {
                            ga *= xu[c];
                            gb *= xv[c];
                        }
                        og_u[c] = finite ? ga : f_nan();
                        og_v[c] = finite ? gb : f_nan();
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < E; ++c) {
                        og_u[c] = (e0 + c < n) ? lds32((PA0 >> 1) + GOFF + 4 * (e0 + c)) : 0.0f;
                        og_v[c] = (e0 + c < m) ? lds32((PB0 >> 1) + GOFF + 4 * (e0 + c)) : 0.0f;
                    }
                }
            }
        }  // active

        // ---- stage 5: stage the gradient rows (FPC rows back to back) and store them -------------
        if constexpr (WITH_GRAD) {
            float* const ou = args.grad_u != nullptr ? args.grad_u + frame0 * n : nullptr;
            float* const ov = args.grad_v != nullptr ? args.grad_v + frame0 * m : nullptr;
            const bool bulk_out = out_aligned && (nfr == FPC);
            fence_async_smem();  // order my generic-proxy accesses to LAND / PAIRS before the TMA that follows
            __syncthreads();     // every group is done with its pairs and its dL/dCDF array
            {
                const long long next = quad + gridDim.x;
                if (threadIdx.x == 0 && next < n_quads && quad_is_bulk(next)) issue_load(next);
            }
            if (active) {
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    if (e0 + c < n) outf[g * n + e0 + c] = og_u[c];
                    if (e0 + c < m) outf[FPC * n + g * m + e0 + c] = og_v[c];
                }
            }
            if (bulk_out) fence_async_smem();
            __syncthreads();
            if (bulk_out) {
                if (threadIdx.x == 0) {
                    if (ou != nullptr) bulk_s2g(ou, outf, 4u * FPC * n);
                    if (ov != nullptr) bulk_s2g(ov, outf + FPC * n, 4u * FPC * m);
                    bulk_commit();
                }
            } else {
                if (ou != nullptr)
                    for (int idx = threadIdx.x; idx < nfr * n; idx += NT) ou[idx] = outf[idx];
                if (ov != nullptr)
                    for (int idx = threadIdx.x; idx < nfr * m; idx += NT) ov[idx] = outf[FPC * n + idx];
                __syncthreads();  // the plain stores have read the staging area
            }
        }
    }
    if constexpr (WITH_GRAD) {
        if (threadIdx.x == 0) bulk_wait_read_all();  // shared memory must outlive the last bulk store
    }
}

}  // namespace sot
