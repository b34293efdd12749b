// SOT frame kernel: per-frame 1-D Wasserstein (W_p^p) between two spectra, forward and fused
// forward+backward, one frame per group of TPF threads, FPC = 4 frames per CTA.
//
// Reference semantics reproduced (file:line into /root/reference):
//   losses.py:172-184  square, mass, safe_divide (cut mode divides the prediction by the
//                      TARGET's mass), utils.py:135-142
//   losses.py:292-293  inclusive CDFs                     -> fp64 block scan, fp32 storage
//   losses.py:295      sort(cat(cu, cv))                  -> merge-path partition + sequential walk
//   losses.py:214-220  searchsorted-left + clamp + gather -> running co-ranks (i, j) of the walk
//   losses.py:301-313  sum_k dq_k * |uq_k - vq_k|^p, strict `qs > 1` mask
//   autograd of all of the above          -> scatter of dL/dCDF, two suffix scans,
//                                                           normalisation chain rule, 2x
// Data layout in shared memory (floats), n = bins of u, m = bins of v:
//   zoneC : FPC*(n+1) + FPC*(m+1)   landing zone of the TMA bulk loads (4 raw rows of u back to
//                                   back, then 4 raw rows of v), later overwritten by the CDF rows
//                                   (stride n+1 / m+1, +inf sentinel at the end of each row)
//   zoneG : FPC*n + FPC*m           (WITH_GRAD) scatter target for dL/dCDF, then suffix sums, then
//                                   the finished gradient rows = source of the TMA bulk store
//   pos   : (n+1) + (m+1) per CTA (shared supports) or per frame; entry [n] repeats [n-1]
//                                   (the reference's clamp of the searchsorted index, :220)
//   scratch / carry / mbarrier
// Thread t of a group owns the E consecutive bins [t*E, t*E+E) of both rows ("blocked"); E is
// odd so that every blocked shared-memory access of a warp is bank-conflict free and no 16-byte
// alignment of the 4*n-byte rows is ever needed (n = 1025 / 257: rows are only 4-byte aligned).
#pragma once
#include "sot_device.cuh"

namespace sot {

// FPC = frames per CTA.  4 rows of 4*n bytes are always a 16-byte multiple (the TMA bulk unit);
// FPC = 1 exists for rows so long that four of them do not fit in shared memory.

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

// Shared-memory carve-up, shared between host (size) and device (pointers).
struct SmemPlan {
    int zoneC, zoneG, posU, posV, scratch, carry, mbar, total;  // byte offsets
};
__host__ __device__ inline SmemPlan smem_plan(int FPC, int n, int m, int tpf, bool with_grad,
                                             bool pos_u_shared, bool pos_v_shared) {
    SmemPlan s;
    int off = 0;
    s.zoneC = off;
    off += 4 * (FPC * (n + 1) + FPC * (m + 1));
    s.zoneG = off;
    if (with_grad) off += 4 * (FPC * n + FPC * m);
    s.posU = off;
    off += 4 * (n + 1) * (pos_u_shared ? 1 : FPC);
    s.posV = off;
    off += 4 * (m + 1) * (pos_v_shared ? 1 : FPC);
    off = (off + 15) & ~15;
    s.scratch = off;
    off += 8 * FPC * 3 * 2 * (tpf / 32);  // 3 collectives x 2 values x warps, per group
    s.carry = off;
    off += 4 * FPC * tpf;
    off = (off + 15) & ~15;
    s.mbar = off;
    off += 16;
    s.total = off;
    return s;
}

// Register budget per thread -> resident CTAs per SM requested from ptxas.  The gradient walk
// wants ~120 registers if left alone, which would leave a single 512-thread CTA per SM.
constexpr int reg_budget(int e, int out) {
    return out == OUT_GRAD ? (e <= 17 ? 64 : 128) : (e <= 9 ? 40 : (e <= 17 ? 64 : 96));
}
constexpr int min_ctas(int threads, int e, int out) {
    const int c = 65536 / (threads * reg_budget(e, out));
    return c < 1 ? 1 : (c > 16 ? 16 : c);
}

template <int TPF, int E, int FPC, int PMODE, int OUT, int MODE>
__global__ void __launch_bounds__(FPC* TPF, min_ctas(FPC* TPF, E, OUT)) sot_frames_kernel(const FrameArgs args) {
    constexpr bool WITH_GRAD = (OUT == OUT_GRAD);
    constexpr int NO_FIX = static_cast<int>(0x80000000);
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = args.n, m = args.m, K = n + m;
    const bool square = args.flags & FLAG_SQUARE;
    const bool cut_scale = args.flags & FLAG_CUT_SCALE;
    const bool limit = args.flags & FLAG_LIMIT;
    const bool shared_pu = args.pos_u_stride == 0, shared_pv = args.pos_v_stride == 0;
    const SmemPlan plan = smem_plan(FPC, n, m, TPF, WITH_GRAD, shared_pu, shared_pv);

    float* const zoneC = reinterpret_cast<float*>(smem + plan.zoneC);
    float* const landU = zoneC;                  // raw rows, stride n
    float* const landV = zoneC + FPC * (n + 1);  // raw rows, stride m
    float* const zoneG = reinterpret_cast<float*>(smem + plan.zoneG);
    float* const posU_all = reinterpret_cast<float*>(smem + plan.posU);
    float* const posV_all = reinterpret_cast<float*>(smem + plan.posV);
    double* const scratch_all = reinterpret_cast<double*>(smem + plan.scratch);
    float* const carry_all = reinterpret_cast<float*>(smem + plan.carry);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(smem + plan.mbar);

    const int g = threadIdx.x / TPF, tid = threadIdx.x % TPF;
    const long long frame0 = static_cast<long long>(blockIdx.x) * FPC;
    const long long left = args.n_frames - frame0;
    const int nfr = left < FPC ? static_cast<int>(left) : FPC;
    const bool active = g < nfr;
    const long long frame = frame0 + g;

    const float* const gu = args.u + frame0 * n;
    const float* const gv = args.v + frame0 * m;
    const bool rows16 = ((FPC * n) % 4 == 0) && ((FPC * m) % 4 == 0);  // always true for FPC == 4
    const bool bulk_in = rows16 && (nfr == FPC) && ((reinterpret_cast<uintptr_t>(gu) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(gv) & 15) == 0);

    // ---- stage 0: bring 4 raw rows of u and of v into shared memory ------------------------
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (bulk_in) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(mbar, 4u * FPC * static_cast<uint32_t>(n + m));
            bulk_g2s(landU, gu, 4u * FPC * n, mbar);
            bulk_g2s(landV, gv, 4u * FPC * m, mbar);
        }
    } else {  // ragged tail CTA or a base pointer that is not 16-byte aligned: plain coalesced loads
        for (int idx = threadIdx.x; idx < nfr * n; idx += FPC * TPF) landU[idx] = gu[idx];
        for (int idx = threadIdx.x; idx < nfr * m; idx += FPC * TPF) landV[idx] = gv[idx];
    }
    // supports (L2-resident after the first CTA), while the bulk copies are in flight
    if (shared_pu) {
        for (int idx = threadIdx.x; idx <= n; idx += FPC * TPF) posU_all[idx] = args.pos_u[min(idx, n - 1)];
    } else {
        for (int idx = threadIdx.x; idx < nfr * (n + 1); idx += FPC * TPF) {
            const int r = idx / (n + 1), c = idx - r * (n + 1);
            posU_all[idx] = args.pos_u[(frame0 + r) * args.pos_u_stride + min(c, n - 1)];
        }
    }
    if (shared_pv) {
        for (int idx = threadIdx.x; idx <= m; idx += FPC * TPF) posV_all[idx] = args.pos_v[min(idx, m - 1)];
    } else {
        for (int idx = threadIdx.x; idx < nfr * (m + 1); idx += FPC * TPF) {
            const int r = idx / (m + 1), c = idx - r * (m + 1);
            posV_all[idx] = args.pos_v[(frame0 + r) * args.pos_v_stride + min(c, m - 1)];
        }
    }
    if (bulk_in) mbar_wait(mbar, 0);
    __syncthreads();

    // ---- stage 1: blocked read of my E bins of each row (conflict free: E is odd) -----------
    float xu[E], xv[E];
    const int e0 = tid * E;
#pragma unroll
    for (int c = 0; c < E; ++c) {
        const int e = e0 + c;
        xu[c] = (active && e < n) ? landU[g * n + e] : 0.0f;
        xv[c] = (active && e < m) ? landV[g * m + e] : 0.0f;
    }
    __syncthreads();  // the landing zone is now dead: CDF rows (stride n+1) may overwrite it

    float* const cu = zoneC + g * (n + 1);
    float* const cv = zoneC + FPC * (n + 1) + g * (m + 1);
    const float* const pu = posU_all + (shared_pu ? 0 : g * (n + 1));
    const float* const pv = posV_all + (shared_pv ? 0 : g * (m + 1));
    double* const scratch = scratch_all + g * (3 * 2 * (TPF / 32));
    float* const carry = carry_all + g * TPF;
    float* const gzu = zoneG + g * n;            // dL/dcu scatter row, later grad_u row
    float* const gzv = zoneG + FPC * n + g * m;  // dL/dcv scatter row, later grad_v row

    double inv_u = 1.0, inv_v = 1.0;  // 1 / clamped mass (fp64 reciprocal of the fp32 mass)
    bool u_live = true, v_live = true;  // mass above the safe_divide floor -> it carries gradient
    bool finite = true;

    if (active) {
        // ---- stage 2: masses and CDFs (fp64 accumulation, one rounding to fp32 per entry) ----
        if constexpr (MODE == MODE_SPECTRA) {
            double tu = 0.0, tv = 0.0;
#pragma unroll
            for (int c = 0; c < E; ++c) {
                tu += static_cast<double>(square ? xu[c] * xu[c] : xu[c]);
                tv += static_cast<double>(square ? xv[c] * xv[c] : xv[c]);
            }
            double off_u = tu, off_v = tv, tot_u, tot_v;
            group_scan2<TPF, false>(off_u, off_v, tot_u, tot_v, scratch, tid, g);
            const float mass_u = static_cast<float>(tot_u), mass_v = static_cast<float>(tot_v);
            finite = (fabsf(mass_u) <= FLT_BIG) && (fabsf(mass_v) <= FLT_BIG);  // false for NaN too
            u_live = mass_u > SAFE_EPS;  // utils.py:137: den <= eps -> eps
            v_live = mass_v > SAFE_EPS;
            inv_u = 1.0 / static_cast<double>(u_live ? mass_u : SAFE_EPS);
            inv_v = cut_scale ? inv_u : 1.0 / static_cast<double>(v_live ? mass_v : SAFE_EPS);
            if (args.flags & FLAG_RAW) {  // no normalisation, hence no mass term in the gradient
                inv_u = inv_v = 1.0;
                u_live = v_live = false;
            }
            double run_u = off_u, run_v = off_v;
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const int e = e0 + c;
                run_u += static_cast<double>(square ? xu[c] * xu[c] : xu[c]);
                run_v += static_cast<double>(square ? xv[c] * xv[c] : xv[c]);
                if (e < n) cu[e] = fminf(static_cast<float>(run_u * inv_u), FLT_BIG);
                if (e < m) cv[e] = fminf(static_cast<float>(run_v * inv_v), FLT_BIG);
            }
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const int e = e0 + c;
                if (e < n) cu[e] = fminf(xu[c], FLT_BIG);
                if (e < m) cv[e] = fminf(xv[c], FLT_BIG);
            }
        }
        if (tid == 0) {
            cu[n] = f_inf();
            cv[m] = f_inf();
        }
        group_sync<TPF>(g);

        // ---- stage 3: merge-path partition, then walk my L consecutive merged slots ----------
        const int L = (K + TPF - 1) / TPF;
        const int k0 = min(tid * L, K), k1 = min(k0 + L, K);
        const int cnt = k1 - k0;
        float acc = 0.0f;
        int fix = NO_FIX;         // zoneG row offset whose dL/dCDF still lacks the inherited m*d
        float carry_out = -1.0f;  // -1 = "my whole range continues a group opened before me"
        if constexpr (OUT == OUT_PLAN) {
            // Transport-plan emitter.  Same partition and tie-group rules as the gradient walk, but
            // the per-group payload is the pair of lower-bound indices (#{cu < q}, #{cv < q}) -- what
            // `searchsorted(cu, qs)` / `searchsorted(cv, qs)` return (losses.py:219) -- instead of m*d.
            int* const icarry = reinterpret_cast<int*>(carry);
            int n_inherited = 0, mine = -1;
            if (args.plan_cu != nullptr)
                for (int e = tid; e < n; e += TPF) args.plan_cu[frame * n + e] = cu[e];
            if (args.plan_cv != nullptr)
                for (int e = tid; e < m; e += TPF) args.plan_cv[frame * m + e] = cv[e];
            if (cnt > 0) {
                int i = merge_path(cu, cv, n, m, k0);
                int j = k0 - i;
                float qprev = (k0 == 0) ? 0.0f : fmaxf(i > 0 ? cu[i - 1] : -f_inf(), j > 0 ? cv[j - 1] : -f_inf());
                int is = 0, js = 0;
                bool inherited = (k0 != 0);
                for (int s = 0; s < cnt; ++s) {
                    const float a = cu[i], b = cv[j];
                    const bool take_v = b < a;
                    const float q = take_v ? b : a;
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
                    if (args.plan_uq != nullptr) args.plan_uq[o] = pu[is];
                    if (args.plan_vq != nullptr) args.plan_vq[o] = pv[js];
                    const bool masked = limit && (q > 1.0f);
                    const float dq = masked ? 0.0f : (q - qprev);
                    acc = fmaf(dq, transport_cost<PMODE>(pu[i], pv[j], args.p), acc);
                    qprev = q;
                    if (take_v) ++j; else ++i;
                }
                mine = inherited ? -1 : ((is << 16) | js);
            }
            icarry[tid] = mine;
            group_sync<TPF>(g);
            if (n_inherited > 0) {
                int s = tid - 1;
                int c = icarry[s];
                while (c < 0) c = icarry[--s];
                const int is = c >> 16, js = c & 0xffff;
                for (int t = 0; t < n_inherited; ++t) {
                    const long long o = frame * K + k0 + t;
                    if (args.plan_iu != nullptr) args.plan_iu[o] = is;
                    if (args.plan_iv != nullptr) args.plan_iv[o] = js;
                    if (args.plan_uq != nullptr) args.plan_uq[o] = pu[is];
                    if (args.plan_vq != nullptr) args.plan_vq[o] = pv[js];
                }
            }
        } else {
        if (cnt > 0) {
            int i = merge_path(cu, cv, n, m, k0);
            int j = k0 - i;
            float a = cu[i], b = cv[j], pa = pu[i], pb = pv[j];
            float qprev, md_prev;
            bool inherited;
            if (k0 == 0) {
                qprev = 0.0f;  // the zero the reference pads in front of qs (losses.py:301)
                md_prev = transport_cost<PMODE>(pa, pb, args.p);
                inherited = false;
            } else {
                qprev = fmaxf(i > 0 ? cu[i - 1] : -f_inf(), j > 0 ? cv[j - 1] : -f_inf());
                md_prev = 0.0f;
                inherited = true;
            }
            int src_prev = 0;
#pragma unroll 4
            for (int s = 0; s < cnt; ++s) {
                const bool take_v = b < a;  // equal values: the u entry first (stable cat order)
                const float q = take_v ? b : a;
                const float fresh = transport_cost<PMODE>(pa, pb, args.p);
                const bool masked = limit && (q > 1.0f);  // strict, losses.py:307
                float dq = q - qprev;
                dq = masked ? 0.0f : dq;
                acc = fmaf(dq, fresh, acc);
                if constexpr (WITH_GRAD) {
                    const bool same = (q == qprev);
                    const float md = same ? md_prev : (masked ? 0.0f : fresh);
                    if (s > 0) {  // dL/dCDF of the previous slot: nonzero only where its tie group ends
                        float* const dst = (src_prev < 0) ? (gzv + ~src_prev) : (gzu + src_prev);
                        *dst = same ? 0.0f : (md_prev - md);
                        if (!same && inherited) fix = src_prev;
                    }
                    inherited = inherited && same;
                    md_prev = md;
                    src_prev = take_v ? ~j : i;
                }
                qprev = q;
                if (take_v) {
                    ++j;
                    b = cv[j];
                    pb = pv[j];
                } else {
                    ++i;
                    a = cu[i];
                    pa = pu[i];
                }
            }
            if constexpr (WITH_GRAD) {
                // peek at slot k1 (first slot of the next thread, or the virtual end slot with m*d = 0)
                bool same = false;
                float md = 0.0f;
                if (k1 < K) {
                    const float q = fminf(a, b);
                    same = (q == qprev);
                    const bool masked = limit && (q > 1.0f);
                    md = same ? md_prev : (masked ? 0.0f : transport_cost<PMODE>(pa, pb, args.p));
                }
                float* const dst = (src_prev < 0) ? (gzv + ~src_prev) : (gzu + src_prev);
                *dst = same ? 0.0f : (md_prev - md);
                if (!same && inherited) fix = src_prev;
                inherited = inherited && same;
                carry_out = inherited ? -1.0f : md_prev;
            }
        }

        }  // OUT != OUT_PLAN

        if constexpr (WITH_GRAD) {
            // look-back: add the m*d of a tie group that was opened by an earlier thread
            carry[tid] = carry_out;
            group_sync<TPF>(g);
            if (fix != NO_FIX) {
                int s = tid - 1;
                float c = carry[s];
                while (c < 0.0f) c = carry[--s];  // thread 0 never carries the marker
                float* const dst = (fix < 0) ? (gzv + ~fix) : (gzu + fix);
                *dst += c;
            }
            group_sync<TPF>(g);
        }

        // ---- loss of the frame ------------------------------------------------------------
        {
            double part = static_cast<double>(acc), unused = 0.0;
            group_sum2<TPF>(part, unused, scratch + 2 * (TPF / 32), tid, g);
            if (tid == 0 && args.loss != nullptr)
                args.loss[frame] = finite ? static_cast<float>(part) : f_nan();
        }

        if constexpr (WITH_GRAD) {
            if constexpr (MODE == MODE_SPECTRA) {
                // ---- stage 4: cumsum transpose = suffix sums of dL/dCDF, fp64 accumulation --------
                double su = 0.0, sv = 0.0;
#pragma unroll
                for (int c = E - 1; c >= 0; --c) {
                    const int e = e0 + c;
                    if (e < n) su += static_cast<double>(gzu[e]);
                    if (e < m) sv += static_cast<double>(gzv[e]);
                }
                double off_u = su, off_v = sv, tot_u, tot_v;
                group_scan2<TPF, true>(off_u, off_v, tot_u, tot_v, scratch + 4 * (TPF / 32), tid, g);
                // suffix sums gw, and the two dot products <gw, a> the normalisation needs
                float gwu[E], gwv[E];  // fp32 copies of the fp64 running suffix sums
                double dot_u = 0.0, dot_v = 0.0;
                double run_u = off_u, run_v = off_v;
#pragma unroll
                for (int c = E - 1; c >= 0; --c) {
                    const int e = e0 + c;
                    if (e < n) run_u += static_cast<double>(gzu[e]);
                    if (e < m) run_v += static_cast<double>(gzv[e]);
                    gwu[c] = static_cast<float>(run_u);
                    gwv[c] = static_cast<float>(run_v);
                    dot_u += run_u * static_cast<double>(square ? xu[c] * xu[c] : xu[c]);
                    dot_v += run_v * static_cast<double>(square ? xv[c] * xv[c] : xv[c]);
                }
                group_sum2<TPF>(dot_u, dot_v, scratch, tid, g);  // slot 0 is free again (see stage 2)
                // w = a * inv  =>  sum_i gw_i w_i = dot * inv ; a clamped mass has no derivative
                const float corr_u = u_live ? static_cast<float>((cut_scale ? dot_u + dot_v : dot_u) * inv_u) : 0.0f;
                const float corr_v = (!cut_scale && v_live) ? static_cast<float>(dot_v * inv_v) : 0.0f;
                const float up = args.upstream != nullptr ? args.upstream[frame] : 1.0f;
                const float ku = static_cast<float>(inv_u) * up, kv = static_cast<float>(inv_v) * up;
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    const int e = e0 + c;
                    float ga = (gwu[c] - corr_u) * ku;
                    float gb = (gwv[c] - corr_v) * kv;
                    if (square) {
                        ga *= 2.0f * xu[c];
                        gb *= 2.0f * xv[c];
                    }
                    if (!finite) ga = gb = f_nan();
                    if (e < n) gzu[e] = ga;  // same thread read gzu[e] above: in place is safe
                    if (e < m) gzv[e] = gb;
                }
            }
        }
    }

    // ---- stage 5: write the gradient rows (4 rows back to back -> one bulk store per side) ----
    if constexpr (WITH_GRAD) {
        float* const ou = args.grad_u != nullptr ? args.grad_u + frame0 * n : nullptr;
        float* const ov = args.grad_v != nullptr ? args.grad_v + frame0 * m : nullptr;
        const bool bulk_out = rows16 && (nfr == FPC) && ((reinterpret_cast<uintptr_t>(ou) & 15) == 0) &&
                              ((reinterpret_cast<uintptr_t>(ov) & 15) == 0);
        if (bulk_out) fence_async_smem();
        __syncthreads();
        if (bulk_out) {
            if (threadIdx.x == 0) {
                if (ou != nullptr) bulk_s2g(ou, zoneG, 4u * FPC * n);
                if (ov != nullptr) bulk_s2g(ov, zoneG + FPC * n, 4u * FPC * m);
                bulk_commit();
                bulk_wait_read_all();
            }
        } else {
            if (ou != nullptr)
                for (int idx = threadIdx.x; idx < nfr * n; idx += FPC * TPF) ou[idx] = zoneG[idx];
            if (ov != nullptr)
                for (int idx = threadIdx.x; idx < nfr * m; idx += FPC * TPF) ov[idx] = zoneG[FPC * n + idx];
        }
    }
}

}  // namespace sot
