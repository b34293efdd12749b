// Per-configuration launcher: one translation unit per (TPF, E, FPC) so the configs compile in
// parallel.  Each exposes `sot_launch_<TPF>_<E>_<FPC>(args, out, mode, stream)`.
#pragma once
#include "sot_kernels.cuh"

namespace sot {

struct LaunchRequest {
    FrameArgs args;
    int out;   // OUT_LOSS / OUT_GRAD / OUT_PLAN
    int mode;  // MODE_SPECTRA / MODE_CDF
};

template <int TPF, int E, int FPC, int PMODE, int OUT, int MODE>
cudaError_t launch_one(const FrameArgs& a, cudaStream_t stream) {
    auto kernel = sot_frames_kernel<TPF, E, FPC, PMODE, OUT, MODE>;
    const SmemPlan plan = smem_plan(FPC, a.n, a.m, TPF, a.pos_u_stride == 0 && a.pos_v_stride == 0);
    // per instantiation (and device): opt-in shared memory size and the persistent grid size
    static int configured_bytes = -1, cached_bytes = -1, cached_grid = 0, cached_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (plan.total > configured_bytes || dev != cached_dev) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.total);
        if (e != cudaSuccess) return e;
        configured_bytes = plan.total;
    }
    if (plan.total != cached_bytes || dev != cached_dev) {
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FPC * TPF, plan.total);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        cached_grid = (per_sm < 1 ? 1 : per_sm) * sms;  // persistent: every CTA slot of the chip, once
        cached_bytes = plan.total;
        cached_dev = dev;
    }
    const long long quads = (a.n_frames + FPC - 1) / FPC;
    const unsigned grid = static_cast<unsigned>(quads < cached_grid ? quads : cached_grid);
    kernel<<<grid, FPC * TPF, plan.total, stream>>>(a);
    return cudaGetLastError();
}

template <int TPF, int E, int FPC>
cudaError_t launch_config(const LaunchRequest& r, cudaStream_t stream) {
    const bool p2 = (r.args.p == 2.0f);
    if (r.mode == MODE_SPECTRA) {
        if (r.out == OUT_LOSS)
            return p2 ? launch_one<TPF, E, FPC, 2, OUT_LOSS, MODE_SPECTRA>(r.args, stream)
                      : launch_one<TPF, E, FPC, 0, OUT_LOSS, MODE_SPECTRA>(r.args, stream);
        if (r.out == OUT_GRAD)
            return p2 ? launch_one<TPF, E, FPC, 2, OUT_GRAD, MODE_SPECTRA>(r.args, stream)
                      : launch_one<TPF, E, FPC, 0, OUT_GRAD, MODE_SPECTRA>(r.args, stream);
        return launch_one<TPF, E, FPC, 0, OUT_PLAN, MODE_SPECTRA>(r.args, stream);
    }
    if (r.out == OUT_LOSS) return launch_one<TPF, E, FPC, 0, OUT_LOSS, MODE_CDF>(r.args, stream);
    if (r.out == OUT_GRAD) return launch_one<TPF, E, FPC, 0, OUT_GRAD, MODE_CDF>(r.args, stream);
    return launch_one<TPF, E, FPC, 0, OUT_PLAN, MODE_CDF>(r.args, stream);
}

}  // namespace sot

#define SOT_DEFINE_CONFIG(TPF, E, FPC)                                                              \
    cudaError_t sot_launch_##TPF##_##E##_##FPC(const sot::LaunchRequest& r, cudaStream_t stream) { \
        return sot::launch_config<TPF, E, FPC>(r, stream);                                          \
    }
#define SOT_DECLARE_CONFIG(TPF, E, FPC) \
    cudaError_t sot_launch_##TPF##_##E##_##FPC(const sot::LaunchRequest& r, cudaStream_t stream);
