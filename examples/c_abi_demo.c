/* Plain-C caller of the SOT C ABI (include/sot_b200.h): host buffers in, per-frame loss and both gradients out,
 * no Python and no torch anywhere.  What a non-Python host (or the reference behind a cgo / JNI / ctypes stub,
 * INTEGRATION.md) links against.
 *
 *   gcc -std=c99 -O2 examples/c_abi_demo.c -Iinclude -L1d-spectral-optimal-transport_b200/_lib -lsot_b200 \
 *       -Wl,-rpath,$PWD/1d-spectral-optimal-transport_b200/_lib -lm -o /tmp/sot_c_demo && /tmp/sot_c_demo
 *
 * Prints the mean loss of 64 synthetic frames of 1025 bins (two shifted Gaussian bumps per frame: the squared
 * spectra are Gaussians whose centres differ by 0.05, so W_2^2 ~ 0.0025) and a checksum of the gradients. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "sot_b200.h"

int main(void) {
    const int n_frames = 64, n_bins = 1025;
    float* x = malloc(sizeof(float) * n_frames * n_bins);
    float* y = malloc(sizeof(float) * n_frames * n_bins);
    float* gx = malloc(sizeof(float) * n_frames * n_bins);
    float* gy = malloc(sizeof(float) * n_frames * n_bins);
    float* pos = malloc(sizeof(float) * n_bins);
    float* loss = malloc(sizeof(float) * n_frames);
    if (!x || !y || !gx || !gy || !pos || !loss) return 2;
    for (int i = 0; i < n_bins; ++i) pos[i] = (float)i / (float)(n_bins - 1);
    for (int f = 0; f < n_frames; ++f) {
        const double cx = 0.30 + 0.004 * f, cy = cx + 0.05, w = 0.02;
        for (int i = 0; i < n_bins; ++i) {
            const double p = pos[i];
            x[f * n_bins + i] = (float)exp(-0.25 * (p - cx) * (p - cx) / (w * w)); /* magnitude: its square is N(cx, w) */
            y[f * n_bins + i] = (float)(0.7 * exp(-0.25 * (p - cy) * (p - cy) / (w * w)));
        }
    }
    sot_problem prob;
    prob.n_frames = n_frames;
    prob.n_u = n_bins;
    prob.n_v = n_bins;
    prob.u = x;
    prob.v = y;
    prob.pos_u = pos;
    prob.pos_v = pos;
    prob.pos_u_stride = 0; /* one support row shared by all frames */
    prob.pos_v_stride = 0;
    prob.p = 2.0f;
    prob.flags = SOT_SQUARE; /* squared magnitudes, both spectra normalised to mass 1 */
    const int rc = sot_loss_grad_host(&prob, NULL, loss, gx, gy, 0);
    if (rc != SOT_OK) {
        fprintf(stderr, "sot_loss_grad_host failed (%d): %s\n", rc, sot_last_error());
        return 1;
    }
    double mean = 0.0, checksum = 0.0;
    for (int f = 0; f < n_frames; ++f) mean += loss[f];
    for (int i = 0; i < n_frames * n_bins; ++i) checksum += fabs((double)gx[i]) + fabs((double)gy[i]);
    printf("{\"abi\": %d, \"frames\": %d, \"mean_loss\": %.9g, \"grad_abs_sum\": %.9g, \"launches\": %lld}\n",
           sot_abi_version(), n_frames, mean / n_frames, checksum, (long long)sot_launch_count());
    sot_host_release(0);
    free(x), free(y), free(gx), free(gy), free(pos), free(loss);
    return 0;
}
