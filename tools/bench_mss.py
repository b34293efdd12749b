"""Row f3 measurement: one MSS term (L1 of magnitudes, the paper's setting) on complex spectrograms, forward +
backward w.r.t. the prediction, fused kernels (csrc/sot_mss.cu) versus the reference's chain of torch ops.
CUDA events, 65 536 x 1025 complex bins (0.5 GB per spectrogram, larger than L2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sot_b200 import mss  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / steps


def main():
    dev = torch.device("cuda", 0)
    n = 65536 * 1025
    zt = torch.view_as_complex(torch.randn(n, 2, device=dev))
    zv = torch.view_as_complex(torch.randn(n, 2, device=dev)).requires_grad_(True)

    def fused():
        zv.grad = None
        mss.mss_term(zt, zv, 1.0, 0.0, "L1").backward()

    def eager():
        zv.grad = None
        (zt.abs() - zv.abs()).abs().mean().backward()

    out = {"elements": n}
    for name, fn, nbytes in (("fused", fused, 16 * n + 24 * n), ("torch_ops", eager, None)):
        ms = timed(fn)
        out[name] = {"ms": round(ms, 4)}
        if nbytes:
            out[name]["algorithmic_gbs"] = round(nbytes / ms / 1e6, 1)  # fwd reads 16 B; bwd reads 16 B, writes 8 B
    print(json.dumps(out))


if __name__ == "__main__":
    main()
