"""Row f1 measurement: the SOT step fed from complex STFT frames, with |z| fused into the kernels
versus materialised by torch (`.abs()` forward + its backward), and the whole audio -> loss -> audio-gradient
wrapper step (`features.Wasserstein1DWithTransform`).  CUDA events, inputs larger than L2.

    python tools/bench_stft_prologue.py [--signals 2048] [--n-fft 2048]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sot_b200 import features, losses, synthetic as S  # noqa: E402


def timed(fn, steps, warmup=5):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--signals", type=int, default=2048)
    ap.add_argument("--n-fft", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(1)
    tgt, f0 = S.harmonic_signals(args.signals, gen, device=dev)
    pred, _ = S.harmonic_signals(args.signals, gen, n_partials=20, f0=f0 * 1.05, peak_normalise=False, amp_lo=0.05,
                                 amp_hi=0.5, device=dev)
    kw = dict(p=2, square_dist=True)
    tkw = dict(type="stft", n_fft=args.n_fft, hop_length=256, window="flattop")
    wrapper = features.Wasserstein1DWithTransform(transform_kwargs=dict(tkw), **kw).to(dev)
    loss_fn = losses.Wasserstein1D(**kw)
    zx = wrapper.transform.complex_frames(tgt).detach().requires_grad_(True)
    zy = wrapper.transform.complex_frames(pred).detach().requires_grad_(True)
    pos = wrapper._positions(dev)
    frames = zx.shape[0] * zx.shape[1]

    def fused_step():
        zx.grad = zy.grad = None
        loss_fn(zx, zy, x_pos=pos, y_pos=pos).backward()

    def abs_step():
        zx.grad = zy.grad = None
        loss_fn(zx.abs(), zy.abs(), x_pos=pos, y_pos=pos).backward()

    xa, ya = tgt.clone().requires_grad_(True), pred.clone().requires_grad_(True)

    def wrapper_step():
        xa.grad = ya.grad = None
        wrapper(xa, ya).backward()

    def wrapper_abs_step():  # the reference's structure: transform() returns magnitudes
        xa.grad = ya.grad = None
        loss_fn(wrapper.transform(xa), wrapper.transform(ya), x_pos=pos, y_pos=pos).backward()

    out = {"frames": frames, "bins": zx.shape[-1]}
    for name, fn in (("frames_fused_abs", fused_step), ("frames_torch_abs", abs_step),
                     ("audio_fused_abs", wrapper_step), ("audio_torch_abs", wrapper_abs_step)):
        ms = timed(fn, args.steps)
        out[name] = {"ms_per_step": round(ms, 4), "frames_per_s": round(frames / ms * 1e3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
