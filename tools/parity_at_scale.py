"""Parity at the paper's batch size: 1024 frames (64 signals x 16) per configuration.  The CUDA path and the
reference's own float32 evaluation (CPU oracle = bit-exact restatement of the reference) are BOTH compared with the
float64 evaluation of the same formula, for the per-frame loss and for the gradients of the batch mean.
Prints one JSON line per configuration (DESIGN.md section 2 quotes them)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sot_oracle as O  # noqa: E402  (the checker, not the product)
from sot_b200 import losses, synthetic as S  # noqa: E402


def rel_l2(a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double())).item()


def main():
    dev = "cuda:0"
    for name, n_fft, cut, grid in (("SOT-2048", 2048, True, "linear"), ("SOT-512", 512, True, "linear"),
                                   ("SOT-512-LogF", 512, True, "logf"), ("SOT-NoCut", 2048, False, "linear")):
        x, y = S.sot_batch(64, n_fft, seed=42)
        pos = S.linear_positions(n_fft) if grid == "linear" else S.logf_positions(n_fft)
        kw = dict(p=2, square=True, cut_scale=cut, limit=cut)
        out = {}
        for tag, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
            xr, yr = x.detach().clone().to(dt).requires_grad_(True), y.detach().clone().to(dt).requires_grad_(True)
            rows = O.sot_per_frame(xr, yr, pos.to(dt), pos.to(dt), stable=True, **kw)
            rows.mean().backward()
            out[tag] = (rows.detach(), xr.grad, yr.grad)
        xg, yg = x.detach().to(dev).requires_grad_(True), y.detach().to(dev).requires_grad_(True)
        rows = losses.sot_frames(xg, yg, pos.to(dev), pos.to(dev), **kw)
        rows.mean().backward()
        out["cuda"] = (rows.detach().cpu(), xg.grad.cpu(), yg.grad.cpu())
        truth = out["ref64"]
        line = {"config": name, "frames": rows.numel(), "mean_loss": truth[0].mean().item(),
                "mean_loss_rel_diff_cuda_vs_ref32": abs(out["cuda"][0].double().mean().item() - out["ref32"][0].double().mean().item())
                / out["ref32"][0].double().mean().item()}
        for tag in ("cuda", "ref32"):
            err = (out[tag][0].double() - truth[0]).abs()
            line[tag] = {"loss_abs_err_vs_fp64_mean": err.mean().item(), "loss_abs_err_vs_fp64_max": err.max().item(),
                         "grad_x_rel_l2_vs_fp64": rel_l2(out[tag][1], truth[1]),
                         "grad_y_rel_l2_vs_fp64": rel_l2(out[tag][2], truth[2])}
        closer = ((out["cuda"][0].double() - truth[0]).abs() <= (out["ref32"][0].double() - truth[0]).abs()).sum().item()
        line["frames_where_cuda_is_at_least_as_close_to_fp64_as_ref32"] = int(closer)
        print(json.dumps(line))


if __name__ == "__main__":
    main()
