"""`metrics.wasserstein_distance` of the reference (metrics.py:144-149) on the CUDA path: hann STFT of both
signals, `Wasserstein1D(p, fixed_x=n_bins)` on the linspace grid; the complex frames go straight into the kernel."""
from __future__ import annotations

import torch

from .features import complex_frames
from .losses import Wasserstein1D

__all__ = ["wasserstein_distance"]


@torch.inference_mode()
def wasserstein_distance(x, x_hat, p=1, n_fft=512):
    zx = complex_frames(x, size=n_fft)
    zx_hat = complex_frames(x_hat, size=n_fft)
    return Wasserstein1D(p=p, fixed_x=zx.shape[-1]).to(x.device)(zx, zx_hat)
