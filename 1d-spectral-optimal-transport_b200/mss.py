"""Multi-scale spectral loss (`MSSLoss`) and `MixOfLosses`, the other half of every SOT paper config
(reference: losses.py:345-425, `mean_difference` :7-36, weights 0.05 / 1 in the YAMLs).

Same constructor and `forward(target_audio, audio, **kwargs)` as the reference.  Per FFT size the two
complex spectrograms go straight from `torch.stft` into ONE reduction kernel (forward) and ONE gradient
kernel (backward) -- `csrc/sot_mss.cu`; the reference materialises `abs()`, the difference, `abs()` again
and the mean (and `safe_log` of both for the log term) as separate tensors, forward and backward.
"""
from __future__ import annotations

import torch

from . import _capi
from .features import stft

__all__ = ["MSSLoss", "MixOfLosses", "mss_term"]


class _MssTerm(torch.autograd.Function):
    """(zt, zv) complex64 spectrograms of one FFT size -> mean-difference term, a 0-dim float32 tensor."""

    @staticmethod
    def forward(ctx, zt, zv, mag_weight, logmag_weight, loss_type):
        ctx.cfg = (float(mag_weight), float(logmag_weight), int(loss_type), 1.0 / max(zt.numel(), 1))
        ctx.need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        ctx.save_for_backward(zt, zv)
        total = _capi.mss_forward(zt, zv, *ctx.cfg)
        return total[0].to(torch.float32)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        zt, zv = ctx.saved_tensors
        scale = grad_out.to(torch.float32).reshape(1).contiguous()
        gt, gv = _capi.mss_backward(zt, zv, *ctx.cfg, scale, ctx.need[0], ctx.need[1])
        return gt, gv, None, None, None


def _is_dense(z: torch.Tensor) -> bool:
    """Some permutation of the dims is contiguous (e.g. torch.stft's transposed view): every element of the
    storage span is visited exactly once, so the elementwise kernels can walk the memory linearly."""
    expect = 1
    for size, stride in sorted(((sz, st) for sz, st in zip(z.shape, z.stride()) if sz > 1), key=lambda t: t[1]):
        if stride != expect:
            return False
        expect *= size
    return True


def _dense(z: torch.Tensor) -> torch.Tensor:
    return z if _is_dense(z) else z.contiguous()


def mss_term(zt, zv, mag_weight=1.0, logmag_weight=0.0, loss_type="L1") -> torch.Tensor:
    """mag_weight * mean d(|zt|, |zv|) + logmag_weight * mean d(safe_log|zt|, safe_log|zv|) for one pair of
    complex spectrograms of equal shape (d = L1 or L2 difference, losses.py:7-36)."""
    kind = loss_type.upper()
    if kind not in ("L1", "L2"):
        raise ValueError('Loss type ({}), must be "L1", "L2" '.format(kind))  # losses.py:36
    zt, zv = _dense(zt), _dense(zv)
    if zt.stride() != zv.stride():
        zv = zv.contiguous()
        zt = zt.contiguous()
    return _MssTerm.apply(zt, zv, mag_weight, logmag_weight, _capi.SOT_MSS_L1 if kind == "L1" else _capi.SOT_MSS_L2)


def _partial_mean_term(zt, zv, mag_weight, logmag_weight, loss_type, dims):
    """`mean_difference(..., dims=dims)` (losses.py:7-36) for a caller that wants per-item values (evaluation):
    a partial mean has no single scalar to reduce into, so this rarely used form runs as torch reductions on the
    magnitudes, like the reference; the training form (dims=None) is the fused kernel pair."""
    kind = loss_type.upper()
    if kind not in ("L1", "L2"):
        raise ValueError('Loss type ({}), must be "L1", "L2" '.format(kind))  # losses.py:36
    tm, vm = zt.abs().float(), zv.abs().float()

    def mean_difference(a, b):
        d = a - b
        return torch.mean(d.abs() if kind == "L1" else d ** 2, dim=dims)

    out = 0.0
    if mag_weight > 0:
        out = out + mag_weight * mean_difference(tm, vm)
    if logmag_weight > 0:
        eps = torch.tensor(1e-5, device=tm.device)  # utils.py:145-151 `safe_log`
        out = out + logmag_weight * mean_difference(torch.log(torch.where(tm <= eps, eps, tm)),
                                                    torch.log(torch.where(vm <= eps, eps, vm)))
    return out


class MSSLoss(torch.nn.Module):
    """losses.py:365-425.  `dims=None` (every training config): one fused reduction kernel and one gradient kernel
    per FFT size; `dims=...` (a partial mean): torch reductions on the magnitudes, like the reference."""

    def __init__(self, fft_sizes=(2048, 1024, 512, 256, 128, 64), loss_type="L1", mag_weight=0.0, logmag_weight=0.0):
        super().__init__()
        self.fft_sizes = fft_sizes
        self.loss_type = loss_type
        self.mag_weight = mag_weight
        self.logmag_weight = logmag_weight

    def forward(self, target_audio, audio, **kwargs):
        dims = kwargs.get("dims", None)
        loss = 0.0
        if not (self.mag_weight > 0 or self.logmag_weight > 0):
            return loss  # the reference adds nothing either (losses.py:409, 415)
        for size in self.fft_sizes:
            zt = stft(target_audio, frame_size=size)  # compute_mag's defaults: overlap 0.75, hann, end padding
            zv = stft(audio, frame_size=size)
            if dims is None:
                loss = loss + mss_term(zt, zv, max(self.mag_weight, 0.0), max(self.logmag_weight, 0.0), self.loss_type)
            else:
                loss = loss + _partial_mean_term(zt, zv, self.mag_weight, self.logmag_weight, self.loss_type, dims)
        return loss


class MixOfLosses(torch.nn.Module):
    """losses.py:345-362: a plain list of losses and weights -> dict of weighted values keyed by class name."""

    def __init__(self, losses, weights=None):
        super().__init__()
        self.losses = losses
        self.weights = weights

    def forward(self, x, y, **kwargs):
        loss = {}
        for loss_fn, weight in zip(self.losses, self.weights):
            loss[loss_fn.__class__.__name__] = loss_fn(x, y, **kwargs) * weight
        return loss
