"""The STFT front end of the SOT loss, kept complex so that `|.|` runs inside the SOT kernel.

Reference: `features.py:191-237` (`stft`, `compute_mag`), `features.py:85-113` (`TorchSTFT`),
`utils.py:252-274` (`pad_for_stft`), `losses.py:316-343` (`Wasserstein1DWithTransform`).

The reference materialises `stft(x).abs()` (one pass over the complex spectrum forward, one
backward) and permutes it to (batch, time, freq) before the loss.  `torch.stft` already lays its
result out frame-major in memory -- its (batch, freq, time) return value is a transposed view --
so `complex_frames` hands those rows to the kernels without a copy, and the magnitude, its square
and their derivatives are fused into the two SOT launches (flag SOT_COMPLEX_INPUT,
`include/sot_b200.h`).  The FFT itself stays cuFFT through `torch.stft`: it is a library
transform, not part of the hot path this repository replaces.
"""
from __future__ import annotations

import functools

import torch

from .losses import Wasserstein1D

__all__ = ["pad_for_stft", "stft", "complex_frames", "compute_mag", "TorchSTFT", "get_transform",
           "Wasserstein1DWithTransform"]


def pad_for_stft(signal: torch.Tensor, frame_size: int, hop_length: int) -> torch.Tensor:
    """Zero-pad the end so the window slides until it is completely beyond the signal
    (utils.py:252-274): ceil(len / hop) frames."""
    length = signal.shape[1]
    frames = -(-length // hop_length)
    pad = max(0, frame_size + hop_length * (frames - 1) - length)
    return signal if pad == 0 else torch.nn.functional.pad(signal, (0, pad))


def _window(window, frame_size: int, device) -> torch.Tensor:
    if window is None:
        return torch.hann_window(int(frame_size), device=device)
    return torch.as_tensor(window, dtype=torch.float32).to(device)


def stft(audio, frame_size=2048, overlap=0.75, center=False, pad_end=True, window=None) -> torch.Tensor:
    """(batch, samples) -> complex64 (batch, freq, time), same arguments as features.py:191-214."""
    audio = torch.as_tensor(audio, dtype=torch.float32)
    hop_length = int(frame_size * (1.0 - overlap))
    if pad_end:
        audio = pad_for_stft(audio, frame_size, hop_length)
    assert frame_size * overlap % 2.0 == 0.0
    return torch.stft(audio, n_fft=int(frame_size), hop_length=hop_length, win_length=int(frame_size),
                      window=_window(window, frame_size, audio.device), center=center, normalized=True,
                      return_complex=True)


def complex_frames(audio, size=2048, overlap=0.75, pad_end=True, center=False, window=None) -> torch.Tensor:
    """(batch, samples) -> complex64 (batch, time, freq): `compute_mag(...).permute(0, 2, 1)` without the
    `.abs()`.  No copy when `torch.stft` returns its usual transposed view."""
    spec = stft(audio, frame_size=size, overlap=overlap, center=center, pad_end=pad_end, window=window)
    return spec.transpose(1, 2).contiguous()


def compute_mag(audio, size=2048, overlap=0.75, pad_end=True, center=False, add_in_sqrt=0.0, window=None):
    """Magnitude spectrogram (batch, freq, time) as in features.py:217-237 (`add_in_sqrt` is unused there too)."""
    return stft(audio, frame_size=size, overlap=overlap, center=center, pad_end=pad_end, window=window).abs().float()


class TorchSTFT(torch.nn.Module):
    """features.py:85-113.  `forward` returns magnitudes (batch, time, freq) like the reference;
    `complex_frames` returns the same frames before `.abs()` for the fused loss."""

    def __init__(self, **kwargs):
        super().__init__()
        self.n_fft = kwargs.pop("n_fft", 1024)
        hop_length = kwargs.pop("hop_length", 256)
        self.sr = kwargs.pop("sr", 16000)
        self.log = kwargs.pop("log", False)
        window = kwargs.pop("window", None)
        if window is not None and not torch.is_tensor(window):
            from scipy.signal import get_window  # features.py:95
            window = get_window(window, self.n_fft)
        self._stft_args = dict(size=self.n_fft, overlap=1 - hop_length / self.n_fft, window=window)
        self.transform = functools.partial(compute_mag, **self._stft_args)

    def complex_frames(self, x) -> torch.Tensor:
        return complex_frames(x, **self._stft_args)

    def forward(self, x, **kwargs):
        x = self.transform(x).permute(0, 2, 1)
        if kwargs.get("reduce", False):
            x = x.mean(dim=1)
        if kwargs.get("log", False) or self.log:
            x = torch.log(torch.clamp(x, min=1e-5))  # safe_log (utils.py:145-151): values <= eps become eps
        return x

    def get_frequencies(self):
        return torch.fft.rfftfreq(self.n_fft, d=1 / self.sr)


def get_transform(transform, sample_rate):
    """features.py:33-61.  "stft" and "identity" are covered; the CQT front end (nnAudio) is outside
    the SOT hot path (out of scope) and is refused by name."""
    if isinstance(transform, dict):
        kwargs = dict(transform)
        name = kwargs.pop("type")
    else:
        name, kwargs = transform, None
    if name == "stft":
        if kwargs is None:
            kwargs = dict(log=False, n_fft=1024, hop_length=256)
        for reference_only in ("center", "output_format"):  # accepted and ignored by the reference's TorchSTFT
            kwargs.pop(reference_only, None)
        kwargs.update({"sr": sample_rate})
        return TorchSTFT(**kwargs)
    if name == "identity":
        return torch.nn.Identity()
    if name == "cqt":
        raise NotImplementedError("sot_b200: the CQT transform is not part of the B200 hot path; compute it with "
                                  "the reference's features.CQT and pass the result to Wasserstein1D")
    raise ValueError(f"Unknown transform {name}")


class Wasserstein1DWithTransform(torch.nn.Module):
    """losses.py:316-343: transform both signals, then the SOT loss on linear positions in [0, 1].

    With the STFT transform (and no log compression) the complex frames go straight into the SOT
    kernels; every other transform takes the reference's route (transform, then `Wasserstein1D`)."""

    def __init__(self, p=1, fixed_x=None, require_sort=True, log_scaled_x=False, transform_kwargs=None, **kwargs):
        super().__init__()
        self.wasserstein = Wasserstein1D(p=p, fixed_x=fixed_x, require_sort=require_sort,
                                         log_scaled_x=log_scaled_x, **kwargs)
        self.transform = get_transform(transform_kwargs, sample_rate=transform_kwargs.pop("sr", 16000))
        self._pos_cache = None

    def _positions(self, device):
        if self._pos_cache is None or self._pos_cache.device != device:
            pos = torch.as_tensor(self.transform.get_frequencies()).to(device)
            self._pos_cache = pos / pos.max()
        return self._pos_cache

    def forward(self, x, y, **kwargs):
        fused = isinstance(self.transform, TorchSTFT) and not self.transform.log
        if fused:
            x, y = self.transform.complex_frames(x), self.transform.complex_frames(y)
        else:
            x, y = self.transform(x), self.transform(y)
        pos = self._positions(x.device)
        return self.wasserstein(x, y, x_pos=pos, y_pos=pos, **kwargs)
