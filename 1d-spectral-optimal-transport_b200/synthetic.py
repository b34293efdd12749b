"""Synthetic harmonic spectra of the paper's shape -- inputs for tests and bench.py.

Not part of the hot path: this stands in for the reference's *upstream* stages (data generator,
DDSP synth, `torch.stft`) so that the loss sees realistic inputs.  The distribution follows the
reference's dataset generator (synthetic_data.py:76-122, 331-345: sr 16 kHz, 4096 samples,
f0 ~ U(40, 1950) Hz, 8 harmonics with amplitudes ~ U(0.4, 1), a random number (1..8) of leading
partials active, peak-normalised to 0.9 -- synthetic_data.py:235-237) and its STFT front end
(features.py:85-113, 191-237: flattop window, hop 256, `normalized=True`, `center=False`, padded
at the end so a 4096-sample signal yields 16 frames -- utils.py:252-275).  The prediction side
mimics the decoder: 20 partials (train_config.yaml:41), detuned f0, not peak-normalised.

`torch.stft` (cuFFT on a GPU) stays upstream of the loss by design (BASELINE.json north_star).
"""
from __future__ import annotations

import math

import torch

SAMPLE_RATE = 16000
N_SAMPLES = 4096
HOP = 256
PAPER_SEEDS = (42, 123, 456, 789, 101112)


def flattop_window(n: int, device=None) -> torch.Tensor:
    """Periodic ('fftbins') flat-top window, the 5-term cosine series scipy uses for 'flattop'."""
    a = (0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368)
    t = torch.arange(n, dtype=torch.float64, device=device) * (2.0 * math.pi / n)
    w = sum(((-1) ** k) * a[k] * torch.cos(k * t) for k in range(5))
    return w.to(torch.float32)


def harmonic_signals(n_signals, gen, *, n_partials=8, min_active=1, f0_lo=40.0, f0_hi=1950.0,
                     amp_lo=0.4, amp_hi=1.0, detune=1.0, peak_normalise=True, f0=None,
                     n_samples=N_SAMPLES, sr=SAMPLE_RATE, device="cpu"):
    """(n_signals, n_samples) float32 sums of harmonics below Nyquist; returns (audio, f0)."""
    if f0 is None:
        f0 = torch.rand(n_signals, 1, generator=gen) * (f0_hi - f0_lo) + f0_lo
    amps = torch.rand(n_signals, n_partials, generator=gen) * (amp_hi - amp_lo) + amp_lo
    n_active = torch.randint(min_active, n_partials + 1, (n_signals, 1), generator=gen)
    order = torch.arange(1, n_partials + 1).unsqueeze(0)
    amps = amps * (order <= n_active)
    f0, amps, order = f0.to(device), amps.to(device), order.to(device)
    freqs = f0 * detune * order  # (S, P)
    amps = amps * (freqs < sr / 2)
    t = torch.arange(n_samples, device=device, dtype=torch.float32) / sr
    audio = torch.zeros(n_signals, n_samples, device=device)
    for k in range(n_partials):  # partial by partial: bounded memory for big sweeps
        audio += amps[:, k:k + 1] * torch.sin(2 * math.pi * freqs[:, k:k + 1] * t)
    if peak_normalise:
        audio = audio / (audio.abs().amax(dim=1, keepdim=True) + 1e-7) * 0.9
    return audio, f0.cpu()


def stft_magnitude(audio: torch.Tensor, n_fft: int, hop: int = HOP) -> torch.Tensor:
    """(B, samples) -> (B, frames, n_fft//2+1) contiguous float32 magnitudes."""
    n = audio.shape[1]
    frames = -(-n // hop)
    pad = max(0, n_fft + hop * (frames - 1) - n)
    audio = torch.nn.functional.pad(audio, (0, pad))
    spec = torch.stft(audio, n_fft=n_fft, hop_length=hop, win_length=n_fft,
                      window=flattop_window(n_fft, audio.device), center=False, normalized=True,
                      return_complex=True)
    return spec.abs().permute(0, 2, 1).contiguous()


def sot_batch(n_signals: int, n_fft: int, seed: int = 42, device="cpu", chunk: int = 512):
    """Target and prediction magnitude spectra, each (n_signals, 16, n_fft//2+1)."""
    gen = torch.Generator().manual_seed(seed)
    xs, ys = [], []
    for lo in range(0, n_signals, chunk):
        b = min(chunk, n_signals - lo)
        tgt, f0 = harmonic_signals(b, gen, device=device)
        jitter = 1.0 + 0.2 * (torch.rand(b, 1, generator=gen) - 0.5)  # +-10 % detune
        pred, _ = harmonic_signals(b, gen, n_partials=20, f0=f0 * jitter, peak_normalise=False,
                                   amp_lo=0.05, amp_hi=0.5, device=device)
        xs.append(stft_magnitude(tgt, n_fft))
        ys.append(stft_magnitude(pred, n_fft))
    return torch.cat(xs), torch.cat(ys)


def linear_positions(n_fft: int, sr: int = SAMPLE_RATE) -> torch.Tensor:
    """Bin frequencies scaled linearly to [0, 1] (trainer.py:193-197)."""
    f = torch.fft.rfftfreq(n_fft, d=1 / sr)
    return f / f.max()


def logf_positions(n_fft: int, sr: int = SAMPLE_RATE, fmin: float = 32.7,
                   fmax: float = 32.7 * 2 ** (284 / 36)) -> torch.Tensor:
    """Log-frequency positions: MIDI pitch of each bin mapped affinely so [fmin, fmax] -> [0, 1]
    (trainer.py:187-191, utils.py:54-106; a 0 Hz bin is given MIDI 0).  Not monotone at
    n_fft=2048 (bin 0 lands above bin 1), which exercises the unsorted-support path."""
    f = torch.fft.rfftfreq(n_fft, d=1 / sr)

    def midi(hz):
        hz = torch.as_tensor(hz, dtype=torch.float32)
        note = 12.0 * (torch.log2(torch.clamp(hz, min=1e-7)) - math.log2(440.0)) + 69.0
        return torch.where(hz <= 0, torch.zeros_like(note), note)

    lo, hi = midi(fmin), midi(fmax)
    return ((midi(f) - lo) / (hi - lo)).to(torch.float32)
