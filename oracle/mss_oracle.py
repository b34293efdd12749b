"""CPU restatement (torch ops) of the reference's multi-scale spectral loss for the parity tests of
`csrc/sot_mss.cu`.  TEST INFRASTRUCTURE ONLY: nothing in the product path imports this module.

Follows /root/reference: losses.py:365-425 (`MSSLoss.forward`), losses.py:7-36 (`mean_difference`),
utils.py:145-151 (`safe_log`), features.py:191-237 (`stft`, `compute_mag`), utils.py:252-274 (`pad_for_stft`).
Pinned against the reference itself through tests/golden/mss_*.npz (tests/golden/make_golden_mss.py).
"""
import torch

LOG_EPS = 1e-5


def guarded_log(x):  # utils.py:145-151
    eps = torch.tensor(LOG_EPS, dtype=x.dtype, device=x.device)
    return torch.log(torch.where(x <= eps, eps, x))


def difference_mean(target, value, loss_type):  # losses.py:7-36, dims=None
    diff = target - value
    kind = loss_type.upper()
    if kind == "L1":
        return diff.abs().mean()
    if kind == "L2":
        return (diff ** 2).mean()
    raise ValueError('Loss type ({}), must be "L1", "L2" '.format(kind))


def magnitude_frames(audio, size, overlap=0.75):  # features.py:191-237 with compute_mag's defaults
    hop = int(size * (1.0 - overlap))
    length = audio.shape[1]
    frames = -(-length // hop)
    pad = max(0, size + hop * (frames - 1) - length)
    audio = torch.nn.functional.pad(audio.float(), (0, pad))
    spec = torch.stft(audio, n_fft=size, hop_length=hop, win_length=size, window=torch.hann_window(size, device=audio.device), center=False,
                      normalized=True, return_complex=True)
    return spec.abs()


def term_from_magnitudes(target_mag, value_mag, mag_weight, logmag_weight, loss_type, start=0.0):
    """`start` + the two weighted means, added one after the other like losses.py:409-423 does."""
    out = start
    if mag_weight > 0:
        out = out + mag_weight * difference_mean(target_mag, value_mag, loss_type)
    if logmag_weight > 0:
        out = out + logmag_weight * difference_mean(guarded_log(target_mag), guarded_log(value_mag), loss_type)
    return out


def mss_loss(target_audio, audio, fft_sizes=(2048, 1024, 512, 256, 128, 64), loss_type="L1", mag_weight=0.0,
             logmag_weight=0.0):
    total = 0.0
    for size in fft_sizes:
        total = term_from_magnitudes(magnitude_frames(target_audio, size), magnitude_frames(audio, size),
                                     mag_weight, logmag_weight, loss_type, start=total)
    return total
