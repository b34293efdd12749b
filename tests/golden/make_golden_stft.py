"""Generates tests/golden/stft_*.npz: the reference's STFT front end + SOT loss on small seeded audio
(`losses.Wasserstein1DWithTransform`, losses.py:316-343; `features.compute_mag`, features.py:217-237),
with autograd gradients w.r.t. the audio AND w.r.t. the complex STFT frames.  They pin the fused
complex-input kernels (flag SOT_COMPLEX_INPUT).  Build-container only, like make_golden.py.

    python tests/golden/make_golden_stft.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import reference_loader  # noqa: E402
from sot_b200 import synthetic as S  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    ref = reference_loader.load()
    feats = sys.modules["features"]
    gen = torch.Generator().manual_seed(2024)
    cases = [
        ("stft512_cut", 512, dict(p=2, square_dist=True, dont_normalize=True, limit_quantile_range=True)),
        ("stft512_nocut", 512, dict(p=2, square_dist=True)),
        ("stft512_p1_abs", 512, dict(p=1, square_dist=False)),
        ("stft2048_cut", 2048, dict(p=2, square_dist=True, dont_normalize=True, limit_quantile_range=True)),
        ("stft2048_nocut", 2048, dict(p=2, square_dist=True)),
    ]
    for name, n_fft, ctor in cases:
        tgt, f0 = S.harmonic_signals(2, gen)
        pred, _ = S.harmonic_signals(2, gen, n_partials=20, f0=f0 * 1.07, peak_normalise=False, amp_lo=0.05, amp_hi=0.5)
        tkw = dict(type="stft", n_fft=n_fft, hop_length=256, window="flattop", log=False)
        mod = ref.Wasserstein1DWithTransform(transform_kwargs=dict(tkw), **ctor)
        # (1) audio in, audio gradients out: the wrapper exactly as the reference runs it
        xa = tgt.clone().requires_grad_(True)
        ya = pred.clone().requires_grad_(True)
        value = mod(xa, ya)
        value.backward()
        # (2) the same loss from the complex frames, gradients w.r.t. the frames
        window = feats.get_window("flattop", n_fft)
        zx = feats.stft(tgt, frame_size=n_fft, overlap=1 - 256 / n_fft, window=window).detach().requires_grad_(True)
        zy = feats.stft(pred, frame_size=n_fft, overlap=1 - 256 / n_fft, window=window).detach().requires_grad_(True)
        pos = torch.fft.rfftfreq(n_fft, d=1 / 16000)
        pos = pos / pos.max()
        value_z = mod.wasserstein(zx.abs().permute(0, 2, 1), zy.abs().permute(0, 2, 1), x_pos=pos, y_pos=pos)
        value_z.backward()
        assert torch.equal(value_z.detach(), value.detach())
        save(name, dict(ctor=ctor, transform=tkw), audio_x=tgt, audio_y=pred, value=value.detach(),
             grad_audio_x=xa.grad, grad_audio_y=ya.grad,
             zx=zx.detach().permute(0, 2, 1).contiguous(), zy=zy.detach().permute(0, 2, 1).contiguous(),
             grad_zx=zx.grad.permute(0, 2, 1).contiguous(), grad_zy=zy.grad.permute(0, 2, 1).contiguous())


if __name__ == "__main__":
    main()
