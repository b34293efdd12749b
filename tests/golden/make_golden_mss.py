"""Generates tests/golden/mss_*.npz by running the UNMODIFIED reference `losses.MSSLoss` (losses.py:365-425) on
small seeded audio, with autograd gradients w.r.t. both signals.  Build-container only, like make_golden.py.

    python tests/golden/make_golden_mss.py
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
    gen = torch.Generator().manual_seed(77)
    tgt, f0 = S.harmonic_signals(3, gen)
    pred, _ = S.harmonic_signals(3, gen, n_partials=20, f0=f0 * 1.04, peak_normalise=False, amp_lo=0.05, amp_hi=0.5)
    cases = [
        ("mss_paper_l1", dict(fft_sizes=(2048, 1024, 512, 256, 128, 64), loss_type="L1", mag_weight=1, logmag_weight=0)),
        ("mss_l1_mag_log", dict(fft_sizes=(1024, 256, 64), loss_type="L1", mag_weight=1.0, logmag_weight=0.5)),
        ("mss_l2_mag_log", dict(fft_sizes=(512, 128), loss_type="L2", mag_weight=0.7, logmag_weight=1.0)),
    ]
    for name, ctor in cases:
        mod = ref.MSSLoss(**ctor)
        x = tgt.clone().requires_grad_(True)
        y = pred.clone().requires_grad_(True)
        value = mod(x, y)
        value.backward()
        save(name, dict(ctor=ctor), audio_x=tgt, audio_y=pred, value=value.detach(), grad_audio_x=x.grad,
             grad_audio_y=y.grad)
    # metrics.wasserstein_distance (metrics.py:144-149) restated with the reference's own pieces (metrics.py itself
    # needs mir_eval, which this image lacks): hann STFT magnitudes -> Wasserstein1D(p, fixed_x=n_bins)
    feats = sys.modules["features"]
    for p in (1, 2):
        with torch.inference_mode():
            mag_x = feats.compute_mag(tgt, size=512).permute(0, 2, 1)
            mag_y = feats.compute_mag(pred, size=512).permute(0, 2, 1)
            value = ref.Wasserstein1D(p=p, fixed_x=mag_x.shape[-1])(mag_x, mag_y)
        save(f"metric_wd_p{p}", dict(p=p, n_fft=512), audio_x=tgt, audio_y=pred, value=value)


if __name__ == "__main__":
    main()
