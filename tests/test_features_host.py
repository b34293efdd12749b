"""Host side of the STFT front end (sot_b200/features.py) against the reference fixtures -- CPU only.
The SOT kernels are not involved: this pins the framing / padding / window / normalisation that feeds them."""
import pytest
import torch

from sot_b200 import _capi, features
from tests import golden_io as G


@pytest.mark.parametrize("name", ["stft512_nocut", "stft2048_cut"])
def test_complex_frames_match_the_reference_stft(name):
    g = G.load(name)
    tkw = dict(g["meta"]["transform"])
    transform = features.get_transform(dict(tkw), sample_rate=16000)
    assert isinstance(transform, features.TorchSTFT)
    frames = transform.complex_frames(g["audio_x"])
    assert frames.shape == g["zx"].shape and frames.dtype == torch.complex64 and frames.is_contiguous()
    assert torch.allclose(torch.view_as_real(frames), torch.view_as_real(g["zx"]), rtol=1e-5, atol=1e-7)
    mag = transform(g["audio_x"])  # (batch, time, freq) like features.py:104-110
    assert torch.allclose(mag, g["zx"].abs(), rtol=1e-5, atol=1e-7)
    freqs = transform.get_frequencies()
    assert freqs.shape[0] == tkw["n_fft"] // 2 + 1 and freqs[-1].item() == 8000.0


def test_pad_for_stft_slides_the_window_past_the_end():
    for length, frame, hop in ((4000, 512, 256), (4096, 2048, 256), (100, 64, 16), (64, 64, 64)):
        out = features.pad_for_stft(torch.zeros(2, length), frame, hop)
        frames = -(-length // hop)
        assert out.shape[1] == max(length, frame + hop * (frames - 1))
        assert (out.shape[1] - frame) // hop + 1 == frames


def test_get_transform_names():
    assert isinstance(features.get_transform("identity", 16000), torch.nn.Identity)
    default = features.get_transform("stft", 16000)
    assert default.n_fft == 1024 and default.sr == 16000
    with pytest.raises(NotImplementedError):
        features.get_transform({"type": "cqt"}, 16000)
    with pytest.raises(ValueError, match="Unknown transform"):
        features.get_transform("mel", 16000)


def test_wrapper_has_no_cpu_path():
    mod = features.Wasserstein1DWithTransform(p=2, square_dist=True,
                                              transform_kwargs=dict(type="stft", n_fft=512, hop_length=256))
    assert isinstance(mod.wasserstein, features.Wasserstein1D)
    with pytest.raises(_capi.SotError, match="CUDA only"):
        mod(torch.randn(2, 4096), torch.randn(2, 4096))
