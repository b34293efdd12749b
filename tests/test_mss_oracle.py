"""The MSS oracle (oracle/mss_oracle.py) against the reference fixtures -- CPU only."""
import pytest
import torch

from oracle import mss_oracle as M
from tests import golden_io as G


@pytest.mark.parametrize("name", ["mss_paper_l1", "mss_l1_mag_log", "mss_l2_mag_log"])
def test_mss_oracle_matches_the_reference(name):
    g = G.load(name)
    ctor = g["meta"]["ctor"]
    x = g["audio_x"].clone().requires_grad_(True)
    y = g["audio_y"].clone().requires_grad_(True)
    value = M.mss_loss(x, y, **ctor)
    value.backward()
    assert torch.equal(value.detach(), g["value"])  # same ops in the same order: bit-exact
    assert torch.equal(x.grad, g["grad_audio_x"]) and torch.equal(y.grad, g["grad_audio_y"])


def test_unknown_loss_type_raises_like_the_reference():
    with pytest.raises(ValueError, match="Loss type"):
        M.difference_mean(torch.zeros(3), torch.ones(3), "cosine")
