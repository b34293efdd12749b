"""Host-side behaviour of the drop-in module that needs no GPU: constructor surface, attribute
and buffer names, pickling, error types, and the refusal to compute anywhere but on CUDA."""
import pickle

import pytest
import torch

from sot_b200 import _capi, losses as L


def test_constructor_surface_matches_reference():
    m = L.Wasserstein1D(p=2, fixed_x=None, require_sort=True, log_scaled_x=False, cumsum_only=False,
                        dont_normalize=True, hinge=False, limit_quantile_range=True, square_dist=True)
    assert (m.p, m.require_sort, m.log_scaled_x) == (2, True, False)
    assert (m.dont_normalize, m.limit_quantile_range, m.hinge, m.square_dist) == (True, True, False, True)
    assert m.fixed_x is None and list(m.state_dict()) == []
    assert type(m).__name__ == "Wasserstein1D"  # trainer.py:216 compares the class name
    d = L.Wasserstein1D()
    assert (d.p, d.dont_normalize, d.limit_quantile_range, d.hinge, d.square_dist) == (1, False, False, False, False)
    f = L.Wasserstein1D(p=1, fixed_x=257)
    assert torch.equal(f.fixed_x, torch.linspace(0, 1, 257)) and list(f.state_dict()) == ["fixed_x"]


def test_pickle_round_trip():
    m = L.Wasserstein1D(p=2, fixed_x=33, square_dist=True, dont_normalize=True)
    m2 = pickle.loads(pickle.dumps(m))
    assert m2.p == 2 and m2.square_dist and m2.dont_normalize and torch.equal(m2.fixed_x, m.fixed_x)


def test_errors_match_reference_types():
    m = L.Wasserstein1D(p=2)
    with pytest.raises(ValueError, match="If fixed_x is not provided, x_pos and y_pos must be provided"):
        m(torch.rand(2, 3, 9), torch.rand(2, 3, 9))
    with pytest.raises(ValueError, match="x_pos and y_pos must be provided"):
        m(torch.rand(2, 3, 9), torch.rand(2, 3, 9), x_pos=torch.linspace(0, 1, 9))


def test_no_cpu_fallback():
    m = L.Wasserstein1D(p=2, fixed_x=9)
    with pytest.raises(_capi.SotError, match="CUDA only"):
        m(torch.rand(2, 3, 9), torch.rand(2, 3, 9))
    with pytest.raises(_capi.SotError, match="CUDA only"):
        L.wasserstein_1d(torch.rand(2, 9), torch.rand(2, 9))
    with pytest.raises(_capi.SotError):
        L.quantile_function(torch.rand(2, 4), torch.rand(2, 9), torch.rand(2, 9))


def test_p_below_one_is_an_assertion():
    with pytest.raises(AssertionError, match="only valid for p>=1"):
        L.sot_frames(torch.rand(2, 9), torch.rand(2, 9), torch.linspace(0, 1, 9), torch.linspace(0, 1, 9), p=0.5)
    with pytest.raises(AssertionError):
        L.wasserstein_1d(torch.rand(2, 9), torch.rand(2, 9), p=0)


def test_product_never_imports_the_oracle():
    import os
    import re
    pkg = os.path.dirname(L.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_sharded_module_pickles_without_its_peer_memory_mailboxes():
    import io
    import pickle
    from sot_b200 import sharding
    mod = sharding.ShardedWasserstein1D(p=2, square_dist=True, fixed_x=65, collective="auto")
    mod._reducer = object()  # stands in for a live PeerReducer (CUDA tensors + symmetric-memory handle)
    clone = pickle.loads(pickle.dumps(mod))
    assert clone._reducer is None and clone.collective == "auto" and clone.fixed_x.shape == (65,)
    buf = io.BytesIO()
    torch.save(mod, buf)
    buf.seek(0)
    assert torch.load(buf, weights_only=False)._reducer is None
    with pytest.raises(ValueError):
        sharding.ShardedWasserstein1D(collective="mpi")
