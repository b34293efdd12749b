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


@pytest.mark.parametrize("p,limit,shared", [(1, False, True), (2, True, True), (2, False, False), (3, True, False)])
def test_position_gradients_from_a_fixed_plan_are_the_references(p, limit, shared):
    """Gradients w.r.t. the SUPPORT POSITIONS (the reference's graph reaches them through `xs[idx]` in
    `quantile_function`, losses.py:219-220): `position_term_from_plan` -- plain torch ops on the plan the kernel
    emits -- evaluated here on the oracle's plan must reproduce the per-frame values and the autograd gradients of
    the reference's formula w.r.t. x_pos and y_pos (values: same arithmetic, exact; gradients: the scatter-adds run
    in another order, rel 1e-5)."""
    from oracle import sot_oracle as O
    gen = torch.Generator().manual_seed(7 * p + limit)
    N, F = 6, 33
    x, y = torch.rand(N, F, generator=gen), 1.3 * torch.rand(N, F, generator=gen)
    base = torch.sort(torch.rand(F, generator=gen))[0] if shared else torch.sort(torch.rand(N, F, generator=gen), 1)[0]
    pu = base.clone().requires_grad_(True)
    pv = (base + 0.01).clone().requires_grad_(True)
    kw = dict(p=p, square=True, cut_scale=limit, limit=limit)
    up = torch.rand(N, generator=gen)
    rows_ref = O.sot_per_frame(x, y, pu, pv, stable=True, **kw)
    g_ref = torch.autograd.grad((rows_ref * up).sum(), (pu, pv))
    uq, vq, qs, cu, cv = O.sot_quantiles(x, y, pu.detach(), pv.detach(), square=True, cut_scale=limit, stable=True)[:5]
    iu, iv = torch.searchsorted(cu, qs.contiguous()), torch.searchsorted(cv, qs.contiguous())
    rows = L.position_term_from_plan(qs, iu.int(), iv.int(), pu, pv, p, limit)
    g = torch.autograd.grad((rows * up).sum(), (pu, pv))
    assert torch.equal(rows.detach(), rows_ref.detach())
    torch.testing.assert_close(g[0], g_ref[0], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(g[1], g_ref[1], rtol=1e-5, atol=1e-7)
    assert g[0].abs().sum() > 0 and g[1].abs().sum() > 0


@pytest.mark.parametrize("case", ["mean, shared linear grid", "hinge, per-frame unsorted supports", "wasserstein_1d values"])
def test_position_gradient_path_host_logic(monkeypatch, case):
    """The whole host path of a call whose POSITIONS require gradients -- hoisted sort with the sorted positions kept in
    the graph, the plan "launch", `value + (term - term.detach())`, hinge gating, the mean -- with the CUDA entry
    points replaced by the oracle (tests/fake_kernels.py).  Against the oracle's own autograd on the same float32
    inputs; the GPU twin of this test is tests/test_gpu_parity.py::test_gradients_wrt_support_positions_*."""
    from oracle import sot_oracle as O
    from tests import fake_kernels
    fake_kernels.install(monkeypatch, _capi, L)
    gen = torch.Generator().manual_seed(len(case))
    N, F = 8, 33
    x, y = torch.rand(N, F, generator=gen) ** 4, torch.rand(N, F, generator=gen) ** 4
    if case == "mean, shared linear grid":
        pu0 = torch.linspace(0, 1, F)
        pv0 = pu0 + 0.003
        kw = dict(p=2, square=True, cut_scale=True, limit=True)
        mod = L.Wasserstein1D(p=2, square_dist=True, dont_normalize=True, limit_quantile_range=True)
        call = lambda xs, ys, a, b: mod(xs, ys, x_pos=a, y_pos=b)  # noqa: E731
        ref = lambda xs, ys, a, b: O.sot_loss(xs, ys, a, b, stable=True, **kw)  # noqa: E731
    elif case == "hinge, per-frame unsorted supports":
        pu0 = torch.rand(N, F, generator=gen)
        pv0 = torch.rand(N, F, generator=gen)
        kw = dict(p=1, square=False, cut_scale=False, limit=False)
        mod = L.Wasserstein1D(p=1, hinge=True)
        call = lambda xs, ys, a, b: mod(xs, ys, x_pos=a, y_pos=b, hinge=0.01)  # noqa: E731
        ref = lambda xs, ys, a, b: O.sot_loss(xs, ys, a, b, hinge_gate=True, hinge_at=0.01, stable=True, **kw)  # noqa: E731
    else:
        pu0 = torch.rand(N, F, generator=gen)
        pv0 = torch.rand(N, F, generator=gen) + 0.2
        call = lambda xs, ys, a, b: L.wasserstein_1d(a, b, p=2).mean()  # noqa: E731
        w = torch.full((N, F), 1.0 / F)
        ref = lambda xs, ys, a, b: O.w1d_rows(a, b, w, w, p=2, require_sort=True, limit=False, stable=True).mean()  # noqa: E731
    leaves = [t.clone().requires_grad_(True) for t in (x, y, pu0, pv0)]
    value = call(*leaves)
    value.backward()
    leaves_ref = [t.clone().requires_grad_(True) for t in (x, y, pu0, pv0)]
    want = ref(*leaves_ref)
    want.backward()
    assert value.item() == pytest.approx(want.item(), rel=1e-6)
    for mine, truth in zip(leaves[2:], leaves_ref[2:]):
        assert mine.grad is not None and mine.grad.abs().sum() > 0
        torch.testing.assert_close(mine.grad, truth.grad, rtol=2e-4, atol=1e-7)
    if case != "wasserstein_1d values":
        for mine, truth in zip(leaves[:2], leaves_ref[:2]):  # (the spectra gradients still arrive; a near-tie order flip
            err = (mine.grad - truth.grad).norm() / truth.grad.norm()  # between the two evaluation orders moves single bins)
            assert err <= 5e-2, err
