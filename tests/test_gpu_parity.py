"""GPU parity: the CUDA path (through the C ABI) against the oracle and the committed reference
fixtures.  Protocol P1-P4 of SURVEY.md App. B.

Tolerances (float32 arithmetic; stated where used):
  * P1 merged grid / searchsorted indices / quantile positions from injected CDFs: BIT-EXACT.
  * P2 per-frame loss from injected CDFs: rel 2e-6 (summation order only); dL/dCDF: BIT-EXACT.
  * loss from magnitudes, no-cut mode: rel 1e-5 per frame and in aggregate vs the reference.
  * loss from magnitudes, cutoff mode: the reference is discontinuous in the last ulp of the
    target CDF (strict `qs > 1` mask, App. B) -> compared on the kernel's own CDFs (which must be
    within 4 ulp of the fp64 CDFs) at rel 2e-6, plus aggregate rel 2e-3 vs the reference value.
  * gradients, two gates: (a) per frame, rel-L2 vs the fp64 continuation of the kernel's own CDFs within
    4 eps32 x the conditioning number + 2e-6; (b) per fixture and at the paper's batch size, rel-L2 error vs the
    reference evaluated in float64 (`ref64`: the oracle on upcast inputs) no worse than
    max(5e-5, r x the error of the reference's own float32 gradients against the same `ref64`), r = 1.5 at
    1024 frames and 3 on the small fixtures, where the single worst bin of a frame (one near-tie order flip) may
    be set aside as long as the full error stays below 5e-4 (see `_assert_grad_gate`, `FLIP_CAP`).
"""
import numpy as np
import pytest
import torch

from oracle import sot_oracle as O
from tests import golden_io as G

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def capi():
    from sot_b200 import _capi
    _capi.load()
    return _capi


@pytest.fixture(scope="module")
def L():
    from sot_b200 import losses
    return losses


def _tune(capi, tuning):
    """Select a kernel configuration; the alternatives of the production choice are only compiled with
    SOT_BUILD_TUNING=1 (`python -m sot_b200.build`): skip them when absent."""
    try:
        capi.set_tuning(*tuning)
    except ValueError:
        pytest.skip("kernel configuration not compiled in (build with SOT_BUILD_TUNING=1)")


# Floor of the gradient gate: on well-conditioned samples the error is plain rounding of the CDFs -- the kernel's
# are within 3 ulp of the fp64 CDF (gated below), the CPU reference accumulates its cumsum in fp64 (<= 0.5 ulp;
# its CUDA cumsum is an fp32 scan like ours) -- measured 3.2e-5 against the reference's 1.4e-5 on 32 frames.
GRAD_FLOOR = 5e-5


def _rel_l2(a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double())).item()


def _ref64_grads(x, y, px, py, kw, scale=1.0):
    """Gradients of scale * mean_n loss_n from the reference's formula evaluated in float64 (stable tie order)."""
    _, gx, gy = O.sot_loss_and_grads(x.double(), y.double(), px.double(), py.double(), stable=True, **kw)
    return gx * scale, gy * scale


# One near-tie order flip (two CDF entries less than their rounding apart trade places) moves ONE gradient bin of ONE
# frame; pooled over the 8 .. 64 frames of a fixture that single bin can be several times the sum of all rounding errors
# (sot2048_logf_unsorted with 33 bins per thread: one bin in each of two frames -> pooled 2.0e-4, without them 7.6e-5;
# the reference's own float32 gradients have frames at 7e-4 and 6e-2 in sot2048_cut).  Which pairs flip is a coin toss
# of the summation order -- tests/kernel_model.py reproduces that number on the CPU and shows 64 x 17, 32 x 33 and the
# reference statistically alike on 64 frames -- so the small-fixture gate sets the worst bin of each frame aside, but
# never lets the full error past FLIP_CAP.
FLIP_CAP = 5e-4


def _rel_l2_without_worst_bin(a, b):
    """Pooled rel-L2 of a against b with, per row, the bin of largest |a - b| left out."""
    d2 = (a.double() - b.double()).reshape(-1, a.shape[-1]) ** 2
    return (torch.sqrt((d2.sum(1) - d2.max(1).values).clamp_min(0).sum()) / torch.linalg.vector_norm(b.double())).item()


def _assert_grad_gate(mine, ref32, ref64, what, ratio=1.5):
    """North star: 'loss and gradients within rel 1e-5 in fp32 (tighter against an fp64 reference run)'.  The
    gradient of this loss is discontinuous in the last ulp of the fp32 CDFs, so the reference's
    own fp32 gradients sit 1e-4 .. 1e-2 from its fp64 ones; the gate is relative to that.  `ratio`: 1.5 at the
    paper's batch size (measured 0.7 .. 1.0: the ill-conditioned frames dominate both errors alike); 3 on the
    16 .. 64-frame fixtures, where the error counts near-tie order flips and the kernel's 3-ulp CDFs flip up to
    2.3 x as many as the CPU reference's fp64-accumulated ones (measured ratios 0.9 .. 2.3) -- there the worst bin
    of each frame may be set aside (see FLIP_CAP); at ratio 1.5 nothing is."""
    e_mine, e_ref = _rel_l2(mine, ref64), _rel_l2(ref32, ref64)
    bound = max(GRAD_FLOOR, ratio * e_ref)
    if e_mine > bound and ratio > 1.5:
        e_trim = _rel_l2_without_worst_bin(mine, ref64)
        assert e_trim <= bound and e_mine <= FLIP_CAP, \
            f"{what}: CUDA {e_mine:.3e} ({e_trim:.3e} without the worst bin per frame) vs fp64, reference fp32 {e_ref:.3e} vs fp64"
        return e_mine, e_ref
    assert e_mine <= bound, f"{what}: CUDA {e_mine:.3e} vs fp64, reference fp32 {e_ref:.3e} vs fp64"
    return e_mine, e_ref


def _flags(capi, kw):
    return ((capi.SOT_SQUARE if kw["square"] else 0) | (capi.SOT_CUT_SCALE if kw["cut_scale"] else 0) |
            (capi.SOT_LIMIT if kw["limit"] else 0))


def _sorted_case(g):
    """Golden case as sorted-support 2-D rows (the C ABI wants ascending supports)."""
    F = g["x"].shape[-1]
    x, y = g["x"].reshape(-1, F), g["y"].reshape(-1, F)
    pos, perm = torch.sort(g["pos_x"], stable=True)
    return x[:, perm].contiguous(), y[:, perm].contiguous(), pos.contiguous(), perm


# ------------------------------------------------------------------------------------------
# P1: bit-exact plan from the reference's own CDFs
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", G.MODULE_CASES)
@pytest.mark.parametrize("tuning", [(0, 0), (32, 33), (64, 9), (64, 17, 1), (64, 17, 2), (128, 9), (128, 17), (256, 17)])
def test_p1_plan_from_reference_cdfs_bit_exact(capi, name, tuning):
    g = G.load(name)
    F = g["x"].shape[-1]
    if tuning[0] * tuning[1] and tuning[0] * tuning[1] < F:
        pytest.skip("row does not fit this configuration")
    cu, cv = g["cu"].reshape(-1, F).contiguous(), g["cv"].reshape(-1, F).contiguous()
    pos = torch.sort(g["pos_x"], stable=True)[0]
    _tune(capi, tuning)
    try:
        uq, vq, qs, _, _, iu, iv = capi.quantiles(cu.to(DEV), cv.to(DEV), pos.to(DEV), pos.to(DEV), 0,
                                                  from_cdf=True, want_indices=True)
    finally:
        capi.set_tuning(0, 0, 0)
    K = 2 * F
    assert torch.equal(qs.cpu(), g["qs"].reshape(-1, K)), "merged quantile grid"
    iu_ref = torch.searchsorted(cu, g["qs"].reshape(-1, K).contiguous())
    iv_ref = torch.searchsorted(cv, g["qs"].reshape(-1, K).contiguous())
    assert torch.equal(iu.cpu().long(), iu_ref), "searchsorted(cu, qs)"
    assert torch.equal(iv.cpu().long(), iv_ref), "searchsorted(cv, qs)"
    assert torch.equal(uq.cpu(), g["uq"].reshape(-1, K)) and torch.equal(vq.cpu(), g["vq"].reshape(-1, K))


def test_p1_unequal_supports_and_heavy_ties(capi):
    rng = np.random.default_rng(3)
    N, n, m = 10, 37, 90
    wu = rng.random((N, n)).astype(np.float32) * (rng.random((N, n)) < 0.3)
    wv = rng.random((N, m)).astype(np.float32) * (rng.random((N, m)) < 0.3)
    wu[:, 0] += 0.1
    wv[:, 0] += 0.1
    wv[3, :n] = wu[3]  # shared values between the two rows -> cross ties
    cu = torch.from_numpy(np.cumsum(wu / wu.sum(1, keepdims=True), axis=1, dtype=np.float64).astype(np.float32))
    cv = torch.from_numpy(np.cumsum(wv / wu.sum(1, keepdims=True), axis=1, dtype=np.float64).astype(np.float32))
    pu = torch.sort(torch.rand(N, n), dim=1)[0]
    pv = torch.sort(torch.rand(N, m), dim=1)[0]
    uq, vq, qs, _, _, iu, iv = capi.quantiles(cu.to(DEV), cv.to(DEV), pu.to(DEV), pv.to(DEV), 0, from_cdf=True,
                                              want_indices=True)
    qs_ref = torch.sort(torch.cat((cu, cv), 1), dim=1, stable=True)[0]
    assert torch.equal(qs.cpu(), qs_ref)
    assert torch.equal(iu.cpu().long(), torch.searchsorted(cu, qs_ref))
    assert torch.equal(iv.cpu().long(), torch.searchsorted(cv, qs_ref))
    uq_ref, _ = O.lower_bound_lookup(qs_ref, cu, pu)
    vq_ref, _ = O.lower_bound_lookup(qs_ref, cv, pv)
    assert torch.equal(uq.cpu(), uq_ref) and torch.equal(vq.cpu(), vq_ref)


# ------------------------------------------------------------------------------------------
# P2: loss and dL/dCDF from injected CDFs vs the closed form (stable-sort tie attribution)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sot512_cut", "sot512_nocut", "sot2048_cut", "sot2048_nocut", "sot512_logf_cut"])
@pytest.mark.parametrize("p", [1.0, 2.0, 3.0])
@pytest.mark.parametrize("tuning", [(0, 0), (32, 33), (64, 9), (64, 17, 1), (64, 17, 2), (128, 9), (128, 17)])
@pytest.mark.parametrize("uniform", [False, True])
def test_p2_loss_and_cdf_gradients_from_reference_cdfs(capi, name, p, tuning, uniform):
    g = G.load(name)
    F = g["x"].shape[-1]
    if tuning[0] * tuning[1] and tuning[0] * tuning[1] < F:
        pytest.skip("row does not fit this configuration")
    limit = bool(g["meta"]["ctor"].get("limit_quantile_range", False))
    cu, cv = g["cu"].reshape(-1, F).contiguous(), g["cv"].reshape(-1, F).contiguous()
    pos = g["pos_x"]
    if uniform and g["meta"]["grid"] != "linear":
        pytest.skip("the uniform-grid kernels need an exact linear grid")
    flags = (capi.SOT_LIMIT if limit else 0) | (capi.SOT_UNIFORM_GRID if uniform else 0)
    _tune(capi, tuning)
    try:
        loss, g_cu, g_cv = capi.loss_from_cdf(cu.to(DEV), cv.to(DEV), pos.to(DEV), pos.to(DEV), p, flags)
        loss_only, _, _ = capi.loss_from_cdf(cu.to(DEV), cv.to(DEV), pos.to(DEV), pos.to(DEV), p, flags,
                                             want_grads=False)
    finally:
        capi.set_tuning(0, 0, 0)
    assert torch.equal(loss, loss_only), "forward-only and fused kernels disagree on the loss"
    for r in range(cu.shape[0]):
        pn = pos.numpy()
        ref_loss, r_cu, r_cv, _, _ = O.closed_form_from_cdfs(cu[r].numpy(), cv[r].numpy(), pn, pn, int(p), limit)
        assert abs(loss[r].item() - float(ref_loss)) <= 2e-6 * abs(float(ref_loss)) + 1e-12, f"frame {r} loss"
        if p == 3.0:  # powf on the device vs numpy power (<= 2 ulp apart): dL/dCDF is a DIFFERENCE of two
            # such values in [0, 1], so compare with an absolute tolerance of a few ulp(1)
            assert np.allclose(g_cu[r].cpu().numpy(), r_cu, rtol=1e-5, atol=5e-7)
            assert np.allclose(g_cv[r].cpu().numpy(), r_cv, rtol=1e-5, atol=5e-7)
        else:
            assert np.array_equal(g_cu[r].cpu().numpy(), r_cu), f"frame {r}: dL/dcu"
            assert np.array_equal(g_cv[r].cpu().numpy(), r_cv), f"frame {r}: dL/dcv"


# ------------------------------------------------------------------------------------------
# end to end from magnitudes: golden fixtures of the reference
# ------------------------------------------------------------------------------------------
def _ulp_distance(a, b):
    ai = a.view(torch.int32).long()
    bi = b.view(torch.int32).long()
    return (ai - bi).abs()


@pytest.mark.parametrize("name", G.MODULE_CASES)
def test_kernel_cdfs_within_ulps_of_fp64(capi, L, name):
    g = G.load(name)
    kw = G.oracle_kwargs(g["meta"]["ctor"])
    x, y, pos, _ = _sorted_case(g)
    out = capi.quantiles(x.to(DEV), y.to(DEV), pos.to(DEV), pos.to(DEV), _flags(capi, dict(kw, limit=False)))
    cu, cv = out[3].cpu(), out[4].cpu()
    wx, wy = O.spectra_to_weights(x.double(), y.double(), kw["square"], kw["cut_scale"])
    cu64, cv64 = torch.cumsum(wx, 1), torch.cumsum(wy, 1)
    # Normalised spectra: the kernel sums each thread's bins in packed fp32 (runs of <= 17 adds: 33 bins per
    # thread are cut into four runs), carries the offsets between threads in fp64 and rounds once more in the
    # final add -- measured <= 3 ulp from the fp64 CDF, like the reference's own float32 CDFs on the CPU (the row
    # end is exact to fp64; tests/test_kernel_model.py holds the CPU model of this stage).  The reference's own
    # CUDA cumsum is an fp32 scan; only its CPU cumsum accumulates in fp64.
    assert _ulp_distance(cu, cu64.float()).max().item() <= 4
    assert _ulp_distance(cv, cv64.float()).max().item() <= 4
    assert (cu[:, 1:] >= cu[:, :-1]).all() and (cv[:, 1:] >= cv[:, :-1]).all(), "CDFs must be non-decreasing"
    # and the plan built on them is exactly what the reference builds on the same CDFs
    qs_ref = torch.sort(torch.cat((cu, cv), 1), dim=1, stable=True)[0]
    assert torch.equal(out[2].cpu(), qs_ref)
    uq_ref, _ = O.lower_bound_lookup(qs_ref, cu, pos.unsqueeze(0).expand_as(cu))
    vq_ref, _ = O.lower_bound_lookup(qs_ref, cv, pos.unsqueeze(0).expand_as(cv))
    assert torch.equal(out[0].cpu(), uq_ref) and torch.equal(out[1].cpu(), vq_ref)


@pytest.mark.parametrize("name", G.MODULE_CASES)
@pytest.mark.parametrize("mode", ["onepass", "recompute"])
@pytest.mark.parametrize("short_rows", ["two frames per warp", "one warp per frame"])
def test_module_forward_backward_vs_reference_fixture(capi, L, name, mode, short_rows):
    """`short_rows`: rows of <= 272 bins run two frames per warp by default (16 threads x 17 bins per frame); the
    quantile taps that give the kernel's own CDFs for gate (a) come from the one-warp-per-frame kernels (32 x 9:
    other local sums, CDFs a few ulp apart), so gate (a) is checked with that configuration selected and the
    default one goes through everything else."""
    g = G.load(name)
    ctor = dict(g["meta"]["ctor"])
    kw = G.oracle_kwargs(ctor)
    F = g["x"].shape[-1]
    default_is_sub = F <= 272
    if short_rows == "one warp per frame" and not default_is_sub:
        pytest.skip("only rows of <= 272 bins have the two configurations")
    same_cdfs = not (default_is_sub and short_rows == "two frames per warp")
    mod = L.Wasserstein1D(**ctor, backward_mode=mode)
    x = g["x"].to(DEV).requires_grad_(True)
    y = g["y"].to(DEV).requires_grad_(True)
    px, py = g["pos_x"].to(DEV), g["pos_y"].to(DEV)
    if short_rows == "one warp per frame":
        _tune(capi, (32, 9, 1))
    try:
        value = mod(x, y, x_pos=px, y_pos=py)
        value.backward()
        with torch.no_grad():
            rows = mod(g["x"].reshape(-1, 1, F).to(DEV), g["y"].reshape(-1, 1, F).to(DEV), x_pos=px, y_pos=py, dims=1)
    finally:
        capi.set_tuning(0, 0, 0)
    assert value.shape == g["value"].shape and rows.shape == g["rows"].shape
    assert x.grad.shape == g["grad_x"].shape and y.grad.shape == g["grad_y"].shape
    ref_rows = g["rows"]
    rel = ((rows.cpu() - ref_rows).abs() / ref_rows.abs().clamp_min(1e-30))
    if not kw["limit"]:
        assert rel.max().item() <= 1e-5, f"per-frame loss, no-cut: {rel}"
        assert abs(value.item() - g["value"].item()) <= 1e-5 * abs(g["value"].item())
    else:
        # cutoff mode: a frame whose target CDF ends within an ulp of 1 keeps or drops its last slots with the
        # rounding of the CDF (App. B) -- the reference's own fp32 value is that far from its fp64 value
        v64 = O.sot_loss(g["x"].double(), g["y"].double(), g["pos_x"].double(), g["pos_y"].double(), **kw).item()
        slack = max(1e-5, 1.5 * abs(g["value"].item() - v64) / abs(v64))
        assert abs(value.item() - v64) <= slack * abs(v64), (value.item(), g["value"].item(), v64)

    # loss and gradients against the fp64 continuation of the KERNEL's CDFs
    xs, ys, pos, perm = _sorted_case(g)
    out = capi.quantiles(xs.to(DEV), ys.to(DEV), pos.to(DEV), pos.to(DEV), _flags(capi, dict(kw, limit=False)))
    cu, cv = out[3].cpu().numpy(), out[4].cpu().numpy()
    a = xs.double() ** 2 if kw["square"] else xs.double()
    b = ys.double() ** 2 if kw["square"] else ys.double()
    mass_u, mass_v = a.sum(1).float().numpy(), b.sum(1).float().numpy()
    gx = x.grad.reshape(-1, F).cpu()[:, perm].numpy() * rows.numel()  # undo the mean's 1/N
    gy = y.grad.reshape(-1, F).cpu()[:, perm].numpy() * rows.numel()
    ref_gx = g["grad_x"].reshape(-1, F)[:, perm].numpy() * rows.numel()
    ref_gy = g["grad_y"].reshape(-1, F)[:, perm].numpy() * rows.numel()
    f8 = np.float64
    for r in range(xs.shape[0]):
        if not same_cdfs:
            break  # (gate (a) needs the CDFs of the configuration that ran: see the docstring)
        tl, tgx, tgy = O.fp64_chain_from_cdfs(xs[r].numpy(), ys[r].numpy(), cu[r], cv[r], pos.numpy(), pos.numpy(),
                                              mass_u[r], mass_v[r], kw["p"], kw["square"], kw["cut_scale"],
                                              kw["limit"])
        assert abs(rows[r].item() - tl) <= 2e-6 * abs(tl) + 1e-12, f"frame {r}: loss vs fp64 chain on kernel CDFs"
        # Conditioning of the gradient: dL/dx = 2x (gw - <gw, w>) / M subtracts two O(|gw|) numbers.
        # An fp32 evaluation (the reference's autograd included) carries eps32 * |uncancelled term|;
        # measured on the CPU model of the kernel the ratio error / (eps32 * cond) stays below 1.3.
        _, _, _, g_wu, g_wv = O.closed_form_from_cdfs(cu[r].astype(f8), cv[r].astype(f8), pos.numpy().astype(f8),
                                                      pos.numpy().astype(f8), kw["p"], kw["limit"])
        eps_m = np.float32(O.SAFE_EPS)
        mu = max(mass_u[r], eps_m)
        mv = mu if kw["cut_scale"] else max(mass_v[r], eps_m)
        lead_x = (2 * xs[r].numpy() if kw["square"] else 1.0) * g_wu / f8(mu)
        lead_y = (2 * ys[r].numpy() if kw["square"] else 1.0) * g_wv / f8(mv)
        for mine, truth, lead, nm in ((gx[r], tgx, lead_x, "x"), (gy[r], tgy, lead_y, "y")):
            nt = np.linalg.norm(truth)
            if nt == 0:
                assert np.linalg.norm(mine) == 0
                continue
            cond = np.linalg.norm(lead) / nt
            e_mine = np.linalg.norm(mine - truth) / nt
            bound = (4.0 if kw["p"] != 3 else 16.0) * 5.96e-8 * cond + 2e-6
            assert e_mine <= bound, f"frame {r} grad_{nm}: rel-L2 {e_mine:.3e} > {bound:.3e} (cond {cond:.1f})"
            # and never further from the fp64 continuation than 2x the reference's own fp32 autograd
            # would be allowed to be under the same conditioning argument
        if kw["limit"]:
            same = (np.array_equal(cu[r], g["cu"].reshape(-1, F)[r].numpy()) and
                    np.array_equal(cv[r], g["cv"].reshape(-1, F)[r].numpy()))
            if same:
                assert abs(rows[r].item() - ref_rows[r].item()) <= 2e-6 * abs(ref_rows[r].item())
    # (b) the reference's own gradients: CUDA must be as close to the float64 evaluation of the reference's formula
    # as the reference's float32 autograd (the fixture) is, per fixture
    t64x, t64y = _ref64_grads(g["x"].reshape(-1, F), g["y"].reshape(-1, F), g["pos_x"], g["pos_y"], kw)
    t64x, t64y = t64x[:, perm].numpy() * rows.numel(), t64y[:, perm].numpy() * rows.numel()
    _assert_grad_gate(torch.from_numpy(gx), torch.from_numpy(ref_gx), torch.from_numpy(t64x), f"{name} grad_x", 3.0)
    _assert_grad_gate(torch.from_numpy(gy), torch.from_numpy(ref_gy), torch.from_numpy(t64y), f"{name} grad_y", 3.0)


@pytest.mark.parametrize("name", ["sot512_cut", "sot2048_cut", "sot2048_nocut", "sot512_p1_nosquare", "sot512_p3"])
def test_uniform_grid_kernels_are_bit_identical_to_the_general_ones(capi, name):
    """Positions computed as pos[0] + i*h (exact for these grids) vs positions loaded from memory."""
    g = G.load(name)
    kw = G.oracle_kwargs(g["meta"]["ctor"])
    F = g["x"].shape[-1]
    x, y = g["x"].reshape(-1, F).to(DEV), g["y"].reshape(-1, F).to(DEV)
    pos = g["pos_x"].to(DEV)
    flags = _flags(capi, kw)
    up = torch.rand(x.shape[0], device=DEV)
    a = capi.forward_backward(x, y, pos, pos, float(kw["p"]), flags, upstream=up)
    b = capi.forward_backward(x, y, pos, pos, float(kw["p"]), flags | capi.SOT_UNIFORM_GRID, upstream=up)
    for s, t2 in zip(a, b):
        assert torch.equal(s, t2)
    assert torch.equal(capi.forward(x, y, pos, pos, float(kw["p"]), flags),
                       capi.forward(x, y, pos, pos, float(kw["p"]), flags | capi.SOT_UNIFORM_GRID))
    # different lengths and offsets on the two sides: pos_v = 0.25 + j*h on 200 bins vs pos_u on F bins
    h = (pos[1] - pos[0]).item()
    pv = 0.25 + torch.arange(200, device=DEV) * h
    yv = y[:, :200].contiguous()
    a = capi.forward_backward(x, yv, pos, pv, float(kw["p"]), flags)
    b = capi.forward_backward(x, yv, pos, pv, float(kw["p"]), flags | capi.SOT_UNIFORM_GRID)
    for s, t2 in zip(a, b):
        assert torch.equal(s, t2)


def test_module_detects_uniform_grids(L):
    from sot_b200 import synthetic as S
    lin = S.linear_positions(2048).to(DEV)
    assert L._uniform_grid(lin, lin.clone(), lin, lin)
    fx = torch.linspace(0, 1, 257, device=DEV)
    assert L._uniform_grid(fx, fx, fx, fx)
    logf = S.logf_positions(512).to(DEV)
    assert not L._uniform_grid(logf, logf, logf, logf)
    odd = torch.linspace(0, 1, 1000, device=DEV)  # step 1/999: not exact in binary
    assert not L._uniform_grid(odd, odd, odd, odd)


def test_fused_mean_path_matches_the_per_frame_path(capi, L):
    """Wasserstein1D(...)(x, y) takes the two-launch path with the mean folded into the kernels;
    it must agree with mean(per-frame rows) and its gradients with the rows path."""
    g = G.load("sot2048_cut")
    ctor = g["meta"]["ctor"]
    px, py = g["pos_x"].to(DEV), g["pos_y"].to(DEV)
    x1 = g["x"].to(DEV).requires_grad_(True)
    y1 = g["y"].to(DEV).requires_grad_(True)
    before = capi.launch_count()
    v1 = L.Wasserstein1D(**ctor)(x1, y1, x_pos=px, y_pos=py)
    (0.5 * v1).backward()
    assert capi.launch_count() - before == 2, "one forward and one backward launch"
    assert v1.ndim == 0 and v1.dtype == torch.float32
    x2 = g["x"].to(DEV).requires_grad_(True)
    y2 = g["y"].to(DEV).requires_grad_(True)
    rows = L.sot_frames(x2, y2, px, py, p=2, square=True, cut_scale=True, limit=True)
    v2 = rows.mean()
    (0.5 * v2).backward()
    assert abs(v1.item() - v2.item()) <= 1e-6 * abs(v2.item())
    assert torch.allclose(x1.grad, x2.grad, rtol=1e-6, atol=0) and torch.allclose(y1.grad, y2.grad, rtol=1e-6, atol=0)
    # inference_mode / no grad: one launch
    before = capi.launch_count()
    with torch.inference_mode():
        v3 = L.Wasserstein1D(**ctor)(g["x"].to(DEV), g["y"].to(DEV), x_pos=px, y_pos=py)
    assert capi.launch_count() - before == 1 and abs(v3.item() - v2.item()) <= 1e-6 * abs(v2.item())


def test_saved_merge_indices_give_bit_identical_gradients(capi):
    g = G.load("sot2048_cut")
    F = 1025
    x, y = g["x"].reshape(-1, F).to(DEV), g["y"].reshape(-1, F).to(DEV)
    pos = g["pos_x"].to(DEV)
    scale = torch.full((1,), 0.125, device=DEV)
    for flags in (capi.SOT_SQUARE, capi.SOT_SQUARE | capi.SOT_CUT_SCALE | capi.SOT_LIMIT | capi.SOT_UNIFORM_GRID):
        total, rows, coranks = capi.forward_sum(x, y, pos, pos, 2.0, flags, want_rows=True, save_coranks=True)
        assert coranks.shape == (x.shape[0], capi.load().sot_coranks_per_frame(F, F)) and coranks.dtype == torch.int16
        assert (coranks >= 0).all() and (coranks <= F).all() and (coranks[:, 0] == 0).all()
        assert (coranks[:, 1:] >= coranks[:, :-1]).all(), "co-ranks are non-decreasing along the merge path"
        assert total.item() == pytest.approx(rows.double().sum().item(), rel=1e-12)
        a = capi.forward_backward_scaled(x, y, pos, pos, 2.0, flags, scale, coranks=coranks)
        b = capi.forward_backward_scaled(x, y, pos, pos, 2.0, flags, scale, coranks=None)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        # co-ranks of another configuration are ignored, not misused
        capi.set_tuning(128, 9, 1)
        try:
            c = capi.forward_backward_scaled(x, y, pos, pos, 2.0, flags, scale, coranks=coranks)
        finally:
            capi.set_tuning(0, 0, 0)
        # (another configuration sums other groups of bins: CDFs a few ulp apart, see the conditioning note)
        assert (c[0] - b[0]).norm() <= 1e-3 * b[0].norm()


def test_fused_and_recompute_modes_agree_bitwise(L):
    g = G.load("sot2048_cut")
    outs = []
    for mode in ("recompute", "fused"):
        mod = L.Wasserstein1D(**g["meta"]["ctor"], backward_mode=mode)
        x = g["x"].to(DEV).requires_grad_(True)
        y = g["y"].to(DEV).requires_grad_(True)
        v = mod(x, y, x_pos=g["pos_x"].to(DEV), y_pos=g["pos_y"].to(DEV))
        (3.0 * v).backward()
        outs.append((v.detach().clone(), x.grad.clone(), y.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    # "fused" rounds the unit gradient and then multiplies by the upstream factor; "recompute" folds
    # the factor into the normalisation constant first: a few roundings apart
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=0)
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-6, atol=0)


@pytest.mark.parametrize("tuning", [(128, 9), (64, 17, 1), (64, 17, 2), (32, 33), (128, 17), (256, 17), (256, 33), (32, 9), (64, 9)])
def test_every_kernel_configuration_gives_the_same_answer(capi, L, tuning):
    g = G.load("sot2048_nocut")
    base = None
    for t in ((0, 0), tuning):
        _tune(capi, t)
        try:
            mod = L.Wasserstein1D(**g["meta"]["ctor"])
            x = g["x"].to(DEV).requires_grad_(True)
            y = g["y"].to(DEV).requires_grad_(True)
            v = mod(x, y, x_pos=g["pos_x"].to(DEV), y_pos=g["pos_y"].to(DEV))
            v.backward()
        finally:
            capi.set_tuning(0, 0, 0)
        cur = (v.item(), x.grad.clone(), y.grad.clone())
        if base is None:
            base = cur
    # configurations differ in how many bins a thread sums locally -> CDFs a few ulp apart -> the
    # (ill-conditioned, test_module_forward_backward_vs_reference_fixture) gradients ~1e-4 apart
    assert abs(cur[0] - base[0]) <= 2e-6 * abs(base[0])
    for a, b in ((cur[1], base[1]), (cur[2], base[2])):
        assert (a - b).norm().item() <= 1e-3 * b.norm().item()


# ------------------------------------------------------------------------------------------
# parity at the paper's batch size (1024 frames = 64 signals x 16), all four BASELINE configurations
# ------------------------------------------------------------------------------------------
PAPER_CONFIGS = {"SOT-2048": (2048, True, "linear"), "SOT-512": (512, True, "linear"),
                 "SOT-512-LogF": (512, True, "logf"), "SOT-NoCut": (2048, False, "linear")}


@pytest.mark.parametrize("config", sorted(PAPER_CONFIGS))
@pytest.mark.parametrize("mode", ["onepass", "recompute"])
@pytest.mark.parametrize("tuning", [(0, 0), (16, 17, 1), (32, 9, 1)])
def test_parity_at_the_papers_batch_size(capi, L, config, mode, tuning):
    """The module as the trainer calls it (trainer.py:209-221) on 1024 synthetic frames, against the reference's
    float32 evaluation (the pinned oracle) and its float64 evaluation.  Bounds = the north star's 1e-5 on the batch
    loss in EVERY mode (measured 1e-7 .. 3.3e-6, profiles/r01j_parity_at_scale.jsonl) and the fp64-relative
    gradient gate."""
    from sot_b200 import synthetic as S
    n_fft, cut, grid = PAPER_CONFIGS[config]
    if tuning[0] and n_fft != 512:
        pytest.skip("the short-row configurations (two frames per warp / one warp per frame) hold 257 bins")
    x, y = S.sot_batch(64, n_fft, seed=42)
    pos = S.linear_positions(n_fft) if grid == "linear" else S.logf_positions(n_fft)
    kw = dict(p=2, square=True, cut_scale=cut, limit=cut)
    rows32, g32x, g32y = O.sot_loss_and_grads(x, y, pos, pos, stable=True, **kw)
    rows64, g64x, g64y = O.sot_loss_and_grads(x.double(), y.double(), pos.double(), pos.double(), stable=True, **kw)
    mod = L.Wasserstein1D(p=2, square_dist=True, dont_normalize=cut, limit_quantile_range=cut, backward_mode=mode)
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    _tune(capi, tuning)
    try:
        value = mod(xd, yd, x_pos=pos.to(DEV), y_pos=pos.to(DEV))
        value.backward()
        with torch.no_grad():
            rows = mod(x.reshape(-1, 1, x.shape[-1]).to(DEV), y.reshape(-1, 1, y.shape[-1]).to(DEV),
                       x_pos=pos.to(DEV), y_pos=pos.to(DEV), dims=1).cpu()
    finally:
        capi.set_tuning(0, 0, 0)
    ref_mean = rows32.double().mean().item()
    assert abs(value.item() - ref_mean) <= 1e-5 * abs(ref_mean), (config, value.item(), ref_mean)
    # per frame: no further from the float64 value than the reference's float32 value is, on average
    e_mine = (rows.reshape(-1).double() - rows64).abs().mean().item()
    e_ref = (rows32.double() - rows64).abs().mean().item()
    assert e_mine <= 1.5 * e_ref + 1e-9, (config, e_mine, e_ref)
    _assert_grad_gate(xd.grad.cpu().reshape(g64x.shape), g32x, g64x, f"{config} d/dtarget")
    _assert_grad_gate(yd.grad.cpu().reshape(g64y.shape), g32y, g64y, f"{config} d/dprediction")


# ------------------------------------------------------------------------------------------
# the rest of the reference's surface
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [1, 2])
def test_fixed_grid_metric_path(L, p):
    g = G.load(f"fixedx65_p{p}")
    mod = L.Wasserstein1D(p=p, fixed_x=65).to(DEV)
    with torch.inference_mode():  # metrics.py:146
        value = mod(g["x"].to(DEV), g["y"].to(DEV))
        per_item = mod(g["x"].to(DEV), g["y"].to(DEV), dims=1)
    assert torch.allclose(value.cpu(), g["value"], rtol=1e-5, atol=0)
    assert per_item.shape == g["per_item"].shape
    assert torch.allclose(per_item.cpu(), g["per_item"], rtol=1e-5, atol=0)
    # buffer left on the CPU (MixOfLosses keeps a plain list, losses.py:354): still works
    cpu_buf = L.Wasserstein1D(p=p, fixed_x=65)
    assert torch.allclose(cpu_buf(g["x"].to(DEV), g["y"].to(DEV)).cpu(), g["value"], rtol=1e-5, atol=0)


def test_hinge_gate_and_threshold(L):
    g = G.load("fixedx65_hinge")
    mod = L.Wasserstein1D(**g["meta"]["ctor"]).to(DEV)
    x = g["x"].to(DEV).requires_grad_(True)
    y = g["y"].to(DEV).requires_grad_(True)
    v = mod(x, y, **g["meta"]["call"])
    v.backward()
    assert torch.allclose(v.cpu(), g["value"], rtol=1e-5, atol=0)
    ctor, call = g["meta"]["ctor"], g["meta"]["call"]
    xd = g["x"].double().requires_grad_(True)
    yd = g["y"].double().requires_grad_(True)
    grid = torch.linspace(0, 1, ctor["fixed_x"]).double()
    rows64 = O.sot_per_frame(xd, yd, grid, grid, p=ctor.get("p", 1), square=bool(ctor.get("square_dist", False)),
                             stable=True)
    torch.relu(rows64 - call.get("hinge", 0.0)).mean().backward()
    for mine, ref, t64, nm in ((x.grad.cpu(), g["grad_x"], xd.grad, "x"), (y.grad.cpu(), g["grad_y"], yd.grad, "y")):
        _assert_grad_gate(mine, ref, t64.reshape(ref.shape), f"hinge grad_{nm}", 3.0)
        assert torch.equal(mine.reshape(30, -1).abs().sum(1) == 0, ref.reshape(30, -1).abs().sum(1) == 0), "hinge gate"


@pytest.mark.parametrize("n_bins", [271, 272, 273, 288, 289, 576, 1055, 1056, 1057, 1088, 1089, 1152])
def test_rows_that_fill_a_kernel_configuration_exactly(L, n_bins):
    """Row lengths at and around threads x bins-per-thread of the configurations (16x17, 32x9, 64x9, 32x33, 64x17, 128x9): every
    thread's bins lie inside the row (the +inf sentinel is then written by thread 0), one bin short, one bin over
    (next configuration).  No-cut mode: the loss is continuous there, so the reference's float32 value is the gate."""
    gen = torch.Generator().manual_seed(n_bins)
    x = torch.rand(6, n_bins, generator=gen) ** 4 + 1e-4
    y = torch.rand(6, n_bins, generator=gen) ** 4 + 1e-4
    pos = torch.linspace(0, 1, n_bins)
    rows32, g32x, g32y = O.sot_loss_and_grads(x, y, pos, pos, stable=True, p=2, square=True)
    _, g64x, g64y = O.sot_loss_and_grads(x.double(), y.double(), pos.double(), pos.double(), stable=True, p=2, square=True)
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    mod = L.Wasserstein1D(p=2, square_dist=True)
    v = mod(xd, yd, x_pos=pos.to(DEV), y_pos=pos.to(DEV))
    v.backward()
    assert abs(v.item() - rows32.mean().item()) <= 1e-5 * rows32.mean().item()
    with torch.no_grad():
        rows = mod(xd.detach()[:, None], yd.detach()[:, None], x_pos=pos.to(DEV), y_pos=pos.to(DEV), dims=1).cpu()
    assert torch.allclose(rows, rows32, rtol=1e-5, atol=0)
    _assert_grad_gate(xd.grad.cpu(), g32x, g64x, f"{n_bins} bins d/dtarget", 3.0)
    _assert_grad_gate(yd.grad.cpu(), g32y, g64y, f"{n_bins} bins d/dprediction", 3.0)


@pytest.mark.parametrize("n_frames", [1, 2, 3, 7, 33])
@pytest.mark.parametrize("p,cut,grid", [(2, True, "linear"), (2, False, "logf"), (1, False, "linear"), (3, True, "linear")])
def test_two_frames_per_warp_configuration(capi, L, n_frames, p, cut, grid):
    """The sub-warp configuration for short rows (16 threads per frame, two frames per warp) against the
    one-warp-per-frame kernels and the oracle: odd frame counts (the spare half-warp redoes the last frame), a
    poisoned frame next to a healthy one in the same warp, both grids, three p."""
    from sot_b200 import synthetic as S
    x, y = S.sot_batch(3, 512, seed=7 + n_frames)
    x, y = x.reshape(-1, 257)[:n_frames].clone(), y.reshape(-1, 257)[:n_frames].clone()
    if n_frames >= 3:
        x[1, 5] = float("nan")  # poisons frame 1 only
    pos = S.linear_positions(512) if grid == "linear" else S.logf_positions(512)
    ctor = dict(p=p, square_dist=True, dont_normalize=cut, limit_quantile_range=cut)
    out = {}
    for name, tuning in (("warp", (32, 9, 1)), ("half", (16, 17, 1))):
        _tune(capi, tuning)
        try:
            xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
            rows = L.Wasserstein1D(**ctor)(xd[:, None], yd[:, None], x_pos=pos.to(DEV), y_pos=pos.to(DEV), dims=1)
            rows.nansum().backward()
            mean = L.Wasserstein1D(**ctor)(x.to(DEV), y.to(DEV), x_pos=pos.to(DEV), y_pos=pos.to(DEV))
            with torch.no_grad():
                fwd = L.Wasserstein1D(**ctor)(x.to(DEV)[:, None], y.to(DEV)[:, None], x_pos=pos.to(DEV),
                                              y_pos=pos.to(DEV), dims=1)
        finally:
            capi.set_tuning(0, 0, 0)
        out[name] = (rows.detach().cpu(), xd.grad.cpu(), yd.grad.cpu(), mean.item(), fwd.cpu())
    a, b = out["warp"], out["half"]
    ok = torch.isfinite(a[0])
    assert torch.equal(ok, torch.isfinite(b[0])) and (n_frames < 3 or not ok[1]) and ok.sum() >= n_frames - 1
    assert torch.equal(b[0][ok], b[4][ok]), "forward-only and fused kernels of the sub-warp configuration disagree"
    # the two configurations sum different groups of bins locally: CDFs a few ulp apart
    assert torch.allclose(a[0][ok], b[0][ok], rtol=(2e-4 if cut else 2e-6), atol=1e-9)
    assert (a[3] != a[3]) == (b[3] != b[3]) and (a[3] != a[3] or abs(a[3] - b[3]) <= (2e-4 if cut else 2e-6) * abs(a[3]))
    for ga, gb in ((a[1], b[1]), (a[2], b[2])):
        assert torch.equal(torch.isnan(ga).any(1), ~ok) and torch.equal(torch.isnan(gb).any(1), ~ok)
        assert _rel_l2(gb[ok], ga[ok]) <= 5e-3
    # and against the reference's formula in float64 (healthy frames)
    kw = dict(p=p, square=True, cut_scale=cut, limit=cut)
    r64 = O.sot_per_frame(x[ok].double(), y[ok].double(), pos.double(), pos.double(), **kw)
    r32 = O.sot_per_frame(x[ok], y[ok], pos, pos, **kw)
    tol = torch.maximum(torch.full_like(r64, 1e-5), 1.5 * (r32.double() - r64).abs() / r64.abs())
    assert ((b[0][ok].double() - r64).abs() <= tol * r64.abs() + 1e-12).all()


def test_float64_inputs_are_accepted_like_the_reference(L):
    """The reference (and its `ref64` evaluation) takes float64 spectra: the drop-in converts, computes in float32
    and hands float64 values and gradients back."""
    g = G.load("sot512_nocut")
    ctor = dict(g["meta"]["ctor"])
    x = g["x"].double().to(DEV).requires_grad_(True)
    y = g["y"].double().to(DEV).requires_grad_(True)
    v = L.Wasserstein1D(**ctor)(x, y, x_pos=g["pos_x"].double().to(DEV), y_pos=g["pos_y"].double().to(DEV))
    v.backward()
    assert v.dtype == torch.float64 and x.grad.dtype == torch.float64 and y.grad.dtype == torch.float64
    assert abs(v.item() - g["value"].item()) <= 1e-5 * abs(g["value"].item())
    rows = L.Wasserstein1D(**ctor)(x.detach(), y.detach(), x_pos=g["pos_x"].to(DEV), y_pos=g["pos_y"].to(DEV), dims=1)
    assert rows.dtype == torch.float64 and rows.shape == g["x"].shape[:1]


@pytest.mark.parametrize("p,limit", [(1, False), (2, True)])
def test_module_level_w1d_unequal_unsorted_supports(L, p, limit):
    g = G.load(f"w1d_n37_m90_p{p}")
    uw = g["u_weights"].to(DEV).requires_grad_(True)
    vw = g["v_weights"].to(DEV).requires_grad_(True)
    rows = L.wasserstein_1d(g["u_values"].to(DEV), g["v_values"].to(DEV), uw, vw, p=p, limit_quantile_range=limit)
    rows.sum().backward()
    assert torch.allclose(rows.detach().cpu(), g["rows"], rtol=1e-5, atol=1e-8)
    # random weights: no exact ties, so the gradient is well defined -> compare directly
    assert (uw.grad.cpu() - g["grad_uw"]).norm() <= 1e-4 * g["grad_uw"].norm()
    assert (vw.grad.cpu() - g["grad_vw"]).norm() <= 1e-4 * g["grad_vw"].norm()
    q = L.wasserstein_1d(g["u_values"].to(DEV), g["v_values"].to(DEV), g["u_weights"].to(DEV), g["v_weights"].to(DEV),
                         p=p, return_quantiles=True)
    assert len(q) == 5
    for got, key in zip(q, ("uq", "vq", "qs", "cu", "cv")):
        assert got.shape == g[key].shape
        assert torch.allclose(got.cpu(), g[key], rtol=0, atol=3e-7), key
    u = G.load("w1d_uniform_p2")
    rows = L.wasserstein_1d(u["u_values"].to(DEV), u["v_values"].to(DEV), p=2)
    assert torch.allclose(rows.cpu(), u["rows"], rtol=1e-5, atol=1e-8)


def test_quantile_function_and_return_quantiles(L):
    g = G.load("quantile_function")
    out = L.quantile_function(g["qs"].to(DEV), g["cws"].to(DEV), g["xs"].to(DEV))
    assert torch.equal(out.cpu(), g["out"])
    c = G.load("sot512_nocut")
    mod = L.Wasserstein1D(**c["meta"]["ctor"])
    q = mod(c["x"].to(DEV), c["y"].to(DEV), x_pos=c["pos_x"].to(DEV), y_pos=c["pos_y"].to(DEV), return_quantiles=True)
    assert isinstance(q, list) and len(q) == 5
    for got, key in zip(q, ("uq", "vq", "qs", "cu", "cv")):
        assert got.shape == c[key].shape, key
    assert torch.allclose(q[2].cpu(), c["qs"], rtol=0, atol=3e-7)
    assert torch.allclose(q[3].cpu(), c["cu"], rtol=0, atol=2e-7)


def test_analytic_known_answers(L):
    k = G.load("kat_diracs")
    x, y = k["x"].to(DEV), k["y"].to(DEV)
    assert L.Wasserstein1D(p=1, fixed_x=9)(x, y).item() == pytest.approx(0.5, abs=1e-7)
    assert L.Wasserstein1D(p=2, fixed_x=9)(x, y).item() == pytest.approx(0.25, abs=1e-7)
    assert L.Wasserstein1D(p=2, fixed_x=9)(x, x).item() == 0.0
    z = torch.zeros(1, 9, device=DEV)
    cut = L.Wasserstein1D(p=2, fixed_x=9, dont_normalize=True, limit_quantile_range=True)
    assert cut(z, y).item() == pytest.approx(k["dead_cut"].item(), abs=1e-7)
    assert L.Wasserstein1D(p=2, fixed_x=9)(z, y).item() == pytest.approx(k["dead_nocut"].item(), rel=1e-6)


# ------------------------------------------------------------------------------------------
# edge cases: ragged batches, unaligned rows, per-frame supports, empty input, NaN
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_frames", [1, 2, 3, 5, 7, 13])
@pytest.mark.parametrize("offset", [0, 1])
def test_ragged_and_unaligned_batches(capi, n_frames, offset):
    g = G.load("sot512_nocut")
    F = 257
    x, y = g["x"].reshape(-1, F), g["y"].reshape(-1, F)
    pos = g["pos_x"].to(DEV)
    full_x, full_y = x.to(DEV), y.to(DEV)
    ref_loss, ref_gu, ref_gv = capi.forward_backward(full_x, full_y, pos, pos, 2.0, capi.SOT_SQUARE)
    # rows [offset, offset + n_frames): with offset 1 the base pointer is only 4-byte aligned,
    # and n_frames not a multiple of 4 leaves a ragged tail CTA -> the non-TMA load/store path
    sx = full_x[offset:offset + n_frames]
    sy = full_y[offset:offset + n_frames]
    assert sx.is_contiguous()
    loss, gu, gv = capi.forward_backward(sx, sy, pos, pos, 2.0, capi.SOT_SQUARE)
    assert torch.equal(loss, ref_loss[offset:offset + n_frames])
    assert torch.equal(gu, ref_gu[offset:offset + n_frames]) and torch.equal(gv, ref_gv[offset:offset + n_frames])
    assert torch.equal(capi.forward(sx, sy, pos, pos, 2.0, capi.SOT_SQUARE), loss)


def test_per_frame_supports_match_shared_support(capi, L):
    g = G.load("sot512_logf_cut")
    F = 257
    x, y = g["x"].reshape(-1, F).to(DEV), g["y"].reshape(-1, F).to(DEV)
    pos = g["pos_x"].to(DEV)
    flags = capi.SOT_SQUARE | capi.SOT_CUT_SCALE | capi.SOT_LIMIT
    a = capi.forward_backward(x, y, pos, pos, 2.0, flags)
    rep = pos.unsqueeze(0).repeat(x.shape[0], 1).contiguous()
    b = capi.forward_backward(x, y, rep, rep, 2.0, flags)
    for s, t in zip(a, b):
        assert torch.equal(s, t)
    # module level: 3-D per-frame positions, unsorted along the row (reversed grid + reversed spectra)
    mod = L.Wasserstein1D(p=2, square_dist=True)
    x3, y3 = g["x"].to(DEV), g["y"].to(DEV)
    p3 = pos.expand_as(x3).contiguous()
    v_sorted = mod(x3, y3, x_pos=p3, y_pos=p3)
    v_flipped = mod(x3.flip(-1), y3.flip(-1), x_pos=p3.flip(-1).contiguous(), y_pos=p3.flip(-1).contiguous())
    assert torch.equal(v_sorted, v_flipped)


def test_empty_batch_and_zero_spectra(capi, L):
    pos = torch.linspace(0, 1, 33, device=DEV)
    e = torch.empty(0, 33, device=DEV)
    assert capi.forward(e, e, pos, pos, 2.0, 0).shape == (0,)
    z = torch.zeros(4, 33, device=DEV, requires_grad=True)
    y = torch.rand(4, 33, device=DEV, requires_grad=True)
    for cut in (False, True):
        v = L.Wasserstein1D(p=2, square_dist=True, dont_normalize=cut, limit_quantile_range=cut)(z, y, pos, pos)
        v.backward()
        ref = O.sot_loss(torch.zeros(4, 33), y.detach().cpu(), pos.cpu(), pos.cpu(), p=2, square=True, cut_scale=cut,
                         limit=cut)
        assert torch.isfinite(v) and v.item() == pytest.approx(ref.item(), rel=1e-5, abs=1e-9)
        assert torch.isfinite(z.grad).all() and torch.isfinite(y.grad).all()


def test_nan_input_poisons_only_its_frame(capi):
    pos = torch.linspace(0, 1, 65, device=DEV)
    x = torch.rand(8, 65, device=DEV)
    y = torch.rand(8, 65, device=DEV)
    clean = capi.forward(x, y, pos, pos, 2.0, capi.SOT_SQUARE)
    x[2, 7] = float("nan")
    loss, gu, gv = capi.forward_backward(x, y, pos, pos, 2.0, capi.SOT_SQUARE)
    assert torch.isnan(loss[2]) and torch.isnan(gu[2]).all()
    keep = [0, 1, 3, 4, 5, 6, 7]
    assert torch.equal(loss[keep], clean[keep]) and torch.isfinite(gu[keep]).all() and torch.isfinite(gv[keep]).all()


def test_inputs_are_not_modified_and_wrong_dtype_raises(capi, L):
    pos = torch.linspace(0, 1, 65, device=DEV)
    x = torch.rand(8, 65, device=DEV)
    y = torch.rand(8, 65, device=DEV)
    x0, y0 = x.clone(), y.clone()
    capi.forward_backward(x, y, pos, pos, 2.0, capi.SOT_SQUARE | capi.SOT_CUT_SCALE | capi.SOT_LIMIT)
    assert torch.equal(x, x0) and torch.equal(y, y0)
    with pytest.raises(TypeError):
        L.Wasserstein1D(p=2, fixed_x=65)(x.long(), y.long())  # (float64 is converted like the reference accepts it)
    half = L.Wasserstein1D(p=2, fixed_x=65).to(DEV)(x.half(), y.half())
    assert torch.isfinite(half)


# ------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json config sizes; the oracle would take minutes here)
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big_batch():
    from sot_b200 import synthetic as S
    x, y = S.sot_batch(1024, 2048, seed=42, device=DEV)  # 16384 frames x 1025 bins
    return x, y, S.linear_positions(2048).to(DEV)


def test_full_size_properties(capi, L, big_batch):
    x, y, pos = big_batch
    F = x.shape[-1]
    xr, yr = x.reshape(-1, F), y.reshape(-1, F)
    sq = capi.SOT_SQUARE
    # identical spectra -> exactly zero, in both modes
    assert capi.forward(xr, xr, pos, pos, 2.0, sq).abs().max().item() == 0.0
    assert capi.forward(xr, xr, pos, pos, 2.0, sq | capi.SOT_CUT_SCALE | capi.SOT_LIMIT).abs().max().item() == 0.0
    loss, gu, gv = capi.forward_backward(xr, yr, pos, pos, 2.0, sq)
    assert torch.isfinite(loss).all() and (loss >= 0).all() and (loss <= 1.0).all()  # supports lie in [0, 1]
    # symmetry of W_2^2 between two unit-mass measures
    swapped = capi.forward(yr, xr, pos, pos, 2.0, sq)
    assert ((loss - swapped).abs() <= 1e-5 * loss.abs() + 1e-9).all()
    # no-cut mode is invariant to the scale of either spectrum (both are normalised): power-of-two
    # scalings are exact in floating point, so the result must not change in a single bit
    assert torch.equal(capi.forward(xr * 4.0, yr * 0.5, pos, pos, 2.0, sq), loss)
    # ... hence (Euler) <x, dL/dx> = 0 and <y, dL/dy> = 0 per frame, up to the fp32 cancellation
    # noise of ill-conditioned frames (see the conditioning bound in the fixture test)
    for gr, sp in ((gu, xr), (gv, yr)):
        ratio = (gr * sp).sum(1).abs() / ((gr.abs() * sp).sum(1) + 1e-30)
        assert ratio.median().item() <= 1e-4 and (ratio <= 1e-2).float().mean().item() >= 0.99
    # cutoff mode: scaling BOTH spectra by the same power of two changes nothing
    cut = sq | capi.SOT_CUT_SCALE | capi.SOT_LIMIT
    lc = capi.forward(xr, yr, pos, pos, 2.0, cut)
    assert torch.equal(capi.forward(xr * 2.0, yr * 2.0, pos, pos, 2.0, cut), lc)
    # forward-only and fused kernels agree on the loss; the module mean equals the row mean
    assert torch.equal(capi.forward(xr, yr, pos, pos, 2.0, sq), loss)
    mod = L.Wasserstein1D(p=2, square_dist=True)
    assert mod(x, y, x_pos=pos, y_pos=pos).item() == pytest.approx(loss.double().mean().item(), rel=1e-5)


def test_full_size_directional_derivative(L, big_batch):
    """Sanity check of the gradient's direction and scale by a central finite difference on a large
    batch.  Only a coarse check is possible: W_p^p between discrete measures on fixed supports is
    PIECEWISE LINEAR in the weights (an LP value), its breakpoints (a target CDF value crossing a
    prediction CDF value) are ~1e-9..1e-3 apart here, and any step large enough for float32 crosses
    many of them -- so the difference quotient averages slopes the analytic gradient does not see."""
    x, y, pos = big_batch
    x, y = x[:256], y[:256]
    mod = L.Wasserstein1D(p=2, square_dist=True)
    yv = y.clone().requires_grad_(True)
    v = mod(x, yv, x_pos=pos, y_pos=pos)
    v.backward()
    gen = torch.Generator(device=y.device).manual_seed(0)
    d = torch.randn(y.shape, generator=gen, device=y.device) * y.abs().mean()
    eps = 1e-2
    with torch.no_grad():
        up = mod(x, y + eps * d, x_pos=pos, y_pos=pos)  # weights are magnitudes squared: signs are harmless
        dn = mod(x, y - eps * d, x_pos=pos, y_pos=pos)
    fd = (up.double() - dn.double()).item() / (2 * eps)
    an = (yv.grad.double() * d.double()).sum().item()
    assert fd * an > 0 and 0.5 <= fd / an <= 2.0, (fd, an)


@pytest.mark.parametrize("case", ["mean, shared linear grid", "hinge, per-frame unsorted supports", "wasserstein_1d values"])
def test_gradients_wrt_support_positions_match_reference_autograd(L, case):
    """The reference's graph reaches x_pos / y_pos through `xs[idx]` (losses.py:219-220).  Here the plan kernel
    supplies merged grid and indices, torch differentiates `position_term_from_plan` (losses._position_term); the
    hinge (threshold 0.03: 4 of the 24 frames fall below it) gates both kinds of gradient; the spectra gradients still come from the fused kernel in the same backward.  Gate: rel-L2 1e-4 against the oracle's
    float64 autograd (position gradients are sums of dq * p|dx|^(p-1): continuous in the CDFs, no tie-order issue)."""
    gen = torch.Generator().manual_seed(len(case))
    N, F = 24, 257
    x, y = torch.rand(N, F, generator=gen) ** 4, torch.rand(N, F, generator=gen) ** 4
    if case == "mean, shared linear grid":
        pu0 = torch.linspace(0, 1, F)
        pv0 = pu0 + 0.003
        # (no cutoff here: the strict `qs > 1` mask makes value and gradients discontinuous in the last ulp of the CDF,
        # which is the subject of other tests; the cutoff form of this path runs in tests/test_host_logic.py)
        kw = dict(p=2, square=True, cut_scale=False, limit=False)
        mod = L.Wasserstein1D(p=2, square_dist=True)
        call = lambda xs, ys, a, b: mod(xs, ys, x_pos=a, y_pos=b)  # noqa: E731
        ref = lambda xs, ys, a, b: O.sot_loss(xs, ys, a, b, stable=True, **kw)  # noqa: E731
    elif case == "hinge, per-frame unsorted supports":
        pu0 = torch.rand(N, F, generator=gen)
        pv0 = torch.rand(N, F, generator=gen)
        kw = dict(p=1, square=False, cut_scale=False, limit=False)
        mod = L.Wasserstein1D(p=1, hinge=True)
        call = lambda xs, ys, a, b: mod(xs, ys, x_pos=a, y_pos=b, hinge=0.03)  # noqa: E731
        ref = lambda xs, ys, a, b: O.sot_loss(xs, ys, a, b, hinge_gate=True, hinge_at=0.03, stable=True, **kw)  # noqa: E731
    else:  # module-level entry: the VALUES are the positions, weights uniform
        pu0 = torch.rand(N, F, generator=gen)
        pv0 = torch.rand(N, F, generator=gen) + 0.2
        call = lambda xs, ys, a, b: L.wasserstein_1d(a, b, p=2).mean()  # noqa: E731
        w = torch.full((N, F), 1.0 / F, dtype=torch.float64)
        ref = lambda xs, ys, a, b: O.w1d_rows(a, b, w, w, p=2, require_sort=True, limit=False, stable=True).mean()  # noqa: E731
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    pu, pv = pu0.to(DEV).requires_grad_(True), pv0.to(DEV).requires_grad_(True)
    value = call(xd, yd, pu, pv)
    value.backward()
    x64, y64 = x.double().requires_grad_(True), y.double().requires_grad_(True)
    pu64, pv64 = pu0.double().requires_grad_(True), pv0.double().requires_grad_(True)
    want = ref(x64, y64, pu64, pv64)
    want.backward()
    assert abs(value.item() - want.item()) <= 5e-5 * abs(want.item())
    for mine, truth, nm in ((pu.grad, pu64.grad, "x_pos"), (pv.grad, pv64.grad, "y_pos")):
        assert mine is not None and mine.shape == truth.shape
        assert _rel_l2(mine.cpu(), truth) <= 1e-4, (case, nm, _rel_l2(mine.cpu(), truth))
    if case != "wasserstein_1d values":  # the spectra gradients still arrive, from the fused kernel
        assert xd.grad is not None and yd.grad is not None and xd.grad.abs().sum() > 0
        assert _rel_l2(yd.grad.cpu(), y64.grad) <= 5e-2  # (sanity only: the spectra gradients have their own gates)


def test_host_buffer_entry_point_matches_device_path(capi):
    g = G.load("sot2048_cut")
    F = 1025
    x = g["x"].reshape(-1, F).repeat(5, 1)[:37].contiguous().pin_memory()
    y = g["y"].reshape(-1, F).repeat(5, 1)[:37].contiguous().pin_memory()
    pos = g["pos_x"].contiguous()
    flags = capi.SOT_SQUARE | capi.SOT_CUT_SCALE | capi.SOT_LIMIT
    up = torch.rand(37)
    loss, gu, gv = capi.loss_grad_host(x, y, pos, pos, 2.0, flags, upstream=up)
    d = capi.forward_backward(x.to(DEV), y.to(DEV), pos.to(DEV), pos.to(DEV), 2.0, flags, upstream=up.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(loss, d[0].cpu()) and torch.equal(gu, d[1].cpu()) and torch.equal(gv, d[2].cpu())


@pytest.mark.parametrize("n_frames,n_bins", [(37, 1025), (300, 257), (5, 513)])
def test_host_buffer_entry_point_zero_copy_matches_device_path(capi, monkeypatch, n_frames, n_bins):
    """SOT_HOST_ZEROCOPY=1 with every buffer pinned: the kernels read the spectra from and write loss and gradients to
    host memory themselves (one launch, no staging copies).  Same numbers as the device path, bit for bit; with a
    pageable buffer among them the call falls back to the copy pipeline (same numbers again)."""
    gen = torch.Generator().manual_seed(n_frames)
    x = torch.rand(n_frames, n_bins, generator=gen).pin_memory()
    y = torch.rand(n_frames, n_bins, generator=gen).pin_memory()
    pos = torch.linspace(0, 1, n_bins)
    flags = capi.SOT_SQUARE | capi.SOT_CUT_SCALE | capi.SOT_LIMIT
    up = torch.rand(n_frames, generator=gen).pin_memory()
    d = capi.forward_backward(x.to(DEV), y.to(DEV), pos.to(DEV), pos.to(DEV), 2.0, flags, upstream=up.to(DEV))
    torch.cuda.synchronize()
    monkeypatch.setenv("SOT_HOST_ZEROCOPY", "1")
    for pinned_out in (True, False):
        out = {"loss": torch.full((n_frames,), -1.0), "grad_u": torch.full_like(x, -1.0), "grad_v": torch.full_like(y, -1.0)}
        if pinned_out:
            out = {k: v.pin_memory() for k, v in out.items()}
        loss, gu, gv = capi.loss_grad_host(x, y, pos, pos, 2.0, flags, upstream=up, out=out)
        assert torch.equal(loss, d[0].cpu()) and torch.equal(gu, d[1].cpu()) and torch.equal(gv, d[2].cpu()), pinned_out


@pytest.mark.parametrize("n_bins", [257, 1025, 2049])
@pytest.mark.parametrize("cut", [False, True])
def test_cdf_rows_are_monotone_on_heavy_tailed_spectra(capi, n_bins, cut):
    """The CDF stage sums each thread's bins in fp32 and stitches the threads together with fp64 offsets; the
    rows must come out non-decreasing (the merge relies on it) whatever the dynamic range: 4096 frames of
    log-normal magnitudes over 12 decades with isolated peaks, plus frames of equal bins (long exact ramps)."""
    gen = torch.Generator().manual_seed(n_bins + cut)
    x = torch.exp(6.0 * torch.randn(4096, n_bins, generator=gen) - 8.0)
    y = torch.exp(6.0 * torch.randn(4096, n_bins, generator=gen) - 8.0)
    x[::3, ::37] += 1.0  # peaks that dominate their thread's local sum
    y[::5, 11::53] += 3.0
    x[7], y[7] = 1.0, 0.25  # constant rows
    pos = torch.linspace(0, 1, n_bins)
    flags = capi.SOT_SQUARE | (capi.SOT_CUT_SCALE if cut else 0)
    out = capi.quantiles(x.to(DEV), y.to(DEV), pos.to(DEV), pos.to(DEV), flags)
    cu, cv, qs = out[3], out[4], out[2]
    assert (cu[:, 1:] >= cu[:, :-1]).all() and (cv[:, 1:] >= cv[:, :-1]).all()
    assert (qs[:, 1:] >= qs[:, :-1]).all()
    assert torch.isfinite(cu).all() and torch.isfinite(cv).all()
    # the target row ends within an ulp of 1 (total * fl64(1 / fl32(total)); the reference's sum of x^2 / mass
    # is no closer) and the fp64 CDF is tracked everywhere, also across 12 decades
    assert (cu[:, -1] - 1.0).abs().max().item() <= 1.2e-7
    w = x.double() ** 2
    c64 = torch.cumsum(w, 1) / w.sum(1, keepdim=True)
    # (worst case of the fp32 local sums: ~E/2 ulp for E bins per thread; measured here: up to 7, on the paper's
    # spectra <= 3 -- test_kernel_cdfs_within_ulps_of_fp64)
    assert _ulp_distance(cu.cpu(), c64.float()).max().item() <= 10


def test_randomised_shapes_and_options_vs_oracle(L):
    """60 seeded random problems: unequal supports (1..300 bins), 1..9 frames, shared / per-frame / unsorted
    positions, p in {1, 2, 3.5}, squared or plain magnitudes, cut scaling, exact zeros and duplicated values.
    Loss per frame against the oracle in float64 (no cutoff mask here: it is discontinuous in the last ulp of
    the CDF and has its own tests): rel 2e-5 + 1e-9 (float32 CDFs of up to 300 entries, |dx|^3.5 via powf)."""
    gen = torch.Generator().manual_seed(20240917)
    for trial in range(60):
        N = int(torch.randint(1, 10, (1,), generator=gen))
        n = int(torch.randint(1, 301, (1,), generator=gen))
        m = n if trial % 3 else int(torch.randint(1, 301, (1,), generator=gen))
        p = (1, 2, 3.5)[trial % 3]
        square, cut = bool(trial & 1), bool(trial & 2)
        x = torch.rand(N, n, generator=gen) ** 4
        y = torch.rand(N, m, generator=gen) ** 4
        x[:, ::7] = 0
        if m > 3:
            y[:, 1] = y[:, 2]
        kind = trial % 4
        if kind == 0:  # shared ascending grids
            px, py = torch.sort(torch.rand(n, generator=gen))[0], torch.sort(torch.rand(m, generator=gen))[0]
        elif kind == 1:  # shared, unsorted (require_sort does the work)
            px, py = torch.rand(n, generator=gen), torch.rand(m, generator=gen)
        elif kind == 2:  # per-frame ascending
            px = torch.sort(torch.rand(N, n, generator=gen), dim=1)[0]
            py = torch.sort(torch.rand(N, m, generator=gen), dim=1)[0]
        else:  # per-frame unsorted with duplicated positions
            px, py = torch.rand(N, n, generator=gen).round(decimals=2), torch.rand(N, m, generator=gen).round(decimals=2)
        want = O.sot_per_frame(x.double(), y.double(), px.double(), py.double(), p=p, square=square, cut_scale=cut,
                               limit=False, stable=True)
        got = L.sot_frames(x.to(DEV), y.to(DEV), px.to(DEV), py.to(DEV), p=p, square=square, cut_scale=cut)
        err = (got.double().cpu() - want).abs()
        assert (err <= 2e-5 * want.abs() + 1e-9).all(), (trial, N, n, m, p, square, cut, kind, err.max().item())
