"""Generates tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference/losses.py`)
on small seeded inputs.  Build-container only (the reference tree does not travel); the fixtures
it writes are committed and are what pins `oracle/sot_oracle.py` and the CUDA kernels.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402
from sot_b200 import synthetic as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_module(ref, x, y, pos_x, pos_y, ctor, call=None):
    """Reference forward value, per-frame values, autograd gradients and the quantile taps."""
    call = call or {}
    mod = ref.Wasserstein1D(**ctor)
    xr = x.clone().requires_grad_(True)
    yr = y.clone().requires_grad_(True)
    value = mod(xr, yr, x_pos=pos_x, y_pos=pos_y, **call)
    value.backward()
    F = x.shape[-1]
    with torch.no_grad():
        rows = mod(x.reshape(-1, 1, F), y.reshape(-1, 1, y.shape[-1]), x_pos=pos_x, y_pos=pos_y, dims=1, **call)
        uq, vq, qs, cu, cv = mod(x, y, x_pos=pos_x, y_pos=pos_y, return_quantiles=True, **call)
    return dict(value=value.detach(), rows=rows, grad_x=xr.grad, grad_y=yr.grad, uq=uq, vq=vq, qs=qs, cu=cu, cv=cv)


def save(name, meta, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array(repr(meta)), **arrays)
    print(name, {k: v.shape for k, v in arrays.items()})


def frames(n_fft, n_frames, seed):
    x, y = S.sot_batch(-(-n_frames // 16), n_fft, seed=seed)
    F = x.shape[-1]
    x, y = x.reshape(-1, F)[:n_frames], y.reshape(-1, F)[:n_frames]
    return x.reshape(n_frames // 4, 4, F).contiguous(), y.reshape(n_frames // 4, 4, F).contiguous()


def main():
    ref = reference_loader.load()
    torch.manual_seed(0)
    paper = dict(p=2, square_dist=True, require_sort=True, fixed_x=None, log_scaled_x=False, cumsum_only=False,
                 hinge=False)
    cases = [
        ("sot512_cut", 512, 16, 42, "linear", dict(paper, dont_normalize=True, limit_quantile_range=True)),
        ("sot512_nocut", 512, 16, 123, "linear", dict(paper, dont_normalize=False, limit_quantile_range=False)),
        ("sot512_logf_cut", 512, 16, 456, "logf", dict(paper, dont_normalize=True, limit_quantile_range=True)),
        ("sot2048_cut", 2048, 8, 789, "linear", dict(paper, dont_normalize=True, limit_quantile_range=True)),
        ("sot2048_nocut", 2048, 8, 101112, "linear", dict(paper, dont_normalize=False, limit_quantile_range=False)),
        ("sot2048_logf_unsorted", 2048, 8, 42, "logf", dict(paper, dont_normalize=True, limit_quantile_range=True)),
        ("sot512_p1_nosquare", 512, 16, 7, "linear", dict(p=1, square_dist=False)),
        ("sot512_p3", 512, 8, 8, "linear", dict(p=3, square_dist=True)),
    ]
    for name, n_fft, n_frames, seed, grid, ctor in cases:
        x, y = frames(n_fft, n_frames, seed)
        pos = S.linear_positions(n_fft) if grid == "linear" else S.logf_positions(n_fft)
        out = run_module(ref, x, y, pos, pos.clone(), ctor)
        save(name, dict(ctor=ctor, n_fft=n_fft, grid=grid, seed=seed), x=x, y=y, pos_x=pos, pos_y=pos, **out)

    # metrics.py:144-149 path: fixed_x grid, hann-like smooth spectra, p in {1, 2}, inference only
    g = torch.Generator().manual_seed(5)
    x = torch.rand(6, 5, 65, generator=g) ** 4
    y = torch.rand(6, 5, 65, generator=g) ** 4
    for p in (1, 2):
        mod = ref.Wasserstein1D(p=p, fixed_x=65)
        with torch.no_grad():
            value = mod(x, y)
            rows = mod(x.reshape(-1, 1, 65), y.reshape(-1, 1, 65), dims=1)
            per_item = mod(x, y, dims=1)
        save(f"fixedx65_p{p}", dict(ctor=dict(p=p, fixed_x=65)), x=x, y=y, value=value, rows=rows, per_item=per_item)

    # hinge gate + threshold (losses.py:203-205)
    mod = ref.Wasserstein1D(p=2, fixed_x=65, hinge=True, square_dist=True)
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    value = mod(xr, yr, hinge=0.02)
    value.backward()
    save("fixedx65_hinge", dict(ctor=dict(p=2, fixed_x=65, hinge=True, square_dist=True), call=dict(hinge=0.02)),
         x=x, y=y, value=value.detach(), grad_x=xr.grad, grad_y=yr.grad)

    # module-level wasserstein_1d: n != m, per-row unsorted supports, weights used as given
    g = torch.Generator().manual_seed(11)
    uv, vv = torch.rand(12, 37, generator=g), torch.rand(12, 90, generator=g) * 1.5 - 0.2
    uw = torch.rand(12, 37, generator=g)
    uw = uw / uw.sum(1, keepdim=True)
    vw = torch.rand(12, 90, generator=g)
    vw = vw / vw.sum(1, keepdim=True) * 1.2
    for p, limit in ((1, False), (2, True)):
        uwr, vwr = uw.clone().requires_grad_(True), vw.clone().requires_grad_(True)
        rows = ref.wasserstein_1d(uv, vv, uwr, vwr, p=p, limit_quantile_range=limit)
        rows.sum().backward()
        q = ref.wasserstein_1d(uv, vv, uw, vw, p=p, return_quantiles=True)
        save(f"w1d_n37_m90_p{p}", dict(p=p, limit=limit), u_values=uv, v_values=vv, u_weights=uw, v_weights=vw,
             rows=rows.detach(), grad_uw=uwr.grad, grad_vw=vwr.grad, uq=q[0], vq=q[1], qs=q[2], cu=q[3], cv=q[4])
    rows = ref.wasserstein_1d(uv, vv, p=2)  # uniform weights
    save("w1d_uniform_p2", dict(p=2), u_values=uv, v_values=vv, rows=rows)

    # quantile_function stand-alone
    cws = torch.cumsum(uw, 1)
    qs = torch.sort(torch.rand(12, 50, generator=g) * 1.1, dim=1)[0]
    save("quantile_function", {}, qs=qs, cws=cws, xs=torch.sort(uv, dim=1)[0],
         out=ref.quantile_function(qs, cws, torch.sort(uv, dim=1)[0]))

    # analytic known answers: Diracs at 0.25 / 0.75 on a 9-point grid
    x = torch.zeros(1, 9)
    y = torch.zeros(1, 9)
    x[0, 2] = 1.0
    y[0, 6] = 1.0
    w1 = ref.Wasserstein1D(p=1, fixed_x=9)(x, y)
    w2 = ref.Wasserstein1D(p=2, fixed_x=9)(x, y)
    zero = ref.Wasserstein1D(p=2, fixed_x=9)(x, x)
    dead_cut = ref.Wasserstein1D(p=2, fixed_x=9, dont_normalize=True, limit_quantile_range=True)(torch.zeros(1, 9), y)
    dead_nocut = ref.Wasserstein1D(p=2, fixed_x=9)(torch.zeros(1, 9), y)
    save("kat_diracs", {}, x=x, y=y, w1=w1, w2=w2, zero=zero, dead_cut=dead_cut, dead_nocut=dead_nocut)


if __name__ == "__main__":
    main()
