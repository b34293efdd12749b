"""The multi-GPU path's host logic on CPU: world_size 2 over gloo.  The per-frame values come
from the oracle here (no GPU); what is under test is the split, the single scalar all-reduce and
the 1/N_global gradient scaling of `sot_b200.sharding`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sot_oracle as O
from sot_b200 import sharding
from tests import golden_io as G


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        g = G.load("sot512_nocut")
        x, y, pos = g["x"].reshape(-1, 257), g["y"].reshape(-1, 257), g["pos_x"]
        lo = sum(counts[:rank])
        hi = lo + counts[rank]
        xs = x[lo:hi].clone().requires_grad_(True)
        ys = y[lo:hi].clone().requires_grad_(True)
        rows = O.sot_per_frame(xs, ys, pos, pos, p=2, square=True)
        value = sharding.global_mean(rows)
        (2.0 * value).backward()
        torch.save(dict(value=value.detach(), gx=xs.grad, gy=ys.grad, lo=lo, hi=hi), f"{out_dir}/r{rank}.pt")
    finally:
        dist.destroy_process_group()


def _fake_kernels(monkeypatch_target):
    """Stand-ins for the two CUDA entry points `_SotMean` calls, with the oracle's arithmetic (CPU tests only): what
    is under test is the HOST logic around the launch -- counts, 1/N_global, the exchange, the in-place backward."""
    def mean_step(u, v, pos_u, pos_v, p, flags, grad_scale, mean_scale, want_gu=True, want_gv=True, want_rows=False,
                  total_out=None, count_value=0.0, post=None, grad_scale_device=None, want_mean=None):
        assert post is None, "the CPU tests run the all-reduce form of the exchange"
        with torch.enable_grad():  # (autograd is off inside Function.forward)
            ur, vr = u.detach().clone().requires_grad_(True), v.detach().clone().requires_grad_(True)
            rows = O.sot_per_frame(ur, vr, pos_u, pos_v, p=p, square=bool(flags & 1))
            total = rows.sum()
            gu, gv = torch.autograd.grad(total, (ur, vr))
        scale = grad_scale * (float(grad_scale_device) if grad_scale_device is not None else 1.0)
        if total_out is not None:
            total_out[0] = total.detach().double()
            total_out[1] = count_value
        mean = (total.detach() * mean_scale).float() if (want_mean or (want_mean is None and total_out is None)) else None
        return mean, None, (gu * scale if want_gu else None), (gv * scale if want_gv else None)

    def scale_inplace(a, b, scale):
        for t in (a, b):
            if t is not None:
                t.mul_(float(scale))

    monkeypatch_target.mean_step = mean_step
    monkeypatch_target.scale_inplace = scale_inplace


def _fused_worker(rank, world, port, counts, equal_shards, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        from sot_b200 import _capi, losses
        _fake_kernels(_capi)
        g = G.load("sot512_nocut")
        x, y, pos = g["x"].reshape(-1, 257), g["y"].reshape(-1, 257), g["pos_x"]
        lo = sum(counts[:rank])
        hi = lo + counts[rank]
        exchange = sharding.MeanExchange(None, "auto", overlap=False, equal_shards=equal_shards)
        out = {}
        failed_next = False
        for mode in ("onepass", "recompute"):
            xs = x[lo:hi].clone().requires_grad_(True)
            ys = y[lo:hi].clone().requires_grad_(True)
            try:
                value = losses._SotMean.apply(xs, ys, pos, pos, 2.0, 1, exchange, mode)
            except RuntimeError:  # the call AFTER one whose counts did not add up (both ranks raise: no collective yet)
                failed_next = True
                break
            (2.0 * value).backward()
            out[mode] = dict(value=value.detach(), gx=xs.grad, gy=ys.grad)
        torch.save(dict(out=out, lo=lo, hi=hi, used=exchange.collective_used, failed_next=failed_next),
                   f"{out_dir}/f{rank}.pt")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts,equal_shards", [((8, 8), True), ((11, 5), False), ((11, 5), True)])
def test_fused_mean_host_logic_on_two_ranks(tmp_path, counts, equal_shards):
    """`losses._SotMean` with a `sharding.MeanExchange` on 2 gloo ranks (the path bench.py runs on the GPUs, with the
    kernels replaced by the oracle): global mean, gradients scaled by 1/N_global inside the "launch", one exchange;
    unequal shards are caught when `equal_shards=True` was promised (NaN value, error on the next call)."""
    mp.spawn(_fused_worker, args=(2, _free_port(), counts, equal_shards, str(tmp_path)), nprocs=2, join=True)
    g = G.load("sot512_nocut")
    x = g["x"].reshape(-1, 257).clone().requires_grad_(True)
    y = g["y"].reshape(-1, 257).clone().requires_grad_(True)
    full = O.sot_per_frame(x, y, g["pos_x"], g["pos_x"], p=2, square=True).mean()
    (2.0 * full).backward()
    consistent = equal_shards is False or counts[0] == counts[1]
    for r in range(2):
        d = torch.load(f"{tmp_path}/f{r}.pt")
        assert d["used"] == "gloo"
        assert ("recompute" in d["out"]) == consistent
        for mode in d["out"]:
            o = d["out"][mode]
            if consistent:
                assert o["value"].item() == pytest.approx(full.item(), rel=1e-6)
                assert torch.allclose(o["gx"], x.grad[d["lo"]:d["hi"]], rtol=1e-5, atol=0)
                assert torch.allclose(o["gy"], y.grad[d["lo"]:d["hi"]], rtol=1e-5, atol=0)
            else:
                assert torch.isnan(o["value"]).item(), "a broken equal_shards promise must not go unnoticed"
        assert d["failed_next"] == (not consistent)


@pytest.mark.parametrize("counts", [(8, 8), (11, 5)])
def test_two_rank_mean_and_gradients_match_single_process(tmp_path, counts):
    mp.spawn(_worker, args=(2, _free_port(), counts, str(tmp_path)), nprocs=2, join=True)
    g = G.load("sot512_nocut")
    x = g["x"].reshape(-1, 257).clone().requires_grad_(True)
    y = g["y"].reshape(-1, 257).clone().requires_grad_(True)
    full = O.sot_per_frame(x, y, g["pos_x"], g["pos_x"], p=2, square=True).mean()
    (2.0 * full).backward()
    for r in range(2):
        d = torch.load(f"{tmp_path}/r{r}.pt")
        assert d["value"].item() == pytest.approx(full.item(), rel=1e-6)
        assert torch.allclose(d["gx"], x.grad[d["lo"]:d["hi"]], rtol=1e-6, atol=0)
        assert torch.allclose(d["gy"], y.grad[d["lo"]:d["hi"]], rtol=1e-6, atol=0)


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 64, 4096):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_global_mean_is_plain_mean():
    rows = torch.rand(13, requires_grad=True)
    v = sharding.global_mean(rows)
    v.backward()
    assert v.item() == pytest.approx(rows.detach().mean().item(), rel=1e-6)
    assert torch.allclose(rows.grad, torch.full((13,), 1 / 13))
