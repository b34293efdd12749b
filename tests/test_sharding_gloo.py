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
