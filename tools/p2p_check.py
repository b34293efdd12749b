"""Multi-GPU check of the peer-memory all-reduce (run under torchrun on >= 2 GPUs of one box):
values against NCCL, latency of both, and the sharded loss with either collective.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sot_b200 import sharding, synthetic as S  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    red = sharding.PeerReducer()
    ok = True
    for it in range(50):
        vals = torch.tensor([rank + 0.25 * it, 1000.0 + rank, -3.5 * rank], dtype=torch.float64, device=dev)
        mine = red.all_reduce(vals)
        ref = vals.clone()
        dist.all_reduce(ref)
        ok = ok and torch.equal(mine, ref)
    times = {}
    for name, fn in (("p2p", lambda v: red.all_reduce(v)), ("nccl", lambda v: dist.all_reduce(v))):
        v = torch.ones(2, dtype=torch.float64, device=dev)
        for _ in range(20):
            fn(v)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            fn(v)
        b.record()
        b.synchronize()
        times[name] = a.elapsed_time(b) / 200 * 1e3
    # the sharded loss with both collectives
    x, y = S.sot_batch(64, 2048, seed=rank, device=dev)
    pos = S.linear_positions(2048).to(dev)
    out = {}
    for coll in ("nccl", "p2p"):
        fn = sharding.ShardedWasserstein1D(p=2, square_dist=True, collective=coll)
        xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        v = fn(xg, yg, x_pos=pos, y_pos=pos)
        v.backward()
        out[coll] = (v.item(), yg.grad.clone())
    # same loss bits; the gradients' upstream scale is grad * fl32(1/N) on one path and fl32(grad / N) on the other
    same = out["nccl"][0] == out["p2p"][0] and torch.allclose(out["nccl"][1], out["p2p"][1], rtol=1e-6, atol=0)
    flag = torch.tensor([float(ok and same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "values_match_nccl": bool(flag.item()), "us_per_call": times,
                          "loss": out["p2p"][0]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
