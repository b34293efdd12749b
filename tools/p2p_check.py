"""Multi-GPU check of the sharded loss (run under torchrun on >= 2 GPUs of one box):

  * the peer-memory all-reduce against NCCL, bit for bit, and the latency of both;
  * `ShardedWasserstein1D` -- the fused path `bench.py` runs -- with every exchange ("nccl" / "p2p", collected in
    stream or overlapped on a side stream, one-pass or recompute backward): the value against the mean that ONE GPU
    computes over all shards, the gradients against the single-GPU gradients of this rank's rows;
  * replay of the whole step from a CUDA graph with the peer-memory exchange (sequence numbers live on the device);
  * one rank arriving 3 s late at the exchange: the others wait, the value is right;
  * unequal shards under `equal_shards=True`: NaN loss and an error on the next call; `equal_shards=False`: correct.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sot_b200 import losses, sharding, synthetic as S  # noqa: E402


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

    # ---- the sharded loss: every rank's shard is known to every rank (seeded), so each can check on its own ----
    kw = dict(p=2, square_dist=True, dont_normalize=True, limit_quantile_range=True)
    shards = [S.sot_batch(16, 2048, seed=100 + r, device=dev) for r in range(world)]
    pos = S.linear_positions(2048).to(dev)
    single = losses.Wasserstein1D(**kw)
    with torch.no_grad():
        want = sum(float(single(xs, ys, x_pos=pos, y_pos=pos).item()) for xs, ys in shards) / world
    x, y = shards[rank]
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    (single(xr, yr, x_pos=pos, y_pos=pos) * 3.0 / world).backward()  # d(3 * global mean) / d(my rows)
    results, worst_val, worst_grad = {}, 0.0, 0.0
    for coll in ("nccl", "p2p"):
        for overlap in (False, True):
            for mode in ("onepass", "recompute"):
                fn = sharding.ShardedWasserstein1D(collective=coll, overlap_exchange=overlap, backward_mode=mode, **kw)
                for _ in range(3):  # several calls: the two-phase mailboxes and the run-ahead bound are exercised
                    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
                    v = fn(xg, yg, x_pos=pos, y_pos=pos)
                    (3.0 * v).backward()
                fn.wait_value(v)
                e_val = abs(v.item() - want) / abs(want)
                e_grad = max(((xg.grad - xr.grad).norm() / xr.grad.norm()).item(),
                             ((yg.grad - yr.grad).norm() / yr.grad.norm()).item())
                worst_val, worst_grad = max(worst_val, e_val), max(worst_grad, e_grad)
                results[f"{coll}/{'overlap' if overlap else 'instream'}/{mode}"] = [e_val, e_grad, fn.exchange.collective_used]
    sharded_ok = worst_val <= 2e-6 and worst_grad <= 2e-6

    # ---- a rank that arrives seconds late (checkpointing, evaluation, a data-loader stall): the others wait, the
    #      value is right -- never a silent NaN (the old 2 s timeout poisoned the loss) ----
    import time
    fn = sharding.ShardedWasserstein1D(collective="p2p", **kw)
    fn(x, y, x_pos=pos, y_pos=pos)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == world - 1:
        time.sleep(3.0)
    v_late = fn(x, y, x_pos=pos, y_pos=pos)
    late_ok = abs(v_late.item() - want) <= 2e-6 * abs(want)

    # ---- CUDA graph of the whole step with the peer-memory exchange ----
    fn = sharding.ShardedWasserstein1D(collective="p2p", **kw)
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    for _ in range(3):
        xg.grad = None
        yg.grad = None
        fn(xg, yg, x_pos=pos, y_pos=pos).backward()
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    xg.grad = None
    yg.grad = None
    with torch.cuda.graph(g):
        v = fn(xg, yg, x_pos=pos, y_pos=pos)
        v.backward()
    graph_vals = []
    for _ in range(4):
        g.replay()
        graph_vals.append(v.item())
    graph_ok = all(abs(t - want) <= 2e-6 * abs(want) for t in graph_vals) and \
        ((yg.grad * 3.0 - yr.grad).norm() / yr.grad.norm()).item() <= 2e-6

    # ---- unequal shards ----
    n_mine = x.shape[0] - (4 if rank == 0 else 0)  # rank 0 holds 4 signals (64 frames) fewer
    xu, yu = x[:n_mine], y[:n_mine]
    with torch.no_grad():
        sums = [float(single(xs[:xs.shape[0] - (4 if r == 0 else 0)], ys[:ys.shape[0] - (4 if r == 0 else 0)],
                             x_pos=pos, y_pos=pos).item()) * 16 * (xs.shape[0] - (4 if r == 0 else 0))
                for r, (xs, ys) in enumerate(shards)]
        n_all = sum(16 * (xs.shape[0] - (4 if r == 0 else 0)) for r, (xs, ys) in enumerate(shards))
    want_unequal = sum(sums) / n_all
    safe = sharding.ShardedWasserstein1D(collective="p2p", equal_shards=False, **kw)
    v_safe = safe(xu, yu, x_pos=pos, y_pos=pos)
    unequal_ok = abs(v_safe.item() - want_unequal) <= 2e-6 * abs(want_unequal)
    optimistic = sharding.ShardedWasserstein1D(collective="p2p", **kw)
    v_opt = optimistic(xu, yu, x_pos=pos, y_pos=pos)
    torch.cuda.synchronize()
    raised = False
    try:
        optimistic(xu, yu, x_pos=pos, y_pos=pos)
    except RuntimeError:
        raised = True
    unequal_ok = unequal_ok and bool(torch.isnan(v_opt).item()) and raised

    flag = torch.tensor([float(ok and sharded_ok and graph_ok and unequal_ok and late_ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "values_match_nccl": bool(ok), "all_ok": bool(flag.item()),
                          "sharded_loss_ok": sharded_ok, "graph_replay_ok": graph_ok, "unequal_shards_ok": unequal_ok, "late_rank_ok": late_ok,
                          "us_per_call": times, "want": want, "worst_value_rel": worst_val,
                          "worst_grad_rel_l2": worst_grad, "cases": results, "graph_values": graph_vals}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
