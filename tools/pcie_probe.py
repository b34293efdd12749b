"""Host<->device copy rates of this box with pinned memory (what bounds bench.py's `e2e` line): H2D alone, D2H alone,
and both directions at once on two streams -- on one GPU, or, under torchrun, on all GPUs of the box AT THE SAME
TIME (every rank copies concurrently between barriers), which is what the sharded end-to-end path does.

    python tools/pcie_probe.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py [--bind]
`--bind` pins each rank to the cores of its GPU's NUMA node (or an own slice of the cores) first, like bench.py.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(mb=512, reps=6):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bind = None
    if "--bind" in sys.argv:
        from sot_b200 import hostbind
        bind = hostbind.bind_to_gpu_node(local)
    n = mb * (1 << 20) // 4
    h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the slowest rank: all ranks were copying together
        return t.item()

    run(True, True)
    gb = mb / 1024
    out = {"ranks_copying_at_once": world, "mb_per_copy": mb, "bind": bind,
           "h2d_gbs_per_gpu": gb / run(True, False), "d2h_gbs_per_gpu": gb / run(False, True)}
    out["both_each_gbs_per_gpu"] = gb / run(True, True)
    out["aggregate_both_directions_gbs"] = 2 * world * out["both_each_gbs_per_gpu"]
    try:
        out["cpus"] = os.cpu_count()
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        pass
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
