#!/usr/bin/env python
"""bench.py -- SOT loss forward+backward throughput (frames/s) on 1..8 B200, with roofline,
end-to-end (host buffers through the C ABI) and CPU-baseline legs.  Prints ONE JSON line (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" = one `loss = Wasserstein1D(...)(x, y, x_pos, y_pos); loss.backward()` over this rank's
shard of the workload with BOTH spectra requiring gradients (24*F+8 algorithmic bytes per frame,
SURVEY.md section 8d).  Weak scaling: every rank owns `--frames` frames (default 65,536 x 1025 bins,
BASELINE.json config "SOT-NoCut large sweep"); inputs + gradients per rank are ~1 GiB, far larger
than the 126 MB L2, so no L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SOT loss fwd+bwd frames/s (n_fft=2048)"
WORKLOADS = {
    # name: (n_fft, cut, grid)
    "sot2048-nocut-sweep": (2048, False, "linear"),  # BASELINE.json configs[3] (the multi-GPU sweep)
    "sot2048-cut": (2048, True, "linear"),           # paper config SOT-2048
    "sot512-cut": (512, True, "linear"),             # configs[1]
    "sot512-logf-cut": (512, True, "logf"),          # configs[2]
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sot2048-nocut-sweep", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=65536, help="frames per GPU (weak scaling)")
    ap.add_argument("--mode", default=None, choices=["recompute", "fused"], help="backward mode (default: library default)")
    ap.add_argument("--tuning", default=None, help="threads_per_frame,bins_per_thread[,chains] kernel override")
    ap.add_argument("--collective", default="auto", choices=["auto", "nccl", "p2p"],
                    help="the scalar all-reduce of the sharded mean: NCCL or the NVLink peer-memory kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=1024, help="frames per CPU-baseline step (64 signals x 16)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_leg(n_fft, cut, grid, frames, steps, warmup, seconds_cap=25.0):
    """The oracle (torch-CPU restatement of losses.py:129-313, bit-identical to the reference in
    the build container) timed on this box's host cores: forward + backward, both grads."""
    from oracle import sot_oracle as O
    from sot_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = S.sot_batch(max(1, frames // 16), n_fft, seed=42)
    pos = S.linear_positions(n_fft) if grid == "linear" else S.logf_positions(n_fft)
    kw = dict(p=2, square=True, cut_scale=cut, limit=cut)

    def step():
        xr = x.clone().requires_grad_(True)
        yr = y.clone().requires_grad_(True)
        O.sot_loss(xr, yr, pos, pos.clone(), **kw).backward()

    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > seconds_cap:
            break
    n = x.shape[0] * x.shape[1]
    total = sum(times)
    return {"value": n * len(times) / total, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} steps x {n} frames x {x.shape[-1]} bins (oracle/sot_oracle.py, torch CPU ops, "
                      f"fwd+bwd, both grads), median {sorted(times)[len(times) // 2] * 1e3:.1f} ms/step"}, total / len(times)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_fft, cut, grid = WORKLOADS[args.workload]
    F = n_fft // 2 + 1
    config = {"workload": args.workload, "frames_per_gpu": args.frames, "bins": F, "p": 2, "square_dist": True,
              "cutoff": cut, "positions": grid, "grads": "both spectra",
              "l2": "inputs+grads per GPU ~%.0f MB >> 126 MB L2, no flush needed" % (16 * F * args.frames / 1e6),
              "parallelism": f"frames sharded over {world} rank(s), one scalar all-reduce ({args.collective})"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb, ms = cpu_reference_leg(n_fft, cut, grid, args.cpu_frames, args.steps, max(args.warmup, 1))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config, "cpu_baseline": cb, "gpu_launches": 0,
                          "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    from sot_b200 import _capi, sharding, synthetic as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the SOT kernels have no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.tuning:
        _capi.set_tuning(*[int(t) for t in args.tuning.split(",")])

    # ---- inputs: synthetic harmonic spectra, resident in HBM ----------------------------------
    signals = -(-args.frames // 16)
    x, y = S.sot_batch(signals, n_fft, seed=42 + rank, device=dev)
    x = x.reshape(-1, F)[:args.frames].contiguous()
    y = y.reshape(-1, F)[:args.frames].contiguous()
    pos = (S.linear_positions(n_fft) if grid == "linear" else S.logf_positions(n_fft)).to(dev)
    pos_y = pos.clone()
    kw = dict(p=2, square_dist=True, dont_normalize=cut, limit_quantile_range=cut)
    if args.mode:
        kw["backward_mode"] = args.mode
    loss_fn = sharding.ShardedWasserstein1D(collective=args.collective, **kw)
    xg = x.clone().requires_grad_(True)
    yg = y.clone().requires_grad_(True)

    def step():
        xg.grad = None
        yg.grad = None
        value = loss_fn(xg, yg, x_pos=pos, y_pos=pos_y)
        value.backward()
        return value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        value = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()  # `ncu --profile-from-start off` sees exactly the timed steps
    ev0.record()
    for _ in range(args.steps):
        value = step()
    ev1.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    launches = _capi.launch_count() - launches0
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    frames_s = args.frames * world / (ms_step * 1e-3)

    # ---- dominant kernel alone (fused forward+backward launch), CUDA events on its stream ----
    flags = _capi.SOT_SQUARE | (_capi.SOT_CUT_SCALE | _capi.SOT_LIMIT if cut else 0)
    if grid == "linear":  # what the module detects for these grids (exact i * 2**-k positions)
        flags |= _capi.SOT_UNIFORM_GRID
    up = torch.full((args.frames,), 1.0 / (args.frames * world), device=dev)
    # the launch the timed step's backward makes: upstream scalar on the device, merge-path co-ranks saved by
    # the forward launch (`fused` mode has no separate backward launch: time the plain fused entry point)
    scale_dev = torch.full((1,), 1.0 / (args.frames * world), device=dev)
    coranks = _capi.forward_sum(x, y, pos, pos_y, 2.0, flags, save_coranks=True)[2]
    k_ms = []
    for it in range(3 + args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if loss_fn.backward_mode == "recompute":
            _capi.forward_backward_scaled(x, y, pos, pos_y, 2.0, flags, scale_dev, coranks=coranks)
        else:
            _capi.forward_backward(x, y, pos, pos_y, 2.0, flags, upstream=up, want_loss=False)
        b.record()
        b.synchronize()
        if it >= 3:
            k_ms.append(a.elapsed_time(b))
    f_ms = []
    for it in range(3 + args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if loss_fn.backward_mode == "recompute":  # the step's forward launch: sum on the device, co-ranks saved
            _capi.forward_sum(x, y, pos, pos_y, 2.0, flags, save_coranks=True)
        else:
            _capi.forward(x, y, pos, pos_y, 2.0, flags)
        b.record()
        b.synchronize()
        if it >= 3:
            f_ms.append(a.elapsed_time(b))
    sampler.stop_flag.set()
    sampler.join()
    k_avg, f_avg = sum(k_ms) / len(k_ms), sum(f_ms) / len(f_ms)
    peak, peak_src = peaks()
    bytes_bwd = (16 * F + 4) * args.frames  # reads u, v, upstream; writes grad_u, grad_v (co-ranks: 128 B, not counted)
    bytes_fwd = (8 * F + 4) * args.frames   # reads u, v; writes loss
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("backward_kernel_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "sot_frames_kernel<OUT_GRAD> (fused forward+backward launch)",
                "achieved": bytes_bwd / (k_avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": bytes_bwd / (k_avg * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_bwd, "kernel_ms": k_avg,
                "forward_kernel": {"ms": f_avg, "achieved": bytes_fwd / (f_avg * 1e-3) / 1e9,
                                   "frac": bytes_fwd / (f_avg * 1e-3) / 1e9 / peak,
                                   "algorithmic_bytes_per_launch": bytes_fwd},
                "step": {"algorithmic_bytes_per_frame": 24 * F + 8,
                         "achieved": (24 * F + 8) * args.frames / (ms_step * 1e-3) / 1e9,
                         "frac": (24 * F + 8) * args.frames / (ms_step * 1e-3) / 1e9 / peak}}

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        hx, hy = x.cpu().pin_memory(), y.cpu().pin_memory()
        hpos = pos.cpu()
        hup = up.cpu().pin_memory()
        out = {"loss": torch.empty(args.frames).pin_memory(), "grad_u": torch.empty_like(hx).pin_memory(),
               "grad_v": torch.empty_like(hy).pin_memory()}
        for _ in range(2):
            _capi.loss_grad_host(hx, hy, hpos, hpos, 2.0, flags, upstream=hup, device=local, out=out)
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            _capi.loss_grad_host(hx, hy, hpos, hpos, 2.0, flags, upstream=hup, device=local, out=out)
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the host path must give the device path's numbers
        check = _capi.forward_backward(x, y, pos, pos_y, 2.0, flags, upstream=up)
        torch.cuda.synchronize()
        assert torch.equal(check[0].cpu(), out["loss"]) and torch.equal(check[1].cpu(), out["grad_u"])
        e2e = {"value": args.frames * world * n_e2e / t.item(), "unit": "frames/s",
               "h2d_bytes_per_step": 4 * args.frames * (2 * F + 1) + 8 * F,
               "d2h_bytes_per_step": 4 * args.frames * (2 * F + 1), "steps": n_e2e,
               "api": "sot_loss_grad_host (C ABI, pinned host buffers, chunked 3-stream pipeline)"}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb, _ = cpu_reference_leg(n_fft, cut, grid, args.cpu_frames, 30, 2, seconds_cap=20.0)

    if rank == 0:
        line = {"metric": METRIC, "value": frames_s, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": sampler.summary(),
                "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "cpu_baseline": cb,
                "loss": float(value.item()), "backward_mode": loss_fn.backward_mode}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
