"""The SOT-2048 training step of the paper's config with this repository's losses inside a plain PyTorch (DDP) loop
-- SURVEY.md section 8 row f2 in reduced form.

What is the reference's and what is a stand-in:
  * calling convention, loss mix and hyper-parameters: the reference's (`trainer.py:153-245`, SOT-2048
    `train_config.yaml`): target audio -> encoder -> (f0, 20 harmonic amplitudes) -> harmonic synth -> x_hat;
    loss = 0.05 * MSSLoss(x, x_hat) + 1.0 * Wasserstein1D(|STFT x|, |STFT x_hat|) with n_fft 2048, hop 256, flattop
    window, p = 2, squared magnitudes, cutoff mode; Adam, lr 1e-4; one process per GPU, DDP over NCCL.
  * encoder and synthesiser: STAND-INS written for this example (the reference's PESTO encoder on a CQT front end and
    its DDSP synthesiser are outside the hot path and are not reproduced): a small 1-D CNN over log-magnitude frames
    with a soft-argmax pitch head, and a stationary harmonic oscillator bank.

    python examples/train_step.py [--steps 30] [--batch 256]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py
Prints one JSON line: steps/s, frames/s (16 frames per signal), first and last loss.
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sot_b200 import features, mss, synthetic as S  # noqa: E402

SR, N_FFT, HOP, N_MODES = 16000, 2048, 256, 20
F_MIN, F_MAX, N_PITCH = 32.7, 2000.0, 256


class PitchEncoder(torch.nn.Module):
    """log|STFT| frames (B, T, 1025) -> per-signal f0 (soft-argmax over a log-frequency grid) and amplitudes."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(
            torch.nn.Conv1d(1, 16, 15, stride=2, padding=7), torch.nn.GELU(),
            torch.nn.Conv1d(16, 32, 15, stride=2, padding=7), torch.nn.GELU(),
            torch.nn.Conv1d(32, 32, 15, stride=2, padding=7), torch.nn.GELU(),
            torch.nn.AdaptiveAvgPool1d(32), torch.nn.Flatten(), torch.nn.Linear(32 * 32, 256), torch.nn.GELU())
        self.pitch = torch.nn.Linear(256, N_PITCH)
        self.amps = torch.nn.Linear(256, N_MODES)
        self.register_buffer("grid", torch.exp(torch.linspace(math.log(F_MIN), math.log(F_MAX), N_PITCH)))

    def forward(self, mag):
        B, T, F = mag.shape
        h = self.net(torch.log(mag.reshape(B * T, 1, F) + 1e-5)).reshape(B, T, -1).mean(1)
        f0 = (torch.softmax(self.pitch(h) / 0.1, dim=-1) * self.grid).sum(-1, keepdim=True)  # temperature 0.1
        return f0, torch.sigmoid(self.amps(h)) * 0.5


def harmonic_synth(f0, amps, n_samples=S.N_SAMPLES):
    order = torch.arange(1, N_MODES + 1, device=f0.device)
    freqs = f0 * order
    amps = amps * (freqs < SR / 2)
    t = torch.arange(n_samples, device=f0.device, dtype=torch.float32) / SR
    return (amps.unsqueeze(-1) * torch.sin(2 * math.pi * freqs.unsqueeze(-1) * t)).sum(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", dest="signals", type=int, default=256, help="signals per GPU and step (16 frames each)")
    ap.add_argument("--lr", type=float, default=1e-4, help="Adam learning rate (paper config: 1e-4)")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = PitchEncoder().to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=1e-4)
    transform = features.TorchSTFT(n_fft=N_FFT, hop_length=HOP, window="flattop", sr=SR)
    sot = features.Wasserstein1DWithTransform(p=2, square_dist=True, dont_normalize=True, limit_quantile_range=True,
                                              transform_kwargs=dict(type="stft", n_fft=N_FFT, hop_length=HOP,
                                                                    window="flattop")).to(dev)
    mix = mss.MixOfLosses([mss.MSSLoss(mag_weight=1.0, logmag_weight=0.0), sot], [0.05, 1.0])
    gen = torch.Generator().manual_seed(42 + rank)
    x, _ = S.harmonic_signals(args.signals, gen, device=dev)  # one fixed batch: the loss must go down on it

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            mag = transform(x)
        f0, amps = model(mag)
        x_hat = harmonic_synth(f0, amps)
        parts = mix(x, x_hat)  # {"MSSLoss": ..., "Wasserstein1DWithTransform": ...} already weighted
        loss = sum(parts.values())
        loss.backward()
        opt.step()
        return loss.detach()

    losses = [step() for _ in range(args.warmup)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    losses += [step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "steps_per_s": args.steps / dt.item(),
                          "frames_per_s": args.steps * args.signals * 16 * world / dt.item(),
                          "signals_per_gpu": args.signals, "first_loss": torch.stack(losses[:3]).mean().item(),
                          "last_loss": torch.stack(losses[-3:]).mean().item(),
                          "losses": "0.05 * MSSLoss (6 FFT sizes) + 1.0 * SOT-2048 (cutoff), both on complex STFT frames"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
