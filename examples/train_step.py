"""The SOT-2048 training step of the paper's config (config #5 of BASELINE.json) with this
repository's losses inside a plain PyTorch DDP loop (the reference drives it with Lightning, absent from this image).

What is the reference's and what is not:
  * the step: `trainer.py:77-143, 153-245` -- target audio -> front end -> encoder -> (pitch, 20 harmonic weights per
    frame) -> harmonic synthesiser -> x_hat; loss = 0.05 * MSSLoss(x, x_hat) + 1.0 * Wasserstein1D(|STFT x|, |STFT x_hat|)
    with n_fft 2048, hop 256, flattop window, p = 2, squared magnitudes, cutoff mode; Adam lr 1e-4, weight decay 1e-4
    (SOT-2048 `train_config.yaml:24-120`); one process per GPU, DDP over NCCL (the reference: `devices: 1`);
  * `--model reference` (default when the files are there): the reference's OWN, unmodified `encoder.PESTOEncoder`
    (46 012 parameters, `encoder.py:73-365`) and `synths.Sinusoidal` / `ddsp.py` oscillator bank (`synths.py:46-128`),
    imported from `/root/reference` or from the verbatim git-ignored copy `baseline/_ref/` that travels to the GPU box;
  * the CQT front end is a SUBSTITUTE: the reference uses nnAudio's CQT (`features.py:116-188`; 285 bins, 36 per octave
    from 32.7 Hz, hop 256), which is not installed here.  `LogFrequencyFrontEnd` gives the encoder the same input shape
    and frequency axis from an STFT magnitude interpolated onto those 285 log-spaced frequencies;
  * `--model standin`: a small CNN + stationary oscillator bank written for this example (no reference files needed).
  * both losses are this repository's CUDA kernels, straight from the complex STFT frames.

    python examples/train_step.py [--steps 30] [--batch 64] [--model reference|standin] [--loss-share]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py
Prints one JSON line: steps/s, frames/s (16 frames per signal), first and last loss, and with `--loss-share` the share
of the step spent in the two losses (the same step with the loss replaced by mean(x_hat^2)).
"""
import argparse
import json
import math
import os
import sys
import time
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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


def reference_root():
    for r in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(r, "encoder.py")) and os.path.isfile(os.path.join(r, "synths.py")):
            return r
    return None


class LogFrequencyFrontEnd(torch.nn.Module):
    """Stand-in for the reference's CQT feature extractor (`features.py:116-188`, nnAudio): (B, samples) ->
    (B, frames, 285) magnitudes on the CQT's frequency axis f_k = 32.7 * 2**(k / 36) Hz, hop 256, centred frames --
    an STFT magnitude (n_fft 2048, hann) linearly interpolated onto those frequencies."""

    def __init__(self, n_bins=285, bins_per_octave=36, fmin=32.7, n_fft=2048, hop=HOP, sr=SR):
        super().__init__()
        self.n_fft, self.hop = n_fft, hop
        self.frequencies = fmin * 2.0 ** (torch.arange(n_bins, dtype=torch.float64) / bins_per_octave)
        pos = (self.frequencies * n_fft / sr).clamp(max=n_fft // 2 - 1e-6)
        lo = pos.floor().long()
        frac = (pos - lo).float()
        w = torch.zeros(n_fft // 2 + 1, n_bins)
        w[lo, torch.arange(n_bins)] = 1.0 - frac
        w[lo + 1, torch.arange(n_bins)] += frac
        self.register_buffer("interp", w)
        self.register_buffer("window", torch.hann_window(n_fft))

    def get_frequencies(self):
        return self.frequencies.numpy()

    def forward(self, audio):
        spec = torch.stft(audio, self.n_fft, self.hop, window=self.window, center=True, pad_mode="constant",
                          return_complex=True)
        return spec.abs().transpose(1, 2) @ self.interp  # (B, frames, 285)


class ReferenceModel(torch.nn.Module):
    """`Trainer.encode` + `Trainer.decode` (trainer.py:77-143) around the reference's unmodified encoder and synth."""

    def __init__(self, root):
        super().__init__()
        for name in ("nnAudio", "nnAudio.features", "librosa"):  # imported at features.py:7,10; absent, unused here
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]
        if root not in sys.path:
            sys.path.append(root)
        import encoder as ref_encoder  # noqa: the reference's modules, unmodified
        import synths as ref_synths
        import utils as ref_utils
        self.unit_to_hz = ref_utils.unit_to_hz
        self.front = LogFrequencyFrontEnd()
        self.encoder = ref_encoder.PESTOEncoder(
            n_modes=N_MODES, estimation_type="soft-argmax", output_splits=["frequency", "weights"], harmonic=True,
            feature_size=512, output_size=285, n_chan_input=1, n_chan_layers=[40, 30, 30, 10, 3], n_prefilt_layers=2,
            residual=True, n_bins_in=285, activation_fn="leaky", num_output_layers=1, a_lrelu=0.3, kernel_size=15)
        self.decoder = ref_synths.Sinusoidal(n_samples=S.N_SAMPLES, sample_rate=SR, amp_scale_fn=None,
                                             freq_scale_fn=None, harmonic=True, amp_resample_method="window",
                                             apply_roll_off=False)
        f = self.front.get_frequencies()
        self.f_lo, self.f_hi = float(f[0]), float(f[-1])  # freq_hz_min / max = "auto" (trainer.py:65-70)

    def forward(self, x):
        feats = self.front(x[:, :-1])  # trainer.py:78
        batch, frames, freq = feats.shape
        z = self.encoder(feats.reshape(batch * frames, freq).unsqueeze(1))
        pitch_unit = self.encoder.predict_pitch(z["frequency"], temperature=0.1)["pitch_unit"]
        pitch_hz = self.unit_to_hz(pitch_unit, self.f_lo, self.f_hi).reshape(batch, frames, -1)
        weights = z["weights"].reshape(batch, frames, -1)
        return self.decoder(weights, pitch_hz)  # (B, samples)


class StandInModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = PitchEncoder()
        self.transform = features.TorchSTFT(n_fft=N_FFT, hop_length=HOP, window="flattop", sr=SR)

    def forward(self, x):
        with torch.no_grad():
            mag = self.transform(x)
        f0, amps = self.encoder(mag)
        return harmonic_synth(f0, amps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", dest="signals", type=int, default=64,
                    help="signals per GPU and step, 16 frames each (paper config: 64)")
    ap.add_argument("--lr", type=float, default=1e-4, help="Adam learning rate (paper config: 1e-4)")
    ap.add_argument("--model", default="auto", choices=["auto", "reference", "standin"])
    ap.add_argument("--loss-share", action="store_true", help="also time the step without the two losses")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    root = reference_root()
    kind = args.model if args.model != "auto" else ("reference" if root is not None else "standin")
    if kind == "reference" and root is None:
        raise SystemExit("train_step.py: --model reference needs /root/reference or baseline/_ref/ (python -c "
                         "'import __graft_entry__ as g; g.build()' stages the copy in the build container)")
    model = (ReferenceModel(root) if kind == "reference" else StandInModel()).to(dev)
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=1e-4)
    sot = features.Wasserstein1DWithTransform(p=2, square_dist=True, dont_normalize=True, limit_quantile_range=True,
                                              transform_kwargs=dict(type="stft", n_fft=N_FFT, hop_length=HOP,
                                                                    window="flattop")).to(dev)
    mix = mss.MixOfLosses([mss.MSSLoss(mag_weight=1.0, logmag_weight=0.0), sot], [0.05, 1.0])
    gen = torch.Generator().manual_seed(42 + rank)
    x, _ = S.harmonic_signals(args.signals, gen, device=dev)  # one fixed batch: the loss must go down on it

    def step(with_losses=True):
        opt.zero_grad(set_to_none=True)
        x_hat = model(x)
        if with_losses:
            parts = mix(x, x_hat)  # {"MSSLoss": ..., "Wasserstein1DWithTransform": ...} already weighted
            loss = sum(parts.values())  # trainer.py:233-238
        else:
            loss = x_hat.pow(2).mean()
        loss.backward()
        opt.step()
        return loss.detach()

    def timed(n, **kw):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = [step(**kw) for _ in range(n)]
        b.record()
        b.synchronize()
        dt = torch.tensor([a.elapsed_time(b) * 1e-3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return out, dt.item()

    losses = [step() for _ in range(args.warmup)]
    more, dt = timed(args.steps)
    losses += more
    share = None
    if args.loss_share:
        for _ in range(3):
            step(with_losses=False)
        _, dt_bare = timed(args.steps, with_losses=False)
        share = {"ms_per_step_without_the_losses": 1e3 * dt_bare / args.steps,
                 "share_of_step_in_the_two_losses": max(0.0, 1.0 - dt_bare / dt)}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "model": kind, "trainable_parameters": n_params,
                          "steps_per_s": args.steps / dt, "ms_per_step": 1e3 * dt / args.steps,
                          "frames_per_s": args.steps * args.signals * 16 * world / dt,
                          "signals_per_gpu": args.signals, "first_loss": torch.stack(losses[:3]).mean().item(),
                          "last_loss": torch.stack(losses[-3:]).mean().item(), "loss_share": share,
                          "losses": "0.05 * MSSLoss (6 FFT sizes) + 1.0 * SOT-2048 (cutoff), both on complex STFT frames",
                          "front_end": "log-frequency STFT interpolation (substitute for nnAudio CQT)"
                          if kind == "reference" else "flattop STFT magnitude"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
