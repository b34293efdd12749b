"""Frame sharding across the GPUs of one box: one process per GPU, each owns a contiguous slice of
the frame axis, and the only data-path collective is ONE all-reduce of two scalars (sum of the
per-frame losses, number of frames) -- NCCL over NVLink on the GPUs, gloo in the CPU tests.

The reference is single-process (`devices: 1`, train_config.yaml:155); its `torch.mean` over all
frames (losses.py:211) becomes: local sum -> all-reduce -> / global frame count.  The backward
needs no collective: each rank scales its own rows by 1 / N_global.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import losses


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced split of `n_items` (signals, so a signal's frames stay together)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _GlobalMean(torch.autograd.Function):
    """mean over the frames of ALL ranks of per-rank rows; differentiable w.r.t. the local rows."""

    @staticmethod
    def forward(ctx, rows, group):
        total = rows.sum(dtype=torch.float64)
        count = float(rows.numel())
        ctx.shape, ctx.dtype = rows.shape, rows.dtype
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # (no host->device copy here: `torch.tensor(count, device=...)` would synchronise the stream)
            stats = torch.stack((total, torch.full_like(total, count)))
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
            ctx.save_for_backward(stats[1:2])
            ctx.count = None
            return (stats[0] / stats[1]).to(rows.dtype)
        ctx.count = count
        return (total / count).to(rows.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.count is None:
            (count,) = ctx.saved_tensors
            g = (grad_out.to(torch.float64) / count).to(ctx.dtype)
        else:
            g = grad_out / ctx.count
        return g.expand(ctx.shape), None


def global_mean(rows: torch.Tensor, group=None) -> torch.Tensor:
    return _GlobalMean.apply(rows, group)


class ShardedWasserstein1D(losses.Wasserstein1D):
    """`Wasserstein1D` whose inputs are this rank's slice of the batch and whose value is the mean
    over the whole (all-rank) batch.  Same constructor plus `process_group` (None = the default
    group).  `dims` other than None and `return_quantiles` stay local."""

    def __init__(self, *args, process_group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.process_group = process_group

    def _mean_group(self):
        return self.process_group

    def forward(self, x, y, x_pos=None, y_pos=None, **kwargs):
        if kwargs.get("dims", None) is not None or kwargs.get("return_quantiles", False):
            return super().forward(x, y, x_pos=x_pos, y_pos=y_pos, **kwargs)
        if not self.hinge and getattr(self, "backward_mode", "recompute") == "recompute" and x.numel() > 0:
            return super().forward(x, y, x_pos=x_pos, y_pos=y_pos, **kwargs)  # fused global mean
        if (x_pos is None or y_pos is None) and self.fixed_x is None:
            raise ValueError("If fixed_x is not provided, x_pos and y_pos must be provided")
        x_pos_ = self.fixed_x if x_pos is None else x_pos
        y_pos_ = self.fixed_x if y_pos is None else y_pos
        need_sort = (bool(self.require_sort) and x_pos is not None, bool(self.require_sort) and y_pos is not None)
        rows = losses.sot_frames(
            x, y, x_pos_.to(x.device), y_pos_.to(y.device), p=self.p, square=bool(self.square_dist),
            cut_scale=bool(kwargs.get("dont_normalize", False) or self.dont_normalize),
            limit=bool(kwargs.get("limit_quantile_range", False) or self.limit_quantile_range),
            require_sort=need_sort, backward_mode=getattr(self, "backward_mode", "recompute"))
        if self.hinge:
            rows = torch.nn.functional.relu(rows - kwargs.get("hinge", 0.0))
        return global_mean(rows, self.process_group)
