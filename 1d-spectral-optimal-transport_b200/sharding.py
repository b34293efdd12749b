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

from . import _capi, losses


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced split of `n_items` (signals, so a signal's frames stay together)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class PeerReducer:
    """All-reduce (sum) of a handful of float64 scalars through NVLink / NVSwitch peer memory: every rank
    stores its values into a mailbox on each peer and spins on its own (`csrc/sot_p2p.cu`, one 32-thread
    kernel, a few microseconds) instead of an NCCL all-reduce (tens of microseconds of launch + protocol that
    sit between the forward and the backward launch of a 0.75 ms step).  The mailboxes are one symmetric
    allocation (`torch.distributed._symmetric_memory`); construction is collective over `group`."""

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.box = symm.empty(_capi.p2p_mailbox_doubles(self.world), dtype=torch.float64, device=device)
        self.box.zero_()
        self.handle = symm.rendezvous(self.box, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every mailbox is zeroed before anybody writes into it

    def all_reduce(self, values: torch.Tensor) -> torch.Tensor:
        self.seq += 1
        return _capi.p2p_allreduce(values, torch.empty_like(values), self.ptrs, self.rank, self.seq)

    def global_mean(self, local_sum: torch.Tensor, local_count: float):
        """(sum over ranks of local_sum) / (sum of local_count) and 1 / (sum of local_count), (1,) float32 each."""
        self.seq += 1
        return _capi.p2p_global_mean(local_sum, local_count, self.ptrs, self.rank, self.seq)


class _GlobalMean(torch.autograd.Function):
    """mean over the frames of ALL ranks of per-rank rows; differentiable w.r.t. the local rows."""

    @staticmethod
    def forward(ctx, rows, group):
        total = rows.sum(dtype=torch.float64)
        count = float(rows.numel())
        ctx.shape, ctx.dtype = rows.shape, rows.dtype
        if dist.is_available() and dist.is_initialized() and _world(group) > 1:
            # (no host->device copy here: `torch.tensor(count, device=...)` would synchronise the stream)
            stats = torch.stack((total, torch.full_like(total, count)))
            if isinstance(group, PeerReducer):
                stats = group.all_reduce(stats)
            else:
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


def _world(group) -> int:
    return group.world if isinstance(group, PeerReducer) else dist.get_world_size(group)


def global_mean(rows: torch.Tensor, group=None) -> torch.Tensor:
    return _GlobalMean.apply(rows, group)


class ShardedWasserstein1D(losses.Wasserstein1D):
    """`Wasserstein1D` whose inputs are this rank's slice of the batch and whose value is the mean
    over the whole (all-rank) batch.  Same constructor plus `process_group` (None = the default
    group).  `dims` other than None and `return_quantiles` stay local."""

    def __init__(self, *args, process_group=None, collective="auto", **kwargs):
        """`collective`: "nccl" = `torch.distributed.all_reduce` on `process_group`; "p2p" = the peer-memory
        kernel (`PeerReducer`); "auto" = p2p when it can be set up (CUDA, symmetric memory), else nccl."""
        super().__init__(*args, **kwargs)
        if collective not in ("auto", "nccl", "p2p"):
            raise ValueError('collective must be "auto", "nccl" or "p2p"')
        self.process_group = process_group
        self.collective = collective
        self._reducer = None

    def __getstate__(self):  # the peer-memory mailboxes belong to this process: never pickled with the module
        state = self.__dict__.copy()
        state["_reducer"] = None
        return state

    def _mean_group(self):
        if self.collective == "nccl" or not (dist.is_available() and dist.is_initialized()):
            return self.process_group
        if self._reducer is None and dist.get_world_size(self.process_group) > 1 and \
                dist.get_backend(self.process_group) == "nccl":
            try:  # (collective: every rank takes the same branch -- same code, same box)
                self._reducer = PeerReducer(self.process_group)
            except Exception:
                if self.collective == "p2p":
                    raise
                self._reducer = False
        return self._reducer if self._reducer else self.process_group

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
