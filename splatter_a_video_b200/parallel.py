"""Frame-parallel data parallelism: every rank holds the full Gaussian set and renders different
frames; after backward ONE all-reduce over a single flat fp32 gradient buffer, plus a 3-float/Gaussian densification
statistics reduce (sum |grad ndc|, sum visible, max radius).  The reference has no gradient collective at all
(`--distributed` only barriers, /root/reference/src/train.py:19-31,210-213), so this is new functionality whose
parity reference is the 1-GPU run accumulating the same (ids1, ids2) pairs.

Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_frames(num_frames: int, rank: int, world: int) -> range:
    """Contiguous block of the clip per rank (keeps a rank's ground-truth frames local)."""
    per = (num_frames + world - 1) // world
    lo = min(rank * per, num_frames)
    return range(lo, min(lo + per, num_frames))


def frame_for_step(step: int, rank: int, world: int, num_frames: int) -> int:
    """Frame rendered by `rank` at global step `step`: rank r walks its own contiguous shard."""
    shard = shard_frames(num_frames, rank, world)
    if len(shard) == 0:
        return (step * world + rank) % max(num_frames, 1)
    return shard[step % len(shard)]


class FlatParams:
    """Named per-Gaussian parameters stored as views of ONE flat fp32 buffer, gradients as views of a second one,
    so the per-step collective is a single all-reduce with no packing copies."""

    def __init__(self, tensors: Dict[str, torch.Tensor]):
        self.names = list(tensors)
        self.shapes = {k: tuple(v.shape) for k, v in tensors.items()}
        sizes = [tensors[k].numel() for k in self.names]
        total = sum(sizes)
        dev = next(iter(tensors.values())).device
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params: Dict[str, torch.Tensor] = {}
        off = 0
        for k, n in zip(self.names, sizes):
            view = self.flat[off:off + n].view(self.shapes[k])
            view.copy_(tensors[k])
            p = view.requires_grad_(True)
            p.grad = self.flat_grad[off:off + n].view(self.shapes[k])  # autograd accumulates in place into this view
            self.params[k] = p
            off += n

    def __getitem__(self, k: str) -> torch.Tensor:
        return self.params[k]

    def zero_grad(self):
        self.flat_grad.zero_()

    def floats_per_gaussian(self, P: int) -> float:
        return self.flat.numel() / max(P, 1)

    def allreduce_grads(self, average: bool = True, async_op: bool = False):
        """One collective per step over the whole flat gradient buffer."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        work = dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, async_op=async_op)
        if average and not async_op:
            self.flat_grad.div_(dist.get_world_size())
        return work


def reduce_densify_stats(grad_norm_sum: torch.Tensor, visible_count: torch.Tensor, max_radius: torch.Tensor):
    """Densification statistics must be identical on every rank so clone/split/prune stay replicated
    (atlas_gs_optimizer.py:110-121): sum, sum, max."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return grad_norm_sum, visible_count, max_radius
    packed = torch.stack([grad_norm_sum.float(), visible_count.float()], 0)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    r = max_radius.clone()
    dist.all_reduce(r, op=dist.ReduceOp.MAX)
    return packed[0], packed[1], r
