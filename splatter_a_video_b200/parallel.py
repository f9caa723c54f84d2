"""Frame-parallel data parallelism: every rank holds the full Gaussian set and renders different
frames; after backward ONE all-reduce over a single flat fp32 gradient buffer, plus a 3-float/Gaussian densification
statistics reduce (sum |grad ndc|, sum visible, max radius).  The reference has no gradient collective at all
(`--distributed` only barriers, /root/reference/src/train.py:19-31,210-213), so this is new functionality whose
parity reference is the 1-GPU run accumulating the same (ids1, ids2) pairs.

Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

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
        self.sizes = sizes
        total = sum(sizes)
        total = (total + 3) // 4 * 4      # float4-friendly tail for the fused optimizer
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

    def zero_grad(self, names: Sequence[str] = None):
        """Zero the whole flat gradient buffer, or only the named parameters' slices."""
        if names is None:
            self.flat_grad.zero_()
        else:
            for k in names:
                self.params[k].grad.zero_()

    def grad_sinks(self, names: Sequence[str]) -> Dict[str, torch.Tensor]:
        """The named parameters' slices of the flat gradient buffer, for operators that can write gradients in place
        (gs.frame.render_ortho_frame(grad_sinks=...)): no zero-fill and no accumulation pass for those parameters."""
        return {k: self.params[k].grad for k in names}

    def floats_per_gaussian(self, P: int) -> float:
        return self.flat.numel() / max(P, 1)

    def allreduce_grads(self, average: bool = True, sparse: Optional[dict] = None, subset: Optional[dict] = None):
        """The step's gradient exchange over the flat gradient buffer; the result always equals the dense all-reduce sum.

        Without hints: ONE all-reduce of the whole buffer.  Two exact volume reductions for this model:

        sparse = {name: (view_shape, dim, index_tensor)}: the gradient is non-zero only in ONE slice along `dim`, a
            different one per rank -- the cubic-spline coefficients: a frame touches the 12 coefficients of its own interval
            out of 4*NI*3 (dynamic_gaussian_with_base_point_cloud.py:239-247).  The active slices travel through an
            all-gather (N*P*12 floats instead of P*4*NI*3) and are scatter-added locally.
        subset = {name: (view_shape, dim, index_tensor)}: the gradient is non-zero only in the SAME slices on every rank --
            the SH coefficients under the renderer's constant view direction (0,0,1) (dptr_ortho_enhanced.py:270-271): only
            bases 0, 2, 6, 12 have a non-zero basis value, 12 of 48 floats.  Only those slices are all-reduced.
        Every other parameter is all-reduced in place, one collective per maximal contiguous run."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        world, rank = dist.get_world_size(), dist.get_rank()
        sparse, subset = sparse or {}, subset or {}
        if not sparse and not subset:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            if average:
                self.flat_grad.div_(world)
            return None
        scale = 1.0 / world if average else 1.0
        # ---- one packed all-reduce: [every dense parameter | the fixed slices of the `subset` parameters]
        pieces, writeback, off = [], [], 0
        gathers = []
        for k, n in zip(self.names, self.sizes):
            g = self.flat_grad[off:off + n]
            if k in sparse:
                shape, dim, index = sparse[k]
                gathers.append((g.view(shape), dim, index))
            elif k in subset:
                shape, dim, index = subset[k]
                gv = g.view(shape)
                idx = index.to(torch.long)
                pieces.append(gv.index_select(dim, idx).reshape(-1))
                writeback.append(("subset", gv, dim, idx))
            else:
                pieces.append(g)
                writeback.append(("dense", g, None, None))
            off += n
        tail = self.flat_grad[off:]                       # float4 padding of the flat buffer
        if pieces:
            comm = torch.cat(pieces)
            if scale != 1.0:
                comm.mul_(scale)
            dist.all_reduce(comm, op=dist.ReduceOp.SUM)
            o = 0
            for (kind, gv, dim, idx), src in zip(writeback, pieces):
                m = src.numel()
                if kind == "dense":
                    gv.copy_(comm[o:o + m])
                else:
                    sel_shape = list(gv.shape); sel_shape[dim] = idx.numel()
                    gv.index_copy_(dim, idx, comm[o:o + m].view(sel_shape))
                o += m
        # ---- one all-gather per sparse parameter: [own active slice | its index]
        for gv, dim, index in gathers:
            idx = index.to(torch.long)
            mine = gv.index_select(dim, idx)
            if scale != 1.0:
                mine = mine * scale
                gv.index_copy_(dim, idx, mine)
            payload = torch.cat([mine.reshape(-1), index.to(torch.float32).reshape(-1)])
            allp = torch.empty(world, payload.numel(), dtype=torch.float32, device=payload.device)
            if dist.get_backend() == "nccl":
                dist.all_gather_into_tensor(allp, payload)
            else:                                           # gloo (CPU tests) has no flat all-gather
                dist.all_gather(list(allp.unbind(0)), payload)
            nsl = mine.numel()
            for r in range(world):
                if r != rank:
                    gv.index_add_(dim, allp[r, nsl:].to(torch.long), allp[r, :nsl].view(mine.shape))
        if tail.numel():
            tail.zero_()
        return None


class FlatAdam:
    """torch.optim.Adam semantics (betas, eps, per-parameter learning rates) as ONE fused kernel over FlatParams
    (`spv_adam_step`).  Role of the per-attribute param groups of src/pointrix/optimizer/__init__.py:27-62."""

    def __init__(self, flat: FlatParams, lrs: Dict[str, float], betas=(0.9, 0.999), eps: float = 1e-15):
        import ctypes
        self.flat, self.betas, self.eps = flat, betas, eps
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        ends, off = [], 0
        for k, n in zip(flat.names, flat.sizes):
            off += n
            ends.append(off)
        ends[-1] = flat.flat.numel()
        self._ends = (ctypes.c_longlong * len(ends))(*ends)
        self._lrs = (ctypes.c_float * len(ends))(*[float(lrs[k]) for k in flat.names])
        self.nseg = len(ends)
        self.t = 0

    def step(self):
        from . import _lib as L
        import ctypes
        self.t += 1
        f = self.flat
        L.call("spv_adam_step", f.flat.numel(), L.ptr(f.flat), L.ptr(f.flat_grad), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
               self.nseg, ctypes.cast(self._ends, ctypes.c_void_p), ctypes.cast(self._lrs, ctypes.c_void_p), float(self.betas[0]),
               float(self.betas[1]), float(self.eps), int(self.t), L.stream())


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
