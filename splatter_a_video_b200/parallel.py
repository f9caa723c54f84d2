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

    def allreduce_grads(self, average: bool = True):
        """ONE dense all-reduce of the whole flat gradient buffer (see GradExchange for the volume-reduced exact variant)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        if average:
            self.flat_grad.div_(dist.get_world_size())
        return None


class GradExchange:
    """The step's gradient exchange over a FlatParams buffer: result == dense all-reduce (sum or mean) of the flat gradient.

    dense parameters are all-reduced; `subset` parameters only in their fixed slices; the `sparse` parameter (at most one:
    the spline coefficients) travels as each rank's <= 2 active slices through an all-gather and is added locally in rank order.

        subset = {name: (view_shape, dim, [slice indices])}          e.g. {"shs": ((P,16,3), 1, [0,2,6,12])}
        sparse = {name: (view_shape, dim, [idx_dev tensors (int32[1], <= 2)])}
        dirty  = int32[17] device list shared with gs.frame.deform_position_pair: exchanged intervals are recorded there so
                 the next backward clears them (CUDA path only).

    CUDA tensors: two kernels of this library (`spv_exchange_pack/unpack`) around ONE all-reduce + ONE all-gather.  CPU tensors
    (gloo tests): the same plan executed with torch indexing."""

    def __init__(self, flat: FlatParams, P: int, subset: Optional[dict] = None, sparse: Optional[dict] = None,
                 dirty: Optional[torch.Tensor] = None, skip: Sequence[str] = ()):
        self.flat, self.P, self.dirty = flat, int(P), dirty
        self.subset, self.sparse = dict(subset or {}), dict(sparse or {})
        if len(self.sparse) > 1:
            raise ValueError("GradExchange: at most one sparse parameter")
        self.segs = []          # (name, flat_off, A, B, C, mode, sel)
        off = 0
        for k, n in zip(flat.names, flat.sizes):
            w = n // self.P
            if k in skip:
                pass
            elif k in self.sparse or k in self.subset:
                shape, dim, sel = (self.sparse if k in self.sparse else self.subset)[k]
                A = 1
                for d in shape[1:dim]:
                    A *= int(d)
                B = int(shape[dim])
                C = w // (A * B)
                if k in self.sparse:
                    if not (1 <= len(sel) <= 2):
                        raise ValueError("GradExchange: 1 or 2 sparse slice indices")
                    self.segs.append((k, off, A, B, C, 2, list(sel)))
                else:
                    self.segs.append((k, off, A, B, C, 1, [int(x) for x in sel]))
            else:
                self.segs.append((k, off, 1, 1, w, 0, [0]))
            off += n
        self._cuda = None

    # ---- CUDA path -------------------------------------------------------------------------------------------------
    def _setup_cuda(self):
        import ctypes
        from . import _lib as L

        class Seg(ctypes.Structure):
            _fields_ = [("flat_offset", ctypes.c_longlong), ("A", ctypes.c_int), ("B", ctypes.c_int), ("C", ctypes.c_int),
                        ("mode", ctypes.c_int), ("nsel", ctypes.c_int), ("sel", ctypes.c_int * 16)]

        arr = (Seg * len(self.segs))()
        idx_ptrs = (ctypes.c_void_p * 16)()
        for q, (k, off, A, B, C, mode, sel) in enumerate(self.segs):
            arr[q].flat_offset, arr[q].A, arr[q].B, arr[q].C, arr[q].mode = off, A, B, C, mode
            arr[q].nsel = B if mode == 0 else len(sel)
            if mode == 1:
                for t, v in enumerate(sel):
                    arr[q].sel[t] = v
            if mode == 2:
                for t, v in enumerate(sel):
                    idx_ptrs[t] = v.data_ptr()
        n_ar, n_ag = ctypes.c_longlong(), ctypes.c_longlong()
        L.call("spv_exchange_sizes", self.P, len(self.segs), ctypes.cast(arr, ctypes.c_void_p), ctypes.byref(n_ar), ctypes.byref(n_ag))
        dev = self.flat.flat_grad.device
        world = dist.get_world_size() if dist.is_initialized() else 1
        self._cuda = dict(arr=arr, idx=idx_ptrs, n_ar=n_ar.value, n_ag=n_ag.value,
                          ar=torch.empty(max(n_ar.value, 1), dtype=torch.float32, device=dev),
                          ag=torch.empty(max(n_ag.value, 1), dtype=torch.float32, device=dev),
                          all=torch.empty(world, max(n_ag.value, 1), dtype=torch.float32, device=dev))

    def pack(self, scale: float = 1.0):
        import ctypes
        from . import _lib as L
        if self._cuda is None:
            self._setup_cuda()
        c = self._cuda
        L.call("spv_exchange_pack", self.P, len(self.segs), ctypes.cast(c["arr"], ctypes.c_void_p), ctypes.cast(c["idx"], ctypes.c_void_p),
               L.ptr(self.flat.flat_grad), float(scale), L.ptr(c["ar"]), L.ptr(c["ag"]), L.stream())
        return c["ar"], c["ag"]

    def unpack(self, reduced: torch.Tensor, gathered: torch.Tensor):
        import ctypes
        from . import _lib as L
        c = self._cuda
        L.call("spv_exchange_unpack", self.P, len(self.segs), ctypes.cast(c["arr"], ctypes.c_void_p), int(gathered.shape[0]),
               L.ptr(reduced), L.ptr(gathered), L.ptr(self.flat.flat_grad), L.ptr(self.dirty), L.stream())

    # ---- the exchange ----------------------------------------------------------------------------------------------
    def run(self, average: bool = True):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        scale = 1.0 / world if average else 1.0
        if self.flat.flat_grad.is_cuda:
            ar, ag = self.pack(scale)
            c = self._cuda
            if c["n_ar"]:
                dist.all_reduce(ar, op=dist.ReduceOp.SUM)
            if c["n_ag"]:
                dist.all_gather_into_tensor(c["all"], ag)
            self.unpack(ar, c["all"])
            return
        self._run_torch(world, dist.get_rank(), scale)

    def _views(self):
        g, P = self.flat.flat_grad, self.P
        for k, off, A, B, C, mode, sel in self.segs:
            yield k, g[off:off + P * A * B * C].view(P, A, B, C), mode, sel

    def _run_torch(self, world, rank, scale):
        """Same plan with torch ops (CPU / gloo): the parity reference of the kernels."""
        pieces, back = [], []
        for k, v, mode, sel in self._views():
            if mode == 2:
                idx = [int(t.item()) for t in sel]
                uniq = list(dict.fromkeys(idx))
                mine = torch.zeros(v.shape[0], v.shape[1], 2, v.shape[3], dtype=v.dtype)
                for t, b in enumerate(uniq):
                    mine[:, :, t] = v[:, :, b] * scale
                tail = torch.tensor(uniq + [-1] * (2 - len(uniq)), dtype=torch.float32)
                payload = torch.cat([mine.reshape(-1), tail])
                allp = [torch.empty_like(payload) for _ in range(world)]
                dist.all_gather(allp, payload)
                touched = sorted({int(b) for p_ in allp for b in p_[-2:].tolist() if b >= 0})
                acc = {b: torch.zeros_like(v[:, :, 0]) for b in touched}
                for p_ in allp:
                    data = p_[:-2].view(mine.shape)
                    for t, b in enumerate(p_[-2:].tolist()):
                        if b >= 0:
                            acc[int(b)] += data[:, :, t]
                for b in touched:
                    v[:, :, b] = acc[b]
            else:
                sl = v if mode == 0 else v[:, :, sel]
                pieces.append((sl * scale).reshape(-1))
                back.append((v, mode, sel, sl.shape))
        if pieces:
            comm = torch.cat(pieces)
            dist.all_reduce(comm, op=dist.ReduceOp.SUM)
            o = 0
            for v, mode, sel, shape in back:
                m = 1
                for d in shape:
                    m *= d
                blk = comm[o:o + m].view(shape)
                if mode == 0:
                    v.copy_(blk)
                else:
                    v[:, :, sel] = blk
                o += m


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
