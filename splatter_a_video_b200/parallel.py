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


def frame_for_step(step: int, rank: int, world: int, num_frames: int, policy: str = "interleaved") -> int:
    """Frame rendered by `rank` at global step `step`.

    "interleaved" (default): frame (step*world + rank) mod F -- the index stream a non-shuffling DistributedSampler deals out
    (indices[rank::world]; the reference swaps one in under --distributed, src/loaders/create_training_dataset.py:168,186).  The
    `world` frames of one step are consecutive, so they fall into one or two spline intervals (a node every 5 frames) and
    have similar tile lists: the exchanged interval gradients stay as compact as on one GPU and the ranks finish together.
    "blocks": rank r walks its own contiguous shard of the clip (keeps a rank's ground-truth frames local)."""
    if policy == "interleaved":
        return (step * world + rank) % max(num_frames, 1)
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

    @classmethod
    def empty_like(cls, other: "FlatParams", P_old: int, P_new: int) -> "FlatParams":
        """Same named parameters for a population of P_new points (uninitialised values, zero gradients): used when the
        structure changes (densification) so no per-parameter temporaries are needed."""
        self = cls.__new__(cls)
        self.names = list(other.names)
        self.shapes = {k: (P_new,) + tuple(other.shapes[k][1:]) for k in self.names}
        self.sizes = [n // max(P_old, 1) * P_new for n in other.sizes]
        total = (sum(self.sizes) + 3) // 4 * 4
        dev = other.flat.device
        self.flat = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
        self.params = {}
        off = 0
        for k, n in zip(self.names, self.sizes):
            p = self.flat[off:off + n].view(self.shapes[k]).requires_grad_(True)
            p.grad = self.flat_grad[off:off + n].view(self.shapes[k])
            self.params[k] = p
            off += n
        return self

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
                 dirty: Optional[torch.Tensor] = None, skip: Sequence[str] = (), deferred: Optional[dict] = None):
        self.flat, self.P, self.dirty = flat, int(P), dirty
        self.exchange_path = "nccl"
        # deferred = {"shs": name of the SH parameter, "node": name of the spline parameter, "NI": interval count}: the two
        # LINEAR tails of the backward pass (colour -> SH coefficients, position -> spline coefficients) run AFTER the
        # exchange on the reduced / gathered upstream gradients -- 3 + 6 floats per Gaussian on the wire instead of 12 + 24.
        self.deferred = dict(deferred) if deferred else None
        if self.deferred:
            skip = tuple(skip) + (self.deferred["shs"], self.deferred["node"])
            subset, sparse = None, None
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
                          ar=torch.zeros(max(n_ar.value, 1), dtype=torch.float32, device=dev),
                          ag=torch.empty(max(n_ag.value, 1), dtype=torch.float32, device=dev),
                          all=torch.empty(world, max(n_ag.value, 1), dtype=torch.float32, device=dev))
        if self.deferred:
            # ONE staging row per rank: [packed dense parameters | colour gradient 3P | position gradients 6P | 4 frame scalars]
            c, P = self._cuda, self.P
            nd = (n_ar.value + 3) // 4 * 4                       # keep every block 16-byte aligned
            n_red = nd + (3 * P + 3) // 4 * 4                    # floats that are SUMMED over ranks (dense + colour)
            row = n_red + (6 * P + 4 + 3) // 4 * 4
            c["row"] = torch.zeros(row, dtype=torch.float32, device=dev)
            c["rows"] = torch.zeros(world, row, dtype=torch.float32, device=dev)
            c["reduced"] = torch.zeros(n_red, dtype=torch.float32, device=dev)
            c["n_red"], c["n_dense"] = n_red, nd
            c["ar"] = c["row"][:max(n_ar.value, 1)]              # pack writes the dense parameters straight into the row
            c["g_rgb"] = c["row"][nd:nd + 3 * P].view(P, 3)
            c["payload"] = c["row"][n_red:n_red + 6 * P + 4]
            c["gathered"] = c["rows"][:, n_red:]                 # per-rank [6P + 4 (+pad)] views, stride = row
            c["clamped"] = torch.zeros(P, 3, dtype=torch.uint8, device=dev)
            c["dirs"] = torch.zeros(P, 3, dtype=torch.float32, device=dev); c["dirs"][:, 2] = 1     # the renderer's constant view direction

    # ---- deferred linear tails (CUDA only) ------------------------------------------------------------------------
    def sh_sink(self):
        """(dL_drgb[P,3], clamped[P,3]) for render_ortho_frame(grad_sinks={"shs_deferred": ...}): the frame backward leaves the
        colour gradient in the all-reduce buffer instead of running the SH backward."""
        if self._cuda is None:
            self._setup_cuda()
        return self._cuda["g_rgb"], self._cuda["clamped"]

    def node_defer(self):
        """Staging buffer for gs.frame.deform_position_pair(defer=...)."""
        if self._cuda is None:
            self._setup_cuda()
        return self._cuda["payload"]

    def _p2p_setup(self) -> bool:
        """Symmetric (peer-mapped) double buffer for the staging rows via torch.distributed._symmetric_memory.  Disabled with
        SPV_EXCHANGE=nccl; falls back to the NCCL all-gather if the rendezvous is not possible on this system."""
        c = self._cuda
        if "p2p" in c:
            return c["p2p"] is not None
        c["p2p"] = None
        import os
        if os.environ.get("SPV_EXCHANGE", "p2p") != "p2p":
            self.exchange_path = "nccl all-gather"
            return False
        try:
            import ctypes
            import torch.distributed._symmetric_memory as symm
            row, n_red = c["row"].numel(), int(c["n_red"])
            world_n = dist.get_world_size()
            # per half: [staging row | the reduced block | NVLS push mode: every rank's gathered tail, slot r written by rank r]
            span = row + n_red + world_n * (row - n_red)
            buf = symm.empty(2 * span, dtype=torch.float32, device=c["row"].device)
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) > 8:
                raise ValueError("the peer-pointer tables of the exchange kernels hold 8 ranks (one NVSwitch domain)")
            arrs, reds = [], []
            for half in range(2):
                a, b = (ctypes.c_void_p * 8)(), (ctypes.c_void_p * 8)()
                for r in range(len(ptrs)):
                    a[r] = ptrs[r] + half * span * 4
                    b[r] = ptrs[r] + (half * span + row) * 4
                arrs.append(a); reds.append(b)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            want = os.environ.get("SPV_EXCHANGE_MODE", "auto")       # auto | oneshot | twophase | nvls
            if want == "auto":
                want = "nvls" if (mc and dist.get_world_size() >= 4) else "oneshot"
            if want == "nvls" and not mc:
                want = "oneshot"
            push = want == "nvls" and os.environ.get("SPV_EXCHANGE_PULL", "0") != "1"
            c["p2p"] = dict(buf=buf, hdl=hdl, arrs=arrs, reds=reds, span=span, step=0, mode=want, mc=mc, push=push)
            self.exchange_path = {"oneshot": "p2p one-shot (peer loads)", "twophase": "p2p two-phase (peer loads)",
                                  "nvls": "nvls (multimem.ld_reduce / multimem.st through the NVSwitch; tails "
                                          + ("pushed by multicast stores)" if push else "pulled by peer loads)")}[want]
        except Exception as e:   # noqa: BLE001 -- any failure here means: no peer mapping on this system
            self.exchange_path = f"nccl ({type(e).__name__}: {str(e)[:80]})"
            import warnings
            warnings.warn(f"GradExchange: symmetric-memory rendezvous failed, falling back to the NCCL all-gather: {e!r}", RuntimeWarning)
        return c["p2p"] is not None

    def _p2p_exchange(self, world: int, scale: float) -> torch.Tensor:
        """Publishes the staging row in the symmetric buffer and returns the tensor holding the summed (scaled) block.
        The gathered tails and the spline backward that consumes them run on a second stream next to the reduction
        (joined by `_finish_deferred`)."""
        import ctypes
        from . import _lib as L
        c = self._cuda
        p = c["p2p"]
        half, row, span, n_red = p["step"] & 1, c["row"].numel(), p["span"], int(c["n_red"])
        p["step"] += 1
        main = torch.cuda.current_stream()
        if "side" not in p:
            p["side"] = torch.cuda.Stream()
        side = p["side"]
        rows_ptr = ctypes.cast(p["arrs"][half], ctypes.c_void_p)
        if p["push"]:
            # publish by push: the summed block into this rank's half, the gathered tail multicast into slot `rank` of every rank's
            # gathered area; after the barrier the spline backward reads all tails LOCALLY, next to the in-switch reduction
            gath, rank = row - n_red, dist.get_rank()
            g0 = half * span + row + n_red                       # float offset of the gathered area inside the symmetric buffer
            L.call("spv_exchange_publish", n_red, row, L.ptr(c["row"]), p["buf"].data_ptr() + half * span * 4,
                   p["mc"] + (g0 + rank * gath) * 4, L.stream())
            p["hdl"].barrier(channel=0)
            gathered = p["buf"][g0:g0 + world * gath].view(world, gath)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._spline_tail(world, scale, gathered)
        else:
            p["buf"][half * span:half * span + row].copy_(c["row"])
            p["hdl"].barrier(channel=0)
            # ---- side stream: gather every rank's position gradients + frame scalars, rebuild the spline-coefficient gradient
            side.wait_stream(main)
            with torch.cuda.stream(side):
                L.call("spv_exchange_gather_peers", n_red, row, world, rows_ptr, L.ptr(c["rows"]), c["rows"].stride(0), L.stream())
                self._spline_tail(world, scale)
        p["side_pending"] = True
        # ---- main stream: the summed block
        if p["mode"] == "oneshot":
            L.call("spv_exchange_reduce_peers", n_red, row, world, rows_ptr, float(scale), L.ptr(c["reduced"]), None, c["rows"].stride(0),
                   L.stream())
            return c["reduced"]
        rank = dist.get_rank()
        red_pub = p["buf"][half * span + row:(half + 1) * span]
        if p["mode"] == "nvls":
            mc_row = p["mc"] + half * span * 4
            L.call("spv_exchange_nvls", n_red, row, rank, world, mc_row, mc_row + row * 4, rows_ptr, float(scale), None,
                   c["rows"].stride(0), L.stream())
            p["hdl"].barrier(channel=1)
            return red_pub                      # every rank's own red area now holds the whole block (written by its owners)
        L.call("spv_exchange_reduce_scatter_peers", n_red, row, rank, world, rows_ptr, float(scale), L.ptr(red_pub), L.ptr(c["reduced"]),
               None, c["rows"].stride(0), L.stream())
        p["hdl"].barrier(channel=1)
        L.call("spv_exchange_fetch_reduced", n_red, rank, world, ctypes.cast(p["reds"][half], ctypes.c_void_p), L.ptr(c["reduced"]),
               L.stream())
        return c["reduced"]

    def _exchange_rows(self, world: int, scale: float) -> torch.Tensor:
        c = self._cuda
        if self._p2p_setup():
            # symmetric-memory path: publish the row, one device-side barrier, then the reduction kernel pulls the peers' rows
            # over NVLink (double-buffered: the barrier of step k+1 also proves every rank finished reading step k's buffer)
            return self._p2p_exchange(world, scale)
        # ONE NCCL all-gather of the staging rows; dense parameters and colour gradient are summed locally in rank order
        dist.all_gather_into_tensor(c["rows"], c["row"])
        self._reduce_rows(world, scale)
        return c["reduced"]

    def _reduce_rows(self, world: int, scale: float):
        from . import _lib as L
        c = self._cuda
        L.call("spv_exchange_reduce", int(c["n_red"]), world, L.ptr(c["rows"]), c["rows"].stride(0), float(scale), L.ptr(c["reduced"]),
               L.stream())

    def _spline_tail(self, world: int, scale: float, gathered: Optional[torch.Tensor] = None):
        from . import _lib as L
        c, d = self._cuda, self.deferred
        node = self.flat.params[d["node"]]
        gathered = c["gathered"] if gathered is None else gathered          # [world, 6P + 4 (+pad)] rows
        L.call("spv_deform_spline_backward_gathered", self.P, int(d["NI"]), int(bool(d.get("interval_major", False))), world,
               gathered.data_ptr(), gathered.stride(0),
               float(scale), L.ptr(self.dirty), L.ptr(node.grad), L.stream())

    def _finish_deferred(self, world: int, scale: float, reduced: torch.Tensor):
        from . import _lib as L
        c, P, d = self._cuda, self.P, self.deferred
        g_rgb_red = reduced[c["n_dense"]:c["n_dense"] + 3 * P]
        shs = self.flat.params[d["shs"]]
        if shs.shape[1] == 4:          # only the bases the constant view direction reaches are kept (gs.frame.sh_z_split)
            L.call("spv_compute_sh_z_backward", P, L.ptr(c["clamped"]), L.ptr(g_rgb_red), L.ptr(shs.grad), L.stream())
        else:
            L.call("spv_compute_sh_backward", P, L.ptr(shs), 3, L.ptr(c["dirs"]), None, L.ptr(c["clamped"]), L.ptr(g_rgb_red), 16,
                   L.ptr(shs.grad), None, L.stream())
        p = c.get("p2p")
        if p and p.get("side_pending"):          # the spline tail already runs on the side stream: join it
            torch.cuda.current_stream().wait_stream(p["side"])
            p["side_pending"] = False
        else:
            self._spline_tail(world, scale)

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
    @property
    def is_collective(self) -> bool:
        """False when run() is a no-op (one process, nothing deferred): the caller may then capture the whole step in one graph."""
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        return bool(multi or self.deferred)

    def run(self, average: bool = True):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            if self.deferred:                      # single process: no collective, but the deferred tails still have to run
                if self._cuda is None:
                    self._setup_cuda()
                self.pack(1.0)
                self._cuda["rows"][0].copy_(self._cuda["row"])
                self._reduce_rows(1, 1.0)
                self.unpack(self._cuda["reduced"], self._cuda["all"])
                self._finish_deferred(1, 1.0, self._cuda["reduced"])
            return
        world = dist.get_world_size()
        scale = 1.0 / world if average else 1.0
        if self.flat.flat_grad.is_cuda:
            ar, ag = self.pack(1.0 if self.deferred else scale)      # deferred: the scale is applied by the local reduction
            c = self._cuda
            if self.deferred:
                reduced = self._exchange_rows(world, scale)
                # two independent consumers of the reduced block: the dense parameters' copy into the flat gradient (aux stream)
                # next to the deferred SH backward (this stream); the spline tail already runs on the exchange's side stream
                main = torch.cuda.current_stream()
                if "aux" not in c:
                    c["aux"] = torch.cuda.Stream()
                c["aux"].wait_stream(main)
                with torch.cuda.stream(c["aux"]):
                    self.unpack(reduced, c["all"])
                self._finish_deferred(world, scale, reduced)
                main.wait_stream(c["aux"])
                return
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

    def __init__(self, flat: FlatParams, lrs: Dict[str, float], betas=(0.9, 0.999), eps: float = 1e-15, device_clock: bool = False,
                 lazy: Optional[dict] = None):
        """device_clock: keep the step counter / bias corrections and the learning rates in device memory so `step()` can be
        captured in a CUDA graph and replayed (every replay advances the clock; `set_lrs` updates the rates outside the graph).

        lazy = {"name": spline-coefficient parameter, "P": points, "NI": intervals, "interval_major": bool, "dirty": int32[17] device
        list shared with gs.frame.deform_position_pair / GradExchange}: interval-lazy update of that parameter (implies
        device_clock).  A step only streams the intervals that hold gradient; call `prepare(idx1_dev, idx2_dev)` before a forward
        pass reads intervals and `flush()` before anything else observes the parameter (densification, checkpoints, rendering
        other frames).  Same parameters as the dense update whenever they are observed that way (tests/test_frame_gpu.py)."""
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
        self.lazy = dict(lazy) if lazy else None
        self.device_clock = bool(device_clock) or self.lazy is not None
        if self.device_clock:
            dev = flat.flat.device
            self.state_dev = torch.zeros(8, dtype=torch.float32, device=dev)      # [step, bc1, bc2_sqrt, -, b1^t, b2^t as 2 doubles]
            self.lr_dev = torch.tensor([float(lrs[k]) for k in flat.names], dtype=torch.float32, device=dev)
        if self.lazy:
            dev = flat.flat.device
            self.lazy["seg"] = flat.names.index(self.lazy["name"])
            self.lazy["layout"] = int(bool(self.lazy.get("interval_major", False)))
            self.last_dev = torch.zeros(int(self.lazy["NI"]), dtype=torch.int32, device=dev)
            self.ring_dev = torch.zeros(2 * 4096, dtype=torch.float32, device=dev)

    def set_lrs(self, lrs: Dict[str, float]):
        """Learning-rate schedule hook (role of ExponLRScheduler, src/pointrix/optimizer/scheduler.py:9-100)."""
        vals = [float(lrs[k]) for k in self.flat.names]
        for i, v in enumerate(vals):
            self._lrs[i] = v
        if self.device_clock:
            self.lr_dev.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=True)

    def _lazy_args(self):
        import ctypes
        from . import _lib as L
        z, f = self.lazy, self.flat
        return (f.flat.numel(), self.nseg, ctypes.cast(self._ends, ctypes.c_void_p), z["seg"], int(z["P"]), int(z["NI"]), z["layout"],
                L.ptr(f.flat), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq))

    def prepare(self, idx1_dev: torch.Tensor, idx2_dev: torch.Tensor):
        """Lazy mode: bring the intervals a forward pass is about to read (device scalars, as given to deform_position_pair) up to
        the current optimizer step.  No-op for the dense optimizer."""
        if not self.lazy:
            return
        from . import _lib as L
        L.call("spv_adam_lazy_prepare", *self._lazy_args(), L.ptr(idx1_dev), L.ptr(idx2_dev), L.ptr(self.last_dev), L.ptr(self.ring_dev),
               L.ptr(self.state_dev), L.ptr(self.lr_dev), float(self.betas[0]), float(self.betas[1]), float(self.eps), L.stream())

    def flush(self):
        """Lazy mode: bring EVERY interval up to the current step (before densification, checkpoints, rendering other frames)."""
        if not self.lazy:
            return
        from . import _lib as L
        L.call("spv_adam_lazy_flush", *self._lazy_args(), L.ptr(self.last_dev), L.ptr(self.ring_dev), L.ptr(self.state_dev),
               L.ptr(self.lr_dev), float(self.betas[0]), float(self.betas[1]), float(self.eps), L.stream())

    def step(self):
        from . import _lib as L
        import ctypes
        self.t += 1
        f = self.flat
        if self.lazy:
            z = self.lazy
            L.call("spv_adam_step_lazy", f.flat.numel(), L.ptr(f.flat), L.ptr(f.flat_grad), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                   self.nseg, ctypes.cast(self._ends, ctypes.c_void_p), L.ptr(self.lr_dev), float(self.betas[0]), float(self.betas[1]),
                   float(self.eps), L.ptr(self.state_dev), z["seg"], int(z["P"]), int(z["NI"]), z["layout"], L.ptr(z["dirty"]),
                   L.ptr(self.last_dev), L.ptr(self.ring_dev), L.stream())
            return
        if self.device_clock:
            L.call("spv_adam_step_device", f.flat.numel(), L.ptr(f.flat), L.ptr(f.flat_grad), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                   self.nseg, ctypes.cast(self._ends, ctypes.c_void_p), L.ptr(self.lr_dev), float(self.betas[0]), float(self.betas[1]),
                   float(self.eps), L.ptr(self.state_dev), L.stream())
            return
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
