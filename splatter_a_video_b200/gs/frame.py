"""Fused per-frame operators (extensions beyond the ``dptr.gs`` surface; used by DPTROrthoEnhancedRenderB200).

``render_ortho_frame``: the whole of ``DPTROrthoEnhancedRender.render_iter``
(/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py:205-383) as ONE autograd node over the C ABI
(`spv_frame_ortho_forward/backward`): no host synchronisation, exact tile culling, single-traversal blending.  Because
nothing on the path reads device memory from the host, a whole training step can be captured in a CUDA graph.

``deform_position``: cubic-spline position of the active model (src/dynamic_gaussian_with_base_point_cloud.py:236-250).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L


class CapacityOverflow(RuntimeError):
    """A frame that was already handed to the caller turned out to have been rendered from a truncated tile list."""


class Capacity:
    """Intersection-capacity policy of the sync-free path.  The count of tile intersections never leaves the device on the
    hot path: every frame's (count, overflow) status is copied to its own pinned host slot asynchronously, and EVERY status is
    eventually inspected -- the previous frame's before the next one is queued (by then its forward has long finished, so the
    wait is free), the current one on demand (`check(wait=True)`).  Overflow is sticky until reported:

      * found for the frame being rendered (`wait=True`): `check()` grows the capacity and returns False, the renderer renders
        the frame again;
      * found LATE (an earlier frame whose images / gradients were already used): the capacity is grown, `late_overflows` is
        counted and `check()` raises CapacityOverflow once, so the training step that consumed the truncated frame is known.

    Head-room is kept proactively (grow at `refill` of the capacity in use, rescale when the Gaussian count changes), so with
    a slowly growing scene the late case needs a > 1/refill jump of the intersection count between two frames."""

    def __init__(self, initial: int = 0, growth: float = 1.5, slack: float = 1.25, refill: float = 0.85):
        self.I_cap = int(initial)
        self.growth, self.slack, self.refill = growth, slack, refill
        self._pending = []          # [(event, pinned host int32[2], capacity the frame was rendered with)]
        self._free = []
        self.last_I = 0
        self.P = None
        self.late_overflows = 0
        self._late = False

    def _slot(self):
        if self._free:
            return self._free.pop()
        t = torch.zeros(2, dtype=torch.int32)
        return t.pin_memory() if torch.cuda.is_available() else t

    def set_population(self, P: int):
        """Rescale the capacity when the number of Gaussians changes (densification / pruning)."""
        if self.P is not None and P != self.P and self.P > 0 and self.I_cap > 0:
            self.I_cap = max(self.I_cap, int(self.I_cap * (P / self.P) * 1.05) + 1024)
        self.P = P

    def _fold(self, wait: bool, current: bool):
        """Read every finished status (all of them when `wait`).  `current`: the newest pending status belongs to the frame
        being rendered right now (it can still be rendered again); every other overflow is a late one.  Returns True if the
        current frame overflowed."""
        current_overflowed = False
        while self._pending:
            ev, host, cap = self._pending[0]
            if not wait and not ev.query():
                break
            ev.synchronize()
            self._pending.pop(0)
            self.last_I = int(host[0])
            over = int(host[1]) != 0
            self._free.append(host)
            if over:
                if cap >= self.I_cap:                      # not already grown past the capacity that overflowed
                    self.I_cap = int(self.I_cap * self.growth) + 1024
                if current and wait and not self._pending:
                    current_overflowed = True
                else:
                    self._late = True
                    self.late_overflows += 1
            elif self.last_I > self.refill * cap and cap >= self.I_cap:
                self.I_cap = int(self.I_cap * self.growth) + 1024      # proactive head-room
        return current_overflowed

    def observe(self, status: Tensor):
        """Queue the status of the frame just enqueued; first settles every EARLIER frame (their kernels ran long ago)."""
        self._fold(wait=True, current=False)
        host = self._slot()
        host.copy_(status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((ev, host, self.I_cap))

    def check(self, wait: bool = False) -> bool:
        """True if the frame observed last can be used (wait=False: as far as is known without blocking).  False: it
        overflowed and the capacity has grown -- render it again.  Raises CapacityOverflow once for an overflow found late."""
        newest = self._fold(wait=wait, current=True)
        if self._late:
            self._late = False
            raise CapacityOverflow(f"an earlier frame overflowed the tile-intersection capacity (now grown to {self.I_cap}); "
                                   "its images and gradients were computed from truncated tile lists")
        return not newest

    def drain(self) -> bool:
        """Block until every observed frame has been inspected (end of an epoch / before a checkpoint); same result as check."""
        return self.check(wait=True)


def _ptr_array(ptrs):
    import ctypes
    return (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)


class _FrameOrtho(torch.autograd.Function):
    """inputs: position, scaling, rotation, opacity, shs, extr, <scalars>, ndc, abs_ndc, sinks, *attribute groups
    outputs: rgb[3,H,W], depth[1,H,W], one image per attribute group, gs_idx, radii, status."""

    @staticmethod
    def forward(ctx, position, scaling, rotation, opacity, shs, extr, W, H, K, bg_rgb, nearest, extent, I_cap, cull,
                ndc, abs_ndc, sinks, *attr_groups):
        import ctypes
        L.need_cuda(position, scaling, rotation, opacity, shs, extr, *attr_groups)
        pos, sc, rot, op, sh = (L.f32c(x) for x in (position, scaling, rotation, opacity, shs))
        # internal channel order: attribute groups that need a gradient first -- the backward kernel then reduces feature
        # gradients only for the leading 4 + (their channels) image channels
        order = sorted(range(len(attr_groups)), key=lambda i: not attr_groups[i].requires_grad)
        groups = [L.f32c(attr_groups[i]) for i in order]
        ex = L.f32c(extr)
        P = pos.shape[0]
        chans = [int(g.shape[1]) for g in groups]
        A = sum(chans)
        if sh.dim() != 3 or sh.shape[1] not in (16, 4) or sh.shape[2] != 3:
            raise ValueError("render_ortho_frame needs degree-3 SH coefficients [P,16,3], or [P,4,3] = the bases 0, 2, 6, 12 alone")
        nb = int(sh.shape[1])
        if len(groups) > 8 or A > 19:
            raise ValueError("render_ortho_frame handles at most 8 attribute groups / 19 attribute channels")
        dev = pos.device
        images = torch.empty(4 + A, H, W, dtype=torch.float32, device=dev)
        gs_idx = torch.empty(H, W, K, dtype=torch.int32, device=dev)
        radii = torch.empty(P, dtype=torch.int32, device=dev)
        status = torch.empty(2, dtype=torch.int32, device=dev)
        nbytes = L.query("spv_frame_workspace_bytes", P, int(I_cap), int(W), int(H), A)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ch_arr = (ctypes.c_int * max(len(chans), 1))(*chans)
        L.call("spv_frame_ortho_forward", P, int(W), int(H), len(groups), ctypes.cast(_ptr_array([g.data_ptr() for g in groups]), ctypes.c_void_p),
               ctypes.cast(ch_arr, ctypes.c_void_p), int(K), int(I_cap), int(bool(cull)), L.ptr(pos), L.ptr(sc), L.ptr(rot), L.ptr(op),
               L.ptr(sh), nb, L.ptr(ex), float(nearest), float(extent), float(bg_rgb), L.ptr(images), L.ptr(gs_idx), L.ptr(radii),
               L.ptr(status), L.ptr(ws), nbytes, L.stream())
        ctx.meta = (P, int(W), int(H), chans, int(I_cap), float(bg_rgb), ndc is not None, abs_ndc is not None, nb)
        ctx.sinks = dict(sinks) if sinks else {}
        ctx.order = order
        ctx.attr_needs = [bool(attr_groups[i].requires_grad) for i in order]
        ctx.save_for_backward(sc, rot, op, sh, ex, ws)
        ctx.mark_non_differentiable(gs_idx, radii, status)
        # outputs without an upstream gradient arrive as None in backward (handled there) instead of as zero-filled tensors:
        # autograd would otherwise materialise one for EVERY output, the [H,W,K] id image included (33 MB of fill per step)
        ctx.set_materialize_grads(False)
        # the per-image views are created inside forward: autograd hands their gradients to backward one by one, so no
        # zero-filled [C,H,W] gradient image is ever assembled
        views, c = [None] * len(order), 4
        for slot, n in zip(order, chans):
            views[slot] = images[c:c + n]; c += n
        return (images[:3], images[3:4], *views, gs_idx, radii, status)

    @staticmethod
    def backward(ctx, *grads):
        import ctypes
        P, W, H, chans, I_cap, bg_rgb, has_ndc, has_abs, nb = ctx.meta
        sc, rot, op, sh, ex, ws = ctx.saved_tensors
        first = not getattr(ctx, "backward_ran", False)     # a second backward over the same graph must clear the packed rows itself
        ctx.backward_ran = True
        dev = sc.device
        sinks = ctx.sinks
        ng = len(chans)
        user_grads = [None if g is None else L.f32c(g) for g in grads[:2 + ng]]
        img_grads = user_grads[:2] + [user_grads[2 + slot] for slot in ctx.order]      # internal channel order
        planes, HW = [], H * W * 4
        for g, n in zip(img_grads, [3, 1] + chans):
            for k in range(n):
                planes.append(None if g is None else g.data_ptr() + k * HW)
        n_grad_channels = 4 + sum(n for n, need in zip(chans, ctx.attr_needs) if need)

        def out(name, *shape):
            """Gradient buffer: a caller-provided sink (written in place, `None` returned to autograd so nothing is
            accumulated on top) or a fresh tensor handed to autograd."""
            t = sinks.get(name)
            if t is not None:
                assert t.is_contiguous() and t.numel() == math.prod(shape) and t.dtype == torch.float32
                return t, None
            t = torch.empty(*shape, dtype=torch.float32, device=dev)
            return t, t

        g_pos = torch.empty(P, 3, dtype=torch.float32, device=dev)
        g_sc, r_sc = out("scaling", P, 3)
        g_rot, r_rot = out("rotation", P, 4)
        g_op, r_op = out("opacity", P, 1)
        defer = sinks.get("shs_deferred")          # (dL_drgb[P,3], clamped[P,3] uint8): SH backward runs after the exchange
        if defer is not None:
            g_sh, r_sh = None, None
        else:
            g_sh, r_sh = out("shs", P, nb, 3)
        # attribute groups: sink key ("attr", i) = the i-th group the caller passed
        g_attr, r_attr = [], []
        for slot, n, need in zip(ctx.order, chans, ctx.attr_needs):
            t, r = out(("attr", slot), P, n) if need else (None, None)
            g_attr.append(t); r_attr.append(r)
        g_ndc = torch.empty(P, 2, dtype=torch.float32, device=dev) if has_ndc else None
        g_abs = torch.empty(P, 2, dtype=torch.float32, device=dev) if has_abs else None
        ch_arr = (ctypes.c_int * max(ng, 1))(*chans)
        L.call("spv_frame_ortho_backward", P, W, H, ng, ctypes.cast(ch_arr, ctypes.c_void_p), n_grad_channels, I_cap, L.ptr(sc), L.ptr(rot), L.ptr(op),
               L.ptr(sh), nb, L.ptr(ex), bg_rgb, ctypes.cast(_ptr_array(planes), ctypes.c_void_p), L.ptr(g_pos), L.ptr(g_sc), L.ptr(g_rot),
               L.ptr(g_op), L.ptr(g_sh), ctypes.cast(_ptr_array([None if t is None else t.data_ptr() for t in g_attr]), ctypes.c_void_p),
               L.ptr(g_ndc), L.ptr(g_abs), L.ptr(defer[0]) if defer is not None else None,
               L.ptr(defer[1]) if defer is not None else None, int(first), L.ptr(ws), ws.numel(), L.stream())
        g_user = [None] * ng
        for slot, t in zip(ctx.order, r_attr):
            g_user[slot] = t
        return (g_pos, r_sc, r_rot, r_op, r_sh, None, None, None, None, None, None, None, None, None, g_ndc, g_abs, None, *g_user)


# The bases of a degree-3 SH expansion that are non-zero along the ortho renderers' constant view direction (0,0,1)
# (dptr_ortho_enhanced.py:270-271).  A trainer that only ever renders through those renderers can keep just these coefficients
# trainable ([P,4,3], `sh_z_split`): the other twelve get gradient 0 on every step, so torch.optim.Adam never moves them.
SH_Z_BASES = (0, 2, 6, 12)


def sh_z_split(shs: Tensor) -> Tuple[Tensor, Tensor]:
    """[P,16,3] -> ([P,4,3] coefficients of SH_Z_BASES, [P,12,3] the others in ascending basis order)."""
    act = list(SH_Z_BASES)
    rest = [b for b in range(16) if b not in act]
    return shs[:, act].contiguous(), shs[:, rest].contiguous()


def sh_z_merge(shs_z: Tensor, shs_rest: Tensor) -> Tensor:
    """Inverse of sh_z_split: the [P,16,3] tensor the reference's checkpoints / PLY files hold."""
    act = list(SH_Z_BASES)
    rest = [b for b in range(16) if b not in act]
    out = torch.empty(shs_z.shape[0], 16, 3, dtype=shs_z.dtype, device=shs_z.device)
    out[:, act] = shs_z
    out[:, rest] = shs_rest.to(out.device)
    return out


def render_ortho_frame(position: Tensor, scaling: Tensor, rotation: Tensor, opacity: Tensor, shs: Tensor,
                       attrs, extr: Tensor, W: int, H: int, K: int, bg_rgb: float, I_cap: int,
                       cull: bool = True, nearest: float = 0.01, extent: float = 1.3, ndc: Optional[Tensor] = None,
                       abs_ndc: Optional[Tensor] = None, grad_sinks: Optional[dict] = None):
    """One frame of DPTROrthoEnhancedRender.render_iter as one autograd node.

    shs: [P,16,3], or [P,4,3] holding the coefficients of SH_Z_BASES alone (bit-identical images and gradients; the SH kernels
    then move 48 instead of 192 bytes per Gaussian).

    attrs: None, one [P,A] tensor, or a list/tuple of per-Gaussian attribute tensors (<= 8 groups, <= 19 channels in total);
    they are read in place (no concatenation) and each receives its own gradient.
    Returns (images, gs_idx[H,W,K], radii[P], status[2] = (intersections, overflow) on device) where `images` is the
    [4+A,H,W] stack rgb|depth|attrs when `attrs` is a tensor / None, or a list [rgb, depth, attr_0, ...] when it is a list.

    grad_sinks (optional): {"scaling"|"rotation"|"opacity"|"shs"|("attr", i): tensor}, ("attr", i) = the i-th attribute tensor
    of `attrs`.  The backward pass WRITES (not accumulates)
    that input's gradient straight into the given buffer -- e.g. the parameter's slice of a flat gradient buffer -- and
    returns no gradient to autograd for it: no zero-fill, no accumulation pass (one backward per step)."""
    as_list = isinstance(attrs, (list, tuple))
    groups = list(attrs) if as_list else ([attrs] if attrs is not None else [])
    res = _FrameOrtho.apply(position, scaling, rotation, opacity, shs, extr, W, H, K, bg_rgb, nearest, extent, I_cap, cull,
                            ndc, abs_ndc, grad_sinks, *groups)
    imgs, (gs_idx, radii, status) = list(res[:-3]), res[-3:]
    if as_list:
        return imgs, gs_idx, radii, status
    return torch.cat(imgs, 0), gs_idx, radii, status


# ------------------------------------------------------------------------------------------------ deformation
def spline_interval(time: float, num_frames: int, interval_num: int) -> Tuple[int, float]:
    """(interval index, in-interval distance) exactly as the reference computes them on the host
    (dynamic_gaussian_with_base_point_cloud.py:66-68,239-245)."""
    intervals_idx = torch.linspace(0, num_frames - 1, interval_num + 1).long()
    intervals = intervals_idx / (num_frames - 1)
    normed_time = time / (num_frames - 1)
    idx = int(torch.searchsorted(intervals, torch.tensor(normed_time - 1e-7), right=False)) - 1
    idx = max(idx, 0)
    return idx, float(normed_time - float(intervals[idx]))


def spline_to_interval_major(pos_cubic_node: Tensor, interval_num: int) -> Tensor:
    """[P, 4*NI*3] in the reference's (coefficient, interval, xyz) order -> the same numbers interval-major (interval, coefficient,
    xyz): the storage the deformation kernels read with `interval_major=True`.  Values are moved, nothing is recomputed."""
    P = pos_cubic_node.shape[0]
    return pos_cubic_node.reshape(P, 4, interval_num, 3).permute(0, 2, 1, 3).reshape(P, -1).contiguous()


def spline_from_interval_major(node_im: Tensor, interval_num: int) -> Tensor:
    """Inverse of spline_to_interval_major (e.g. before writing a reference-layout checkpoint)."""
    P = node_im.shape[0]
    return node_im.reshape(P, interval_num, 4, 3).permute(0, 2, 1, 3).reshape(P, -1).contiguous()


class _DeformSpline(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, coeff, idx_dev, dist_dev, NI, sink, layout=0):
        L.need_cuda(base, coeff, idx_dev, dist_dev)
        b, c = L.f32c(base), L.f32c(coeff)
        P = b.shape[0]
        pos = torch.empty(P, 3, dtype=torch.float32, device=b.device)
        L.call("spv_deform_spline_forward", P, int(NI), int(layout), L.ptr(b), L.ptr(c), L.ptr(idx_dev), L.ptr(dist_dev), L.ptr(pos), L.stream())
        ctx.meta = (P, int(NI), tuple(coeff.shape), base.requires_grad)
        ctx.layout = int(layout)
        ctx.sink = sink
        ctx.save_for_backward(idx_dev, dist_dev)
        return pos

    @staticmethod
    def backward(ctx, g_pos):
        P, NI, shape, base_grad = ctx.meta
        idx_dev, dist_dev = ctx.saved_tensors
        g_coeff = ctx.sink if ctx.sink is not None else torch.empty(shape, dtype=torch.float32, device=g_pos.device)
        gp = L.f32c(g_pos)
        L.call("spv_deform_spline_backward", P, NI, ctx.layout, L.ptr(idx_dev), L.ptr(dist_dev), L.ptr(gp), L.ptr(g_coeff), 0, L.stream())
        return (gp if base_grad else None), (None if ctx.sink is not None else g_coeff), None, None, None, None, None


def deform_position(base: Tensor, pos_cubic_node: Tensor, idx_dev: Tensor, dist_dev: Tensor, interval_num: int,
                    grad_sink: Optional[Tensor] = None, interval_major: bool = False) -> Tensor:
    """position(t) = base + cubic spline; `pos_cubic_node` is [P, 4*interval_num*3]; idx_dev (int32[1]) / dist_dev
    (float32[1]) are device scalars produced from `spline_interval` (update them in place to replay a CUDA graph).
    grad_sink: optional buffer the coefficient gradient is written into (see render_ortho_frame)."""
    return _DeformSpline.apply(base, pos_cubic_node, idx_dev, dist_dev, interval_num, grad_sink, int(bool(interval_major)))


class _DeformSplinePair(torch.autograd.Function):
    """(pos at ids1, pos at ids2) in one kernel; backward writes both active intervals into a sink kept clean through a
    device-side dirty list (no per-step clear of the whole [P, 4*NI*3] gradient), or into a fresh zero tensor."""

    @staticmethod
    def forward(ctx, base, coeff, idx1, dist1, idx2, dist2, NI, sink, dirty, defer, layout=0):
        L.need_cuda(base, coeff, idx1, dist1, idx2, dist2)
        ctx.defer = defer
        ctx.layout = int(layout)
        b, c = L.f32c(base), L.f32c(coeff)
        P = b.shape[0]
        pos1 = torch.empty(P, 3, dtype=torch.float32, device=b.device)
        pos2 = torch.empty(P, 3, dtype=torch.float32, device=b.device)
        L.call("spv_deform_spline_forward2", P, int(NI), int(layout), L.ptr(b), L.ptr(c), L.ptr(idx1), L.ptr(dist1), L.ptr(idx2), L.ptr(dist2),
               L.ptr(pos1), L.ptr(pos2), L.stream())
        ctx.meta = (P, int(NI), tuple(coeff.shape), base.requires_grad)
        ctx.sink, ctx.dirty = sink, dirty
        ctx.save_for_backward(idx1, dist1, idx2, dist2)
        return pos1, pos2

    @staticmethod
    def backward(ctx, g1, g2):
        P, NI, shape, base_grad = ctx.meta
        idx1, dist1, idx2, dist2 = ctx.saved_tensors
        dev = idx1.device
        if g1 is None:
            g1 = torch.zeros(P, 3, dtype=torch.float32, device=dev)
        g1 = L.f32c(g1)
        g2 = None if g2 is None else L.f32c(g2)
        if ctx.defer is not None:      # frame-parallel: only stage the position gradients; GradExchange finishes the backward
            L.call("spv_deform_defer", P, L.ptr(g1), L.ptr(g2), L.ptr(idx1), L.ptr(dist1), L.ptr(idx2), L.ptr(dist2), L.ptr(ctx.defer),
                   L.stream())
            return (g1 if g2 is None else g1 + g2) if base_grad else None, None, None, None, None, None, None, None, None, None, None
        if ctx.sink is not None:
            g_coeff, dirty, ret = ctx.sink, ctx.dirty, None
        else:
            g_coeff = torch.zeros(shape, dtype=torch.float32, device=dev)
            dirty, ret = torch.zeros(17, dtype=torch.int32, device=dev), g_coeff
        L.call("spv_deform_spline_backward2", P, NI, ctx.layout, L.ptr(idx1), L.ptr(dist1), L.ptr(idx2), L.ptr(dist2), L.ptr(g1), L.ptr(g2),
               L.ptr(dirty), L.ptr(g_coeff), L.stream())
        g_base = None
        if base_grad:
            g_base = g1 if g2 is None else g1 + g2
        return g_base, ret, None, None, None, None, None, None, None, None, None


def deform_position_pair(base: Tensor, pos_cubic_node: Tensor, idx1: Tensor, dist1: Tensor, idx2: Tensor, dist2: Tensor,
                         interval_num: int, grad_sink: Optional[Tensor] = None, dirty: Optional[Tensor] = None,
                         defer: Optional[Tensor] = None, interval_major: bool = False):
    """Positions at the two frame times of a training step (ids1 rendered; ids2 = the `track_gs` attribute,
    src/trainer_fragGS.py:486-508) from ONE pass over the spline coefficients.  Both outputs are differentiable.
    grad_sink + dirty: the coefficient gradient is WRITTEN into `grad_sink` ([P, 4*NI*3], zero-initialised once) and
    `dirty` (int32[17] on the device, zero-initialised once) tracks which intervals hold gradient so only those are cleared on
    the next call.
    defer (frame-parallel training): float32[6P + 4] staging buffer of parallel.GradExchange -- the backward only stores the two
    position gradients there; the coefficient gradient of EVERY rank's frames is rebuilt after the all-gather
    (`spv_deform_spline_backward_gathered`), which is 4x less traffic than exchanging coefficient gradients.
    interval_major: `pos_cubic_node` (and its gradient) are stored interval-major, see spline_to_interval_major."""
    if (grad_sink is None) != (dirty is None):
        raise ValueError("deform_position_pair: grad_sink and dirty go together")
    return _DeformSplinePair.apply(base, pos_cubic_node, idx1, dist1, idx2, dist2, interval_num, grad_sink, dirty, defer,
                                   int(bool(interval_major)))


def rotation_basis(time: float, start_frame_id: int, time_len: int) -> Tensor:
    """[t^0..t^3 | cos(t*pi*(1..4)) | sin(t*pi*(1..4))] with t = (time - start)/time_len (get_rotation, :186-193)."""
    t = (time - start_frame_id) / time_len
    k = torch.arange(1, 5, dtype=torch.float32)
    return torch.cat([torch.pow(torch.tensor(float(t)), torch.arange(4, dtype=torch.float32)), torch.cos(t * k * math.pi),
                      torch.sin(t * k * math.pi)])


class _DeformRotation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rotation, rot_poly_feat, rot_fourier_feat, basis_dev):
        L.need_cuda(rotation, rot_poly_feat, rot_fourier_feat, basis_dev)
        r, pf, ff = L.f32c(rotation), L.f32c(rot_poly_feat), L.f32c(rot_fourier_feat)
        P = r.shape[0]
        out = torch.empty(P, 4, dtype=torch.float32, device=r.device)
        inv = torch.empty(P, dtype=torch.float32, device=r.device)
        L.call("spv_deform_rotation_forward", P, L.ptr(r), L.ptr(pf), L.ptr(ff), L.ptr(basis_dev), L.ptr(out), L.ptr(inv), L.stream())
        ctx.save_for_backward(out, inv)
        return out

    @staticmethod
    def backward(ctx, g):
        out, inv = ctx.saved_tensors
        P = out.shape[0]
        gr = torch.empty(P, 4, dtype=torch.float32, device=out.device)
        L.call("spv_deform_rotation_backward", P, L.ptr(out), L.ptr(inv), L.ptr(L.f32c(g)), L.ptr(gr), L.stream())
        return gr, None, None, None


def deform_rotation(rotation: Tensor, rot_poly_feat: Tensor, rot_fourier_feat: Tensor, basis_dev: Tensor) -> Tensor:
    """Unit quaternion of every Gaussian at frame time t (src/dynamic_gaussian_with_base_point_cloud.py:184-198); the
    poly / Fourier features enter detached, as in the reference.  `basis_dev` = rotation_basis(...).cuda()."""
    return _DeformRotation.apply(rotation, rot_poly_feat, rot_fourier_feat, basis_dev)


class _DeformPolyFourier(torch.autograd.Function):
    @staticmethod
    def forward(ctx, position, pos_poly_feat, pos_fourier_feat, basis_dev):
        L.need_cuda(position, pos_poly_feat, pos_fourier_feat, basis_dev)
        p, pf, ff = L.f32c(position), L.f32c(pos_poly_feat), L.f32c(pos_fourier_feat)
        P = p.shape[0]
        if pf.numel() != P * 12 or ff.numel() != P * 24:
            raise ValueError("deform_position_polyfourier expects pos_poly_feat [P,4,3] and pos_fourier_feat [P,8,3]")
        out = torch.empty(P, 3, dtype=torch.float32, device=p.device)
        L.call("spv_deform_polyfourier_forward", P, L.ptr(p), L.ptr(pf), L.ptr(ff), L.ptr(basis_dev), L.ptr(out), L.stream())
        ctx.save_for_backward(basis_dev)
        ctx.shapes = (tuple(pos_poly_feat.shape), tuple(pos_fourier_feat.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        (basis_dev,) = ctx.saved_tensors
        g = L.f32c(g)
        P, dev = g.shape[0], g.device
        need_p, need_pf, need_ff = ctx.needs_input_grad[:3]
        gp = torch.empty(P, 3, dtype=torch.float32, device=dev) if need_p else None
        gpf = torch.empty(ctx.shapes[0], dtype=torch.float32, device=dev) if need_pf else None
        gff = torch.empty(ctx.shapes[1], dtype=torch.float32, device=dev) if need_ff else None
        L.call("spv_deform_polyfourier_backward", P, L.ptr(basis_dev), L.ptr(g), L.ptr(gp), L.ptr(gpf), L.ptr(gff), L.stream())
        return gp, gpf, gff, None


def deform_position_polyfourier(position: Tensor, pos_poly_feat: Tensor, pos_fourier_feat: Tensor, basis_dev: Tensor,
                                detach_pos: bool = False) -> Tensor:
    """Position of every Gaussian at frame time t in the ALTERNATIVE model (src/dynamic_gaussian_points.py:170-186, get_position):
    position + polynomial(4) + Fourier(8) terms, `basis_dev` = rotation_basis(time, start_frame_id, time_len).cuda() (the same 12
    numbers the rotation uses).  All three tensors are differentiable; `detach_pos` stops the gradient to `position` like the
    reference's flag."""
    return _DeformPolyFourier.apply(position.detach() if detach_pos else position, pos_poly_feat, pos_fourier_feat, basis_dev)
