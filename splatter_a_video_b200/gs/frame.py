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


class Capacity:
    """Intersection-capacity policy of the sync-free path.  The count of tile intersections never leaves the device on the
    hot path; it is copied to pinned host memory asynchronously and inspected one call later.  On overflow the affected
    frame was rendered from a truncated list, so `check()` raises and the caller re-renders with the grown capacity."""

    def __init__(self, initial: int = 0, growth: float = 1.5, slack: float = 1.25):
        self.I_cap = int(initial)
        self.growth, self.slack = growth, slack
        self._host = torch.zeros(2, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else torch.zeros(2, dtype=torch.int32)
        self._event = None
        self.last_I = 0

    def observe(self, status: Tensor):
        self._host.copy_(status, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def check(self, wait: bool = False) -> bool:
        """True if the last observed frame fitted.  Grows the capacity when it did not."""
        if self._event is None:
            return True
        if not wait and not self._event.query():
            return True
        self._event.synchronize()
        self._event = None
        self.last_I = int(self._host[0])
        if int(self._host[1]) != 0:
            self.I_cap = int(self.I_cap * self.growth) + 1024
            return False
        return True


class _FrameOrtho(torch.autograd.Function):
    @staticmethod
    def forward(ctx, position, scaling, rotation, opacity, shs, attrs, extr, W, H, K, bg_rgb, nearest, extent, I_cap, cull,
                ndc, abs_ndc, sinks):
        L.need_cuda(position, scaling, rotation, opacity, shs, attrs, extr)
        pos, sc, rot, op, sh = (L.f32c(x) for x in (position, scaling, rotation, opacity, shs))
        at = L.f32c(attrs) if attrs is not None else None
        ex = L.f32c(extr)
        P = pos.shape[0]
        A = 0 if at is None else at.shape[1]
        if sh.shape[1] != 16:
            raise ValueError("render_ortho_frame needs degree-3 SH coefficients [P,16,3]")
        dev = pos.device
        images = torch.empty(4 + A, H, W, dtype=torch.float32, device=dev)
        gs_idx = torch.empty(H, W, K, dtype=torch.int32, device=dev)
        radii = torch.empty(P, dtype=torch.int32, device=dev)
        status = torch.empty(2, dtype=torch.int32, device=dev)
        nbytes = L.query("spv_frame_workspace_bytes", P, int(I_cap), int(W), int(H), A)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        L.call("spv_frame_ortho_forward", P, int(W), int(H), A, int(K), int(I_cap), int(bool(cull)), L.ptr(pos), L.ptr(sc),
               L.ptr(rot), L.ptr(op), L.ptr(sh), L.ptr(at), L.ptr(ex), float(nearest), float(extent), float(bg_rgb),
               L.ptr(images), L.ptr(gs_idx), L.ptr(radii), L.ptr(status), L.ptr(ws), nbytes, L.stream())
        ctx.meta = (P, int(W), int(H), A, int(I_cap), float(bg_rgb), ndc is not None, abs_ndc is not None)
        ctx.sinks = dict(sinks) if sinks else {}
        ctx.save_for_backward(sc, rot, op, sh, ex, ws)
        ctx.mark_non_differentiable(gs_idx, radii, status)
        return images, gs_idx, radii, status

    @staticmethod
    def backward(ctx, g_images, _g, _r, _s):
        P, W, H, A, I_cap, bg_rgb, has_ndc, has_abs = ctx.meta
        sc, rot, op, sh, ex, ws = ctx.saved_tensors
        dev = sc.device
        sinks = ctx.sinks

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
        g_sh, r_sh = out("shs", P, 16, 3)
        g_at = torch.empty(P, max(A, 1), dtype=torch.float32, device=dev)
        g_ndc = torch.empty(P, 2, dtype=torch.float32, device=dev) if has_ndc else None
        g_abs = torch.empty(P, 2, dtype=torch.float32, device=dev) if has_abs else None
        L.call("spv_frame_ortho_backward", P, W, H, A, I_cap, L.ptr(sc), L.ptr(rot), L.ptr(op), L.ptr(sh), L.ptr(ex), bg_rgb,
               L.ptr(L.f32c(g_images)), L.ptr(g_pos), L.ptr(g_sc), L.ptr(g_rot), L.ptr(g_op), L.ptr(g_sh), L.ptr(g_at),
               L.ptr(g_ndc), L.ptr(g_abs), L.ptr(ws), ws.numel(), L.stream())
        return (g_pos, r_sc, r_rot, r_op, r_sh, g_at if A > 0 else None, None, None, None, None, None, None, None, None, None,
                g_ndc, g_abs, None)


def render_ortho_frame(position: Tensor, scaling: Tensor, rotation: Tensor, opacity: Tensor, shs: Tensor,
                       attrs: Optional[Tensor], extr: Tensor, W: int, H: int, K: int, bg_rgb: float, I_cap: int,
                       cull: bool = True, nearest: float = 0.01, extent: float = 1.3, ndc: Optional[Tensor] = None,
                       abs_ndc: Optional[Tensor] = None, grad_sinks: Optional[dict] = None
                       ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (images[4+A,H,W] = rgb|depth|attrs, gs_idx[H,W,K], radii[P], status[2] = (intersections, overflow) on device).

    grad_sinks (optional): {"scaling"|"rotation"|"opacity"|"shs": tensor}.  The backward pass WRITES (not accumulates)
    that input's gradient straight into the given buffer -- e.g. the parameter's slice of a flat gradient buffer -- and
    returns no gradient to autograd for it: no zero-fill, no accumulation pass (one backward per step)."""
    return _FrameOrtho.apply(position, scaling, rotation, opacity, shs, attrs, extr, W, H, K, bg_rgb, nearest, extent, I_cap,
                             cull, ndc, abs_ndc, grad_sinks)


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


class _DeformSpline(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, coeff, idx_dev, dist_dev, NI, sink):
        L.need_cuda(base, coeff, idx_dev, dist_dev)
        b, c = L.f32c(base), L.f32c(coeff)
        P = b.shape[0]
        pos = torch.empty(P, 3, dtype=torch.float32, device=b.device)
        L.call("spv_deform_spline_forward", P, int(NI), L.ptr(b), L.ptr(c), L.ptr(idx_dev), L.ptr(dist_dev), L.ptr(pos), L.stream())
        ctx.meta = (P, int(NI), tuple(coeff.shape), base.requires_grad)
        ctx.sink = sink
        ctx.save_for_backward(idx_dev, dist_dev)
        return pos

    @staticmethod
    def backward(ctx, g_pos):
        P, NI, shape, base_grad = ctx.meta
        idx_dev, dist_dev = ctx.saved_tensors
        g_coeff = ctx.sink if ctx.sink is not None else torch.empty(shape, dtype=torch.float32, device=g_pos.device)
        gp = L.f32c(g_pos)
        L.call("spv_deform_spline_backward", P, NI, L.ptr(idx_dev), L.ptr(dist_dev), L.ptr(gp), L.ptr(g_coeff), 0, L.stream())
        return (gp if base_grad else None), (None if ctx.sink is not None else g_coeff), None, None, None, None


def deform_position(base: Tensor, pos_cubic_node: Tensor, idx_dev: Tensor, dist_dev: Tensor, interval_num: int,
                    grad_sink: Optional[Tensor] = None) -> Tensor:
    """position(t) = base + cubic spline; `pos_cubic_node` is [P, 4*interval_num*3]; idx_dev (int32[1]) / dist_dev
    (float32[1]) are device scalars produced from `spline_interval` (update them in place to replay a CUDA graph).
    grad_sink: optional buffer the coefficient gradient is written into (see render_ortho_frame)."""
    return _DeformSpline.apply(base, pos_cubic_node, idx_dev, dist_dev, interval_num, grad_sink)
