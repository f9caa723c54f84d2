"""Fused image losses of the training step behind `spv_loss_*` (csrc/loss.cu).

Each function mirrors one expression of the reference trainer and returns a scalar that autograd can differentiate w.r.t. the
rendered image; the gradient is produced by the same C call as the value (no second pass over the images, no host sync), so
render -> loss -> backward stays on one stream and can sit inside one CUDA graph.

  rgb_loss(pred[3,H,W], gt[H,W,3])        (1-l) * l1_loss + l * (1 - ssim)          trainer_fragGS.py:573-578
  depth_loss_dpt(pred[H,W,*], gt[H,W,*])  median / mean-abs-deviation normalised MSE  src/loss.py:184-206
  track_loss(track[3,H,W], ...)           trimmed, confidence-weighted L1 / max(H,W)  trainer_fragGS.py:531-571

There is no CPU fallback: a CPU tensor raises, a missing library raises (see _lib.load).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib as L


def _ws(nbytes: int, dev) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


class _RGBLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_chw: Tensor, gt_hwc: Tensor, lambda_dssim: float, weight: float):
        L.need_cuda(pred_chw, gt_hwc)
        if pred_chw.dim() != 3 or pred_chw.shape[0] != 3 or gt_hwc.shape != (pred_chw.shape[1], pred_chw.shape[2], 3):
            raise ValueError(f"rgb_loss expects pred [3,H,W] and gt [H,W,3], got {tuple(pred_chw.shape)} and {tuple(gt_hwc.shape)}")
        p, g = L.f32c(pred_chw.detach()), L.f32c(gt_hwc.detach())
        H, W = int(p.shape[1]), int(p.shape[2])
        need_grad = ctx.needs_input_grad[0]
        out = torch.empty(3, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p) if need_grad else None
        nbytes = L.query("spv_loss_rgb_workspace_bytes", W, H)
        ws = _ws(nbytes, p.device)
        L.call("spv_loss_rgb", W, H, L.ptr(p), L.ptr(g), float(weight), float(lambda_dssim), L.ptr(out),
               L.ptr(grad) if need_grad else None, L.ptr(ws), nbytes, L.stream())
        ctx.grad = grad
        l1, ssim = out[1], out[2]
        ctx.mark_non_differentiable(l1, ssim)
        return out[0], l1, ssim

    @staticmethod
    def backward(ctx, g_total, _g_l1, _g_ssim):
        return ctx.grad * g_total, None, None, None


def rgb_loss(pred_chw: Tensor, gt_hwc: Tensor, lambda_dssim: float = 0.2, weight: float = 1.0) -> Tuple[Tensor, Tensor, Tensor]:
    """`weight * ((1 - lambda) * l1_loss(p, g) + lambda * (1 - ssim(p, g)))` exactly as trainer_fragGS.py:573-578 evaluates it
    on `[1,H,W,3]` tensors (so `ssim` windows over (x, colour) inside each row, pointrix/model/loss.py:83).  `pred_chw` is the
    renderer's `rgb[0]` ([3,H,W], no permute needed), `gt_hwc` the ground-truth frame as the trainer stores it.
    Returns (loss, l1, ssim); only `loss` carries gradient."""
    return _RGBLoss.apply(pred_chw, gt_hwc, lambda_dssim, weight)


class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: Tensor, gt: Tensor, weight: float):
        L.need_cuda(pred, gt)
        if pred.numel() != gt.numel() or pred.numel() == 0:
            raise ValueError("depth_loss_dpt expects two non-empty maps of the same size")
        p, g = L.f32c(pred.detach()), L.f32c(gt.detach())
        n = p.numel()
        need_grad = ctx.needs_input_grad[0]
        out = torch.empty(1, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p) if need_grad else None
        nbytes = L.query("spv_loss_depth_workspace_bytes", n)
        ws = _ws(nbytes, p.device)
        L.call("spv_loss_depth_dpt", n, L.ptr(p), L.ptr(g), float(weight), L.ptr(out), L.ptr(grad) if need_grad else None,
               L.ptr(ws), nbytes, L.stream())
        ctx.grad = grad
        return out[0]

    @staticmethod
    def backward(ctx, g_out):
        return ctx.grad * g_out, None, None


def depth_loss_dpt(pred_depth: Tensor, gt_depth: Tensor, weight: float = 1.0) -> Tensor:
    """src/loss.py:184-206 without the optional weight map (the trainer passes none, trainer_fragGS.py:599-601)."""
    return _DepthLoss.apply(pred_depth, gt_depth, weight)


class _TrackLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, track_chw: Tensor, query_xy: Tensor, target_xy: Tensor, visible: Tensor, weights: Tensor, quantile: float,
                weight: float):
        L.need_cuda(track_chw, query_xy, target_xy, visible, weights)
        if track_chw.dim() != 3 or track_chw.shape[0] < 2:
            raise ValueError("track_loss expects the rendered track image [>=2,H,W]")
        t = L.f32c(track_chw.detach())
        H, W = int(t.shape[1]), int(t.shape[2])
        q = query_xy.to(torch.int32).contiguous()
        n = int(q.shape[0])
        tgt, wts = L.f32c(target_xy), L.f32c(weights.reshape(-1))
        vis = visible.reshape(-1).to(torch.uint8).contiguous()
        if not (tgt.shape == (n, 2) and wts.numel() == n and vis.numel() == n):
            raise ValueError("track_loss: query_xy [n,2], target_xy [n,2], visible [n], weights [n] must agree")
        need_grad = ctx.needs_input_grad[0]
        out = torch.empty(1, dtype=torch.float32, device=t.device)
        grad = torch.zeros_like(t) if need_grad else None       # channels >= 2 (the depth of the track) get no gradient
        nbytes = L.query("spv_loss_track_workspace_bytes", n)
        ws = _ws(nbytes, t.device)
        L.call("spv_loss_track", n, W, H, L.ptr(t), L.ptr(q), L.ptr(tgt), L.ptr(vis), L.ptr(wts), float(quantile), float(weight),
               L.ptr(out), L.ptr(grad) if need_grad else None, L.ptr(ws), nbytes, L.stream())
        ctx.grad = grad
        return out[0]

    @staticmethod
    def backward(ctx, g_out):
        return ctx.grad * g_out, None, None, None, None, None, None


def track_loss(track_chw: Tensor, query_xy: Tensor, target_xy: Tensor, visible: Tensor, weights: Tensor, quantile: float = 0.98,
               weight: float = 1.0) -> Tensor:
    """The optical-flow loss of trainer_fragGS.py:531-571: `masked_l1_loss(denormalize(track)[query][visible], target[visible],
    mask=weights[visible], quantile=0.98) / max(H, W)` (src/criterion.py:46-51, src/util.py:82).  Point i of `query_xy` (integer
    pixel x,y) pairs with row i of the other arrays; zero loss and no gradient when nothing is visible."""
    return _TrackLoss.apply(track_chw, query_xy, target_xy, visible, weights, quantile, weight)


# ---- value + gradient image in one call, no autograd node ----------------------------------------------------------------------
# A training step that hands `dL/dimage` straight to the rasterizer's backward (torch.autograd.backward(images, grads)) does not
# need the losses on the autograd tape: one C call per loss gives the scalar(s) and the gradient image, with nothing multiplied
# by an upstream 1.0 afterwards.
def rgb_loss_grad(pred_chw: Tensor, gt_hwc: Tensor, lambda_dssim: float = 0.2, weight: float = 1.0, buffers: dict = None) -> Tuple[Tensor, Tensor]:
    """-> (float32[3] = (loss, l1, ssim) on the device, dL/dpred [3,H,W]).  `buffers`: optional dict the call fills with its output /
    gradient / workspace tensors on first use and re-uses afterwards (lets a caller run the losses on side streams without
    allocating there)."""
    L.need_cuda(pred_chw, gt_hwc)
    p, g = L.f32c(pred_chw.detach()), L.f32c(gt_hwc)
    H, W = int(p.shape[1]), int(p.shape[2])
    nbytes = L.query("spv_loss_rgb_workspace_bytes", W, H)
    buffers = buffers if buffers is not None else {}
    if "out" not in buffers:
        buffers.update(out=torch.empty(3, dtype=torch.float32, device=p.device), grad=torch.empty_like(p), ws=_ws(nbytes, p.device))
    out, grad, ws = buffers["out"], buffers["grad"], buffers["ws"]
    L.call("spv_loss_rgb", W, H, L.ptr(p), L.ptr(g), float(weight), float(lambda_dssim), L.ptr(out), L.ptr(grad), L.ptr(ws), nbytes, L.stream())
    return out, grad


def depth_loss_grad(pred_depth: Tensor, gt_depth: Tensor, weight: float = 1.0, buffers: dict = None) -> Tuple[Tensor, Tensor]:
    """-> (float32[1] loss on the device, dL/dpred with pred's shape).  `buffers`: see rgb_loss_grad."""
    L.need_cuda(pred_depth, gt_depth)
    p, g = L.f32c(pred_depth.detach()), L.f32c(gt_depth)
    n = p.numel()
    nbytes = L.query("spv_loss_depth_workspace_bytes", n)
    buffers = buffers if buffers is not None else {}
    if "out" not in buffers:
        buffers.update(out=torch.empty(1, dtype=torch.float32, device=p.device), grad=torch.empty_like(p), ws=_ws(nbytes, p.device))
    out, grad, ws = buffers["out"], buffers["grad"], buffers["ws"]
    L.call("spv_loss_depth_dpt", n, L.ptr(p), L.ptr(g), float(weight), L.ptr(out), L.ptr(grad), L.ptr(ws), nbytes, L.stream())
    return out, grad


def track_loss_grad(track_chw: Tensor, query_xy: Tensor, target_xy: Tensor, visible: Tensor, weights: Tensor, quantile: float = 0.98,
                    weight: float = 1.0, grad: Tensor = None, buffers: dict = None) -> Tuple[Tensor, Tensor]:
    """-> (float32[1] loss on the device, dL/dtrack [C,H,W] with channels >= 2 zero).  query_xy int32 [n,2], visible uint8 [n]
    (already in the kernel's dtypes: no conversion kernels on the hot path).  `grad`: optional caller-owned [C,H,W] buffer whose
    channels >= 2 are already zero (only the two coordinate planes are cleared and rewritten)."""
    L.need_cuda(track_chw, query_xy, target_xy, visible, weights)
    t = L.f32c(track_chw.detach())
    H, W = int(t.shape[1]), int(t.shape[2])
    n = int(query_xy.shape[0])
    if query_xy.dtype != torch.int32 or visible.dtype != torch.uint8:
        raise ValueError("track_loss_grad expects int32 query pixels and a uint8 visibility mask")
    nbytes = L.query("spv_loss_track_workspace_bytes", n)
    buffers = buffers if buffers is not None else {}
    if "out" not in buffers:
        buffers.update(out=torch.empty(1, dtype=torch.float32, device=t.device), ws=_ws(nbytes, t.device))
    out, ws = buffers["out"], buffers["ws"]
    if grad is None:
        grad = torch.zeros_like(t)
    L.call("spv_loss_track", n, W, H, L.ptr(t), L.ptr(query_xy), L.ptr(L.f32c(target_xy)), L.ptr(visible), L.ptr(L.f32c(weights.reshape(-1))),
           float(quantile), float(weight), L.ptr(out), L.ptr(grad), L.ptr(ws), nbytes, L.stream())
    return out, grad
