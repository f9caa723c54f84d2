"""Single-traversal blending of the trainer's three passes (extension on top of the ``dptr.gs`` surface).

The reference renderer walks every tile list three times per frame -- RGB (+first-K ids, bg=white_bg), depth (bg=1.0) and
the attribute stack (bg=0, ``opacity.detach()``, ``ndc.detach()``) -- /root/reference/src/pointrix/renderer/
dptr_ortho_enhanced.py:342-376.  ``blend_rgb_depth_attrs`` returns the same three images, ids and gradients from ONE
forward and ONE backward traversal over ``cat([rgb, depth, attrs])``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L
from . import alpha_blending, alpha_blending_enhanced

MAX_FUSED_CHANNELS = 23


class _BlendGroups(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, idx_sorted, tile_range, bg_rgb, W, H, ndc, abs_ndc, K):
        L.need_cuda(uv, conic, opacity, feature, idx_sorted, tile_range)
        uv_c, conic_c, op_c, feat_c = L.f32c(uv), L.f32c(conic), L.f32c(opacity), L.f32c(feature)
        idx_c = idx_sorted.to(torch.int32).contiguous()
        tr_c = tile_range.to(torch.int32).contiguous()
        P, C = feat_c.shape
        dev = feat_c.device
        rendered = torch.empty(C, H, W, dtype=torch.float32, device=dev)
        final_T = torch.empty(H, W, dtype=torch.float32, device=dev)
        ncontrib = torch.empty(H, W, dtype=torch.int32, device=dev)
        gs_idx = torch.empty(H, W, K, dtype=torch.int32, device=dev)
        L.call("spv_alpha_blend_groups_forward", P, C, int(W), int(H), int(K), L.ptr(uv_c), L.ptr(conic_c), L.ptr(op_c),
               L.ptr(feat_c), L.ptr(idx_c), L.ptr(tr_c), float(bg_rgb), 1.0, 0.0, L.ptr(rendered), L.ptr(final_T),
               L.ptr(ncontrib), L.ptr(gs_idx), L.stream())
        ctx.meta = (float(bg_rgb), int(W), int(H), ndc is not None, abs_ndc is not None)
        ctx.save_for_backward(uv_c, conic_c, op_c, feat_c, idx_c, tr_c, final_T, ncontrib)
        ctx.mark_non_differentiable(ncontrib, gs_idx)
        return rendered, ncontrib, gs_idx

    @staticmethod
    def backward(ctx, dL_drendered, _n, _g):
        bg_rgb, W, H, has_ndc, has_abs = ctx.meta
        uv, conic, opacity, feature, idx_sorted, tile_range, final_T, ncontrib = ctx.saved_tensors
        P, C = feature.shape
        dev = feature.device
        g_uv = torch.empty(P, 2, dtype=torch.float32, device=dev)
        g_uv_rgb = torch.empty(P, 2, dtype=torch.float32, device=dev)
        g_abs = torch.empty(P, 2, dtype=torch.float32, device=dev)
        g_conic = torch.empty(P, 3, dtype=torch.float32, device=dev)
        g_op = torch.empty(P, 1, dtype=torch.float32, device=dev)
        g_feat = torch.empty(P, C, dtype=torch.float32, device=dev)
        ws = torch.empty(L.query("spv_alpha_blend_groups_backward_workspace_bytes", P), dtype=torch.uint8, device=dev)
        L.call("spv_alpha_blend_groups_backward", P, C, W, H, L.ptr(uv), L.ptr(conic), L.ptr(opacity), L.ptr(feature),
               L.ptr(idx_sorted), L.ptr(tile_range), bg_rgb, 1.0, 0.0, L.ptr(final_T), L.ptr(ncontrib),
               L.ptr(L.f32c(dL_drendered)), L.ptr(g_uv), L.ptr(g_uv_rgb), L.ptr(g_abs), L.ptr(g_conic), L.ptr(g_op),
               L.ptr(g_feat), L.ptr(ws), ws.numel(), L.stream())
        g_ndc = g_abs_ndc = None
        if has_ndc or has_abs:
            scale = torch.tensor([0.5 * W, 0.5 * H], dtype=torch.float32, device=dev)
            if has_ndc:
                g_ndc = g_uv_rgb * scale[None, :]
            if has_abs:
                g_abs_ndc = g_abs * scale[None, :]
        return g_uv, g_conic, g_op, g_feat, None, None, None, None, None, g_ndc, g_abs_ndc, None


def blend_rgb_depth_attrs(uv: Tensor, conic: Tensor, opacity: Tensor, rgb: Tensor, depth: Tensor, attrs: Optional[Tensor],
                          idx_sorted: Tensor, tile_range: Tensor, bg_color: float, W: int, H: int,
                          ndc: Optional[Tensor] = None, abs_ndc: Optional[Tensor] = None, K: int = 10
                          ) -> Tuple[Tensor, Tensor, Optional[Tensor], Tensor]:
    """-> (rgb_img[3,H,W], depth_img[1,H,W], attr_img[A,H,W] | None, gs_idx[H,W,K])."""
    A = 0 if attrs is None else attrs.shape[1]
    if rgb.shape[1] != 3 or depth.shape[1] != 1 or 4 + A > MAX_FUSED_CHANNELS or K <= 0:
        # shapes outside the fused kernel's envelope: the reference's three calls
        img, _, gs_idx = alpha_blending_enhanced(uv, conic, opacity, rgb, idx_sorted, tile_range, bg_color, W, H, ndc, abs_ndc, K=K)
        dimg = alpha_blending(uv, conic, opacity, depth, idx_sorted, tile_range, 1.0, W, H, None if ndc is None else ndc.detach())
        aimg = None
        if attrs is not None:
            aimg = alpha_blending(uv, conic, opacity.detach(), attrs, idx_sorted, tile_range, 0.0, W, H,
                                  None if ndc is None else ndc.detach())
        return img, dimg, aimg, gs_idx
    feats = [rgb, depth] + ([attrs] if attrs is not None else [])
    feature = torch.cat(feats, dim=1)
    img, _, gs_idx = _BlendGroups.apply(uv, conic, opacity, feature, idx_sorted, tile_range, bg_color, W, H, ndc, abs_ndc, K)
    return img[:3], img[3:4], (img[4:] if attrs is not None else None), gs_idx
