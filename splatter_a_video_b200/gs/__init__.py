"""Drop-in replacement for the reference's ``dptr.gs`` operator module (boundary B1).

Same function names, positional signatures, dtypes, return arity and autograd contract as
/root/reference/src/submodules/dptr/dptr/gs/{__init__,project_point,compute_cov3d,ewa_project,sort_gaussian,
compute_sh,compute_sh_free,alpha_blending,alpha_blending_enhanced,alpha_blending_with_bias}.py, so
``import dptr.gs as gs`` call sites (src/trainer_fragGS.py:29, src/pointrix/renderer/dptr*.py:3) run unmodified
through the ``dptr`` alias package at the repo root.  Every op is a ``torch.autograd.Function`` whose forward /
backward call the sm_100a kernels through the C ABI (include/spv_b200.h) on torch's current stream.

There is no CPU path: like the reference (include/utils.h:9-10) inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L

__all__ = [
    "project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "compute_sh", "compute_sh_free",
    "alpha_blending", "rasterization", "alpha_blending_enhanced", "alpha_blending_with_bias",
    "project_point_ortho", "ewa_project_ortho",
]

_NUM_SH = (1, 4, 9, 16)


def _vis_u8(visible: Optional[Tensor], P: int, device) -> Tensor:
    if visible is None:
        return torch.ones(P, dtype=torch.uint8, device=device)
    v = visible.reshape(-1)
    if v.dtype == torch.bool:
        return v.contiguous().view(torch.uint8)
    return (v != 0).contiguous().view(torch.uint8)


def _extr12(extr: Tensor) -> Tensor:
    """[3,4] or [4,4] extrinsics -> contiguous fp32 buffer whose first 12 floats are the row-major 3x4 [R|t]."""
    e = L.f32c(extr)
    if e.numel() < 12:
        raise ValueError("extr must hold at least 3x4 values")
    return e


def _grad_like_extr(g12: Tensor, shape) -> Tensor:
    n = 1
    for d in shape:
        n *= d
    out = torch.zeros(n, dtype=torch.float32, device=g12.device)
    out[:12] = g12
    return out.reshape(shape)


# ------------------------------------------------------------------------------------------------ project_point
class _ProjectPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, intr, extr, W, H, nearest, extent):
        L.need_cuda(xyz, intr, extr)
        xyz_c, intr_c, extr_c = L.f32c(xyz), L.f32c(intr), _extr12(extr)
        P = xyz_c.shape[0]
        uv = torch.empty(P, 2, dtype=torch.float32, device=xyz.device)
        depth = torch.empty(P, 1, dtype=torch.float32, device=xyz.device)
        L.call("spv_project_point_forward", P, L.ptr(xyz_c), L.ptr(intr_c), L.ptr(extr_c), int(W), int(H),
               float(nearest), float(extent), L.ptr(uv), L.ptr(depth), L.stream())
        ctx.save_for_backward(xyz_c, intr_c, extr_c, depth)
        ctx.extr_shape = tuple(extr.shape)
        ctx.need = (intr.requires_grad, extr.requires_grad)
        return uv, depth

    @staticmethod
    def backward(ctx, dL_duv, dL_ddepth):
        xyz, intr, extr, depth = ctx.saved_tensors
        P = xyz.shape[0]
        need_intr, need_extr = ctx.need
        dL_dxyz = torch.empty(P, 3, dtype=torch.float32, device=xyz.device)
        g_intr = torch.empty(4, dtype=torch.float32, device=xyz.device) if need_intr else None
        g_extr = torch.empty(12, dtype=torch.float32, device=xyz.device) if need_extr else None
        L.call("spv_project_point_backward", P, L.ptr(xyz), L.ptr(intr), L.ptr(extr), L.ptr(depth),
               L.ptr(L.f32c(dL_duv)), L.ptr(L.f32c(dL_ddepth)), L.ptr(dL_dxyz), L.ptr(g_intr), L.ptr(g_extr),
               L.stream())
        return (dL_dxyz, g_intr, _grad_like_extr(g_extr, ctx.extr_shape) if need_extr else None,
                None, None, None, None)


def project_point(xyz: Tensor, intr: Tensor, extr: Tensor, W: int, H: int, nearest: float = 0.2,
                  extent: float = 1.3) -> Tuple[Tensor, Tensor]:
    """Perspective projection + near/extent culling (gs/project_point.py:8-45; K1/K2).
    Returns ``(uv[P,2], depth[P,1])``; culled rows are exactly 0."""
    return _ProjectPoint.apply(xyz, intr, extr, W, H, nearest, extent)


class _ProjectPointOrtho(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, extr, W, H, nearest, extent):
        L.need_cuda(xyz, extr)
        xyz_c, extr_c = L.f32c(xyz), _extr12(extr)
        P = xyz_c.shape[0]
        uv = torch.empty(P, 2, dtype=torch.float32, device=xyz.device)
        depth = torch.empty(P, 1, dtype=torch.float32, device=xyz.device)
        L.call("spv_project_point_ortho_forward", P, L.ptr(xyz_c), L.ptr(extr_c), int(W), int(H), float(nearest),
               float(extent), L.ptr(uv), L.ptr(depth), L.stream())
        ctx.save_for_backward(extr_c, depth)
        ctx.WH = (int(W), int(H))
        return uv, depth

    @staticmethod
    def backward(ctx, dL_duv, dL_ddepth):
        extr, depth = ctx.saved_tensors
        P = depth.shape[0]
        g = torch.empty(P, 3, dtype=torch.float32, device=depth.device)
        L.call("spv_project_point_ortho_backward", P, L.ptr(extr), ctx.WH[0], ctx.WH[1], L.ptr(depth),
               L.ptr(L.f32c(dL_duv)), L.ptr(L.f32c(dL_ddepth)), L.ptr(g), L.stream())
        return g, None, None, None, None, None


def project_point_ortho(xyz: Tensor, extr: Tensor, W: int, H: int, nearest: float = 0.2,
                        extent: float = 1.3) -> Tuple[Tensor, Tensor]:
    """Orthographic projection of the video trainer as ONE kernel (the reference runs ~12 torch kernels:
    src/pointrix/renderer/dptr_ortho_enhanced.py:145-202).  Extension: not part of the reference ``dptr.gs``."""
    return _ProjectPointOrtho.apply(xyz, extr, W, H, nearest, extent)


# ------------------------------------------------------------------------------------------------ compute_cov3d
class _ComputeCov3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales, uquats, visible):
        L.need_cuda(scales, uquats, visible)
        s, q = L.f32c(scales), L.f32c(uquats)
        P = s.shape[0]
        vis = _vis_u8(visible, P, s.device)
        cov3d = torch.empty(P, 6, dtype=torch.float32, device=s.device)
        L.call("spv_compute_cov3d_forward", P, L.ptr(s), L.ptr(q), L.ptr(vis), L.ptr(cov3d), L.stream())
        ctx.save_for_backward(s, q, vis)
        return cov3d

    @staticmethod
    def backward(ctx, dL_dcov3d):
        s, q, vis = ctx.saved_tensors
        P = s.shape[0]
        gs = torch.empty(P, 3, dtype=torch.float32, device=s.device)
        gq = torch.empty(P, 4, dtype=torch.float32, device=s.device)
        L.call("spv_compute_cov3d_backward", P, L.ptr(s), L.ptr(q), L.ptr(vis), L.ptr(L.f32c(dL_dcov3d)), L.ptr(gs),
               L.ptr(gq), L.stream())
        return gs, gq, None


def compute_cov3d(scales: Tensor, uquats: Tensor, visible: Optional[Tensor] = None) -> Tensor:
    """Sigma = R S^2 R^T, upper triangle ``[P,6]`` (gs/compute_cov3d.py:7-33; K3/K4)."""
    if visible is None:
        visible = torch.ones_like(scales[:, 0], dtype=torch.bool)
    return _ComputeCov3D.apply(scales, uquats, visible)


# ------------------------------------------------------------------------------------------------ ewa_project
class _EWAProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, cov3d, intr, extr, uv, W, H, visible):
        L.need_cuda(xyz, cov3d, intr, extr, uv, visible)
        xyz_c, cov_c, intr_c, extr_c, uv_c = L.f32c(xyz), L.f32c(cov3d), L.f32c(intr), _extr12(extr), L.f32c(uv)
        P = xyz_c.shape[0]
        vis = _vis_u8(visible, P, xyz.device)
        conic = torch.empty(P, 3, dtype=torch.float32, device=xyz.device)
        radius = torch.empty(P, dtype=torch.int32, device=xyz.device)
        tiles = torch.empty(P, dtype=torch.int32, device=xyz.device)
        L.call("spv_ewa_project_forward", P, L.ptr(xyz_c), L.ptr(cov_c), L.ptr(intr_c), L.ptr(extr_c), L.ptr(uv_c),
               int(W), int(H), L.ptr(vis), L.ptr(conic), L.ptr(radius), L.ptr(tiles), L.stream())
        ctx.save_for_backward(xyz_c, cov_c, intr_c, extr_c, radius)
        ctx.extr_shape = tuple(extr.shape)
        ctx.need = (intr.requires_grad, extr.requires_grad)
        ctx.mark_non_differentiable(radius, tiles)
        return conic, radius, tiles

    @staticmethod
    def backward(ctx, dL_dconic, _r, _t):
        xyz, cov3d, intr, extr, radius = ctx.saved_tensors
        P = xyz.shape[0]
        need_intr, need_extr = ctx.need
        gx = torch.empty(P, 3, dtype=torch.float32, device=xyz.device)
        gc = torch.empty(P, 6, dtype=torch.float32, device=xyz.device)
        g_intr = torch.empty(4, dtype=torch.float32, device=xyz.device) if need_intr else None
        g_extr = torch.empty(12, dtype=torch.float32, device=xyz.device) if need_extr else None
        L.call("spv_ewa_project_backward", P, L.ptr(xyz), L.ptr(cov3d), L.ptr(intr), L.ptr(extr), L.ptr(radius),
               L.ptr(L.f32c(dL_dconic)), L.ptr(gx), L.ptr(gc), L.ptr(g_intr), L.ptr(g_extr), L.stream())
        return (gx, gc, g_intr, _grad_like_extr(g_extr, ctx.extr_shape) if need_extr else None,
                None, None, None, None)


def ewa_project(xyz: Tensor, cov3d: Tensor, intr: Tensor, extr: Tensor, uv: Tensor, W: int, H: int,
                visible: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Perspective EWA splat: ``(conic[P,3], radius[P] int32, tiles[P] int32)`` (gs/ewa_project.py:8-53; K5/K6)."""
    if visible is None:
        visible = torch.ones_like(uv[:, 0], dtype=torch.bool)
    return _EWAProject.apply(xyz, cov3d, intr, extr, uv, W, H, visible)


class _EWAProjectOrtho(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cov3d, extr, uv, W, H, visible):
        L.need_cuda(cov3d, extr, uv, visible)
        cov_c, extr_c, uv_c = L.f32c(cov3d), _extr12(extr), L.f32c(uv)
        P = cov_c.shape[0]
        vis = _vis_u8(visible, P, cov3d.device)
        conic = torch.empty(P, 3, dtype=torch.float32, device=cov3d.device)
        radius = torch.empty(P, dtype=torch.int32, device=cov3d.device)
        tiles = torch.empty(P, dtype=torch.int32, device=cov3d.device)
        L.call("spv_ewa_project_ortho_forward", P, L.ptr(cov_c), L.ptr(extr_c), L.ptr(uv_c), int(W), int(H),
               L.ptr(vis), L.ptr(conic), L.ptr(radius), L.ptr(tiles), L.stream())
        ctx.save_for_backward(cov_c, extr_c, radius)
        ctx.WH = (int(W), int(H))
        ctx.mark_non_differentiable(radius, tiles)
        return conic, radius, tiles

    @staticmethod
    def backward(ctx, dL_dconic, _r, _t):
        cov3d, extr, radius = ctx.saved_tensors
        P = cov3d.shape[0]
        gc = torch.empty(P, 6, dtype=torch.float32, device=cov3d.device)
        L.call("spv_ewa_project_ortho_backward", P, L.ptr(cov3d), L.ptr(extr), ctx.WH[0], ctx.WH[1], L.ptr(radius),
               L.ptr(L.f32c(dL_dconic)), L.ptr(gc), L.stream())
        return gc, None, None, None, None, None


def ewa_project_ortho(cov3d: Tensor, extr: Tensor, uv: Tensor, W: int, H: int,
                      visible: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Orthographic EWA of the video trainer as ONE kernel (the reference runs ~40 torch kernels:
    ``ewa_project_torch_impl``, src/pointrix/renderer/dptr_ortho_enhanced.py:18-111).  Extension."""
    if visible is None:
        visible = torch.ones_like(uv[:, 0], dtype=torch.bool)
    return _EWAProjectOrtho.apply(cov3d, extr, uv, W, H, visible)


# ------------------------------------------------------------------------------------------------ sort_gaussian
def sort_gaussian(uv: Tensor, depth: Tensor, W: int, H: int, radius: Tensor, tiles: Tensor) -> Tuple[Tensor, Tensor]:
    """``(idx_sorted[I] int32, tile_range[ntiles,2] int32)`` (gs/sort_gaussian.py:8-54; K11-K14).
    One host sync (the size of ``idx_sorted``) instead of the reference's two ``.item()`` calls."""
    L.need_cuda(uv, depth, radius, tiles)
    uv_c, depth_c = L.f32c(uv.detach()), L.f32c(depth.detach())
    radius_c = radius.to(torch.int32).contiguous()
    tiles_c = tiles.to(torch.int32).contiguous()
    P = uv_c.shape[0]
    dev = uv.device
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    tile_range = torch.empty(ntiles, 2, dtype=torch.int32, device=dev)
    if P == 0:
        tile_range.zero_()
        return torch.empty(0, dtype=torch.int32, device=dev), tile_range
    offsets = torch.empty(P, dtype=torch.int32, device=dev)
    ws1 = torch.empty(L.query("spv_sort_scan_workspace_bytes", P), dtype=torch.uint8, device=dev)
    L.call("spv_sort_scan", P, L.ptr(tiles_c), L.ptr(offsets), L.ptr(ws1), ws1.numel(), L.stream())
    I = int(offsets[-1].item())
    idx_sorted = torch.empty(max(I, 0), dtype=torch.int32, device=dev)
    ws2 = torch.empty(L.query("spv_sort_workspace_bytes", P, I), dtype=torch.uint8, device=dev)
    L.call("spv_sort_gaussian", P, I, L.ptr(uv_c), L.ptr(depth_c), L.ptr(radius_c), L.ptr(offsets), int(W), int(H),
           L.ptr(idx_sorted), L.ptr(tile_range), L.ptr(ws2), ws2.numel(), L.stream())
    return idx_sorted, tile_range


# ------------------------------------------------------------------------------------------------ compute_sh
class _ComputeSH(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shs, degree, view_dirs, visible, free):
        L.need_cuda(shs, view_dirs, visible)
        shs_c, dirs_c = L.f32c(shs), L.f32c(view_dirs)
        P = shs_c.shape[0]
        if shs_c.shape[1] < _NUM_SH[degree]:
            raise ValueError(f"shs has {shs_c.shape[1]} bases, degree {degree} needs {_NUM_SH[degree]}")
        vis = _vis_u8(visible, P, shs.device)
        colors = torch.empty(P, 3, dtype=torch.float32, device=shs.device)
        clamped = None if free else torch.empty(P, 3, dtype=torch.uint8, device=shs.device)
        L.call("spv_compute_sh_forward", P, L.ptr(shs_c), int(degree), L.ptr(dirs_c), L.ptr(vis), int(free),
               L.ptr(colors), L.ptr(clamped), L.stream())
        ctx.degree, ctx.free = int(degree), bool(free)
        ctx.save_for_backward(shs_c, dirs_c, vis, clamped if clamped is not None else vis)
        return colors

    @staticmethod
    def backward(ctx, dL_dcolor):
        shs, dirs, vis, clamped = ctx.saved_tensors
        P, S = shs.shape[0], shs.shape[1]
        g_shs = torch.empty(P, S, 3, dtype=torch.float32, device=shs.device)
        g_dirs = torch.empty(P, 3, dtype=torch.float32, device=shs.device)
        L.call("spv_compute_sh_backward", P, L.ptr(shs), ctx.degree, L.ptr(dirs), L.ptr(vis),
               None if ctx.free else L.ptr(clamped), L.ptr(L.f32c(dL_dcolor)), S, L.ptr(g_shs), L.ptr(g_dirs),
               L.stream())
        return g_shs, None, g_dirs, None, None


def compute_sh(shs: Tensor, degree: int, view_dirs: Tensor, visible: Optional[Tensor] = None) -> Tensor:
    """SH -> RGB with +0.5 and clamp at 0 (gs/compute_sh.py:8-36; K7/K8)."""
    if visible is None:
        visible = torch.ones_like(shs[:, 0, 0], dtype=torch.bool)
    return _ComputeSH.apply(shs, degree, view_dirs, visible, False)


def compute_sh_free(shs: Tensor, degree: int, view_dirs: Tensor, visible: Optional[Tensor] = None) -> Tensor:
    """SH evaluation without offset/clamp (gs/compute_sh_free.py:8-36; K9/K10)."""
    if visible is None:
        visible = torch.ones_like(shs[:, 0, 0], dtype=torch.bool)
    return _ComputeSH.apply(shs, degree, view_dirs, visible, True)


# ------------------------------------------------------------------------------------------------ alpha blending
def _blend_forward(uv, conic, opacity, feature, opacity_bias, idx_sorted, tile_range, bg, W, H, K, trunc):
    L.need_cuda(uv, conic, opacity, feature, idx_sorted, tile_range, opacity_bias)
    uv_c, conic_c, op_c, feat_c = L.f32c(uv), L.f32c(conic), L.f32c(opacity), L.f32c(feature)
    bias_c = L.f32c(opacity_bias)
    idx_c = idx_sorted.to(torch.int32).contiguous()
    tr_c = tile_range.to(torch.int32).contiguous()
    P, C = feat_c.shape[0], feat_c.shape[1]
    dev = feat_c.device
    rendered = torch.empty(C, H, W, dtype=torch.float32, device=dev)
    final_T = torch.empty(H, W, dtype=torch.float32, device=dev)
    ncontrib = torch.empty(H, W, dtype=torch.int32, device=dev)
    gs_idx = torch.empty(H, W, K, dtype=torch.int32, device=dev) if K > 0 else None
    L.call("spv_alpha_blend_forward", P, C, int(W), int(H), int(K), int(bool(trunc)), L.ptr(uv_c), L.ptr(conic_c),
           L.ptr(op_c), L.ptr(feat_c), L.ptr(bias_c), L.ptr(idx_c), L.ptr(tr_c), float(bg), L.ptr(rendered),
           L.ptr(final_T), L.ptr(ncontrib), L.ptr(gs_idx), L.stream())
    saved = (uv_c, conic_c, op_c, feat_c, idx_c, tr_c, final_T, ncontrib) + ((bias_c,) if bias_c is not None else ())
    return rendered, ncontrib, gs_idx, saved


def _blend_backward(saved, has_bias, bg, W, H, dL_drendered):
    uv, conic, opacity, feature, idx_sorted, tile_range, final_T, ncontrib = saved[:8]
    bias = saved[8] if has_bias else None
    P, C = feature.shape
    dev = feature.device
    g_uv = torch.empty(P, 2, dtype=torch.float32, device=dev)
    g_abs = torch.empty(P, 2, dtype=torch.float32, device=dev)
    g_conic = torch.empty(P, 3, dtype=torch.float32, device=dev)
    g_op = torch.empty(P, 1, dtype=torch.float32, device=dev)
    g_feat = torch.empty(P, C, dtype=torch.float32, device=dev)
    g_bias = torch.empty(P, 1, dtype=torch.float32, device=dev) if has_bias else None
    ws = torch.empty(L.query("spv_alpha_blend_backward_workspace_bytes", P, C), dtype=torch.uint8, device=dev)
    L.call("spv_alpha_blend_backward", P, C, int(W), int(H), L.ptr(uv), L.ptr(conic), L.ptr(opacity), L.ptr(feature),
           L.ptr(bias), L.ptr(idx_sorted), L.ptr(tile_range), float(bg), L.ptr(final_T), L.ptr(ncontrib),
           L.ptr(L.f32c(dL_drendered)), L.ptr(g_uv), L.ptr(g_abs), L.ptr(g_conic), L.ptr(g_op), L.ptr(g_feat),
           L.ptr(g_bias), L.ptr(ws), ws.numel(), L.stream())
    return g_uv, g_conic, g_op, g_feat, g_abs, g_bias


def _ndc_grads(ctx_ndc, ctx_abs_ndc, g_uv, g_abs, W, H):
    """gs/alpha_blending.py:112-120: the dummy ndc / abs_ndc inputs receive dL_duv * [W/2, H/2]."""
    scale = None
    g_ndc = g_abs_ndc = None
    if ctx_ndc is not None or ctx_abs_ndc is not None:
        scale = torch.tensor([0.5 * W, 0.5 * H], dtype=g_uv.dtype, device=g_uv.device)
    if ctx_ndc is not None:
        g_ndc = g_uv * scale[None, :]
    if ctx_abs_ndc is not None:
        g_abs_ndc = g_abs * scale[None, :]
    return g_ndc, g_abs_ndc


class _AlphaBlending(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc, abs_ndc):
        rendered, _, _, saved = _blend_forward(uv, conic, opacity, feature, None, idx_sorted, tile_range, bg, W, H, 0, False)
        ctx.meta = (float(bg), int(W), int(H), ndc is not None, abs_ndc is not None)
        ctx.save_for_backward(*saved)
        return rendered

    @staticmethod
    def backward(ctx, dL_drendered):
        bg, W, H, has_ndc, has_abs = ctx.meta
        g_uv, g_conic, g_op, g_feat, g_abs, _ = _blend_backward(ctx.saved_tensors, False, bg, W, H, dL_drendered)
        g_ndc, g_abs_ndc = _ndc_grads(True if has_ndc else None, True if has_abs else None, g_uv, g_abs, W, H)
        return g_uv, g_conic, g_op, g_feat, None, None, None, None, None, g_ndc, g_abs_ndc


def alpha_blending(uv: Tensor, conic: Tensor, opacity: Tensor, feature: Tensor, idx_sorted: Tensor,
                   title_bins: Tensor, bg: float, W: int, H: int, ndc: Optional[Tensor] = None,
                   abs_ndc: Optional[Tensor] = None) -> Tensor:
    """Tile-based front-to-back alpha blending -> ``feature_map[C,H,W]`` (gs/alpha_blending.py:7-57; K15/K16)."""
    return _AlphaBlending.apply(uv, conic, opacity, feature, idx_sorted, title_bins, bg, W, H, ndc, abs_ndc)


class _AlphaBlendingEnhanced(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc, abs_ndc, K, enable_truncation):
        rendered, ncontrib, gs_idx, saved = _blend_forward(uv, conic, opacity, feature, None, idx_sorted, tile_range,
                                                           bg, W, H, int(K), enable_truncation)
        ctx.meta = (float(bg), int(W), int(H), ndc is not None, abs_ndc is not None)
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(ncontrib, gs_idx)
        return rendered, ncontrib, gs_idx

    @staticmethod
    def backward(ctx, dL_drendered, _n, _g):
        bg, W, H, has_ndc, has_abs = ctx.meta
        g_uv, g_conic, g_op, g_feat, g_abs, _ = _blend_backward(ctx.saved_tensors, False, bg, W, H, dL_drendered)
        g_ndc, g_abs_ndc = _ndc_grads(True if has_ndc else None, True if has_abs else None, g_uv, g_abs, W, H)
        return g_uv, g_conic, g_op, g_feat, None, None, None, None, None, g_ndc, g_abs_ndc, None, None


def alpha_blending_enhanced(uv: Tensor, conic: Tensor, opacity: Tensor, feature: Tensor, idx_sorted: Tensor,
                            title_bins: Tensor, bg: float, W: int, H: int, ndc: Optional[Tensor] = None,
                            abs_ndc: Optional[Tensor] = None, K: int = 10,
                            enable_truncation: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """Blending + first-K contributing ids: ``(feature_map[C,H,W], ncontrib[H,W] int32, gs_idx[H,W,K] int32)``
    (gs/alpha_blending_enhanced.py:7-66; K17/K18)."""
    return _AlphaBlendingEnhanced.apply(uv, conic, opacity, feature, idx_sorted, title_bins, bg, W, H, ndc, abs_ndc, K,
                                        enable_truncation)


class _AlphaBlendingWithBias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, opacity_bias, idx_sorted, tile_range, bg, W, H, ndc, abs_ndc):
        rendered, _, _, saved = _blend_forward(uv, conic, opacity, feature, opacity_bias, idx_sorted, tile_range, bg, W, H,
                                               0, False)
        ctx.meta = (float(bg), int(W), int(H), ndc is not None, abs_ndc is not None)
        ctx.save_for_backward(*saved)
        return rendered

    @staticmethod
    def backward(ctx, dL_drendered):
        bg, W, H, has_ndc, has_abs = ctx.meta
        g_uv, g_conic, g_op, g_feat, g_abs, g_bias = _blend_backward(ctx.saved_tensors, True, bg, W, H, dL_drendered)
        g_ndc, g_abs_ndc = _ndc_grads(True if has_ndc else None, True if has_abs else None, g_uv, g_abs, W, H)
        return g_uv, g_conic, g_op, g_feat, g_bias, None, None, None, None, None, g_ndc, g_abs_ndc


def alpha_blending_with_bias(uv: Tensor, conic: Tensor, opacity: Tensor, feature: Tensor, opacity_bias: Tensor,
                             idx_sorted: Tensor, title_bins: Tensor, bg: float, W: int, H: int,
                             ndc: Optional[Tensor] = None, abs_ndc: Optional[Tensor] = None) -> Tensor:
    """Blending with a per-Gaussian additive alpha bias (gs/alpha_blending_with_bias.py:7-60; K19/K20)."""
    return _AlphaBlendingWithBias.apply(uv, conic, opacity, feature, opacity_bias, idx_sorted, title_bins, bg, W, H, ndc,
                                        abs_ndc)


# ------------------------------------------------------------------------------------------------ pipeline
def rasterization(xyz: Tensor, scale: Tensor, rotate: Tensor, opacity: Tensor, feature: Tensor, intr: Tensor,
                  extr: Tensor, W: int, H: int, bg: float, ndc: Optional[Tensor] = None) -> Tensor:
    """Vanilla 3DGS pipeline -> ``feature_map[C,H,W]`` (gs/__init__.py:28-100)."""
    uv, depth = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles_touched = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    idx_sorted, tile_range = sort_gaussian(uv, depth, W, H, radius, tiles_touched)
    return alpha_blending(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc)
