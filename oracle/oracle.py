"""ctypes/numpy front-end of the C oracle (oracle/spv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(splatter_a_video_b200/) never imports this module.

Every function mirrors one stage of the reference ``dptr.gs`` API
(/root/reference/src/submodules/dptr/dptr/gs/*.py) with numpy arrays in and out.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspv_oracle.so")
_lib = None

f32p = ctypes.POINTER(ctypes.c_float)
i32p = ctypes.POINTER(ctypes.c_int)
i64p = ctypes.POINTER(ctypes.c_int64)
u8p = ctypes.POINTER(ctypes.c_ubyte)


def build(force: bool = False) -> str:
    """Compile libspv_oracle.so with the committed Makefile (gcc, no fast-math, no FMA)."""
    src = os.path.join(_HERE, "spv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libspv_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_count_intersections.restype = ctypes.c_int64
    return _lib


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _b(a):
    return np.ascontiguousarray(np.asarray(a).astype(np.uint8).reshape(-1))


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _extr12(extr):
    """[3,4] or [4,4] -> 12 floats with row stride 4 (the kernels index extr[0..11])."""
    e = _f(extr).reshape(-1)
    return np.ascontiguousarray(e[:12])


# --------------------------------------------------------------------------- stages
def project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3):
    xyz = _f(xyz, (-1, 3)); P = xyz.shape[0]
    intr = _f(intr); e = _extr12(extr)
    uv = np.empty((P, 2), np.float32); depth = np.empty((P, 1), np.float32)
    lib().orc_project_point_fwd(P, _p(xyz, f32p), _p(intr, f32p), _p(e, f32p), int(W), int(H),
                                ctypes.c_float(nearest), ctypes.c_float(extent), _p(uv, f32p), _p(depth, f32p))
    return uv, depth


def project_point_backward(xyz, intr, extr, depth, dL_duv, dL_ddepth, need_intr=False, need_extr=False):
    xyz = _f(xyz, (-1, 3)); P = xyz.shape[0]
    intr = _f(intr); e = _extr12(extr)
    g = np.empty((P, 3), np.float32)
    gi = np.zeros(4, np.float32) if need_intr else None
    ge = np.zeros(12, np.float32) if need_extr else None
    lib().orc_project_point_bwd(P, _p(xyz, f32p), _p(intr, f32p), _p(e, f32p), _p(_f(depth), f32p),
                                _p(_f(dL_duv), f32p), _p(_f(dL_ddepth), f32p), _p(g, f32p), _p(gi, f32p), _p(ge, f32p))
    return g, gi, (None if ge is None else ge.reshape(3, 4))


def project_point_ortho(xyz, extr, W, H, nearest=0.2, extent=1.3):
    xyz = _f(xyz, (-1, 3)); P = xyz.shape[0]
    e = _extr12(extr)
    uv = np.empty((P, 2), np.float32); depth = np.empty((P, 1), np.float32)
    lib().orc_project_point_ortho_fwd(P, _p(xyz, f32p), _p(e, f32p), int(W), int(H),
                                      ctypes.c_float(nearest), ctypes.c_float(extent), _p(uv, f32p), _p(depth, f32p))
    return uv, depth


def compute_cov3d(scales, uquats, visible=None):
    scales = _f(scales, (-1, 3)); P = scales.shape[0]
    vis = _b(np.ones(P, bool) if visible is None else visible)
    out = np.empty((P, 6), np.float32)
    lib().orc_compute_cov3d_fwd(P, _p(scales, f32p), _p(_f(uquats), f32p), _p(vis, u8p), _p(out, f32p))
    return out


def compute_cov3d_backward(scales, uquats, visible, dL_dcov3d):
    scales = _f(scales, (-1, 3)); P = scales.shape[0]
    vis = _b(np.ones(P, bool) if visible is None else visible)
    gs = np.empty((P, 3), np.float32); gq = np.empty((P, 4), np.float32)
    lib().orc_compute_cov3d_bwd(P, _p(scales, f32p), _p(_f(uquats), f32p), _p(vis, u8p), _p(_f(dL_dcov3d), f32p),
                                _p(gs, f32p), _p(gq, f32p))
    return gs, gq


def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None):
    xyz = _f(xyz, (-1, 3)); P = xyz.shape[0]
    vis = _b(np.ones(P, bool) if visible is None else visible)
    conic = np.empty((P, 3), np.float32); radius = np.empty(P, np.int32); tiles = np.empty(P, np.int32)
    lib().orc_ewa_project_fwd(P, _p(xyz, f32p), _p(_f(cov3d), f32p), _p(_f(intr), f32p), _p(_extr12(extr), f32p),
                              _p(_f(uv), f32p), int(W), int(H), _p(vis, u8p), _p(conic, f32p), _p(radius, i32p),
                              _p(tiles, i32p))
    return conic, radius, tiles


def ewa_project_backward(xyz, cov3d, intr, extr, radius, dL_dconic, need_intr=False, need_extr=False):
    xyz = _f(xyz, (-1, 3)); P = xyz.shape[0]
    gx = np.empty((P, 3), np.float32); gc = np.empty((P, 6), np.float32)
    gi = np.zeros(4, np.float32) if need_intr else None
    ge = np.zeros(12, np.float32) if need_extr else None
    lib().orc_ewa_project_bwd(P, _p(xyz, f32p), _p(_f(cov3d), f32p), _p(_f(intr), f32p), _p(_extr12(extr), f32p),
                              _p(_i(radius), i32p), _p(_f(dL_dconic), f32p), _p(gx, f32p), _p(gc, f32p),
                              _p(gi, f32p), _p(ge, f32p))
    return gx, gc, gi, (None if ge is None else ge.reshape(3, 4))


def ewa_project_ortho(cov3d, extr, uv, W, H, visible):
    cov3d = _f(cov3d, (-1, 6)); P = cov3d.shape[0]
    vis = _b(visible)
    conic = np.empty((P, 3), np.float32); radius = np.empty(P, np.int32); tiles = np.empty(P, np.int32)
    lib().orc_ewa_project_ortho_fwd(P, _p(cov3d, f32p), _p(_extr12(extr), f32p), _p(_f(uv), f32p), int(W), int(H),
                                    _p(vis, u8p), _p(conic, f32p), _p(radius, i32p), _p(tiles, i32p))
    return conic, radius, tiles


def compute_sh(shs, degree, view_dirs, visible=None, free=False):
    shs = _f(shs); P = shs.shape[0]
    vis = _b(np.ones(P, bool) if visible is None else visible)
    colors = np.empty((P, 3), np.float32)
    clamped = None if free else np.empty((P, 3), np.uint8)
    lib().orc_compute_sh_fwd(P, _p(shs, f32p), int(degree), _p(_f(view_dirs), f32p), _p(vis, u8p), int(bool(free)),
                             _p(colors, f32p), _p(clamped, u8p))
    return colors, clamped


def compute_sh_backward(shs, degree, view_dirs, visible, clamped, dL_dcolors):
    shs = _f(shs); P, S = shs.shape[0], shs.shape[1]
    vis = _b(np.ones(P, bool) if visible is None else visible)
    cl = None if clamped is None else _b(clamped)
    gsh = np.empty((P, S, 3), np.float32); gd = np.empty((P, 3), np.float32)
    lib().orc_compute_sh_bwd(P, _p(shs, f32p), int(degree), _p(_f(view_dirs), f32p), _p(vis, u8p), _p(cl, u8p),
                             _p(_f(dL_dcolors), f32p), int(S), _p(gsh, f32p), _p(gd, f32p))
    return gsh, gd


def sort_gaussian(uv, depth, W, H, radius, tiles, return_keys=False):
    uv = _f(uv, (-1, 2)); P = uv.shape[0]
    radius = _i(radius).reshape(-1); tiles = _i(tiles).reshape(-1)
    I = int(lib().orc_count_intersections(P, _p(tiles, i32p)))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    idx_sorted = np.zeros(max(I, 0), np.int32)
    keys = np.zeros(max(I, 0), np.int64) if return_keys else None
    tile_range = np.zeros((gx * gy, 2), np.int32)
    lib().orc_sort_gaussian(P, _p(uv, f32p), _p(_f(depth), f32p), int(W), int(H), _p(radius, i32p), _p(tiles, i32p),
                            ctypes.c_int64(I), _p(idx_sorted, i32p), _p(keys, i64p), _p(tile_range, i32p))
    if return_keys:
        return idx_sorted, tile_range, keys
    return idx_sorted, tile_range


def alpha_blending_forward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H,
                           K: int = 0, enable_truncation: bool = False, opacity_bias=None,
                           frag_eps: float = 0.0):
    """-> dict(rendered[C,H,W], final_T[H,W], ncontrib[H,W], gs_idx[H,W,K]|None, fragile[H,W])."""
    feature = _f(feature); P, C = feature.shape
    rendered = np.empty((C, H, W), np.float32); final_T = np.empty((H, W), np.float32)
    ncontrib = np.empty((H, W), np.int32)
    gs_idx = np.empty((H, W, K), np.int32) if K > 0 else None
    fragile = np.zeros((H, W), np.uint8)
    ob = None if opacity_bias is None else _f(opacity_bias)
    lib().orc_alpha_blend_fwd(P, C, int(W), int(H), int(K), int(bool(enable_truncation)), _p(_f(uv), f32p),
                              _p(_f(conic), f32p), _p(_f(opacity), f32p), _p(feature, f32p), _p(ob, f32p),
                              _p(_i(idx_sorted), i32p), _p(_i(tile_range), i32p), ctypes.c_float(bg),
                              _p(rendered, f32p), _p(final_T, f32p), _p(ncontrib, i32p), _p(gs_idx, i32p),
                              _p(fragile, u8p), ctypes.c_float(frag_eps))
    return dict(rendered=rendered, final_T=final_T, ncontrib=ncontrib, gs_idx=gs_idx, fragile=fragile.astype(bool))


def alpha_blending_backward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, final_T, ncontrib,
                            dL_drendered, opacity_bias=None):
    """-> dict(dL_duv, dL_dconic, dL_dopacity, dL_dfeature, dL_dabs_uv[, dL_dopacity_bias])."""
    feature = _f(feature); P, C = feature.shape
    duv = np.empty((P, 2), np.float32); dabs = np.empty((P, 2), np.float32); dconic = np.empty((P, 3), np.float32)
    dop = np.empty((P, 1), np.float32); dfeat = np.empty((P, C), np.float32)
    ob = None if opacity_bias is None else _f(opacity_bias)
    dob = None if opacity_bias is None else np.empty((P, 1), np.float32)
    lib().orc_alpha_blend_bwd(P, C, int(W), int(H), _p(_f(uv), f32p), _p(_f(conic), f32p), _p(_f(opacity), f32p),
                              _p(feature, f32p), _p(ob, f32p), _p(_i(idx_sorted), i32p), _p(_i(tile_range), i32p),
                              ctypes.c_float(bg), _p(_f(final_T), f32p), _p(_i(ncontrib), i32p),
                              _p(_f(dL_drendered), f32p), _p(duv, f32p), _p(dabs, f32p), _p(dconic, f32p),
                              _p(dop, f32p), _p(dfeat, f32p), _p(dob, f32p))
    out = dict(dL_duv=duv, dL_dconic=dconic, dL_dopacity=dop, dL_dfeature=dfeat, dL_dabs_uv=dabs)
    if dob is not None:
        out["dL_dopacity_bias"] = dob
    return out


# --------------------------------------------------------------------------- pipelines
def rasterization(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg, nearest=0.2, extent=1.3):
    """gs/__init__.py:28-100 (perspective pipeline), forward only.  Returns (feature_map, aux dict)."""
    uv, depth = project_point(xyz, intr, extr, W, H, nearest, extent)
    visible = depth.reshape(-1) != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    idx_sorted, tile_range = sort_gaussian(uv, depth, W, H, radius, tiles)
    out = alpha_blending_forward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H)
    aux = dict(uv=uv, depth=depth, cov3d=cov3d, conic=conic, radius=radius, tiles=tiles, idx_sorted=idx_sorted,
               tile_range=tile_range, **out)
    return out["rendered"], aux
