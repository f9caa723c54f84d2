"""Shared test utilities: scene -> numpy inputs, oracle pipelines, tolerances.

Tolerances (BASELINE.json north_star): pixels <= 1e-4 max-abs, gradients <= 1e-3 relative, tile indices bit-exact.
Discrete decisions (alpha >= 1/255, T >= 1e-4, power <= 0) flip under a 1e-7 relative change of exp(); pixels whose
decisions sit within FRAG_EPS of a threshold are reported by the oracle ("fragile") and excluded from the strict
pixel / index comparisons -- their count is bounded instead.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from splatter_a_video_b200 import synth  # noqa: E402

PIX_TOL = 1e-4
GRAD_RTOL = 1e-3
FRAG_EPS = 2e-5
GOLDEN = os.path.join(ROOT, "tests", "golden")


def scene_np(P, W, H, seed=1234, frame=0, frames=10, **kw):
    sc = synth.make_scene(P, frames, W, H, seed=seed, **kw)
    d = dict(P=P, W=W, H=H, xyz=sc.frame_position(frame).numpy(), scaling=sc.scaling.numpy(),
             rotation=sc.rotation.numpy(), opacity=sc.opacity.numpy(), shs=sc.shs.numpy(),
             attrs=sc.attr_features(min(frame + 1, frames - 1)).numpy(), extr=sc.extr.numpy(), intr=sc.intr.numpy())
    return d


def oracle_ortho(s, K=20, frag_eps=FRAG_EPS):
    """C-oracle forward of the trainer's ortho frame (dptr_ortho_enhanced.py:270-376)."""
    P, W, H = s["P"], s["W"], s["H"]
    dirs = np.zeros((P, 3), np.float32); dirs[:, 2] = 1
    rgb, clamped = O.compute_sh(s["shs"], 3, dirs)
    uv, depth = O.project_point_ortho(s["xyz"], s["extr"], W, H, nearest=0.01)
    vis = depth.reshape(-1) != 0
    cov3d = O.compute_cov3d(s["scaling"], s["rotation"], vis)
    conic, radius, tiles = O.ewa_project_ortho(cov3d, s["extr"], uv, W, H, vis)
    idx, tr = O.sort_gaussian(uv, depth, W, H, radius, tiles)
    out = dict(dirs=dirs, rgb=rgb, clamped=clamped, uv=uv, depth=depth, vis=vis, cov3d=cov3d, conic=conic,
               radius=radius, tiles=tiles, idx_sorted=idx, tile_range=tr)
    out["blend_rgb"] = O.alpha_blending_forward(uv, conic, s["opacity"], rgb, idx, tr, 0.0, W, H, K=K, frag_eps=frag_eps)
    return out


def oracle_persp(s, nearest=0.2, extent=1.3):
    P, W, H = s["P"], s["W"], s["H"]
    uv, depth = O.project_point(s["xyz"], s["intr"], s["extr"], W, H, nearest, extent)
    vis = depth.reshape(-1) != 0
    cov3d = O.compute_cov3d(s["scaling"], s["rotation"], vis)
    conic, radius, tiles = O.ewa_project(s["xyz"], cov3d, s["intr"], s["extr"], uv, W, H, vis)
    idx, tr = O.sort_gaussian(uv, depth, W, H, radius, tiles)
    return dict(uv=uv, depth=depth, vis=vis, cov3d=cov3d, conic=conic, radius=radius, tiles=tiles, idx_sorted=idx,
                tile_range=tr)


def t(a, device, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return x if dtype is None else x.to(dtype)


def n(x):
    return x.detach().cpu().numpy()


def grad_err(a, b):
    """(normwise relative error, max elementwise violation of |a-b| <= rtol|b| + 1e-4 max|b|)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    nb = np.linalg.norm(b)
    rel = np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
    bound = GRAD_RTOL * np.abs(b) + 1e-4 * (np.abs(b).max() if b.size else 0.0) + 1e-30
    worst = float((np.abs(a - b) / bound).max()) if b.size else 0.0
    return rel, worst


def assert_grad_close(a, b, name, norm_tol=1e-4):
    rel, worst = grad_err(a, b)
    assert rel <= norm_tol, f"{name}: normwise rel err {rel:.3e} > {norm_tol}"
    assert worst <= 1.0, f"{name}: elementwise tolerance exceeded x{worst:.2f}"


def assert_pixels_close(img, ref, fragile, name, tol=PIX_TOL, frag_tol=0.05, max_frag_frac=2e-3):
    """img/ref [C,H,W] (or [H,W]); fragile [H,W] bool."""
    img = np.asarray(img); ref = np.asarray(ref)
    d = np.abs(img - ref)
    if d.ndim == 3:
        d = d.max(0)
    frac = fragile.mean()
    assert frac <= max_frag_frac, f"{name}: fragile pixel fraction {frac:.2e}"
    bad = d[~fragile].max() if (~fragile).any() else 0.0
    assert bad <= tol, f"{name}: max abs pixel error {bad:.3e} on non-fragile pixels"
    if fragile.any():
        assert d[fragile].max() <= frag_tol, f"{name}: fragile pixel error {d[fragile].max():.3e}"


def ref_module():
    """The UNMODIFIED reference rasterizer compiled for sm_100a (oracle/_ref/_C.so), or None."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", "_C.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_C", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
