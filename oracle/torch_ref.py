"""Pure-PyTorch (CPU, fp32) differentiable restatement of the rasterizer hot path.

TEST INFRASTRUCTURE ONLY (same rule as oracle/spv_oracle.c).  Two jobs:

1. BASELINE.json ``configs[0]``: "1k Gaussians, 2x64x64 frames, pure-PyTorch projection +
   alpha-blend on CPU (correctness ref, no GPU)" -- this file *is* that path, timed by
   bench.py's ``cpu_baseline`` leg on the host cores.
2. An independent gradient reference: torch autograd through the restated forward validates the
   C oracle's hand-derived backward (tests/test_oracle_cpu.py).

Op order follows the reference (file:line cited per function; paths relative to
/root/reference/src/).  Nothing here is copied from the reference; the torch-only pieces of the
reference (ortho projection / ortho EWA) are re-expressed with the same arithmetic.
"""
from __future__ import annotations

from typing import Tuple

import torch

BLOCK = 16
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def project_point_ortho(xyz, extr, W, H, nearest=0.2, extent=1.3):
    """pointrix/renderer/dptr_ortho_enhanced.py:177-202."""
    R, t = extr[:3, :3], extr[:3, 3]
    cam = xyz @ R.t() + t
    u = (cam[:, 0] + 1.0) * W / 2 - 0.5
    v = (cam[:, 1] + 1.0) * H / 2 - 0.5
    d = torch.nan_to_num(cam[:, 2])
    mask = (d <= nearest) | (u < (1 - extent) * W * 0.5) | (u > (1 + extent) * W * 0.5) \
        | (v < (1 - extent) * H * 0.5) | (v > (1 + extent) * H * 0.5)
    keep = (~mask).to(xyz.dtype)
    uv = torch.stack([u, v], -1) * keep[:, None]
    return uv, (d * keep)[:, None]


def project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3):
    """submodules/dptr/dptr/gs/src/project_point.cu:27-56."""
    e = extr.reshape(-1)[:12].reshape(3, 4)
    cam = xyz @ e[:, :3].t() + e[:, 3]
    inv = 1.0 / (cam[:, 2] + 1e-7)
    u = intr[0] * cam[:, 0] * inv + intr[2] - 0.5
    v = intr[1] * cam[:, 1] * inv + intr[3] - 0.5
    mask = torch.zeros_like(u, dtype=torch.bool)
    if nearest > 0:
        mask |= cam[:, 2] <= nearest
    if extent > 0:
        mask |= (u < (1 - extent) * W * 0.5) | (u > (1 + extent) * W * 0.5) \
            | (v < (1 - extent) * H * 0.5) | (v > (1 + extent) * H * 0.5)
    keep = (~mask).to(xyz.dtype)
    return torch.stack([u, v], -1) * keep[:, None], (cam[:, 2] * keep)[:, None]


def quat_to_rot(q):
    r, x, y, z = q.unbind(-1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, -1).reshape(-1, 3, 3)  # row-major R (same entries as compute_cov3d.cu:24-40)


def compute_cov3d(scales, uquats, visible=None):
    """submodules/dptr/dptr/gs/src/compute_cov3d.cu:42-58: Sigma = R S^2 R^T, upper triangle."""
    R = quat_to_rot(uquats)
    M = R * scales[:, None, :]
    S = M @ M.transpose(1, 2)
    c = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)
    if visible is not None:
        c = c * visible.reshape(-1, 1).to(c.dtype)
    return c


def _sym3(c):
    return torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], -1),
                        torch.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                        torch.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], -2)


def _finish_ewa(c00, c01, c11, uv, W, H, visible, use_division):
    det = c00 * c11 - c01 * c01
    if use_division:   # ortho torch path, dptr_ortho_enhanced.py:56-63
        conic = torch.stack([c11 / det, -c01 / det, c00 / det], -1)
    else:              # ewa_project.cu:74-77
        inv = 1.0 / det
        conic = torch.stack([c11 * inv, -c01 * inv, c00 * inv], -1)
    mid = 0.5 * (c00 + c11)
    root = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(mid + root, mid - root))).detach()
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    u, v = uv[:, 0].detach(), uv[:, 1].detach()
    x0 = ((u - radius) / BLOCK).to(torch.int32).clamp(0, gx)
    y0 = ((v - radius) / BLOCK).to(torch.int32).clamp(0, gy)
    x1 = ((u + radius + BLOCK - 1) / BLOCK).to(torch.int32).clamp(0, gx)
    y1 = ((v + radius + BLOCK - 1) / BLOCK).to(torch.int32).clamp(0, gy)
    tiles = (x1 - x0) * (y1 - y0)
    mask = (tiles != 0) & (det.detach() != 0) & visible.reshape(-1)
    conic = torch.where(mask[:, None], conic, torch.zeros_like(conic))
    return conic, (radius * mask).to(torch.int32), (tiles * mask).to(torch.int32)


def ewa_project_ortho(cov3d, extr, uv, W, H, visible):
    """pointrix/renderer/dptr_ortho_enhanced.py:26-111 with J = [[W/2,0,0],[0,H/2,0]]."""
    J = torch.tensor([[W / 2, 0, 0], [0, H / 2, 0]], dtype=cov3d.dtype)
    T = J @ extr[:3, :3]
    c2 = T @ _sym3(cov3d) @ T.t()
    return _finish_ewa(c2[:, 0, 0] + 0.3, c2[:, 0, 1], c2[:, 1, 1] + 0.3, uv, W, H, visible, True)


def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible):
    """submodules/dptr/dptr/gs/src/ewa_project.cu:34-82."""
    e = extr.reshape(-1)[:12].reshape(3, 4)
    t = xyz @ e[:, :3].t() + e[:, 3]
    fx, fy = intr[0], intr[1]
    z = torch.zeros_like(t[:, 0])
    J = torch.stack([torch.stack([fx / t[:, 2], z, -(fx * t[:, 0]) / (t[:, 2] * t[:, 2])], -1),
                     torch.stack([z, fy / t[:, 2], -(fy * t[:, 1]) / (t[:, 2] * t[:, 2])], -1)], -2)
    T = J @ e[:, :3]
    c2 = T @ _sym3(cov3d) @ T.transpose(1, 2)
    return _finish_ewa(c2[:, 0, 0] + 0.3, c2[:, 0, 1], c2[:, 1, 1] + 0.3, uv, W, H, visible, False)


def compute_sh(shs, degree, dirs, visible=None, free=False):
    """submodules/dptr/dptr/gs/src/compute_sh.cu:43-79 (requires shs.shape[1] == (degree+1)**2)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    r = SH_C0 * shs[:, 0]
    if degree > 0:
        r = r - SH_C1 * y * shs[:, 1] + SH_C1 * z * shs[:, 2] - SH_C1 * x * shs[:, 3]
    if degree > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = r + SH_C2[0] * xy * shs[:, 4] + SH_C2[1] * yz * shs[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * shs[:, 6] \
            + SH_C2[3] * xz * shs[:, 7] + SH_C2[4] * (xx - yy) * shs[:, 8]
    if degree > 2:
        r = r + SH_C3[0] * y * (3 * xx - yy) * shs[:, 9] + SH_C3[1] * xy * z * shs[:, 10] \
            + SH_C3[2] * y * (4 * zz - xx - yy) * shs[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12] \
            + SH_C3[4] * x * (4 * zz - xx - yy) * shs[:, 13] + SH_C3[5] * z * (xx - yy) * shs[:, 14] \
            + SH_C3[6] * x * (xx - 3 * yy) * shs[:, 15]
    if not free:
        r = torch.clamp_min(r + 0.5, 0.0)
    if visible is not None:
        r = r * visible.reshape(-1, 1).to(r.dtype)
    return r


def sort_gaussian(uv, depth, W, H, radius, tiles) -> Tuple[torch.Tensor, torch.Tensor]:
    """submodules/dptr/dptr/gs/sort_gaussian.py:41-54 + src/sort_gaussian.cu:15-69."""
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    r = radius.to(torch.float32)
    x0 = ((uv[:, 0] - r) / BLOCK).to(torch.int64).clamp(0, gx)
    y0 = ((uv[:, 1] - r) / BLOCK).to(torch.int64).clamp(0, gy)
    x1 = ((uv[:, 0] + r + BLOCK - 1) / BLOCK).to(torch.int64).clamp(0, gx)
    y1 = ((uv[:, 1] + r + BLOCK - 1) / BLOCK).to(torch.int64).clamp(0, gy)
    dbits = depth.reshape(-1).contiguous().view(torch.int32).to(torch.int64)
    keys, ids = [], []
    for i in torch.nonzero(radius > 0).reshape(-1).tolist():
        ys = torch.arange(int(y0[i]), int(y1[i]))
        xs = torch.arange(int(x0[i]), int(x1[i]))
        tid = (ys[:, None] * gx + xs[None, :]).reshape(-1)
        keys.append((tid << 32) | dbits[i])
        ids.append(torch.full_like(tid, i))
    tile_range = torch.zeros(gx * gy, 2, dtype=torch.int32)
    if not keys:
        return torch.zeros(0, dtype=torch.int32), tile_range
    keys = torch.cat(keys); ids = torch.cat(ids)
    ks, order = torch.sort(keys, stable=True)
    idx_sorted = ids[order].to(torch.int32)
    t = (ks >> 32)
    n = t.numel()
    starts = torch.nonzero(torch.cat([torch.tensor([True]), t[1:] != t[:-1]])).reshape(-1)
    ends = torch.cat([starts[1:], torch.tensor([n])])
    tile_range[t[starts], 0] = starts.to(torch.int32)
    tile_range[t[starts], 1] = ends.to(torch.int32)
    return idx_sorted, tile_range


def alpha_blending(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, K=0, opacity_bias=None):
    """submodules/dptr/dptr/gs/src/alpha_blending_enhanced.cu:57-133, vectorised over the 256 pixels of a
    tile, sequential over its Gaussians.  Differentiable w.r.t. uv, conic, opacity, feature(, bias).
    Returns (rendered[C,H,W], final_T[H,W], ncontrib[H,W], gs_idx[H,W,K] or None)."""
    C = feature.shape[1]
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    out = torch.zeros(C, gy * BLOCK, gx * BLOCK, dtype=feature.dtype)
    Tout = torch.ones(gy * BLOCK, gx * BLOCK, dtype=feature.dtype)
    nc = torch.zeros(gy * BLOCK, gx * BLOCK, dtype=torch.int32)
    gsi = torch.full((gy * BLOCK, gx * BLOCK, max(K, 1)), -1, dtype=torch.int32)
    ly, lx = torch.meshgrid(torch.arange(BLOCK), torch.arange(BLOCK), indexing="ij")
    for tile in range(gx * gy):
        tx, ty = tile % gx, tile // gx
        px = (tx * BLOCK + lx).reshape(-1).to(feature.dtype)
        py = (ty * BLOCK + ly).reshape(-1).to(feature.dtype)
        r0, r1 = int(tile_range[tile, 0]), int(tile_range[tile, 1])
        T = torch.ones(BLOCK * BLOCK, dtype=feature.dtype)
        F = torch.zeros(BLOCK * BLOCK, C, dtype=feature.dtype)
        done = torch.zeros(BLOCK * BLOCK, dtype=torch.bool)
        last = torch.zeros(BLOCK * BLOCK, dtype=torch.int32)
        layer = torch.zeros(BLOCK * BLOCK, dtype=torch.int64)
        ids = torch.full((BLOCK * BLOCK, max(K, 1)), -1, dtype=torch.int32)
        for n, k in enumerate(range(r0, r1)):
            g = int(idx_sorted[k])
            vx, vy = uv[g, 0] - px, uv[g, 1] - py
            power = -0.5 * (conic[g, 0] * vx * vx + conic[g, 2] * vy * vy) - conic[g, 1] * vx * vy
            a = opacity[g, 0] * torch.exp(power)
            if opacity_bias is not None:
                a = a + opacity_bias[g, 0]
            # min(0.99, .) with the reference's straight-through gradient (the backward kernel ignores the clamp)
            alpha = a + (torch.clamp(a, max=0.99) - a).detach()
            valid = (~done) & (power <= 0) & (alpha.detach() >= 1.0 / 255.0)
            next_T = T * (1 - alpha)
            stop = valid & (next_T.detach() < 0.0001)
            done = done | stop
            app = valid & ~stop
            F = F + torch.where(app[:, None], feature[g][None, :] * (alpha * T)[:, None], torch.zeros(()))
            T = torch.where(app, next_T, T)
            last = torch.where(app, torch.tensor(n + 1, dtype=torch.int32), last)
            if K > 0:
                rec = app & (layer < K)
                rows = torch.nonzero(rec).reshape(-1)
                ids[rows, layer[rows]] = g
                layer = layer + rec.to(torch.int64)
            if bool(done.all()):
                break
        ys = slice(ty * BLOCK, (ty + 1) * BLOCK); xs = slice(tx * BLOCK, (tx + 1) * BLOCK)
        out[:, ys, xs] = (F + (T * bg)[:, None]).t().reshape(C, BLOCK, BLOCK)
        Tout[ys, xs] = T.detach().reshape(BLOCK, BLOCK)
        nc[ys, xs] = last.reshape(BLOCK, BLOCK)
        gsi[ys, xs] = ids.reshape(BLOCK, BLOCK, -1)
    return out[:, :H, :W], Tout[:H, :W], nc[:H, :W], (gsi[:H, :W, :K] if K > 0 else None)


def render_ortho_frame(position, scaling, rotation, opacity, shs, attrs, extr, W, H, K=20):
    """The trainer's per-frame forward (pointrix/renderer/dptr_ortho_enhanced.py:270-376): SH(deg 3, dir z) ->
    ortho projection (nearest 0.01) -> cov3d -> ortho EWA -> sort -> RGB(K) / depth(bg 1) / attributes(bg 0,
    opacity detached).  Returns dict of images + intermediates."""
    dirs = torch.zeros_like(position); dirs[:, 2] = 1.0
    rgb = compute_sh(shs, 3, dirs)
    uv, depth = project_point_ortho(position, extr, W, H, nearest=0.01)
    visible = depth != 0
    cov3d = compute_cov3d(scaling, rotation, visible)
    conic, radius, tiles = ewa_project_ortho(cov3d, extr, uv, W, H, visible.reshape(-1))
    idx_sorted, tile_range = sort_gaussian(uv.detach(), depth.detach(), W, H, radius, tiles)
    img, fT, nc, gs_idx = alpha_blending(uv, conic, opacity, rgb, idx_sorted, tile_range, 0.0, W, H, K=K)
    dimg, _, _, _ = alpha_blending(uv, conic, opacity, depth, idx_sorted, tile_range, 1.0, W, H)
    out = dict(rgb=img, depth=dimg, gs_idx=gs_idx, ncontrib=nc, final_T=fT, uv=uv, depth_pts=depth, conic=conic,
               radius=radius, tiles=tiles, idx_sorted=idx_sorted, tile_range=tile_range, colors=rgb, cov3d=cov3d)
    if attrs is not None:
        aimg, _, _, _ = alpha_blending(uv, conic, opacity.detach(), attrs, idx_sorted, tile_range, 0.0, W, H)
        out["attrs"] = aimg
    return out
