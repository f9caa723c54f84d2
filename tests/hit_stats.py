"""Applied (pixel, Gaussian) pairs and active (pixel-block, entry) pairs per warp footprint for one frame of config A, counted on
the CPU from the oracle's tile lists (numpy; ~10 s).  Used for the instruction budgets in DESIGN.md sections 3 and 8.
    python tests/hit_stats.py      (lives under tests/ because it uses the oracle, which only test infrastructure may import)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from splatter_a_video_b200 import synth
from oracle import oracle as O
sc = synth.make_config("cfg2_davis480p")
P, W, H = sc.P, sc.W, sc.H
xyz = sc.frame_position(0).numpy()
uv, depth = O.project_point_ortho(xyz, sc.extr.numpy(), W, H, nearest=0.01)
vis = depth.reshape(-1) != 0
cov3d = O.compute_cov3d(sc.scaling.numpy(), sc.rotation.numpy(), vis)
conic, radius, tiles = O.ewa_project_ortho(cov3d, sc.extr.numpy(), uv, W, H, vis)
idx, tr = O.sort_gaussian(uv, depth, W, H, radius, tiles)
op = sc.opacity.numpy().reshape(-1)
gx = (W + 15) // 16
t0 = time.time()
tot = dict(entries=0, hits=0, b8x4=0, b8x8=0, b16x8=0, b16x16=0, b16x4=0, b4x4=0)
yy, xx = np.meshgrid(np.arange(16), np.arange(16), indexing='ij')
for t in range(tr.shape[0]):
    a, b = tr[t]
    n = b - a
    if n <= 0: continue
    ids = idx[a:b]
    tx, ty = t % gx, t // gx
    px = (tx * 16 + xx).reshape(-1).astype(np.float32); py = (ty * 16 + yy).reshape(-1).astype(np.float32)
    inside = (px < W) & (py < H)
    dx = uv[ids, 0][:, None] - px[None]; dy = uv[ids, 1][:, None] - py[None]
    c = conic[ids]
    power = -0.5 * (c[:, 0:1] * dx * dx + c[:, 2:3] * dy * dy) - c[:, 1:2] * dx * dy
    alpha = np.minimum(0.99, op[ids][:, None] * np.exp(power))
    ok = (power <= 0) & (alpha >= 1.0 / 255.0) & inside[None]
    al = np.where(ok, alpha, 0.0)
    Tn = np.cumprod(1.0 - al, axis=0)
    term = Tn < 1e-4
    # applied = ok and not yet terminated (termination happens BEFORE applying the offending one)
    dead = np.maximum.accumulate(term, axis=0)
    applied = ok & ~dead
    hit = applied.reshape(n, 16, 16)
    tot['entries'] += n; tot['hits'] += int(applied.sum())
    tot['b8x4'] += int(hit.reshape(n, 4, 4, 2, 8).any(axis=(2, 4)).sum())
    tot['b8x8'] += int(hit.reshape(n, 2, 8, 2, 8).any(axis=(2, 4)).sum())
    tot['b16x8'] += int(hit.reshape(n, 2, 8, 16).any(axis=(2, 3)).sum())
    tot['b16x4'] += int(hit.reshape(n, 4, 4, 16).any(axis=(2, 3)).sum())
    tot['b4x4'] += int(hit.reshape(n, 4, 4, 4, 4).any(axis=(2, 4)).sum())
    tot['b16x16'] += int(hit.any(axis=(1, 2)).sum())
print(tot, round(time.time() - t0, 1), 's')
