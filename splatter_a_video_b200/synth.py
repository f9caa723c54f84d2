"""Synthetic DAVIS-shaped scenes -- the only data source of bench.py and the tests.

There is no dataset in the reference tree (data paths are hard-coded to the author's disk,
/root/reference/src/trainer_fragGS.py:258,285), so every measured workload is generated here:
per-Gaussian SoA tensors with the shapes of ``FragModel.forward``'s ``render_dict``
(/root/reference/src/frag_model.py:122-137) plus the trainer's attribute list
(track_gs 3 + mask 1 + pos_poly_feat 12 + dino 3 = 19 channels, frag_gs_v10.yaml:115-118).
All generation happens on the CPU with a fixed seed so CPU oracle and GPU see identical bits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

# name -> (P, frames, W, H): BASELINE.json "configs"
CONFIGS = {
    "cfg1_tiny": dict(P=1_000, frames=2, W=64, H=64),
    "cfg2_davis480p": dict(P=200_000, frames=50, W=854, H=480),
    "cfg3_480p_500k": dict(P=500_000, frames=80, W=854, H=480),
    "cfg4_1080p_2m": dict(P=2_000_000, frames=120, W=1920, H=1080),
}

ATTR_CHANNELS = {"track_gs": 3, "mask_attribute": 1, "pos_poly_feat": 12, "dino_attribute": 3}


@dataclass
class Scene:
    """One clip worth of Gaussians.  ``position`` is the frame-0 (base) position; per-frame positions
    come from ``frame_position(t)`` (cubic spline over nodes every 5 frames, mirroring
    dynamic_gaussian_with_base_point_cloud.py:66-78,236-250)."""
    P: int
    frames: int
    W: int
    H: int
    position: torch.Tensor        # [P,3]
    scaling: torch.Tensor         # [P,3]  (already exp-activated)
    rotation: torch.Tensor        # [P,4]  unit quaternion (r,x,y,z)
    opacity: torch.Tensor         # [P,1]  (already sigmoid-activated)
    shs: torch.Tensor             # [P,16,3]
    attrs: Dict[str, torch.Tensor]
    nodes: torch.Tensor           # [P, n_nodes, 3] displacement spline nodes
    extr: torch.Tensor            # [4,4] world->camera (identity)
    intr: torch.Tensor            # [4] fx,fy,cx,cy (perspective sweeps only)

    def frame_position(self, t: int) -> torch.Tensor:
        return self.position + eval_spline(self.nodes, float(t) / 5.0)

    def to(self, device) -> "Scene":
        kw = {}
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                kw[k] = v.to(device)
            elif isinstance(v, dict):
                kw[k] = {a: b.to(device) for a, b in v.items()}
            else:
                kw[k] = v
        return Scene(**kw)

    def attr_features(self, t2: int) -> torch.Tensor:
        """[P,19] attribute matrix the trainer blends in its third pass
        (trainer_fragGS.py:506-512: track_gs = position at frame ids2)."""
        a = dict(self.attrs)
        a["track_gs"] = self.frame_position(t2)
        return torch.cat([a[k] for k in ATTR_CHANNELS], dim=1).contiguous()


def eval_spline(nodes: torch.Tensor, u: float) -> torch.Tensor:
    """Uniform Catmull-Rom cubic through ``nodes[:, k]`` at parameter u (in node units)."""
    n = nodes.shape[1]
    k = min(max(int(math.floor(u)), 0), n - 2)
    s = u - k
    p0 = nodes[:, max(k - 1, 0)]
    p1 = nodes[:, k]
    p2 = nodes[:, k + 1]
    p3 = nodes[:, min(k + 2, n - 1)]
    s2, s3 = s * s, s * s * s
    return 0.5 * ((2 * p1) + (-p0 + p2) * s + (2 * p0 - 5 * p1 + 4 * p2 - p3) * s2 + (-p0 + 3 * p1 - 3 * p2 + p3) * s3)


def make_scene(P: int, frames: int, W: int, H: int, seed: int = 1234, sigma_px: float = 1.5,
               heavy_frac: float = 0.02, heavy_mult: float = 8.0) -> Scene:
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(P, 2, generator=g) * 2 - 1
    z = torch.rand(P, 1, generator=g) * 1.5 + 0.5
    position = torch.cat([xy, z], 1)
    s0 = 2.0 * sigma_px / W                     # sigma_px pixels after the ortho Jacobian diag(W/2, H/2)
    log_s = math.log(s0) + 0.35 * torch.randn(P, 3, generator=g)
    heavy = torch.rand(P, 1, generator=g) < heavy_frac
    scaling = torch.exp(log_s) * torch.where(heavy, torch.tensor(heavy_mult), torch.tensor(1.0))
    q = torch.randn(P, 4, generator=g)
    rotation = q / q.norm(dim=1, keepdim=True)
    opacity = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g))
    shs = torch.empty(P, 16, 3)
    shs[:, 0] = (torch.rand(P, 3, generator=g) * 2 - 1) / 0.28209479177387814
    shs[:, 1:] = 0.05 * torch.randn(P, 15, 3, generator=g)
    attrs = {k: torch.rand(P, c, generator=g) for k, c in ATTR_CHANNELS.items()}
    n_nodes = max(2, math.ceil(frames / 5) + 1)
    nodes = 0.02 * torch.randn(P, n_nodes, 3, generator=g)
    nodes[:, 0] = 0
    extr = torch.eye(4)
    intr = torch.tensor([W / 2.0, W / 2.0, W / 2.0, H / 2.0])
    return Scene(P, frames, W, H, position.contiguous(), scaling.contiguous(), rotation.contiguous(),
                 opacity.contiguous(), shs.contiguous(), attrs, nodes.contiguous(), extr, intr)


def make_config(name: str, seed: int = 1234, **overrides) -> Scene:
    cfg = dict(CONFIGS[name]); cfg.update(overrides)
    return make_scene(cfg["P"], cfg["frames"], cfg["W"], cfg["H"], seed=seed)
