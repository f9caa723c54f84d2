"""Shim for the `diff_gaussian_rasterization` call site (boundary B2).

The reference only *imports* this package in one renderer that its config never selects
(/root/reference/src/pointrix/renderer/base_splatting.py:17,123-174); the package itself (graphdeco-inria's rasterizer) is a
third-party dependency that is NOT vendored in the reference tree and no version is pinned, so there is nothing to pin
parity against: **parity at B2 is unpinned** (DESIGN.md).  This shim keeps the call site working by mapping the
`GaussianRasterizationSettings` / `GaussianRasterizer` surface onto the same sm_100a kernels the B1 operators use:
perspective projection from (viewmatrix, tanfov), SH evaluated inside with `sh_degree`, a 3-vector background, screen-space
gradients delivered through `means2D.grad`.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import nn

from splatter_a_video_b200 import gs as _gs


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor      # 4x4 world->view, stored transposed (row-vector convention), as 3DGS code bases pass it
    projmatrix: torch.Tensor      # unused: the symmetric pinhole it encodes is rebuilt from tanfov
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if cov3D_precomp is not None:
            raise NotImplementedError("cov3D_precomp is not supported by this shim (scales + rotations are)")
        H, W = int(rs.image_height), int(rs.image_width)
        dev = means3D.device
        extr = rs.viewmatrix.to(dev).t()[:3, :4].contiguous()
        intr = torch.tensor([W / (2.0 * rs.tanfovx), H / (2.0 * rs.tanfovy), W / 2.0, H / 2.0], dtype=torch.float32, device=dev)
        if colors_precomp is None:
            d = means3D - rs.campos.to(dev).reshape(1, 3)
            d = d / d.norm(dim=1, keepdim=True)
            nb = (rs.sh_degree + 1) ** 2
            colors = _gs.compute_sh(shs[:, :nb].contiguous(), rs.sh_degree, d)
        else:
            colors = colors_precomp
        uv, depth = _gs.project_point(means3D, intr, extr, W, H)          # near 0.2 / extent 1.3: the 3DGS frustum test
        visible = depth != 0
        cov3d = _gs.compute_cov3d(scales * rs.scale_modifier, rotations, visible)
        conic, radius, tiles = _gs.ewa_project(means3D, cov3d, intr, extr, uv, W, H, visible)
        idx_sorted, tile_range = _gs.sort_gaussian(uv, depth, W, H, radius, tiles)
        feat = torch.cat([colors, torch.ones_like(colors[:, :1])], dim=1)
        img = _gs.alpha_blending(uv, conic, opacities, feat, idx_sorted, tile_range, 0.0, W, H, means2D[:, :2] if means2D is not None else None)
        bg = rs.bg.to(dev).reshape(3, 1, 1)
        rendered = img[:3] + (1.0 - img[3:4]) * bg                         # T_final * bg per channel
        return rendered, radius
