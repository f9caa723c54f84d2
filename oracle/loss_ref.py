"""TEST INFRASTRUCTURE — CPU restatement (differentiable PyTorch, float32) of the image losses that follow the rasterizer in the
reference training step.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this; the product (splatter_a_video_b200/losses.py -> csrc/loss.cu) never does.

Pinned: tests/golden/golden_losses.npz holds inputs, values and autograd gradients produced by the reference's own functions
(imported from /root/reference/src by tests/golden/make_loss_golden.py); tests/test_oracle_cpu.py checks this file against it.

  rgb_loss        trainer_fragGS.py:573-578 = (1-l) * l1_loss (pointrix/model/loss.py:22-38) + l * (1 - ssim (:62-112))
  depth_loss_dpt  src/loss.py:184-206 (no weight map)
  track_loss      trainer_fragGS.py:531-571 with masked_l1_loss (src/criterion.py:46-51) and denormalize_coords (src/util.py:82)
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def window_1d(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """Normalised Gaussian taps; float32 after a double-precision exp, summed in float32 (pointrix/model/loss.py:58-60)."""
    taps = torch.tensor([math.exp(-((i - size // 2) ** 2) / (2.0 * sigma ** 2)) for i in range(size)], dtype=torch.float32)
    return taps / taps.sum()


def rowwise_ssim_map(p_hwc: torch.Tensor, g_hwc: torch.Tensor, size: int = 11) -> torch.Tensor:
    """SSIM map as the trainer gets it: tensors are [1,H,W,3] when `ssim` is called, `channel = size(-3) = H`
    (loss.py:83), so every image row is a depthwise channel whose 2-D plane is (x, colour): the 11x11 window covers 11 pixels
    along x and the zero-padded 3-wide colour axis.  Returns [H,W,3]."""
    H, W, _ = p_hwc.shape
    w1 = window_1d(size).to(p_hwc.device)
    w2 = (w1[:, None] @ w1[None, :])[None, None]                       # loss.py:64 (outer product in float32)
    stack = torch.stack([p_hwc, g_hwc, p_hwc * p_hwc, g_hwc * g_hwc, p_hwc * g_hwc])        # 5 window means at once
    m = F.conv2d(stack.reshape(5 * H, 1, W, 3), w2, padding=size // 2).reshape(5, H, W, 3)
    mu_p, mu_g = m[0], m[1]
    var_p, var_g, cov = m[2] - mu_p * mu_p, m[3] - mu_g * mu_g, m[4] - mu_p * mu_g
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu_p * mu_g + c1) * (2 * cov + c2)) / ((mu_p * mu_p + mu_g * mu_g + c1) * (var_p + var_g + c2))


def rgb_loss(pred_chw: torch.Tensor, gt_hwc: torch.Tensor, lambda_dssim: float = 0.2, weight: float = 1.0):
    """Returns (loss, l1, ssim)."""
    p = pred_chw.permute(1, 2, 0)                                       # trainer_fragGS.py:573
    l1 = (p - gt_hwc).abs().mean()
    ssim = rowwise_ssim_map(p, gt_hwc).mean()
    return weight * ((1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim)), l1, ssim


def _lower_median(x: torch.Tensor) -> torch.Tensor:
    """torch.median semantics (lower middle element) with the gradient sent to the first element that holds the value."""
    flat = x.reshape(-1)
    value = flat.detach().sort().values[(flat.numel() - 1) // 2]
    first = int((flat.detach() == value).nonzero()[0, 0])
    return flat[first]


def depth_loss_dpt(pred: torch.Tensor, gt: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    t_p, t_g = _lower_median(pred), _lower_median(gt)
    s_p, s_g = (pred - t_p).abs().mean(), (gt - t_g).abs().mean()
    return weight * (((pred - t_p) / s_p - (gt - t_g) / s_g) ** 2).mean()


def track_loss(track_chw, query_xy, target_xy, visible, weights, quantile: float = 0.98, weight: float = 1.0) -> torch.Tensor:
    _, H, W = track_chw.shape
    vis = visible.reshape(-1).bool()
    if int(vis.sum()) == 0:
        return track_chw.sum() * 0.0
    q = query_xy.long()
    norm = track_chw[:2, q[:, 1], q[:, 0]].t()                          # [n,2] normalised (x,y) at the query pixels
    pix = (norm + 1.0) * torch.tensor([W, H], dtype=norm.dtype, device=norm.device) / 2.0   # util.py:82
    per_point = (pix[vis] - target_xy[vis]).abs().mean(dim=-1)
    thr = torch.quantile(per_point.detach(), quantile)
    keep = per_point.detach() <= thr
    w = weights.reshape(-1)[vis]
    loss = (per_point * w)[keep].sum() / (w[keep].sum() + 1e-8)         # criterion.py:51 with ndim = 1
    return weight * loss / max(H, W)
