"""Generates tests/golden/golden_losses.npz by running the REFERENCE's own loss functions (imported, unmodified, from
/root/reference/src) on small seeded inputs, with autograd gradients.  Runs only in the authoring container (the reference
checkout does not travel); the fixture it writes is committed.

    python tests/golden/make_loss_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ploss = load("ref_pointrix_loss", os.path.join(REF, "pointrix/model/loss.py"))      # l1_loss, ssim
    dloss = load("ref_loss", os.path.join(REF, "loss.py"))                              # depth_loss_dpt
    crit = load("ref_criterion", os.path.join(REF, "criterion.py"))                     # masked_l1_loss
    out = {}
    g = torch.Generator().manual_seed(20240607)
    # ---- rgb: the exact expression of trainer_fragGS.py:573-578 (tensors reshaped to [1,H,W,3])
    for tag, (H, W) in {"a": (12, 40), "b": (7, 150)}.items():
        pred = torch.rand(3, H, W, generator=g).requires_grad_(True)
        gt = (pred.detach().permute(1, 2, 0) + 0.15 * torch.randn(H, W, 3, generator=g)).clamp(0, 1).contiguous()
        pred_rgb1 = pred.permute(1, 2, 0).reshape(1, -1, 3)
        gt_rgb1 = gt.reshape(1, -1, 3)
        l1 = ploss.l1_loss(pred_rgb1.reshape(-1, H, W, 3), gt_rgb1.reshape(-1, H, W, 3))
        ss = ploss.ssim(pred_rgb1.reshape(-1, H, W, 3), gt_rgb1.reshape(-1, H, W, 3))
        loss = (1.0 - 0.2) * l1 + 0.2 * (1 - ss)
        (grad,) = torch.autograd.grad(loss, pred)
        out.update({f"rgb_{tag}_pred": pred.detach().numpy(), f"rgb_{tag}_gt": gt.numpy(), f"rgb_{tag}_loss": np.float32(loss.item()),
                    f"rgb_{tag}_l1": np.float32(l1.item()), f"rgb_{tag}_ssim": np.float32(ss.item()), f"rgb_{tag}_grad": grad.numpy()})
    # ---- depth: depth_loss_dpt(depth[h,w,1], gt_depth[h,w,1]) (trainer_fragGS.py:593-601); odd and even element counts
    for tag, (H, W) in {"a": (9, 21), "b": (16, 30)}.items():
        pred = (0.5 + 1.5 * torch.rand(H, W, 1, generator=g)).requires_grad_(True)
        gt = 0.3 + 2.0 * torch.rand(H, W, 1, generator=g) + 0.5 * pred.detach()
        loss = dloss.depth_loss_dpt(pred, gt)
        (grad,) = torch.autograd.grad(loss, pred)
        out.update({f"depth_{tag}_pred": pred.detach().numpy(), f"depth_{tag}_gt": gt.numpy(), f"depth_{tag}_loss": np.float32(loss.item()),
                    f"depth_{tag}_grad": grad.numpy()})
    # ---- track: trainer_fragGS.py:551-571 (mask scatter + boolean gather on raster-ordered unique query pixels, util.py:82)
    H, W, step = 24, 36, 4
    ys, xs = torch.meshgrid(torch.arange(1, H, step), torch.arange(2, W, step), indexing="ij")
    query = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)                            # raster order, unique
    n = query.shape[0]
    track = (torch.rand(1, 3, H, W, generator=g) * 2 - 1).requires_grad_(True)           # render_results['track_gs']
    gt_tracks = torch.stack([torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H], -1)
    visibles = torch.rand(n, generator=g) > 0.25
    conf = torch.rand(n, generator=g)
    w_interval = torch.exp(torch.tensor(-2 * 7.0 / 50))
    predicted_track_gs = track.permute(0, 2, 3, 1)
    predicted_track_2d = (predicted_track_gs[..., :2] + 1.) * torch.tensor([W, H]) / 2.   # util.denormalize_coords (util.py:82)
    track_weights = conf[..., None] * w_interval
    masks_flatten = torch.zeros_like(predicted_track_2d[..., 0])
    qp = query.to(torch.int64)
    masks_flatten[0, qp[:, 1], qp[:, 0]] = 1.0
    masks_flatten = masks_flatten.reshape(-1, H * W) > 0.5
    p2 = predicted_track_2d.reshape(-1, H * W, 2)
    loss = crit.masked_l1_loss(p2[masks_flatten][visibles], gt_tracks[visibles], mask=track_weights[visibles], quantile=0.98) / max(H, W)
    (grad,) = torch.autograd.grad(loss, track)
    out.update({"track_img": track.detach()[0].numpy(), "track_query": query.numpy().astype(np.int32), "track_target": gt_tracks.numpy(),
                "track_visible": visibles.numpy(), "track_weights": track_weights.reshape(-1).numpy(), "track_loss": np.float32(loss.item()),
                "track_grad": grad[0].numpy()})
    path = os.path.join(HERE, "golden_losses.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
