"""-m gpu: the renderer plugin (boundary B0) end to end -- forward images and the full backward chain to every
render_dict tensor -- against the differentiable pure-PyTorch restatement (tiny config) and against the C oracle
plus size-independent properties at BASELINE.json's full 480p size."""
import numpy as np
import pytest
import torch

import helpers as Hh
from helpers import O, n, t
from oracle import torch_ref as TR
from splatter_a_video_b200 import synth

pytestmark = pytest.mark.gpu

ATTRS = ["track_gs", "mask_attribute", "pos_poly_feat", "dino_attribute"]


def _batch(sc, dev, K=20):
    return {"height": sc.H, "width": sc.W, "extrinsic_matrix": sc.extr.to(dev), "intrinsic_matrix": sc.intr.to(dev),
            "camera_center": torch.zeros(3, device=dev), "render_attributes_list": list(ATTRS), "num_idx": K}


def _render_dict(sc, dev, frame=0):
    d = {"position": sc.frame_position(frame), "opacity": sc.opacity, "scaling": sc.scaling, "rotation": sc.rotation, "shs": sc.shs,
         "track_gs": sc.frame_position(frame + 1), "mask_attribute": sc.attrs["mask_attribute"],
         "pos_poly_feat": sc.attrs["pos_poly_feat"], "dino_attribute": sc.attrs["dino_attribute"]}
    return {k: v.to(dev).clone().requires_grad_(True) for k, v in d.items()}


@pytest.mark.parametrize("name", ["DPTROrthoEnhancedRender", "DPTROrthoEnhancedRenderB200"])
def test_render_batch_matches_torch_restatement(cuda, name):
    from splatter_a_video_b200.renderer import parse_renderer
    try:
        from splatter_a_video_b200.gs import fused  # noqa: F401
    except ImportError:
        if name.endswith("B200"):
            pytest.skip("fused path not built")
    sc = synth.make_config("cfg1_tiny")
    W, H = sc.W, sc.H
    rd = _render_dict(sc, cuda)
    rnd = parse_renderer({"name": name}, white_bg=False, device=cuda)
    out = rnd.render_batch(rd, [_batch(sc, cuda)])
    assert out["rgb"].shape == (1, 3, H, W) and out["depth"].shape == (1, 1, H, W) and out["pos_poly_feat"].shape == (1, 12, H, W)
    assert out["gs_idx"].shape == (1, H, W, 20) and out["radii"].dtype == torch.int32 and out["visibility"].dtype == torch.bool
    assert len(out["viewspace_points"]) == 1

    # reference semantics in differentiable torch (CPU)
    cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in rd.items()}
    attrs_cpu = torch.cat([cpu[k] for k in ATTRS], 1)
    r = TR.render_ortho_frame(cpu["position"], cpu["scaling"], cpu["rotation"], cpu["opacity"], cpu["shs"], attrs_cpu, sc.extr, W, H, K=20)
    f = O.alpha_blending_forward(r["uv"].detach().numpy(), r["conic"].detach().numpy(), sc.opacity.numpy(), r["colors"].detach().numpy(),
                                 r["idx_sorted"].numpy(), r["tile_range"].numpy(), 0.0, W, H, K=20, frag_eps=Hh.FRAG_EPS)
    frag = f["fragile"]
    Hh.assert_pixels_close(n(out["rgb"][0]), r["rgb"].detach().numpy(), frag, "rgb")
    Hh.assert_pixels_close(n(out["depth"][0]), r["depth"].detach().numpy(), frag, "depth")
    got_attr = torch.cat([out[k][0] for k in ATTRS], 0)
    Hh.assert_pixels_close(n(got_attr), r["attrs"].detach().numpy(), frag, "attrs")
    assert np.array_equal(n(out["gs_idx"][0])[~frag], r["gs_idx"].numpy()[~frag])
    assert np.array_equal(n(out["radii"]), r["radius"].numpy())
    assert np.array_equal(n(out["visibility"]), (r["radius"] > 0).numpy())

    g = torch.Generator().manual_seed(0)
    grgb, gdep, gatt = torch.randn(3, H, W, generator=g), torch.randn(1, H, W, generator=g), torch.randn(19, H, W, generator=g)
    for x in (grgb, gdep, gatt):
        x[:, torch.from_numpy(frag)] = 0
    torch.autograd.backward([out["rgb"][0], out["depth"][0], got_attr], [grgb.to(cuda), gdep.to(cuda), gatt.to(cuda)])
    torch.autograd.backward([r["rgb"], r["depth"], r["attrs"]], [grgb, gdep, gatt])
    for k in ("position", "scaling", "rotation", "opacity", "shs", "track_gs", "mask_attribute", "pos_poly_feat", "dino_attribute"):
        Hh.assert_grad_close(n(rd[k].grad), cpu[k].grad.numpy(), f"d/d{k}", norm_tol=2e-4)
    # densification statistic: ndc.grad = dL_duv(RGB pass only) * [W/2, H/2]
    ndc = out["viewspace_points"][0]
    assert ndc.grad is not None and ndc.grad.shape == (sc.P, 2)
    uv_leaf = r["uv"].detach().clone().requires_grad_(True)
    img2, _, _, _ = TR.alpha_blending(uv_leaf, r["conic"].detach(), sc.opacity, r["colors"].detach(), r["idx_sorted"], r["tile_range"], 0.0, W, H)
    img2.backward(grgb)
    Hh.assert_grad_close(n(ndc.grad), (uv_leaf.grad * torch.tensor([0.5 * W, 0.5 * H])).numpy(), "ndc.grad", norm_tol=2e-4)


def test_full_size_frame_properties_and_oracle(cuda):
    """BASELINE.json configs[1] size (200k Gaussians, 854x480), one frame."""
    from splatter_a_video_b200 import gs
    s = Hh.scene_np(200_000, 854, 480, seed=1234, frames=50)
    W, H, P = s["W"], s["H"], s["P"]
    oo = Hh.oracle_ortho(s, K=20)
    dirs = torch.zeros(P, 3, device=cuda); dirs[:, 2] = 1
    rgb = gs.compute_sh(t(s["shs"], cuda), 3, dirs)
    uv, depth = gs.project_point_ortho(t(s["xyz"], cuda), t(s["extr"], cuda), W, H, nearest=0.01)
    vis = depth != 0
    cov3d = gs.compute_cov3d(t(s["scaling"], cuda), t(s["rotation"], cuda), vis)
    conic, radius, tiles = gs.ewa_project_ortho(cov3d, t(s["extr"], cuda), uv, W, H, vis.squeeze(-1))
    idx, tr = gs.sort_gaussian(uv, depth, W, H, radius, tiles)
    # bit-exact integer path at full size
    assert np.array_equal(n(radius), oo["radius"]) and np.array_equal(n(tiles), oo["tiles"])
    assert np.array_equal(n(idx), oo["idx_sorted"]) and np.array_equal(n(tr), oo["tile_range"])
    # properties: ranges tile the list exactly; inside a tile depth is non-decreasing, ties by ascending id
    trn = n(tr).astype(np.int64); idn = n(idx).astype(np.int64); dn = n(depth).reshape(-1)
    lens = trn[:, 1] - trn[:, 0]
    assert lens.min() >= 0 and lens.sum() == idn.size == int(n(tiles).sum())
    nz = lens > 0
    assert np.array_equal(trn[nz, 0], np.cumsum(lens[nz]) - lens[nz])   # tiles appear in id order, back to back
    same_tile = np.ones(idn.size, bool); same_tile[trn[nz, 0]] = False
    dd = dn[idn]
    inner = same_tile[1:]
    assert (dd[1:][inner] >= dd[:-1][inner]).all()
    tie = inner & (dd[1:] == dd[:-1])
    assert (idn[1:][tie] > idn[:-1][tie]).all()
    # blend: linearity in the features + transmittance bound + oracle parity
    op = t(s["opacity"], cuda)
    img, ncontrib, gs_idx = gs.alpha_blending_enhanced(uv, conic, op, rgb, idx, tr, 0.0, W, H, None, None, K=20)
    f = oo["blend_rgb"]
    Hh.assert_pixels_close(n(img), f["rendered"], f["fragile"], "full-size rgb")
    ok = ~f["fragile"]
    assert np.array_equal(n(ncontrib)[ok], f["ncontrib"][ok]) and np.array_equal(n(gs_idx)[ok], f["gs_idx"][ok])
    ones = torch.ones(P, 1, device=cuda)
    acc = gs.alpha_blending(uv, conic, op, ones, idx, tr, 0.0, W, H)            # accumulated alpha = 1 - T
    assert float(acc.max()) <= 1.0 - 1e-4 + 1e-6 and float(acc.min()) >= 0.0
    img2 = gs.alpha_blending(uv, conic, op, 2.0 * rgb + 1.0, idx, tr, 0.0, W, H)
    assert float((img2 - (2.0 * img + acc)).abs().max()) <= 2e-5
    white = gs.alpha_blending(uv, conic, op, rgb, idx, tr, 1.0, W, H)
    assert float((white - (img + (1.0 - acc))).abs().max()) <= 2e-5
    # gradient parity at full size (19 attribute channels)
    feat_np = s["attrs"]
    g_np = np.random.default_rng(5).standard_normal((19, H, W)).astype(np.float32)
    fo = O.alpha_blending_forward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], 0.0, W, H,
                                  frag_eps=Hh.FRAG_EPS)
    g_np[:, fo["fragile"]] = 0
    leaves = [x.detach().clone().requires_grad_(True) for x in (uv, conic, op, t(feat_np, cuda))]
    gs.alpha_blending(*leaves, idx, tr, 0.0, W, H).backward(t(g_np, cuda))
    b = O.alpha_blending_backward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], 0.0, W, H,
                                  fo["final_T"], fo["ncontrib"], g_np)
    for name, got in zip(("dL_duv", "dL_dconic", "dL_dopacity", "dL_dfeature"), leaves):
        Hh.assert_grad_close(n(got.grad), b[name], f"full-size {name}")


def test_diff_gaussian_rasterization_shim(cuda):
    """Boundary B2 (call site only, parity unpinned): the shim must behave like the B1 perspective pipeline with a colour
    background, return int radii and deliver screen-space gradients through means2D.grad."""
    import math
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from splatter_a_video_b200 import gs
    s = Hh.scene_np(3000, 96, 64, seed=21)
    W, H = 96, 64
    fx = fy = W / 2.0
    view = torch.eye(4)
    settings = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=W / (2 * fx), tanfovy=H / (2 * fy),
                                             bg=torch.tensor([0.2, 0.4, 0.6], device=cuda), scale_modifier=1.0, viewmatrix=view.to(cuda),
                                             projmatrix=torch.eye(4, device=cuda), sh_degree=3, campos=torch.zeros(3, device=cuda))
    means3D = t(s["xyz"], cuda).requires_grad_(True)
    means2D = torch.zeros(s["P"], 3, device=cuda, requires_grad=True)
    img, radii = GaussianRasterizer(settings)(means3D=means3D, means2D=means2D, shs=t(s["shs"], cuda), colors_precomp=None,
                                              opacities=t(s["opacity"], cuda), scales=t(s["scaling"], cuda), rotations=t(s["rotation"], cuda),
                                              cov3D_precomp=None)
    assert img.shape == (3, H, W) and radii.dtype == torch.int32 and radii.shape == (s["P"],)
    # same thing through the B1 operators
    intr = torch.tensor([fx, fy, W / 2.0, H / 2.0], device=cuda)
    extr = torch.eye(4, device=cuda)[:3]
    d = means3D.detach(); d = d / d.norm(dim=1, keepdim=True)
    col = gs.compute_sh(t(s["shs"], cuda), 3, d)
    ones = torch.ones(s["P"], 1, device=cuda)
    ref = gs.rasterization(means3D.detach(), t(s["scaling"], cuda), t(s["rotation"], cuda), t(s["opacity"], cuda),
                           torch.cat([col, ones], 1), intr, extr, W, H, 0.0)
    want = ref[:3] + (1 - ref[3:4]) * settings.bg.reshape(3, 1, 1)
    assert float((img - want).abs().max()) <= 1e-6
    img.sum().backward()
    assert means2D.grad is not None and float(means2D.grad[:, :2].abs().sum()) > 0 and float(means2D.grad[:, 2].abs().sum()) == 0
    assert torch.isfinite(means3D.grad).all()
