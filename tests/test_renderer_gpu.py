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


# ---------------------------------------------------------------------------------------------------------------------------
# The path bench.py actually times -- DPTROrthoEnhancedRenderB200 in frame mode, tile culling on, record-staged kernels --
# held DIRECTLY against the oracle at full size (forward images / ids / radii and the whole backward chain to every leaf).
def _oracle_frame_backward(s, oo, grads, W, H, K=20):
    """C-oracle blend backward of the three passes (dptr_ortho_enhanced.py:342-376: RGB bg 0, depth bg 1, attributes bg 0 with
    opacity.detach()), then the per-Gaussian chain (uv, depth, conic, rgb) -> (position, scaling, rotation, shs) through the
    differentiable torch restatement on the CPU (itself pinned bit-exactly to the reference's own ortho functions)."""
    P = s["P"]
    uv, conic, op, idx, tr = oo["uv"], oo["conic"], s["opacity"], oo["idx_sorted"], oo["tile_range"]
    passes = {"rgb": (oo["rgb"], 0.0), "depth": (oo["depth"], 1.0), "attrs": (s["attrs"], 0.0)}
    fwd, bwd = {}, {}
    fragile = np.zeros((H, W), bool)
    for name, (feat, bg) in passes.items():
        fwd[name] = O.alpha_blending_forward(uv, conic, op, feat, idx, tr, bg, W, H, K=(K if name == "rgb" else 0), frag_eps=Hh.FRAG_EPS)
        fragile |= fwd[name]["fragile"]
    for name, (feat, bg) in passes.items():
        g = grads[name].copy()
        g[:, fragile] = 0
        bwd[name] = O.alpha_blending_backward(uv, conic, op, feat, idx, tr, bg, W, H, fwd[name]["final_T"], fwd[name]["ncontrib"], g)
    g_uv = bwd["rgb"]["dL_duv"] + bwd["depth"]["dL_duv"] + bwd["attrs"]["dL_duv"]
    g_conic = bwd["rgb"]["dL_dconic"] + bwd["depth"]["dL_dconic"] + bwd["attrs"]["dL_dconic"]
    out = {"opacity": bwd["rgb"]["dL_dopacity"] + bwd["depth"]["dL_dopacity"], "attrs": bwd["attrs"]["dL_dfeature"],
           "ndc": bwd["rgb"]["dL_duv"] * np.array([0.5 * W, 0.5 * H], np.float32)}
    leaves = {k: torch.from_numpy(s[k]).clone().requires_grad_(True) for k in ("xyz", "scaling", "rotation", "shs")}
    dirs = torch.zeros(P, 3); dirs[:, 2] = 1.0
    rgb = TR.compute_sh(leaves["shs"], 3, dirs)
    uv_t, depth_t = TR.project_point_ortho(leaves["xyz"], torch.from_numpy(s["extr"]), W, H, nearest=0.01)
    vis = depth_t.detach() != 0
    cov3d = TR.compute_cov3d(leaves["scaling"], leaves["rotation"], vis)
    conic_t, radius_t, _ = TR.ewa_project_ortho(cov3d, torch.from_numpy(s["extr"]), uv_t, W, H, vis.reshape(-1))
    assert np.array_equal(radius_t.numpy(), oo["radius"])
    torch.autograd.backward([uv_t, depth_t, conic_t, rgb],
                            [torch.from_numpy(g_uv), torch.from_numpy(bwd["depth"]["dL_dfeature"]), torch.from_numpy(g_conic),
                             torch.from_numpy(bwd["rgb"]["dL_dfeature"])])
    out.update(position=leaves["xyz"].grad.numpy(), scaling=leaves["scaling"].grad.numpy(), rotation=leaves["rotation"].grad.numpy(),
               shs=leaves["shs"].grad.numpy())
    return fwd, out, fragile


@pytest.mark.parametrize("P,W,H,track_grad", [(200_000, 854, 480, True), (200_000, 854, 480, False), (600_000, 1920, 1080, True)],
                         ids=["cfg2-bwd24x14", "cfg2-bwd24x8", "1080p-600k"])
def test_benched_frame_path_against_oracle_full_size(cuda, P, W, H, track_grad):
    """`track_grad` selects the backward dispatch: with a gradient on track_gs 11 feature channels are reduced
    (blend_rec_bwd<24,14>), without it 8 (blend_rec_bwd<24,8>).  The 1080p case covers 8160-tile grids and > 2 M keys."""
    from splatter_a_video_b200.renderer import parse_renderer
    s = Hh.scene_np(P, W, H, seed=1234, frames=50)
    oo = Hh.oracle_ortho(s, K=20)
    rng = np.random.default_rng(11)
    grads = {"rgb": rng.standard_normal((3, H, W)).astype(np.float32), "depth": rng.standard_normal((1, H, W)).astype(np.float32),
             "attrs": rng.standard_normal((19, H, W)).astype(np.float32)}
    fwd, want, fragile = _oracle_frame_backward(s, oo, grads, W, H)
    for g in grads.values():
        g[:, fragile] = 0

    a = s["attrs"]
    rd = {"position": s["xyz"], "opacity": s["opacity"], "scaling": s["scaling"], "rotation": s["rotation"], "shs": s["shs"],
          "track_gs": a[:, 0:3], "mask_attribute": a[:, 3:4], "pos_poly_feat": a[:, 4:16], "dino_attribute": a[:, 16:19]}
    rd = {k: t(v, cuda).clone().requires_grad_(track_grad or k != "track_gs") for k, v in rd.items()}
    rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    assert rnd.frame and rnd.cull
    batch = {"height": H, "width": W, "extrinsic_matrix": t(s["extr"], cuda), "intrinsic_matrix": t(s["intr"], cuda),
             "camera_center": torch.zeros(3, device=cuda), "render_attributes_list": list(ATTRS), "num_idx": 20}
    out = rnd.render_batch(rd, [batch])
    assert int(rnd.last_status.cpu()[1]) == 0
    I_culled, I_ref = int(rnd.last_status.cpu()[0]), int(oo["idx_sorted"].size)
    assert 0 < I_culled < I_ref                         # culling shortened the lists ...
    if W >= 1920:
        assert I_ref > 2_000_000 and ((W + 15) // 16) * ((H + 15) // 16) == 8160
    # ... without moving a pixel: images, first-K ids, radii against the oracle's UN-culled traversal
    ok = ~fragile
    Hh.assert_pixels_close(n(out["rgb"][0]), fwd["rgb"]["rendered"], fragile, "rgb")
    Hh.assert_pixels_close(n(out["depth"][0]), fwd["depth"]["rendered"], fragile, "depth")
    got_attr = torch.cat([out[k][0] for k in ATTRS], 0)
    Hh.assert_pixels_close(n(got_attr), fwd["attrs"]["rendered"], fragile, "attrs")
    assert np.array_equal(n(out["gs_idx"][0])[ok], fwd["rgb"]["gs_idx"][ok])
    assert np.array_equal(n(out["radii"]), oo["radius"]) and np.array_equal(n(out["visibility"]), oo["radius"] > 0)
    torch.autograd.backward([out["rgb"][0], out["depth"][0], got_attr], [t(grads["rgb"], cuda), t(grads["depth"], cuda), t(grads["attrs"], cuda)])
    for k in ("position", "scaling", "rotation", "opacity", "shs"):
        Hh.assert_grad_close(n(rd[k].grad), want[k], f"d/d{k}", norm_tol=1e-4)
    ga = want["attrs"]
    for k, sl in (("track_gs", slice(0, 3)), ("mask_attribute", slice(3, 4)), ("pos_poly_feat", slice(4, 16)), ("dino_attribute", slice(16, 19))):
        if rd[k].requires_grad:
            Hh.assert_grad_close(n(rd[k].grad), ga[:, sl], f"d/d{k}", norm_tol=1e-4)
        else:
            assert rd[k].grad is None
    Hh.assert_grad_close(n(out["viewspace_points"][0].grad), want["ndc"], "ndc.grad", norm_tol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------------
# The other two renderer mirrors (boundary B0): DPTRRender (perspective, src/pointrix/renderer/dptr.py:42-169) and
# DPTROrthoRender (src/pointrix/renderer/dptr_ortho.py): view-dependent SH, ONE blend of cat(rgb, depth[, pixel_flow]).
@pytest.mark.parametrize("name", ["DPTRRender", "DPTROrthoRender"])
def test_single_blend_renderer_mirrors_match_torch_restatement(cuda, name):
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_config("cfg1_tiny")
    W, H, P = sc.W, sc.H, sc.P
    persp = name == "DPTRRender"
    g = torch.Generator().manual_seed(5)
    flow = torch.rand(P, 2, generator=g)
    campos = torch.tensor([0.05, -0.02, -0.3])
    src = {"position": sc.frame_position(0), "opacity": sc.opacity, "scaling": sc.scaling * (2.0 if persp else 1.0), "rotation": sc.rotation,
           "shs": sc.shs, "pixel_flow": flow}
    rd = {k: v.to(cuda).clone().requires_grad_(True) for k, v in src.items()}
    rnd = parse_renderer({"name": name}, white_bg=True, device=cuda)
    batch = {"height": H, "width": W, "extrinsic_matrix": sc.extr.to(cuda), "intrinsic_matrix": sc.intr.to(cuda), "camera_center": campos.to(cuda)}
    out = rnd.render_batch(rd, [batch])
    assert out["rgb"].shape == (1, 3, H, W) and out["depth"].shape == (1, 1, H, W) and out["pixel_flow"].shape == (1, 2, H, W)
    assert out["radii"].dtype == torch.int32 and out["visibility"].dtype == torch.bool and len(out["viewspace_points"]) == 1

    cpu = {k: v.clone().requires_grad_(True) for k, v in src.items()}
    d = cpu["position"] - campos.reshape(1, 3)
    d = d / d.norm(dim=1, keepdim=True)
    rgb = TR.compute_sh(cpu["shs"], 3, d)
    if persp:
        uv, depth = TR.project_point(cpu["position"], sc.intr, sc.extr, W, H, nearest=0.01)
    else:
        uv, depth = TR.project_point_ortho(cpu["position"], sc.extr, W, H, nearest=0.01)
    vis = depth.detach() != 0
    cov3d = TR.compute_cov3d(cpu["scaling"], cpu["rotation"], vis)
    if persp:
        conic, radius, tiles = TR.ewa_project(cpu["position"], cov3d, sc.intr, sc.extr, uv, W, H, vis)
    else:
        conic, radius, tiles = TR.ewa_project_ortho(cov3d, sc.extr, uv, W, H, vis.reshape(-1))
    idx, tr = TR.sort_gaussian(uv.detach(), depth.detach(), W, H, radius, tiles)
    feat = torch.cat([rgb, depth, cpu["pixel_flow"]], 1)
    img, _, _, _ = TR.alpha_blending(uv, conic, cpu["opacity"], feat, idx, tr, 1.0, W, H)
    f = O.alpha_blending_forward(uv.detach().numpy(), conic.detach().numpy(), sc.opacity.numpy(), feat.detach().numpy(), idx.numpy(), tr.numpy(),
                                 1.0, W, H, frag_eps=Hh.FRAG_EPS)
    frag = f["fragile"]
    assert np.array_equal(n(out["radii"]), radius.numpy()) and np.array_equal(n(out["visibility"]), (radius > 0).numpy())
    got = torch.cat([out["rgb"][0], out["depth"][0], out["pixel_flow"][0]], 0)
    Hh.assert_pixels_close(n(got), img.detach().numpy(), frag, name)
    gimg = torch.randn(6, H, W, generator=g)
    gimg[:, torch.from_numpy(frag)] = 0
    got.backward(gimg.to(cuda)); img.backward(gimg)
    for k in src:
        Hh.assert_grad_close(n(rd[k].grad), cpu[k].grad.numpy(), f"{name} d/d{k}", norm_tol=2e-4)
