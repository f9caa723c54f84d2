"""-m gpu parity tests: every `dptr.gs`-compatible op (through the C ABI) against the CPU oracle.

Integer / index outputs must be bit-exact; pixels <= 1e-4; gradients <= 1e-3 relative (helpers.py).
"""
import numpy as np
import pytest
import torch

import helpers as Hh
from helpers import O, n, t

pytestmark = pytest.mark.gpu

SIZES = [  # (P, W, H, seed)
    (1_000, 64, 64, 1234),       # BASELINE.json configs[0] shape
    (30_000, 333, 250, 7),       # ragged: W, H not multiples of 16
]


@pytest.fixture(scope="module", params=SIZES, ids=lambda s: f"P{s[0]}_{s[1]}x{s[2]}")
def case(request, cuda):
    P, W, H, seed = request.param
    s = Hh.scene_np(P, W, H, seed=seed)
    return s, Hh.oracle_ortho(s), Hh.oracle_persp(s, nearest=0.01)


def gs_mod():
    import dptr.gs as gs  # the alias package: what reference call sites import
    return gs


# ------------------------------------------------------------------------------------------------ per-Gaussian stages
def test_project_point_perspective(case, cuda):
    s, _, op = case
    gs = gs_mod()
    xyz = t(s["xyz"], cuda).requires_grad_(True)
    intr = t(s["intr"], cuda).requires_grad_(True)
    extr = t(s["extr"][:3], cuda).requires_grad_(True)
    uv, depth = gs.project_point(xyz, intr, extr, s["W"], s["H"], nearest=0.01)
    assert np.array_equal(n(uv), op["uv"]) and np.array_equal(n(depth), op["depth"]), "uv/depth must be bit-exact"
    g = torch.Generator(device="cpu").manual_seed(3)
    guv = torch.randn(uv.shape, generator=g).to(cuda); gd = torch.randn(depth.shape, generator=g).to(cuda)
    torch.autograd.backward([uv, depth], [guv, gd])
    ox, oi, oe = O.project_point_backward(s["xyz"], s["intr"], s["extr"], op["depth"], n(guv), n(gd), True, True)
    Hh.assert_grad_close(n(xyz.grad), ox, "dL_dxyz")
    Hh.assert_grad_close(n(intr.grad), oi, "dL_dintr", norm_tol=1e-3)
    Hh.assert_grad_close(n(extr.grad), oe, "dL_dextr", norm_tol=1e-3)


def test_project_point_default_culling(case, cuda):
    s, _, _ = case
    gs = gs_mod()
    uv, depth = gs.project_point(t(s["xyz"], cuda), t(s["intr"], cuda), t(s["extr"], cuda), s["W"], s["H"])
    ouv, od = O.project_point(s["xyz"], s["intr"], s["extr"], s["W"], s["H"])
    assert np.array_equal(n(uv), ouv) and np.array_equal(n(depth), od)
    assert (od == 0).any() and (od != 0).any(), "case must exercise culling"


def test_project_point_ortho(case, cuda):
    s, oo, _ = case
    from splatter_a_video_b200 import gs
    xyz = t(s["xyz"], cuda).requires_grad_(True)
    uv, depth = gs.project_point_ortho(xyz, t(s["extr"], cuda), s["W"], s["H"], nearest=0.01)
    assert np.array_equal(n(uv), oo["uv"]) and np.array_equal(n(depth), oo["depth"])
    g = torch.Generator().manual_seed(5)
    guv = torch.randn(uv.shape, generator=g); gd = torch.randn(depth.shape, generator=g)
    torch.autograd.backward([uv, depth], [guv.to(cuda), gd.to(cuda)])
    from oracle import torch_ref as TR
    x = torch.from_numpy(s["xyz"]).requires_grad_(True)
    u2, d2 = TR.project_point_ortho(x, torch.from_numpy(s["extr"]), s["W"], s["H"], nearest=0.01)
    torch.autograd.backward([u2, d2], [guv, gd])
    Hh.assert_grad_close(n(xyz.grad), x.grad.numpy(), "ortho dL_dxyz")


def test_compute_cov3d(case, cuda):
    s, oo, _ = case
    gs = gs_mod()
    sc = t(s["scaling"], cuda).requires_grad_(True); q = t(s["rotation"], cuda).requires_grad_(True)
    vis = oo["vis"].copy(); vis[::7] = False
    cov = gs.compute_cov3d(sc, q, t(vis, cuda).reshape(-1, 1))
    ref = O.compute_cov3d(s["scaling"], s["rotation"], vis)
    assert np.array_equal(n(cov), ref), "cov3d must be bit-exact (IEEE fp32, no FMA)"
    gcov = torch.randn(cov.shape, generator=torch.Generator().manual_seed(1))
    cov.backward(gcov.to(cuda))
    gs_, gq_ = O.compute_cov3d_backward(s["scaling"], s["rotation"], vis, gcov.numpy())
    assert np.array_equal(n(sc.grad), gs_) and np.array_equal(n(q.grad), gq_)
    assert np.all(n(cov)[::7] == 0)


def test_ewa_project_perspective(case, cuda):
    s, _, op = case
    gs = gs_mod()
    xyz = t(s["xyz"], cuda).requires_grad_(True)
    cov3d = t(op["cov3d"], cuda).requires_grad_(True)
    intr = t(s["intr"], cuda).requires_grad_(True); extr = t(s["extr"][:3], cuda).requires_grad_(True)
    conic, radius, tiles = gs.ewa_project(xyz, cov3d, intr, extr, t(op["uv"], cuda), s["W"], s["H"],
                                          t(op["vis"], cuda).reshape(-1, 1))
    assert radius.dtype == torch.int32 and tiles.dtype == torch.int32
    assert np.array_equal(n(radius), op["radius"]) and np.array_equal(n(tiles), op["tiles"])
    assert np.array_equal(n(conic), op["conic"])
    gc = torch.randn(conic.shape, generator=torch.Generator().manual_seed(2))
    conic.backward(gc.to(cuda))
    ox, oc, oi, oe = O.ewa_project_backward(s["xyz"], op["cov3d"], s["intr"], s["extr"], op["radius"], gc.numpy(), True, True)
    assert np.array_equal(n(xyz.grad), ox) and np.array_equal(n(cov3d.grad), oc)
    Hh.assert_grad_close(n(intr.grad)[:2], oi[:2], "ewa dL_dintr", norm_tol=1e-3)
    Hh.assert_grad_close(n(extr.grad), oe, "ewa dL_dextr", norm_tol=1e-3)


def test_ewa_project_ortho(case, cuda):
    s, oo, _ = case
    from splatter_a_video_b200 import gs
    cov3d = t(oo["cov3d"], cuda).requires_grad_(True)
    conic, radius, tiles = gs.ewa_project_ortho(cov3d, t(s["extr"], cuda), t(oo["uv"], cuda), s["W"], s["H"], t(oo["vis"], cuda))
    assert np.array_equal(n(radius), oo["radius"]) and np.array_equal(n(tiles), oo["tiles"])
    assert np.array_equal(n(conic), oo["conic"])
    gc = torch.randn(conic.shape, generator=torch.Generator().manual_seed(2))
    conic.backward(gc.to(cuda))
    from oracle import torch_ref as TR
    c = torch.from_numpy(oo["cov3d"]).requires_grad_(True)
    k, _, _ = TR.ewa_project_ortho(c, torch.from_numpy(s["extr"]), torch.from_numpy(oo["uv"]), s["W"], s["H"], torch.from_numpy(oo["vis"]))
    k.backward(gc)
    Hh.assert_grad_close(n(cov3d.grad), c.grad.numpy(), "ortho dL_dcov3d")


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
@pytest.mark.parametrize("free", [False, True])
def test_compute_sh(case, cuda, deg, free):
    s, oo, _ = case
    gs = gs_mod()
    P = s["P"]
    nb = (deg + 1) ** 2
    shs_np = np.ascontiguousarray(s["shs"][:, :nb])
    rng = np.random.default_rng(deg)
    dirs_np = rng.standard_normal((P, 3)).astype(np.float32); dirs_np /= np.linalg.norm(dirs_np, axis=1, keepdims=True)
    vis = np.ones(P, bool); vis[::5] = False
    shs = t(shs_np, cuda).requires_grad_(True); dirs = t(dirs_np, cuda).requires_grad_(True)
    fn = gs.compute_sh_free if free else gs.compute_sh
    col = fn(shs, deg, dirs, t(vis, cuda))
    ref, clamped = O.compute_sh(shs_np, deg, dirs_np, vis, free=free)
    assert np.array_equal(n(col), ref)
    g = torch.randn(col.shape, generator=torch.Generator().manual_seed(4))
    col.backward(g.to(cuda))
    gsh, gd = O.compute_sh_backward(shs_np, deg, dirs_np, vis, clamped, g.numpy())
    assert np.array_equal(n(shs.grad), gsh) and np.array_equal(n(dirs.grad), gd)


# ------------------------------------------------------------------------------------------------ binning
def test_sort_gaussian(case, cuda):
    s, oo, op = case
    gs = gs_mod()
    for o in (oo, op):
        idx, tr = gs.sort_gaussian(t(o["uv"], cuda), t(o["depth"], cuda), s["W"], s["H"], t(o["radius"], cuda), t(o["tiles"], cuda))
        assert idx.dtype == torch.int32 and tr.dtype == torch.int32
        assert np.array_equal(n(idx), o["idx_sorted"]), "idx_sorted must be bit-exact"
        assert np.array_equal(n(tr), o["tile_range"]), "tile_range must be bit-exact"


def test_sort_gaussian_empty(cuda):
    gs = gs_mod()
    P = 64
    z = torch.zeros(P, 2, device=cuda)
    idx, tr = gs.sort_gaussian(z, torch.zeros(P, 1, device=cuda), 64, 48, torch.zeros(P, dtype=torch.int32, device=cuda),
                               torch.zeros(P, dtype=torch.int32, device=cuda))
    assert idx.numel() == 0 and tr.shape == (12, 2) and int(tr.abs().sum()) == 0


def test_sort_depth_ties_resolve_by_id(cuda):
    """Equal (tile, depth) keys must come out in ascending Gaussian id (stable sort of the emission order)."""
    gs = gs_mod()
    P, W, H = 500, 32, 32
    uv = np.full((P, 2), 15.5, np.float32); depth = np.full((P, 1), 1.0, np.float32)
    radius = np.full(P, 3, np.int32); tiles = np.full(P, 4, np.int32)
    idx, tr = gs.sort_gaussian(t(uv, cuda), t(depth, cuda), W, H, t(radius, cuda), t(tiles, cuda))
    oi, otr = O.sort_gaussian(uv, depth, W, H, radius, tiles)
    assert np.array_equal(n(idx), oi) and np.array_equal(n(tr), otr)
    assert np.array_equal(n(idx)[:P], np.arange(P))


# ------------------------------------------------------------------------------------------------ blending
def _blend_inputs(s, oo, C, seed):
    rng = np.random.default_rng(seed)
    feat = rng.random((s["P"], C), dtype=np.float32)
    gimg = rng.standard_normal((C, s["H"], s["W"])).astype(np.float32)
    return feat, gimg


@pytest.mark.parametrize("C,bg", [(1, 1.0), (3, 0.0), (4, 0.5), (8, 0.0), (19, 0.0), (23, 1.0), (24, 0.0), (37, 0.25), (70, 0.0)])
def test_alpha_blending_fwd_bwd(case, cuda, C, bg):
    s, oo, _ = case
    gs = gs_mod()
    W, H = s["W"], s["H"]
    feat_np, g_np = _blend_inputs(s, oo, C, C)
    uv = t(oo["uv"], cuda).requires_grad_(True); conic = t(oo["conic"], cuda).requires_grad_(True)
    op = t(s["opacity"], cuda).requires_grad_(True); feat = t(feat_np, cuda).requires_grad_(True)
    ndc = torch.zeros_like(uv, requires_grad=True); abs_ndc = torch.zeros_like(uv, requires_grad=True)
    idx, tr = t(oo["idx_sorted"], cuda), t(oo["tile_range"], cuda)
    img = gs.alpha_blending(uv, conic, op, feat, idx, tr, bg, W, H, ndc, abs_ndc)
    f = O.alpha_blending_forward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], bg, W, H,
                                 frag_eps=Hh.FRAG_EPS)
    assert img.shape == (C, H, W)
    Hh.assert_pixels_close(n(img), f["rendered"], f["fragile"], f"blend C={C}")
    # gradients are compared where the forward decisions provably agree: zero the upstream gradient on fragile pixels
    g_np[:, f["fragile"]] = 0
    img.backward(t(g_np, cuda))
    b = O.alpha_blending_backward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], bg, W, H,
                                  f["final_T"], f["ncontrib"], g_np)
    Hh.assert_grad_close(n(uv.grad), b["dL_duv"], "dL_duv")
    Hh.assert_grad_close(n(conic.grad), b["dL_dconic"], "dL_dconic")
    Hh.assert_grad_close(n(op.grad), b["dL_dopacity"], "dL_dopacity")
    Hh.assert_grad_close(n(feat.grad), b["dL_dfeature"], "dL_dfeature")
    half = np.array([0.5 * W, 0.5 * H], np.float32)
    Hh.assert_grad_close(n(ndc.grad), b["dL_duv"] * half, "dL_dndc")
    Hh.assert_grad_close(n(abs_ndc.grad), b["dL_dabs_uv"] * half, "dL_dabs_ndc")


@pytest.mark.parametrize("K,trunc", [(20, False), (10, False), (4, True)])
def test_alpha_blending_enhanced(case, cuda, K, trunc):
    s, oo, _ = case
    gs = gs_mod()
    W, H = s["W"], s["H"]
    uv = t(oo["uv"], cuda).requires_grad_(True); conic = t(oo["conic"], cuda).requires_grad_(True)
    op = t(s["opacity"], cuda).requires_grad_(True); feat = t(oo["rgb"], cuda).requires_grad_(True)
    ndc = torch.zeros_like(uv, requires_grad=True)
    img, ncontrib, gs_idx = gs.alpha_blending_enhanced(uv, conic, op, feat, t(oo["idx_sorted"], cuda), t(oo["tile_range"], cuda),
                                                       0.0, W, H, ndc, None, K=K, enable_truncation=trunc)
    f = O.alpha_blending_forward(oo["uv"], oo["conic"], s["opacity"], oo["rgb"], oo["idx_sorted"], oo["tile_range"], 0.0, W, H,
                                 K=K, enable_truncation=trunc, frag_eps=Hh.FRAG_EPS)
    ok = ~f["fragile"]
    Hh.assert_pixels_close(n(img), f["rendered"], f["fragile"], "enhanced rgb")
    assert ncontrib.dtype == torch.int32 and gs_idx.dtype == torch.int32 and gs_idx.shape == (H, W, K)
    assert np.array_equal(n(ncontrib)[ok], f["ncontrib"][ok]), "ncontrib must be exact"
    assert np.array_equal(n(gs_idx)[ok], f["gs_idx"][ok]), "gs_idx must be exact"
    g_np = np.random.default_rng(0).standard_normal((3, H, W)).astype(np.float32)
    g_np[:, f["fragile"]] = 0
    img.backward(t(g_np, cuda))
    b = O.alpha_blending_backward(oo["uv"], oo["conic"], s["opacity"], oo["rgb"], oo["idx_sorted"], oo["tile_range"], 0.0, W, H,
                                  f["final_T"], f["ncontrib"], g_np)
    for name, got in (("dL_duv", uv.grad), ("dL_dconic", conic.grad), ("dL_dopacity", op.grad), ("dL_dfeature", feat.grad)):
        Hh.assert_grad_close(n(got), b[name], f"enhanced {name}")


@pytest.mark.parametrize("C", [3, 12])
def test_alpha_blending_with_bias(case, cuda, C):
    s, oo, _ = case
    gs = gs_mod()
    W, H = s["W"], s["H"]
    feat_np, g_np = _blend_inputs(s, oo, C, 100 + C)
    bias_np = (0.05 * np.random.default_rng(9).random((s["P"], 1))).astype(np.float32)
    uv = t(oo["uv"], cuda).requires_grad_(True); conic = t(oo["conic"], cuda).requires_grad_(True)
    op = t(s["opacity"], cuda).requires_grad_(True); feat = t(feat_np, cuda).requires_grad_(True)
    bias = t(bias_np, cuda).requires_grad_(True)
    img = gs.alpha_blending_with_bias(uv, conic, op, feat, bias, t(oo["idx_sorted"], cuda), t(oo["tile_range"], cuda), 0.3, W, H)
    f = O.alpha_blending_forward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], 0.3, W, H,
                                 opacity_bias=bias_np, frag_eps=Hh.FRAG_EPS)
    Hh.assert_pixels_close(n(img), f["rendered"], f["fragile"], "bias blend")
    g_np[:, f["fragile"]] = 0
    img.backward(t(g_np, cuda))
    b = O.alpha_blending_backward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], 0.3, W, H,
                                  f["final_T"], f["ncontrib"], g_np, opacity_bias=bias_np)
    for name, got in (("dL_duv", uv.grad), ("dL_dconic", conic.grad), ("dL_dopacity", op.grad), ("dL_dfeature", feat.grad),
                      ("dL_dopacity_bias", bias.grad)):
        Hh.assert_grad_close(n(got), b[name], f"bias {name}")


def test_blend_masked_gradients(case, cuda):
    """Gradient parity with the upstream gradient zeroed on fragile pixels (always comparable)."""
    s, oo, _ = case
    gs = gs_mod()
    W, H, C = s["W"], s["H"], 19
    feat_np, g_np = _blend_inputs(s, oo, C, 77)
    f = O.alpha_blending_forward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], 0.0, W, H,
                                 frag_eps=Hh.FRAG_EPS)
    g_np[:, f["fragile"]] = 0
    uv = t(oo["uv"], cuda).requires_grad_(True); conic = t(oo["conic"], cuda).requires_grad_(True)
    op = t(s["opacity"], cuda).requires_grad_(True); feat = t(feat_np, cuda).requires_grad_(True)
    img = gs.alpha_blending(uv, conic, op, feat, t(oo["idx_sorted"], cuda), t(oo["tile_range"], cuda), 0.0, W, H)
    img.backward(t(g_np, cuda))
    b = O.alpha_blending_backward(oo["uv"], oo["conic"], s["opacity"], feat_np, oo["idx_sorted"], oo["tile_range"], 0.0, W, H,
                                  f["final_T"], f["ncontrib"], g_np)
    for name, got in (("dL_duv", uv.grad), ("dL_dconic", conic.grad), ("dL_dopacity", op.grad), ("dL_dfeature", feat.grad)):
        Hh.assert_grad_close(n(got), b[name], name)


def test_rasterization_pipeline(case, cuda):
    s, _, _ = case
    gs = gs_mod()
    W, H = s["W"], s["H"]
    feat_np = np.random.default_rng(1).random((s["P"], 3), dtype=np.float32)
    img = gs.rasterization(t(s["xyz"], cuda), t(s["scaling"], cuda), t(s["rotation"], cuda), t(s["opacity"], cuda),
                           t(feat_np, cuda), t(s["intr"], cuda), t(s["extr"][:3], cuda), W, H, 1.0)
    ref, aux = O.rasterization(s["xyz"], s["scaling"], s["rotation"], s["opacity"], feat_np, s["intr"], s["extr"], W, H, 1.0)
    f = O.alpha_blending_forward(aux["uv"], aux["conic"], s["opacity"], feat_np, aux["idx_sorted"], aux["tile_range"], 1.0, W, H,
                                 frag_eps=Hh.FRAG_EPS)
    Hh.assert_pixels_close(n(img), ref, f["fragile"], "rasterization")


def test_empty_and_errors(cuda):
    gs = gs_mod()
    W, H = 40, 24
    z2 = torch.zeros(0, 2, device=cuda)
    tr = torch.zeros(6, 2, dtype=torch.int32, device=cuda)
    img = gs.alpha_blending(z2, torch.zeros(0, 3, device=cuda), torch.zeros(0, 1, device=cuda), torch.zeros(0, 3, device=cuda),
                            torch.zeros(0, dtype=torch.int32, device=cuda), tr, 0.75, W, H)
    assert img.shape == (3, H, W) and torch.all(img == 0.75)
    with pytest.raises(RuntimeError):
        gs.compute_cov3d(torch.ones(4, 3), torch.ones(4, 4))  # CPU tensors: no CPU path, like the reference
