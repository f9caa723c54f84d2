"""-m gpu: the fused image losses (csrc/loss.cu via splatter_a_video_b200.losses) against
(i) the committed outputs of the reference's own loss functions (tests/golden/golden_losses.npz) and
(ii) the CPU restatement oracle/loss_ref.py on seeded inputs up to the benchmark's 854x480 frames, plus size-independent
properties.  Tolerances: loss value 1e-5 relative (+2e-6 absolute), gradients 1e-3 of the largest gradient magnitude."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_ref as LR

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_losses.npz")
LOSS_RTOL, LOSS_ATOL, GRAD_RTOL = 1e-5, 2e-6, 1e-3


def _close_loss(got, want, what):
    assert abs(float(got) - float(want)) <= LOSS_ATOL + LOSS_RTOL * abs(float(want)), f"{what}: {float(got)} vs {float(want)}"


def _close_grad(got, want, what, rtol=GRAD_RTOL):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    err, scale = np.abs(got - want).max(), np.abs(want).max()
    assert err <= rtol * scale + 1e-12, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def _cuda_rgb(cuda, pred, gt, lam=0.2, weight=1.0):
    from splatter_a_video_b200 import losses
    p = torch.from_numpy(np.ascontiguousarray(pred)).to(cuda).requires_grad_(True)
    loss, l1, ssim = losses.rgb_loss(p, torch.from_numpy(np.ascontiguousarray(gt)).to(cuda), lam, weight)
    (g,) = torch.autograd.grad(loss, p)
    return float(loss), float(l1), float(ssim), g.cpu().numpy()


def _oracle_rgb(pred, gt, lam=0.2, weight=1.0):
    p = torch.from_numpy(np.ascontiguousarray(pred)).requires_grad_(True)
    loss, l1, ssim = LR.rgb_loss(p, torch.from_numpy(np.ascontiguousarray(gt)), lam, weight)
    (g,) = torch.autograd.grad(loss, p)
    return float(loss), float(l1), float(ssim), g.numpy()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rgb_loss_matches_reference_golden(cuda, tag):
    G = np.load(GOLD)
    loss, l1, ssim, g = _cuda_rgb(cuda, G[f"rgb_{tag}_pred"], G[f"rgb_{tag}_gt"])
    _close_loss(loss, G[f"rgb_{tag}_loss"], "loss"); _close_loss(l1, G[f"rgb_{tag}_l1"], "l1"); _close_loss(ssim, G[f"rgb_{tag}_ssim"], "ssim")
    _close_grad(g, G[f"rgb_{tag}_grad"], "dL/drgb")


@pytest.mark.parametrize("H,W", [(1, 1), (3, 7), (5, 128), (4, 129), (9, 300), (480, 854)])
def test_rgb_loss_matches_oracle(cuda, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    pred = rng.random((3, H, W), dtype=np.float32)
    gt = np.clip(pred.transpose(1, 2, 0) + 0.2 * rng.standard_normal((H, W, 3)).astype(np.float32), 0, 1).astype(np.float32)
    got, want = _cuda_rgb(cuda, pred, gt, 0.2, 50.0), _oracle_rgb(pred, gt, 0.2, 50.0)      # --loss_rgb_weight style scaling
    for a, b, what in zip(got[:3], want[:3], ["loss", "l1", "ssim"]):
        _close_loss(a, b, what)
    _close_grad(got[3], want[3], "dL/drgb")


def test_rgb_loss_identical_images_and_upstream_scale(cuda):
    from splatter_a_video_b200 import losses
    rng = np.random.default_rng(7)
    img = rng.random((3, 33, 200), dtype=np.float32)
    loss, l1, ssim, g = _cuda_rgb(cuda, img, np.ascontiguousarray(img.transpose(1, 2, 0)))
    assert abs(loss) <= 1e-6 and l1 == 0.0 and abs(ssim - 1.0) <= 1e-6
    assert np.abs(g).max() <= 1e-7                                   # SSIM is stationary at p == g and sign(0) == 0
    # autograd plumbing: d(3 * loss) = 3 * d(loss); evaluation without gradient skips the backward kernel
    p = torch.from_numpy(img).to(cuda).requires_grad_(True)
    gt = torch.from_numpy(np.ascontiguousarray(np.roll(img, 3, 2).transpose(1, 2, 0))).to(cuda)
    (g1,) = torch.autograd.grad(losses.rgb_loss(p, gt)[0], p)
    (g3,) = torch.autograd.grad(3.0 * losses.rgb_loss(p, gt)[0], p)
    assert torch.allclose(g3, 3.0 * g1, rtol=1e-6, atol=0)
    with torch.no_grad():
        l_eval = losses.rgb_loss(p, gt)[0]
    assert float(l_eval) == float(losses.rgb_loss(p, gt)[0])


def _cuda_depth(cuda, pred, gt, weight=1.0):
    from splatter_a_video_b200 import losses
    p = torch.from_numpy(pred).to(cuda).requires_grad_(True)
    loss = losses.depth_loss_dpt(p, torch.from_numpy(gt).to(cuda), weight)
    (g,) = torch.autograd.grad(loss, p)
    return float(loss), g.cpu().numpy()


def _oracle_depth(pred, gt, weight=1.0):
    p = torch.from_numpy(pred).requires_grad_(True)
    loss = LR.depth_loss_dpt(p, torch.from_numpy(gt), weight)
    (g,) = torch.autograd.grad(loss, p)
    return float(loss), g.numpy()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_depth_loss_matches_reference_golden(cuda, tag):
    G = np.load(GOLD)
    loss, g = _cuda_depth(cuda, G[f"depth_{tag}_pred"], G[f"depth_{tag}_gt"])
    _close_loss(loss, G[f"depth_{tag}_loss"], "depth loss")
    _close_grad(g, G[f"depth_{tag}_grad"], "dL/ddepth")


@pytest.fixture(params=["fused", "staged"])
def depth_path(request):
    """spv_loss_depth_dpt as one kernel with device-side barriers (images up to ~600 k pixels) or as the chain of six kernels
    larger images fall back to; both must match the oracle."""
    from splatter_a_video_b200 import _lib as L
    L.set_option("depth_staged", 1 if request.param == "staged" else 0)
    yield request.param
    L.set_option("depth_staged", 0)


@pytest.mark.parametrize("H,W,bg_frac", [(1, 2, 0.0), (1, 3, 0.0), (17, 23, 0.0), (64, 64, 0.6), (480, 854, 0.0), (480, 854, 0.55),
                                         (1080, 1920, 0.3)])
def test_depth_loss_matches_oracle(cuda, depth_path, H, W, bg_frac):
    """bg_frac > 0.5 puts the median inside the tie group of background pixels (rendered depth == bg == 1.0 exactly)."""
    rng = np.random.default_rng(H + W)
    pred = (0.5 + 1.5 * rng.random((H, W, 1))).astype(np.float32)
    gt = (0.3 + 2.0 * rng.random((H, W, 1)) + 0.5 * pred).astype(np.float32)
    if bg_frac > 0:
        pred[rng.random((H, W, 1)) < bg_frac] = 1.0
    got, want = _cuda_depth(cuda, pred, gt, 2.0), _oracle_depth(pred, gt, 2.0)
    _close_loss(got[0], want[0], "depth loss")
    _close_grad(got[1], want[1], "dL/ddepth")
    # invariances of the loss (size-independent): shifting or scaling the prediction leaves it unchanged
    g = got[1].astype(np.float64).reshape(-1)
    assert abs(g.sum()) <= 1e-4 * np.abs(g).sum()
    assert abs((g * pred.reshape(-1)).sum()) <= 1e-4 * np.abs(g * pred.reshape(-1)).sum()


def _track_inputs(rng, H, W, n, vis_frac=0.8, raster=True):
    if raster:
        flat = np.sort(rng.choice(H * W, size=n, replace=False))
    else:
        flat = rng.integers(0, H * W, size=n)                           # unordered, with repeats
    query = np.stack([flat % W, flat // W], -1).astype(np.int32)
    img = (rng.random((3, H, W)) * 2 - 1).astype(np.float32)
    target = np.stack([rng.random(n) * W, rng.random(n) * H], -1).astype(np.float32)
    visible = rng.random(n) < vis_frac
    weights = rng.random(n).astype(np.float32)
    return img, query, target, visible, weights


def _cuda_track(cuda, img, query, target, visible, weights, q=0.98):
    from splatter_a_video_b200 import losses
    t = torch.from_numpy(img).to(cuda).requires_grad_(True)
    loss = losses.track_loss(t, torch.from_numpy(query).to(cuda), torch.from_numpy(target).to(cuda), torch.from_numpy(visible).to(cuda),
                             torch.from_numpy(weights).to(cuda), q)
    (g,) = torch.autograd.grad(loss, t, allow_unused=True)
    return float(loss), (g.cpu().numpy() if g is not None else np.zeros_like(img))


def _oracle_track(img, query, target, visible, weights, q=0.98):
    t = torch.from_numpy(img).requires_grad_(True)
    loss = LR.track_loss(t, torch.from_numpy(query), torch.from_numpy(target), torch.from_numpy(visible), torch.from_numpy(weights), q)
    (g,) = torch.autograd.grad(loss, t)
    return float(loss), g.numpy()


def test_track_loss_matches_reference_golden(cuda):
    G = np.load(GOLD)
    loss, g = _cuda_track(cuda, G["track_img"], G["track_query"], G["track_target"], G["track_visible"], G["track_weights"])
    _close_loss(loss, G["track_loss"], "track loss")
    _close_grad(g, G["track_grad"], "dL/dtrack", rtol=1e-5)


@pytest.mark.parametrize("H,W,n,vis,raster", [(24, 36, 1, 1.0, True), (24, 36, 2, 1.0, True), (48, 64, 300, 0.7, True),
                                              (480, 854, 5000, 0.8, True), (48, 64, 500, 0.9, False)])
def test_track_loss_matches_oracle(cuda, H, W, n, vis, raster):
    rng = np.random.default_rng(n)
    args = _track_inputs(rng, H, W, n, vis, raster)
    got, want = _cuda_track(cuda, *args), _oracle_track(*args)
    _close_loss(got[0], want[0], "track loss")
    _close_grad(got[1], want[1], "dL/dtrack", rtol=1e-5)


def test_track_loss_without_visible_points_is_zero(cuda):
    rng = np.random.default_rng(3)
    img, query, target, visible, weights = _track_inputs(rng, 20, 30, 40)
    loss, g = _cuda_track(cuda, img, query, target, np.zeros_like(visible), weights)
    assert loss == 0.0 and np.count_nonzero(g) == 0                      # trainer_fragGS.py:569-570


def test_losses_reject_cpu_tensors():
    from splatter_a_video_b200 import losses
    with pytest.raises(RuntimeError):
        losses.rgb_loss(torch.zeros(3, 4, 4), torch.zeros(4, 4, 3))
    with pytest.raises(RuntimeError):
        losses.depth_loss_dpt(torch.zeros(4, 4), torch.ones(4, 4))


def test_loss_timing_report(cuda):
    """Not a parity test: times the three fused losses (value + gradient) at the benchmark frame size against the same
    expressions written with torch ops on the same GPU (what the reference trainer executes), and leaves the numbers in
    gpurun_out/loss_timing.json for profiles/."""
    import json
    from splatter_a_video_b200 import losses
    H, W, n = 480, 854, 4096
    rng = np.random.default_rng(0)
    rgb = torch.from_numpy(rng.random((3, H, W), dtype=np.float32)).to(cuda).requires_grad_(True)
    gt = torch.from_numpy(rng.random((H, W, 3), dtype=np.float32)).to(cuda)
    depth = torch.from_numpy((0.5 + rng.random((1, H, W))).astype(np.float32)).to(cuda).requires_grad_(True)
    gt_depth = torch.from_numpy((0.5 + rng.random((1, H, W))).astype(np.float32)).to(cuda)
    img, query, target, visible, weights = _track_inputs(rng, H, W, n)
    track = torch.from_numpy(img).to(cuda).requires_grad_(True)
    targs = [torch.from_numpy(a).to(cuda) for a in (query, target, visible, weights)]

    def ours():
        total = (losses.rgb_loss(rgb, gt, 0.2, 10.0)[0] + losses.depth_loss_dpt(depth, gt_depth) + losses.track_loss(track, *targs, 0.98, 2.0))
        return torch.autograd.grad(total, [rgb, depth, track])

    def torch_ops():
        total = (LR.rgb_loss(rgb, gt, 0.2, 10.0)[0] + LR.depth_loss_dpt(depth, gt_depth) + LR.track_loss(track, *targs, 0.98, 2.0))
        return torch.autograd.grad(total, [rgb, depth, track])

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    a, b = ours(), torch_ops()
    for x, y, what in zip(a, b, ["rgb", "depth", "track"]):
        _close_grad(x.cpu().numpy(), y.cpu().numpy(), f"timing inputs: dL/d{what}")
    res = {"frame": f"{W}x{H}", "track_points": n, "fused_ms": timed(ours), "torch_ops_ms": timed(torch_ops)}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/loss_timing.json", "w") as f:
        json.dump(res, f)
    print("loss timing:", res)
    assert res["fused_ms"] < res["torch_ops_ms"]
