"""-m gpu: the fused per-frame path (one C call forward, one backward, no host sync, exact tile culling) against the staged
`dptr.gs` op sequence (already pinned to the oracle / reference by the other suites), the deformation op against torch,
and CUDA-graph replay of a whole step against eager execution."""
import numpy as np
import pytest
import torch

import helpers as Hh
from helpers import n
from splatter_a_video_b200 import synth

pytestmark = pytest.mark.gpu

ATTRS = ["track_gs", "mask_attribute", "pos_poly_feat", "dino_attribute"]
LEAVES = ["position", "opacity", "scaling", "rotation", "shs", "track_gs", "mask_attribute", "pos_poly_feat", "dino_attribute"]


def _batch(sc, dev, K=20):
    return {"height": sc.H, "width": sc.W, "extrinsic_matrix": sc.extr.to(dev), "intrinsic_matrix": sc.intr.to(dev),
            "camera_center": torch.zeros(3, device=dev), "render_attributes_list": list(ATTRS), "num_idx": K}


def _rd(sc, dev, frame=0):
    d = {"position": sc.frame_position(frame), "opacity": sc.opacity, "scaling": sc.scaling, "rotation": sc.rotation, "shs": sc.shs,
         "track_gs": sc.frame_position(frame + 1), "mask_attribute": sc.attrs["mask_attribute"],
         "pos_poly_feat": sc.attrs["pos_poly_feat"], "dino_attribute": sc.attrs["dino_attribute"]}
    return {k: v.to(dev).clone().requires_grad_(True) for k, v in d.items()}


def _run(rnd, sc, dev, gimgs):
    rd = _rd(sc, dev)
    out = rnd.render_batch(rd, [_batch(sc, dev)])
    keys = ["rgb", "depth"] + ATTRS
    torch.autograd.backward([out[k][0] for k in keys], [gimgs[k] for k in keys])
    return out, rd


@pytest.mark.parametrize("P,W,H", [(1000, 64, 64), (40_000, 333, 250)])
@pytest.mark.parametrize("cull", [False, True])
def test_frame_path_equals_staged_ops(cuda, P, W, H, cull):
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_scene(P, 6, W, H, seed=3)
    g = torch.Generator().manual_seed(1)
    chans = {"rgb": 3, "depth": 1, "track_gs": 3, "mask_attribute": 1, "pos_poly_feat": 12, "dino_attribute": 3}
    gimgs = {k: torch.randn(c, H, W, generator=g).to(cuda) for k, c in chans.items()}
    staged = parse_renderer({"name": "DPTROrthoEnhancedRender"}, white_bg=False, device=cuda)
    frame = parse_renderer({"name": "DPTROrthoEnhancedRenderB200", "cull": cull}, white_bg=False, device=cuda)
    o1, r1 = _run(staged, sc, cuda, gimgs)
    o2, r2 = _run(frame, sc, cuda, gimgs)
    for k in ["rgb", "depth"] + ATTRS:
        assert float((o1[k] - o2[k]).abs().max()) <= 1e-6, k     # same kernels, shorter lists: images must not move
    assert torch.equal(o1["gs_idx"], o2["gs_idx"]) and torch.equal(o1["radii"], o2["radii"]) and torch.equal(o1["visibility"], o2["visibility"])
    for k in LEAVES:
        Hh.assert_grad_close(n(r2[k].grad), n(r1[k].grad), f"frame d/d{k}", norm_tol=2e-5)
    Hh.assert_grad_close(n(o2["viewspace_points"][0].grad), n(o1["viewspace_points"][0].grad), "ndc.grad", norm_tol=2e-5)
    st = frame.last_status.cpu()
    assert int(st[1]) == 0
    if cull:
        from splatter_a_video_b200 import gs
        assert int(st[0]) < 0.95 * int(frame.capacity.I_cap)   # lists got shorter than the un-culled count + slack


def test_capacity_overflow_recovers(cuda):
    from splatter_a_video_b200.gs.frame import Capacity
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_scene(5000, 4, 128, 96, seed=9)
    ref = parse_renderer({"name": "DPTROrthoEnhancedRender"}, white_bg=False, device=cuda)
    small = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    small.capacity = Capacity(initial=100)          # far too small: must grow and re-render
    rd = {k: v.detach() for k, v in _rd(sc, cuda).items()}
    o1 = ref.render_batch(dict(rd), [_batch(sc, cuda)])
    o2 = small.render_batch(dict(rd), [_batch(sc, cuda)])
    assert small.capacity.I_cap > 100
    assert float((o1["rgb"] - o2["rgb"]).abs().max()) <= 1e-6 and torch.equal(o1["gs_idx"], o2["gs_idx"])


def test_deform_position_matches_reference_formula(cuda):
    from splatter_a_video_b200.gs.frame import deform_position, spline_interval
    P, T = 3000, 50
    NI = -(-T // 5)
    g = torch.Generator().manual_seed(4)
    base = torch.randn(P, 3, generator=g)
    node = 0.1 * torch.randn(P, 4 * NI * 3, generator=g)
    for time in (0, 1, 7, 24, 25, 49):
        idx, dist = spline_interval(time, T, NI)
        # the reference formula on CPU (dynamic_gaussian_with_base_point_cloud.py:239-250)
        cpu_node = node.clone().requires_grad_(True)
        coeff = cpu_node.reshape(-1, 4, NI, 3)
        want = coeff[:, 3, idx] + coeff[:, 2, idx] * dist + coeff[:, 1, idx] * dist ** 2 + coeff[:, 0, idx] * dist ** 3 + base
        gnode = node.to(cuda).requires_grad_(True)
        got = deform_position(base.to(cuda), gnode, torch.tensor([idx], dtype=torch.int32, device=cuda),
                              torch.tensor([dist], dtype=torch.float32, device=cuda), NI)
        np.testing.assert_allclose(n(got), want.detach().numpy(), rtol=1e-6, atol=1e-6)
        gp = torch.randn(P, 3, generator=g)
        want.backward(gp); got.backward(gp.to(cuda))
        np.testing.assert_allclose(n(gnode.grad), cpu_node.grad.numpy(), rtol=1e-6, atol=1e-7)


def test_cuda_graph_replay_of_a_full_step(cuda):
    """No host synchronisation on the frame path => a whole forward+backward captures into a CUDA graph and replays for
    other frames by updating two device scalars."""
    from splatter_a_video_b200.gs.frame import deform_position, spline_interval
    from splatter_a_video_b200.renderer import parse_renderer
    P, W, H, T = 8000, 160, 128, 20
    NI = -(-T // 5)
    sc = synth.make_scene(P, T, W, H, seed=2).to(cuda)
    g = torch.Generator().manual_seed(7)
    node = (0.02 * torch.randn(P, 4 * NI * 3, generator=g)).to(cuda).requires_grad_(True)
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("opacity", "scaling", "rotation", "shs")}
    idx_dev = torch.zeros(1, dtype=torch.int32, device=cuda); dist_dev = torch.zeros(1, device=cuda)
    gimg = torch.randn(4, H, W, generator=g).to(cuda)
    rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    batch = {"height": H, "width": W, "extrinsic_matrix": sc.extr, "intrinsic_matrix": sc.intr, "camera_center": torch.zeros(3, device=cuda),
             "num_idx": 8}

    def step():
        for p in [node, *leaves.values()]:
            p.grad = None
        pos = deform_position(sc.position, node, idx_dev, dist_dev, NI)
        out = rnd.render_batch({"position": pos, **leaves}, [dict(batch)])
        torch.autograd.backward([out["rgb"][0], out["depth"][0]], [gimg[:3], gimg[3:]])
        return out["rgb"]

    def set_frame(t):
        i, d = spline_interval(t, T, NI)
        idx_dev.fill_(i); dist_dev.fill_(d)

    set_frame(3)
    step()                                   # settles the capacity (one sync, outside the graph)
    rnd.capacity.I_cap = int(rnd.capacity.I_cap * 1.5)
    rnd.observe_capacity = False
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        img_static = step()
    grads_static = {"node": node.grad, **{k: v.grad for k, v in leaves.items()}}
    for t in (11, 0, 19):
        set_frame(t)
        graph.replay()
        torch.cuda.synchronize()
        got_img = img_static.clone(); got = {k: v.clone() for k, v in grads_static.items()}
        assert int(rnd.last_status.cpu()[1]) == 0
        eager_img = step().clone()
        want = {"node": node.grad.clone(), **{k: v.grad.clone() for k, v in leaves.items()}}
        assert float((got_img - eager_img).abs().max()) <= 1e-6
        for k in want:
            Hh.assert_grad_close(n(got[k]), n(want[k]), f"graph replay d/d{k}", norm_tol=2e-5)


@pytest.mark.parametrize("device_clock", [False, True])
def test_flat_adam_matches_torch_adam(cuda, device_clock):
    """device_clock: step counter, bias corrections and learning rates in device memory (the CUDA-graph capturable entry)."""
    from splatter_a_video_b200.parallel import FlatAdam, FlatParams
    g = torch.Generator().manual_seed(0)
    tensors = {"a": torch.randn(1000, 3, generator=g), "b": torch.randn(1000, 16, 3, generator=g), "c": torch.randn(1001, 1, generator=g)}
    lrs = {"a": 1e-3, "b": 2.5e-3, "c": 5e-2}
    flat = FlatParams({k: v.to(cuda) for k, v in tensors.items()})
    opt = FlatAdam(flat, lrs, eps=1e-15, device_clock=device_clock)
    ref = {k: v.clone().to(cuda).requires_grad_(True) for k, v in tensors.items()}
    topt = torch.optim.Adam([{"params": [ref[k]], "lr": lrs[k]} for k in ref], eps=1e-15)
    for it in range(5):
        for k in tensors:
            gk = torch.randn(tensors[k].shape, generator=g).to(cuda)
            flat[k].grad.copy_(gk); ref[k].grad = gk.clone()
        opt.step(); topt.step()
        for k in tensors:
            np.testing.assert_allclose(n(flat[k]), n(ref[k]), rtol=2e-6, atol=1e-7)


def test_grad_sinks_write_in_place(cuda):
    """Frame path with gradient sinks: gradients land in the caller's flat buffer slices, equal to the autograd ones."""
    from splatter_a_video_b200.parallel import FlatParams
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_scene(6000, 4, 160, 96, seed=12)
    g = torch.Generator().manual_seed(3)
    gimg = torch.randn(4, sc.H, sc.W, generator=g).to(cuda)
    batch = {"height": sc.H, "width": sc.W, "extrinsic_matrix": sc.extr.to(cuda), "intrinsic_matrix": sc.intr.to(cuda),
             "camera_center": torch.zeros(3, device=cuda), "num_idx": 8}
    names = ["scaling", "rotation", "opacity", "shs"]

    def run(use_sinks):
        flat = FlatParams({k: getattr(sc, k).to(cuda) for k in ["position"] + names})
        flat.flat_grad.fill_(123.0)                      # stale content must be overwritten, not accumulated
        flat["position"].grad.zero_()
        rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
        b = dict(batch)
        if use_sinks:
            b["grad_sinks"] = flat.grad_sinks(names)
        else:
            flat.zero_grad()
        out = rnd.render_batch({k: flat[k] for k in ["position"] + names}, [b])
        torch.autograd.backward([out["rgb"][0], out["depth"][0]], [gimg[:3], gimg[3:]])
        return flat.flat_grad.clone()

    a, b = run(False), run(True)
    Hh.assert_grad_close(n(b), n(a), "sinks vs autograd accumulation", norm_tol=2e-5)


def test_attribute_gradient_sinks_by_name(cuda):
    """grad_sinks keyed by render_dict names: attribute tensors' gradients are written into the caller's buffers (stale content
    overwritten), the autograd leaves get none, values equal the autograd path."""
    from splatter_a_video_b200.renderer import parse_renderer
    P, W, H = 8000, 160, 96
    sc = synth.make_scene(P, 6, W, H, seed=23)
    g = torch.Generator().manual_seed(4)
    chans = {"rgb": 3, "depth": 1, "track_gs": 3, "mask_attribute": 1, "pos_poly_feat": 12, "dino_attribute": 3}
    gimgs = {k: torch.randn(c, H, W, generator=g).to(cuda) for k, c in chans.items()}
    keys = ["rgb", "depth"] + ATTRS

    def run(sinks):
        rd = _rd(sc, cuda)
        rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
        b = _batch(sc, cuda)
        if sinks:
            b["grad_sinks"] = sinks
        out = rnd.render_batch(rd, [b])
        torch.autograd.backward([out[k][0] for k in keys], [gimgs[k] for k in keys])
        return rd

    want = run(None)
    sinks = {k: torch.full((P, chans[k]), 7.0, device=cuda) for k in ("mask_attribute", "dino_attribute")}
    got = run(sinks)
    for k in sinks:
        assert got[k].grad is None
        Hh.assert_grad_close(n(sinks[k]), n(want[k].grad), f"sink {k}", norm_tol=2e-5)
    for k in ("position", "track_gs", "pos_poly_feat", "scaling"):
        Hh.assert_grad_close(n(got[k].grad), n(want[k].grad), f"d/d{k}", norm_tol=2e-5)


def test_deform_rotation_matches_reference_formula(cuda):
    from splatter_a_video_b200.gs.frame import deform_rotation, rotation_basis
    P = 2000
    g = torch.Generator().manual_seed(8)
    rot = torch.randn(P, 4, generator=g); poly = 0.1 * torch.randn(P, 16, generator=g); four = 0.1 * torch.randn(P, 32, generator=g)
    for time in (0, 13, 49):
        normed = (time - 0) / 49
        # reference (dynamic_gaussian_with_base_point_cloud.py:184-198) on CPU
        r = rot.clone().requires_grad_(True)
        pb = torch.pow(torch.tensor(normed), torch.arange(4).float())[None, :, None]
        k = torch.arange(4).float() + 1
        fb = torch.cat([torch.cos(normed * k * np.pi), torch.sin(normed * k * np.pi)], 0)[None, :, None]
        want = torch.nn.functional.normalize(r + torch.sum(poly.reshape(P, 4, 4) * pb, dim=1).detach()
                                             + torch.sum(four.reshape(P, 8, 4) * fb, dim=1).detach())
        gr = rot.to(cuda).requires_grad_(True)
        got = deform_rotation(gr, poly.to(cuda), four.to(cuda), rotation_basis(time, 0, 49).to(cuda))
        np.testing.assert_allclose(n(got), want.detach().numpy(), rtol=1e-5, atol=1e-6)
        go = torch.randn(P, 4, generator=g)
        want.backward(go); got.backward(go.to(cuda))
        np.testing.assert_allclose(n(gr.grad), r.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_frame_path_with_gradient_free_attribute_groups(cuda):
    """The trainer's real situation: track_gs is detached and pos_poly_feat frozen, only mask/dino (4 channels) carry a
    gradient -> the backward reduces 8 feature channels (16-wide network); results must equal the staged ops."""
    from splatter_a_video_b200.renderer import parse_renderer
    P, W, H = 20_000, 256, 160
    sc = synth.make_scene(P, 6, W, H, seed=17)
    g = torch.Generator().manual_seed(2)
    chans = {"rgb": 3, "depth": 1, "track_gs": 3, "mask_attribute": 1, "pos_poly_feat": 12, "dino_attribute": 3}
    gimgs = {k: torch.randn(c, H, W, generator=g).to(cuda) for k, c in chans.items()}
    frozen = ("track_gs", "pos_poly_feat")

    def run(name):
        rd = _rd(sc, cuda)
        for k in frozen:
            rd[k] = rd[k].detach()
        rnd = parse_renderer({"name": name}, white_bg=False, device=cuda)
        out = rnd.render_batch(rd, [_batch(sc, cuda)])
        keys = ["rgb", "depth"] + ATTRS
        torch.autograd.backward([out[k][0] for k in keys], [gimgs[k] for k in keys])
        return out, rd

    o1, r1 = run("DPTROrthoEnhancedRender")
    o2, r2 = run("DPTROrthoEnhancedRenderB200")
    for k in ["rgb", "depth"] + ATTRS:
        assert float((o1[k] - o2[k]).abs().max()) <= 1e-6, k
    for k in LEAVES:
        if k in frozen:
            assert r2[k].grad is None
        else:
            Hh.assert_grad_close(n(r2[k].grad), n(r1[k].grad), f"d/d{k}", norm_tol=2e-5)
    Hh.assert_grad_close(n(o2["viewspace_points"][0].grad), n(o1["viewspace_points"][0].grad), "ndc.grad", norm_tol=2e-5)


def test_kernel_timer_brackets_the_blend_kernels(cuda):
    """spv_kernel_timer_*: off by default; when on, one frame step leaves a positive device time for both blend kernels that
    is smaller than the event-timed step around it."""
    import ctypes
    from splatter_a_video_b200 import _lib as L
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_scene(20_000, 4, 256, 192, seed=5)
    g = torch.Generator().manual_seed(2)
    chans = {"rgb": 3, "depth": 1, "track_gs": 3, "mask_attribute": 1, "pos_poly_feat": 12, "dino_attribute": 3}
    gimgs = {k: torch.randn(c, 192, 256, generator=g).to(cuda) for k, c in chans.items()}
    rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    _run(rnd, sc, cuda, gimgs)
    ms = ctypes.c_float()
    with pytest.raises(RuntimeError):
        L.call("spv_kernel_timer_read", 0, ctypes.byref(ms))       # never enabled: no events exist
    L.call("spv_kernel_timer_enable", 1)
    try:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _run(rnd, sc, cuda, gimgs); e1.record()
        torch.cuda.synchronize()
        got = []
        for slot in (0, 1):
            L.call("spv_kernel_timer_read", slot, ctypes.byref(ms))
            got.append(ms.value)
        assert 0.0 < got[0] < e0.elapsed_time(e1) and 0.0 < got[1] < e0.elapsed_time(e1)
    finally:
        L.call("spv_kernel_timer_enable", 0)


def _bin(name, uv, depth, radius, conic, opacity, cull, W, H, I_cap):
    """Direct C-ABI call of one of the two capacity-bounded binning entry points."""
    from splatter_a_video_b200 import _lib as L
    P, dev = uv.shape[0], uv.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    idx = torch.full((I_cap,), -7, dtype=torch.int32, device=dev)
    tr = torch.empty(T, 2, dtype=torch.int32, device=dev)
    st = torch.empty(2, dtype=torch.int32, device=dev)
    nb = (L.query("spv_bin_tiles_workspace_bytes", P, I_cap, W, H) if name == "spv_bin_tiles"
          else L.query("spv_bin_capacity_workspace_bytes", P, I_cap))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    L.call(name, P, I_cap, L.ptr(uv), L.ptr(depth), L.ptr(radius), L.ptr(conic), L.ptr(opacity), int(cull), W, H, L.ptr(idx),
           L.ptr(tr), L.ptr(st), L.ptr(ws), nb, L.stream())
    torch.cuda.synchronize()
    return idx, tr, st.cpu()


# (P, W, H, scale multiplier): the last two make every tile segment exceed the small-sort limit (200 KB shared-memory sort) and
# 25600 keys (in-place global sort) respectively
@pytest.mark.parametrize("P,W,H,big", [(3000, 96, 80, 1.0), (60_000, 333, 250, 1.0), (12_000, 48, 32, 6.0), (150_000, 32, 32, 8.0)])
@pytest.mark.parametrize("cull", [0, 1])
def test_tile_segment_binning_equals_radix_binning(cuda, P, W, H, big, cull):
    """spv_bin_tiles (per-tile histogram + scan + scatter + per-tile shared-memory sort) must give bit-identical idx_sorted /
    tile_range / status to spv_bin_capacity (global radix sort), and without culling to gs.sort_gaussian."""
    from splatter_a_video_b200 import gs
    sc = synth.make_scene(P, 2, W, H, seed=11)
    pos, scaling = sc.frame_position(0).to(cuda), (sc.scaling * big).to(cuda)
    # duplicate depths inside tiles: ties must fall back to ascending Gaussian id
    pos[1::7, 2] = pos[0::7, 2][: pos[1::7].shape[0]]
    uv, depth = gs.project_point_ortho(pos, sc.extr.to(cuda), W, H, 0.01)
    vis = depth != 0
    cov3d = gs.compute_cov3d(scaling, sc.rotation.to(cuda), vis)
    conic, radius, tiles = gs.ewa_project_ortho(cov3d, sc.extr.to(cuda), uv, W, H, vis.squeeze(-1))
    opacity = sc.opacity.to(cuda)
    total = int(tiles.sum())
    I_cap = total + 1000
    a_idx, a_tr, a_st = _bin("spv_bin_capacity", uv, depth, radius, conic, opacity, cull, W, H, I_cap)
    b_idx, b_tr, b_st = _bin("spv_bin_tiles", uv, depth, radius, conic, opacity, cull, W, H, I_cap)
    I = int(a_st[0])
    assert int(b_st[0]) == I and int(a_st[1]) == 0 and int(b_st[1]) == 0
    assert torch.equal(a_tr, b_tr)
    assert torch.equal(a_idx[:I], b_idx[:I])
    if big > 1:
        assert int((a_tr[:, 1] - a_tr[:, 0]).max()) > (25600 if big >= 8 else 4096)
    if not cull:
        assert I == total
        r_idx, r_tr = gs.sort_gaussian(uv, depth, W, H, radius, tiles)
        assert torch.equal(r_idx, b_idx[:I]) and torch.equal(r_tr, b_tr)
    # overflow: flagged, nothing written out of bounds, ranges clipped to the capacity
    small = max(I // 3, 1)
    c_idx, c_tr, c_st = _bin("spv_bin_tiles", uv, depth, radius, conic, opacity, cull, W, H, small)
    assert int(c_st[1]) == 1 and int(c_st[0]) == small and int(c_tr.max()) <= small


def test_frame_backward_twice_and_gradient_free_images(cuda):
    """(a) a second backward over the same graph (retain_graph) must reproduce the first -- the packed gradient rows are cleared
    by the forward call for the first backward only; (b) rendered images that receive NO upstream gradient (the trainer puts no
    loss on its mask / dino / pos_poly_feat images) are pruned from the backward traversal: same gradients as explicit zeros."""
    from splatter_a_video_b200.gs.frame import render_ortho_frame
    sc = synth.make_scene(20_000, 4, 256, 192, seed=21)
    W, H, P = sc.W, sc.H, sc.P
    g = torch.Generator().manual_seed(4)
    g_rgb, g_depth, g_track = (torch.randn(c, H, W, generator=g).to(cuda) for c in (3, 1, 3))

    def leaves():
        d = {"position": sc.frame_position(0), "scaling": sc.scaling, "rotation": sc.rotation, "opacity": sc.opacity, "shs": sc.shs,
             "track": sc.frame_position(1), "mask": sc.attrs["mask_attribute"], "poly": sc.attrs["pos_poly_feat"]}
        return {k: v.to(cuda).clone().requires_grad_(k != "poly") for k, v in d.items()}

    def render(L_):
        imgs, _, _, status = render_ortho_frame(L_["position"], L_["scaling"], L_["rotation"], L_["opacity"], L_["shs"],
                                                [L_["track"], L_["mask"], L_["poly"]], sc.extr.to(cuda), W, H, 20, 0.0, 8 * P)
        assert int(status.cpu()[1]) == 0
        return imgs      # [rgb, depth, track, mask, poly]

    # (a)
    A_ = leaves()
    imgs = render(A_)
    outs, grads = [imgs[0], imgs[1], imgs[2]], [g_rgb, g_depth, g_track]
    first = torch.autograd.grad(outs, [A_[k] for k in ("position", "opacity", "shs", "track")], grads, retain_graph=True)
    second = torch.autograd.grad(outs, [A_[k] for k in ("position", "opacity", "shs", "track")], grads)
    for a, b in zip(first, second):
        Hh.assert_grad_close(n(b), n(a), "second backward", norm_tol=2e-6)
    # (b) no gradient on the mask / poly images  ==  explicit zero gradients on them
    B_ = leaves()
    imgs = render(B_)
    zeros = [torch.zeros_like(imgs[3]), torch.zeros_like(imgs[4])]
    ref = torch.autograd.grad(imgs, [B_[k] for k in ("position", "scaling", "opacity", "shs", "track", "mask")], grads + zeros)
    C_ = leaves()
    imgs = render(C_)
    got = torch.autograd.grad([imgs[0], imgs[1], imgs[2]], [C_[k] for k in ("position", "scaling", "opacity", "shs", "track", "mask")],
                              grads, allow_unused=True)
    for name, a, b in zip(("position", "scaling", "opacity", "shs", "track", "mask"), ref, got):
        if b is None:
            assert float(a.abs().max()) == 0.0, name
        elif name == "mask":
            assert float(b.abs().max()) == 0.0 and float(a.abs().max()) == 0.0
        else:
            Hh.assert_grad_close(n(b), n(a), f"pruned d/d{name}", norm_tol=2e-6)


@pytest.mark.parametrize("grad_groups", ["mask+dino", "track+mask+dino", "track+mask+dino+extra"])   # 8 / 11 / 13 feature-gradient channels
def test_backward_with_and_without_the_abs_pair(cuda, grad_groups):
    """The |RGB-pass dL_duv| pair is only reduced when the caller asks for it (abs_ndc given): without it the RGB-pass pair
    rides in its network slots and the second reduction network shrinks (blend_rec_bwd_kernel<CH,CG,ABS>: CG = 8 / 12 / 14).
    Every other gradient must be the same either way, and equal to the staged ops' (round 1's kernels)."""
    from splatter_a_video_b200.gs.frame import render_ortho_frame
    from splatter_a_video_b200 import gs
    from splatter_a_video_b200.gs import fused as F_
    sc = synth.make_scene(40_000, 4, 333, 250, seed=12)
    W, H, P = sc.W, sc.H, sc.P
    g = torch.Generator().manual_seed(6)
    extra = torch.rand(P, 2, generator=g)
    chans = (3, 1, 3, 1, 10, 3, 2)
    gimg = [torch.randn(c, H, W, generator=g).to(cuda) for c in chans]
    need = {"track": "track" in grad_groups, "mask": True, "poly": False, "dino": True, "extra": "extra" in grad_groups}
    names = ("position", "scaling", "rotation", "opacity", "shs", "track", "mask", "dino", "extra")

    def leaves():
        d = {"position": sc.frame_position(0), "scaling": sc.scaling, "rotation": sc.rotation, "opacity": sc.opacity, "shs": sc.shs,
             "track": sc.frame_position(1), "mask": sc.attrs["mask_attribute"], "poly": sc.attrs["pos_poly_feat"][:, :10].contiguous(),
             "dino": sc.attrs["dino_attribute"], "extra": extra}
        return {k: v.to(cuda).clone().requires_grad_(need.get(k, True)) for k, v in d.items()}

    def run(with_abs):
        L_ = leaves()
        ndc = torch.zeros(P, 2, device=cuda, requires_grad=True)
        abs_ndc = torch.zeros(P, 2, device=cuda, requires_grad=True) if with_abs else None
        imgs, _, _, status = render_ortho_frame(L_["position"], L_["scaling"], L_["rotation"], L_["opacity"], L_["shs"],
                                                [L_["track"], L_["mask"], L_["poly"], L_["dino"], L_["extra"]], sc.extr.to(cuda), W, H, 20, 0.0,
                                                8 * P, ndc=ndc, abs_ndc=abs_ndc)
        assert int(status.cpu()[1]) == 0
        torch.autograd.backward(imgs, gimg)
        out = {k: L_[k].grad for k in names if L_[k].requires_grad}
        out["ndc"] = ndc.grad
        if with_abs:
            out["abs_ndc"] = abs_ndc.grad
        return out

    with_abs, without = run(True), run(False)
    assert float(with_abs["abs_ndc"].abs().sum()) > 0
    for k in without:
        Hh.assert_grad_close(n(without[k]), n(with_abs[k]), f"{grad_groups}: no-abs d/d{k}", norm_tol=2e-5)
    # the staged single-traversal ops (blend.cu) on the same frame
    L_ = leaves()
    dirs = torch.zeros(P, 3, device=cuda); dirs[:, 2] = 1
    rgb = gs.compute_sh(L_["shs"], 3, dirs)
    uv, depth = gs.project_point_ortho(L_["position"], sc.extr.to(cuda), W, H, 0.01)
    vis = depth != 0
    cov3d = gs.compute_cov3d(L_["scaling"], L_["rotation"], vis)
    conic, radius, tiles = gs.ewa_project_ortho(cov3d, sc.extr.to(cuda), uv, W, H, vis.squeeze(-1))
    idx, tr = gs.sort_gaussian(uv, depth, W, H, radius, tiles)
    ndc = torch.zeros(P, 2, device=cuda, requires_grad=True); abs_ndc = torch.zeros(P, 2, device=cuda, requires_grad=True)
    attrs = torch.cat([L_["track"], L_["mask"], L_["poly"], L_["dino"], L_["extra"]], 1)
    img, dimg, aimg, _ = F_.blend_rgb_depth_attrs(uv, conic, L_["opacity"], rgb, depth, attrs, idx, tr, 0.0, W, H, ndc, abs_ndc, K=20)
    torch.autograd.backward([img, dimg, aimg], [gimg[0], gimg[1], torch.cat(gimg[2:], 0)])
    for k in names:
        if L_[k].requires_grad:
            Hh.assert_grad_close(n(with_abs[k]), n(L_[k].grad), f"{grad_groups}: frame vs staged d/d{k}", norm_tol=2e-5)
    Hh.assert_grad_close(n(with_abs["ndc"]), n(ndc.grad), "ndc", norm_tol=2e-5)
    Hh.assert_grad_close(n(with_abs["abs_ndc"]), n(abs_ndc.grad), "abs_ndc", norm_tol=2e-5)


def test_capacity_overflow_after_the_first_frame_is_reported(cuda):
    """An overflow on a LATER frame (the capacity was settled on an easier one) must not pass silently: the frame that follows
    raises CapacityOverflow, the capacity has grown, and the overflowing frame rendered again matches the staged ops.  The
    truncated launch itself must not write outside its clipped tile segments (sort.cu emit pass)."""
    from splatter_a_video_b200.gs.frame import CapacityOverflow
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_scene(20_000, 4, 256, 192, seed=33)
    ref = parse_renderer({"name": "DPTROrthoEnhancedRender"}, white_bg=False, device=cuda)
    rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    easy = {k: v.detach() for k, v in _rd(sc, cuda).items()}
    hard = dict(easy); hard["scaling"] = easy["scaling"] * 3.0          # ~9x the tile intersections
    rnd.render_batch(dict(easy), [_batch(sc, cuda)])                    # settles the capacity (waits once)
    cap0 = rnd.capacity.I_cap
    o_trunc = rnd.render_batch(dict(hard), [_batch(sc, cuda)])          # overflows; not known yet (no host wait)
    torch.cuda.synchronize()
    assert int(rnd.last_status.cpu()[1]) == 1
    assert torch.isfinite(o_trunc["rgb"]).all()
    with pytest.raises(CapacityOverflow):
        rnd.render_batch(dict(easy), [_batch(sc, cuda)])                # the earlier frame's status is inspected here
    assert rnd.capacity.I_cap > cap0 and rnd.capacity.late_overflows == 1
    for _ in range(16):                                                 # grown (x1.5 per report) until the hard frame fits
        o2 = rnd.render_batch(dict(hard), [_batch(sc, cuda)])
        if rnd.capacity.drain():                                        # False: this frame overflowed, capacity grown
            break
    o1 = ref.render_batch(dict(hard), [_batch(sc, cuda)])
    assert int(rnd.last_status.cpu()[1]) == 0
    assert float((o1["rgb"] - o2["rgb"]).abs().max()) <= 1e-6 and torch.equal(o1["gs_idx"], o2["gs_idx"])


def test_capacity_follows_the_population(cuda):
    """P changes after densification: the capacity is rescaled instead of staying at the first frame's value."""
    from splatter_a_video_b200.renderer import parse_renderer
    rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    small = synth.make_scene(5_000, 4, 128, 96, seed=9)
    big = synth.make_scene(15_000, 4, 128, 96, seed=9)
    rnd.render_batch({k: v.detach() for k, v in _rd(small, cuda).items()}, [_batch(small, cuda)])
    cap0 = rnd.capacity.I_cap
    out = rnd.render_batch({k: v.detach() for k, v in _rd(big, cuda).items()}, [_batch(big, cuda)])
    assert rnd.capacity.I_cap >= 3 * (cap0 - 4096)
    rnd.capacity.drain()
    assert int(rnd.last_status.cpu()[1]) == 0 and torch.isfinite(out["rgb"]).all()


def test_flat_adam_replays_inside_a_cuda_graph(cuda):
    """A captured optimizer step must advance its bias corrections on every replay (the host-clock entry would freeze them at
    capture time) and pick up learning rates changed between replays."""
    from splatter_a_video_b200.parallel import FlatAdam, FlatParams
    g = torch.Generator().manual_seed(1)
    tensors = {"a": torch.randn(512, 3, generator=g), "b": torch.randn(512, 4, generator=g)}
    lrs = {"a": 1e-2, "b": 3e-3}
    flat = FlatParams({k: v.to(cuda) for k, v in tensors.items()})
    opt = FlatAdam(flat, lrs, eps=1e-15, device_clock=True)
    ref = {k: v.clone().to(cuda).requires_grad_(True) for k, v in tensors.items()}
    topt = torch.optim.Adam([{"params": [ref[k]], "lr": lrs[k]} for k in ref], eps=1e-15)
    grads = [{k: torch.randn(tensors[k].shape, generator=g).to(cuda) for k in tensors} for _ in range(6)]
    static = {k: torch.zeros_like(grads[0][k]) for k in tensors}

    def step():
        for k in tensors:
            flat[k].grad.copy_(static[k])
        opt.step()

    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            step()
    torch.cuda.current_stream().wait_stream(s)
    for it in range(6):
        if it == 3:
            lrs = {"a": 5e-3, "b": 1e-3}
            opt.set_lrs(lrs)
            for grp, k in zip(topt.param_groups, ref):
                grp["lr"] = lrs[k]
        for k in tensors:
            static[k].copy_(grads[it][k]); ref[k].grad = grads[it][k].clone()
        gr.replay(); topt.step()
        torch.cuda.synchronize()
        for k in tensors:
            np.testing.assert_allclose(n(flat[k]), n(ref[k]), rtol=3e-6, atol=1e-7)


@pytest.mark.parametrize("F", [50, 7])
def test_deform_position_polyfourier_matches_reference_golden(cuda, F):
    """gs.frame.deform_position_polyfourier (spv_deform_polyfourier_*) against get_position of the ALTERNATIVE model executed on the
    CPU (golden_deform.npz, tests/golden/make_deform_golden.py): every frame forward, all three gradients at one frame, and the
    detach_pos flag."""
    import os
    from splatter_a_video_b200.gs.frame import deform_position_polyfourier, rotation_basis
    G = np.load(os.path.join(Hh.GOLDEN, "golden_deform.npz"))
    pre = f"ALT{F}_"
    leaves = [torch.from_numpy(G[pre + k]).to(cuda).requires_grad_(True) for k in ("position", "poly", "fourier")]
    for t in range(F):
        with torch.no_grad():
            got = deform_position_polyfourier(*leaves, rotation_basis(t, 0, F - 1).to(cuda))
        np.testing.assert_allclose(n(got), G[pre + "pos_t"][t], rtol=0, atol=2e-6, err_msg=f"frame {t}")
    t_g = int(G[pre + "t_grad"])
    w = torch.from_numpy(G[pre + "w"]).to(cuda)
    (deform_position_polyfourier(*leaves, rotation_basis(t_g, 0, F - 1).to(cuda)) * w).sum().backward()
    for leaf, k in zip(leaves, ("g_position", "g_poly", "g_fourier")):
        np.testing.assert_allclose(n(leaf.grad), G[pre + k], rtol=1e-6, atol=1e-7)
    for leaf in leaves:
        leaf.grad = None
    (deform_position_polyfourier(*leaves, rotation_basis(t_g, 0, F - 1).to(cuda), detach_pos=True) * w).sum().backward()
    assert leaves[0].grad is None
    np.testing.assert_allclose(n(leaves[1].grad), G[pre + "g_poly"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("defer", [False, True])
def test_interval_major_spline_storage_equals_reference_layout(cuda, defer):
    """The deformation ops on interval-major coefficients ([P,NI,4,3], gs.frame.spline_to_interval_major) give the same positions
    and -- after converting back -- the same coefficient gradients as on the reference's [P,4,NI,3] layout: single evaluation, the
    two-frame op with its incrementally cleaned sink, and the gathered (frame-parallel) backward."""
    from splatter_a_video_b200 import _lib as L
    from splatter_a_video_b200.gs.frame import (deform_position, deform_position_pair, spline_from_interval_major, spline_interval,
                                                spline_to_interval_major)
    P, T = 4001, 50
    NI = -(-T // 5)
    g = torch.Generator().manual_seed(14)
    base = torch.randn(P, 3, generator=g).to(cuda)
    node = (0.1 * torch.randn(P, 4 * NI * 3, generator=g)).to(cuda)
    node_im = spline_to_interval_major(node, NI)
    assert torch.equal(spline_from_interval_major(node_im, NI), node)
    dev_i = lambda v: torch.tensor([v], dtype=torch.int32, device=cuda)
    dev_f = lambda v: torch.tensor([v], dtype=torch.float32, device=cuda)
    # single evaluation
    for time in (0, 24, 49):
        i, d = spline_interval(time, T, NI)
        a = node.clone().requires_grad_(True); b = node_im.clone().requires_grad_(True)
        pa = deform_position(base, a, dev_i(i), dev_f(d), NI)
        pb = deform_position(base, b, dev_i(i), dev_f(d), NI, interval_major=True)
        assert torch.equal(pa, pb)
        gp = torch.randn(P, 3, generator=g).to(cuda)
        pa.backward(gp); pb.backward(gp)
        assert torch.equal(spline_from_interval_major(b.grad, NI), a.grad)
    # two frame times into sinks kept clean through the dirty lists, several steps in a row (intervals change between steps)
    sinks = [torch.zeros(P, 4 * NI * 3, device=cuda) for _ in range(2)]
    dirty = [torch.zeros(17, dtype=torch.int32, device=cuda) for _ in range(2)]
    for (t1, t2) in [(3, 4), (4, 5), (30, 31), (49, 49)]:
        (i1, d1), (i2, d2) = spline_interval(t1, T, NI), spline_interval(t2, T, NI)
        g1, g2 = torch.randn(P, 3, generator=g).to(cuda), torch.randn(P, 3, generator=g).to(cuda)
        outs = []
        for which, (coef, im) in enumerate(((node, False), (node_im, True))):
            c = coef.clone().requires_grad_(True)
            if defer:
                payload = torch.zeros(6 * P + 4, device=cuda)
                p1, p2 = deform_position_pair(base, c, dev_i(i1), dev_f(d1), dev_i(i2), dev_f(d2), NI, sinks[which], dirty[which], payload,
                                              interval_major=im)
                torch.autograd.backward([p1, p2], [g1, g2])
                rows = payload.reshape(1, -1).contiguous()
                L.call("spv_deform_spline_backward_gathered", P, NI, int(im), 1, L.ptr(rows), rows.stride(0), 1.0, L.ptr(dirty[which]),
                       L.ptr(sinks[which]), L.stream())
            else:
                p1, p2 = deform_position_pair(base, c, dev_i(i1), dev_f(d1), dev_i(i2), dev_f(d2), NI, sinks[which], dirty[which],
                                              interval_major=im)
                torch.autograd.backward([p1, p2], [g1, g2])
            outs.append((p1.detach(), p2.detach()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert torch.equal(spline_from_interval_major(sinks[1], NI), sinks[0])
        assert torch.equal(dirty[0], dirty[1])


@pytest.mark.parametrize("interval_major", [False, True])
def test_interval_lazy_adam_equals_dense_adam(cuda, interval_major):
    """FlatAdam(lazy=...): a step streams only the spline intervals that hold gradient and replays the zero-gradient updates of
    the others when they are next needed.  Against the dense device-clock optimizer (itself held to torch.optim.Adam above) on the
    same gradient schedule: the intervals a forward pass is about to read are identical after prepare(), everything is identical
    after flush() -- parameters and both moments, including a learning-rate change on the way."""
    from splatter_a_video_b200.parallel import FlatAdam, FlatParams
    P, NI = 257 * 4, 7
    g = torch.Generator().manual_seed(5)
    tensors = {"node": 0.1 * torch.randn(P, 4 * NI * 3, generator=g), "scaling": torch.randn(P, 3, generator=g), "opacity": torch.randn(P, 1, generator=g)}
    lrs = {"node": 2e-3, "scaling": 5e-3, "opacity": 5e-2}
    dense_flat = FlatParams({k: v.to(cuda) for k, v in tensors.items()})
    lazy_flat = FlatParams({k: v.to(cuda) for k, v in tensors.items()})
    dirty = torch.zeros(17, dtype=torch.int32, device=cuda)
    dense = FlatAdam(dense_flat, lrs, eps=1e-15, device_clock=True)
    lazy = FlatAdam(lazy_flat, lrs, eps=1e-15, lazy={"name": "node", "P": P, "NI": NI, "interval_major": interval_major, "dirty": dirty})

    def view(t):          # [P, 4*NI*3] -> [P, NI, 12] (interval, slot) whatever the storage order
        return t.reshape(P, NI, 4, 3).reshape(P, NI, 12) if interval_major else t.reshape(P, 4, NI, 3).permute(0, 2, 1, 3).reshape(P, NI, 12)

    schedule = [(0, 0), (0, 1), (1, 1), (5, 6), (6, 6), (2, 3), (0, 0), (3, 3), (6, 5), (1, 2), (4, 4), (0, 6)] * 3
    dev_i = lambda v: torch.tensor([v], dtype=torch.int32, device=cuda)
    for it, (i1, i2) in enumerate(schedule):
        lazy.prepare(dev_i(i1), dev_i(i2))
        for b in {i1, i2}:        # what the forward pass of this step would read
            assert torch.equal(view(lazy_flat["node"].detach())[:, b], view(dense_flat["node"].detach())[:, b]), (it, b)
        if it == 17:
            lrs2 = {"node": 1e-3, "scaling": 2e-3, "opacity": 1e-2}
            dense.set_lrs(lrs2); lazy.set_lrs(lrs2)
        # this step's gradient: dense for the small parameters, the two intervals only for the coefficients
        gn = torch.zeros(P, NI, 12)
        for b in {i1, i2}:
            gn[:, b] = torch.randn(P, 12, generator=g)
        gn = (gn.reshape(P, NI, 4, 3) if interval_major else gn.reshape(P, NI, 4, 3).permute(0, 2, 1, 3)).reshape(P, -1)
        for flat in (dense_flat, lazy_flat):
            flat["node"].grad.copy_(gn.to(cuda))
        for k in ("scaling", "opacity"):
            gk = torch.randn(tensors[k].shape, generator=g).to(cuda)
            dense_flat[k].grad.copy_(gk); lazy_flat[k].grad.copy_(gk)
        dirty.copy_(torch.tensor([2, i1, i2] + [0] * 14, dtype=torch.int32))
        dense.step(); lazy.step()
    # un-flushed, idle intervals lag behind; flushed, everything is the dense optimizer's state
    lazy.flush()
    torch.cuda.synchronize()
    assert torch.equal(lazy.last_dev.cpu(), torch.full((NI,), len(schedule), dtype=torch.int32))
    for a, b, name in ((lazy_flat.flat, dense_flat.flat, "param"), (lazy.exp_avg, dense.exp_avg, "exp_avg"), (lazy.exp_avg_sq, dense.exp_avg_sq, "exp_avg_sq")):
        np.testing.assert_allclose(n(a), n(b), rtol=2e-6, atol=1e-9, err_msg=name)


def test_sh_along_z_from_four_bases_is_bit_identical(cuda):
    """spv_compute_sh_z_forward / _backward on the [P,4,3] coefficients of the bases (0, 2, 6, 12) == compute_sh(deg 3) along the
    renderer's constant direction (0,0,1) on the full [P,16,3] tensor: colours and clamp pattern bit for bit, coefficient
    gradients bit for bit on the four bases and exactly zero on the other twelve."""
    from splatter_a_video_b200 import _lib as L
    from splatter_a_video_b200 import gs
    from splatter_a_video_b200.gs.frame import SH_Z_BASES, sh_z_merge, sh_z_split
    P = 70_001
    g = torch.Generator().manual_seed(31)
    shs = torch.randn(P, 16, 3, generator=g).to(cuda)         # about half of the colours clamp at zero
    dirs = torch.zeros(P, 3, device=cuda); dirs[:, 2] = 1
    full = shs.clone().requires_grad_(True)
    want = gs.compute_sh(full, 3, dirs)
    gcol = torch.randn(P, 3, generator=g).to(cuda)
    want.backward(gcol)
    shs_z, rest = sh_z_split(shs)
    assert torch.equal(sh_z_merge(shs_z, rest), shs)
    col = torch.empty(P, 3, device=cuda); clamped = torch.empty(P, 3, dtype=torch.uint8, device=cuda)
    L.call("spv_compute_sh_z_forward", P, L.ptr(shs_z), L.ptr(col), L.ptr(clamped), L.stream())
    assert torch.equal(col, want.detach())
    assert bool((want.detach()[clamped.bool()] == 0).all()) and int(clamped.sum()) > P // 10      # clamped => colour 0; the mask is exercised
    gz = torch.empty(P, 4, 3, device=cuda)
    L.call("spv_compute_sh_z_backward", P, L.ptr(clamped), L.ptr(gcol), L.ptr(gz), L.stream())
    act = list(SH_Z_BASES)
    assert torch.equal(gz, full.grad[:, act])
    dead = [b for b in range(16) if b not in act]
    assert float(full.grad[:, dead].abs().max()) == 0.0


def test_frame_path_with_the_four_reachable_sh_bases(cuda):
    """DPTROrthoEnhancedRenderB200 fed shs = [P,4,3] (gs.frame.sh_z_split): the same images and gradients as with [P,16,3], the SH
    gradient restricted to the four bases; the staged renderers refuse the short tensor."""
    from splatter_a_video_b200.gs.frame import SH_Z_BASES, sh_z_split
    from splatter_a_video_b200.renderer import parse_renderer
    P, W, H = 20_000, 256, 160
    sc = synth.make_scene(P, 6, W, H, seed=19)
    g = torch.Generator().manual_seed(6)
    keys = ["rgb", "depth"] + ATTRS
    chans = {"rgb": 3, "depth": 1, "track_gs": 3, "mask_attribute": 1, "pos_poly_feat": 12, "dino_attribute": 3}
    gimgs = {k: torch.randn(c, H, W, generator=g).to(cuda) for k, c in chans.items()}

    def run(short):
        rd = _rd(sc, cuda)
        if short:
            rd["shs"] = sh_z_split(rd["shs"].detach())[0].requires_grad_(True)
        rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
        out = rnd.render_batch(rd, [_batch(sc, cuda)])
        torch.autograd.backward([out[k][0] for k in keys], [gimgs[k] for k in keys])
        return out, rd

    o16, r16 = run(False)
    o4, r4 = run(True)
    for k in keys:
        assert torch.equal(o16[k], o4[k]), k
    Hh.assert_grad_close(n(r4["shs"].grad), n(r16["shs"].grad[:, list(SH_Z_BASES)]), "d/dshs", norm_tol=2e-5)
    for k in ("position", "scaling", "rotation", "opacity", "mask_attribute"):
        Hh.assert_grad_close(n(r4[k].grad), n(r16[k].grad), f"d/d{k}", norm_tol=2e-5)
    rd = _rd(sc, cuda)
    rd["shs"] = sh_z_split(rd["shs"].detach())[0]
    with pytest.raises(ValueError):
        parse_renderer({"name": "DPTROrthoEnhancedRender"}, white_bg=False, device=cuda).render_batch(rd, [_batch(sc, cuda)])
