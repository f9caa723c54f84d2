"""-m gpu: the gradient-exchange pack / unpack kernels (`spv_exchange_*`) and the two-frame deformation op.  The collectives
themselves are NCCL's; here W simulated ranks on one GPU are packed, reduced / gathered with torch, and unpacked -- the result
must equal the dense sum of the ranks' flat gradient buffers (the property tests/test_parallel_cpu.py checks over gloo)."""
import pytest
import torch

import helpers as Hh  # noqa: F401
from splatter_a_video_b200.parallel import FlatParams, GradExchange

pytestmark = pytest.mark.gpu


def _rank_state(cuda, P, NI, rank, pair, seed):
    g = torch.Generator().manual_seed(seed + rank)
    flat = FlatParams({"node": torch.zeros(P, 4 * NI * 3, device=cuda), "scaling": torch.zeros(P, 3, device=cuda),
                       "rotation": torch.zeros(P, 4, device=cuda), "shs": torch.zeros(P, 16, 3, device=cuda),
                       "mask": torch.zeros(P, 1, device=cuda)})
    i1, i2 = pair
    gn = torch.zeros(P, 4, NI, 3)
    for b in {i1, i2}:
        gn[:, :, b] = torch.randn(P, 4, 3, generator=g)
    gsh = torch.zeros(P, 16, 3); gsh[:, [0, 2, 6, 12]] = torch.randn(P, 4, 3, generator=g)
    flat["node"].grad.copy_(gn.reshape(P, -1).to(cuda)); flat["shs"].grad.copy_(gsh.to(cuda))
    for k, w in (("scaling", 3), ("rotation", 4), ("mask", 1)):
        flat[k].grad.copy_(torch.randn(P, w, generator=g).to(cuda))
    idx = [torch.tensor([i1], dtype=torch.int32, device=cuda), torch.tensor([i2], dtype=torch.int32, device=cuda)]
    dirty = torch.zeros(17, dtype=torch.int32, device=cuda)
    ex = GradExchange(flat, P, subset={"shs": ((P, 16, 3), 1, [0, 2, 6, 12])}, sparse={"node": ((P, 4, NI, 3), 2, idx)}, dirty=dirty)
    return flat, ex, dirty


@pytest.mark.parametrize("pairs", [[(2, 3), (5, 5)], [(1, 2), (2, 4), (0, 0), (9, 8)], [(0, 0)] * 8])
def test_exchange_pack_unpack_equals_dense_sum(cuda, pairs):
    P, NI, W = 3001, 10, len(pairs)
    ranks = [_rank_state(cuda, P, NI, r, pairs[r], seed=7) for r in range(W)]
    dense = sum(f.flat_grad.clone() for f, _, _ in ranks) / W
    packed = [ex.pack(1.0 / W) for _, ex, _ in ranks]
    reduced = sum(ar.clone() for ar, _ in packed)                       # what the all-reduce delivers
    gathered = torch.stack([ag.clone() for _, ag in packed]).contiguous()   # what the all-gather delivers
    assert reduced.numel() == P * (3 + 4 + 12 + 1) and gathered.shape[1] == P * 24 + 16
    for r, (flat, ex, dirty) in enumerate(ranks):
        ex.unpack(reduced, gathered)
        torch.cuda.synchronize()
        assert float((flat.flat_grad - dense).abs().max()) <= 1e-6 * float(dense.abs().max())
        d = dirty.cpu().tolist()
        assert d[0] == 2 * W and d[1:1 + 2 * W] == [b for p in pairs for b in p]
    # bit-identical on every rank (same summation order)
    for flat, _, _ in ranks[1:]:
        assert torch.equal(flat.flat_grad, ranks[0][0].flat_grad)


def test_deform_pair_matches_two_single_evaluations_and_keeps_the_sink_clean(cuda):
    from splatter_a_video_b200.gs.frame import deform_position, deform_position_pair
    P, NI = 5000, 10
    g = torch.Generator().manual_seed(3)
    base = torch.randn(P, 3, generator=g).to(cuda)
    node = (0.1 * torch.randn(P, 4 * NI * 3, generator=g)).to(cuda)
    sink = torch.zeros(P, 4 * NI * 3, device=cuda)
    dirty = torch.zeros(17, dtype=torch.int32, device=cuda)
    for (i1, d1, i2, d2) in [(2, 0.03, 2, 0.05), (2, 0.07, 3, 0.0), (7, 0.01, 1, 0.02), (4, 0.02, 4, 0.02)]:
        t = lambda v, dt: torch.tensor([v], dtype=dt, device=cuda)
        a1, b1, a2, b2 = t(i1, torch.int32), t(d1, torch.float32), t(i2, torch.int32), t(d2, torch.float32)
        n_ref = node.clone().requires_grad_(True)
        r1 = deform_position(base, n_ref, a1, b1, NI)
        r2 = deform_position(base, n_ref, a2, b2, NI)
        n_new = node.clone().requires_grad_(True)
        p1, p2 = deform_position_pair(base, n_new, a1, b1, a2, b2, NI, sink, dirty)
        assert torch.equal(p1, r1) and torch.equal(p2, r2)
        g1, g2 = torch.randn(P, 3, generator=g).to(cuda), torch.randn(P, 3, generator=g).to(cuda)
        torch.autograd.backward([r1, r2], [g1, g2])
        torch.autograd.backward([p1, p2], [g1, g2])
        assert n_new.grad is None                                   # written into the sink instead
        # every interval touched by earlier iterations was cleared: the sink equals this step's gradient exactly
        assert float((sink - n_ref.grad).abs().max()) <= 1e-6 * float(n_ref.grad.abs().max())
        assert dirty.cpu().tolist()[:3] == ([1, i1, i1] if i1 == i2 else [2, i1, i2])
    # without a sink: a fresh gradient tensor through autograd; ids2 used without gradient
    n3 = node.clone().requires_grad_(True)
    p1, p2 = deform_position_pair(base, n3, a1, b1, a2, b2, NI)
    p1.sum().backward()
    n4 = node.clone().requires_grad_(True)
    deform_position(base, n4, a1, b1, NI).sum().backward()
    assert torch.equal(n3.grad, n4.grad)


def test_deferred_tails_equal_direct_backward_single_rank(cuda):
    """GradExchange(deferred=...) at world size 1: the frame backward leaves dL_drgb / dL_dpos, run() finishes the SH and
    spline backward -- every gradient must equal the direct (non-deferred) backward."""
    from splatter_a_video_b200 import synth
    from splatter_a_video_b200.gs.frame import deform_position_pair, render_ortho_frame
    sc = synth.make_scene(20_000, 50, 200, 150, seed=5)
    P, NI, W, H = sc.P, 10, sc.W, sc.H
    g = torch.Generator().manual_seed(2)
    node0 = (0.02 * torch.randn(P, 4 * NI * 3, generator=g)).to(cuda)
    gimg = [torch.randn(c, H, W, generator=g).to(cuda) for c in (3, 1, 3, 1)]
    t = lambda v, dt: torch.tensor([v], dtype=dt, device=cuda)
    i1, d1, i2, d2 = t(3, torch.int32), t(0.04, torch.float32), t(4, torch.int32), t(0.0, torch.float32)

    def run(deferred):
        flat = FlatParams({"pos_cubic_node": node0.clone(), "scaling": sc.scaling.to(cuda), "rotation": sc.rotation.to(cuda),
                           "opacity": sc.opacity.to(cuda), "shs": sc.shs.to(cuda), "mask_attribute": sc.attrs["mask_attribute"].to(cuda)})
        dirty = torch.zeros(17, dtype=torch.int32, device=cuda)
        sinks = dict(flat.grad_sinks(["scaling", "rotation", "opacity"]))
        ex = None
        if deferred:
            ex = GradExchange(flat, P, dirty=dirty, deferred={"shs": "shs", "node": "pos_cubic_node", "NI": NI})
            sinks["shs_deferred"] = ex.sh_sink()
        else:
            sinks["shs"] = flat["shs"].grad
        pos, track = deform_position_pair(sc.position.to(cuda), flat["pos_cubic_node"], i1, d1, i2, d2, NI, flat["pos_cubic_node"].grad,
                                          dirty, ex.node_defer() if deferred else None)
        imgs, _, _, status = render_ortho_frame(pos, flat["scaling"], flat["rotation"], flat["opacity"], flat["shs"],
                                                [track, flat["mask_attribute"]], sc.extr.to(cuda), W, H, 20, 0.0, 8 * P, grad_sinks=sinks)
        assert int(status.cpu()[1]) == 0
        torch.autograd.backward(imgs, gimg)
        if deferred:
            ex.run()
        torch.cuda.synchronize()
        return flat

    a, b = run(False), run(True)
    for k in a.names:
        ga, gb = a[k].grad, b[k].grad
        assert float(ga.abs().max()) > 0, k
        assert float((ga - gb).abs().max()) <= 2e-5 * float(ga.abs().max()) + 1e-7, k


def test_spline_backward_gathered_sums_ranks_in_order(cuda):
    from splatter_a_video_b200 import _lib as L
    P, NI, W = 4001, 10, 3
    g = torch.Generator().manual_seed(9)
    frames = [(2, 0.01, 3, 0.0), (3, 0.02, 3, 0.04), (9, 0.0, 2, 0.05)]
    gathered = torch.zeros(W, 6 * P + 4, device=cuda)
    want = torch.zeros(P, 4, NI, 3, dtype=torch.float64)
    for r, (i1, d1, i2, d2) in enumerate(frames):
        g1, g2 = torch.randn(P, 3, generator=g), torch.randn(P, 3, generator=g)
        gathered[r, :3 * P] = g1.reshape(-1).to(cuda); gathered[r, 3 * P:6 * P] = g2.reshape(-1).to(cuda)
        tail = torch.tensor([i1, i2], dtype=torch.int32).view(torch.float32)
        gathered[r, 6 * P:] = torch.tensor([float(tail[0]), d1, float(tail[1]), d2]).to(cuda)
        for gg, ii, dd in ((g1, i1, d1), (g2, i2, d2)):
            for a_, pw in enumerate((3, 2, 1, 0)):
                want[:, a_, ii] += 0.5 * gg.double() * dd ** pw
    sink = torch.full((P, 4 * NI * 3), 7.0, device=cuda)            # stale content in the listed dirty intervals only
    sink.view(P, 4, NI, 3)[:, :, [0, 1, 4, 6, 7, 8]] = 0            # 5 is dirty but not re-written: must be cleared
    dirty = torch.tensor([4, 2, 3, 9, 5] + [0] * 12, dtype=torch.int32, device=cuda)
    L.call("spv_deform_spline_backward_gathered", P, NI, 0, W, L.ptr(gathered), gathered.stride(0), 0.5, L.ptr(dirty), L.ptr(sink), L.stream())
    torch.cuda.synchronize()
    got = sink.view(P, 4, NI, 3).cpu().double()
    assert float((got - want).abs().max()) <= 1e-5
    assert dirty.cpu().tolist()[:4] == [3, 2, 3, 9]
