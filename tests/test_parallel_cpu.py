"""CPU-only, gloo, world_size 2: the frame-parallel host logic (frame sharding, flat gradient buffer, ONE all-reduce per
step, densification-statistics reduce).  The renderer needs CUDA, so the differentiable CPU restatement
(oracle/torch_ref.py -- test infrastructure) stands in for it; the property checked is the one DESIGN.md section 6 states:
N ranks rendering N different frames + one all-reduce == one rank accumulating the same frames."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as Hh  # noqa: F401  (sys.path)
from splatter_a_video_b200 import synth
from splatter_a_video_b200.parallel import FlatParams, GradExchange, frame_for_step, reduce_densify_stats, shard_frames


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _scene():
    return synth.make_scene(120, 4, 32, 32, seed=5)


def _loss_for_frame(sc, flat, frame):
    from oracle import torch_ref as TR
    pos = flat["position"] + synth.eval_spline(sc.nodes, frame / 5.0)
    out = TR.render_ortho_frame(pos, flat["scaling"], flat["rotation"], flat["opacity"], flat["shs"], None, sc.extr, sc.W, sc.H, K=4)
    w = torch.linspace(0.5, 1.5, 3 * sc.H * sc.W).reshape(3, sc.H, sc.W)
    return (out["rgb"] * w).sum() + out["depth"].sum(), out


def _params(sc):
    return FlatParams({"position": sc.position.clone(), "scaling": sc.scaling.clone(), "rotation": sc.rotation.clone(),
                       "opacity": sc.opacity.clone(), "shs": sc.shs.clone()})


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sc = _scene()
    flat = _params(sc)
    frame = frame_for_step(0, rank, world, sc.frames)
    loss, out = _loss_for_frame(sc, flat, frame)
    flat.zero_grad()
    loss.backward()
    flat.allreduce_grads(average=False)
    radius = out["radius"].float()
    gsum, vis, rmax = reduce_densify_stats(torch.full((sc.P,), float(rank + 1)), (out["radius"] > 0).float(), radius)
    if rank == 0:
        q.put((flat.flat_grad.clone().numpy(), gsum.numpy(), vis.numpy(), rmax.numpy(), frame))
    dist.barrier()
    dist.destroy_process_group()


def _sparse_case(rank, case):
    """Interval pairs (ids1, ids2) per rank: distinct, equal on one rank, and overlapping across ranks."""
    return {0: [(2, 3), (5, 5)], 1: [(1, 2), (2, 4)], 2: [(0, 0), (0, 0)]}[case][rank]


def _sparse_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, NI = 50, 6
    res = []
    for case in range(3):
        g = torch.Generator().manual_seed(100 + rank + 10 * case)
        flat = FlatParams({"node": torch.zeros(P, 4 * NI * 3), "other": torch.zeros(P, 7), "shs": torch.zeros(P, 16, 3),
                           "tail": torch.zeros(P, 2)})
        i1, i2 = _sparse_case(rank, case)
        idx = [torch.tensor([i1], dtype=torch.int32), torch.tensor([i2], dtype=torch.int32)]   # the rank's two frame times
        gn = torch.zeros(P, 4, NI, 3)
        for b in {i1, i2}:
            gn[:, :, b] = torch.randn(P, 4, 3, generator=g)
        go = torch.randn(P, 7, generator=g)
        sh_idx = [0, 2, 6, 12]
        gsh = torch.zeros(P, 16, 3); gsh[:, sh_idx] = torch.randn(P, 4, 3, generator=g)
        flat["node"].grad.copy_(gn.reshape(P, -1)); flat["other"].grad.copy_(go); flat["shs"].grad.copy_(gsh)
        flat["tail"].grad.copy_(torch.randn(P, 2, generator=g))
        dense = flat.flat_grad.clone()
        dist.all_reduce(dense)
        ex = GradExchange(flat, P, subset={"shs": ((P, 16, 3), 1, sh_idx)}, sparse={"node": ((P, 4, NI, 3), 2, idx)})
        ex.run(average=False)
        res.append((flat.flat_grad.numpy().copy(), dense.numpy().copy()))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sparse_interval_exchange_equals_dense_allreduce():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=200)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for got, want in res:
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


def test_shard_frames_partition():
    for n, w in ((50, 4), (80, 8), (7, 2), (3, 8)):
        seen = []
        for r in range(w):
            seen += list(shard_frames(n, r, w))
        assert seen == list(range(n))
    assert frame_for_step(0, 1, 2, 50, policy="blocks") == 25 and frame_for_step(26, 1, 2, 50, policy="blocks") == 26
    # default = DistributedSampler order: the ranks of one step render consecutive frames, every frame once per epoch
    assert [frame_for_step(3, r, 4, 50) for r in range(4)] == [12, 13, 14, 15]
    assert sorted(frame_for_step(s_, r, 2, 50) for s_ in range(25) for r in range(2)) == list(range(50))


def test_flat_params_views_and_grads():
    sc = _scene()
    flat = _params(sc)
    assert flat.flat.numel() == sc.P * (3 + 3 + 4 + 1 + 48)
    (flat["scaling"].sum() * 2 + flat["shs"].sum()).backward()
    off = sc.P * 3
    assert torch.all(flat.flat_grad[off:off + sc.P * 3] == 2) and torch.all(flat.flat_grad[:off] == 0)
    assert flat["scaling"].grad.data_ptr() == flat.flat_grad[off:].data_ptr()   # in-place accumulation into the flat buffer


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank_accumulating_both_frames():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, gsum, vis, rmax, frame0 = q.get(timeout=500)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process reference: accumulate the two frames the ranks rendered
    sc = _scene()
    flat = _params(sc)
    flat.zero_grad()
    vis_ref = torch.zeros(sc.P); rmax_ref = torch.zeros(sc.P)
    for r in range(world):
        loss, out = _loss_for_frame(sc, flat, frame_for_step(0, r, world, sc.frames))
        loss.backward()
        vis_ref += (out["radius"] > 0).float(); rmax_ref = torch.maximum(rmax_ref, out["radius"].float())
    np.testing.assert_allclose(got, flat.flat_grad.numpy(), rtol=1e-5, atol=1e-6 * np.abs(got).max())
    assert np.array_equal(gsum, np.full(sc.P, 3.0)) and np.array_equal(vis, vis_ref.numpy()) and np.array_equal(rmax, rmax_ref.numpy())
    assert frame0 == 0
