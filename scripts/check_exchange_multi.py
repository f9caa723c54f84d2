"""torchrun check (N GPUs): frame-parallel gradient exchange == dense all-reduce of the ranks' direct gradients.

Every rank renders its own frame pair of a small scene twice: (a) direct backward + dense all-reduce of the whole flat gradient
buffer (the definition), (b) deferred linear tails + GradExchange (p2p or nccl path).  Prints max relative error.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_exchange_multi.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter_a_video_b200 import synth  # noqa: E402
from splatter_a_video_b200.gs.frame import deform_position_pair, render_ortho_frame, spline_interval  # noqa: E402
from splatter_a_video_b200.parallel import FlatParams, GradExchange, frame_for_step  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sc = synth.make_scene(30_000, 50, 320, 240, seed=5)
    P, NI, W, H = sc.P, 10, sc.W, sc.H
    g = torch.Generator().manual_seed(2)
    node0 = (0.02 * torch.randn(P, 4 * NI * 3, generator=g)).to(dev)
    worst = 0.0
    for step in range(3):
        frame = frame_for_step(step * 7, rank, world, sc.frames)
        gg = torch.Generator().manual_seed(100 + rank + 10 * step)
        gimg = [torch.randn(c, H, W, generator=gg).to(dev) for c in (3, 1, 3, 1)]
        (a1, b1), (a2, b2) = spline_interval(frame, sc.frames, NI), spline_interval(min(frame + 1, sc.frames - 1), sc.frames, NI)
        t = lambda v, dt: torch.tensor([v], dtype=dt, device=dev)
        i1, d1, i2, d2 = t(a1, torch.int32), t(b1, torch.float32), t(a2, torch.int32), t(b2, torch.float32)
        flats = []
        for deferred in (False, True):
            flat = FlatParams({"pos_cubic_node": node0.clone(), "scaling": sc.scaling.to(dev), "rotation": sc.rotation.to(dev),
                               "opacity": sc.opacity.to(dev), "shs": sc.shs.to(dev), "mask_attribute": sc.attrs["mask_attribute"].to(dev)})
            dirty = torch.zeros(17, dtype=torch.int32, device=dev)
            sinks = dict(flat.grad_sinks(["scaling", "rotation", "opacity"]))
            ex = None
            if deferred:
                ex = GradExchange(flat, P, dirty=dirty, deferred={"shs": "shs", "node": "pos_cubic_node", "NI": NI})
                sinks["shs_deferred"] = ex.sh_sink()
            else:
                sinks["shs"] = flat["shs"].grad
            pos, track = deform_position_pair(sc.position.to(dev), flat["pos_cubic_node"], i1, d1, i2, d2, NI, flat["pos_cubic_node"].grad,
                                              dirty, ex.node_defer() if deferred else None)
            imgs, _, _, _ = render_ortho_frame(pos, flat["scaling"], flat["rotation"], flat["opacity"], flat["shs"],
                                               [track, flat["mask_attribute"]], sc.extr.to(dev), W, H, 20, 0.0, 8 * P, grad_sinks=sinks)
            torch.autograd.backward(imgs, gimg)
            if deferred:
                ex.run(average=True)
                ex.run(average=True) if False else None
                path = ex.exchange_path
            else:
                flat.allreduce_grads(average=True)
            torch.cuda.synchronize()
            flats.append(flat)
        ref, got = flats
        for k in ref.names:
            e = float((ref[k].grad - got[k].grad).abs().max()) / (float(ref[k].grad.abs().max()) + 1e-30)
            worst = max(worst, e)
        # every rank must hold bit-identical gradients
        mine = got.flat_grad.clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        same = bool(torch.equal(mine, other))
        if rank == 0:
            print(f"step {step}: path={path} max rel err vs dense all-reduce = {worst:.2e}; identical across ranks: ", end="")
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(bool(flag.item()))
        assert worst <= 5e-5 and bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("exchange check ok")


if __name__ == "__main__":
    main()
