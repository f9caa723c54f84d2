"""Micro-benchmark of the gradient-exchange pack / unpack kernels on one GPU (W simulated ranks): CUDA-event times at the
bench's shape (P = 200k, 50 frames).  python scripts/exchange_micro.py [W]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter_a_video_b200.parallel import FlatParams, GradExchange  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
P, NI = 200_000, 10
flat = FlatParams({"pos_cubic_node": torch.zeros(P, 4 * NI * 3, device=dev), "scaling": torch.zeros(P, 3, device=dev),
                   "rotation": torch.zeros(P, 4, device=dev), "opacity": torch.zeros(P, 1, device=dev),
                   "shs": torch.zeros(P, 16, 3, device=dev), "mask_attribute": torch.zeros(P, 1, device=dev),
                   "dino_attribute": torch.zeros(P, 3, device=dev)})
flat.flat_grad.normal_()
idx = [torch.tensor([3], dtype=torch.int32, device=dev), torch.tensor([4], dtype=torch.int32, device=dev)]
dirty = torch.zeros(17, dtype=torch.int32, device=dev)
ex = GradExchange(flat, P, subset={"shs": ((P, 16, 3), 1, [0, 2, 6, 12])}, sparse={"pos_cubic_node": ((P, 4, NI, 3), 2, idx)}, dirty=dirty)
ar, ag = ex.pack(1.0 / W)
gathered = ag[None].repeat(W, 1).contiguous()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("pack", lambda: ex.pack(1.0 / W)), ("unpack", lambda: ex.unpack(ar, gathered))):
    for cold in (False, True):
        ts = []
        for _ in range(7):
            if cold:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{name} ({'cold' if cold else 'warm'} L2): {sorted(ts)[len(ts) // 2] * 1e3:.1f} us")
