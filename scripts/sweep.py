#!/usr/bin/env python
"""BASELINE.json configs[4]: render sweep 100k-5M Gaussians x {480p, 1080p}, forward + backward, 1 GPU, HBM-roofline report.

One frame per point through the fused per-frame path (gs.frame.render_ortho_frame: RGB K=20 + depth + 19 attribute
channels).  Prints one JSON line per (P, resolution): CUDA-event times of forward and backward (median of --reps after
warm-up, 512 MiB L2 flush before each), the intersection count, and achieved algorithmic GB/s of the whole frame against
MEASURED_PEAKS.json.  Usage: python scripts/sweep.py [--quick] > profiles/rNN_sweep.jsonl"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter_a_video_b200 import synth  # noqa: E402
from splatter_a_video_b200.gs import frame as F  # noqa: E402


def frame_bytes(P, I, W, H, C=23, K=20):
    """Algorithmic bytes of one frame: preprocess fwd 280 P + bwd 504 P, emit 20 P + 12 I, sort 2*12 I (lower
    bound), blend fwd I(28+4C) + HW(4C+8+4K), blend bwd I(28+4C) + HW(4C+8) + 4P(8+C)."""
    fwd = 280 * P + 20 * P + 12 * I + 24 * I + I * (28 + 4 * C) + H * W * (4 * C + 8 + 4 * K)
    bwd = 504 * P + I * (28 + 4 * C) + H * W * (4 * C + 8) + 4 * P * (8 + C)
    return fwd, bwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    Ps = [100_000, 1_000_000] if a.quick else [100_000, 200_000, 500_000, 1_000_000, 2_000_000, 5_000_000]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for (W, H) in ((854, 480), (1920, 1080)):
        for P in Ps:
            sc = synth.make_scene(P, 2, W, H, seed=1234).to(dev)
            leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("position", "scaling", "rotation", "opacity", "shs")}
            attrs = sc.attr_features(1).clone().requires_grad_(True)
            g = torch.randn(23, H, W, device=dev)
            cap = 8 * P
            tf, tb, I = [], [], 0
            for rep in range(a.reps + 2):
                flush.zero_()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                img, gs_idx, radii, status = F.render_ortho_frame(leaves["position"], leaves["scaling"], leaves["rotation"],
                                                                  leaves["opacity"], leaves["shs"], attrs, sc.extr, W, H, 20, 0.0, cap)
                e[1].record()
                img.backward(g)
                e[2].record()
                torch.cuda.synchronize()
                if rep == 0:
                    st = status.cpu()
                    I = int(st[0]); assert int(st[1]) == 0
                    cap = int(1.1 * I) + 4096
                for p in [*leaves.values(), attrs]:
                    p.grad = None
                if rep >= 2:
                    tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
            fb, bb = frame_bytes(P, I, W, H)
            mf, mb = float(np.median(tf)), float(np.median(tb))
            print(json.dumps({"P": P, "W": W, "H": H, "I_culled": I, "fwd_ms": mf, "bwd_ms": mb, "fps_fwd": 1000.0 / mf,
                              "it_per_s_fwd_bwd": 1000.0 / (mf + mb), "algorithmic_GB": (fb + bb) / 1e9,
                              "achieved_GBps": (fb + bb) / 1e6 / (mf + mb), "frac_of_hbm_peak": (fb + bb) / 1e6 / (mf + mb) / peak,
                              "peak_GBps": peak}), flush=True)
            del sc, leaves, attrs, g
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
