"""In-situ kernel durations of one training step (frame mode) with torch.profiler (CUPTI activity records): unlike the ncu
launch list (cold L2, serialised replays) these are the durations the kernels have INSIDE the step, with the previous
kernel's outputs still in L2.  Writes gpurun_out/<tag>_kernels_{graph,eager}.json: per kernel name count / total / mean us.

    python scripts/profile_step.py <tag> [steps]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def profile(wl, steps, flush):
    from torch.profiler import ProfilerActivity, profile as tprofile
    for i in range(3):
        wl.step_resident(i)
    torch.cuda.synchronize()
    with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(steps):
            flush.zero_()
            wl.step_resident(3 + i)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type is None or "cuda" not in str(ev.device_type).lower():
            continue
        r = rows.setdefault(ev.name, [0, 0.0])
        r[0] += 1; r[1] += float(ev.device_time if hasattr(ev, "device_time") else ev.cuda_time)
    out = [{"kernel": k, "launches_per_step": n / steps, "us_per_step": t / steps, "mean_us": t / n} for k, (n, t) in rows.items()]
    out.sort(key=lambda r: -r["us_per_step"])
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "prof"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for graph in (True, False):
        wl = bench.Workload("cfg2_davis480p", dev, mode="frame", graph=graph)
        rows = profile(wl, steps, flush)
        total = sum(r["us_per_step"] for r in rows if "Memset" not in r["kernel"] or True)
        name = os.path.join(ROOT, "gpurun_out", f"{tag}_kernels_{'graph' if graph else 'eager'}.json")
        with open(name, "w") as f:
            json.dump({"steps": steps, "sum_us_per_step": total, "note": "the 512 MiB L2 flush between steps is listed too (vectorized fill kernel)",
                       "kernels": rows}, f, indent=1)
        print(("graph" if graph else "eager"), f"sum of device activity {total:.1f} us/step")
        for r in rows[:40]:
            print(f"  {r['us_per_step']:9.1f} us  x{r['launches_per_step']:5.1f}  {r['kernel'][:110]}")
        del wl
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
