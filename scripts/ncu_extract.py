#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full) into profiles/: a trimmed CSV of the metrics DESIGN.md / bench.py quote and
the matching entry of profiles/ncu_traffic.json (per-launch DRAM bytes -> roofline.traffic, warp instructions -> roofline_issue).

    python scripts/ncu_extract.py gpurun_out/prof_bwd_r02.ncu-rep blend_records_backward profiles/r02_ncu_blend_rec_bwd.csv
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]


def main():
    rep, key, out_csv = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    row = data[-1]                                      # the last captured launch (warm)
    col = {h: i for i, h in enumerate(hdr)}
    get = lambda name: row[col[name]] if name in col else ""
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["Kernel Name", "", get("Kernel Name")])
        for k in KEEP:
            if k in col:
                w.writerow([k, units[col[k]], row[col[k]]])
    num = lambda name: float(get(name).replace(",", "")) if get(name) else 0.0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = num("dram__bytes_read.sum") * scale.get(units[col["dram__bytes_read.sum"]], 1.0)
    wr = num("dram__bytes_write.sum") * scale.get(units[col["dram__bytes_write.sum"]], 1.0)
    dur_unit = units[col["gpu__time_duration.sum"]]
    dur_us = num("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(dur_unit, 1.0)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path))
    table[key] = {"bytes": int(rd + wr), "read": int(rd), "write": int(wr), "inst": int(num("smsp__inst_executed.sum")),
                  "capture": os.path.relpath(out_csv, ROOT), "kernel": get("Kernel Name")[:120], "duration_us": round(dur_us, 2)}
    json.dump(table, open(path, "w"), indent=1)
    print(key, json.dumps(table[key]))


if __name__ == "__main__":
    main()
