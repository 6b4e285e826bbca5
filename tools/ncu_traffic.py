#!/usr/bin/env python
"""Turn an `ncu --set full` capture into the figures bench.py reports as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/<tag>.ncu-rep --config cfg4 --frames 15 [--csv profiles/<tag>_ncu_full.csv]

Reads the report with `ncu -i ... --page raw --csv`, keeps one row per kernel (the last launch of each name),
converts dram__bytes_read.sum + dram__bytes_write.sum to bytes per FRAME (the capture ran --batch <frames>
--lanes 1, so one launch = <frames> frames) and merges them into profiles/traffic_per_frame.json under the
configuration's name.  With --csv it also writes the selected ncu columns of those launches (the evidence the
JSON is derived from).  Nothing here runs on the data path.
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__cluster_max_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def short_name(full):
    m = re.search(r"(k_[a-z0-9_]+)", full)
    return m.group(1) if m else full


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--frames", type=int, default=15, help="frames per launch of the captured run")
    ap.add_argument("--csv", default=None, help="also write the selected ncu columns here")
    ap.add_argument("--out", default=str(ROOT / "profiles" / "traffic_per_frame.json"))
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    kernels, times, picked = {}, {}, {}
    for r in data:
        name = short_name(r[col["Kernel Name"]])
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
        ms = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * TIME[units[col["gpu__time_duration.sum"]]]
        kernels[name] = (rd + wr) / a.frames
        times[name] = ms
        picked[name] = r
    out = Path(a.out)
    doc = json.loads(out.read_text()) if out.exists() else {}
    doc[a.config] = {"source": f"{a.csv or a.report} (ncu --set full --clock-control none, {a.frames} frames per launch, "
                               "dram__bytes_read.sum + dram__bytes_write.sum per launch / frames)",
                     "frames_per_launch": a.frames, "kernels": kernels, "ncu_ms_per_launch": times}
    out.write_text(json.dumps(doc, indent=1, sort_keys=True) + "\n")
    if a.csv:
        keep = [k for k in KEEP if k in col]
        with open(a.csv, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(keep)
            w.writerow([units[col[k]] for k in keep])
            for r in picked.values():
                w.writerow([r[col[k]] for k in keep])
    tot = sum(kernels.values())
    for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]):
        print(f"{k:28s} {v / 1e6:9.1f} MB/frame  {times[k]:7.3f} ms/launch")
    print(f"{'total':28s} {tot / 1e6:9.1f} MB/frame")


if __name__ == "__main__":
    sys.exit(main())
