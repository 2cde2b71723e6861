#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv, *.ncu-rep} into the small, committed summaries under profiles/.

  python tools/ncu_summary.py --tag r01_f32_1g --rep gpurun_out/prof_f32_1g.ncu-rep \
         --launches gpurun_out/launches.csv [--traffic-key f32_1g_dram_bytes_per_launch]
"""
import argparse
import collections
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launches_summary(path, tag):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        us = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        name = r[ki]
        short = name if len(name) < 120 else name[:117] + "..."
        key = (short, r[gi], r[bi])
        a = agg.setdefault(key, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = min(a[2], us)
        a[3] = max(a[3], us)
    total = sum(a[1] for a in agg.values())
    out = os.path.join(PROF, f"{tag}_launches.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "total_us", "min_us", "max_us", "share_pct"])
        for (k, g, b), a in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, g, b, a[0], round(a[1], 1), round(a[2], 2), round(a[3], 2),
                        round(100 * a[1] / total, 2)])
    return out


def full_summary(rep, tag, traffic_key):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = os.path.join(PROF, f"{tag}_ncu_full_metrics.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])
    if traffic_key:
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

        def to_bytes(val, unit):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
            return float(val.replace(",", "")) * scale
        per = [to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for r in data]
        tj = os.path.join(PROF, "traffic.json")
        cur = {}
        if os.path.exists(tj):
            cur = json.load(open(tj))
        cur[traffic_key] = round(sum(per) / len(per))
        cur[traffic_key + "_source"] = f"profiles/{tag}_ncu_full_metrics.csv (dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(per)} launches)"
        json.dump(cur, open(tj, "w"), indent=1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--traffic-key", default="")
    a = ap.parse_args()
    os.makedirs(PROF, exist_ok=True)
    if a.launches:
        print(launches_summary(a.launches, a.tag))
    if a.rep:
        print(full_summary(a.rep, a.tag, a.traffic_key))


if __name__ == "__main__":
    main()
