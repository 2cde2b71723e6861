#!/usr/bin/env python
"""Small-n latency sweep (BASELINE config 5): us/call for f32 n = 2^10 .. 2^20 (and beyond)
through the synchronous C-ABI call (host wall clock, median) and kernel-only (CUDA events over
back-to-back launches).  --tune also sweeps load shape / CTA size per n."""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def measure(ds, calls):
    for _ in range(50):
        ds.argminmax()
    lat = []
    for _ in range(calls):
        t0 = time.perf_counter()
        ds.argminmax()
        lat.append((time.perf_counter() - t0) * 1e6)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        ds.launch(0)
    e1.record()
    e1.synchronize()
    ds.wait(0)
    return statistics.median(lat), min(lat), e0.elapsed_time(e1) / 200 * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--calls", type=int, default=1000)
    ap.add_argument("--tune", action="store_true")
    ap.add_argument("--max-log2", type=int, default=24)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    lib = amm.lib()
    es = amm.DTYPE_SIZE[a.dtype]
    nmax = 1 << a.max_log2
    t = torch.randint(-2**31, 2**31 - 1, (nmax * es // 4 + 4,), dtype=torch.int32, device="cuda:0")
    if a.dtype == "f32":
        f = t.view(torch.float32)
        f[torch.isnan(f)] = 0
    raw = t.view(torch.uint8)
    ws = amm.Workspace(0)
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    for lg in range(10, a.max_log2 + 1):
        n = 1 << lg
        ds = amm.DeviceSlice.from_torch(raw[: n * es], dtype=a.dtype, workspace=ws, stream=stream)
        if a.tune:
            best = None
            for v in range(lib.amm_variant_count()):
                for th in (128, 256, 512):
                    ws.set_tuning(th, 8, v)
                    med, mn, dev = measure(ds, 200)
                    row = {"n": n, "variant": lib.amm_variant_name(v).decode(), "threads": th,
                           "us_call_median": round(med, 2), "us_call_min": round(mn, 2),
                           "us_kernel": round(dev, 2), "grid": ws.last_launch()["grid"]}
                    rows.append(row)
                    if best is None or row["us_kernel"] < best["us_kernel"]:
                        best = row
            print("BEST", json.dumps(best), flush=True)
        else:
            ws.set_tuning(0, 0, -1)
            med, mn, dev = measure(ds, a.calls)
            row = {"n": n, "us_call_median": round(med, 2), "us_call_min": round(mn, 2),
                   "us_kernel": round(dev, 2), "GBps_kernel": round(n * es / dev / 1e3, 1),
                   "grid": ws.last_launch()["grid"], "threads": ws.last_launch()["threads"]}
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
