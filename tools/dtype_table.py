#!/usr/bin/env python
"""GB/s of the scan for every dtype x NaN mode at a fixed byte size (default 4 GiB), with the
library's automatic geometry, plus optional explicit variants.  Kernel-only, CUDA events."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1 << 32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--variants", default="", help="extra explicit variants, e.g. 0,1,5")
    ap.add_argument("--out", default="")
    ap.add_argument("--dtypes", default="", help="comma list, default all")
    a = ap.parse_args()
    lib = amm.lib()
    g = torch.Generator(device="cuda:0").manual_seed(3)
    words = a.bytes // 4
    t = torch.empty(words, dtype=torch.int32, device="cuda:0")
    for i in range(0, words, 1 << 28):
        t[i:i + (1 << 28)].random_(-2**31, 2**31 - 1, generator=g)
    raw = t.view(torch.uint8)
    ws = amm.Workspace(0)
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    configs = [(-1, 0, 0)] + [(int(v), 256, 8) for v in a.variants.split(",") if v]
    for dt in ([d for d in a.dtypes.split(',') if d] or list(amm.DTYPES)):
        modes = (0, 1) if dt in amm.FLOAT_DTYPES else (0,)
        n = a.bytes // amm.DTYPE_SIZE[dt]
        ds = amm.DeviceSlice.from_torch(raw, dtype=dt, workspace=ws, stream=stream)
        for mode in modes:
            for (v, th, c) in configs:
                ws.set_tuning(th, c, v)
                for _ in range(3):
                    ds.launch(mode)
                res = ds.wait(mode)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    ds.launch(mode)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                tun = ws.get_tuning()
                row = {"dtype": dt, "mode": "return_nan" if mode else "ignore_nan", "n": n,
                       "variant": tun["variant_name"], "threads": tun["threads"],
                       "grid": ws.last_launch()["grid"], "ms": round(ms, 4),
                       "GBps": round(a.bytes / ms / 1e6, 1), "Gelem_s": round(n / ms / 1e6, 1),
                       "result": list(res)}
                rows.append(row)
                print(json.dumps(row), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
