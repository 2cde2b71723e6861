#!/usr/bin/env python
"""GB/s of the windowed / segmented form for fixed window lengths (kernel-only, CUDA events)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--dtypes", default="f32,u8,f16,f64")
    ap.add_argument("--windows", default="64,256,1000,4096,20000,1000000")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ragged", action="store_true",
                    help="ragged windows (offsets): lengths uniform in [W/2, 3W/2], the down-sampling-by-time case")
    ap.add_argument("--outlier", type=int, default=0, help="with --ragged: one window of this many elements among them")
    a = ap.parse_args()
    g = torch.Generator(device="cuda:0").manual_seed(3)
    t = torch.empty(a.bytes // 4, dtype=torch.int32, device="cuda:0")
    t.random_(-2**31, 2**31 - 1, generator=g)
    raw = t.view(torch.uint8)
    ws = amm.Workspace(0)
    stream = torch.cuda.current_stream().cuda_stream
    for dt in a.dtypes.split(","):
        ds = amm.DeviceSlice.from_torch(raw, dtype=dt, workspace=ws, stream=stream)
        for w in [int(x) for x in a.windows.split(",")]:
            arg = w
            if a.ragged:
                n = ds.n
                nwin = max(1, n // w)
                lens = torch.randint(max(1, w // 2), w + w // 2 + 1, (nwin,), device="cuda:0", dtype=torch.int64, generator=g)
                # scale the running sum to end exactly at n: no giant / empty windows at the end
                csum = torch.cumsum(lens, 0).to(torch.float64)
                offs = torch.zeros(nwin + 1, dtype=torch.int64, device="cuda:0")
                offs[1:] = torch.floor(csum * (n / float(csum[-1]))).to(torch.int64)
                offs[-1] = n
                if a.outlier:
                    # one window of `outlier` elements in the middle of the small ones
                    mid = nwin // 2
                    shift = min(a.outlier, n - int(offs[mid + 1]))
                    offs[mid + 1:] = torch.clamp(offs[mid + 1:] + shift, max=n)
                arg = offs
            for _ in range(2):
                out = ds.argminmax_windows(arg)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                out = ds.argminmax_windows(arg)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            print(json.dumps({"dtype": dt, "window": w, "ragged": bool(a.ragged), "outlier": a.outlier, "n_windows": out.shape[0], "ms": round(ms, 4),
                              "GBps": round(a.bytes / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
