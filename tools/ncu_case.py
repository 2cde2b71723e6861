#!/usr/bin/env python
"""One (dtype, n) case of serialised scans on cold L2 -- the command that is run plain and then
under ncu (launch list / --set full) for profiles/.  Prints one JSON line with the CUDA-event time
per launch (the plain run's number; a number printed under ncu is not a measurement)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--iters", type=int, default=12)
    a = ap.parse_args()
    es = amm.DTYPE_SIZE[a.dtype]
    nbytes = a.n * es
    copies = 1 if nbytes >= (1 << 30) else max(2, -(-(512 << 20) // nbytes))
    g = torch.Generator(device="cuda:0").manual_seed(3)
    ws = amm.Workspace(0)
    stream = torch.cuda.current_stream().cuda_stream
    bufs = []
    for _ in range(copies):
        t = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device="cuda:0")
        t.random_(-2**31, 2**31 - 1, generator=g)
        if a.dtype == "f32":
            f = t.view(torch.float32)
            f[torch.isnan(f)] = 0
        bufs.append(amm.DeviceSlice.from_torch(t.view(torch.uint8)[:nbytes], dtype=a.dtype, workspace=ws, stream=stream))
    for b in bufs[:3]:
        b.launch(a.mode)
    res = bufs[0].wait(a.mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        bufs[i % copies].launch(a.mode)
    e1.record()
    e1.synchronize()
    bufs[0].wait(a.mode)
    us = e0.elapsed_time(e1) / a.iters * 1e3
    ll = ws.last_launch()
    print(json.dumps({"dtype": a.dtype, "n": a.n, "bytes": nbytes, "us_per_launch": round(us, 2),
                      "GBps": round(nbytes / us / 1e3, 1), "grid": ll["grid"], "block": ll["threads"],
                      "tuning": ws.get_tuning(), "rotating_buffers": copies, "result": list(res)}), flush=True)


if __name__ == "__main__":
    main()
