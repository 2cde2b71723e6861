#!/usr/bin/env python
"""Measurements for the BASELINE.md section-4 table on ONE B200 (kernel-only, CUDA events):
  config 1: f32 x 10M, cold L2 (8 rotating buffers) and warm L2, serialised and pipelined;
            plus the reference bench size n = 102 400
  config 2: all 11 dtypes x 100M elements, both NaN modes (uniform random bit patterns)
Prints JSON lines."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def timed(slices, mode, iters, ws, pipelined):
    ws.set_pipelined(pipelined)
    for i in range(3):
        slices[i % len(slices)].launch(mode)
    slices[0].wait(mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        slices[i % len(slices)].launch(mode)
    e1.record()
    e1.synchronize()
    slices[0].wait(mode)
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    g = torch.Generator(device="cuda:0").manual_seed(11)
    ws = amm.Workspace(0)
    stream = torch.cuda.current_stream().cuda_stream

    def rand_bytes(nbytes):
        t = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device="cuda:0")
        t.random_(-2**31, 2**31 - 1, generator=g)
        return t.view(torch.uint8)[:nbytes]

    # ---- config 1
    for n in (102_400, 10_000_000):
        bufs = []
        for _ in range(8):
            raw = rand_bytes(n * 4)
            f = raw.view(torch.float32)
            f[torch.isnan(f)] = 0
            bufs.append(amm.DeviceSlice.from_torch(raw, dtype="f32", workspace=ws, stream=stream))
        for label, sl in (("cold_L2_8_rotating_buffers", bufs), ("warm_L2_same_buffer", bufs[:1])):
            for pip in (False, True):
                us = timed(sl, 0, 400, ws, pip)
                print(json.dumps({"config": 1, "dtype": "f32", "n": n, "cache": label,
                                  "pipelined": pip, "us_per_call": round(us, 2),
                                  "GBps": round(n * 4 / us / 1e3, 1)}), flush=True)
        del bufs
    # ---- config 2
    n = 100_000_000
    for dt in amm.DTYPES:
        es = amm.DTYPE_SIZE[dt]
        # rotate over enough distinct buffers that consecutive scans cannot hit in the 126 MB L2
        copies = max(1, -(-(512 << 20) // (n * es)))
        raws = [rand_bytes(n * es) for _ in range(copies)]
        dss = [amm.DeviceSlice.from_torch(r, dtype=dt, workspace=ws, stream=stream) for r in raws]
        for mode in ((0, 1) if dt in amm.FLOAT_DTYPES else (0,)):
            row = {"config": 2, "dtype": dt, "mode": "return_nan" if mode else "ignore_nan", "n": n,
                   "bytes": n * es, "rotating_buffers": copies}
            for pip in (False, True):
                us = timed(dss, mode, 60, ws, pip)
                row["us_pipelined" if pip else "us_serialised"] = round(us, 2)
                row["GBps_pipelined" if pip else "GBps_serialised"] = round(n * es / us / 1e3, 1)
            row["Gelem_per_s_serialised"] = round(n / row["us_serialised"] / 1e3, 1)
            print(json.dumps(row), flush=True)
        del raws, dss


if __name__ == "__main__":
    main()
