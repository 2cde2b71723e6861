#!/usr/bin/env python
"""us per synchronous amm_argminmax call, timed INSIDE the library (amm_measure_call_latency: no
binding overhead), for several dtypes and n = 2^10 .. 2^max.  Environment knobs of the library
(AMM_DIRECT_BYTES, AMM_DIRECT_VPT, ...) apply, so two runs give an A/B of the small-input kernels."""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtypes", default="f32,u8,i16,f16,i64,f64")
    ap.add_argument("--modes", default="0")
    ap.add_argument("--calls", type=int, default=500)
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=23)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    lib = amm.lib()
    t = torch.randint(-2**31, 2**31 - 1, ((8 << a.max_log2) // 4 + 4,), dtype=torch.int32, device="cuda:0")
    raw = t.view(torch.uint8)
    ws = amm.Workspace(0)
    stream = torch.cuda.current_stream().cuda_stream
    out = (ctypes.c_double * 3)()
    for dt in a.dtypes.split(","):
        for mode in [int(m) for m in a.modes.split(",")]:
            if mode == 1 and dt not in ("f16", "f32", "f64"):
                continue
            row = {"tag": a.tag, "dtype": dt, "mode": mode, "us_median": {}}
            for lg in range(a.min_log2, a.max_log2 + 1):
                n = 1 << lg
                rc = lib.amm_measure_call_latency(amm.DTYPES[dt], raw.data_ptr(), n, mode, ws.handle,
                                                  stream, a.calls, out)
                if rc:
                    raise RuntimeError(f"rc={rc}")
                row["us_median"][str(lg)] = round(out[0], 2)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
