#!/usr/bin/env python
"""Tuning sweep on one B200: GB/s of the scan kernel for every load shape x CTA size x CTAs/SM.
Kernel-only timing with CUDA events on the launching stream (input resident in HBM, larger
than L2).  Usage: python tools/sweep.py [--dtype f32] [--n 1073741824] [--iters 10]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def make_data(dtype: str, n: int):
    es = amm.DTYPE_SIZE[dtype]
    g = torch.Generator(device="cuda:0").manual_seed(1)
    words = (n * es + 3) // 4
    t = torch.randint(-2**31, 2**31 - 1, (words,), dtype=torch.int32, device="cuda:0", generator=g)
    if dtype == "f32":
        f = t.view(torch.float32)
        f[torch.isnan(f)] = 0
    elif dtype == "f64":
        f = t[: (words // 2) * 2].view(torch.float64)
        f[torch.isnan(f)] = 0
    elif dtype == "f16":
        f = t.view(torch.float16)
        f[torch.isnan(f)] = 0
    return t.view(torch.uint8)[: n * es]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--n", type=int, default=1 << 30)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--variants", default="")
    ap.add_argument("--threads", default="128,256,512")
    ap.add_argument("--ctas", default="1,2,3,4,6,8")
    ap.add_argument("--out", default="")
    ap.add_argument("--rotate", type=int, default=1,
                    help="cycle through R distinct copies of the input (R x bytes > L2 => cold)")
    a = ap.parse_args()
    lib = amm.lib()
    data = make_data(a.dtype, a.n)
    ws = amm.Workspace(0)
    ws.set_direct_threshold(0)  # this tool sweeps the streaming kernel only
    ws.set_pipelined(bool(int(os.environ.get("AMM_PIPELINED", "0"))))
    copies = [data] + [data.clone() for _ in range(a.rotate - 1)]
    dss = [amm.DeviceSlice.from_torch(c, dtype=a.dtype, workspace=ws,
                                      stream=torch.cuda.current_stream().cuda_stream) for c in copies]
    ds = dss[0]
    nbytes = a.n * amm.DTYPE_SIZE[a.dtype]
    variants = [int(v) for v in a.variants.split(",")] if a.variants else range(lib.amm_variant_count())
    rows = []
    ref = None
    for v in variants:
        for th in [int(x) for x in a.threads.split(",")]:
            for c in [int(x) for x in a.ctas.split(",")]:
                ws.set_tuning(th, c, v)
                for _ in range(3):
                    ds.launch(a.mode)
                res = ds.wait(a.mode)
                if ref is None:
                    ref = res
                assert res == ref, (res, ref)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for it in range(a.iters):
                    dss[it % len(dss)].launch(a.mode)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                ll = ws.last_launch()
                row = {"variant": lib.amm_variant_name(v).decode(), "threads": th, "ctas_per_sm": c,
                       "grid": ll["grid"], "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}
                rows.append(row)
                print(json.dumps(row), flush=True)
    best = max(rows, key=lambda r: r["GBps"])
    print("BEST", json.dumps(best))
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"dtype": a.dtype, "n": a.n, "mode": a.mode, "rows": rows, "best": best}, f, indent=1)


if __name__ == "__main__":
    main()
