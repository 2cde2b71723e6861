#!/usr/bin/env python
"""In-process A/B of two builds of the library on the synchronous-call latency: both .so files are
loaded side by side (each with its own workspace), and for every (dtype, n) the two are timed in
alternation (A, B, A, B, ...), `rounds` times `calls` calls each; the medians of the per-round
medians are printed.  Process-to-process and box-to-box drift (+-0.5 us) cancels out.
  python tools/exp/build_prev.py HEAD && python tools/exp/latency_ab.py"""
import argparse
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DT = {"i8": 0, "i16": 1, "i32": 2, "i64": 3, "u8": 4, "u16": 5, "u32": 6, "u64": 7, "f16": 8, "f32": 9, "f64": 10}


def load(path, env=""):
    """Load `path` and create a workspace with the `K=V,K=V` environment overrides in force (the
    library reads its knobs when a workspace is created)."""
    saved = {}
    for kv in [x for x in env.split(",") if x]:
        k, v = kv.split("=", 1)
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        return _load(path)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _load(path):
    lib = ctypes.CDLL(path)
    lib.amm_workspace_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.amm_workspace_destroy.argtypes = [ctypes.c_void_p]
    lib.amm_measure_call_latency.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    ws = ctypes.c_void_p()
    rc = lib.amm_workspace_create(0, ctypes.byref(ws))
    assert rc == 0, rc
    return lib, ws


def main():
    from argminmax_b200 import _build, DTYPES
    ap = argparse.ArgumentParser()
    ap.add_argument("--a", default=os.path.join(_build.LIB_DIR, "libargminmax_b200_prev.so"))
    ap.add_argument("--b", default=_build.LIB_PATH)
    ap.add_argument("--dtypes", default="f32,u8,f16,i64,f64")
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=22)
    ap.add_argument("--calls", type=int, default=300)
    ap.add_argument("--rounds", type=int, default=7)
    ap.add_argument("--env-a", default="", help="K=V,K=V knobs in force while A's workspace is created")
    ap.add_argument("--env-b", default="")
    a = ap.parse_args()
    t = torch.randint(-2**31, 2**31 - 1, ((8 << a.max_log2) // 4 + 4,), dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    la, wa = load(a.a, a.env_a)
    lb, wb = load(a.b, a.env_b)
    out = (ctypes.c_double * 3)()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for dt in a.dtypes.split(","):
        if a.mode == 1 and dt not in ("f16", "f32", "f64"):
            continue
        row = {"dtype": dt, "mode": a.mode, "a": os.path.basename(a.a) + " " + a.env_a, "b": os.path.basename(a.b) + " " + a.env_b, "us_a": {}, "us_b": {}, "b_minus_a": {}}
        for lg in range(a.min_log2, a.max_log2 + 1):
            n = 1 << lg
            ra, rb = [], []
            for _ in range(a.rounds):
                for lib, ws, acc in ((la, wa, ra), (lb, wb, rb)):
                    rc = lib.amm_measure_call_latency(DTYPES[dt], t.data_ptr(), n, a.mode, ws, stream, a.calls, out)
                    assert rc == 0, rc
                    acc.append(out[0])
            ma, mb = statistics.median(ra), statistics.median(rb)
            row["us_a"][str(lg)] = round(ma, 2)
            row["us_b"][str(lg)] = round(mb, 2)
            row["b_minus_a"][str(lg)] = round(mb - ma, 2)
        print(json.dumps(row), flush=True)
    la.amm_workspace_destroy(wa)
    lb.amm_workspace_destroy(wb)


if __name__ == "__main__":
    main()
