#!/usr/bin/env python
"""How fast is amm_host_argminmax on PAGEABLE host memory (what a Vec<T> / numpy array is)
compared with pinned memory?"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np
import torch

import argminmax_b200 as amm

n = 1 << 28  # 1 GiB of f32
rng = np.random.default_rng(0)
a = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32)
a[np.isnan(a)] = 0
ws = amm.Workspace(0)
out = (ctypes.c_uint64 * 2)()
L = amm.lib()


def run(ptr, tag, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = L.amm_host_argminmax(amm.DTYPES["f32"], ptr, n, 0, ws.handle, out)
        ts.append(time.perf_counter() - t0)
        assert rc == 0
    print(json.dumps({"memory": tag, "GBps_best": round(n * 4 / min(ts) / 1e9, 2),
                      "GBps_median": round(n * 4 / sorted(ts)[len(ts) // 2] / 1e9, 2),
                      "result": [out[0], out[1]]}), flush=True)


run(a.ctypes.data, "pageable (numpy)")
pinned = torch.empty(n, dtype=torch.float32, pin_memory=True)
pinned.numpy()[:] = a
run(pinned.data_ptr(), "pinned (cudaHostAlloc)")
run(a.ctypes.data, "pageable again")
