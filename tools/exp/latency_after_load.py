#!/usr/bin/env python
"""Experiment: is the synchronous-call latency different right after a burst of bandwidth-bound
work?  (bench.py measured 12.9 us at 1 Mi f32 in-process after its streaming loops while a fresh
process measures 10.8 us.)"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch

import argminmax_b200 as amm

lib = amm.lib()
n_big = 1 << 30
t = torch.randint(-2**31, 2**31 - 1, (n_big,), dtype=torch.int32, device="cuda:0")
f = t.view(torch.float32)
f[torch.isnan(f)] = 0
small = torch.randint(-2**31, 2**31 - 1, (1 << 20,), dtype=torch.int32, device="cuda:0").view(torch.float32)
small[torch.isnan(small)] = 0
ws = amm.Workspace(0)
stream = torch.cuda.current_stream().cuda_stream
big = amm.DeviceSlice.from_torch(f, dtype="f32", workspace=ws, stream=stream)
out = (ctypes.c_double * 3)()


def lat(ptr, tag):
    rc = lib.amm_measure_call_latency(amm.DTYPES["f32"], ptr, 1 << 20, 0, ws.handle, stream, 1000, out)
    assert rc == 0
    print(json.dumps({"when": tag, "us_median": round(out[0], 2), "us_min": round(out[1], 2)}), flush=True)


lat(small.data_ptr(), "fresh, own 4 MiB buffer")
lat(f.data_ptr(), "fresh, first 4 MiB of the 4 GiB buffer")
for _ in range(400):
    big.launch(0)
big.wait(0)
lat(f.data_ptr(), "right after 400 x 4 GiB scans, 4 GiB buffer")
lat(small.data_ptr(), "right after, own buffer")
time.sleep(1.0)
lat(small.data_ptr(), "1 s later, own buffer")
host = torch.empty(1 << 30, dtype=torch.float32, pin_memory=True)
lat(small.data_ptr(), "after pinning 4 GiB of host memory")
del host
lat(small.data_ptr(), "after freeing it")
