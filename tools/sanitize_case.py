#!/usr/bin/env python
"""A small, fast set of calls for compute-sanitizer (memcheck / racecheck / initcheck): every
dtype x mode through the direct kernel, the streaming kernel (several load shapes, partial
tiles, unaligned base) and the host-slice pipeline, each checked against the oracle.
  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import argminmax_b200 as amm
from oracle import oracle
from tests.gpu_util import random_array, to_device
from tests import golden_util as G


def main():
    rng = np.random.default_rng(0)
    lib = amm.lib()
    ws = amm.Workspace(0)
    checks = 0
    for dt in G.ALL_DTYPES:
        es = np.dtype(G.NP[dt]).itemsize
        for n, off in ((1, 0), (37, es), (4099, 3 * es), (70_001, 5 * es)):
            a = random_array(dt, n, rng)
            if dt in G.FLOAT_DTYPES and n > 10:
                a[n // 2] = np.nan
            ds, _keep = to_device(a, byte_offset=off)
            ds._ws = ws
            for mode in ((0, 1) if dt in G.FLOAT_DTYPES else (0,)):
                exp = oracle.argminmax(a, mode)
                ws.set_direct_threshold(1 << 62)          # one-pass kernel
                ws.set_tuning(0, 0, -1)
                got = ds.nanargminmax() if mode else ds.argminmax()
                assert got == exp, (dt, n, mode, "direct", got, exp)
                ws.set_direct_threshold(0)                # streaming kernel, every load shape
                for v in range(lib.amm_variant_count()):
                    ws.set_tuning(128, 2, v)
                    got = ds.nanargminmax() if mode else ds.argminmax()
                    assert got == exp, (dt, n, mode, v, got, exp)
                    checks += 1
    ws.set_direct_threshold(8 << 20)
    ws.set_tuning(0, 0, -1)
    os.environ.setdefault("AMM_HOST_CHUNK_MB", "1")
    a = random_array("f32", 700_003, rng)
    assert amm.HostSlice.from_numpy(a, workspace=ws).argminmax() == oracle.argminmax(a, 0)
    print("sanitize_case ok:", checks, "streaming checks")


if __name__ == "__main__":
    main()
