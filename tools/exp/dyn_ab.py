#!/usr/bin/env python
"""NEEDS tools/exp/dynamic_tile_claims.patch applied (it adds amm_workspace_set_dynamic): the
product keeps the static schedule, see profiles/r01d_schedule_experiment.md.

A/B of the big-input scan's tile schedule (static grid-stride vs dynamic claims) inside ONE
process on the SAME buffers, interleaved in short blocks so that box / allocation / thermal
differences cancel.  Prints one JSON line per (buffer, schedule, mode) with the per-block
GB/s samples.  Usage: python tools/exp/dyn_ab.py [--blocks 8] [--launches 50] [--buffers 2]"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch

import argminmax_b200 as amm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--launches", type=int, default=50)
    ap.add_argument("--buffers", type=int, default=2)
    ap.add_argument("--n", type=int, default=1 << 30)
    ap.add_argument("--batches", default="4,8")
    a = ap.parse_args()
    stream = torch.cuda.current_stream().cuda_stream
    ws = amm.Workspace(0)
    scheds = [("static", False, 0)] + [(f"dyn{b}", True, int(b)) for b in a.batches.split(",")]
    bufs = []
    for k in range(a.buffers):
        g = torch.Generator(device="cuda:0").manual_seed(k)
        t = torch.randint(-2**31, 2**31 - 1, (a.n,), dtype=torch.int32, device="cuda:0", generator=g)
        f = t.view(torch.float32)
        f[torch.isnan(f)] = 0
        bufs.append(f)
        if k + 1 < a.buffers:  # shift the next allocation's placement
            bufs.append(None)
            _pad = torch.empty((37 << 20) + 4096 * k, dtype=torch.uint8, device="cuda:0")
            bufs[-1] = _pad
    data = [b for b in bufs if b is not None and b.dtype == torch.float32]
    nbytes = a.n * 4
    for bi, f in enumerate(data):
        ds = amm.DeviceSlice.from_torch(f, dtype="f32", workspace=ws, stream=stream)
        want = None
        samples = {(s[0], m): [] for s in scheds for m in ("pipelined", "serialised")}
        for blk in range(a.blocks):
            for name, on, batch in scheds:
                ws.set_dynamic(on, batch)
                for mode in ("pipelined", "serialised"):
                    ws.set_pipelined(mode == "pipelined")
                    for _ in range(3):
                        ds.launch(amm.IGNORE_NAN)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.launches):
                        ds.launch(amm.IGNORE_NAN)
                    e1.record()
                    e1.synchronize()
                    got = ds.wait(amm.IGNORE_NAN)
                    want = want or got
                    assert got == want, (name, mode, got, want)
                    samples[(name, mode)].append(nbytes * a.launches / e0.elapsed_time(e1) / 1e6)
        for (name, mode), v in samples.items():
            print(json.dumps({"buffer": bi, "ptr": hex(f.data_ptr()), "sched": name, "mode": mode,
                              "GBps_median": round(statistics.median(v), 1), "GBps_min": round(min(v), 1),
                              "GBps_max": round(max(v), 1), "samples": [round(x) for x in v]}), flush=True)


if __name__ == "__main__":
    main()
