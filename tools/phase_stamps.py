#!/usr/bin/env python
"""Where does the fixed per-call cost of a scan go?  Runs the AMM_PHASE_STAMPS twin of the library
(thread 0 of every CTA records %globaltimer at the phase boundaries) on cold-L2 rotating buffers
and prints, per (dtype, n, knobs), the phase boundaries relative to the earliest CTA entry, next
to the event-timed duration of the same serialised launches.

  AMM_LIB_PATH is set by this tool; run it in its own process:
  python tools/phase_stamps.py [--cases f32:10000000,f32:100000000,u8:100000000,f32:1073741824]
"""
import argparse
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from argminmax_b200 import _build  # noqa: E402

os.environ["AMM_LIB_PATH"] = _build.build_stamps()

import numpy as np  # noqa: E402
import torch  # noqa: E402

import argminmax_b200 as amm  # noqa: E402

NAMES = ["entry", "first_fold", "stream_done", "cta_reduced", "ticket", "grid_folded", "exact_done", "published"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="f32:10000000,f32:100000000,u8:100000000,f32:1073741824")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--tag", default="")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--direct", action="store_true", help="keep the library's kernel choice (latency kernel for small n)")
    ap.add_argument("--copies", type=int, default=0, help="rotating buffers (1 = L2-resident input); 0 = enough for 512 MB")
    a = ap.parse_args()
    lib = amm.lib()
    ws = amm.Workspace(0)
    ws.set_tuning(a.threads, a.ctas, a.variant)
    if not a.direct:
        ws.set_direct_threshold(0)  # streaming kernel only
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda:0").manual_seed(5)
    for case in a.cases.split(","):
        dt, n = case.split(":")
        n = int(n)
        es = amm.DTYPE_SIZE[dt]
        copies = max(2, -(-(512 << 20) // (n * es)))
        if n * es >= (1 << 30):
            copies = 1
        if a.copies > 0:
            copies = a.copies
        bufs = []
        for _ in range(copies):
            t = torch.empty((n * es + 3) // 4, dtype=torch.int32, device="cuda:0")
            t.random_(-2**31, 2**31 - 1, generator=g)
            if dt == "f32":
                f = t.view(torch.float32)
                f[torch.isnan(f)] = 0
            bufs.append(amm.DeviceSlice.from_torch(t.view(torch.uint8)[: n * es], dtype=dt, workspace=ws, stream=stream))
        for b in bufs[:3]:
            b.launch(0)
        bufs[0].wait(0)
        # event-timed serialised launches (the number profiles/ and RESULTS.md quote)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 200 if n * es < (1 << 30) else 20
        e0.record()
        for i in range(iters):
            bufs[i % copies].launch(0)
        e1.record()
        e1.synchronize()
        bufs[0].wait(0)
        us_events = e0.elapsed_time(e1) / iters * 1e3
        rows = []
        for r in range(a.reps):
            b = bufs[r % copies]
            b.launch(0)
            b.wait(0)
            torch.cuda.synchronize()
            grid = ws.last_launch()["grid"]
            out = (ctypes.c_uint64 * (grid * 8))()
            rc = lib.amm_debug_read_stamps(ws.handle, out, grid)
            assert rc == 0, rc
            st = np.frombuffer(out, dtype=np.uint64).reshape(grid, 8).astype(np.int64)
            t0 = st[:, 0].min()
            last = int(np.argmax(st[:, 7]))  # the CTA that published (last ticket holder, or CTA 0 of the slot flow)
            rows.append({
                "entry_spread": int(st[:, 0].max() - t0),
                "first_fold_min": int(st[:, 1][st[:, 1] > 0].min() - t0) if (st[:, 1] > 0).any() else None,
                "first_fold_max": int(st[:, 1].max() - t0) if (st[:, 1] > 0).any() else None,
                "stream_done_min": int(st[:, 2].min() - t0), "stream_done_med": int(np.median(st[:, 2]) - t0),
                "stream_done_max": int(st[:, 2].max() - t0),
                "cta_reduced_max": int(st[:, 3].max() - t0), "ticket_max": int(st[:, 4].max() - t0),
                "grid_folded": int(st[last, 5] - t0), "exact_done": int(st[last, 6] - t0),
                "published": int(st[last, 7] - t0),
            })
        med = {k: (statistics.median([r[k] for r in rows if r[k] is not None]) if any(r[k] is not None for r in rows) else None)
               for k in rows[0]}
        ll = ws.last_launch()
        print(json.dumps({"tag": a.tag, "dtype": dt, "n": n, "bytes": n * es, "grid": ll["grid"], "threads": ll["threads"],
                          "tuning": ws.get_tuning(), "us_per_launch_events_serialised": round(us_events, 2),
                          "ideal_us_at_7300GBps": round(n * es / 7.3e6, 2),
                          "ns_after_first_cta_entry_median": med}), flush=True)
        del bufs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
