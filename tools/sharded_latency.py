#!/usr/bin/env python
"""Small-n latency of ONE logical array sharded over G GPUs of one box, driven from one host
thread (amm_sharded_*: per-GPU scan fused with the in-kernel peer exchange, or NCCL all-gather
with AMM_SHARDED_EXCHANGE=nccl).  BASELINE config 5's "1K-1M latency sweep at G = 8"."""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import argminmax_b200 as amm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--calls", type=int, default=500)
    ap.add_argument("--max-log2", type=int, default=24)
    a = ap.parse_args()
    g = a.gpus
    for lg in range(10, a.max_log2 + 1, 2):
        n = 1 << lg
        per = n // g
        shards, keep = [], []
        for d in range(g):
            t = torch.rand(per, dtype=torch.float32, device=f"cuda:{d}")
            keep.append(t)
            shards.append(amm.DeviceSlice.from_torch(t))
        sh = amm.ShardedDeviceSlice(shards)
        for _ in range(30):
            sh.argminmax()
        lat = []
        for _ in range(a.calls):
            t0 = time.perf_counter()
            sh.argminmax()
            lat.append((time.perf_counter() - t0) * 1e6)
        print(json.dumps({"gpus": g, "n_total": per * g, "exchange": "peer" if sh.uses_peer_exchange else
                          ("nccl" if sh.uses_nccl else "host"),
                          "us_call_median": round(statistics.median(lat), 2), "us_call_min": round(min(lat), 2)}),
              flush=True)
        sh.close()


if __name__ == "__main__":
    main()
