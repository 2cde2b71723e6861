#!/usr/bin/env python
"""N ranks, one per GPU, all streaming 4 GiB of pinned host memory to their GPU at once: where does
the end-to-end host-slice call (amm_host_argminmax, chunked H2D || scan) lose against plain
copies?  Prints, per configuration, the aggregate GB/s (sum over ranks of each rank's own rate)
and the lock-step rate (N x bytes / slowest rank).
  torchrun --nproc-per-node N tools/exp/h2d_multi.py"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch
import torch.distributed as dist

import argminmax_b200 as amm


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    host.uniform_(-1, 1)
    dst = torch.empty(1 << 28, dtype=torch.float32, device=dev)

    def rates(seconds):
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        allt = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, t)
        ts = allt.cpu().tolist()
        return sum(n * 4 / x / 1e9 for x in ts), world * n * 4 / max(ts) / 1e9, [round(n * 4 / x / 1e9, 1) for x in ts]

    def plain(chunk_elems, nstreams):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        k = 0
        for i in range(0, n, chunk_elems):
            m = min(chunk_elems, n - i)
            off = (k % nstreams) * (dst.numel() // nstreams)
            with torch.cuda.stream(streams[k % nstreams]):
                dst[off:off + m].copy_(host[i:i + m], non_blocking=True)
            k += 1
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    out = []
    for label, chunk, ns in (("plain 1 GiB copies, 1 stream", 1 << 28, 1), ("plain 32 MiB copies, 2 streams", 1 << 23, 2),
                             ("plain 128 MiB copies, 2 streams", 1 << 25, 2), ("plain 32 MiB copies, 1 stream", 1 << 23, 1)):
        best = min(plain(chunk, ns) for _ in range(3))
        s, lock, per = rates(best)
        out.append({"what": label, "sum_gbs": round(s, 1), "lockstep_gbs": round(lock, 1), "per_rank": per})
    for chunk_mb in (32, 128, 512):
        os.environ["AMM_HOST_CHUNK_MB"] = str(chunk_mb)
        ws = amm.Workspace(rank)
        hs = amm.HostSlice.from_torch(host, dtype="f32", device=rank, workspace=ws)
        rec = amm.AmmRecord()
        def step():
            rc = amm.lib().amm_host_argminmax_record(amm.DTYPES["f32"], hs.host_ptr, hs.n, 0, 0, ws.handle, ctypes.byref(rec))
            assert rc == 0, rc
        step()
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            step()
            best = min(best, time.perf_counter() - t0)
        s, lock, per = rates(best)
        out.append({"what": f"amm_host_argminmax_record, {chunk_mb} MiB chunks", "sum_gbs": round(s, 1),
                    "lockstep_gbs": round(lock, 1), "per_rank": per})
        ws.close()
    if rank == 0:
        for o in out:
            print(json.dumps({"gpus": world, **o}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
