"""N>1 host logic on CPU: two gloo ranks exchange 64-byte records and run the deterministic
first-index combine (amm_combine_records).  The per-shard records are produced by the ORACLE
here (no GPU), laid out exactly as the kernel lays them out; the GPU twin of this test
(tests/test_gpu_multi.py) produces them with the kernel."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _key_f32(x: np.float32) -> int:
    """canonical ord key of policies.cuh (-0 -> +0, then ((v >> 31) & 0x7fffffff) ^ v)."""
    b = int(np.array([x], dtype=np.float32).view(np.int32)[0])
    if b == -2**31:
        b = 0
    return b ^ ((b >> 31) & 0x7FFFFFFF)


def _record_for_shard(arr: np.ndarray, offset: int, mode: int):
    from argminmax_b200 import AmmRecord
    from oracle import oracle
    NO = 2**64 - 1
    if arr.size == 0:
        return AmmRecord(0, 0, 0, 0, 0, 0, 0, 0)
    nan_pos = np.flatnonzero(np.isnan(arr))
    nan_idx = int(nan_pos[0]) + offset if nan_pos.size and mode == 1 else NO
    if np.isnan(arr).all():
        return AmmRecord(0, NO, 0, NO, nan_idx, 2 if nan_idx != NO else 0, arr.size, 1)
    imin, imax = oracle.argminmax(arr, 0)
    flags = 1 | (2 if nan_idx != NO else 0)
    return AmmRecord(_key_f32(arr[imin]), imin + offset, _key_f32(arr[imax]), imax + offset,
                     nan_idx, flags, arr.size, 1)


def _worker(rank: int, world: int, port: int, case: str, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from argminmax_b200 import combine_across_ranks, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same data on every rank
        n = 100_003
        a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
        a[np.isnan(a)] = 0
        mode = 0
        if case == "dup":       # duplicate extremes in both shards: the first shard must win
            a[[10, 70_000]] = np.float32(-3.0e38)
            a[[20_000, 99_000]] = np.float32(3.0e38)
        elif case == "nan_ret":  # NaNs in both shards, return-NaN -> the first one
            a[[60_001, 4_242]] = np.nan
            mode = 1
        elif case == "nan_ign":  # second shard entirely NaN, ignore-NaN
            a[50_016:] = np.nan
        elif case == "all_nan":
            a[:] = np.nan
        b, e = shard_bounds(n, world, rank, align_elems=8)
        rec = _record_for_shard(a[b:e], b, mode)
        got = combine_across_ranks(rec, mode)
        q.put((rank, case, got, a.tobytes() if rank == 0 else None, mode))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("case", ["plain", "dup", "nan_ret", "nan_ign", "all_nan"])
def test_two_rank_gloo_exchange_and_combine(case):
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = next(r[3] for r in results if r[3] is not None)
    mode = results[0][4]
    a = np.frombuffer(data, dtype=np.float32)
    exp = oracle.argminmax(a, mode)
    for _rank, _case, got, _d, _m in results:
        assert got == exp, (case, got, exp)


def test_shard_bounds_cover_the_array():
    from argminmax_b200 import shard_bounds
    for n in (0, 1, 7, 1000, 1 << 20, (1 << 30) + 5):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = shard_bounds(n, world, r, align_elems=8)
                assert b == prev and b <= e <= n
                assert b % 8 == 0 or b == n
                prev = e
            assert prev == n
