"""Sharded paths on real GPUs.  With one GPU the in-process ShardedDeviceSlice is exercised
with several shards on the same device (records then travel through the pinned host slots,
NCCL refuses duplicate devices); with >= 2 GPUs the NCCL all-gather path runs, in-process
(amm_sharded_*) and as one rank per GPU under torch.distributed."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.gpu_util import random_array

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shards_on(arr, cuts, devices):
    import torch

    from argminmax_b200 import DeviceSlice
    from tests.gpu_util import NP_NAME
    name = NP_NAME[arr.dtype]
    out, keep = [], []
    bounds = [0, *cuts, arr.size]
    for g, dev in enumerate(devices):
        part = np.ascontiguousarray(arr[bounds[g]:bounds[g + 1]])
        t = torch.from_numpy(part.view(np.uint8).copy()).to(f"cuda:{dev}")
        if t.numel() == 0:
            t = torch.zeros(16, dtype=torch.uint8, device=f"cuda:{dev}")
            ds = DeviceSlice(t.data_ptr(), 0, name, device=dev, owner=t)
        else:
            ds = DeviceSlice.from_torch(t, dtype=name)
        keep.append(t)
        out.append(ds)
    return out, keep


@pytest.mark.parametrize("dt", ["f32", "u8", "i64", "f16", "f64"])
def test_sharded_same_device(dt):
    from argminmax_b200 import ShardedDeviceSlice
    rng = np.random.default_rng(31)
    n = 3_000_001
    a = random_array(dt, n, rng)
    for cuts in ([n // 2], [1, 2], [1_000_000, 1_000_000, 2_500_000], [n - 1]):
        shards, _keep = _shards_on(a, cuts, [0] * (len(cuts) + 1))
        sh = ShardedDeviceSlice(shards)
        assert sh.argminmax() == oracle.argminmax(a, 0), cuts
        if dt in G.FLOAT_DTYPES:
            b = a.copy()
            b[[n - 2, 1_700_000]] = np.nan
            shards, _keep2 = _shards_on(b, cuts, [0] * (len(cuts) + 1))
            sh2 = ShardedDeviceSlice(shards)
            assert sh2.nanargminmax() == (1_700_000, 1_700_000)
            assert sh2.argminmax() == oracle.argminmax(b, 0)
            sh2.close()
        sh.close()


def test_sharded_duplicates_first_shard_wins():
    from argminmax_b200 import ShardedDeviceSlice
    a = np.ones(2_000_000, dtype=np.float32)
    a[[500_000, 1_500_000]] = -7.0
    a[[700_000, 1_200_000]] = 9.0
    shards, _keep = _shards_on(a, [1_000_000], [0, 0])
    sh = ShardedDeviceSlice(shards)
    assert sh.argminmax() == (500_000, 700_000)
    sh.close()


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_sharded_multi_gpu_in_process(exchange, monkeypatch):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    from argminmax_b200 import ShardedDeviceSlice
    monkeypatch.setenv("AMM_SHARDED_EXCHANGE", exchange)
    g = min(_ngpu(), 8)
    rng = np.random.default_rng(41)
    n = 8_000_003
    a = random_array("f32", n, rng)
    cuts = [n * (i + 1) // g for i in range(g - 1)]
    shards, _keep = _shards_on(a, cuts, list(range(g)))
    sh = ShardedDeviceSlice(shards)
    assert sh.uses_peer_exchange if exchange == "peer" else sh.uses_nccl
    for _ in range(5):
        assert sh.argminmax() == oracle.argminmax(a, 0)
    a[[n - 1, cuts[0] + 5]] = np.nan
    shards, _keep2 = _shards_on(a, cuts, list(range(g)))
    sh2 = ShardedDeviceSlice(shards)
    assert sh2.nanargminmax() == (cuts[0] + 5, cuts[0] + 5)
    assert sh2.argminmax() == oracle.argminmax(a, 0)
    sh.close()
    sh2.close()
    # an empty shard in the middle still takes part in the exchange
    b = random_array("i16", 1_000_000, rng)
    cuts2 = [400_000] + [400_000] * (g - 2) if g > 2 else [1_000_000]
    shards, _keep3 = _shards_on(b, cuts2, list(range(g)))
    sh3 = ShardedDeviceSlice(shards)
    assert sh3.argminmax() == oracle.argminmax(b, 0)
    sh3.close()


_RANK_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import argminmax_b200 as amm
from oracle import oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
rng = np.random.default_rng(77)
n = 6_000_011
a = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32).copy()
a[np.isnan(a)] = 0
a[[10, n - 10]] = np.float32(-3.3e38)      # duplicate min in the first and the last shard
for mode, b in ((0, a), (1, None)):
    if b is None:
        b = a.copy(); b[[n // 2 + 1, n - 3]] = np.nan
    lo, hi = amm.shard_bounds(n, world, rank, align_elems=8)
    t = torch.from_numpy(b[lo:hi].copy()).cuda()
    ds = amm.DistributedShardedSlice(amm.DeviceSlice.from_torch(t))
    assert ds.exchange == os.environ["AMM_SHARDED_EXCHANGE"], ds.exchange
    got = ds.nanargminmax() if mode else ds.argminmax()
    exp = oracle.argminmax(b, mode)
    assert got == exp, (rank, mode, got, exp)
    for _ in range(4):   # back-to-back collective calls (parity banks of the exchange buffer)
        assert ds.argminmax() == oracle.argminmax(b, 0)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_distributed_one_rank_per_gpu(tmp_path, exchange):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    g = min(_ngpu(), 8)
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    env = dict(os.environ, AMM_SHARDED_EXCHANGE=exchange)
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
