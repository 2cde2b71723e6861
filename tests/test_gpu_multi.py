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


def test_in_process_sharded_leaves_the_current_device_alone():
    """Creating, calling and destroying a ShardedDeviceSlice walks over all devices inside the
    library; the caller's current device must be what it was -- also when it IS device 0 (the
    guard once only restored when it had switched on entry: bench.py's rank 0 came back from
    the in-process check with cuda:1 current)."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch

    from argminmax_b200 import ShardedDeviceSlice
    g = min(_ngpu(), 8)
    rng = np.random.default_rng(47)
    a = random_array("f32", 100_000 * g, rng)
    cuts = [a.size * (i + 1) // g for i in range(g - 1)]
    for cur in (0, g - 1):
        torch.cuda.set_device(cur)
        shards, _keep = _shards_on(a, cuts, list(range(g)))
        assert torch.cuda.current_device() == cur
        sh = ShardedDeviceSlice(shards)
        assert torch.cuda.current_device() == cur
        assert sh.argminmax() == oracle.argminmax(a, 0)
        assert torch.cuda.current_device() == cur
        sh.close()
        assert torch.cuda.current_device() == cur
    torch.cuda.set_device(0)


def test_in_process_bad_shard_is_rejected_before_any_launch(monkeypatch):
    """a shard that fails validation (here: a misaligned pointer on the LAST device)
    must not leave the first devices running a collective the others never join -- the call fails
    up front and the very next call on the same handle works."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    import ctypes

    import argminmax_b200 as amm
    from argminmax_b200 import ShardedDeviceSlice
    monkeypatch.setenv("AMM_SHARDED_EXCHANGE", "peer")
    g = min(_ngpu(), 8)
    rng = np.random.default_rng(43)
    n = 1_000_000 * g
    a = random_array("f32", n, rng)
    cuts = [n * (i + 1) // g for i in range(g - 1)]
    shards, _keep = _shards_on(a, cuts, list(range(g)))
    sh = ShardedDeviceSlice(shards)
    assert sh.uses_peer_exchange
    assert sh.argminmax() == oracle.argminmax(a, 0)
    ptrs = (ctypes.c_void_p * g)(*[s.data_ptr for s in shards])
    lens = (ctypes.c_uint64 * g)(*[s.n for s in shards])
    out = (ctypes.c_uint64 * 2)()
    ptrs[g - 1] = shards[g - 1].data_ptr + 2          # f32 at an odd 2-byte address
    rc = amm.lib().amm_sharded_argminmax(sh._h, amm.DTYPES["f32"], ptrs, lens, 0, out)
    assert rc == 3                                     # AMM_ERR_ALIGN, nothing was launched
    ptrs[g - 1] = None                                 # NULL with n > 0
    assert amm.lib().amm_sharded_argminmax(sh._h, amm.DTYPES["f32"], ptrs, lens, 0, out) == 2
    for _ in range(3):                                 # the handle is still in step
        assert sh.argminmax() == oracle.argminmax(a, 0)
    assert sh.uses_peer_exchange
    sh.close()


_TIMEOUT_SCRIPT = r"""
import os, sys, time
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import argminmax_b200 as amm
from oracle import oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
rng = np.random.default_rng(5)
n = 2_000_000
a = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32).copy()
a[np.isnan(a)] = 0
lo, hi = amm.shard_bounds(n, world, rank, align_elems=8)
t = torch.from_numpy(a[lo:hi].copy()).cuda()
shard = amm.DeviceSlice.from_torch(t)
ds = amm.DistributedShardedSlice(shard)
assert ds.exchange == "peer"
assert ds.argminmax() == oracle.argminmax(a, 0)
# a second slice on the same device gets its own exchange (the first one stays usable)
ds2 = amm.DistributedShardedSlice(shard)
assert ds2.exchange == "peer" and ds2.workspace is not ds.workspace
assert ds2.argminmax() == ds.argminmax() == oracle.argminmax(a, 0)
ds2.close()
assert ds.argminmax() == oracle.argminmax(a, 0)
# the LAST rank shows up 1.5 s late with a 200 ms timeout: EVERY rank must fail, not just the
# early ones (the late rank finds the early ranks' poisoned records)
ds.set_timeout(200)
dist.barrier()
if rank == world - 1:
    time.sleep(1.5)
failed = False
try:
    ds.argminmax()
except amm.AmmError as e:
    failed = e.status == 7
flags = torch.tensor([1 if failed else 0], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
assert int(flags.item()) == 1, (rank, failed)
# the exchange is unusable until re-created ...
try:
    ds.argminmax()
    raise SystemExit("a broken exchange accepted a call")
except amm.AmmError as e:
    assert e.status == 7
ds.close()
# ... and a fresh one works
ds3 = amm.DistributedShardedSlice(shard)
assert ds3.exchange == "peer"
assert ds3.argminmax() == oracle.argminmax(a, 0)
ds3.close()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_distributed_exchange_ownership_and_collective_timeout(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    g = min(_ngpu(), 8)
    script = tmp_path / "rank_timeout.py"
    script.write_text(_TIMEOUT_SCRIPT.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}",
           "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)]
    env = dict(os.environ, AMM_SHARDED_EXCHANGE="peer")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]


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
