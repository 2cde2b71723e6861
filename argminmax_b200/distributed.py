"""One rank per GPU (torch.distributed): a logical array sharded into consecutive index
ranges over the ranks of one job.

Data path per call, on the rank's compute stream, nothing synchronous in between.
Default ("peer"): ONE kernel per rank -- the scan kernel's last CTA stores the rank's 64-byte
record into every peer's exchange buffer over NVLink (CUDA-IPC-mapped memory), waits for the
other ranks' records, folds them and publishes the global answer (amm_exchange_*).
Fallback / AMM_SHARDED_EXCHANGE=nccl:
    scan kernel (64-byte record, indices already global)
 -> all_gather_into_tensor of the records   (NCCL over NVLink/NVSwitch: world x 64 B)
 -> combine kernel (ascending rank order, strict compares => first occurrence)
Either way (argmin, argmax) is read from the workspace's pinned host slot.
torch.distributed is plumbing only (rendezvous + the collective).  With the gloo backend
(CPU tests) the record exchange and the host-side combine are exercised without a GPU.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

from . import _ffi
from ._ffi import DTYPES, FLOAT_DTYPES, IGNORE_NAN, RETURN_NAN, check, lib
from .device_slice import DeviceSlice, combine_records


def shard_bounds(n_total: int, world: int, rank: int, align_elems: int = 1) -> tuple[int, int]:
    """Contiguous shard [begin, end) of rank `rank`: ceil(n/world) elements per rank, rounded up
    to `align_elems` so every shard starts on a vector boundary."""
    per = -(-n_total // world)
    per = -(-per // align_elems) * align_elems
    begin = min(n_total, per * rank)
    end = min(n_total, begin + per)
    return begin, end


def exchange_records(local: _ffi.AmmRecord, group=None) -> list[_ffi.AmmRecord]:
    """All-gather one 64-byte record per rank (host path; works with gloo and nccl)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    raw = bytes(local)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    if backend == "nccl":
        t = t.cuda()
    out = torch.empty(world * 64, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    buf = out.cpu().numpy().tobytes()
    return [_ffi.AmmRecord.from_buffer_copy(buf[i * 64:(i + 1) * 64]) for i in range(world)]


def combine_across_ranks(local: _ffi.AmmRecord, mode: int, group=None) -> tuple[int, int]:
    """Host-side: exchange records, then amm_combine_records in ascending rank order."""
    return combine_records(exchange_records(local, group), mode)


# workspaces that currently hold a peer exchange (a workspace holds at most one)
_exchange_owners: dict[int, "DistributedShardedSlice"] = {}  # keyed by the amm_workspace handle


class DistributedShardedSlice:
    """This rank's shard of a logical array; every rank must call the same methods in the same
    order (they contain a collective).  Mirrors the ArgMinMax / NaNArgMinMax surface.

    The fused exchange lives in a workspace and a workspace holds one exchange.  The slice uses
    its shard's workspace when that is free and a private one otherwise (a second slice on the
    same device), so two slices never share -- or tear down -- each other's exchange buffers.
    `close()` is a collective too: every rank unmaps its peers' buffers, a barrier, then each
    rank frees its own (CUDA requires importers to close before the exporter frees)."""

    def __init__(self, shard: DeviceSlice, group=None):
        import torch
        import torch.distributed as dist

        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", shard.device)
        lens = torch.zeros(self.world, dtype=torch.int64, device=dev)
        mine = torch.tensor([shard.n], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(lens, mine, group=group)
        lens = lens.cpu().tolist()
        self.global_offset = int(sum(lens[: self.rank]))
        self.n_total = int(sum(lens))
        self._rec = torch.zeros(64, dtype=torch.uint8, device=dev)
        self._gathered = torch.zeros(64 * self.world, dtype=torch.uint8, device=dev)
        self.exchange = "nccl"
        self._ws = shard.workspace
        self._owns_exchange = False
        want = os.environ.get("AMM_SHARDED_EXCHANGE", "peer")
        if self.world > 1 and want == "peer":
            self._setup_peer_exchange(dev)

    @property
    def workspace(self):
        """The workspace the collective scans run on (tuning / pipelined mode apply here)."""
        return self._ws

    def _setup_peer_exchange(self, dev) -> None:
        """Fused in-kernel exchange: allocate this rank's buffer, swap CUDA IPC handles and GPU
        UUIDs once (the only collective in the setup), map the peers' buffers.  Falls back to the
        NCCL all-gather path if any rank cannot set it up -- e.g. two ranks share one GPU -- and
        all ranks agree on the outcome."""
        import torch
        import torch.distributed as dist

        from .device_slice import Workspace

        if self._ws.handle.value in _exchange_owners:  # the shard's workspace already carries an exchange
            self._ws = Workspace(self.shard.device)
        ws = self._ws
        handle = (ctypes.c_uint8 * 64)()
        uuid = (ctypes.c_uint8 * 16)()
        ok = lib().amm_exchange_create(ws.handle, self.world, self.rank, handle) == _ffi.AMM_OK
        created = ok
        ok = ok and lib().amm_device_uuid(self.shard.device, uuid) == _ffi.AMM_OK
        mine = torch.tensor(list(bytes(handle)) + list(bytes(uuid)) + [1 if ok else 0],
                            dtype=torch.uint8, device=dev)
        allh = torch.zeros(81 * self.world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        rows = allh.view(self.world, 81).cpu()
        ok = bool(rows[:, 80].min().item() == 1)
        if ok:
            table = bytes(rows[:, :64].contiguous().numpy().tobytes())
            ids = bytes(rows[:, 64:80].contiguous().numpy().tobytes())
            buf = (ctypes.c_uint8 * len(table)).from_buffer_copy(table)
            idbuf = (ctypes.c_uint8 * len(ids)).from_buffer_copy(ids)
            ok = lib().amm_exchange_connect_ipc(ws.handle, buf, idbuf) == _ffi.AMM_OK
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.exchange = "peer"
            self._owns_exchange = True
            _exchange_owners[ws.handle.value] = self
        elif created:
            # some rank failed: unmap whatever was mapped, then free (same order as close())
            lib().amm_exchange_disconnect(ws.handle)
            dist.barrier(group=self.group)
            lib().amm_exchange_destroy(ws.handle)

    def set_timeout(self, milliseconds: int) -> None:
        """How long the in-kernel exchange waits for a peer's record (0 = for ever, like NCCL)."""
        check(lib().amm_exchange_set_timeout(self._ws.handle, milliseconds), "amm_exchange_set_timeout")

    def close(self) -> None:
        """Collective teardown of the fused exchange (no-op on the NCCL path)."""
        import torch
        import torch.distributed as dist

        if not self._owns_exchange:
            return
        self._owns_exchange = False
        _exchange_owners.pop(self._ws.handle.value, None)
        torch.cuda.synchronize(self.shard.device)
        lib().amm_exchange_disconnect(self._ws.handle)
        dist.barrier(group=self.group)
        lib().amm_exchange_destroy(self._ws.handle)
        self.exchange = "nccl"

    def __len__(self):
        return self.n_total

    def launch(self, mode: int = IGNORE_NAN) -> None:
        """Enqueue scan -> all-gather -> combine on the current stream; no host sync."""
        import torch
        import torch.distributed as dist

        stream = torch.cuda.current_stream().cuda_stream
        s = self.shard
        if self.exchange == "peer":
            rc = lib().amm_argminmax_launch_exchange(DTYPES[s.dtype], s.data_ptr, s.n, mode,
                                                     self.global_offset, self._ws.handle, stream)
            check(rc, "amm_argminmax_launch_exchange")
            return
        if s.n == 0:
            self._rec.zero_()  # no value, no NaN
        else:
            rc = lib().amm_argminmax_launch(DTYPES[s.dtype], s.data_ptr, s.n, mode,
                                            self.global_offset, self._ws.handle, stream,
                                            self._rec.data_ptr())
            check(rc, "amm_argminmax_launch")
        dist.all_gather_into_tensor(self._gathered, self._rec, group=self.group)
        rc = lib().amm_combine_records_launch(self._gathered.data_ptr(), self.world, mode,
                                              self._ws.handle, stream)
        check(rc, "amm_combine_records_launch")

    def wait(self, mode: int = IGNORE_NAN) -> tuple[int, int]:
        import torch
        out = (ctypes.c_uint64 * 2)()
        check(lib().amm_workspace_wait(self._ws.handle,
                                       torch.cuda.current_stream().cuda_stream, mode, out),
              "amm_workspace_wait")
        return int(out[0]), int(out[1])

    def _run(self, mode: int) -> tuple[int, int]:
        if self.n_total == 0:
            raise _ffi.EmptyInputError(_ffi.AMM_ERR_EMPTY, "DistributedShardedSlice")
        self.launch(mode)
        return self.wait(mode)

    def argminmax(self):
        return self._run(IGNORE_NAN)

    def argmin(self):
        return self._run(IGNORE_NAN)[0]

    def argmax(self):
        return self._run(IGNORE_NAN)[1]

    def nanargminmax(self):
        if self.shard.dtype not in FLOAT_DTYPES:
            raise TypeError("NaNArgMinMax is only implemented for float dtypes")
        return self._run(RETURN_NAN)

    def nanargmin(self):
        return self.nanargminmax()[0]

    def nanargmax(self):
        return self.nanargminmax()[1]
