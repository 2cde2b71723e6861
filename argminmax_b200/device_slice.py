"""Host-side mirror of the reference's trait surface on device-resident data.

`DeviceSlice` plays the role of `impl ArgMinMax for DeviceSlice<T>` /
`impl NaNArgMinMax for DeviceSlice<T>` (the Rust shim in rust/src/cuda/mod.rs binds
the same C symbols): same method names, argument meaning and error behaviour as
src/lib.rs:134-244 -- `argminmax/argmin/argmax` ignore NaNs, `nanargminmax/
nanargmin/nanargmax` return the first NaN and exist for float dtypes only, empty
input raises (the reference panics).  PyTorch appears only as the owner of device
memory and streams; every result comes from the CUDA library through the C ABI.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

from . import _ffi
from ._ffi import DTYPES, DTYPE_SIZE, FLOAT_DTYPES, IGNORE_NAN, RETURN_NAN, check, lib


def _torch_dtype_name(t) -> str:
    import torch
    table = {
        torch.int8: "i8", torch.int16: "i16", torch.int32: "i32", torch.int64: "i64",
        torch.uint8: "u8", torch.float16: "f16", torch.float32: "f32", torch.float64: "f64",
    }
    for name, short in (("uint16", "u16"), ("uint32", "u32"), ("uint64", "u64")):
        if hasattr(torch, name):
            table[getattr(torch, name)] = short
    if t.dtype not in table:
        raise TypeError(f"unsupported tensor dtype {t.dtype}")
    return table[t.dtype]


class Workspace:
    """amm_workspace: per-device scratch + pinned result slot (one per concurrent caller)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        check(lib().amm_workspace_create(device, ctypes.byref(self._h)), "amm_workspace_create")
        self.device = device

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().amm_workspace_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def set_tuning(self, threads: int = 0, ctas_per_sm: int = 0, variant: int = -1):
        check(lib().amm_workspace_set_tuning(self._h, threads, ctas_per_sm, variant),
              "amm_workspace_set_tuning")

    def set_pipelined(self, on: bool = True):
        """Let consecutive scans on one stream overlap (see amm_workspace_set_pipelined)."""
        check(lib().amm_workspace_set_pipelined(self._h, 1 if on else 0),
              "amm_workspace_set_pipelined")

    def set_direct_threshold(self, nbytes: int):
        check(lib().amm_workspace_set_direct_threshold(self._h, nbytes),
              "amm_workspace_set_direct_threshold")

    def get_tuning(self) -> dict:
        t, c, v, s = (ctypes.c_int() for _ in range(4))
        check(lib().amm_workspace_get_tuning(self._h, t, c, v, s), "amm_workspace_get_tuning")
        return {"threads": t.value, "ctas_per_sm": c.value, "variant": v.value,
                "variant_name": lib().amm_variant_name(v.value).decode(), "sm_count": s.value}

    def last_launch(self) -> dict:
        g, t = ctypes.c_int(), ctypes.c_int()
        total = ctypes.c_uint64()
        check(lib().amm_workspace_last_launch(self._h, g, t, total), "amm_workspace_last_launch")
        return {"grid": g.value, "threads": t.value, "launches_total": total.value}

    def host_record(self) -> _ffi.AmmRecord:
        p = ctypes.POINTER(_ffi.AmmRecord)()
        check(lib().amm_workspace_host_record(self._h, ctypes.byref(p)), "amm_workspace_host_record")
        return p.contents

    def device_record_ptr(self) -> int:
        p = ctypes.c_void_p()
        check(lib().amm_workspace_device_record(self._h, ctypes.byref(p)),
              "amm_workspace_device_record")
        return p.value


_default_ws: dict[int, Workspace] = {}


def default_workspace(device: int) -> Workspace:
    if device not in _default_ws:
        _default_ws[device] = Workspace(device)
    return _default_ws[device]


class DeviceSlice:
    """A contiguous 1-D array of `n` elements of `dtype` in the HBM of one B200."""

    def __init__(self, data_ptr: int, n: int, dtype: str, device: int = 0,
                 workspace: Optional[Workspace] = None, stream: int = 0, owner=None):
        if dtype not in DTYPES:
            raise TypeError(f"unknown dtype {dtype!r}")
        self.data_ptr = int(data_ptr)
        self.n = int(n)
        self.dtype = dtype
        self.device = device
        self.stream = stream
        self._owner = owner  # keeps the backing allocation alive
        self._ws = workspace

    @classmethod
    def from_torch(cls, tensor, dtype: Optional[str] = None, workspace: Optional[Workspace] = None,
                   stream: int = 0) -> "DeviceSlice":
        """Wrap a CUDA tensor (no copy).  `dtype` reinterprets the bytes (e.g. a uint8 tensor
        holding u32 data); like the reference's ndarray adapter (src/lib.rs:730) this refuses
        non-contiguous input."""
        if not tensor.is_cuda:
            raise ValueError("DeviceSlice needs a CUDA tensor (use HostSlice for host memory)")
        if tensor.dim() != 1 or not tensor.is_contiguous():
            raise ValueError("DeviceSlice needs a contiguous 1-D tensor")
        name = dtype or _torch_dtype_name(tensor)
        nbytes = tensor.numel() * tensor.element_size()
        if nbytes % DTYPE_SIZE[name]:
            raise ValueError("byte length is not a multiple of the element size")
        return cls(tensor.data_ptr(), nbytes // DTYPE_SIZE[name], name, tensor.device.index or 0,
                   workspace, stream, owner=tensor)

    @classmethod
    def from_dlpack(cls, obj, dtype: Optional[str] = None, workspace: Optional[Workspace] = None,
                    stream: int = 0) -> "DeviceSlice":
        """Any producer of a CUDA DLPack capsule (`__dlpack__`): cupy, jax, numba, torch ..."""
        import torch
        return cls.from_torch(torch.from_dlpack(obj), dtype=dtype, workspace=workspace, stream=stream)

    @classmethod
    def from_cuda_array_interface(cls, obj, workspace: Optional[Workspace] = None,
                                  stream: int = 0, device: int = 0) -> "DeviceSlice":
        """Objects exposing `__cuda_array_interface__` (numba / cupy device arrays), no copy."""
        cai = obj.__cuda_array_interface__
        shape, typestr = cai["shape"], cai["typestr"]
        if len(shape) != 1 or cai.get("strides") not in (None, (int(typestr[2:]),)):
            raise ValueError("DeviceSlice needs a contiguous 1-D array")
        kinds = {"i": "i", "u": "u", "f": "f"}
        if typestr[0] not in "<|" or typestr[1] not in kinds:
            raise TypeError(f"unsupported typestr {typestr}")
        name = f"{kinds[typestr[1]]}{int(typestr[2:]) * 8}"
        return cls(cai["data"][0], shape[0], name, device, workspace, stream, owner=obj)

    def __len__(self):
        return self.n

    @property
    def workspace(self) -> Workspace:
        return self._ws or default_workspace(self.device)

    def _run(self, mode: int) -> tuple[int, int]:
        out = (ctypes.c_uint64 * 2)()
        rc = lib().amm_argminmax(DTYPES[self.dtype], self.data_ptr, self.n, mode,
                                 self.workspace.handle, self.stream, out)
        check(rc, f"amm_argminmax_{self.dtype}")
        return int(out[0]), int(out[1])

    def launch(self, mode: int = IGNORE_NAN, global_offset: int = 0, record_ptr: int = 0) -> None:
        """Enqueue only (amm_argminmax_launch); pair with `wait`."""
        rc = lib().amm_argminmax_launch(DTYPES[self.dtype], self.data_ptr, self.n, mode,
                                        global_offset, self.workspace.handle, self.stream,
                                        record_ptr or None)
        check(rc, "amm_argminmax_launch")

    def wait(self, mode: int = IGNORE_NAN) -> tuple[int, int]:
        out = (ctypes.c_uint64 * 2)()
        check(lib().amm_workspace_wait(self.workspace.handle, self.stream, mode, out),
              "amm_workspace_wait")
        return int(out[0]), int(out[1])

    # --- windowed / segmented form ---------------------------
    def argminmax_windows(self, windows, mode: int = IGNORE_NAN):
        """(n_windows, 2) int64 CUDA tensor of absolute (argmin, argmax) per window, one launch.
        `windows`: an int (fixed window length) or a 1-D int64/uint64 CUDA tensor of n_windows+1
        ascending offsets.  Empty windows give -1 (AMM_NO_INDEX)."""
        import torch
        dev = torch.device("cuda", self.device)
        if isinstance(windows, int):
            nwin = -(-self.n // windows)
            out = torch.empty((nwin, 2), dtype=torch.int64, device=dev)
            rc = lib().amm_argminmax_fixed_windows(DTYPES[self.dtype], self.data_ptr, self.n,
                                                   windows, mode, self.workspace.handle,
                                                   self.stream, out.data_ptr())
        else:
            if windows.dim() != 1 or windows.element_size() != 8 or not windows.is_cuda:
                raise ValueError("offsets must be a 1-D 64-bit integer CUDA tensor")
            nwin = windows.numel() - 1
            out = torch.empty((nwin, 2), dtype=torch.int64, device=dev)
            rc = lib().amm_argminmax_windows(DTYPES[self.dtype], self.data_ptr, self.n,
                                             windows.data_ptr(), nwin, mode,
                                             self.workspace.handle, self.stream, out.data_ptr())
        check(rc, "amm_argminmax_windows")
        return out

    # --- trait ArgMinMax (src/lib.rs:134-186) ---------------------------------
    def argminmax(self) -> tuple[int, int]:
        return self._run(IGNORE_NAN)

    def argmin(self) -> int:
        return self._run(IGNORE_NAN)[0]

    def argmax(self) -> int:
        return self._run(IGNORE_NAN)[1]

    # --- trait NaNArgMinMax (src/lib.rs:194-244), floats only -----------------
    def _need_float(self):
        if self.dtype not in FLOAT_DTYPES:
            raise TypeError(f"NaNArgMinMax is only implemented for float dtypes, not {self.dtype}")

    def nanargminmax(self) -> tuple[int, int]:
        self._need_float()
        return self._run(RETURN_NAN)

    def nanargmin(self) -> int:
        return self.nanargminmax()[0]

    def nanargmax(self) -> int:
        return self.nanargminmax()[1]


class HostSlice:
    """`&[T]` / Vec<T> in host memory (src/lib.rs:679-712): chunked H2D copies overlapped
    with the device scan (amm_host_argminmax).  PCIe-bound; pin the memory for full speed."""

    def __init__(self, host_ptr: int, n: int, dtype: str, device: int = 0,
                 workspace: Optional[Workspace] = None, owner=None):
        if dtype not in DTYPES:
            raise TypeError(f"unknown dtype {dtype!r}")
        self.host_ptr, self.n, self.dtype, self.device = int(host_ptr), int(n), dtype, device
        self._ws, self._owner = workspace, owner

    @classmethod
    def from_numpy(cls, arr, device: int = 0, workspace: Optional[Workspace] = None) -> "HostSlice":
        import numpy as np
        from_np = {np.dtype(k): v for k, v in {
            "int8": "i8", "int16": "i16", "int32": "i32", "int64": "i64", "uint8": "u8",
            "uint16": "u16", "uint32": "u32", "uint64": "u64", "float16": "f16",
            "float32": "f32", "float64": "f64"}.items()}
        if arr.ndim != 1 or not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("HostSlice needs a contiguous 1-D array")
        return cls(arr.ctypes.data, arr.size, from_np[arr.dtype], device, workspace, owner=arr)

    @classmethod
    def from_arrow(cls, arr, device: int = 0, workspace: Optional[Workspace] = None) -> "HostSlice":
        """Arrow `PrimitiveArray` -> its values buffer (the reference's arrow adapter,
        src/lib.rs:764-810: `self.values()`, i.e. the validity bitmap is IGNORED and null
        slots are scanned with whatever value they hold).  Zero-copy; honours the array offset."""
        import pyarrow as pa
        if isinstance(arr, pa.ChunkedArray):
            if arr.num_chunks != 1:
                raise ValueError("HostSlice.from_arrow needs a single contiguous chunk")
            arr = arr.chunk(0)
        names = {pa.int8(): "i8", pa.int16(): "i16", pa.int32(): "i32", pa.int64(): "i64",
                 pa.uint8(): "u8", pa.uint16(): "u16", pa.uint32(): "u32", pa.uint64(): "u64",
                 pa.float16(): "f16", pa.float32(): "f32", pa.float64(): "f64"}
        if arr.type not in names:
            raise TypeError(f"unsupported arrow type {arr.type}")
        name = names[arr.type]
        values = arr.buffers()[1]
        ptr = values.address + arr.offset * DTYPE_SIZE[name]
        return cls(ptr, len(arr), name, device, workspace, owner=(arr, values))

    @classmethod
    def from_torch(cls, tensor, dtype: Optional[str] = None, device: int = 0,
                   workspace: Optional[Workspace] = None) -> "HostSlice":
        if tensor.is_cuda or tensor.dim() != 1 or not tensor.is_contiguous():
            raise ValueError("HostSlice needs a contiguous 1-D CPU tensor")
        name = dtype or _torch_dtype_name(tensor)
        nbytes = tensor.numel() * tensor.element_size()
        return cls(tensor.data_ptr(), nbytes // DTYPE_SIZE[name], name, device, workspace,
                   owner=tensor)

    @property
    def workspace(self) -> Workspace:
        return self._ws or default_workspace(self.device)

    def _run(self, mode: int) -> tuple[int, int]:
        out = (ctypes.c_uint64 * 2)()
        rc = lib().amm_host_argminmax(DTYPES[self.dtype], self.host_ptr, self.n, mode,
                                      self.workspace.handle, out)
        check(rc, "amm_host_argminmax")
        return int(out[0]), int(out[1])

    def argminmax(self):
        return self._run(IGNORE_NAN)

    def argmin(self):
        return self._run(IGNORE_NAN)[0]

    def argmax(self):
        return self._run(IGNORE_NAN)[1]

    def nanargminmax(self):
        if self.dtype not in FLOAT_DTYPES:
            raise TypeError(f"NaNArgMinMax is only implemented for float dtypes, not {self.dtype}")
        return self._run(RETURN_NAN)

    def nanargmin(self):
        return self.nanargminmax()[0]

    def nanargmax(self):
        return self.nanargminmax()[1]


class ShardedDeviceSlice:
    """One logical array split into consecutive index ranges over the GPUs of ONE box,
    driven from this process (amm_sharded_*): per-GPU scan, one NCCL all-gather of the
    64-byte records, deterministic first-index combine."""

    def __init__(self, shards: Sequence[DeviceSlice]):
        if not shards:
            raise ValueError("need at least one shard")
        if len({s.dtype for s in shards}) != 1:
            raise TypeError("all shards must share one dtype")
        self.shards = list(shards)
        self.dtype = shards[0].dtype
        devs = (ctypes.c_int * len(shards))(*[s.device for s in shards])
        self._h = ctypes.c_void_p()
        check(lib().amm_sharded_create(len(shards), devs, ctypes.byref(self._h)),
              "amm_sharded_create")

    def close(self):
        if self._h:
            lib().amm_sharded_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def uses_nccl(self) -> bool:
        return bool(lib().amm_sharded_uses_nccl(self._h))

    @property
    def uses_peer_exchange(self) -> bool:
        return bool(lib().amm_sharded_uses_peer_exchange(self._h))

    def __len__(self):
        return sum(s.n for s in self.shards)

    def _run(self, mode: int) -> tuple[int, int]:
        g = len(self.shards)
        ptrs = (ctypes.c_void_p * g)(*[s.data_ptr for s in self.shards])
        lens = (ctypes.c_uint64 * g)(*[s.n for s in self.shards])
        out = (ctypes.c_uint64 * 2)()
        check(lib().amm_sharded_argminmax(self._h, DTYPES[self.dtype], ptrs, lens, mode, out),
              "amm_sharded_argminmax")
        return int(out[0]), int(out[1])

    def argminmax(self):
        return self._run(IGNORE_NAN)

    def argmin(self):
        return self._run(IGNORE_NAN)[0]

    def argmax(self):
        return self._run(IGNORE_NAN)[1]

    def nanargminmax(self):
        if self.dtype not in FLOAT_DTYPES:
            raise TypeError(f"NaNArgMinMax is only implemented for float dtypes, not {self.dtype}")
        return self._run(RETURN_NAN)

    def nanargmin(self):
        return self.nanargminmax()[0]

    def nanargmax(self):
        return self.nanargminmax()[1]


def combine_records(records: Sequence[_ffi.AmmRecord], mode: int) -> tuple[int, int]:
    """amm_combine_records on host records given in ascending shard order."""
    arr = (_ffi.AmmRecord * len(records))(*records)
    out = (ctypes.c_uint64 * 2)()
    check(lib().amm_combine_records(arr, len(records), mode, out), "amm_combine_records")
    return int(out[0]), int(out[1])
