"""ctypes binding of include/argminmax_b200.h.  There is NO fallback: if the CUDA
library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os

from . import _build

DTYPES = {
    "i8": 0, "i16": 1, "i32": 2, "i64": 3,
    "u8": 4, "u16": 5, "u32": 6, "u64": 7,
    "f16": 8, "f32": 9, "f64": 10,
}
DTYPE_SIZE = {"i8": 1, "u8": 1, "i16": 2, "u16": 2, "f16": 2, "i32": 4, "u32": 4, "f32": 4,
              "i64": 8, "u64": 8, "f64": 8}
FLOAT_DTYPES = ("f16", "f32", "f64")
IGNORE_NAN = 0
RETURN_NAN = 1

AMM_OK, AMM_ERR_EMPTY, AMM_ERR_ARG, AMM_ERR_ALIGN, AMM_ERR_CUDA, AMM_ERR_NCCL, AMM_ERR_TOO_LARGE, AMM_ERR_PEER = range(8)


class AmmRecord(ctypes.Structure):
    _fields_ = [
        ("min_key", ctypes.c_int64), ("min_idx", ctypes.c_uint64),
        ("max_key", ctypes.c_int64), ("max_idx", ctypes.c_uint64),
        ("nan_idx", ctypes.c_uint64), ("flags", ctypes.c_uint64),
        ("n", ctypes.c_uint64), ("seq", ctypes.c_uint64),
    ]


class AmmError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib().amm_status_string(status).decode()
        if status == AMM_ERR_CUDA:
            msg += f" [cudaError {lib().amm_last_cuda_error()}]"
        super().__init__(f"{where}: {msg}")


class EmptyInputError(AmmError, ValueError):
    """n == 0 -- the reference panics (assert!(!arr.is_empty()), src/scalar/generic.rs:182)."""


_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_u64 = ctypes.c_uint64

# every symbol include/argminmax_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "amm_workspace_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "amm_workspace_destroy": (_i, [_vp]),
    "amm_argminmax": (_i, [_i, _vp, _u64, _i, _vp, _vp, _u64p]),
    **{f"amm_argminmax_{d}": (_i, [_vp, _u64, _i, _vp, _vp, _u64p]) for d in DTYPES},
    "amm_argmin": (_i, [_i, _vp, _u64, _i, _vp, _vp, _u64p]),
    "amm_argmax": (_i, [_i, _vp, _u64, _i, _vp, _vp, _u64p]),
    "amm_argminmax_launch": (_i, [_i, _vp, _u64, _i, _u64, _vp, _vp, _vp]),
    "amm_workspace_wait": (_i, [_vp, _vp, _i, _u64p]),
    "amm_workspace_device_record": (_i, [_vp, ctypes.POINTER(_vp)]),
    "amm_workspace_host_record": (_i, [_vp, ctypes.POINTER(ctypes.POINTER(AmmRecord))]),
    "amm_combine_records": (_i, [ctypes.POINTER(AmmRecord), _i, _i, _u64p]),
    "amm_exchange_create": (_i, [_vp, _i, _i, _vp]),
    "amm_exchange_connect_ipc": (_i, [_vp, _vp, _vp]),
    "amm_exchange_set_timeout": (_i, [_vp, _u64]),
    "amm_exchange_disconnect": (_i, [_vp]),
    "amm_exchange_destroy": (_i, [_vp]),
    "amm_device_uuid": (_i, [_i, _vp]),
    "amm_debug_read_stamps": (_i, [_vp, _u64p, _i]),
    "amm_debug_plan_scan": (_i, [_u64, _u64, _i, _i, _i, _i, _u64, _i, _i, _i, _i, _u64p]),
    "amm_debug_plan_direct": (_i, [_u64, _u64, _i, _i, _i, _i, _i, _i, _u64p]),
    "amm_exchange_connect_ptrs": (_i, [_vp, ctypes.POINTER(_vp)]),
    "amm_exchange_local_buffer": (_i, [_vp, ctypes.POINTER(_vp)]),
    "amm_argminmax_launch_exchange": (_i, [_i, _vp, _u64, _i, _u64, _vp, _vp]),
    "amm_host_argminmax": (_i, [_i, _vp, _u64, _i, _vp, _u64p]),
    "amm_host_argminmax_record": (_i, [_i, _vp, _u64, _i, _u64, _vp, ctypes.POINTER(AmmRecord)]),
    "amm_combine_records_launch": (_i, [_vp, _i, _i, _vp, _vp]),
    "amm_sharded_create": (_i, [_i, ctypes.POINTER(_i), ctypes.POINTER(_vp)]),
    "amm_sharded_destroy": (_i, [_vp]),
    "amm_sharded_argminmax": (_i, [_vp, _i, ctypes.POINTER(_vp), _u64p, _i, _u64p]),
    "amm_sharded_uses_nccl": (_i, [_vp]),
    "amm_sharded_uses_peer_exchange": (_i, [_vp]),
    "amm_sharded_workspace": (_i, [_vp, _i, ctypes.POINTER(_vp)]),
    "amm_argminmax_windows": (_i, [_i, _vp, _u64, _vp, _u64, _i, _vp, _vp, _vp]),
    "amm_argminmax_fixed_windows": (_i, [_i, _vp, _u64, _u64, _i, _vp, _vp, _vp]),
    "amm_device_count": (_i, [ctypes.POINTER(_i)]),
    "amm_device_alloc": (_i, [_i, ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "amm_device_free": (_i, [_i, _vp]),
    "amm_copy_to_device": (_i, [_i, _vp, _vp, ctypes.c_size_t]),
    "amm_copy_to_host": (_i, [_i, _vp, _vp, ctypes.c_size_t]),
    "amm_status_string": (ctypes.c_char_p, [_i]),
    "amm_last_cuda_error": (_i, []),
    "amm_dtype_size": (ctypes.c_size_t, [_i]),
    "amm_version": (_i, []),
    "amm_workspace_set_tuning": (_i, [_vp, _i, _i, _i]),
    "amm_workspace_set_direct_threshold": (_i, [_vp, _u64]),
    "amm_workspace_set_pipelined": (_i, [_vp, _i]),
    "amm_workspace_get_tuning": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i),
                                      ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "amm_variant_count": (_i, []),
    "amm_variant_name": (ctypes.c_char_p, [_i]),
    "amm_measure_call_latency": (_i, [_i, _vp, _u64, _i, _vp, _vp, _i, ctypes.POINTER(ctypes.c_double)]),
    "amm_measure_exchange_latency": (_i, [_i, _vp, _u64, _i, _u64, _vp, _vp, _i, ctypes.POINTER(ctypes.c_double)]),
    "amm_workspace_last_launch": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), _u64p]),
    "amm_debug_plan_window_chunks": (_i, [_u64p, _u64p, _u64, ctypes.c_size_t, _i, _i, _u64p, _u64, _u64p,
                                          _u64p]),
}


def lib_path() -> str:
    return _build.LIB_PATH


def lib():
    """Load libargminmax_b200.so (building it first if nvcc is here and it is stale)."""
    global _lib
    if _lib is None:
        path = os.environ.get("AMM_LIB_PATH") or _build.LIB_PATH  # override: A/B experiments only
        if path == _build.LIB_PATH and not _build.is_current():
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001 - re-raised below if nothing usable exists
                if not os.path.exists(path):
                    raise RuntimeError(
                        "argminmax_b200: the CUDA library is not built and cannot be built here "
                        f"({e}); there is no CPU fallback") from e
        cdll = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(cdll, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = cdll
    return _lib


def check(status: int, where: str) -> None:
    if status == AMM_OK:
        return
    if status == AMM_ERR_EMPTY:
        raise EmptyInputError(status, where)
    raise AmmError(status, where)
