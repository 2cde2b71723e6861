"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under argminmax_b200/ does.

`argminmax(arr, mode)` runs oracle/argminmax_oracle.c (the scalar restatement of
the reference, src/scalar/generic.rs:181-217 and src/scalar/scalar_f16.rs:20-136)
on a numpy array; f16 arrays are passed as their raw uint16 bits.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# dtype codes shared with include/argminmax_b200.h (enum amm_dtype)
DTYPE_CODES = {
    "i8": 0, "i16": 1, "i32": 2, "i64": 3,
    "u8": 4, "u16": 5, "u32": 6, "u64": 7,
    "f16": 8, "f32": 9, "f64": 10,
}
NP_TO_NAME = {
    np.dtype(np.int8): "i8", np.dtype(np.int16): "i16", np.dtype(np.int32): "i32",
    np.dtype(np.int64): "i64", np.dtype(np.uint8): "u8", np.dtype(np.uint16): "u16",
    np.dtype(np.uint32): "u32", np.dtype(np.uint64): "u64", np.dtype(np.float16): "f16",
    np.dtype(np.float32): "f32", np.dtype(np.float64): "f64",
}
IGNORE_NAN = 0
RETURN_NAN = 1


def build(force: bool = False) -> None:
    """Compile liboracle.so / libcpusimd.so with the committed Makefile (gcc only)."""
    targets = [os.path.join(_HERE, "liboracle.so"), os.path.join(_HERE, "libcpusimd.so"),
               os.path.join(_HERE, "libsimdemul.so")]
    srcs = [os.path.join(_HERE, "argminmax_oracle.c"), os.path.join(_HERE, "cpu_simd_f32.c"),
            os.path.join(_HERE, "simd_emul.c")]
    stale = force or any(
        (not os.path.exists(t)) or os.path.getmtime(t) < os.path.getmtime(s)
        for t, s in zip(targets, srcs)
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)


_lib = None
_simd = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        for fn in ("oracle_argminmax",):
            f = getattr(_lib, fn)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                          ctypes.POINTER(ctypes.c_uint64)]
        for fn in ("oracle_argmin", "oracle_argmax"):
            f = getattr(_lib, fn)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                          ctypes.POINTER(ctypes.c_uint64)]
        _lib.oracle_argminmax_windows.restype = ctypes.c_int
        _lib.oracle_argminmax_windows.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                                  ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    return _lib


def _load_simd():
    global _simd
    if _simd is None:
        build()
        _simd = ctypes.CDLL(os.path.join(_HERE, "libcpusimd.so"))
        _simd.cpu_simd_isa.restype = ctypes.c_int
        _simd.cpu_simd_argminmax_f32_ignore.restype = ctypes.c_int
        _simd.cpu_simd_argminmax_f32_ignore.argtypes = [
            ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
        _simd.cpu_simd_argminmax_f32_ignore_mt.restype = ctypes.c_int
        _simd.cpu_simd_argminmax_f32_ignore_mt.argtypes = [
            ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
    return _simd


def dtype_name(arr: np.ndarray) -> str:
    return NP_TO_NAME[arr.dtype]


def _prep(arr: np.ndarray):
    arr = np.ascontiguousarray(arr)
    if arr.ndim != 1:
        raise ValueError("oracle expects a 1-D array")
    return arr, DTYPE_CODES[NP_TO_NAME[arr.dtype]]


def argminmax(arr: np.ndarray, mode: int = IGNORE_NAN) -> tuple[int, int]:
    """(argmin, argmax) per the reference's scalar semantics. Raises on empty input
    (the reference panics, src/scalar/generic.rs:182)."""
    arr, code = _prep(arr)
    out = (ctypes.c_uint64 * 2)()
    rc = _load().oracle_argminmax(code, arr.ctypes.data, arr.size, mode, out)
    if rc == 1:
        raise ValueError("empty input")
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return int(out[0]), int(out[1])


def argminmax_windows(arr: np.ndarray, windows, mode: int = IGNORE_NAN) -> np.ndarray:
    """(n_windows, 2) int64 array of absolute (argmin, argmax) per window -- `argminmax` called
    once per window, in C.  `windows`: an int (fixed length, last window may be short) or an
    array of n_windows + 1 ascending offsets.  Empty windows give -1."""
    arr, code = _prep(arr)
    if isinstance(windows, (int, np.integer)):
        nwin = -(-arr.size // int(windows))
        offs, window = None, int(windows)
    else:
        offs = np.ascontiguousarray(windows, dtype=np.uint64)
        nwin, window = offs.size - 1, 0
    out = np.empty((nwin, 2), dtype=np.uint64)
    rc = _load().oracle_argminmax_windows(code, arr.ctypes.data, arr.size,
                                          offs.ctypes.data if offs is not None else None, window, nwin, mode,
                                          out.ctypes.data)
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return out.view(np.int64)


def argmin(arr: np.ndarray, mode: int = IGNORE_NAN) -> int:
    arr, code = _prep(arr)
    out = (ctypes.c_uint64 * 1)()
    rc = _load().oracle_argmin(code, arr.ctypes.data, arr.size, mode, out)
    if rc:
        raise ValueError("empty input" if rc == 1 else f"oracle error {rc}")
    return int(out[0])


def argmax(arr: np.ndarray, mode: int = IGNORE_NAN) -> int:
    arr, code = _prep(arr)
    out = (ctypes.c_uint64 * 1)()
    rc = _load().oracle_argmax(code, arr.ctypes.data, arr.size, mode, out)
    if rc:
        raise ValueError("empty input" if rc == 1 else f"oracle error {rc}")
    return int(out[0])


_emul = None


def simd_emul_argminmax(arr: np.ndarray, mode: int = IGNORE_NAN, isa: int = 0) -> tuple[int, int]:
    """Lane-exact emulation of the reference's DISPATCHED SIMD path (oracle/simd_emul.c):
    isa 0 = what an AVX-512 host runs, isa 1 = what an AVX2 host runs."""
    global _emul
    if _emul is None:
        build()
        _emul = ctypes.CDLL(os.path.join(_HERE, "libsimdemul.so"))
        _emul.simd_emul_argminmax.restype = ctypes.c_int
        _emul.simd_emul_argminmax.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                              ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_uint64)]
    arr, code = _prep(arr)
    out = (ctypes.c_uint64 * 2)()
    rc = _emul.simd_emul_argminmax(code, arr.ctypes.data, arr.size, mode, isa, out)
    if rc:
        raise ValueError("empty input" if rc == 1 else f"simd_emul error {rc}")
    return int(out[0]), int(out[1])


ISA_NAMES = {0: "scalar", 1: "avx", 2: "avx512f"}


def cpu_simd_isa() -> str:
    """ISA the reference's dispatch ladder (src/lib.rs:427-443) would select here."""
    return ISA_NAMES[_load_simd().cpu_simd_isa()]


def cpu_simd_argminmax_f32(arr: np.ndarray, force_isa: int = -1) -> tuple[int, int]:
    """The reference's dispatched SIMD path for &[f32] ArgMinMax, single thread."""
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    out = (ctypes.c_uint64 * 2)()
    rc = _load_simd().cpu_simd_argminmax_f32_ignore(arr.ctypes.data, arr.size, force_isa, out)
    if rc:
        raise ValueError("empty input" if rc == 1 else f"cpu_simd error {rc}")
    return int(out[0]), int(out[1])


def cpu_simd_argminmax_f32_mt(arr: np.ndarray, nthreads: int) -> tuple[int, int]:
    """NOT the reference: contiguous shards over `nthreads` pthreads + ordered combine."""
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    out = (ctypes.c_uint64 * 2)()
    rc = _load_simd().cpu_simd_argminmax_f32_ignore_mt(arr.ctypes.data, arr.size, nthreads, out)
    if rc:
        raise ValueError("empty input" if rc == 1 else f"cpu_simd error {rc}")
    return int(out[0]), int(out[1])
