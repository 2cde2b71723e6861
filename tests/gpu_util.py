"""Helpers for the -m gpu tests: move numpy arrays into B200 HBM (torch is only the
allocator here) and call the CUDA path through the C ABI."""
from __future__ import annotations

import numpy as np

from oracle import oracle
from tests import golden_util as G

NP_NAME = {np.dtype(v): k for k, v in G.NP.items()}


def to_device(arr: np.ndarray, byte_offset: int = 0):
    """-> (DeviceSlice over a device copy of `arr`, owner tensor).  `byte_offset` shifts the
    start of the data inside a larger allocation (must be a multiple of the element size) so
    that unaligned heads are exercised."""
    import torch

    from argminmax_b200 import DeviceSlice

    arr = np.ascontiguousarray(arr)
    name = NP_NAME[arr.dtype]
    raw = torch.from_numpy(arr.view(np.uint8).copy())
    buf = torch.empty(raw.numel() + byte_offset + 64, dtype=torch.uint8, device="cuda:0")
    view = buf[byte_offset:byte_offset + raw.numel()]
    view.copy_(raw)
    torch.cuda.synchronize()
    return DeviceSlice.from_torch(view, dtype=name), buf


def random_array(dt: str, n: int, rng: np.random.Generator, f16_bits: bool = False) -> np.ndarray:
    """The reference's bench/test distributions (dev_utils/src/utils.rs:29-84), seeded:
    ints uniform over the full range; f32/f64 uniform BIT PATTERNS with NaN -> 0.0; f16 uniform
    by value (or, with f16_bits, uniform bit patterns with NaN -> 0.0 to stress subnormals/-0)."""
    T = G.NP[dt]
    if dt == "f16" and not f16_bits:
        return rng.uniform(-65504.0, 65504.0, n).astype(np.float16)
    if dt == "f16":
        a = rng.integers(0, 2**16, n, dtype=np.uint16).view(np.float16).copy()
    elif dt == "f32":
        a = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32).copy()
    elif dt == "f64":
        a = rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64).copy()
    else:
        info = np.iinfo(T)
        return rng.integers(info.min, info.max, n, dtype=T, endpoint=True)
    a[np.isnan(a)] = 0
    return a


def modes_for(dt: str):
    return (0, 1) if dt in G.FLOAT_DTYPES else (0,)


def run_gpu(ds, mode: int):
    return ds.nanargminmax() if mode == 1 else ds.argminmax()


_path_ws = {}


def _workspace_for(path: str):
    """'default': size-based choice; 'stream': always the streaming (tile) kernel, automatic
    tile partition; 'contig' / 'gridstride' / 'dynamic': the streaming kernel with the
    contiguous-range / static grid-stride / grid-stride + claimed-tail partition forced (the last
    with tiny tiles so that small test arrays reach it); 'ring': the TMA-ring kernel of the
    HBM-sized inputs for every size; 'direct': the one-trip latency kernel whenever the input fits."""
    import os

    from argminmax_b200 import Workspace
    if path not in _path_ws:
        # the tile partition of the streaming kernel is a creation-time knob of the workspace
        knobs = {"contig": {"AMM_SCAN_CONTIG": "1"},
                 "gridstride": {"AMM_SCAN_CONTIG": "0", "AMM_SCAN_DYN_PCT": "0"},
                 "dynamic": {"AMM_SCAN_CONTIG": "0", "AMM_SCAN_DYN_PCT": "50"},
                 "ring": {"AMM_SCAN_RING": "1"}}.get(path, {})
        old = {k: os.environ.get(k) for k in knobs}
        os.environ.update(knobs)
        try:
            ws = Workspace(0)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        if path == "dynamic":
            # 2 KiB tiles, one CTA per SM: the claimed tail is in play from ~600 KB upwards
            ws.set_tuning(64, 1, 4)
        if path in ("stream", "contig", "gridstride", "dynamic", "ring"):
            ws.set_direct_threshold(0)
        elif path == "direct":
            ws.set_direct_threshold(1 << 62)
        _path_ws[path] = ws
    return _path_ws[path]


def assert_parity(arr: np.ndarray, mode: int, byte_offset: int = 0, what: str = "",
                  paths=("default", "stream", "direct", "contig", "gridstride", "dynamic", "ring")):
    ds, _keep = to_device(arr, byte_offset)
    exp = oracle.argminmax(arr, mode)
    got = None
    for path in paths:
        ds._ws = _workspace_for(path)
        got = run_gpu(ds, mode)
        assert got == exp, (f"{what} dtype={arr.dtype} n={arr.size} mode={mode} off={byte_offset} "
                            f"path={path}: gpu {got} != oracle {exp}")
    return got
