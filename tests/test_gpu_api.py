"""C-ABI surface details on the GPU: typed symbols, single-output entry points, stream
ordering, independent workspaces, error statuses, fuzzed special values."""
import ctypes

import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.gpu_util import assert_parity, modes_for, random_array, to_device

pytestmark = pytest.mark.gpu


def test_typed_symbols_and_single_output_entry_points():
    """amm_argminmax_<dt> (one per `impl ArgMinMax for &[T]`, src/lib.rs:657-664) and
    amm_argmin / amm_argmax (src/lib.rs:170,185,227,243) agree with the generic entry point."""
    from argminmax_b200 import DTYPES, default_workspace, lib
    L = lib()
    rng = np.random.default_rng(55)
    ws = default_workspace(0)
    for dt in G.ALL_DTYPES:
        a = random_array(dt, 70_001, rng)
        if dt in G.FLOAT_DTYPES:
            a[[5000, 60_000]] = np.nan
        ds, _keep = to_device(a)
        for mode in modes_for(dt):
            exp = oracle.argminmax(a, mode)
            out = (ctypes.c_uint64 * 2)()
            rc = getattr(L, f"amm_argminmax_{dt}")(ds.data_ptr, ds.n, mode, ws.handle, None, out)
            assert rc == 0 and (out[0], out[1]) == exp
            one = (ctypes.c_uint64 * 1)()
            assert L.amm_argmin(DTYPES[dt], ds.data_ptr, ds.n, mode, ws.handle, None, one) == 0
            assert one[0] == exp[0]
            assert L.amm_argmax(DTYPES[dt], ds.data_ptr, ds.n, mode, ws.handle, None, one) == 0
            assert one[0] == exp[1]


def test_error_statuses():
    from argminmax_b200 import _ffi, default_workspace, lib
    import torch
    L = lib()
    ws = default_workspace(0)
    t = torch.zeros(64, dtype=torch.float32, device="cuda:0")
    out = (ctypes.c_uint64 * 2)()
    assert L.amm_argminmax(9, t.data_ptr(), 0, 0, ws.handle, None, out) == _ffi.AMM_ERR_EMPTY
    assert L.amm_argminmax(99, t.data_ptr(), 4, 0, ws.handle, None, out) == _ffi.AMM_ERR_ARG
    assert L.amm_argminmax(9, t.data_ptr(), 4, 7, ws.handle, None, out) == _ffi.AMM_ERR_ARG
    assert L.amm_argminmax(9, None, 4, 0, ws.handle, None, out) == _ffi.AMM_ERR_ARG
    assert L.amm_argminmax(9, t.data_ptr() + 2, 4, 0, ws.handle, None, out) == _ffi.AMM_ERR_ALIGN
    assert L.amm_argminmax(9, t.data_ptr(), 4, 0, None, None, out) == _ffi.AMM_ERR_ARG
    h = ctypes.c_void_p()
    assert L.amm_workspace_create(1000, ctypes.byref(h)) == _ffi.AMM_ERR_ARG
    # waiting on a workspace that never launched anything is an argument error, not a hang
    from argminmax_b200 import Workspace
    fresh = Workspace(0)
    assert L.amm_workspace_wait(fresh.handle, None, 0, out) == _ffi.AMM_ERR_ARG
    fresh.close()


def test_stream_ordering_without_host_sync():
    """The scan is enqueued on the caller's stream: data written by an earlier kernel on that
    stream is seen without any host synchronisation in between."""
    import torch

    from argminmax_b200 import DeviceSlice, Workspace
    s = torch.cuda.Stream()
    ws = Workspace(0)
    n = 50_000_000
    with torch.cuda.stream(s):
        t = torch.zeros(n, dtype=torch.float32, device="cuda:0")
        for k in range(5):
            t.fill_(1.0)
            t[1234 + k] = -5.0
            t[n - 77 - k] = 9.0
            ds = DeviceSlice.from_torch(t, workspace=ws, stream=s.cuda_stream)
            assert ds.argminmax() == (1234 + k, n - 77 - k)
    ws.close()


def test_independent_workspaces_on_two_streams():
    import torch

    from argminmax_b200 import DeviceSlice, Workspace
    rng = np.random.default_rng(77)
    a = random_array("f32", 30_000_001, rng)
    b = random_array("i16", 60_000_003, rng)
    da, _ka = to_device(a)
    db, _kb = to_device(b)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    w1, w2 = Workspace(0), Workspace(0)
    da._ws, da.stream = w1, s1.cuda_stream
    db._ws, db.stream = w2, s2.cuda_stream
    ea, eb = oracle.argminmax(a, 0), oracle.argminmax(b, 0)
    for _ in range(10):
        da.launch(0)
        db.launch(0)
        da.launch(0)
        db.launch(0)
        assert db.wait(0) == eb
        assert da.wait(0) == ea
    w1.close()
    w2.close()


SPECIALS = {
    "f16": [np.float16(x) for x in (np.nan, np.inf, -np.inf, 0.0, -0.0, 65504, -65504, 6e-8, -6e-8, 1, -1)],
    "f32": [np.float32(x) for x in (np.nan, np.inf, -np.inf, 0.0, -0.0, 3.4028235e38, -3.4028235e38,
                                    1e-45, -1e-45, 1, -1)],
    "f64": [np.float64(x) for x in (np.nan, np.inf, -np.inf, 0.0, -0.0, 1.7976931348623157e308,
                                    -1.7976931348623157e308, 5e-324, -5e-324, 1, -1)],
}


# non-canonical NaNs: negative, signalling, all-ones payload (SURVEY A.3 #3/#4 policy: the CUDA
# path treats every NaN bit pattern like the reference's SCALAR code does)
SPECIALS["f16"] += list(np.array([0xFE00, 0x7C01, 0xFFFF, 0xFC01], dtype=np.uint16).view(np.float16))
SPECIALS["f32"] += list(np.array([0xFFC00000, 0x7F800001, 0xFFFFFFFF, 0xFF800001], dtype=np.uint32).view(np.float32))
SPECIALS["f64"] += list(np.array([0xFFF8000000000000, 0x7FF0000000000001, 0xFFFFFFFFFFFFFFFF,
                                  0xFFF0000000000001], dtype=np.uint64).view(np.float64))


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_fuzz_special_values_sizes_alignments(dt):
    """Seeded fuzz: random length, random base alignment, a random sprinkling of special values
    (NaN, +-inf, +-0, type min/max, subnormals; ints: type min/max/0) with many duplicates."""
    rng = np.random.default_rng(1000 + G.ALL_DTYPES.index(dt))
    T = G.NP[dt]
    es = np.dtype(T).itemsize
    for _case in range(60):
        n = int(rng.integers(1, 300_000)) if rng.random() < 0.7 else int(rng.integers(1, 200))
        if rng.random() < 0.5:
            a = random_array(dt, n, rng)
        else:  # small value range => lots of ties
            a = rng.integers(0, 4, n).astype(T)
        if dt in G.FLOAT_DTYPES:
            specials = SPECIALS[dt]
        else:
            info = np.iinfo(T)
            specials = [T(info.min), T(info.max), T(0)]
        k = int(rng.integers(0, max(2, n // 50)))
        if k:
            pos = rng.integers(0, n, k)
            sp = np.array(specials, dtype=T)
            # assign through the raw bits so that NaN payloads survive
            U = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[es]
            a.view(U)[pos] = sp.view(U)[rng.integers(0, len(sp), k)]
        off = int(rng.integers(0, 64 // es)) * es
        for mode in modes_for(dt):
            assert_parity(a, mode, byte_offset=off, what="fuzz")


def test_arrow_and_dlpack_adapters():
    """Row (f) of SURVEY 8: Arrow values buffer (validity bitmap ignored, like src/lib.rs:777)
    and DLPack / __cuda_array_interface__ producers, all zero-copy views."""
    import pyarrow as pa
    import torch

    from argminmax_b200 import DeviceSlice, HostSlice
    rng = np.random.default_rng(91)
    a = random_array("f64", 2_000_003, rng)
    arr = pa.array(a)
    assert HostSlice.from_arrow(arr).argminmax() == oracle.argminmax(a, 0)
    sl = arr.slice(12_345, 1_000_000)       # non-zero offset
    assert HostSlice.from_arrow(sl).argminmax() == oracle.argminmax(a[12_345:1_012_345], 0)
    # nulls are ignored exactly like the reference's adapter: the slot's stored value counts
    b = rng.integers(-1000, 1000, 100_000).astype(np.int32)
    b[777] = -5000
    masked = pa.array(b, mask=(np.arange(b.size) == 777))
    got = HostSlice.from_arrow(masked).argminmax()
    stored = np.frombuffer(masked.buffers()[1], dtype=np.int32)[: b.size]
    assert got == oracle.argminmax(stored, 0)
    # DLPack + __cuda_array_interface__
    t = torch.from_numpy(a).cuda()
    assert DeviceSlice.from_dlpack(t).argminmax() == oracle.argminmax(a, 0)

    class Cai:
        def __init__(self, t):
            self.t = t
            self.__cuda_array_interface__ = {"shape": (t.numel(),), "typestr": "<f8",
                                             "data": (t.data_ptr(), False), "version": 3}
    assert DeviceSlice.from_cuda_array_interface(Cai(t)).nanargminmax() == oracle.argminmax(a, 1)


def test_pipelined_back_to_back_scans_keep_every_result():
    """Pipelined mode (PDL: the streaming phase of scan k+1 overlaps the tail of scan k on the
    same workspace): every call must still deliver its own record, in order."""
    import torch

    from argminmax_b200 import AmmRecord, DTYPES, Workspace, lib
    L = lib()
    rng = np.random.default_rng(202)
    arrays = [random_array("f32", n, rng) for n in (5_000_011, 3_000_000, 12_345, 9_999_999, 64, 2_500_001)]
    arrays.append(random_array("u8", 20_000_003, rng).view(np.uint8))
    expected = [oracle.argminmax(a, 0) for a in arrays]
    devs = [to_device(a) for a in arrays]
    for pipelined in (True, False):
        ws = Workspace(0)
        ws.set_pipelined(pipelined)
        s = torch.cuda.Stream()
        reps = 6
        recs = torch.zeros((reps * len(arrays), 64), dtype=torch.uint8, device="cuda:0")
        with torch.cuda.stream(s):
            k = 0
            for _ in range(reps):
                for (ds, _keep), a in zip(devs, arrays):
                    rc = L.amm_argminmax_launch(DTYPES[ds.dtype], ds.data_ptr, ds.n, 0, 0, ws.handle,
                                                s.cuda_stream, recs[k].data_ptr())
                    assert rc == 0
                    k += 1
        s.synchronize()
        raw = recs.cpu().numpy().tobytes()
        for k in range(reps * len(arrays)):
            r = AmmRecord.from_buffer_copy(raw[64 * k:64 * k + 64])
            assert (r.min_idx, r.max_idx) == expected[k % len(arrays)], (pipelined, k)
            assert r.seq == k + 1
        out = (ctypes.c_uint64 * 2)()
        assert L.amm_workspace_wait(ws.handle, s.cuda_stream, 0, out) == 0
        assert (out[0], out[1]) == expected[-1]
        ws.close()


def test_current_device_is_left_untouched():
    """Every entry point restores the caller's current CUDA device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from argminmax_b200 import DeviceSlice, Workspace
    torch.cuda.set_device(0)
    t = torch.arange(100_000, dtype=torch.float32, device="cuda:1")
    ws = Workspace(1)
    ds = DeviceSlice.from_torch(t, workspace=ws)
    assert ds.argminmax() == (0, 99_999)
    x = torch.zeros(4, device="cuda")      # allocates on the *current* device
    assert x.device.index == 0
    import ctypes
    lib = ctypes.CDLL("libcudart.so.12")
    cur = ctypes.c_int(-1)
    lib.cudaGetDevice(ctypes.byref(cur))
    assert cur.value == 0
    ws.close()


def test_default_stream_with_another_current_device():
    """stream = NULL means 'default stream of the current device': a call on a workspace of
    device 0 must not be confused by the caller having made device 1 current (long kernels
    exercise the stream-query fallback of the result poll)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from argminmax_b200 import DeviceSlice, Workspace
    rng = np.random.default_rng(13)
    a = random_array("f32", 60_000_000, rng)
    ds, _keep = to_device(a)            # lives on cuda:0
    ds._ws = Workspace(0)
    exp = oracle.argminmax(a, 0)
    torch.cuda.set_device(1)
    try:
        for _ in range(20):
            assert ds.argminmax() == exp
    finally:
        torch.cuda.set_device(0)
    ds._ws.close()


@pytest.mark.parametrize("dt", ["i64", "f64", "f32"])
def test_fresh_workspace_on_recycled_memory(dt):
    """A workspace's per-CTA slots (grid step of the 64-bit-key kernels) are validated by an
    integrity word seeded with the call's sequence number -- and every workspace starts counting
    at 1.  A new workspace that gets a destroyed one's memory back must not accept the old slots:
    same sizes, same sequence numbers, different data, workspaces created and destroyed in turn."""
    from argminmax_b200 import Workspace
    rng = np.random.default_rng(123)
    n = 3_000_000
    arrays = [random_array(dt, n, rng) for _ in range(4)]
    slices = [to_device(a) for a in arrays]
    expect = [oracle.argminmax(a, 0) for a in arrays]
    for rep in range(6):
        ws = Workspace(0)
        for k in range(3):
            i = (rep + k) % len(arrays)
            ds = slices[i][0]
            ds._ws = ws
            assert ds.argminmax() == expect[i], (rep, k)
        ws.close()
