"""The reference's randomized property test at the reference's own scale
(src/simd/test_utils.rs:13-18, 44-84: NB_RUNS = 10 000 random arrays of 129 elements plus one of
8193 per dtype and backend, scalar == SIMD; tests/argminmax_test.rs:240-262: 500 arrays of 5000).

On the GPU 10 000 arrays of 129 elements are ONE windowed launch over a 1 290 000-element array
(amm_argminmax_fixed_windows), checked window by window against the oracle (`argminmax` once per
window, oracle_argminmax_windows); the same number of arrays also goes through single calls of
the latency kernel for f32 and u8.  Every array is drawn like the reference draws it
(dev_utils/src/utils.rs:29-84), seeded."""
import ctypes

import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.gpu_util import modes_for, random_array, to_device

pytestmark = pytest.mark.gpu

NB_RUNS = 10_000      # src/simd/test_utils.rs:13
ARR_LEN = 129         # src/simd/test_utils.rs:16 (odd: every array has a ragged tail)
LONG_LEN = 8193       # src/simd/test_utils.rs:65-84 (above the i8/u8 overflow chunk)


def _with_float_specials(a, dt, rng, nan_windows, window):
    """Sprinkle what the reference's float tests sprinkle: NaNs (some whole windows of them) and
    +-inf (src/simd/test_utils.rs:224-285, 302-527)."""
    if dt not in G.FLOAT_DTYPES:
        return a
    n = a.size
    a[rng.integers(0, n, n // 500)] = np.nan
    a[rng.integers(0, n, n // 2000)] = np.inf
    a[rng.integers(0, n, n // 2000)] = -np.inf
    for w in nan_windows:
        a[w * window:(w + 1) * window] = np.nan
    return a


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_ten_thousand_arrays_of_129_in_one_launch(dt):
    rng = np.random.default_rng(20_000 + G.ALL_DTYPES.index(dt))
    a = random_array(dt, NB_RUNS * ARR_LEN, rng)
    a = _with_float_specials(a, dt, rng, nan_windows=(0, 17, NB_RUNS - 1), window=ARR_LEN)
    # element-size byte offset: windows start at every alignment (129 is odd)
    ds, _keep = to_device(a, byte_offset=np.dtype(G.NP[dt]).itemsize)
    for mode in modes_for(dt):
        got = ds.argminmax_windows(ARR_LEN, mode).cpu().numpy()
        exp = oracle.argminmax_windows(a, ARR_LEN, mode)
        bad = np.nonzero((got != exp).any(axis=1))[0]
        assert bad.size == 0, (dt, mode, int(bad[0]), got[bad[0]].tolist(), exp[bad[0]].tolist())


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_five_hundred_arrays_of_5000_in_one_launch(dt):
    """tests/argminmax_test.rs:240-262 (500 x 5000), plus a 8193-element tail window."""
    rng = np.random.default_rng(30_000 + G.ALL_DTYPES.index(dt))
    n = 500 * 5000
    a = random_array(dt, n + LONG_LEN, rng, f16_bits=True)
    a = _with_float_specials(a, dt, rng, nan_windows=(3,), window=5000)
    offs = np.concatenate([np.arange(0, n + 1, 5000), [n + LONG_LEN]]).astype(np.int64)
    import torch
    ds, _keep = to_device(a)
    d_offs = torch.from_numpy(offs).cuda()
    for mode in modes_for(dt):
        got = ds.argminmax_windows(d_offs, mode).cpu().numpy()
        exp = oracle.argminmax_windows(a, offs, mode)
        bad = np.nonzero((got != exp).any(axis=1))[0]
        assert bad.size == 0, (dt, mode, int(bad[0]), got[bad[0]].tolist(), exp[bad[0]].tolist())


@pytest.mark.parametrize("dt", ["f32", "u8"])
def test_ten_thousand_single_calls(dt):
    """The same 10 000 arrays through 10 000 synchronous amm_argminmax calls (the latency
    kernel), every call checked; plus 8193-element arrays through the same entry point."""
    import argminmax_b200 as amm
    rng = np.random.default_rng(40_000 + G.ALL_DTYPES.index(dt))
    a = random_array(dt, NB_RUNS * ARR_LEN + LONG_LEN, rng)
    a = _with_float_specials(a, dt, rng, nan_windows=(5,), window=ARR_LEN)
    ds, _keep = to_device(a)
    es = np.dtype(G.NP[dt]).itemsize
    lib = amm.lib()
    ws = amm.default_workspace(0)
    out = (ctypes.c_uint64 * 2)()
    for mode in modes_for(dt):
        exp = oracle.argminmax_windows(a[: NB_RUNS * ARR_LEN], ARR_LEN, mode)
        for w in range(NB_RUNS):
            rc = lib.amm_argminmax(amm.DTYPES[dt], ds.data_ptr + w * ARR_LEN * es, ARR_LEN, mode, ws.handle, None, out)
            assert rc == 0
            assert (out[0] + w * ARR_LEN, out[1] + w * ARR_LEN) == tuple(exp[w]), (dt, mode, w)
        for start in (0, 1, 7, NB_RUNS * ARR_LEN):   # 8193-element arrays at several alignments
            rc = lib.amm_argminmax(amm.DTYPES[dt], ds.data_ptr + start * es, LONG_LEN, mode, ws.handle, None, out)
            assert rc == 0
            assert (out[0], out[1]) == oracle.argminmax(a[start:start + LONG_LEN], mode), (dt, mode, start)
