"""Windowed / segmented argmin+argmax (one launch for many sub-slices): per window the answer
must equal the oracle's on that sub-slice (+ offset) -- the semantics of calling the
reference's `argminmax` once per bin."""
import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.gpu_util import modes_for, random_array, to_device

pytestmark = pytest.mark.gpu


def _expected(a, bounds, mode):
    out = []
    for b, e in bounds:
        if b == e:
            out.append((-1, -1))
        else:
            i, j = oracle.argminmax(a[b:e], mode)
            out.append((i + b, j + b))
    return out


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_fixed_windows(dt):
    rng = np.random.default_rng(600 + G.ALL_DTYPES.index(dt))
    n = 1_000_003
    a = random_array(dt, n, rng)
    if dt in G.FLOAT_DTYPES:
        a[rng.integers(0, n, 2000)] = np.nan
        a[5000:5600] = np.nan          # whole windows of NaN
    ds, _keep = to_device(a, byte_offset=np.dtype(G.NP[dt]).itemsize)
    for window in (1, 7, 100, 500, 4096, 100_000, n, 2 * n):
        if window == 1 and dt not in ("f32", "u8"):
            continue
        nwin = -(-n // window)
        bounds = [(w * window, min(n, (w + 1) * window)) for w in range(nwin)]
        for mode in modes_for(dt):
            got = [tuple(r) for r in ds.argminmax_windows(window, mode).cpu().tolist()]
            if window == 1:
                assert got == [(i, i) for i in range(n)]
            else:
                assert got == _expected(a, bounds, mode), (dt, window, mode)


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_fixed_windows_odd_sizes_on_an_aligned_array(dt):
    """Fixed windows whose byte size is NOT a multiple of 16 on a 16-byte aligned array take the
    32-bit-geometry staged kernel (argminmax_windows_fixed_kernel): every window starts at a
    different phase of the 16-byte grid, so heads, tails and windows shorter than one vector all
    occur; the last windows (rounded end outside the array, short last window) go to the general
    kernel.  Few distinct values => ties inside and across units; NaN runs for floats."""
    rng = np.random.default_rng(700 + G.ALL_DTYPES.index(dt))
    es = np.dtype(G.NP[dt]).itemsize
    n = 400_003
    a = random_array(dt, n, rng)
    dups = rng.integers(0, 4, n).astype(G.NP[dt])
    a[::3] = dups[::3]                      # many ties
    if dt in G.FLOAT_DTYPES:
        a[rng.integers(0, n, 3000)] = np.nan
        a[7000:7900] = np.nan               # whole windows of NaN
    ds, _keep = to_device(a, byte_offset=0)
    assert ds.data_ptr % 16 == 0
    sizes = [w for w in (17, 33, 129, 500, 1001, 2047) if (w * es) % 16 != 0 and w * es >= 16]
    sizes += [w for w in (3, 5, 9) if w * es >= 16 and (w * es) % 16 != 0]
    assert sizes
    for window in sizes:
        for mode in modes_for(dt):
            got = ds.argminmax_windows(window, mode).cpu().numpy()
            exp = oracle.argminmax_windows(a, window, mode)
            bad = np.nonzero((got != exp).any(axis=1))[0]
            assert bad.size == 0, (dt, window, mode, int(bad[0]), got[bad[0]].tolist(), exp[bad[0]].tolist())
    # the array ends exactly at a window end that is not 16-byte aligned, and at an aligned one
    for n2 in (129 * 1000, 129 * 1024):
        ds2, _k2 = to_device(a[:n2], byte_offset=0)
        for mode in modes_for(dt):
            got = ds2.argminmax_windows(129, mode).cpu().numpy()
            assert (got == oracle.argminmax_windows(a[:n2], 129, mode)).all(), (dt, n2, mode)


@pytest.mark.parametrize("dt", ["f32", "i16", "u8", "f64", "f16"])
def test_ragged_windows_with_empty_and_duplicates(dt):
    import torch
    rng = np.random.default_rng(77)
    n = 300_000
    a = rng.integers(0, 5, n).astype(G.NP[dt])     # many ties: first occurrence per window
    ds, _keep = to_device(a)
    cuts = np.sort(rng.integers(0, n + 1, 2000))
    cuts = np.concatenate([[0, 0], cuts, [n, n]])   # empty windows at both ends too
    offs = torch.from_numpy(cuts.astype(np.int64)).cuda()
    bounds = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
    for mode in modes_for(dt):
        got = [tuple(r) for r in ds.argminmax_windows(offs, mode).cpu().tolist()]
        assert got == _expected(a, bounds, mode)


def test_huge_windows_take_the_streaming_path():
    """Few windows of >= 32 MiB each: one pipelined streaming scan per window + a conversion
    kernel (windows_huge in libamm.cu); answers as for any other window size."""
    import torch
    rng = np.random.default_rng(5)
    n = 35_000_003
    a = random_array("f32", n, rng)
    a[rng.integers(0, n, 50)] = np.nan
    ds, _keep = to_device(a)
    window = 10_000_000
    bounds = [(w * window, min(n, (w + 1) * window)) for w in range(-(-n // window))]
    for mode in (0, 1):
        got = [tuple(r) for r in ds.argminmax_windows(window, mode).cpu().tolist()]
        assert got == _expected(a, bounds, mode)
    cuts = np.array([0, 9_000_001, 9_000_001, 20_000_000, n], dtype=np.int64)   # one empty window
    offs = torch.from_numpy(cuts).cuda()
    bounds = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
    got = [tuple(r) for r in ds.argminmax_windows(offs, 0).cpu().tolist()]
    assert got == _expected(a, bounds, 0)
    b = random_array("u8", 150_000_000, rng)
    db, _kb = to_device(b)
    w8 = 50_000_000
    got = [tuple(r) for r in db.argminmax_windows(w8, 0).cpu().tolist()]
    assert got == _expected(b, [(0, w8), (w8, 2 * w8), (2 * w8, 3 * w8)], 0)


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_short_windows_every_lane_group(dt):
    """Windows of 16 B ... 1 KiB take the TMA-staged kernel with 1, 2, 4 or 8 lanes per window
    (windows_kernel.cuh): every lane-group size, aligned and unaligned bases, window sizes that
    are / are not multiples of 16 bytes, NaN runs that cover whole windows, many ties."""
    es = np.dtype(G.NP[dt]).itemsize
    rng = np.random.default_rng(900 + G.ALL_DTYPES.index(dt))
    n = 200_003
    a = random_array(dt, n, rng)
    ties = rng.integers(0, 3, n).astype(G.NP[dt])     # first occurrence inside short windows
    if dt in G.FLOAT_DTYPES:
        a[rng.integers(0, n, 3000)] = np.nan
        a[70_000:70_900] = np.nan
        ties[rng.integers(0, n, 3000)] = np.nan
    for arr in (a, ties):
        for off in (0, es, 16 - es if es < 16 else 0):
            ds, _keep = to_device(arr, byte_offset=off)
            for wbytes in (16, 48, 100, 128, 136, 256, 264, 512, 520, 1000, 1024):
                window = max(2, wbytes // es)
                nwin = -(-n // window)
                bounds = [(w * window, min(n, (w + 1) * window)) for w in range(nwin)]
                for mode in modes_for(dt):
                    got = [tuple(r) for r in ds.argminmax_windows(window, mode).cpu().tolist()]
                    assert got == _expected(arr, bounds, mode), (dt, off, window, mode)


@pytest.mark.parametrize("dt", ["f32", "u8", "f64", "i16"])
def test_short_ragged_windows_outliers_and_unordered_offsets(dt):
    """Ragged short windows: (i) a few windows far larger than the staging buffer among many
    tiny ones, (ii) offsets that are NOT ascending (reversed pairs give empty windows, windows may
    overlap), (iii) an output buffer that is only 8-byte aligned.  The staged kernel must fall
    back to global loads for whatever its stage does not hold."""
    import torch
    rng = np.random.default_rng(31)
    n = 400_000
    a = random_array(dt, n, rng)
    ds, _keep = to_device(a, byte_offset=np.dtype(G.NP[dt]).itemsize)
    # (i) mean window ~ 40 elements, three outliers of 30k elements
    small = rng.integers(1, 80, 7000)
    sizes = np.concatenate([small[:2000], [30_000], small[2000:4000], [30_000, 0, 0, 30_000], small[4000:]])
    cuts = np.concatenate([[0], np.cumsum(sizes)])
    cuts = cuts[cuts <= n]
    cuts = np.concatenate([cuts, [n]])
    offs = torch.from_numpy(cuts.astype(np.int64)).cuda()
    bounds = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
    got = [tuple(r) for r in ds.argminmax_windows(offs, 0).cpu().tolist()]
    assert got == _expected(a, bounds, 0)
    # (ii) unordered offsets
    cuts = rng.integers(0, n + 1, 5001)
    cuts[::7] = np.sort(cuts[::7])
    offs = torch.from_numpy(cuts.astype(np.int64)).cuda()
    near = np.sort(rng.integers(0, n - 300, 3000))          # short, mostly ascending, some reversed
    cuts2 = np.stack([near, near + rng.integers(-50, 200, 3000)], axis=1).reshape(-1)
    cuts2 = np.clip(cuts2, 0, n)
    for c in (cuts, cuts2):
        offs = torch.from_numpy(c.astype(np.int64)).cuda()
        bounds = [(min(b, min(e, n)), min(e, n)) for b, e in zip(c[:-1].tolist(), c[1:].tolist())]
        got = [tuple(r) for r in ds.argminmax_windows(offs, 0).cpu().tolist()]
        assert got == _expected(a, bounds, 0)
    # (iii) output pointer that is 8- but not 16-byte aligned
    from argminmax_b200 import DTYPES, lib
    window = 24
    nwin = -(-n // window)
    out = torch.full((2 * nwin + 1,), -7, dtype=torch.int64, device="cuda:0")
    rc = lib().amm_argminmax_fixed_windows(DTYPES[dt], ds.data_ptr, ds.n, window, 0,
                                           ds.workspace.handle, ds.stream, out[1:].data_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    got = [tuple(r) for r in out[1:].view(-1, 2).cpu().tolist()]
    assert got == _expected(a, [(w * window, min(n, (w + 1) * window)) for w in range(nwin)], 0)
    assert int(out[0]) == -7


@pytest.mark.parametrize("dt", ["f32", "u8", "f64", "f16"])
def test_few_large_windows_are_chunked(dt):
    """Fewer than 4 windows per SM of 1-64 MiB each are cut into chunks that run through the
    ordinary window kernels, plus a combine kernel (windows_chunked in libamm.cu): duplicate
    extremes in different chunks of one window (first must win), NaNs, an all-NaN chunk, ragged
    windows with empty ones in between; chunked == not chunked == oracle."""
    import torch

    from argminmax_b200 import Workspace
    es = np.dtype(G.NP[dt]).itemsize
    rng = np.random.default_rng(1200 + es)
    n = (48 << 20) // es + 12_345
    a = random_array(dt, n, rng)
    T = G.NP[dt]
    if dt in G.FLOAT_DTYPES:
        a[~np.isfinite(a)] = 0
        big = T(6.0e4)
        a[np.abs(a) >= big] = 0
        lo, hi = -big, big
    else:
        info = np.iinfo(T)
        lo, hi = T(info.min), T(info.max)
        mid = T(info.min // 2 + info.max // 2 + 1)
        a[a == lo] = mid
        a[a == hi] = mid
    window = (5 << 20) // es + 77          # ~5 MiB windows, unaligned length
    nwin = -(-n // window)
    for w in range(nwin):                   # duplicate extremes far apart inside every window
        b, e = w * window, min(n, (w + 1) * window)
        p = np.sort(rng.choice(np.arange(b, e), 6, replace=False))
        a[p[[1, 3, 5]]] = lo
        a[p[[0, 2, 4]]] = hi
    if dt in G.FLOAT_DTYPES:
        a[window + 100_000: window + 100_000 + (200 << 10) // es] = np.nan   # a whole chunk of NaN
        a[3 * window + 5] = np.nan
    ds, _keep = to_device(a, byte_offset=es)
    fixed = [(w * window, min(n, (w + 1) * window)) for w in range(nwin)]
    cuts = np.array([0, 0, 2_000_000 // es * 4, 2_000_000 // es * 4, n // 2, n - 3, n, n], dtype=np.int64)
    ragged = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
    offs = torch.from_numpy(cuts).cuda()
    import os
    os.environ["AMM_WIN_CHUNKED"] = "0"     # read when a workspace is created
    try:
        plain = Workspace(0)                # the same windows through the un-chunked paths
    finally:
        del os.environ["AMM_WIN_CHUNKED"]
    for mode in modes_for(dt):
        exp_fixed, exp_ragged = _expected(a, fixed, mode), _expected(a, ragged, mode)
        for ws in (ds.workspace, plain):
            ds._ws = ws
            assert [tuple(r) for r in ds.argminmax_windows(window, mode).cpu().tolist()] == exp_fixed
            assert [tuple(r) for r in ds.argminmax_windows(offs, mode).cpu().tolist()] == exp_ragged
    plain.close()


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_fuzz_windows(dt):
    """Seeded fuzz over the windowed form: random length, base alignment, window shape (fixed
    of any size, ragged with empty windows and outliers), NaN / +-inf / type extremes sprinkled
    in, many ties.  Every window must equal the oracle on its sub-slice, whatever kernel the
    dispatcher picks (staged 1..32 lanes incl. the geometry-free form, 8 lanes from global, warp,
    CTA, chunked)."""
    import torch
    rng = np.random.default_rng(7000 + G.ALL_DTYPES.index(dt))
    T = G.NP[dt]
    es = np.dtype(T).itemsize
    for _case in range(24):
        n = int(rng.integers(1, 400_000)) if rng.random() < 0.8 else int(rng.integers(1, 300))
        a = random_array(dt, n, rng) if rng.random() < 0.5 else rng.integers(0, 4, n).astype(T)
        if dt in G.FLOAT_DTYPES:
            k = int(rng.integers(0, max(2, n // 40)))
            a[rng.integers(0, n, k)] = rng.choice(np.array([np.nan, np.inf, -np.inf, 0.0], dtype=T), k)
            if rng.random() < 0.3 and n > 100:
                s0 = int(rng.integers(0, n - 50))
                a[s0:s0 + int(rng.integers(1, 3000))] = np.nan       # NaN runs spanning windows
        off = int(rng.integers(0, 32 // es)) * es
        ds, _keep = to_device(a, byte_offset=off)
        if rng.random() < 0.6:
            window = int(rng.choice([1, 2, 3, 4, 8, 16, 33, 64, 100, 128, 250, 256, 1000, 1024, 5000,
                                     70_000, n, n + 5]))
            if window == 1 and n > 50_000:
                window = 2
            nwin = -(-n // window)
            bounds = [(w * window, min(n, (w + 1) * window)) for w in range(nwin)]
            arg = window
        else:
            mean = int(rng.choice([3, 20, 100, 700, 5000]))
            nwin = max(1, min(n // mean, 20_000))
            cuts = np.sort(rng.integers(0, n + 1, nwin + 1))
            cuts[0] = 0 if rng.random() < 0.5 else cuts[0]
            if rng.random() < 0.3 and nwin > 4:      # an outlier: one window swallows a big range
                cuts[nwin // 2:] = np.maximum(cuts[nwin // 2:], min(n, cuts[nwin // 2 - 1] + n // 3))
            bounds = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
            arg = torch.from_numpy(cuts.astype(np.int64)).cuda()
        for mode in modes_for(dt):
            got = [tuple(r) for r in ds.argminmax_windows(arg, mode).cpu().tolist()]
            assert got == _expected(a, bounds, mode), (dt, _case, n, off, mode,
                                                       arg if isinstance(arg, int) else "ragged")


def test_large_mean_but_all_windows_empty():
    """Few windows over a big array whose offsets all collapse to one point: every window is
    empty -> (-1, -1) pairs (the chunked path has nothing to cut and must hand over)."""
    import torch
    a = np.arange(8_000_000, dtype=np.float32)
    ds, _keep = to_device(a)
    for point in (0, 4_000_000, a.size):
        offs = torch.full((5,), point, dtype=torch.int64, device="cuda:0")
        got = [tuple(r) for r in ds.argminmax_windows(offs, 0).cpu().tolist()]
        assert got == [(-1, -1)] * 4
    offs = torch.tensor([0, 0, a.size, a.size], dtype=torch.int64, device="cuda:0")
    got = [tuple(r) for r in ds.argminmax_windows(offs, 0).cpu().tolist()]
    assert got == [(-1, -1), (0, a.size - 1), (-1, -1)]


@pytest.mark.parametrize("dt", ["f32", "u8", "f64", "f16", "i64"])
@pytest.mark.parametrize("mean", [40, 400, 1100])
def test_ragged_windows_with_giant_outliers(dt, mean):
    """The window kernel is chosen by the MEAN window size (1-8 lanes per window from a staged
    block, 8 lanes or one warp per window from global memory).  Windows far beyond that -- here
    0.3-1.5 M elements among tens of thousands of ~`mean`-element ones, at the front, in the
    middle and at the end -- are handed to the whole CTA (window_by_cta) instead of keeping one
    lane group busy for milliseconds; the answers do not change."""
    import time

    import torch
    rng = np.random.default_rng(1000 + mean)
    n = 6_000_000
    a = random_array(dt, n, rng)
    if dt in G.FLOAT_DTYPES:
        a[rng.integers(0, n, 40)] = np.nan
    ds, _keep = to_device(a)
    giants = [1_500_000, 300_000, 700_000, 400_000]
    n_small = (n - sum(giants)) // mean
    small = rng.integers(max(1, mean // 2), mean + mean // 2 + 1, n_small)
    k = n_small // 3
    sizes = np.concatenate([[giants[0]], small[:k], [giants[1], 0, giants[2]], small[k:2 * k], small[2 * k:], [giants[3]]])
    cuts = np.concatenate([[0], np.cumsum(sizes)])
    cuts = cuts[cuts < n]
    cuts = np.concatenate([cuts, [n]])
    offs = torch.from_numpy(cuts.astype(np.int64)).cuda()
    bounds = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
    for mode in modes_for(dt):
        out = ds.argminmax_windows(offs, mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ds.argminmax_windows(offs, mode)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        got = [tuple(r) for r in out.cpu().tolist()]
        assert got == _expected(a, bounds, mode)
        # a lane group walking 1.5 M 8-byte elements alone took 15-200 ms; the CTA takes 1-3 (wide margin:
        # this is a guard against the pathology coming back, not a benchmark)
        assert ms < 30.0, (dt, mean, mode, ms)
