"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the
reference's own known-answer tests.  Bit-exact: indices are integers and no
arithmetic is performed on values, so there is no tolerance anywhere."""
import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.gpu_util import assert_parity, modes_for, random_array, run_gpu, to_device

pytestmark = pytest.mark.gpu

VECTORS = G.load()


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded():
    import argminmax_b200
    argminmax_b200.lib()  # raises if the CUDA extension is missing: no fallback exists
    import torch
    assert torch.cuda.is_available()


@pytest.mark.parametrize("v,arr", VECTORS, ids=[G.vector_id(v) for v, _ in VECTORS])
def test_reference_known_answers(v, arr):
    ds, _keep = to_device(arr)
    got = run_gpu(ds, v["mode"])
    assert list(got) == v["expected"], v["cite"]
    G.check_constraint(v, got)
    # single-output entry points agree with the fused one (test_utils.rs:108-114)
    if v["mode"] == 0:
        assert (ds.argmin(), ds.argmax()) == got
    else:
        assert (ds.nanargmin(), ds.nanargmax()) == got


SIZES = [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 129, 255, 257, 1000, 1027,
         4095, 4096, 4097, 8193, 32768, 65535, 65537, 131072, 1 << 20, (1 << 20) + 3, 3_000_017,
         10_000_019]


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_random_all_sizes(dt):
    """P1 (src/simd/test_utils.rs:44-84): kernel == scalar on random data, many lengths
    (ragged heads/tails, below one vector, below one tile, several tiles)."""
    rng = np.random.default_rng(sum(map(ord, dt)) + 17)
    for n in SIZES:
        a = random_array(dt, n, rng)
        for mode in modes_for(dt):
            assert_parity(a, mode, what="random")


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_many_small_random(dt):
    """P1: 10 000 x 129 in the reference; 300 x {129, 130..} here is enough to hit every
    alignment / tail combination, the rest of the budget goes to large sizes."""
    rng = np.random.default_rng(99)
    for i in range(300):
        a = random_array(dt, 129 + (i % 40), rng)
        for mode in modes_for(dt):
            assert_parity(a, mode, what="small-random")


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_unaligned_base_pointer(dt):
    """Slices may start anywhere (any multiple of the element size): the head before the first
    32-byte boundary and the sub-vector tail go through the exact scan."""
    rng = np.random.default_rng(5)
    es = np.dtype(G.NP[dt]).itemsize
    for k in range(0, 64 // es + 1):
        off = k * es
        for n in (1, 5, 37, 1027, 70001):
            a = random_array(dt, n, rng)
            for mode in modes_for(dt):
                assert_parity(a, mode, byte_offset=off, what="unaligned")


@pytest.mark.parametrize("dt", ["i8", "u8", "i16", "u16", "f16"])
def test_narrow_types_many_duplicates(dt):
    """Hard part #2 of SURVEY 7.3: 8/16-bit random data holds ~n/256 (n/65536) copies of each
    extreme; every reduction level must break ties by first occurrence.  Also the reference's
    index-overflow sizes 2^(bits+1) (P2, test_utils.rs:158-200)."""
    rng = np.random.default_rng(11)
    for n in (512, 131072, 1 << 22, (1 << 22) + 12345):
        a = random_array(dt, n, rng)
        for mode in modes_for(dt):
            assert_parity(a, mode, what="dups")


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_few_distinct_values_every_partition(dt):
    """Arrays of 2-3 distinct values: EVERY tile, every CTA and both the pre-assigned and the
    claimed (dynamic) part of the tile sequence hold copies of both extremes, so any level that
    breaks a tie the wrong way (later tile, later CTA, claimed tile before static tile, partial
    tile before full tile) shows up as a wrong index.  First occurrences sit at odd places."""
    rng = np.random.default_rng(4242)
    T = G.NP[dt]
    for n in (700_001, 3_000_017):
        a = np.full(n, 5, dtype=T)
        a[rng.integers(0, n, n // 3)] = 7
        a[rng.integers(0, n, n // 3)] = 3
        first = 1 + int(rng.integers(0, 40))
        a[:first] = 5                       # neither extreme before `first`
        a[first] = 3
        a[first + 1] = 7
        for mode in modes_for(dt):
            got = assert_parity(a, mode, what="few-values")
            assert got == (first, first + 1)
    a = np.full(2_000_003, 9, dtype=T)      # all equal: (0, 0) on every path
    for mode in modes_for(dt):
        assert assert_parity(a, mode, what="all-equal") == (0, 0)


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_duplicate_extrema_first_occurrence_across_tiles(dt):
    """Planted duplicate extremes in different tiles / CTAs: the earliest must win."""
    rng = np.random.default_rng(3)
    n = 5_000_000
    T = G.NP[dt]
    if dt in G.FLOAT_DTYPES:
        a = rng.uniform(-1000, 1000, n).astype(T)
        lo, hi = T(-60000.0), T(60000.0)
    else:
        info = np.iinfo(T)
        a = rng.integers(info.min // 2 + 1, info.max // 2, n, dtype=T)  # never hits lo / hi
        lo, hi = info.min, info.max
    pos_lo = [1_234_567, 1_234_568, 2_999_999, 4_999_999]
    pos_hi = [777_777, 2_000_000, 3_141_592, 4_999_998]
    a[pos_lo] = lo
    a[pos_hi] = hi
    for mode in modes_for(dt):
        got = assert_parity(a, mode, what="planted")
        assert got == (pos_lo[0], pos_hi[0])
    a[:] = a[0]  # all equal -> (0, 0)
    for mode in modes_for(dt):
        assert assert_parity(a, mode, what="all-equal") == (0, 0)


@pytest.mark.parametrize("dt", G.FLOAT_DTYPES)
def test_nan_placements(dt):
    """Config-3 style adversarial cases at a multi-tile size: NaN at 0 / middle / end / first
    100 / last 100 / random 0.1 %, +-inf, all-NaN, all +-inf; both semantics."""
    rng = np.random.default_rng(8)
    n = 2_000_003
    T = G.NP[dt]
    base = random_array(dt, n, rng)
    base[~np.isfinite(base)] = 0
    nan = T(np.nan)
    cases = {
        "nan0": [0], "nan_mid": [n // 2], "nan_last": [n - 1],
        "nan_first100": list(range(100)), "nan_last100": list(range(n - 100, n)),
        "nan_random": list(rng.choice(n, n // 1000, replace=False)),
        "nan_two_tiles": [700_001, 1_900_000],
    }
    for name, pos in cases.items():
        a = base.copy()
        a[pos] = nan
        for mode in (0, 1):
            got = assert_parity(a, mode, what=name)
            if mode == 1:
                assert got == (min(pos), min(pos))
    a = base.copy()
    a[1_500_000] = T(np.inf)
    a[400_000] = T(-np.inf)
    a[1_600_000] = T(np.inf)
    a[500_000] = T(-np.inf)
    for mode in (0, 1):
        assert assert_parity(a, mode, what="inf") == (400_000, 1_500_000)
    for fill, exp in ((nan, (0, 0)), (T(np.inf), (0, 0)), (T(-np.inf), (0, 0))):
        a = np.full(n, fill, dtype=T)
        for mode in (0, 1):
            assert assert_parity(a, mode, what="const") == exp
    # leading NaNs then values: scalar semantics = first non-NaN (SURVEY A.3 #1 policy)
    a = np.full(n, T(np.inf), dtype=T)
    a[:1000] = nan
    assert assert_parity(a, 0, what="nan-then-inf") == (1000, 1000)
    a = np.full(n, nan, dtype=T)
    a[n - 7] = T(3.0)
    assert assert_parity(a, 0, what="single-value-among-nan") == (n - 7, n - 7)


@pytest.mark.parametrize("dt", G.FLOAT_DTYPES)
def test_signed_zeros_and_subnormals(dt):
    """Policy zone: the CUDA path follows the reference's SCALAR code:
    f32/f64 compare -0.0 == +0.0 (first occurrence wins), f16 orders -0.0 < +0.0 through
    the ord key; subnormals are ordinary values (no flush to zero)."""
    T = G.NP[dt]
    n = 300_000
    for first, second in ((0.0, -0.0), (-0.0, 0.0)):
        a = np.ones(n, dtype=T)
        a[1000] = T(first)
        a[250_000] = T(second)
        for mode in (0, 1):
            assert_parity(a, mode, what="zeros-min")
        b = -np.ones(n, dtype=T)
        b[1000] = T(first)
        b[250_000] = T(second)
        for mode in (0, 1):
            assert_parity(b, mode, what="zeros-max")
    tiny = np.finfo(T).smallest_subnormal
    a = np.ones(n, dtype=T)
    a[123_456] = tiny
    a[234_567] = T(0.0)
    a[11] = -tiny
    for mode in (0, 1):
        assert assert_parity(a, mode, what="subnormal")[0] == 11
    if dt == "f16":
        rng = np.random.default_rng(21)
        a = random_array("f16", 1 << 21, rng, f16_bits=True)
        for mode in (0, 1):
            assert_parity(a, mode, what="f16-bit-patterns")


def test_worst_case_array_is_order_independent():
    """get_worst_case_array (dev_utils/src/utils.rs:10-27): alternating new-min / new-max."""
    n = 1_000_001
    i = np.arange(n)
    a = np.where(i % 2 == 0, -(i // 2), i // 2).astype(np.float32)
    assert_parity(a, 0, what="worst-case")
    assert_parity(a.astype(np.int32), 0, what="worst-case")
    assert_parity(np.arange(n, 0, -1).astype(np.float32), 0, what="decreasing")


def test_empty_input_is_an_error():
    """Reference: assert!(!arr.is_empty()) -> panic (src/scalar/generic.rs:182)."""
    import torch

    from argminmax_b200 import DeviceSlice, EmptyInputError
    t = torch.zeros(16, dtype=torch.float32, device="cuda:0")
    ds = DeviceSlice(t.data_ptr(), 0, "f32", owner=t)
    with pytest.raises(EmptyInputError):
        ds.argminmax()
    with pytest.raises(TypeError):
        DeviceSlice.from_torch(torch.zeros(4, dtype=torch.int32, device="cuda:0")).nanargminmax()


def test_all_tuning_variants_agree():
    """Every load shape / CTA size computes the same indices."""
    from argminmax_b200 import Workspace, lib
    rng = np.random.default_rng(4)
    for dt in ("u8", "i16", "f16", "f32", "f64", "i64"):
        a = random_array(dt, 3_333_333, rng)
        exp = oracle.argminmax(a, 0)
        ds, _keep = to_device(a, byte_offset=np.dtype(G.NP[dt]).itemsize)
        ws = Workspace(0)
        ds._ws = ws
        for variant in range(lib().amm_variant_count()):
            for threads in (128, 256, 512):
                for ctas in (1, 4):
                    ws.set_tuning(threads, ctas, variant)
                    assert ds.argminmax() == exp, (dt, variant, threads, ctas)
        ws.close()


def test_host_slice_pipeline():
    """HostSlice (chunked H2D + scan, amm_host_argminmax) == oracle, incl. chunk borders."""
    from argminmax_b200 import HostSlice
    rng = np.random.default_rng(6)
    for dt, n in (("f32", 20_000_003), ("u8", 70_000_001), ("f64", 9_000_001), ("f16", 1000)):
        a = random_array(dt, n, rng)
        hs = HostSlice.from_numpy(a)
        assert hs.argminmax() == oracle.argminmax(a, 0)
        if dt in G.FLOAT_DTYPES:
            a[n // 3] = np.nan
            a[n - 5] = np.nan
            assert hs.nanargminmax() == (n // 3, n // 3)
            assert hs.argminmax() == oracle.argminmax(a, 0)


def test_split_phase_and_record():
    """amm_argminmax_launch / amm_workspace_wait and the 64-byte record (global_offset)."""
    from argminmax_b200 import Workspace
    rng = np.random.default_rng(12)
    a = random_array("f32", 1_000_000, rng)
    ds, _keep = to_device(a)
    ws = Workspace(0)
    ds._ws = ws
    ds.launch(0, global_offset=5_000_000_000)
    got = ds.wait(0)
    exp = oracle.argminmax(a, 0)
    assert got == (exp[0] + 5_000_000_000, exp[1] + 5_000_000_000)
    rec = ws.host_record()
    assert rec.n == a.size and rec.flags == 1 and rec.nan_idx == 2**64 - 1
    assert rec.min_idx == got[0] and rec.max_idx == got[1]
    ws.close()


def test_cpp_device_slice_mirror():
    """The C++ mirror of the trait surface (include/argminmax_b200.hpp) through a plain-g++
    host program: the reference's doc examples and known answers (tests/cpp)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "bin", "device_slice_demo")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp")], check=True)
    proc = subprocess.run([exe, "check"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert proc.stdout.startswith("OK ")


def test_host_slice_pageable_bounce_ring():
    """Pageable host memory (numpy / Vec<T>) is bounced through a ring of pinned buffers filled by a
    team of copy threads (host_slice.cu); pinned memory is read by the DMA engine directly.  Both
    must give the oracle's answer, with extremes at chunk borders, in the last partial chunk and
    in every bounce buffer of the ring, for several team sizes."""
    import os

    import torch

    from argminmax_b200 import HostSlice, Workspace
    rng = np.random.default_rng(61)
    chunk = (32 << 20) // 4                       # elements per 32 MiB chunk
    n = 7 * chunk + 12_345                        # 8 chunks: the 3-buffer ring wraps twice
    a = rng.uniform(-1.0, 1.0, n).astype(np.float32)
    for pos_lo, pos_hi in ((chunk - 1, chunk), (5 * chunk + 7, 3 * chunk - 1), (n - 1, n - 2), (0, 6 * chunk)):
        b = a.copy()
        b[[pos_lo, min(n - 1, pos_lo + 2 * chunk)]] = -2.0     # duplicates: the first must win
        b[[pos_hi, min(n - 1, pos_hi + chunk + 1)]] = 2.0
        exp = oracle.argminmax(b, 0)
        pinned = torch.empty(n, dtype=torch.float32, pin_memory=True)
        pinned.numpy()[:] = b
        for threads in ("1", "3", ""):
            if threads:
                os.environ["AMM_HOST_COPY_THREADS"] = threads
            try:
                ws = Workspace(0)
                assert HostSlice.from_numpy(b, workspace=ws).argminmax() == exp, ("pageable", threads)
                assert HostSlice.from_torch(pinned, workspace=ws).argminmax() == exp, ("pinned", threads)
                b[n // 2] = np.nan
                assert HostSlice.from_numpy(b, workspace=ws).nanargminmax() == (n // 2, n // 2)
                b[n // 2] = 0.5
                ws.close()
            finally:
                os.environ.pop("AMM_HOST_COPY_THREADS", None)


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_latency_kernel_regimes(dt):
    """The one-trip latency kernel switches shape with the number of 16-byte vectors: one CTA of 256
    or 512 threads (up to 2048 vectors and 16 Ki elements), then several CTAs with 1, 2, 3 or 4
    vectors per thread (fused / select-based / lazy locate).  Every boundary, with a ragged tail and
    an unaligned head, with duplicated extremes so that the first occurrence is what is checked."""
    rng = np.random.default_rng(4242 + sum(map(ord, dt)))
    ve = 16 // np.dtype(G.NP[dt]).itemsize
    step = 148 * 2 * 256  # vectors per "vectors per thread" step on a 148-SM part
    n_vecs = [255, 256, 257, 1023, 1024, 1025, 2047, 2048, 2049, 4096, step, step + 1, 2 * step, 2 * step + 1,
              3 * step, 3 * step + 1, 4 * step - 7, 4 * step]
    sizes = sorted({nv * ve + t for nv in n_vecs for t in (0, ve - 1)} | {16383, 16384, 16385})
    for i, n in enumerate(sizes):
        a = random_array(dt, n, rng)
        if n > 64:
            # duplicate the extremes at a few places behind their first occurrence
            lo, hi = oracle.argminmax(a, 0)
            for j in rng.integers(0, n, 6):
                if j > lo and j != hi:
                    a[j] = a[lo]
            for j in rng.integers(0, n, 6):
                if j > hi and j != lo and a[j] != a[lo]:
                    a[j] = a[hi]
        for mode in modes_for(dt):
            b = a
            if mode == 1 and i % 3 == 0 and n > 8:
                b = a.copy()
                b[[n // 2, n - 2]] = np.nan
            assert_parity(b, mode, byte_offset=(i % 4) * np.dtype(G.NP[dt]).itemsize, what="latency regimes",
                          paths=("default", "direct"))
