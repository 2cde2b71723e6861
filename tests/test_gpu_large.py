"""BASELINE.json configs 2, 3 and 4/5-style sizes on the GPU.

config 2: all 11 dtypes x 100M elements, both semantics, bit-exact vs the oracle.
config 3: adversarial f16/f32/f64 at 100M (NaNs, +-inf, all-equal, duplicate extremes).
config 4/5: sizes where a CPU oracle pass is too slow are checked through size-independent
properties: planted extremes / NaNs at known positions (incl. indices >= 2^32), shard
decomposition == whole, and idempotence across load shapes."""
import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.gpu_util import assert_parity, modes_for, random_array, run_gpu, to_device

pytestmark = pytest.mark.gpu
N100M = 100_000_000


@pytest.mark.slow
@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_config2_all_dtypes_100m(dt):
    rng = np.random.default_rng(2000 + G.ALL_DTYPES.index(dt))
    a = random_array(dt, N100M, rng)
    ds, _keep = to_device(a)
    for mode in modes_for(dt):
        got = run_gpu(ds, mode)
        assert got == oracle.argminmax(a, mode), (dt, mode)
        # ... and equal to the reference's dispatched SIMD path (lane-exact emulation of what
        # an AVX-512 host computes, oracle/simd_emul.c): "scalar and SIMD" of north_star
        assert got == oracle.simd_emul_argminmax(a, mode, 0), (dt, mode, "simd")
    if dt == "f16":  # (ii) uniform bit patterns: subnormals, -0.0, every exponent
        a = random_array(dt, N100M, rng, f16_bits=True)
        ds, _keep = to_device(a)
        for mode in (0, 1):
            assert run_gpu(ds, mode) == oracle.argminmax(a, mode)


@pytest.mark.slow
@pytest.mark.parametrize("dt", G.FLOAT_DTYPES)
def test_config3_adversarial_100m(dt):
    T = G.NP[dt]
    rng = np.random.default_rng(3000 + G.FLOAT_DTYPES.index(dt))
    n = N100M
    base = random_array(dt, n, rng)
    base[~np.isfinite(base)] = 0
    nan = T(np.nan)  # canonical positive quiet NaN
    ds, keep = to_device(base)

    import torch

    def check(arr, expect_ret=None):
        keep[: arr.nbytes].copy_(torch.from_numpy(arr.view(np.uint8)))
        torch.cuda.synchronize()
        for mode in (0, 1):
            got = run_gpu(ds, mode)
            assert got == oracle.argminmax(arr, mode), (dt, mode)
            if mode == 1 and expect_ret is not None:
                assert got == (expect_ret, expect_ret)

    for pos in ([0], [n // 2], [n - 1], list(range(100)), list(range(n - 100, n)),
                sorted(rng.choice(n, n // 1000, replace=False).tolist())):
        a = base.copy()
        a[pos] = nan
        check(a, expect_ret=pos[0])
    a = base.copy()
    inf_pos = rng.choice(n, 64, replace=False)
    a[inf_pos[:32]] = T(np.inf)
    a[inf_pos[32:]] = T(-np.inf)
    check(a)
    a[n // 7] = nan
    check(a, expect_ret=n // 7)
    big = T(60000.0)
    a = base.copy()
    a[np.abs(a) >= big] = 0
    dups = np.sort(rng.choice(n, 6, replace=False))
    a[dups[[1, 3, 5]]] = -big   # duplicate extremes spread over different CTAs
    a[dups[[0, 2, 4]]] = big
    check(a)
    for fill in (T(1.5), T(np.inf), T(-np.inf), nan):
        check(np.full(n, fill, dtype=T))


@pytest.mark.slow
def test_f32_1g_planted_and_shard_decomposition():
    """config 4 size (2^30 f32 = 4 GiB): planted duplicate extremes / NaNs at known places;
    the answer over the whole array equals the combine of two half-array records."""
    import torch

    from argminmax_b200 import DeviceSlice, Workspace, combine_records
    n = 1 << 30
    g = torch.Generator(device="cuda:0").manual_seed(7)
    t = torch.empty(n, dtype=torch.float32, device="cuda:0")
    t.uniform_(-1.0e6, 1.0e6, generator=g)
    lo = [123_456_789, 900_000_001, n - 1]
    hi = [5, 777_777_777, n - 2]
    t[lo] = -2.0e6
    t[hi] = 2.0e6
    ws = Workspace(0)
    ds = DeviceSlice.from_torch(t, workspace=ws)
    assert ds.argminmax() == (lo[0], hi[0])
    assert ds.nanargminmax() == (lo[0], hi[0])
    half = n // 2
    recs = []
    for part, off in ((t[:half], 0), (t[half:], half)):
        d = DeviceSlice.from_torch(part, workspace=ws)
        d.launch(0, global_offset=off)
        d.wait(0)
        r = ws.host_record()
        recs.append(type(r).from_buffer_copy(bytes(r)))
    assert combine_records(recs, 0) == (lo[0], hi[0])
    t[[1_000_000_000, 600_000_000]] = float("nan")
    assert ds.nanargminmax() == (600_000_000, 600_000_000)
    assert ds.argminmax() == (lo[0], hi[0])
    # unaligned view of the same data (element offset 3): indices shift by 3
    v = DeviceSlice.from_torch(t[3:], workspace=ws)
    assert v.argminmax() == (lo[0] - 3, hi[0] - 3)
    ws.close()


@pytest.mark.slow
def test_f32_1g_full_oracle_pass():
    """config 4 as bench.py times it (f32 x 2^30 uniform random bit patterns, NaN -> 0): a FULL
    pass of the scalar oracle and of the restated dispatched SIMD path over the same 4 GiB, both
    semantics, then with NaNs and duplicate extremes injected.  ~10 s of CPU per oracle pass."""
    import torch

    from argminmax_b200 import DeviceSlice
    n = 1 << 30
    rng = np.random.default_rng(2026)
    a = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32)
    a[np.isnan(a)] = 0
    t = torch.empty(n, dtype=torch.float32, device="cuda:0")
    t.copy_(torch.from_numpy(a))
    ds = DeviceSlice.from_torch(t)
    exp = oracle.argminmax(a, 0)
    assert ds.argminmax() == exp
    assert oracle.cpu_simd_argminmax_f32(a) == exp          # the reference's SIMD path agrees
    assert ds.nanargminmax() == oracle.argminmax(a, 1) == exp  # no NaN: both semantics coincide
    # duplicate the extremes LATER in the array (first occurrence must still win), a NaN far out
    lo, hi = exp
    dup_lo, dup_hi, nan_at = n - 12_345, n - 3, 987_654_321
    for arr_idx, src in ((dup_lo, lo), (dup_hi, hi)):
        a[arr_idx] = a[src]
        t[arr_idx] = t[src]
    a[nan_at] = np.nan
    t[nan_at] = float("nan")
    torch.cuda.synchronize()
    assert ds.argminmax() == oracle.argminmax(a, 0) == exp
    assert ds.nanargminmax() == oracle.argminmax(a, 1) == (nan_at, nan_at)


@pytest.mark.slow
def test_indices_beyond_2_pow_32():
    """config 5 needs 64-bit global indices (16 Gi elements per GPU at 2 GPUs): a u8 array of
    2^32 + 2^20 + 5 elements with the extremes planted past index 2^32."""
    import torch

    from argminmax_b200 import DeviceSlice
    n = (1 << 32) + (1 << 20) + 5
    t = torch.full((n,), 100, dtype=torch.uint8, device="cuda:0")
    t[(1 << 32) + 77] = 3
    t[(1 << 32) + 1000] = 3
    t[(1 << 32) + 99_999] = 250
    t[n - 1] = 250
    ds = DeviceSlice.from_torch(t)
    assert ds.argminmax() == ((1 << 32) + 77, (1 << 32) + 99_999)
    t[5] = 250
    assert ds.argminmax() == ((1 << 32) + 77, 5)
