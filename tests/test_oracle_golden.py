"""Pin the CPU oracle (oracle/argminmax_oracle.c) against the reference's own
known-answer tests (tests/golden/, transcribed from src/simd/test_utils.rs,
tests/argminmax_test.rs, src/lib.rs docs, src/scalar/scalar_f16.rs).  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G

VECTORS = G.load()


@pytest.mark.parametrize("v,arr", VECTORS, ids=[G.vector_id(v) for v, _ in VECTORS])
def test_oracle_matches_reference_known_answers(v, arr):
    got = oracle.argminmax(arr, v["mode"])
    assert list(got) == v["expected"], v["cite"]
    G.check_constraint(v, got)


@pytest.mark.parametrize("v,arr", VECTORS, ids=[G.vector_id(v) for v, _ in VECTORS])
def test_oracle_single_ops_agree_with_fused(v, arr):
    """The reference asserts argmin()/argmax() == argminmax().0/.1 in the same tests
    (e.g. src/simd/test_utils.rs:108-114, tests/argminmax_test.rs:127-130)."""
    # the f16 ignore-NaN singles have no first-non-NaN rule: they
    # legitimately differ from the fused op when only +-inf follow leading NaNs.
    got = oracle.argminmax(arr, v["mode"])
    assert oracle.argmin(arr, v["mode"]) == got[0]
    assert oracle.argmax(arr, v["mode"]) == got[1]


def test_oracle_empty_is_an_error():
    # reference: assert!(!arr.is_empty()), src/scalar/generic.rs:182
    with pytest.raises(ValueError):
        oracle.argminmax(np.zeros(0, dtype=np.float32), 0)


def _random(dt, n, rng):
    T = G.NP[dt]
    if dt == "f16":
        return rng.uniform(-65504, 65504, n).astype(np.float16)
    if dt == "f32":
        a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
        a[np.isnan(a)] = 0
        return a
    if dt == "f64":
        a = rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64).copy()
        a[np.isnan(a)] = 0
        return a
    info = np.iinfo(T)
    return rng.integers(info.min, info.max, n, dtype=T, endpoint=True)


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_oracle_vs_numpy_on_random(dt):
    """Second opinion: numpy's argmin/argmax also return the first
    occurrence; on NaN-free data they must agree with the oracle (P1-style:
    src/simd/test_utils.rs:44-84 sizes 8193 and many x 129)."""
    rng = np.random.default_rng(1234)
    for n in [1, 2, 3, 129, 8193] + [129] * 200:
        a = _random(dt, n, rng)
        if dt in ("f32", "f64"):
            # numpy compares -0.0 == +0.0 like the reference's float compares
            exp = (int(np.argmin(a)), int(np.argmax(a)))
        elif dt == "f16":
            # f16 is ordered through the i16 ord key: -0.0 < +0.0 (scalar_f16.rs:10-13)
            k = a.view(np.int16).astype(np.int32)
            k = np.where(k < 0, k ^ 0x7FFF, k)
            exp = (int(np.argmin(k)), int(np.argmax(k)))
        else:
            exp = (int(np.argmin(a)), int(np.argmax(a)))
        for mode in (0, 1):
            assert oracle.argminmax(a, mode) == exp


def test_cpu_simd_restatement_matches_oracle():
    """The AVX-512 / AVX restatement of the dispatched f32 path (oracle/cpu_simd_f32.c)
    must give the scalar oracle's indices on ordinary data, incl. the 2^24 chunk
    restart (simd_f32_ignore_nan.rs:69) and a ragged tail."""
    rng = np.random.default_rng(7)
    for n in [1, 7, 15, 16, 17, 31, 33, 129, 8193, (1 << 24) + 16 + 5]:
        a = _random("f32", n, rng)
        exp = oracle.argminmax(a, 0)
        for isa in (-1, 0, 1, 2):
            try:
                got = oracle.cpu_simd_argminmax_f32(a, isa)
            except ValueError:
                continue  # ISA not available on this host
            assert got == exp, (n, isa)
        assert oracle.cpu_simd_argminmax_f32_mt(a, 3) == exp
    a = _random("f32", 5000, rng)
    a[::7] = np.nan
    assert oracle.cpu_simd_argminmax_f32(a) == oracle.argminmax(a, 0)
    assert oracle.cpu_simd_argminmax_f32_mt(a, 4) == oracle.argminmax(a, 0)
