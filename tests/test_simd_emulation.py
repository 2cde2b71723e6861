"""The lane-exact emulation of the reference's dispatched SIMD path (oracle/simd_emul.c):
(1) it satisfies the reference's own known-answer tests, which the reference runs against its
SIMD backends (src/simd/test_utils.rs); (2) it equals the scalar oracle wherever the reference
itself asserts scalar == SIMD (random data, P1/P2 of SURVEY App. B); (3) it reproduces the
DOCUMENTED places where the reference's SIMD path differs from its scalar path, which shows
the emulation is faithful rather than a copy of the scalar code.  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G
from tests.test_oracle_golden import _random

VECTORS = G.load()


@pytest.mark.parametrize("isa", [0, 1])
@pytest.mark.parametrize("v,arr", VECTORS, ids=[G.vector_id(v) for v, _ in VECTORS])
def test_emulation_matches_reference_known_answers(v, arr, isa):
    got = oracle.simd_emul_argminmax(arr, v["mode"], isa)
    assert list(got) == v["expected"], v["cite"]


@pytest.mark.parametrize("dt", G.ALL_DTYPES)
def test_emulation_equals_scalar_on_random_data(dt):
    """test_return_same_result (test_utils.rs:26-85) + the index-overflow sizes 2^(bits+1)
    (test_utils.rs:158-200; 1<<25 for f32 ignore-NaN, simd_f32_ignore_nan.rs:473)."""
    rng = np.random.default_rng(4321)
    sizes = [1, 15, 16, 17, 31, 33, 63, 65, 129, 513, 1027, 8193, 131072 + 5] + [129] * 100
    if dt == "f32":
        sizes.append((1 << 25) + 7)
    for n in sizes:
        a = _random(dt, n, rng)
        for mode in ((0, 1) if dt[0] == "f" else (0,)):
            exp = oracle.argminmax(a, mode)
            for isa in (0, 1):
                assert oracle.simd_emul_argminmax(a, mode, isa) == exp, (dt, n, mode, isa)


def test_f32_emulation_equals_the_intrinsics_restatement():
    """Two independent restatements of the same reference path must agree lane for lane,
    including inputs where scalar and SIMD differ."""
    rng = np.random.default_rng(99)
    for n in (16, 17, 100, 1027, 70_001):
        a = _random("f32", n, rng)
        a[rng.integers(0, n, max(1, n // 10))] = np.nan
        a[rng.integers(0, n, max(1, n // 20))] = np.inf
        a[: rng.integers(0, 5)] = np.nan
        for isa, force in ((0, 2), (1, 1)):
            try:
                ref = oracle.cpu_simd_argminmax_f32(a, force)
            except ValueError:
                continue
            assert oracle.simd_emul_argminmax(a, 0, isa) == ref, (n, isa)


def test_documented_scalar_simd_divergences_are_reproduced():
    """SURVEY A.3.  These are the reference's own documented / latent self-inconsistencies; the
    CUDA path follows the SCALAR side (see tests/test_gpu_parity.py)."""
    # 1: leading NaN, every other value +inf (README.md:124-125: 'index 0 is returned')
    a = np.full(64, np.inf, dtype=np.float32)
    a[0] = np.nan
    assert oracle.argminmax(a, 0) == (1, 1)
    assert oracle.simd_emul_argminmax(a, 0, 0) == (0, 1)
    # 2: f32 return-NaN with mixed-sign zeros: ord compare inside a chunk sees -0 < +0
    b = np.ones(64, dtype=np.float32)
    b[3] = 0.0
    b[9] = -0.0
    assert oracle.argminmax(b, 1)[0] == 3
    assert oracle.simd_emul_argminmax(b, 1, 0)[0] == 9
    # 4/5: f16 ignore-NaN applies its NaN test to the transformed register: -0.0 is dropped
    c = np.full(64, -1.0, dtype=np.float16)
    c[7] = -0.0
    assert oracle.argminmax(c, 0)[1] == 7
    assert oracle.simd_emul_argminmax(c, 0, 0)[1] == 0
