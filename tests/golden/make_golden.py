#!/usr/bin/env python
"""Generate tests/golden/golden_vectors.{json,npz}.

The reference (a Rust crate) cannot be built or run in this image (no cargo, no
crates.io), so these fixtures are a TRANSCRIPTION of the reference's own
known-answer tests: each vector carries the input construction and the expected
(argmin, argmax) that the reference's test asserts, with the file:line of that
assertion (paths relative to the reference root; table in SURVEY.md App. B).

Where the reference test asserts "equal to the scalar implementation" on random
data plus an inequality (e.g. "index != 0"), the expected pair is computed here
with numpy on the non-NaN sub-array (np.argmin/np.argmax return the first
occurrence) -- independently of oracle/ -- and the reference's inequality is
stored as a constraint checked by the tests as well.

Random base arrays are seeded (numpy PCG64) and STORED in the .npz, so the
fixtures do not depend on numpy's generator staying stable.

Run:  python tests/golden/make_golden.py        (rewrites both files)
"""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

INT_DTYPES = {
    "i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64,
    "u8": np.uint8, "u16": np.uint16, "u32": np.uint32, "u64": np.uint64,
}
FLOAT_DTYPES = {"f16": np.float16, "f32": np.float32, "f64": np.float64}
ALL = {**INT_DTYPES, **FLOAT_DTYPES}
FLOAT_UINT = {"f16": np.uint16, "f32": np.uint32, "f64": np.uint64}

arrays: dict[str, np.ndarray] = {}
vectors: list[dict] = []


def add(vid, dtype, arr, mode, expected, cite, constraint=None, recipe=None):
    """recipe = {"arange_mod": [n, m]} stores `arange(n) % m` by construction instead
    of by value (keeps the 100 000 / 200 000-element monotonic fixtures tiny)."""
    key = f"{vid}_{dtype}"
    if key not in arrays and recipe is None:
        arr = np.ascontiguousarray(arr, dtype=ALL[dtype])
        # floats are stored as raw bits so NaN payloads survive any round trip
        arrays[key] = arr.view(FLOAT_UINT[dtype]) if dtype in FLOAT_DTYPES else arr
    vectors.append({
        "id": vid, "dtype": dtype, "array": None if recipe else key, "recipe": recipe,
        "mode": mode,
        "expected": [int(expected[0]), int(expected[1])],
        "cite": cite, "constraint": constraint,
    })


def type_min_max(dtype):
    if dtype in INT_DTYPES:
        info = np.iinfo(INT_DTYPES[dtype])
        return info.min, info.max
    info = np.finfo(FLOAT_DTYPES[dtype])
    return info.min, info.max   # finite min/max, num_traits::Bounded for floats


def random_finite(dtype, n, seed):
    """dev_utils/src/utils.rs:29-84 distributions, seeded: floats are uniform bit
    patterns with NaN -> 0.0 (f32/f64) / uniform by value (f16); +-inf removed so
    that the planted infinities are the only ones."""
    rng = np.random.default_rng(seed)
    if dtype == "f16":
        a = rng.uniform(-65504.0, 65504.0, n).astype(np.float16)
    elif dtype == "f32":
        a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
    elif dtype == "f64":
        a = rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64).copy()
    else:
        info = np.iinfo(INT_DTYPES[dtype])
        return rng.integers(info.min, info.max, n, dtype=INT_DTYPES[dtype], endpoint=True)
    a[~np.isfinite(a)] = 0
    return a


def first_minmax_ignoring_nan(a):
    ok = ~np.isnan(a)
    idx = np.flatnonzero(ok)
    sub = a[ok].astype(np.float64)
    return int(idx[np.argmin(sub)]), int(idx[np.argmax(sub)])


IGN, RET = 0, 1
TU = "src/simd/test_utils.rs"

# ---- G1 / G2: first index on identical values, test_utils.rs:106-150 --------------
for dt in ALL:
    ones = np.ones(64, dtype=ALL[dt])
    modes = [IGN] + ([RET] if dt in FLOAT_DTYPES else [])
    for m in modes:
        add("G1", dt, ones, m, (0, 0), f"{TU}:106-123")
    lo, hi = type_min_max(dt)
    d = ones.copy()
    d[[5, 13, 41]] = lo
    d[[7, 17, 31]] = hi
    for m in modes:
        add("G2", dt, d, m, (5, 7), f"{TU}:125-150")

N = 1027  # test_utils.rs:206 (FLOAT_ARR_LEN)
for i, dt in enumerate(FLOAT_DTYPES):
    T = FLOAT_DTYPES[dt]
    nan, inf = T(np.nan), T(np.inf)
    # ---- G3-G5: infinities, test_utils.rs:212-286 ---------------------------------
    for m in (IGN, RET):
        add("G3", dt, np.full(N, inf, dtype=T), m, (0, 0), f"{TU}:223-243")
        add("G4", dt, np.full(N, -inf, dtype=T), m, (0, 0), f"{TU}:245-264")
    r = random_finite(dt, N, 100 + i)
    d = r.copy(); d[100] = inf; d[200] = -inf
    for m in (IGN, RET):
        add("G5", dt, d, m, (200, 100), f"{TU}:266-285")

    # ---- G6-G14: ignore-NaN family, test_utils.rs:291-528 --------------------------
    d = random_finite(dt, N, 110 + i); d[0] = nan
    add("G6", dt, d, IGN, first_minmax_ignoring_nan(d), f"{TU}:302-327", "both != 0")
    d = np.ones(N, dtype=T); d[0] = nan
    add("G7", dt, d, IGN, (1, 1), f"{TU}:329-350")
    d = np.arange(N).astype(T); d[0] = nan
    add("G8", dt, d, IGN, (1, N - 1), f"{TU}:352-373")
    d = (N - np.arange(N)).astype(T); d[0] = nan
    add("G9", dt, d, IGN, (N - 1, 1), f"{TU}:375-396")
    d = (N - np.arange(N)).astype(T); d[:100] = nan
    add("G10", dt, d, IGN, (N - 1, 100), f"{TU}:398-424", "both > 99")
    d = random_finite(dt, N, 120 + i); d[N - 1] = nan
    add("G11", dt, d, IGN, first_minmax_ignoring_nan(d), f"{TU}:426-451", "both != 1026")
    d = d.copy(); d[N - 100:] = nan
    add("G12", dt, d, IGN, first_minmax_ignoring_nan(d), f"{TU}:453-479", "both < 927")
    d = random_finite(dt, N, 130 + i); d[123] = nan
    add("G13", dt, d, IGN, first_minmax_ignoring_nan(d), f"{TU}:481-506", "both != 123")
    add("G14", dt, np.full(N, nan, dtype=T), IGN, (0, 0), f"{TU}:508-527")

    # ---- G15-G22: return-NaN family, test_utils.rs:534-708 -------------------------
    d = random_finite(dt, N, 140 + i); d[0] = nan
    add("G15", dt, d, RET, (0, 0), f"{TU}:545-563")
    d = d.copy(); d[:100] = nan
    add("G16", dt, d, RET, (0, 0), f"{TU}:565-584")
    d = random_finite(dt, N, 150 + i); d[N - 1] = nan
    add("G17", dt, d, RET, (N - 1, N - 1), f"{TU}:586-604")
    d = d.copy(); d[N - 100:] = nan
    add("G18", dt, d, RET, (N - 100, N - 100), f"{TU}:606-625")
    d = random_finite(dt, N, 160 + i); d[123] = nan
    add("G19", dt, d, RET, (123, 123), f"{TU}:627-645")
    d = d.copy(); d[N - 100:] = nan
    add("G20", dt, d, RET, (123, 123), f"{TU}:647-666")
    add("G21", dt, np.full(N, nan, dtype=T), RET, (0, 0), f"{TU}:668-687")
    d = random_finite(dt, 128, 170 + i); d[17] = nan
    add("G22", dt, d, RET, (17, 17), f"{TU}:689-707")

# ---- G23 / G24: monotonic x % max_index, tests/argminmax_test.rs:100-165 ------------
ARRAY_LENGTH = 100_000  # tests/argminmax_test.rs:25
for dt in ALL:
    if dt == "f16":
        mx = 1 << 11   # f16::from_usize(1 << MANTISSA_DIGITS), tests/argminmax_test.rs:62
    elif dt in FLOAT_DTYPES:
        mx = ARRAY_LENGTH
    else:
        mx = min(ARRAY_LENGTH, int(np.iinfo(INT_DTYPES[dt]).max))
    rec = {"arange_mod": [ARRAY_LENGTH, mx]}
    add("G23", dt, None, IGN, (0, mx - 1), "tests/argminmax_test.rs:113-140", recipe=rec)
    if dt in FLOAT_DTYPES:
        add("G24", dt, None, RET, (0, mx - 1), "tests/argminmax_test.rs:142-165", recipe=rec)

# ---- G25-G27: documentation examples -------------------------------------------------
add("G25", "i32", np.arange(6), IGN, (0, 5), "src/lib.rs:50-53")
d = np.array([np.nan, 1, np.nan, 3, 4, 5], dtype=np.float32)
add("G26", "f32", d, IGN, (1, 5), "src/lib.rs:61-64")
add("G26", "f32", d, RET, (0, 0), "src/lib.rs:65-67")
add("G27", "i32", None, IGN, (0, 199_999), "README.md:53-55",
    recipe={"arange_mod": [200_000, 200_000]})

# ---- G28: f16 vs its exact f32 widening, src/scalar/scalar_f16.rs:246-347 ------------
rng = np.random.default_rng(28)
base = rng.uniform(-65504.0, 65504.0, 1025).astype(np.float16)
cases = {"plain": [], "nan0": [0], "nan512": [512], "nan1024": [1024], "allnan": list(range(1025))}
for name, pos in cases.items():
    d = base.copy()
    d[pos] = np.float16(np.nan)
    wide = d.astype(np.float32)
    for m in (IGN, RET):
        if np.isnan(wide).all():
            exp = (0, 0)
        elif m == RET and np.isnan(wide).any():
            k = int(np.flatnonzero(np.isnan(wide))[0]); exp = (k, k)
        else:
            exp = first_minmax_ignoring_nan(wide)
        add(f"G28{name}", "f16", d, m, exp, "src/scalar/scalar_f16.rs:246-347")
        add(f"G28{name}", "f32", wide, m, exp, "src/scalar/scalar_f16.rs:246-347")

np.savez_compressed(os.path.join(HERE, "golden_vectors.npz"), **arrays)
with open(os.path.join(HERE, "golden_vectors.json"), "w") as f:
    json.dump({"float_arrays_stored_as": "raw uint bits", "vectors": vectors}, f, indent=1)
print(f"{len(vectors)} vectors, {len(arrays)} arrays")
