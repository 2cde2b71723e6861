"""Loader for tests/golden/golden_vectors.{json,npz} (see tests/golden/make_golden.py)."""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NP = {
    "i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64,
    "u8": np.uint8, "u16": np.uint16, "u32": np.uint32, "u64": np.uint64,
    "f16": np.float16, "f32": np.float32, "f64": np.float64,
}
ALL_DTYPES = list(NP)
INT_DTYPES = [d for d in NP if d[0] != "f"]
FLOAT_DTYPES = ["f16", "f32", "f64"]

_cache = None


def load():
    """-> list of (vector_dict, numpy array in its real dtype)."""
    global _cache
    if _cache is None:
        with open(os.path.join(HERE, "golden", "golden_vectors.json")) as f:
            meta = json.load(f)
        store = np.load(os.path.join(HERE, "golden", "golden_vectors.npz"))
        out = []
        for v in meta["vectors"]:
            T = NP[v["dtype"]]
            if v["recipe"]:
                n, m = v["recipe"]["arange_mod"]
                arr = (np.arange(n) % m).astype(T)
            else:
                arr = store[v["array"]].view(T)
            out.append((v, arr))
        _cache = out
    return _cache


def vector_id(v):
    return f'{v["id"]}-{v["dtype"]}-{"ret" if v["mode"] else "ign"}'


def check_constraint(v, got):
    """The reference test's own inequality (e.g. 'index != 0'), when it states one."""
    c = v.get("constraint")
    if not c:
        return
    a, b = got
    if c == "both != 0":
        assert a != 0 and b != 0
    elif c == "both > 99":
        assert a > 99 and b > 99
    elif c == "both != 1026":
        assert a != 1026 and b != 1026
    elif c == "both < 927":
        assert a < 927 and b < 927
    elif c == "both != 123":
        assert a != 123 and b != 123
    else:
        raise AssertionError(f"unknown constraint {c}")
