"""Host logic of the chunked window path (few large adjacent windows are cut into equal chunks,
libamm.cu: plan_window_chunks) -- pure CPU, through the library's test hook."""
import ctypes

import numpy as np
import pytest

import argminmax_b200 as amm


def plan(begins, ends, es=4, sm_count=148, per_sm=16):
    L = amm.lib()
    b = np.asarray(begins, dtype=np.uint64)
    e = np.asarray(ends, dtype=np.uint64)
    cap = 1 << 16
    cuts = np.zeros(cap, dtype=np.uint64)
    first = np.zeros(len(b) + 1, dtype=np.uint64)
    n_cuts = ctypes.c_uint64()
    p64 = ctypes.POINTER(ctypes.c_uint64)
    rc = L.amm_debug_plan_window_chunks(b.ctypes.data_as(p64), e.ctypes.data_as(p64), len(b), es, sm_count,
                                        per_sm, cuts.ctypes.data_as(p64), cap, ctypes.byref(n_cuts),
                                        first.ctypes.data_as(p64))
    return rc, cuts[: n_cuts.value].astype(np.int64), first.astype(np.int64)


def check(begins, ends, es=4, sm_count=148, per_sm=16):
    rc, cuts, first = plan(begins, ends, es, sm_count, per_sm)
    assert rc == 0
    assert np.all(np.diff(cuts) > 0)                      # chunks are non-empty and ascending
    assert cuts[0] == begins[0] and cuts[-1] == max(ends[-1], begins[0])
    for w, (b, e) in enumerate(zip(begins, ends)):
        lo, hi = first[w], first[w + 1]
        if b == e:
            assert lo == hi                               # an empty window owns no chunk
            continue
        assert cuts[lo] == b and cuts[hi] == e            # its chunks tile the window exactly
        sizes = np.diff(cuts[lo:hi + 1])
        assert sizes.min() > 0 and sizes[:-1].max(initial=0) - sizes.min() <= 256 * len(sizes)
        assert np.all(sizes[:-1] % 256 == 0)              # piece starts stay vector-aligned
        assert np.all(sizes[:-1] == sizes[0])             # equal pieces, a shorter last one at most
    n_chunks = len(cuts) - 1
    total = ends[-1] - begins[0]
    assert n_chunks <= sm_count * per_sm + 2 * len(begins)
    if total * es >= (sm_count * per_sm) * (128 << 10):   # enough data: close to the chunk target
        assert n_chunks >= sm_count * per_sm // 2
    return cuts, first


def test_fixed_windows_are_cut_into_equal_aligned_pieces():
    n, w = 268_435_456, 5_000_000                         # 54 windows of 20 MB (f32)
    begins = list(range(0, n, w))
    ends = [min(n, b + w) for b in begins]
    cuts, first = check(begins, ends)
    assert 1500 < len(cuts) - 1 < 2600


@pytest.mark.parametrize("es", [1, 2, 4, 8])
def test_ragged_windows_with_empty_ones(es):
    rng = np.random.default_rng(es)
    n = 400_000_000 // es
    inner = np.sort(rng.integers(0, n, 40))
    edges = np.concatenate([[0, 0], inner, [inner[-1]], [n, n]])   # empty windows included
    check(edges[:-1].tolist(), edges[1:].tolist(), es=es)


def test_small_total_keeps_chunks_near_128_kib_or_more():
    begins, ends = [0, 3_000_000], [3_000_000, 4_000_000]          # 16 MB of f32 in two windows
    cuts, first = check(begins, ends)
    for w in range(2):                                    # all but the last piece of each window
        sizes = np.diff(cuts[first[w]:first[w + 1] + 1])[:-1]
        # equalising the pieces of a window may shave a little off the 128 KiB floor, never much
        assert sizes.min() * 4 >= 0.9 * (128 << 10)


def test_not_adjacent_windows_are_refused():
    assert plan([0, 10], [5, 20])[0] == 1000               # a gap
    assert plan([0, 3], [5, 20])[0] == 1000                # an overlap
    assert plan([], [])[0] != 0
