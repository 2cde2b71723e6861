"""Host logic of the one-trip latency kernel's launch shape (scan_launch.cuh: plan_direct),
without a GPU: which inputs take one CTA, how many vectors a thread holds, and that the
(grid, threads, vectors per thread) the kernel is launched with covers every whole vector of the
array exactly the way the kernel indexes it (vector v0 + u * threads of thread t in CTA b,
v0 = b * threads * vpt + t, u < vpt)."""
import ctypes

import pytest

import argminmax_b200 as amm

KEYS = ["head_elems", "n_vec", "tail_start", "grid", "threads", "vpt", "fits"]
SM, CTAS, THREADS = 148, 2, 256
CAP_VECS = SM * CTAS * THREADS  # vectors per "vectors per thread" step


def plan(n, es, addr=0, sm=SM, ctas=CTAS, threads=THREADS, min_vpt=1, one_cta_vecs=2048):
    out = (ctypes.c_uint64 * 8)()
    rc = amm.lib().amm_debug_plan_direct(addr, n, es, sm, ctas, threads, min_vpt, one_cta_vecs, out)
    assert rc == 0, rc
    return dict(zip(KEYS, [int(x) for x in out]))


@pytest.mark.parametrize("es", [1, 2, 4, 8])
@pytest.mark.parametrize("misalign", [0, 1, 3])
def test_shape_covers_every_vector_once(es, misalign):
    ve = 16 // es
    n_vecs = [0, 1, 255, 256, 257, 1023, 1024, 1025, 2047, 2048, 2049, 4096, CAP_VECS - 1, CAP_VECS, CAP_VECS + 1,
              2 * CAP_VECS, 2 * CAP_VECS + 1, 3 * CAP_VECS + 5, 4 * CAP_VECS - 1, 4 * CAP_VECS]
    for nv in n_vecs:
        for tail in (0, 1, ve - 1):
            head = ((16 - (misalign * es) % 16) % 16) // es
            n = head + nv * ve + tail
            if n == 0:
                continue
            pl = plan(n, es, addr=0x7F0000000000 + misalign * es)
            # head | vectors | tail account for every element
            assert pl["head_elems"] == min(n, head)
            assert pl["tail_start"] == pl["head_elems"] + pl["n_vec"] * ve
            assert 0 <= n - pl["tail_start"] < ve
            if not pl["fits"]:
                assert n // ve > 4 * CAP_VECS
                continue
            t, vpt, grid = pl["threads"], pl["vpt"], pl["grid"]
            assert t % 32 == 0 and 64 <= t <= 512 and 1 <= vpt <= 4 and grid >= 1
            # the kernel's indexing: CTA b covers vectors [b * t * vpt, (b + 1) * t * vpt)
            assert grid * t * vpt >= pl["n_vec"]
            assert (grid - 1) * t * vpt < max(pl["n_vec"], 1)          # no CTA without work
            assert grid <= SM * CTAS or vpt == 4 or pl["n_vec"] <= t * 4  # more vectors per thread before more CTAs


def test_one_cta_rules():
    # 32-bit elements: one CTA of 256 threads up to 1024 vectors, of 512 up to 2048 vectors / 16 Ki elements
    assert plan(4096, 4)["grid"] == 1 and plan(4096, 4)["threads"] == 256 and plan(4096, 4)["vpt"] == 4
    assert plan(8192, 4)["grid"] == 1 and plan(8192, 4)["threads"] == 512 and plan(8192, 4)["vpt"] == 4
    assert plan(8192 + 4, 4)["grid"] > 1 and plan(8192 + 4, 4)["threads"] == 256
    assert plan(1000, 4)["vpt"] == 1 and plan(1100, 4)["vpt"] == 2      # ceil(n_vec / 256)
    # bytes: 2048 vectors of u8 are 32 Ki elements -> above the element limit, several CTAs
    assert plan(16384, 1)["grid"] == 1 and plan(32768, 1)["grid"] > 1
    # 8-byte elements: 2048 vectors = 4 Ki elements -> one CTA of 512 threads
    assert plan(4096, 8)["grid"] == 1 and plan(4096, 8)["threads"] == 512
    assert plan(4098, 8)["grid"] > 1
    # the knob switches the 512-thread form off
    assert plan(8192, 4, one_cta_vecs=0)["grid"] > 1


def test_vectors_per_thread_steps_and_limit():
    for k in (1, 2, 3, 4):
        pl = plan(k * CAP_VECS * 4, 4)          # f32: 4 elements per vector
        assert pl["fits"] == 1 and pl["vpt"] == k and pl["grid"] == SM * CTAS
    assert plan(4 * CAP_VECS * 4 + 4, 4)["fits"] == 0
    assert plan(CAP_VECS * 4 + 4, 4)["vpt"] == 2
    # minimum vectors per thread (AMM_DIRECT_VPT)
    assert plan(100_000, 4, min_vpt=2)["vpt"] == 2
