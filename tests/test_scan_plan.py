"""Host logic of the streaming kernel's tile plan (scan_launch.cuh:plan_scan), without a GPU.

The test replays the index formulas of argminmax_scan_kernel (scan_kernel.cuh) on the plan the
library computes -- which vectors CTA b visits as static tiles, as claimed tiles and as its
partial tile -- and checks the invariants the kernel's first-occurrence logic rests on:

  * every whole vector of the array is visited exactly once, whatever the partition
    (grid-stride, contiguous ranges, grid-stride + claimed tail) and whoever claims what;
  * head + vectors + tail account for every element;
  * per thread, slot ids ascend with the vector index (so a strict compare keeps the first
    occurrence), claimed tiles lie behind every static tile, the partial tile behind everything;
  * slot ids fit 32 bits and static slot ids stay below the claimed range."""
import ctypes
import random

import numpy as np
import pytest

import argminmax_b200 as amm

KEYS = ["head_elems", "n_vec", "tail_start", "per_vecs", "grid", "n_full_tiles", "rem_vecs", "rem_owner",
        "dyn_tiles", "dyn_koff", "kind", "status"]


def plan(addr, n, es, vec, u, threads, grid_cap, sm_count=148, pipelined=0, contig=-1, dyn_pct=-1):
    out = (ctypes.c_uint64 * 12)()
    rc = amm.lib().amm_debug_plan_scan(addr, n, es, vec, u, threads, grid_cap, sm_count, pipelined, contig, dyn_pct, out)
    assert rc == 0, rc
    return dict(zip(KEYS, [int(x) for x in out]))


def replay(pl, u, threads, rng):
    """-> (visits per vector, list of (vector, order key) per thread sample) by the kernel's formulas."""
    n_vec, grid = pl["n_vec"], pl["grid"]
    tile_vecs = threads * u
    seen = np.zeros(n_vec, dtype=np.uint8)

    def mark(lo, hi):
        assert 0 <= lo <= hi <= n_vec, (lo, hi, n_vec)
        seen[lo:hi] += 1

    # claimed tiles: any assignment of tiles to CTAs -- the claim counter hands each id out once
    dyn_of = {b: [] for b in range(grid)}
    for d in range(pl["dyn_tiles"]):
        dyn_of[rng.randrange(grid)].append(d)
    order_ok = True
    for b in range(grid):
        slots = []  # (slot id, vector index of thread 0's vector) in visiting order
        if pl["per_vecs"]:
            vb = b * pl["per_vecs"]
            ve = min(n_vec, vb + pl["per_vecs"])
            length = max(0, ve - vb)
            n_full, rem = length // tile_vecs, length % tile_vecs
            base, kstride, rem_base = vb, tile_vecs, vb + n_full * tile_vecs
        else:
            nft = pl["n_full_tiles"]
            n_full = (nft - b + grid - 1) // grid if b < nft else 0
            rem = pl["rem_vecs"] if (pl["rem_vecs"] and b == pl["rem_owner"]) else 0
            base, kstride = b * tile_vecs, grid * tile_vecs
            rem_base = (nft + pl["dyn_tiles"]) * tile_vecs
        for k in range(n_full):  # static tiles, local tile id k
            lo = base + k * kstride
            mark(lo, lo + tile_vecs)
            slots.append((k * u, lo))
            assert k < pl["dyn_koff"]
        for d in sorted(dyn_of[b]):  # claimed tiles ascend per CTA
            lo = (pl["n_full_tiles"] + d) * tile_vecs
            mark(lo, lo + tile_vecs)
            slots.append(((pl["dyn_koff"] + d) * u, lo))
            assert (pl["dyn_koff"] + d + 1) * u < 2**32 - 1
        if rem:
            mark(rem_base, rem_base + rem)
            # the partial tile is merged last ("main wins ties"): it must lie behind all the rest
            assert all(v < rem_base for _, v in slots)
        # per thread the slot ids and the vector indices ascend together
        order_ok &= all(s0 < s1 and v0 < v1 for (s0, v0), (s1, v1) in zip(slots, slots[1:]))
    return seen, order_ok


CASES = [
    # (n, es, vec, u, threads, grid_cap, pipelined, contig, dyn_pct)
    (10_000_019, 4, 32, 2, 256, 592, 0, -1, -1),   # 40 MB f32: claimed tail (30 %)
    (10_000_019, 4, 32, 2, 256, 592, 1, -1, -1),   # pipelined, below 0.5 GiB: contiguous ranges
    (10_000_019, 4, 16, 4, 256, 296, 1, -1, -1),
    (3_000_017, 1, 16, 4, 256, 592, 0, -1, -1),    # fewer than two tiles per CTA: contiguous
    (300_000_007, 2, 32, 2, 256, 592, 0, -1, -1),  # 600 MB: claimed tail (8 %)
    (300_000_007, 2, 32, 2, 256, 592, 1, -1, -1),  # pipelined and HBM-sized: claimed tail too
    (5_000_003, 8, 32, 2, 256, 592, 0, 0, 0),      # static grid-stride forced
    (5_000_003, 8, 32, 2, 512, 148, 0, 1, -1),     # contiguous forced
    (700_001, 4, 16, 2, 64, 148, 0, 0, 50),        # the tests' tiny-tile "dynamic" path
    (700_001, 4, 16, 2, 64, 148, 0, 0, 90),
    (1000, 4, 32, 2, 256, 592, 0, -1, -1),         # less than one tile
    (17, 1, 16, 4, 256, 592, 0, -1, -1),           # less than one vector
]


@pytest.mark.parametrize("case", CASES, ids=[f"n{c[0]}_es{c[1]}_v{c[2]}x{c[3]}_t{c[4]}_g{c[5]}_p{c[6]}_c{c[7]}_d{c[8]}" for c in CASES])
@pytest.mark.parametrize("misalign", [0, 1, 3])
def test_every_vector_once_and_in_order(case, misalign):
    n, es, vec, u, threads, grid_cap, pipelined, contig, dyn_pct = case
    addr = 0x7F0000000000 + misalign * es   # element-aligned, vector-unaligned for misalign > 0
    pl = plan(addr, n, es, vec, u, threads, grid_cap, pipelined=pipelined, contig=contig, dyn_pct=dyn_pct)
    assert pl["status"] == 0
    vec_elems = vec // es
    # head | vectors | tail account for every element
    assert pl["head_elems"] == min(n, ((vec - addr % vec) % vec) // es)
    assert pl["tail_start"] == pl["head_elems"] + pl["n_vec"] * vec_elems
    assert 0 <= n - pl["tail_start"] < vec_elems
    assert 1 <= pl["grid"] <= max(1, grid_cap)
    tile_vecs = threads * u
    if not pl["per_vecs"]:
        assert (pl["n_full_tiles"] + pl["dyn_tiles"]) * tile_vecs + pl["rem_vecs"] == pl["n_vec"]
    if pl["kind"] == 2:
        assert pl["dyn_tiles"] >= pl["grid"]                        # one claim per CTA goes out on entry
        assert not pipelined or n * es >= (512 << 20)               # pipelined: HBM-sized inputs only
        assert pl["dyn_koff"] > (pl["n_full_tiles"] + pl["grid"] - 1) // pl["grid"]
    if pl["per_vecs"]:
        assert pl["per_vecs"] % threads == 0 and pl["kind"] == 1
    seen, order_ok = replay(pl, u, threads, random.Random(n + misalign))
    assert order_ok
    assert seen.size == pl["n_vec"]
    assert (seen == 1).all(), (int((seen == 0).sum()), int((seen > 1).sum()))


def test_partition_choice_follows_the_rules():
    # serialised, >= 2 tiles per CTA -> claimed tail; pipelined below 0.5 GiB -> no claims
    a = plan(0, 100_000_000, 4, 32, 2, 256, 592)
    b = plan(0, 100_000_000, 4, 32, 2, 256, 592, pipelined=1)
    c = plan(0, 1 << 30, 4, 32, 2, 256, 592, pipelined=1)
    assert a["kind"] == 2 and b["kind"] in (0, 1) and b["dyn_tiles"] == 0 and c["kind"] == 2
    # 30 % claimed below 0.5 GiB, 8 % above
    full_a = a["n_full_tiles"] + a["dyn_tiles"]
    assert abs(a["dyn_tiles"] / full_a - 0.30) < 0.01
    full_c = c["n_full_tiles"] + c["dyn_tiles"]
    assert abs(c["dyn_tiles"] / full_c - 0.08) < 0.01
    # a grid smaller than the SM count never uses the claim counters (ring-of-64 argument)
    d = plan(0, 100_000_000, 4, 32, 2, 256, 100)
    assert d["kind"] != 2
    # too large for 32-bit slot ids
    e = plan(0, 1 << 46, 1, 16, 2, 64, 148)
    assert e["status"] == 6   # AMM_ERR_TOO_LARGE
