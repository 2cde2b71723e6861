"""Memory-safety check without compute-sanitizer (closed on this pool): the library is also
built with -DAMM_BOUNDS_CHECK, where every global load of input data is verified to lie inside
the caller's array; tools/sanitize_case.py then drives every dtype x mode x kernel x load shape
at ragged sizes / unaligned bases through that build.  A violation fails the call."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_out_of_bounds_loads():
    from argminmax_b200 import _build
    dbg = _build.build_boundscheck()  # no-op when the in-tree debug twin is current
    env = dict(os.environ, AMM_LIB_PATH=dbg, AMM_HOST_CHUNK_MB="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")],
                          capture_output=True, text=True, timeout=900, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "sanitize_case ok" in proc.stdout


_SELFTEST = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import argminmax_b200 as amm
ok = 0
for n, thr in ((100_003, 0), (3_000_001, 0), (777, 1 << 62)):
    t = torch.arange(n, dtype=torch.float32, device="cuda:0")
    ws = amm.Workspace(0)
    ws.set_direct_threshold(thr)
    ds = amm.DeviceSlice.from_torch(t, workspace=ws)
    try:
        ds.argminmax()
    except amm.AmmError:
        ok += 1
print("detected", ok)
"""


def test_bounds_check_build_detects_a_violation(tmp_path):
    """The checker itself works: with AMM_DEBUG_SHRINK_BYTES the debug build is told that the
    array ends 4 bytes early while the kernels still scan all of it -- every call must fail."""
    from argminmax_b200 import _build
    dbg = _build.build_boundscheck()
    script = tmp_path / "selftest.py"
    script.write_text(_SELFTEST.format(root=ROOT))
    env = dict(os.environ, AMM_LIB_PATH=dbg, AMM_DEBUG_SHRINK_BYTES="4")
    proc = subprocess.run([sys.executable, str(script)], capture_output=True, text=True,
                          timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "detected 3" in proc.stdout, proc.stdout
