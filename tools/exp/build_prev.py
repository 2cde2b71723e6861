#!/usr/bin/env python
"""A/B helper: build the library from the sources of a git revision (default HEAD) into
argminmax_b200/lib/libargminmax_b200_prev.so, so that one gpurun call can time the working tree
against it on the same box (AMM_LIB_PATH=... python tools/latency_c.py).
  python tools/exp/build_prev.py [rev]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from argminmax_b200 import _build  # noqa: E402


def main():
    rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
    dst = os.path.join(ROOT, "build", "prev_src")
    subprocess.run(["rm", "-rf", dst], check=True)
    os.makedirs(dst)
    tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "argminmax_b200/csrc", "include"], check=True, capture_output=True)
    subprocess.run(["tar", "-x", "-C", dst], input=tar.stdout, check=True)
    _build.CSRC = os.path.join(dst, "argminmax_b200", "csrc")
    out = os.path.join(_build.LIB_DIR, "libargminmax_b200_prev.so")
    print(_build._build_into(out, os.path.join(_build.LIB_DIR, "build_prev.stamp"), _build.OBJ_DIR + "_prev", [], False))


if __name__ == "__main__":
    main()
