"""Build libargminmax_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libargminmax_b200.so")
STAMP = os.path.join(LIB_DIR, "build.stamp")
SOURCES = ["libamm.cu", "host_slice.cu", "sharded.cu", "scan_inst.cu"]
HEADERS = ["amm_internal.h", "record_slot.h", "policies.cuh", "scan_kernel.cuh", "ring_kernel.cuh", "scan_launch.cuh",
           "windows_kernel.cuh",
           os.path.join("..", "..", "include", "argminmax_b200.h")]
N_GROUPS = 7  # scan_inst.cu is compiled once per policy group (-DAMM_GROUP=k), in parallel
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(os.path.dirname(HERE), "build", "obj")  # build/ is git- and gpurun-ignored


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libargminmax_b200.so cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _fingerprint()


def _compile(nvcc: str, src: str, obj: str, extra: list[str], verbose: bool) -> str:
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src} {extra}:\n" + proc.stdout + proc.stderr)
    return proc.stderr


DEBUG_LIB_PATH = os.path.join(LIB_DIR, "libargminmax_b200_boundscheck.so")
DEBUG_STAMP = os.path.join(LIB_DIR, "build_boundscheck.stamp")


def _build_into(lib_path: str, stamp: str, obj_dir: str, defines: list[str], verbose: bool) -> str:
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    jobs = [(src, os.path.join(obj_dir, src.replace(".cu", ".o")), list(defines))
            for src in SOURCES if src != "scan_inst.cu"]
    jobs += [("scan_inst.cu", os.path.join(obj_dir, f"scan_inst_g{k}.o"),
              [f"-DAMM_GROUP={k}", *defines]) for k in range(N_GROUPS)]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(lambda j: _compile(nvcc, j[0], j[1], j[2], verbose), jobs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path,
            *[j[1] for j in jobs], "-ldl"]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    with open(stamp, "w") as f:
        f.write(_fingerprint())
    if verbose:
        print("\n".join(logs))
    return lib_path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if the sources changed since the last build. Returns its path.
    Every translation unit is built with -gencode arch=compute_100a,code=sm_100a -lineinfo."""
    if not force and is_current():
        return LIB_PATH
    return _build_into(LIB_PATH, STAMP, OBJ_DIR, [], verbose)


def build_boundscheck(force: bool = False) -> str:
    """Debug twin of the library (-DAMM_BOUNDS_CHECK): every global load of input data is
    checked against [base, base + n*sizeof(T)); a violation makes the call fail.  Used by
    tests/test_gpu_bounds.py in place of compute-sanitizer (closed on this pool)."""
    if not force and os.path.exists(DEBUG_LIB_PATH) and os.path.exists(DEBUG_STAMP):
        with open(DEBUG_STAMP) as f:
            if f.read().strip() == _fingerprint():
                return DEBUG_LIB_PATH
    return _build_into(DEBUG_LIB_PATH, DEBUG_STAMP, OBJ_DIR + "_boundscheck",
                       ["-DAMM_BOUNDS_CHECK"], False)


STAMPS_LIB_PATH = os.path.join(LIB_DIR, "libargminmax_b200_stamps.so")
STAMPS_STAMP = os.path.join(LIB_DIR, "build_stamps.stamp")


def build_stamps(force: bool = False) -> str:
    """Debug twin (-DAMM_PHASE_STAMPS): thread 0 of every CTA records %globaltimer at the phase
    boundaries of a scan (amm_debug_read_stamps).  Used by tools/phase_stamps.py to attribute
    the fixed per-call cost; never loaded by the product path."""
    if not force and os.path.exists(STAMPS_LIB_PATH) and os.path.exists(STAMPS_STAMP):
        with open(STAMPS_STAMP) as f:
            if f.read().strip() == _fingerprint():
                return STAMPS_LIB_PATH
    return _build_into(STAMPS_LIB_PATH, STAMPS_STAMP, OBJ_DIR + "_stamps", ["-DAMM_PHASE_STAMPS"], False)


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
