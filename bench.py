#!/usr/bin/env python
"""bench.py -- fused argmin+argmax throughput on B200 against the HBM roofline.

Workload (BASELINE.json metric / configs[3]): f32, 2^30 uniform-random-bit-pattern elements
(NaN -> 0.0, the reference's bench generator dev_utils/src/utils.rs:55-77) = 4 GiB per GPU,
device-resident, ArgMinMax (ignore-NaN).  A "step" is one fused argmin+argmax over that
array.  4 GiB >> the 126 MB L2, so every step streams from HBM (no explicit L2 flush needed).

  python bench.py [--gpus 1] [--steps K] [--warmup W]          # our arm, N=1
  torchrun ... bench.py --gpus N --steps K --warmup W           # one rank per GPU, weak scaling
  python bench.py --impl reference ...                          # the reference's CPU path

Prints ONE JSON line (rank 0).  `value` = aggregate GB/s with inputs resident in HBM,
device-timed (CUDA events on the launching stream, max over ranks).  `e2e` = the same metric
through the public host-slice API with HOST (pinned) input: H2D copies inside the timed
region.  `roofline` = algorithmic bytes / kernel time against MEASURED_PEAKS.json.
`cpu_baseline` = the AVX-512/AVX restatement of the reference's dispatched f32 path
(oracle/cpu_simd_f32.c), single thread like the reference, on this box's host.

Every answer that is timed is first checked against the CPU oracle (`verified_vs_oracle`): the
2^30 scan against both the SIMD restatement and the scalar restatement of the reference on a
pinned host copy of the same bytes; at N > 1 additionally a verification step with duplicate
extremes planted in the first and the last shard, a NaN case in return-NaN mode and an empty
shard, and an in-process amm_sharded_argminmax over all N GPUs driven by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "argminmax GB/s (f32 1G elems per GPU, device-resident) vs HBM roofline"
N_ELEMS = 1 << 30
DTYPE = "f32"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def workload_config(n: int, world: int) -> dict:
    """`config` of the JSON line -- the SAME dict in both arms (ours and --impl reference)."""
    return {
        "workload": "configs[3]: f32 1G (2^30) elements = 4 GiB per GPU, ArgMinMax (ignore-NaN), "
                    "uniform random bit patterns NaN->0 (reference bench generator)",
        "elements_per_gpu": n, "bytes_per_gpu": n * 4,
        "l2": "input (4 GiB) >> 126 MB L2, every step streams from memory (no flush needed)",
        "parallelism": f"shard{world}" if world > 1 else "single",
    }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """dram bytes per launch from the committed ncu --set full capture (profiles/), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("f32_1g_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    """Polls SM clock + throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (t, sm_mhz, reasons_mask, mem_mhz, watts)
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    mem = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM)
                    watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:  # noqa: BLE001
                    mem, watts = None, None
                self.samples.append((time.perf_counter(), mhz, reasons, mem, watts))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def summary(self, t0: float, t1: float) -> dict:
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        mask = 0
        for s in inside:
            mask |= s[2]
        mems = [s[3] for s in inside if s[3] is not None]
        watts = [s[4] for s in inside if s[4] is not None]
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz,
                "reasons": [k for k, bit in names.items() if mask & bit], "samples": len(inside),
                "mem_mhz": statistics.median(mems) if mems else None,
                "power_w_max": round(max(watts), 1) if watts else None}


# ------------------------------------------------------------------------ NUMA
def parse_cpulist(spec: str) -> set:
    """sysfs cpulist ("0-3,8-11,16") -> set of CPU ids."""
    cpus = set()
    for part in spec.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def cpus_from_affinity_mask(words, bits_per_word: int = 64) -> set:
    """NVML CPU-affinity bitmask (array of unsigned longs, LSB = CPU 0) -> set of CPU ids."""
    cpus = set()
    for wi, w in enumerate(words):
        w = int(w)
        for b in range(bits_per_word):
            if (w >> b) & 1:
                cpus.add(wi * bits_per_word + b)
    return cpus


def bind_to_gpu_numa_node(torch, index: int) -> dict:
    """Run this rank (and therefore first-touch its pinned host buffers) on the CPUs next to its GPU
    -- what `numactl --cpunodebind` does for one-rank-per-GPU jobs.  With 8 ranks streaming host
    memory to 8 GPUs at once, buffers that all sit on one socket make the other socket's GPUs pull
    their input across the inter-socket link.  Source of the CPU set: sysfs (numa_node of the
    GPU's PCI function) and, when that says -1 (VMs), NVML's ideal CPU affinity of the device.
    No-op when neither knows more than "all CPUs"."""
    info = {"node": None, "cpus": None, "source": None}
    allowed = os.sched_getaffinity(0)
    try:
        pr = torch.cuda.get_device_properties(index)
        sys_id = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{sys_id}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = parse_cpulist(f.read()) & allowed
            if cpus and cpus != allowed:
                os.sched_setaffinity(0, cpus)
                info.update(cpus=len(cpus), source="sysfs numa_node")
                return info
    except Exception as e:  # noqa: BLE001 - placement is an optimisation, never a failure
        info["error"] = type(e).__name__
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        nwords = (max(allowed) // 64) + 1
        cpus = cpus_from_affinity_mask(pynvml.nvmlDeviceGetCpuAffinity(h, nwords)) & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info.update(cpus=len(cpus), source="nvmlDeviceGetCpuAffinity")
        else:
            info.update(cpus=len(allowed), source="none (NVML affinity = all CPUs: single NUMA domain)")
    except Exception as e:  # noqa: BLE001
        info["nvml_error"] = type(e).__name__
    return info


# ------------------------------------------------------------------------ data
def make_device_data(torch, device, n: int, seed: int):
    """Uniform random bit patterns viewed as f32, NaN -> 0.0 (dev_utils/src/utils.rs:55-77)."""
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.empty(n, dtype=torch.int32, device=device)
    chunk = 1 << 28
    for i in range(0, n, chunk):
        t[i:i + chunk].random_(-2**31, 2**31 - 1, generator=g)
    f = t.view(torch.float32)
    for i in range(0, n, chunk):
        v = f[i:i + chunk]
        v[torch.isnan(v)] = 0
    return f


def make_host_data(np, n: int, seed: int = 0):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32)
    a[np.isnan(a)] = 0
    return a


# ------------------------------------------------------------------- CPU baseline
def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(host_np, budget_s: float, all_cores: bool = True) -> dict:
    """Single-thread AVX-512/AVX restatement of the reference's dispatched f32 path on the SAME
    2^30-element workload (a bounded number of passes); the all-cores variant is NOT something the
    reference does (labelled so)."""
    from oracle import oracle
    nbytes = host_np.nbytes
    times = []
    t_start = time.perf_counter()
    res = None
    while True:
        t0 = time.perf_counter()
        res = oracle.cpu_simd_argminmax_f32(host_np)
        times.append(time.perf_counter() - t0)
        if len(times) >= 3 and (time.perf_counter() - t_start > budget_s or len(times) >= 20):
            break
    out = {
        "value": round(nbytes / min(times) / 1e9, 2), "median": round(nbytes / statistics.median(times) / 1e9, 2),
        "unit": "GB/s", "cores": 1, "kind": "port", "isa": oracle.cpu_simd_isa(),
        "sample": f"the whole workload: f32 x {host_np.size} elements ({nbytes >> 20} MiB, DRAM-resident), best of "
                  f"{len(times)} passes, oracle/cpu_simd_f32.c (restatement of src/simd/simd_f32_ignore_nan.rs + "
                  "generic.rs + task.rs), 1 thread as in the reference",
        "host_cpus": os.cpu_count(), "cpu_model": cpu_model(), "result": list(res),
    }
    if all_cores:
        nt = os.cpu_count() or 1
        mt = []
        for _ in range(3):
            t0 = time.perf_counter()
            r2 = oracle.cpu_simd_argminmax_f32_mt(host_np, nt)
            mt.append(time.perf_counter() - t0)
        out["all_cores_not_reference"] = {"value": round(nbytes / min(mt) / 1e9, 2), "unit": "GB/s",
                                          "threads": nt, "agrees": list(r2) == list(res)}
    return out


def cpu_small_n_latency(np, n: int, calls: int = 20000) -> dict:
    """The reference's own bench size (dev_utils/src/config.rs:4: n = 102 400 f32, cache-resident):
    us per call of the CPU port, timed in C-call granularity (ctypes overhead ~1 us included)."""
    from oracle import oracle
    a = make_host_data(np, n, seed=3)
    for _ in range(200):
        oracle.cpu_simd_argminmax_f32(a)
    t0 = time.perf_counter()
    for _ in range(calls):
        oracle.cpu_simd_argminmax_f32(a)
    us = (time.perf_counter() - t0) / calls * 1e6
    return {"n": n, "us_per_call": round(us, 2), "calls": calls, "isa": oracle.cpu_simd_isa(),
            "note": "cache-resident, 1 thread, through ctypes (adds ~0.5-1 us); the reference publishes 6.975 us on an "
                    "i7-1185G7 (benches/results:15)"}


def run_reference(args) -> int:
    """--impl reference: the reference's own CPU path (restated in C, see oracle/) on this
    box's host cores, on the SAME workload as our arm (f32 x 2^30 per step).  The crate is
    single-threaded, so one thread is all the host threads it can use; an all-cores figure is
    added and labelled as not the reference.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from oracle import oracle
    n = args.elems
    a = make_host_data(np, n, seed=0)
    for _ in range(min(args.warmup, 3)):
        oracle.cpu_simd_argminmax_f32(a)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = oracle.cpu_simd_argminmax_f32(a)
    dt = time.perf_counter() - t0
    verified = res == oracle.argminmax(a, 0)
    assert verified, "SIMD restatement disagrees with the scalar oracle"
    gbs = a.nbytes * args.steps / dt / 1e9
    nt = os.cpu_count() or 1
    t1 = time.perf_counter()
    for _ in range(3):
        oracle.cpu_simd_argminmax_f32_mt(a, nt)
    mt_gbs = a.nbytes * 3 / (time.perf_counter() - t1) / 1e9
    sample_desc = (f"the whole workload per step: f32 x {n} elements ({a.nbytes >> 20} MiB), single thread (the "
                   f"reference crate is single-threaded), ISA {oracle.cpu_simd_isa()} chosen like src/lib.rs:427-443")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, args.gpus),
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": 1, "kind": "port",
                         "isa": oracle.cpu_simd_isa(), "sample": sample_desc, "host_cpus": nt, "cpu_model": cpu_model(),
                         "all_cores_not_reference": {"value": round(mt_gbs, 2), "threads": nt}},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "verified_vs_oracle": verified, "result": list(res),
        "where": "host CPU only (the reference has no GPU path); under torchrun rank 0 alone runs",
    }
    emit(line)
    return 0


# ----------------------------------------------------------------- helper probes
def events_us(torch, fn, iters: int, batch: int = 0) -> float:
    """Mean device time per call of `fn` over `iters` back-to-back calls, CUDA events on the current
    stream.  `batch` > 0: the calls go out in batches of that many with a sync in between, so that
    the host never runs hundreds of launches ahead of the device (with 200 scans queued at once
    the same serialised 4 GiB scans measure 0.593 ms instead of 0.575)."""
    total_ms, done = 0.0, 0
    step = batch if batch > 0 else iters
    while done < iters:
        m = min(step, iters - done)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(done, done + m):
            fn(i)
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        done += m
    return total_ms / iters * 1e3


def serialised_probes(torch, amm, ws, data, stream) -> dict:
    """Single-call (serialised, non-pipelined) device time of the scan at BASELINE config 1 / 2
    sizes, cold L2: consecutive launches walk through distinct segments of the 4 GiB array."""
    ws.set_pipelined(False)
    out = {}
    raw = data.view(torch.uint8)
    total = raw.numel()
    for key, dt, n in (("f32_10m", "f32", 10_000_000), ("f32_100m", "f32", 100_000_000),
                       ("u8_100m", "u8", 100_000_000), ("f32_102400", "f32", 102_400)):
        nbytes = n * amm.DTYPE_SIZE[dt]
        stride = -(-nbytes // 4096) * 4096
        segs = max(1, min(64, total // stride))
        # enough distinct segments that a revisit finds its lines evicted (>= 512 MB walked)
        slices = [amm.DeviceSlice.from_torch(raw[i * stride: i * stride + nbytes], dtype=dt, workspace=ws, stream=stream)
                  for i in range(segs)]
        for s in slices[:3]:
            s.launch(amm.IGNORE_NAN)
        slices[0].wait(amm.IGNORE_NAN)
        iters = 200 if nbytes <= (64 << 20) else 60
        us = events_us(torch, lambda i: slices[i % segs].launch(amm.IGNORE_NAN), iters, batch=50)
        slices[0].wait(amm.IGNORE_NAN)
        out[key] = {"us": round(us, 2), "gbs": round(nbytes / us / 1e3, 1), "segments": segs,
                    "cold_l2": segs * stride >= (256 << 20)}
    out["10m_us"] = out["f32_10m"]["us"]
    out["100m_gbs"] = out["f32_100m"]["gbs"]
    out["what"] = ("device time per call, CUDA events over back-to-back NON-pipelined launches (each scan alone on the "
                   "GPU), rotating over distinct segments of the resident array so that L2 is cold")
    return out


def dtype_table(torch, np, amm, ws, data, host_np, stream, peak: float, verify: bool) -> list:
    """All 14 (dtype, mode) instantiations: serialised GB/s at 4 GiB and at 100 M elements (BASELINE
    config 2), the raw bytes of the f32 array reinterpreted (performance is data-independent:
    the loop is branch-free).  The 100 M rows are checked against the oracle on the host copy."""
    from oracle import oracle
    ws.set_pipelined(False)
    raw = data.view(torch.uint8)
    np_dt = {"i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64, "u8": np.uint8, "u16": np.uint16,
             "u32": np.uint32, "u64": np.uint64, "f16": np.float16, "f32": np.float32, "f64": np.float64}
    rows = []
    for dt in amm.DTYPES:
        es = amm.DTYPE_SIZE[dt]
        for mode in ((0, 1) if dt in amm.FLOAT_DTYPES else (0,)):
            row = {"dtype": dt, "mode": "return_nan" if mode else "ignore_nan"}
            full = amm.DeviceSlice.from_torch(raw, dtype=dt, workspace=ws, stream=stream)
            for _ in range(2):
                full.launch(mode)
            full.wait(mode)
            us = events_us(torch, lambda i: full.launch(mode), 8)
            full.wait(mode)
            row["gbs_4gib"] = round(raw.numel() / us / 1e3, 1)
            row["frac_4gib"] = round(raw.numel() / us / 1e3 / peak, 4)
            n = 100_000_000
            nbytes = n * es
            stride = -(-nbytes // 4096) * 4096
            segs = max(1, raw.numel() // stride)
            sl = [amm.DeviceSlice.from_torch(raw[i * stride: i * stride + nbytes], dtype=dt, workspace=ws, stream=stream)
                  for i in range(segs)]
            for s in sl[:2]:
                s.launch(mode)
            got = sl[0].wait(mode) if segs < 2 else None
            if got is None:
                sl[0].launch(mode)
                got = sl[0].wait(mode)
            us = events_us(torch, lambda i: sl[i % segs].launch(mode), 24)
            sl[0].wait(mode)
            row["us_100m"] = round(us, 2)
            row["gbs_100m"] = round(nbytes / us / 1e3, 1)
            row["frac_100m"] = round(nbytes / us / 1e3 / peak, 4)
            if verify and host_np is not None:
                ref = oracle.argminmax(host_np.view(np.uint8)[:nbytes].view(np_dt[dt]), mode)
                row["parity_100m"] = (tuple(got) == tuple(ref))
                if not row["parity_100m"]:
                    raise AssertionError(f"dtype table: {dt} mode {mode}: gpu {got} oracle {ref}")
            rows.append(row)
    return rows


def h2d_ceiling(torch, host, device, barrier, reps: int = 3) -> float:
    """Plain pinned cudaMemcpyAsync of the whole host buffer, no scan: seconds for this rank's
    4 GiB (best of `reps`), every repetition started at a barrier so that all N ranks pull from
    host memory AT THE SAME TIME, as they do in the e2e step."""
    dst = torch.empty(1 << 28, dtype=torch.float32, device=device)  # 1 GiB landing buffer, reused
    n = host.numel()
    chunk = dst.numel()
    best = float("inf")
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        for i in range(0, n, chunk):
            m = min(chunk, n - i)
            dst[:m].copy_(host[i:i + m], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


_REAL_STDOUT = None


def protect_stdout() -> None:
    """The driver parses ONE JSON line from stdout.  Libraries print there too (NCCL's version
    banner on communicator creation, for one), so fd 1 is pointed at stderr for the whole run and
    the line is written to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# --------------------------------------------------------------------------- main
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=N_ELEMS, help="elements per GPU (default 2^30)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the dtype table, serialised probes, strong scaling, config 5 (ncu runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    protect_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import argminmax_b200 as amm
    from oracle import oracle  # the checker: verifies the timed answers, and the cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    host_group = None
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=device)
        host_group = dist.new_group(backend="gloo")  # CPU-side barriers that leave the GPUs idle
    amm.lib()  # fail loudly if the extension is missing
    numa = bind_to_gpu_numa_node(torch, local_rank)

    n = args.elems
    nbytes = n * 4
    data = make_device_data(torch, device, n, seed=1234 + rank)
    torch.cuda.synchronize()

    ws = amm.Workspace(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    shard = amm.DeviceSlice.from_torch(data, dtype=DTYPE, workspace=ws, stream=stream)
    sharded = amm.DistributedShardedSlice(shard) if distributed else None
    xws = sharded.workspace if sharded is not None else ws

    def enqueue_step():
        if sharded is not None:
            sharded.launch(amm.IGNORE_NAN)
        else:
            shard.launch(amm.IGNORE_NAN)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag: bool) -> bool:
        if not distributed:
            return flag
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def gather_rows(vals):
        """every rank contributes a list of float64; returns world x len rows on every rank"""
        mine = torch.tensor(vals, dtype=torch.float64, device=device)
        if not distributed:
            return [mine.cpu().tolist()]
        allv = torch.zeros(world * len(vals), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allv, mine)
        return allv.view(world, len(vals)).cpu().tolist()

    # pinned host copy of this rank's shard: the oracle's input, and later the e2e input
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    host.copy_(data)
    torch.cuda.synchronize()
    host_np = host.numpy()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # Consecutive steps scan a resident, unchanging array: let step k+1's streaming phase overlap
    # step k's fold/publish tail (programmatic dependent launch, amm_workspace_set_pipelined).
    xws.set_pipelined(True)
    ws.set_pipelined(True)

    # ---- warm-up (untimed) + the answer we are about to time, checked against the oracle
    for _ in range(args.warmup):
        enqueue_step()
    result = (sharded or shard).wait(amm.IGNORE_NAN)
    verification = {}
    t_v = time.perf_counter()
    local_simd = oracle.cpu_simd_argminmax_f32(host_np)      # the reference's dispatched SIMD path
    local_gpu = shard.argminmax()
    ok_local = tuple(local_gpu) == tuple(local_simd)
    if rank == 0:
        local_scalar = oracle.argminmax(host_np, oracle.IGNORE_NAN)  # ... and its scalar path
        ok_local = ok_local and tuple(local_gpu) == tuple(local_scalar)
        verification["oracles"] = ["oracle.cpu_simd_argminmax_f32 (every rank)", "oracle.argminmax scalar (rank 0)"]
    # expected global answer = lexicographic (value, index) fold of the ORACLE's per-shard answers
    li, lj = local_simd
    rows = gather_rows([float(host_np[li]), float(rank * n + li), float(host_np[lj]), float(rank * n + lj)])
    exp = (int(min(rows, key=lambda r: (r[0], r[1]))[1]), int(min(rows, key=lambda r: (-r[2], r[3]))[3]))
    ok_global = tuple(result) == exp
    verified = all_true(ok_local and ok_global)
    verification.update({"timed_answer": list(result), "oracle_answer": list(exp), "local_ok": ok_local,
                         "seconds": round(time.perf_counter() - t_v, 2)})
    if not verified:
        raise SystemExit(f"rank {rank}: timed answer {result} / local {local_gpu} disagrees with the oracle "
                         f"(global {exp}, local {local_simd})")
    launches0 = xws.last_launch()["launches_total"]

    # ---- timed region: exactly K steps, device-timed, barrier + sync on both sides
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        enqueue_step()
    e1.record()
    e1.synchronize()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    result2 = (sharded or shard).wait(amm.IGNORE_NAN)
    if tuple(result2) != exp:
        raise SystemExit(f"rank {rank}: answer after the timed region {result2} != oracle {exp}")
    launches = xws.last_launch()["launches_total"] - launches0
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / args.steps
    value = world * nbytes / ms_per_step / 1e6  # GB/s, aggregate
    clocks = sampler.summary(t_wall0, t_wall1)
    main_launch = xws.last_launch()
    tuning = xws.get_tuning()  # geometry of the timed scans (later probes use other shapes)

    # ---- kernel-only duration of the dominant kernel (roofline), this rank, same stream,
    #      launches strictly serialised (no overlap between consecutive kernels)
    ws.set_pipelined(False)
    xws.set_pipelined(False)
    for _ in range(2):
        shard.launch(amm.IGNORE_NAN)
    kernel_ms = events_us(torch, lambda i: shard.launch(amm.IGNORE_NAN), args.steps, batch=20) / 1e3
    shard.wait(amm.IGNORE_NAN)
    peak, peak_src = measured_peak()
    achieved = nbytes / kernel_ms / 1e6
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": committed_traffic(),
                "traffic_source": "committed ncu --set full capture (profiles/traffic.json), not measured in this run",
                "kernel": ("argminmax_ring_kernel<PolF32<ignore>>" if "ring" in str(tuning.get("variant_name", ""))
                           else "argminmax_scan_kernel<PolF32<ignore>>"),
                "kernel_ms": round(kernel_ms, 5),
                "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                "note": "kernel_ms: serialised launches (each kernel alone on the GPU); `value` is measured with "
                        "consecutive steps pipelined (PDL), which hides each kernel's fold/publish tail",
                "frac_of_nominal_8TBps": round(achieved / 8000.0, 4)}

    extras = {}
    if not args.no_extras:
        # the probes below are single calls: let the GPU settle after the streaming loops above (a
        # B200 that has just streamed for a while answers short kernels ~1.7 us slower for some
        # hundred ms: tools/exp/latency_after_load.py)
        time.sleep(0.5)
        # ---- strong scaling: 2^30 elements IN TOTAL over the N GPUs (each scans 2^30/N of its
        #      buffer), serialised collective calls; N=1 reference = kernel_ms above
        if distributed:
            per = n // world
            part = amm.DeviceSlice.from_torch(data[:per], dtype=DTYPE, workspace=ws, stream=stream)
            sp = amm.DistributedShardedSlice(part)  # private workspace (the first slice owns ws)
            for _ in range(3):
                sp.launch(amm.IGNORE_NAN)
            sp.wait(amm.IGNORE_NAN)
            barrier()
            strong_ms = max_over_ranks(events_us(torch, lambda i: sp.launch(amm.IGNORE_NAN), 50) / 1e3)
            sp.wait(amm.IGNORE_NAN)
            sp.workspace.set_pipelined(True)
            for _ in range(3):
                sp.launch(amm.IGNORE_NAN)
            barrier()
            strong_pipe_ms = max_over_ranks(events_us(torch, lambda i: sp.launch(amm.IGNORE_NAN), 50) / 1e3)
            sp.wait(amm.IGNORE_NAN)
            sp.workspace.set_pipelined(False)
            # host-visible latency of one synchronous collective call, and where sharding starts
            # to pay: total sizes 2^18 .. 2^30 against the single-GPU call on rank 0's buffer
            # (timed inside the library: no Python between the calls)
            cross = []
            out3 = (ctypes.c_double * 3)()
            for lg in (18, 20, 22, 24, 26, 28, 30):
                tot = 1 << lg
                pr = tot // world
                us_sh = None
                if sp.exchange == "peer":
                    barrier()
                    rc = amm.lib().amm_measure_exchange_latency(amm.DTYPES[DTYPE], data.data_ptr(), pr, amm.IGNORE_NAN,
                                                                rank * pr, sp.workspace.handle, stream, 200, out3)
                    if rc:
                        raise RuntimeError(f"amm_measure_exchange_latency rc={rc}")
                    us_sh = max_over_ranks(out3[0])
                rc = amm.lib().amm_measure_call_latency(amm.DTYPES[DTYPE], data.data_ptr(), tot, amm.IGNORE_NAN,
                                                        ws.handle, stream, 200, out3)
                if rc:
                    raise RuntimeError(f"amm_measure_call_latency rc={rc}")
                rows1 = gather_rows([out3[0]])
                cross.append({"total_elems_log2": lg, "us_sharded": round(us_sh, 2) if us_sh is not None else None,
                              "us_one_gpu": round(rows1[0][0], 2)})
            first_win = next((c["total_elems_log2"] for c in cross
                              if c["us_sharded"] is not None and c["us_sharded"] < c["us_one_gpu"]), None)
            sp.close()
            # this box's single-GPU pipelined time for the whole 2^30 (denominator of the pipelined efficiency)
            ws.set_pipelined(True)
            for _ in range(3):
                shard.launch(amm.IGNORE_NAN)
            n1_pipe_ms = events_us(torch, lambda i: shard.launch(amm.IGNORE_NAN), 20) / 1e3
            shard.wait(amm.IGNORE_NAN)
            ws.set_pipelined(False)
            extras["strong"] = {
                "total_elems": n, "per_gpu_elems": per, "ms": round(strong_ms, 5),
                "gbs": round(nbytes / strong_ms / 1e6, 1),
                "efficiency_vs_n1": round(kernel_ms / (world * strong_ms), 4),
                "ms_pipelined": round(strong_pipe_ms, 5),
                "efficiency_vs_n1_pipelined": round(n1_pipe_ms / (world * strong_pipe_ms), 4),
                "n1_ms_same_box": round(kernel_ms, 5), "n1_ms_pipelined_same_box": round(n1_pipe_ms, 5),
                "what": "one logical f32 array of 2^30 elements split over the N GPUs; device time per serialised "
                        "collective call (scan fused with the in-kernel record exchange), max over ranks; "
                        "efficiency = t(1 GPU, 2^30, serialised) / (N x t(N GPUs)); sync_call_us_by_total_size: median host "
                        "wall clock of the synchronous collective call vs the single-GPU call on the same total size, timed "
                        "inside the library",
                "sync_call_us_by_total_size": cross, "sharding_pays_from_log2_elems": first_win,
            }
            # ---- config 5: 32 Gi f32 elements (128 GiB) IN TOTAL over the N GPUs: this rank's 4 GiB
            #      block tiled 32/N times; the expected answer follows from the oracle-checked block
            reps = 32 // world
            try:
                big = data.repeat(reps)
                bshard = amm.DeviceSlice.from_torch(big, dtype=DTYPE, workspace=ws, stream=stream)
                bs = amm.DistributedShardedSlice(bshard)
                bs.workspace.set_pipelined(True)
                for _ in range(2):
                    bs.launch(amm.IGNORE_NAN)
                got5 = bs.wait(amm.IGNORE_NAN)
                per5 = n * reps
                rows5 = gather_rows([float(host_np[li]), float(rank * per5 + li), float(host_np[lj]), float(rank * per5 + lj)])
                exp5 = (int(min(rows5, key=lambda r: (r[0], r[1]))[1]), int(min(rows5, key=lambda r: (-r[2], r[3]))[3]))
                ok5 = all_true(tuple(got5) == exp5)
                barrier()
                ms5 = max_over_ranks(events_us(torch, lambda i: bs.launch(amm.IGNORE_NAN), 10) / 1e3)
                bs.wait(amm.IGNORE_NAN)
                bs.close()
                extras["config5"] = {"total_elems": per5 * world, "per_gpu_gib": per5 * 4 / 2**30, "ms": round(ms5, 4),
                                     "gbs": round(world * per5 * 4 / ms5 / 1e6, 1), "verified_vs_oracle": ok5,
                                     "how_verified": "each GPU holds its oracle-checked 4 GiB block tiled; first occurrences "
                                                     "lie in the first tile, so the expected global answer is the "
                                                     "lexicographic fold of the per-block oracle answers"}
                del big, bshard, bs
                torch.cuda.empty_cache()
            except torch.OutOfMemoryError:
                extras["config5"] = {"skipped": "not enough free HBM for 128 GiB / N per GPU next to the resident buffers"}

            # ---- multi-GPU parity step (before any number of this section is trusted): duplicate
            #      extremes in the first and the last shard, NaN in return mode, an empty shard
            extras["multi_gpu_parity"] = multi_gpu_parity(torch, np, dist, amm, oracle, device, rank, world, ws, stream,
                                                          gather_rows, all_true)
            # ---- in-process amm_sharded_argminmax over all N GPUs, driven by rank 0 alone
            extras["sharded_in_process"] = sharded_in_process(torch, np, dist, amm, oracle, rank, world, host_group)
        else:
            extras["serialised"] = serialised_probes(torch, amm, ws, data, stream)
            extras["dtype_table"] = dtype_table(torch, np, amm, ws, data, host_np, stream, peak, verify=True)

    # the clock sampler belongs to the timed regions above; its NVML polling thread would
    # contend with the launch path of the latency probe below
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    # ---- small-n latency: us per call, synchronous C-ABI call timed inside the library
    #      (amm_measure_call_latency: host wall clock, no Python in the loop)
    # A B200 that has just streamed at full HBM rate answers small calls ~1.7 us slower for a few
    # hundred ms (tools/exp/latency_after_load.py: 12.8 us right after 400 x 4 GiB scans, 11.0 us
    # fresh and again 1 s later).  The probe measures the settled state a latency-bound caller sees.
    time.sleep(1.0)
    lat3 = (ctypes.c_double * 3)()
    lat_by_n = {}
    for nn in (1 << 20, 1 << 17, 102_400):
        rc = amm.lib().amm_measure_call_latency(amm.DTYPES[DTYPE], data.data_ptr(), nn, amm.IGNORE_NAN,
                                                ws.handle, stream, 2000, lat3)
        if rc:
            raise RuntimeError(f"amm_measure_call_latency rc={rc}")
        lat_by_n[nn] = {"median": round(lat3[0], 2), "min": round(lat3[1], 2), "p99": round(lat3[2], 2)}
    us_1m = lat_by_n[1 << 20]["median"]
    small = amm.DeviceSlice.from_torch(data[: 1 << 20], dtype=DTYPE, workspace=ws, stream=stream)
    for _ in range(20):
        small.argminmax()
    lat = []
    for _ in range(300):
        t0 = time.perf_counter()
        small.argminmax()
        lat.append((time.perf_counter() - t0) * 1e6)
    us_1m_python = statistics.median(lat)

    # ---- e2e: the public host-slice call on pinned HOST input (H2D inside the timed region)
    e2e = None
    if not args.no_e2e:
        hs = amm.HostSlice.from_torch(host, dtype=DTYPE, device=local_rank, workspace=ws)
        rec = amm.AmmRecord()

        def e2e_step():
            rc = amm.lib().amm_host_argminmax_record(amm.DTYPES[DTYPE], hs.host_ptr, hs.n, amm.IGNORE_NAN,
                                                     rank * n, ws.handle, ctypes.byref(rec))
            if rc:
                raise RuntimeError(f"amm_host_argminmax_record rc={rc}")
            if distributed:
                return amm.combine_across_ranks(rec, amm.IGNORE_NAN)
            return amm.combine_records([rec], amm.IGNORE_NAN)

        e2e_steps = args.e2e_steps or min(args.steps, 10)
        r_e2e = e2e_step()
        if tuple(r_e2e) != exp:
            raise SystemExit(f"e2e answer {r_e2e} != oracle {exp}")
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        # the ceiling of this leg: plain pinned H2D copies on all N GPUs at once, no scan.  The e2e
        # step is lock-step (its answer is a combine over all ranks), so the comparable ceiling is
        # N x bytes / the SLOWEST rank's plain-copy time; the sum of the ranks' own rates is
        # reported next to it (on these 8-GPU boxes four GPUs get 23 GB/s and four 36 GB/s).
        ceil_rows = gather_rows([h2d_ceiling(torch, host, device, barrier)])
        per_gpu = [nbytes / r[0] / 1e9 for r in ceil_rows]
        ceiling = world * nbytes / max(r[0] for r in ceil_rows) / 1e9
        chunks = -(-nbytes // (32 << 20))
        e2e_val = world * nbytes * e2e_steps / dt / 1e9
        e2e = {"value": round(e2e_val, 2), "unit": "GB/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 64 * chunks,
               "steps": e2e_steps, "ms_per_step": round(dt / e2e_steps * 1e3, 3),
               "h2d_ceiling_gbs": round(ceiling, 2), "frac_of_h2d_ceiling": round(e2e_val / ceiling, 4),
               "h2d_ceiling_sum_of_rank_rates_gbs": round(sum(per_gpu), 2),
               "h2d_ceiling_per_gpu_gbs": [round(x, 2) for x in per_gpu],
               "verified_vs_oracle": True,
               "api": "amm_host_argminmax_record (HostSlice): pinned host f32 -> chunked H2D overlapped with the scan; "
                      "PCIe-bound; h2d_ceiling_gbs = plain pinned cudaMemcpyAsync of the same buffers on all N GPUs, "
                      "every repetition started at a barrier, no scan: N x bytes / slowest rank"}

    # ---- CPU baseline beside it (rank 0, N=1 only), on the same 2^30-element workload
    cpu = None
    small_cpu = None
    if rank == 0 and not distributed and not args.no_cpu_baseline:
        cpu = cpu_baseline(host_np, budget_s=6.0)
        small_cpu = cpu_small_n_latency(np, 102_400)

    ll = main_launch
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(n, world),
            "launch": {
                "residency": "device-resident (HBM) for `value`; pinned host memory for `e2e`",
                "collective": (None if not distributed else
                               "fused in-kernel peer exchange of 64-byte records over NVLink (CUDA IPC), no NCCL on the step"
                               if sharded.exchange == "peer" else
                               "nccl all_gather of 64-byte records + combine kernel"),
                "tuning": tuning, "grid": ll["grid"], "block": ll["threads"],
                "host_numa_binding": numa,
            },
            "elements_per_s": round(world * n / (ms_per_step / 1e3), 1),
            "verified_vs_oracle": verified, "verification": verification,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "us_per_call_1m": round(us_1m, 2),
            "us_per_call_1m_detail": {**lat_by_n[1 << 20],
                                      "through_python_ctypes_median": round(us_1m_python, 2),
                                      "what": "f32 n=2^20, synchronous amm_argminmax, host wall clock, GPU idle for 1 s "
                                              "before the probe (right after the streaming loops: +1.7 us)"},
            "us_per_call": {"128Ki": lat_by_n[1 << 17], "102400": lat_by_n[102_400],
                            "cpu_port_102400": small_cpu,
                            "what": "synchronous C-ABI call (amm_measure_call_latency); 102 400 is the reference's own bench "
                                    "size (dev_utils/src/config.rs:4), where the CPU port is timed beside it"},
            **extras,
            "result": list(result), "clocks": clocks,
        }
        emit(line)
    if distributed:
        if sharded is not None:
            sharded.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _lex_fold(rows):
    """rows: per shard (has_value, min_val, min_idx, max_val, max_idx, nan_idx or -1) in shard
    order -> (argmin, argmax, first_nan or -1) by the reference's merge rule."""
    have = [r for r in rows if r[0]]
    nans = [r[5] for r in rows if r[5] >= 0]
    first_nan = int(min(nans)) if nans else -1
    if not have:
        return 0, 0, first_nan
    imin = int(min(have, key=lambda r: (r[1], r[2]))[2])
    imax = int(min(have, key=lambda r: (-r[3], r[4]))[4])
    return imin, imax, first_nan


def multi_gpu_parity(torch, np, dist, amm, oracle, device, rank, world, ws, stream, gather_rows, all_true) -> dict:
    """One rank per GPU, fused exchange: cases the random timed array cannot cover.  Every rank
    checks its LOCAL answer against the oracle on its shard and the GLOBAL answer against the
    lexicographic fold of the oracle's per-shard answers (src/simd/generic.rs:513-520: the
    earlier chunk wins ties; src/simd/test_utils.rs:89-151: first index)."""
    m = 1 << 22
    cases = []
    base = make_host_data(np, m, seed=77 + rank)
    # A: the SAME min and max value in shard 0 and in shard N-1 (and a decoy pair in between)
    a = base.copy()
    if rank in (0, world - 1):
        a[m // 3] = -np.inf
        a[m // 3 + 7] = np.inf
        a[m - 5] = -np.inf
        a[m - 9] = np.inf
    cases.append(("dup_extremes_first_and_last_shard", a, amm.IGNORE_NAN))
    cases.append(("dup_extremes_return_mode_no_nan", a, amm.RETURN_NAN))
    # B: NaNs in two shards, return mode -> the first NaN of the lowest shard wins for both
    b = base.copy()
    if rank in (world - 1, max(0, world - 2)):
        b[1000 + rank] = np.nan
        b[m - 1] = np.nan
    cases.append(("nan_in_last_two_shards_return_mode", b, amm.RETURN_NAN))
    cases.append(("nan_in_last_two_shards_ignore_mode", b, amm.IGNORE_NAN))
    # C: one empty shard (the middle rank, or the last one at N=2)
    empty_rank = world // 2 if world > 2 else world - 1
    c = a if rank != empty_rank else a[:0]
    cases.append(("one_empty_shard", c, amm.IGNORE_NAN))
    # D: all equal -> (0, 0) globally
    d = np.full(m, 1.5, dtype=np.float32)
    cases.append(("all_equal", d, amm.IGNORE_NAN))
    report = {}
    ok_all = True
    for name, arr, mode in cases:
        dev = torch.from_numpy(arr.copy()).to(device) if arr.size else torch.empty(0, dtype=torch.float32, device=device)
        sl = amm.DeviceSlice.from_torch(dev, dtype="f32", workspace=ws, stream=stream) if arr.size else \
            amm.DeviceSlice(0, 0, "f32", device.index, ws, stream)
        ds = amm.DistributedShardedSlice(sl)
        got = ds.nanargminmax() if mode == amm.RETURN_NAN else ds.argminmax()
        off = ds.global_offset
        if arr.size:
            lo = oracle.argminmax(arr, mode)
            lg = sl.nanargminmax() if mode == amm.RETURN_NAN else sl.argminmax()
            local_ok = tuple(lg) == tuple(lo)
            isn = np.isnan(arr)
            nan_idx = int(np.argmax(isn)) + off if isn.any() else -1
            oi = oracle.argminmax(arr, amm.IGNORE_NAN)
            has = not isn.all()
            row = [1.0 if has else 0.0, float(arr[oi[0]]) if has else 0.0, float(off + oi[0]),
                   float(arr[oi[1]]) if has else 0.0, float(off + oi[1]), float(nan_idx)]
        else:
            local_ok = True
            row = [0.0, 0.0, 0.0, 0.0, 0.0, -1.0]
        rows = gather_rows(row)
        imin, imax, first_nan = _lex_fold(rows)
        exp = (first_nan, first_nan) if (mode == amm.RETURN_NAN and first_nan >= 0) else (imin, imax)
        ok = all_true(local_ok and tuple(got) == exp)
        report[name] = {"ok": ok, "got": list(got), "expected": list(exp), "exchange": ds.exchange}
        ok_all = ok_all and ok
        ds.close()
    report["parity"] = ok_all
    if not ok_all:
        raise SystemExit(f"multi-GPU parity step failed: {json.dumps(report)}")
    return report


def sharded_in_process(torch, np, dist, amm, oracle, rank, world, host_group) -> dict:
    """Rank 0 alone drives an in-process amm_sharded_argminmax (what the Rust
    ShardedDeviceSlice<T> binds) over all N GPUs, fused peer exchange and NCCL all-gather, on small
    oracle-checked arrays; the other ranks wait at a CPU-side barrier with their GPUs idle."""
    out = None
    if rank == 0:
        out = {"parity": True}
        m = 1 << 20
        host = make_host_data(np, m * world, seed=4242)
        host[5] = -np.inf
        host[m * (world - 1) + 11] = -np.inf       # duplicate min in the first and the last shard
        host[17] = np.inf
        host[m * (world - 1) + 3] = np.inf
        with_nan = host.copy()
        with_nan[m * (world - 1) + 100] = np.nan
        with_nan[m + 9 if world > 1 else 9] = np.nan
        for mode_name in ("peer", "nccl"):
            os.environ["AMM_SHARDED_EXCHANGE"] = mode_name
            rep = {}
            try:
                def shards_of(arr, lens):
                    res, o = [], 0
                    for g, ln in enumerate(lens):
                        t = torch.from_numpy(arr[o:o + ln].copy()).to(f"cuda:{g}") if ln else \
                            torch.empty(0, dtype=torch.float32, device=f"cuda:{g}")
                        res.append(amm.DeviceSlice.from_torch(t, dtype="f32") if ln else amm.DeviceSlice(0, 0, "f32", g))
                        o += ln
                    return res
                even = [m] * world
                ragged = [m] * world
                ragged[world // 2 if world > 2 else world - 1] = 0      # an empty shard
                ragged[0] = m * world - sum(ragged[1:])
                for label, arr, lens, mode in (("dup_extremes", host, even, 0), ("nan_return", with_nan, even, 1),
                                               ("nan_ignore", with_nan, even, 0), ("empty_shard", host, ragged, 0)):
                    sh = amm.ShardedDeviceSlice(shards_of(arr, lens))
                    got = sh.nanargminmax() if mode else sh.argminmax()
                    exp = oracle.argminmax(arr, mode)
                    rep[label] = {"ok": tuple(got) == tuple(exp), "got": list(got), "expected": list(exp)}
                    rep["uses_peer"], rep["uses_nccl"] = sh.uses_peer_exchange, sh.uses_nccl
                    if label == "dup_extremes":
                        for _ in range(20):
                            sh.argminmax()
                        t0 = time.perf_counter()
                        for _ in range(200):
                            sh.argminmax()
                        rep["us_per_call"] = round((time.perf_counter() - t0) / 200 * 1e6, 2)
                    sh.close()
                    out["parity"] = out["parity"] and rep[label]["ok"]
            except Exception as e:  # noqa: BLE001 - reported, and fails the run below
                rep["error"] = f"{type(e).__name__}: {e}"
                out["parity"] = False
            out[mode_name] = rep
        os.environ.pop("AMM_SHARDED_EXCHANGE", None)
        out["us_per_call"] = out.get("peer", {}).get("us_per_call")
        out["what"] = (f"rank 0 alone, one host thread, {world} GPUs x 2^20 f32 per call, answers vs oracle.argminmax over "
                       "the concatenated host array")
    dist.barrier(group=host_group)
    if rank == 0 and not out["parity"]:
        raise SystemExit(f"in-process sharded parity failed: {json.dumps(out)}")
    return out


if __name__ == "__main__":
    sys.exit(main())
