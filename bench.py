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
"""
from __future__ import annotations

import argparse
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


def bind_to_gpu_numa_node(torch, index: int) -> dict:
    """Run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA
    node its GPU hangs off -- what `numactl --cpunodebind` does for one-rank-per-GPU jobs.  With 8
    ranks streaming host memory to 8 GPUs at once, buffers that all sit on one socket make the
    other socket's GPUs pull their input across the inter-socket link.  No-op when sysfs has no
    NUMA information (single-node VMs report -1)."""
    info = {"node": None, "cpus": None}
    try:
        pr = torch.cuda.get_device_properties(index)
        sys_id = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{sys_id}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = parse_cpulist(spec) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception as e:  # noqa: BLE001 - placement is an optimisation, never a failure
        info["error"] = type(e).__name__
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
    """Single-thread AVX-512/AVX restatement of the reference's dispatched f32 path on a
    bounded sample; the all-cores variant is NOT something the reference does (labelled so)."""
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
        "sample": f"f32 x {host_np.size} elements ({nbytes >> 20} MiB, DRAM-resident), best of {len(times)} passes, "
                  "oracle/cpu_simd_f32.c (restatement of src/simd/simd_f32_ignore_nan.rs + generic.rs + task.rs), "
                  "1 thread as in the reference",
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


def run_reference(args) -> int:
    """--impl reference: the reference's own CPU path (restated in C, see oracle/) on this
    box's host cores.  Under torchrun only rank 0 works; the others exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from oracle import oracle
    sample = 1 << 27  # 512 MiB of the workload per step: DRAM-resident, ~40 ms per pass
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2**32, sample, dtype=np.uint32).view(np.float32).copy()
    a[np.isnan(a)] = 0
    for _ in range(args.warmup):
        oracle.cpu_simd_argminmax_f32(a)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = oracle.cpu_simd_argminmax_f32(a)
    dt = time.perf_counter() - t0
    assert res == oracle.argminmax(a, 0)
    gbs = a.nbytes * args.steps / dt / 1e9
    nt = os.cpu_count() or 1
    t1 = time.perf_counter()
    for _ in range(3):
        oracle.cpu_simd_argminmax_f32_mt(a, nt)
    mt_gbs = a.nbytes * 3 / (time.perf_counter() - t1) / 1e9
    sample_desc = (f"f32 x 2^27 elements (512 MiB) of the f32 1G workload per step, single thread (the "
                   f"reference crate is single-threaded), ISA {oracle.cpu_simd_isa()} chosen like src/lib.rs:427-443")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "f32 1G elems ArgMinMax (ignore-NaN); reference arm runs a 2^27-element sample per step on the host CPU",
                   "elements_per_step": sample},
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": 1, "kind": "port",
                         "isa": oracle.cpu_simd_isa(), "sample": sample_desc, "host_cpus": nt, "cpu_model": cpu_model(),
                         "all_cores_not_reference": {"value": round(mt_gbs, 2), "threads": nt}},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- main
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=N_ELEMS, help="elements per GPU (default 2^30)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import argminmax_b200 as amm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=device)
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

    def enqueue_step():
        if sharded is not None:
            sharded.launch(amm.IGNORE_NAN)
        else:
            shard.launch(amm.IGNORE_NAN)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # Consecutive steps scan a resident, unchanging array: let step k+1's streaming phase overlap
    # step k's fold/publish tail (programmatic dependent launch, amm_workspace_set_pipelined).
    ws.set_pipelined(True)

    # ---- warm-up (untimed) + correctness of the answer we are about to time
    for _ in range(args.warmup):
        enqueue_step()
    result = (sharded or shard).wait(amm.IGNORE_NAN)
    if distributed:
        # cross-check the exchange + combine: every rank's LOCAL answer (single-GPU path, parity-
        # tested against the oracle) and the values there are gathered; the global answer must be
        # the lexicographic (value, index) extreme over ranks.
        li, lj = shard.argminmax()
        mine = torch.tensor([float(data[li]), float(rank * n + li), float(data[lj]), float(rank * n + lj)],
                            dtype=torch.float64, device=device)
        allv = torch.zeros(world * 4, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allv, mine)
        rows = allv.view(world, 4).cpu().tolist()
        exp = (int(min(rows, key=lambda r: (r[0], r[1]))[1]), int(min(rows, key=lambda r: (-r[2], r[3]))[3]))
        assert result == exp, f"distributed result {result} != expected {exp}"
    launches0 = ws.last_launch()["launches_total"]

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
    assert result2 == result
    launches = ws.last_launch()["launches_total"] - launches0
    if distributed:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * nbytes / ms_per_step / 1e6  # GB/s, aggregate
    clocks = sampler.summary(t_wall0, t_wall1)
    main_launch = ws.last_launch()
    tuning = ws.get_tuning()  # geometry of the timed scans (later probes use other shapes)

    # ---- kernel-only duration of the dominant kernel (roofline), this rank, same stream,
    #      launches strictly serialised (no overlap between consecutive kernels)
    ws.set_pipelined(False)
    for _ in range(2):
        shard.launch(amm.IGNORE_NAN)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        shard.launch(amm.IGNORE_NAN)
    k1.record()
    k1.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps
    peak, peak_src = measured_peak()
    achieved = nbytes / kernel_ms / 1e6
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": committed_traffic(),
                "kernel": "argminmax_scan_kernel<PolF32<ignore>>", "kernel_ms": round(kernel_ms, 5),
                "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                "note": "kernel_ms: serialised launches (each kernel alone on the GPU); `value` is measured with "
                        "consecutive steps pipelined (PDL), which hides each kernel's fold/publish tail",
                "frac_of_nominal_8TBps": round(achieved / 8000.0, 4)}

    # the clock sampler belongs to the timed regions above; its NVML polling thread would
    # contend with the launch path of the latency probe below
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    # ---- small-n latency: us per call at 1M elements, synchronous C-ABI call timed inside the
    #      library (amm_measure_call_latency: host wall clock, no Python in the loop)
    import ctypes
    # A B200 that has just streamed at full HBM rate answers small calls ~1.7 us slower for a few
    # hundred ms (tools/exp/latency_after_load.py: 12.8 us right after 400 x 4 GiB scans, 11.0 us
    # fresh and again 1 s later).  The probe measures the settled state a latency-bound caller sees.
    time.sleep(1.0)
    lat3 = (ctypes.c_double * 3)()
    rc = amm.lib().amm_measure_call_latency(amm.DTYPES[DTYPE], data.data_ptr(), 1 << 20, amm.IGNORE_NAN,
                                            ws.handle, stream, 2000, lat3)
    if rc:
        raise RuntimeError(f"amm_measure_call_latency rc={rc}")
    us_1m = lat3[0]
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
    host = None
    if not args.no_e2e:
        host = torch.empty(n, dtype=torch.float32, pin_memory=True)
        host.copy_(data)
        torch.cuda.synchronize()
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
        assert r_e2e == result, (r_e2e, result)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if distributed:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        chunks = -(-nbytes // (32 << 20))
        e2e = {"value": round(world * nbytes * e2e_steps / dt / 1e9, 2), "unit": "GB/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 64 * chunks,
               "steps": e2e_steps, "ms_per_step": round(dt / e2e_steps * 1e3, 3),
               "api": "amm_host_argminmax_record (HostSlice): pinned host f32 -> chunked H2D overlapped with the scan; PCIe-bound"}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and not distributed and not args.no_cpu_baseline:
        if host is not None:
            sample = host.numpy()[: 1 << 28]
        else:
            sample = data[: 1 << 28].cpu().numpy()
        cpu = cpu_baseline(np.ascontiguousarray(sample), budget_s=12.0)

    sampler.stop_flag = True
    ll = main_launch
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": "configs[3]: f32 1G (2^30) elements = 4 GiB per GPU, device-resident, ArgMinMax (ignore-NaN), "
                            "uniform random bit patterns NaN->0 (reference bench generator)",
                "elements_per_gpu": n, "bytes_per_gpu": nbytes, "l2": "input (4 GiB) >> 126 MB L2, every step streams from HBM",
                "parallelism": f"shard{world}" if distributed else "single",
                "collective": (None if not distributed else
                               "fused in-kernel peer exchange of 64-byte records over NVLink (CUDA IPC), no NCCL on the step"
                               if sharded.exchange == "peer" else
                               "nccl all_gather of 64-byte records + combine kernel"),
                "tuning": tuning, "grid": ll["grid"], "block": ll["threads"],
                "host_numa_binding": numa,
            },
            "elements_per_s": round(world * n / (ms_per_step / 1e3), 1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "us_per_call_1m": round(us_1m, 2),
            "us_per_call_1m_detail": {"median": round(lat3[0], 2), "min": round(lat3[1], 2), "p99": round(lat3[2], 2),
                                      "through_python_ctypes_median": round(us_1m_python, 2),
                                      "what": "f32 n=2^20, synchronous amm_argminmax, host wall clock, GPU idle for 1 s "
                                              "before the probe (right after the streaming loops: +1.7 us)"},
            "result": list(result), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
