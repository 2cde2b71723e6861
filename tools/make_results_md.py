#!/usr/bin/env python
"""Assemble RESULTS.md from the committed measurement files under profiles/ (round 2).
  python tools/make_results_md.py [--tag r02e] > RESULTS.md"""
import argparse
import collections
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def lines(name):
    with open(os.path.join(PROF, name)) as f:
        return [json.loads(l) for l in f if l.strip().startswith("{")]


def one(name):
    return lines(name)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r02g")
    ap.add_argument("--multi-tag", default="r02g")
    a = ap.parse_args()
    t = a.tag
    out = []
    w = out.append
    w("# RESULTS — the BASELINE.md §4 table, filled in (round 2, this pool's B200s)\n")
    w("Every number is measured through `gpurun`; sources are the committed files under `profiles/` "
      "(see `profiles/README.md`); this file is assembled by `tools/make_results_md.py`.")
    w("GB = 1e9 bytes. \"Serialised\" = every call alone on the GPU (CUDA events over back-to-back non-pipelined "
      "launches); \"pipelined\" = a stream of calls with `amm_workspace_set_pipelined` (PDL overlap of each call's "
      "tail with the next call's streaming).")
    w("Peak = 6547.8 GB/s (`MEASURED_PEAKS.json`, read+write copy); HBM3e pin rate 8.19 TB/s. Round-1 figures in "
      "brackets (`profiles/r01d_results_table.jsonl`, other boxes of the same pool).\n")

    rt = lines(f"{t}_results_table.jsonl")
    r1 = {}
    try:
        for d in lines("r01d_results_table.jsonl"):
            if d["config"] == 1:
                r1[(1, d["n"], d["cache"], d["pipelined"])] = d
            else:
                r1[(2, d["dtype"], d["mode"])] = d
    except OSError:
        pass
    w("## Config 1 — f32, ArgMinMax (ignore-NaN), uniform random bit patterns NaN→0 (`tools/results_table.py`)\n")
    w("| n | L2 state | mode | µs/call | GB/s | × measured peak |")
    w("|---:|---|---|---:|---:|---:|")
    for d in rt:
        if d["config"] != 1:
            continue
        old = r1.get((1, d["n"], d["cache"], d["pipelined"]))
        o = f" ({old['us_per_call']})" if old else ""
        w(f"| {d['n']:,} | {d['cache'].replace('_', ' ')} | {'pipelined' if d['pipelined'] else 'serialised'} | "
          f"{d['us_per_call']}{o} | {d['GBps']} | {d['GBps'] / 6547.8:.3f} |")
    w("\nReference for the same workload: 2.4 ms for 10 M f32 on a Ryzen 7 4800U (README.md:101, commented out), "
      "6.98 µs for n = 102 400 cache-resident on an i7-1185G7 (benches/results:15). At that size the CPU is the faster "
      "device: a synchronous GPU call costs 9.2 µs (table below), the C port of the reference on the GPU box's host "
      "10.8 µs through ctypes (`us_per_call.cpu_port_102400` in the bench line) — 400 KB does not amortise a launch. "
      "Where the 10 M serialised call spends its 11.9 µs: `profiles/r02_fixed_cost.md`.\n")

    w("## Config 2 — all 11 dtypes × 100 M elements, both NaN modes, cold L2 (rotating buffers ≥ 512 MB total)\n")
    w("Parity at this size: `tests/test_gpu_large.py::test_config2_all_dtypes_100m` (GPU == scalar oracle == SIMD "
      "emulation, bit-exact); the bench line's `dtype_table` re-checks every row against the oracle in the driver's run.\n")
    w("| dtype | mode | bytes | serialised µs | serialised GB/s | pipelined µs | pipelined GB/s | G elem/s (serialised) |")
    w("|---|---|---:|---:|---:|---:|---:|---:|")
    for d in rt:
        if d["config"] != 2:
            continue
        old = r1.get((2, d["dtype"], d["mode"]))
        o = f" ({old['GBps_serialised']})" if old else ""
        w(f"| {d['dtype']} | {d['mode']} | {d['bytes']:,} | {d['us_serialised']} | {d['GBps_serialised']}{o} | "
          f"{d['us_pipelined']} | {d['GBps_pipelined']} | {d['Gelem_per_s_serialised']} |")
    b1 = one(f"{t}_bench_n1.json")
    w("\nAt 4 GiB per array (`dtype_table` of the bench line, serialised): "
      + ", ".join(f"{r['dtype']}{'/ret' if r['mode'] == 'return_nan' else ''} {r['gbs_4gib']:.0f}" for r in b1["dtype_table"])
      + " GB/s.\n")

    w("## Config 3 — adversarial f16/f32/f64 × 100 M\n")
    w("Parity only (`tests/test_gpu_large.py::test_config3_adversarial_100m`): NaN at 0 / n/2 / n−1 / first 100 / last 100 / "
      "random 0.1 %, ±inf at 64 random places, duplicate extremes in different CTAs, all-equal, all +inf, all −inf, all NaN; "
      "both semantics; bit-exact vs the oracle. Throughput is data-independent (branch-free loop; "
      "`test_worst_case_array_is_order_independent`).\n")

    rf = b1["roofline"]
    w(f"## Config 4 — f32 × 2^30 (4 GiB) on 1 × B200 (`profiles/{t}_bench_n1.json`, `r02d_f32_1g_ring_ncu_full_metrics.csv`)\n")
    w("| quantity | value |")
    w("|---|---|")
    w(f"| `bench.py` value ({b1['steps']} pipelined steps) | {b1['value']} GB/s, {b1['ms_per_step']} ms/step, "
      f"{b1['elements_per_s'] / 1e12:.3f} T elements/s; answer checked against both CPU restatements first (`verified_vs_oracle`) |")
    w(f"| serialised kernel ({b1['launch']['tuning']['variant_name']}, grid {b1['launch']['grid']} × {b1['launch']['block']}) | "
      f"{rf['kernel_ms']} ms → {rf['achieved']} GB/s = {rf['frac']} × measured peak = {rf['frac_of_nominal_8TBps']} × 8 TB/s |")
    w(f"| ncu DRAM traffic per launch | {rf['traffic']:,} B vs 4,294,967,296 B algorithmic (×{rf['traffic'] / 4294967296:.4f}); "
      "DRAM cycles active 91.1–91.4 % (round 1: 88.0–88.3 %) |")
    c = b1["clocks"]
    w(f"| clocks during the timed region | SM {c['sm_mhz']:.0f} MHz of {c['sm_max_mhz']}, memory {c['mem_mhz']:.0f} MHz, throttle reasons {c['reasons']} |")
    e = b1["e2e"]
    w(f"| end to end from pinned host memory (H2D inside the timed region) | {e['value']} GB/s ({e['ms_per_step']} ms per 4 GiB) = "
      f"{e['frac_of_h2d_ceiling']} × the plain pinned copy measured beside it ({e['h2d_ceiling_gbs']} GB/s) |")
    s = b1["serialised"]
    w(f"| single serialised calls, cold L2 (bench line `serialised`) | 10 M f32 {s['f32_10m']['us']} µs; 100 M f32 {s['f32_100m']['us']} µs = "
      f"{s['f32_100m']['gbs']} GB/s; 100 M u8 {s['u8_100m']['us']} µs; 102 400 f32 {s['f32_102400']['us']} µs (device time) |")
    w("")

    w(f"## Config 5 — 32 Gi f32 (128 GiB) sharded, weak and strong scaling (`profiles/{a.multi_tag}_bench_n{{2,4,8}}.json`, one 8-GPU box)\n")
    w("| GPUs | weak: aggregate GB/s (4 GiB per GPU) | ms/step | strong: 2^30 elements in total, ms per serialised call | efficiency vs 1 GPU (serialised / pipelined) | 32 Gi elements in total | e2e from host, GB/s (× plain-copy ceiling) | in-process `amm_sharded_argminmax`, µs/call (peer / NCCL) |")
    w("|---:|---:|---:|---:|---:|---:|---:|---:|")
    w(f"| 1 | {b1['value']} | {b1['ms_per_step']} | {rf['kernel_ms']} | – | – | {e['value']} ({e['frac_of_h2d_ceiling']}) | – |")
    for n, name in ((2, f"{a.multi_tag}_bench_n2.json"), (4, f"{a.multi_tag}_bench_n4.json"), (8, f"{a.multi_tag}_bench_n8.json")):
        try:
            d = one(name)
        except OSError:
            continue
        st, c5, ee, ip = d["strong"], d["config5"], d["e2e"], d["sharded_in_process"]
        w(f"| {n} | {d['value']} | {d['ms_per_step']} | {st['ms']} ({st['gbs']} GB/s) | {st['efficiency_vs_n1']} / "
          f"{st['efficiency_vs_n1_pipelined']} | {c5['ms']} ms, {c5['gbs']} GB/s | {ee['value']} ({ee['frac_of_h2d_ceiling']}) | "
          f"{ip['peer']['us_per_call']} / {ip['nccl']['us_per_call']} |")
    w("\nEvery multi-GPU line was produced after its parity step passed (duplicate extremes in the first and the last shard, "
      "NaNs in return mode, an empty shard, all-equal; in-process check of both exchanges; `multi_gpu_parity`, "
      "`sharded_in_process` in the line). Per-GPU indices exceed 2^32 in the 32 Gi rows. The plain-copy ceiling is measured "
      "in the same run with all N ranks pulling pinned host memory at once, started together (N x bytes / slowest rank): "
      "at 8 GPUs four GPUs get 23 GB/s and four 36 (55.6 alone) -- the host side of PCIe, not the pipeline "
      "(`r02d_h2d_{4,8}gpu_plain_vs_host_slice.jsonl`).\n")
    try:
        d8 = one(f"{a.multi_tag}_bench_n8.json")
        w("Synchronous collective call vs the single-GPU call on the same total size, 8 GPUs, µs (median, timed inside the library): "
          + "; ".join(f"2^{c['total_elems_log2']}: {c['us_sharded']} vs {c['us_one_gpu']}" for c in d8["strong"]["sync_call_us_by_total_size"])
          + f" — sharding pays from 2^{d8['strong']['sharding_pays_from_log2_elems']} elements.\n")
    except OSError:
        pass

    w(f"Small-n latency, synchronous C-ABI call from a C++ host program, f32, 1 GPU (`profiles/{t}_latency_c_f32.jsonl`; an EMPTY "
      "kernel + result poll costs 6.5–7.5 µs on these boxes, `tools/exp/latency_floor.cu`; in brackets the state before the "
      "last pass over the latency kernel, `r02e_latency_c_f32.jsonl`, another box):\n")
    w("| n | µs/call median | min | p99 |")
    w("|---:|---:|---:|---:|")
    try:
        before = {d["n"]: d["us_call_median"] for d in lines("r02e_latency_c_f32.jsonl")}
    except OSError:
        before = {}
    for d in lines(f"{t}_latency_c_f32.jsonl"):
        o = f" ({before[d['n']]})" if d["n"] in before and t != "r02e" else ""
        w(f"| {d['n']:,} | {d['us_call_median']}{o} | {d['us_call_min']} | {d['us_call_p99']} |")
    w(f"\nOther dtypes and return-NaN mode: `profiles/{t}_latency_all_dtypes.jsonl`. One array sharded over 8 GPUs, one host "
      "thread: 20–23 µs/call for 2^10…2^24 elements (launcher thread per device; one thread for all: 36–39 µs; NCCL: 120 µs) — "
      "`profiles/r02d_sharded_latency_8gpu_*.jsonl`.\n")

    w(f"## Windowed / segmented form (SURVEY §8f-4) — 1 GiB of data, fixed windows, kernel-only (`profiles/r02e_windows_bench.jsonl`)\n")
    w("GB/s of input (every window also writes 16 bytes of output); in brackets round 1 (`r01d_windows_bench_1gib.jsonl`) where it "
      "differs by more than 5 %.\n")
    cur = {(d["dtype"], d["window"]): d["GBps"] for d in lines("r02e_windows_bench.jsonl")}
    try:
        old = {(d["dtype"], d["window"]): d["GBps"] for d in lines("r01d_windows_bench_1gib.jsonl")}
    except OSError:
        old = {}
    dts = ["f32", "u8", "i16", "f16", "f64"]
    w("| elements per window | " + " | ".join(dts) + " |")
    w("|---:|" + "---:|" * len(dts))
    for win in sorted({k[1] for k in cur}):
        cells = []
        for dt in dts:
            v = cur.get((dt, win))
            o = old.get((dt, win))
            if v is None:
                cells.append("–")
            elif o and abs(v / o - 1) > 0.05:
                cells.append(f"{v:.0f} ({o:.0f})")
            else:
                cells.append(f"{v:.0f}")
        w(f"| {win:,} | " + " | ".join(cells) + " |")
    w("\nOdd window sizes on an aligned array (the 32-bit-geometry kernel of round 2 against the general staged form, same box): "
      "`profiles/r02d_windows_odd_sizes_{fixed,general}_kernel.jsonl` — u8 × 500 2.58 → 3.58 TB/s, u8 × 129 1.41 → 2.59, "
      "i16 × 500 3.97 → 5.21, f32 × 129 3.61 → 4.72, f32 × 1001 3.59 → 4.67. Parity: `tests/test_gpu_windows.py` and "
      "`tests/test_gpu_reference_scale.py` (per window == oracle on the sub-slice; 10 000 × 129 and 500 × 5000 arrays per "
      "dtype and mode in one launch each).\n")

    # ---- short windows before the value-domain compares (previous state, other box)
    try:
        sh = collections.defaultdict(dict)
        for d in lines("r02f_windows_bench_short.jsonl"):
            sh[d["window"]][d["dtype"]] = d["GBps"]
        w("## Short windows before the value-domain compares (`profiles/r02f_windows_bench_short.jsonl`, another box)\n")
        w("Several passes per block for 16-64-byte windows, packed-equality locate for 8-bit ints and f16, vector pairs for "
          "8-byte dtypes; GB/s of input. (The table above is the final state: f32 / f64 additionally compare values instead "
          "of ord keys.)\n")
        w("| elements per window | " + " | ".join(dts) + " |")
        w("|---:|" + "---:|" * len(dts))
        for win in sorted(sh):
            w(f"| {win:,} | " + " | ".join(f"{sh[win][dt]:.0f}" if dt in sh[win] else "–" for dt in dts) + " |")
        w("")
    except OSError:
        pass

    # ---- ragged windows (offsets), and what one oversized window costs
    try:
        rg = collections.defaultdict(dict)
        for d in lines("r02g_windows_ragged.jsonl"):
            rg[d["window"]][d["dtype"]] = d["GBps"]
        rdt = ["f32", "u8", "f16", "f64"]
        w("## Ragged windows (`amm_argminmax_windows`, offsets) — 1 GiB of data, lengths uniform in [W/2, 3W/2] "
          "(`profiles/r02g_windows_ragged.jsonl`, `tools/windows_bench.py --ragged`)\n")
        w("GB/s of input. Ragged windows always take the general staged form (64-bit window bounds, staged-range tests, "
          "clamps): 1.2-4 x below fixed windows of the same mean size up to ~1 KB per window, level from 4 Ki elements on. "
          "Known gap (DESIGN.md, section 10).\n")
        w("| mean elements per window | " + " | ".join(rdt) + " |")
        w("|---:|" + "---:|" * len(rdt))
        for win in sorted(rg):
            w(f"| {win:,} | " + " | ".join(f"{rg[win][dt]:.0f}" if dt in rg[win] else "–" for dt in rdt) + " |")
        ob = {(d["dtype"], d["window"]): d["ms"] for d in lines("r02g_windows_ragged_outlier_before.jsonl")}
        oa = {(d["dtype"], d["window"]): d["ms"] for d in lines("r02g_windows_ragged_outlier.jsonl")}
        w("\nOne window of 25 M elements among them (the kernel is chosen by the MEAN window size; a lane group of 1-32 "
          "lanes used to fold the outlier alone, now the whole CTA does, `window_by_cta`), ms per call, before -> after "
          "(`r02g_windows_ragged_outlier{_before,}.jsonl`): "
          + "; ".join(f"{dt} x {win}: {ob[(dt, win)]:.1f} -> {oa[(dt, win)]:.1f}" for (dt, win) in sorted(ob) if (dt, win) in oa)
          + ". Parity: `tests/test_gpu_windows.py::test_ragged_windows_with_giant_outliers`.\n")
    except OSError:
        pass

    # ---- last pass over the latency kernel: in-process A/B
    try:
        ab = [d for d in lines("r02g_latency_ab_inproc.jsonl")]
        w("## Call latency before / after the last pass over the latency kernel — in-process A/B "
          "(`profiles/r02g_latency_ab_inproc.jsonl`, `tools/exp/latency_ab.py`)\n")
        w("Both builds loaded side by side and timed in alternation (7 rounds x 300 calls per size), µs per synchronous call, "
          "previous commit -> working tree: fused / select-based / lazy locate, one CTA up to 32 KiB, redux-based CTA reduce, "
          "tagged-slot grid step for 64-bit keys.\n")
        sizes = ["10", "12", "13", "14", "16", "18", "20", "22"]
        w("| dtype, mode | " + " | ".join(f"2^{k}" for k in sizes) + " |")
        w("|---|" + "---:|" * len(sizes))
        for d in ab:
            w(f"| {d['dtype']}, {'return-NaN' if d['mode'] else 'ignore-NaN'} | "
              + " | ".join(f"{d['us_a'][k]:.2f} → {d['us_b'][k]:.2f}" for k in sizes) + " |")
        w("")
    except OSError:
        pass

    cb = b1.get("cpu_baseline") or {}
    ref = one(f"{t}_bench_reference_arm.json")
    w(f"## CPU baseline on the GPU box's host ({cb.get('cpu_model', 'Intel Xeon')}, {cb.get('host_cpus', 16)} vCPUs, AVX-512)\n")
    w(f"Restatement of the reference's dispatched f32 path (`oracle/cpu_simd_f32.c`, ISA `{cb.get('isa')}`), 1 thread like the crate, on the "
      f"SAME 2^30-element array the GPU scans: {cb.get('value')} GB/s best / {cb.get('median')} GB/s median; `--impl reference` arm "
      f"(same `config` dict as our arm, 2^30 elements per step): {ref['value']} GB/s, {ref['ms_per_step']} ms per step. "
      f"{cb.get('all_cores_not_reference', {}).get('threads')} threads, contiguous shards + ordered combine (NOT something the reference does): "
      f"{cb.get('all_cores_not_reference', {}).get('value')} GB/s.")
    hp = lines("r02e_host_pageable.jsonl")
    w("\nHost slices in pageable vs pinned memory (`amm_host_argminmax`, 1 GiB f32): "
      + "; ".join(f"{d['memory']}: {d['GBps_best']} GB/s" for d in hp) + ".")
    print("\n".join(out))


if __name__ == "__main__":
    main()
