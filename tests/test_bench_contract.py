"""bench.py on a CPU-only box: the reference arm (the reference's CPU path restated in oracle/,
timed on the host cores) must print ONE JSON line with the contract's keys; host-side helpers."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_one_contract_line():
    # the driver runs the full 2^30-element workload; the CPU suite uses a smaller --elems
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "3", "--elems", str(1 << 24)], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["steps"] == 1 and d["warmup"] == 3 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["isa"] and cb["sample"] and cb["cpu_model"]
    assert "workload" in d["config"]
    # both arms print the SAME config dict (the driver compares them)
    b = _bench_module()
    assert d["config"] == b.workload_config(1 << 24, 1)
    assert d["verified_vs_oracle"] is True


def test_both_arms_share_one_config():
    b = _bench_module()
    for world in (1, 2, 8):
        c = b.workload_config(1 << 30, world)
        assert c["elements_per_gpu"] == 1 << 30 and c["bytes_per_gpu"] == 4 << 30
        assert c["parallelism"] == ("single" if world == 1 else f"shard{world}")
        assert "L2" in c["l2"]


def test_affinity_mask_and_lex_fold():
    b = _bench_module()
    assert b.cpus_from_affinity_mask([0b1011, 0b1]) == {0, 1, 3, 64}
    # rows: (has_value, min_val, min_idx, max_val, max_idx, nan_idx)
    rows = [[1, -5.0, 3, 9.0, 4, -1], [0, 0, 0, 0, 0, 17], [1, -5.0, 40, 9.0, 41, 12]]
    assert b._lex_fold(rows) == (3, 4, 12)      # lowest shard wins ties; first NaN = min over shards
    assert b._lex_fold([[0, 0, 0, 0, 0, -1]]) == (0, 0, -1)


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=120,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_parse_cpulist():
    b = _bench_module()
    assert b.parse_cpulist("0-3,8-11,16\n") == {0, 1, 2, 3, 8, 9, 10, 11, 16}
    assert b.parse_cpulist("5") == {5}
    assert b.parse_cpulist("") == set()
