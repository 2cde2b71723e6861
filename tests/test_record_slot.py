"""The integrity word of the 64-byte record slots (host slot polled by amm_workspace_wait, peer
slots of the fused exchange) is position-sensitive: a slot torn between two records is never
accepted as the new one (a plain XOR accepted old (1,0) / new (0,1) mixtures).
Compiled with plain g++ from the header the kernels use; needs no GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mix_checksum_rejects_torn_slots(tmp_path):
    exe = tmp_path / "record_slot_check"
    src = os.path.join(ROOT, "tests", "cpp", "record_slot_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), src], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "record_slot_check ok" in out.stdout
