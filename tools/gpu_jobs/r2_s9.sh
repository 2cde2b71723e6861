#!/bin/bash
# round 2, GPU call 9 (1 GPU): parity with the TMA-ring kernel; ring vs register path on 4 GiB (bench + dtype table)
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not large" > $O/r2s9_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s9_pytest.log
tail -n 5 $O/r2s9_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2s9_bench_ring.json 2> $O/r2s9_bench_ring.err; echo "bench ring rc=$?"
AMM_SCAN_RING=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2s9_bench_noring.json 2> $O/r2s9_bench_noring.err; echo "bench noring rc=$?"
timeout 300 python tools/phase_stamps.py --tag ring --cases f32:1073741824,f32:268435456,u8:1073741824 >> $O/r2s9_stamps.jsonl 2>> $O/r2s9_err.log
AMM_SCAN_RING=0 timeout 300 python tools/phase_stamps.py --tag noring --cases f32:1073741824,f32:268435456,u8:1073741824 >> $O/r2s9_stamps.jsonl 2>> $O/r2s9_err.log
AMM_SCAN_RING=1 timeout 300 python tools/phase_stamps.py --tag ring_all --cases f32:10000000,f32:100000000,u8:100000000 >> $O/r2s9_stamps.jsonl 2>> $O/r2s9_err.log
AMM_RING_CTAS=1 timeout 300 python tools/phase_stamps.py --tag ring_1cta --cases f32:1073741824 >> $O/r2s9_stamps.jsonl 2>> $O/r2s9_err.log
tail -n 5 $O/r2s9_err.log $O/r2s9_bench_ring.err 2>/dev/null | tail -n 12
