#!/bin/bash
# round 2, GPU call 43 (1 GPU): ragged offsets prefetch, fixed: window tests + ragged throughput
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 150 python -m pytest tests/test_gpu_windows.py -x -q ) > $O/r2s43_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 2 $O/r2s43_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 60 python tools/windows_bench.py --ragged --dtypes f32,u8,f16,f64 --windows 16,64,129,256,500,1000,4096 --iters 5 > $O/r2s43_windows_ragged.jsonl 2> $O/r2s43_err.log; echo "ragged rc=$?"
