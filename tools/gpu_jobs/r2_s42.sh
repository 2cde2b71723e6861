#!/bin/bash
# round 2, GPU call 42 (1 GPU): ragged windows with offsets prefetched into shared memory: window tests + ragged throughput
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 240 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_bounds.py -x -q ) > $O/r2s42_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 2 $O/r2s42_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 120 python tools/windows_bench.py --ragged --dtypes f32,u8,f16,f64 --windows 16,64,129,256,500,1000,4096,20000 > $O/r2s42_windows_ragged.jsonl 2> $O/r2s42_err.log; echo "ragged rc=$?"
