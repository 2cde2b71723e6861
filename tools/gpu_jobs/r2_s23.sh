#!/bin/bash
# round 2, GPU call 23 (1 GPU): f64 windows compare values instead of keys: parity + table
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_bounds.py -m gpu -x -q > $O/r2s23_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s23_pytest.log
tail -n 3 $O/r2s23_pytest.log
python tools/windows_bench.py --dtypes f64,i64,f32 --windows 16,32,64,100,128,129,256,500,1000,4096 > $O/r2s23_windows_bench.jsonl 2> $O/r2s23_err.log
tail -n 3 $O/r2s23_err.log
