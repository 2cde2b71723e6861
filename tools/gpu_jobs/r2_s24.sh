#!/bin/bash
# round 2, GPU call 24 (1 GPU): value-domain compares in every window kernel for f64: parity + table
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_bounds.py -m gpu -x -q > $O/r2s24_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s24_pytest.log
tail -n 3 $O/r2s24_pytest.log
python tools/windows_bench.py --dtypes f32,u8,i16,f16,f64 --windows 16,64,256,500,1000,4096,40000,250000,1000000,5000000,10000000,100000000 > $O/r2s24_windows_bench.jsonl 2> $O/r2s24_err.log
tail -n 3 $O/r2s24_err.log
