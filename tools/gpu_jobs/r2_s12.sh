#!/bin/bash
# round 2, GPU call 12 (1 GPU): fixed-unaligned window kernel: parity + bench on/off
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_bounds.py -m gpu -x -q > $O/r2s12_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s12_pytest.log
tail -n 8 $O/r2s12_pytest.log
W="--dtypes f32,u8,i16,f16,f64 --windows 17,33,100,129,500,1000,1001,2047"
timeout 600 python tools/windows_bench.py $W > $O/r2s12_windows_fixed1.jsonl 2> $O/r2s12_err.log
AMM_WIN_FIXED=0 timeout 600 python tools/windows_bench.py $W > $O/r2s12_windows_fixed0.jsonl 2>> $O/r2s12_err.log
tail -n 3 $O/r2s12_err.log
