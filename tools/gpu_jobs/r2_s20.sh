#!/bin/bash
# round 2, GPU call 20 (1 GPU): packed-equality locate kept for 8-bit ints and f16 only: parity + window table
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2s20_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s20_pytest.log
tail -n 3 $O/r2s20_pytest.log
python tools/windows_bench.py --dtypes f32,u8,i16,f16,f64 --windows 16,32,64,128,256,500,1000,4096 > $O/r2s20_windows_bench.jsonl 2> $O/r2s20_err.log
AMM_WIN_FAST=0 python tools/windows_bench.py --dtypes u8,i16,f16 --windows 16,32,64 > $O/r2s20_windows_bench_nofast.jsonl 2>> $O/r2s20_err.log
tail -n 3 $O/r2s20_err.log
