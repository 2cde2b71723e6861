#!/bin/bash
# round 2, GPU call 38 (1 GPU): oversized windows handed to the whole CTA: window tests, ragged throughput with / without an outlier
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_bounds.py -x -q ) > $O/r2s38_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 $O/r2s38_pytest.log
timeout 300 python tools/windows_bench.py --ragged --dtypes f32,u8,f16,f64 --windows 16,64,129,256,500,1000,4096,20000 > $O/r2s38_windows_ragged.jsonl 2> $O/r2s38_err.log; echo "ragged rc=$?"
timeout 300 python tools/windows_bench.py --ragged --outlier 25000000 --dtypes f32,u8,f64 --windows 16,129,500,4096 > $O/r2s38_windows_ragged_outlier.jsonl 2>> $O/r2s38_err.log; echo "outlier rc=$?"
AMM_LIB_PATH=$PWD/argminmax_b200/lib/libargminmax_b200_prev.so timeout 600 python tools/windows_bench.py --ragged --outlier 25000000 --dtypes f32,u8 --windows 129,4096 --iters 2 > $O/r2s38_windows_ragged_outlier_prev.jsonl 2>> $O/r2s38_err.log; echo "outlier prev rc=$?"
tail -n 3 $O/r2s38_err.log
