#!/bin/bash
# round 2, GPU call 15 (1 GPU): 8-byte pairs in the staged window kernels + pipelined small-scan shape: parity, tables
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_reference_scale.py tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > $O/r2s15_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s15_pytest.log
tail -n 5 $O/r2s15_pytest.log
python tools/results_table.py > $O/r2s15_results_table.jsonl 2> $O/r2s15_err.log
python tools/windows_bench.py --dtypes f32,u8,i16,f16,f64 --windows 16,64,256,500,1000,4096,40000,250000,1000000,5000000,10000000,100000000 > $O/r2s15_windows_bench.jsonl 2>> $O/r2s15_err.log
python tools/windows_bench.py --dtypes f64,i64 --windows 17,33,64,100,129,256,500 > $O/r2s15_windows_bench_8byte.jsonl 2>> $O/r2s15_err.log
python bench.py > $O/r2s15_bench_n1_default.json 2> $O/r2s15_bench_n1_default.err; echo "bench default rc=$?"
tail -n 3 $O/r2s15_err.log
