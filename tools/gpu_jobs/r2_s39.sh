#!/bin/bash
# round 2, GPU call 39 (1 GPU): final state.  Risky part first under a short limit (oversized windows handed
# to the CTA), then the driver's own sequence: full pytest -m gpu, smoke, both bench arms; latency + results tables
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 150 python -m pytest tests/test_gpu_windows.py -x -q -k "outlier" ) > $O/r2s39_pytest_outliers.log 2>&1; rc=$?; echo "outlier tests rc=$rc"; tail -n 3 $O/r2s39_pytest_outliers.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 60 python tools/windows_bench.py --ragged --outlier 25000000 --dtypes f32,u8,f64 --windows 16,129,500,4096 --iters 3 > $O/r2s39_windows_ragged_outlier.jsonl 2> $O/r2s39_err.log; echo "outlier bench rc=$?"
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r2s39_pytest_full.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 5 $O/r2s39_pytest_full.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s39_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/r2s39_smoke.log
timeout 400 python bench.py > $O/r2s39_bench_n1.json 2> $O/r2s39_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > $O/r2s39_bench_ref.json 2> $O/r2s39_bench_ref.err; echo "ref rc=$?"
make -C tests/cpp -B > /dev/null 2>&1; timeout 120 tests/cpp/bin/device_slice_demo latency > $O/r2s39_latency_c_f32.jsonl 2>> $O/r2s39_err.log; echo "latency rc=$?"
timeout 200 python tools/latency_c.py --dtypes f32,u8,i16,f16,i64,f64 --modes 0,1 --max-log2 22 > $O/r2s39_latency_all_dtypes.jsonl 2>> $O/r2s39_err.log; echo "latency all rc=$?"
timeout 300 python tools/results_table.py > $O/r2s39_results_table.jsonl 2>> $O/r2s39_err.log; echo "table rc=$?"
tail -n 3 $O/r2s39_err.log
