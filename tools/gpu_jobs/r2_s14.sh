#!/bin/bash
# round 2, GPU call 14 (1 GPU): the driver's own sequence on the final state -- full pytest -m gpu, smoke, bench (both arms) --
# plus the tables of RESULTS.md
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > $O/r2s14_pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/r2s14_pytest_full.log
tail -n 8 $O/r2s14_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s14_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2s14_bench_ref.json 2> $O/r2s14_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r2s14_bench_n1.json 2> $O/r2s14_bench_n1.err; echo "bench rc=$?"
python bench.py > $O/r2s14_bench_n1_default.json 2> $O/r2s14_bench_n1_default.err; echo "bench default rc=$?"
python tools/results_table.py > $O/r2s14_results_table.jsonl 2> $O/r2s14_err.log
make -C tests/cpp -B > /dev/null 2>&1; tests/cpp/bin/device_slice_demo latency > $O/r2s14_latency_c_f32.jsonl 2>> $O/r2s14_err.log
python tools/latency_c.py --calls 1000 --max-log2 22 --modes 0,1 > $O/r2s14_latency_all_dtypes.jsonl 2>> $O/r2s14_err.log
python tools/windows_bench.py --dtypes f32,u8,i16,f16,f64 --windows 16,64,256,500,1000,4096,40000,250000,1000000,5000000,10000000,100000000 > $O/r2s14_windows_bench.jsonl 2>> $O/r2s14_err.log
python tools/exp/host_pageable.py > $O/r2s14_host_pageable.jsonl 2>> $O/r2s14_err.log
tail -n 5 $O/r2s14_err.log
