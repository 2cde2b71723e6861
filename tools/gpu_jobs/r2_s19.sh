#!/bin/bash
# round 2, GPU call 19 (1 GPU): tiny-window passes + packed-equality locate: full parity, window table, latency
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > $O/r2s19_pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/r2s19_pytest_full.log
tail -n 8 $O/r2s19_pytest_full.log
python tools/windows_bench.py --dtypes f32,u8,i16,f16,f64 --windows 16,32,64,128,256,500,1000,4096 > $O/r2s19_windows_bench.jsonl 2> $O/r2s19_err.log
python tools/latency_c.py --dtypes f32,u8,i16,f16 --calls 2000 --min-log2 10 --max-log2 21 --tag final > $O/r2s19_lat.jsonl 2>> $O/r2s19_err.log
tail -n 3 $O/r2s19_err.log
