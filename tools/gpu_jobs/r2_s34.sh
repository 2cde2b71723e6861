#!/bin/bash
# round 2, GPU call 34 (1 GPU): compact latency kernel (one locate path with selects, redux-based CTA reduce,
# tagged-slot grid step): full GPU test-suite + in-process A/B against the previous commit's library
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > $O/r2s34_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -n 6 $O/r2s34_pytest_full.log
timeout 600 python tools/exp/latency_ab.py --mode 0 > $O/r2s34_latency_ab_inproc.jsonl 2> $O/r2s34_err.log; echo "ab rc=$?"
timeout 600 python tools/exp/latency_ab.py --mode 1 --dtypes f32,f16,f64 >> $O/r2s34_latency_ab_inproc.jsonl 2>> $O/r2s34_err.log; echo "ab rc=$?"
tail -n 3 $O/r2s34_err.log
