#!/bin/bash
# round 2, GPU call 31 (1 GPU): tagged-slot grid step + lazy locate for 64-bit keys, one CTA up to 4096 vectors:
# full GPU test-suite, then latency A/B against the previous commit's library on the same box
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > $O/r2s31_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -n 6 $O/r2s31_pytest_full.log
PREV=$PWD/argminmax_b200/lib/libargminmax_b200_prev.so
: > $O/r2s31_latency_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python tools/latency_c.py --dtypes f32,u8,i64,f64 --modes 0,1 --min-log2 10 --max-log2 23 --calls 1000 --tag $tag >> $O/r2s31_latency_ab.jsonl 2>> $O/r2s31_err.log; }
for rep in 1 2 3; do
  run prev AMM_LIB_PATH=$PREV
  run new AMM_X=0
  run new_e64_2p18 AMM_DIRECT_ELEMS64=262144
  run new_e64_2p19 AMM_DIRECT_ELEMS64=524288
done
tail -n 3 $O/r2s31_err.log
