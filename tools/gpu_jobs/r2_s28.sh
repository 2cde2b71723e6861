#!/bin/bash
# round 2, GPU call 28 (1 GPU): latency kernel with winner-lane locate + packed words on their own
# lines: parity, then A/B against the previous commit's library on the same box
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_bounds.py -x -q ) > $O/r2s28_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 $O/r2s28_pytest.log
PREV=$PWD/argminmax_b200/lib/libargminmax_b200_prev.so
: > $O/r2s28_latency_ab.jsonl
for rep in 1 2 3; do
  AMM_LIB_PATH=$PREV timeout 200 python tools/latency_c.py --dtypes f32,u8,f16 --modes 0,1 --min-log2 12 --max-log2 21 --calls 1000 --tag prev >> $O/r2s28_latency_ab.jsonl 2>> $O/r2s28_err.log
  timeout 200 python tools/latency_c.py --dtypes f32,u8,f16 --modes 0,1 --min-log2 12 --max-log2 21 --calls 1000 --tag new_spread >> $O/r2s28_latency_ab.jsonl 2>> $O/r2s28_err.log
  AMM_PACK_SPREAD=0 timeout 200 python tools/latency_c.py --dtypes f32,u8,f16 --modes 0,1 --min-log2 12 --max-log2 21 --calls 1000 --tag new_same_line >> $O/r2s28_latency_ab.jsonl 2>> $O/r2s28_err.log
done
echo "ab rc=$?"; tail -n 3 $O/r2s28_err.log
