#!/bin/bash
# round 2, GPU call 32 (1 GPU): single-CTA rule of the latency kernel (size limit x preferred depth), A/B with
# the previous commit's library, 5 interleaved repetitions
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_bounds.py -x -q ) > $O/r2s32_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 2 $O/r2s32_pytest.log
PREV=$PWD/argminmax_b200/lib/libargminmax_b200_prev.so
: > $O/r2s32_latency_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python tools/latency_c.py --dtypes f32,u8,f16,i64 --modes 0 --min-log2 10 --max-log2 18 --calls 1000 --tag $tag >> $O/r2s32_latency_ab.jsonl 2>> $O/r2s32_err.log; }
for rep in 1 2 3 4 5; do
  run prev AMM_LIB_PATH=$PREV
  run v2048_d2 AMM_X=0
  run v2048_d1 AMM_DIRECT_ONE_CTA_VPT=1
  run v2048_d4 AMM_DIRECT_ONE_CTA_VPT=4
  run v1024_d1 AMM_DIRECT_ONE_CTA_VECS=1024 AMM_DIRECT_ONE_CTA_VPT=1
  run v1024_d4 AMM_DIRECT_ONE_CTA_VECS=1024 AMM_DIRECT_ONE_CTA_VPT=4
  run v4096_d4 AMM_DIRECT_ONE_CTA_VECS=4096 AMM_DIRECT_ONE_CTA_VPT=4
done
tail -n 3 $O/r2s32_err.log
