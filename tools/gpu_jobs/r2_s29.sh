#!/bin/bash
# round 2, GPU call 29 (1 GPU): geometry of the latency kernel after the winner-lane locate
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bounds.py -x -q ) > $O/r2s29_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 2 $O/r2s29_pytest.log
: > $O/r2s29_latency_geom.jsonl
run() { tag=$1; shift; env "$@" timeout 200 python tools/latency_c.py --dtypes f32,u8 --modes 0 --min-log2 12 --max-log2 22 --calls 1000 --tag $tag >> $O/r2s29_latency_geom.jsonl 2>> $O/r2s29_err.log; }
for rep in 1 2 3; do
  run t256_c2 AMM_X=0
  run t512_c2 AMM_DIRECT_THREADS=512
  run t256_c4_2m AMM_DIRECT_CTAS=4 AMM_DIRECT_ELEMS32=2097152
  run t512_c4_2m AMM_DIRECT_THREADS=512 AMM_DIRECT_CTAS=4 AMM_DIRECT_ELEMS32=2097152
  run t128_c4 AMM_DIRECT_THREADS=128 AMM_DIRECT_CTAS=4
done
tail -n 3 $O/r2s29_err.log
