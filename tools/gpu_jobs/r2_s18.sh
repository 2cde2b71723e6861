#!/bin/bash
# round 2, GPU call 18 (1 GPU): ring kernel with 32-byte vectors for 8-byte dtypes: parity (all paths) + A/B at 4 GiB
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bounds.py tests/test_gpu_api.py -m gpu -x -q > $O/r2s18_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s18_pytest.log
tail -n 3 $O/r2s18_pytest.log
for dt in i64 u64 f64; do
  for mode in 0 1; do
    [ $mode = 1 ] && [ $dt != f64 ] && continue
    for ring in 1 0; do
      AMM_SCAN_RING=$ring timeout 300 python tools/ncu_case.py --dtype $dt --n 536870912 --mode $mode --iters 10 >> $O/r2s18_ab.jsonl 2>> $O/r2s18_err.log
    done
  done
done
AMM_SCAN_RING=1 timeout 300 python tools/ncu_case.py --dtype f64 --n 8589934592 --iters 5 >> $O/r2s18_ab.jsonl 2>> $O/r2s18_err.log
AMM_SCAN_RING=0 timeout 300 python tools/ncu_case.py --dtype f64 --n 8589934592 --iters 5 >> $O/r2s18_ab.jsonl 2>> $O/r2s18_err.log
cat $O/r2s18_ab.jsonl | cut -c1-140; tail -n 3 $O/r2s18_err.log
