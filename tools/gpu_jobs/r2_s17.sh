#!/bin/bash
# round 2, GPU call 17 (1 GPU): new f64 return-NaN policy (parity + rate), ring vs register path at 16 / 32 / 64 GiB
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_windows.py tests/test_gpu_api.py -m gpu -x -q -k "f64 or golden or known" > $O/r2s17_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s17_pytest.log
tail -n 3 $O/r2s17_pytest.log
timeout 600 python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "f64" >> $O/r2s17_pytest.log 2>&1; echo "pytest large rc=$?" >> $O/r2s17_pytest.log
tail -n 3 $O/r2s17_pytest.log
for n in 4294967296 17179869184 34359738368 68719476736; do
  for ring in 1 0; do
    AMM_SCAN_RING=$ring timeout 300 python tools/ncu_case.py --dtype u8 --n $n --iters 8 >> $O/r2s17_big_ab.jsonl 2>> $O/r2s17_err.log
  done
done
for mode in 0 1; do
  timeout 300 python tools/ncu_case.py --dtype f64 --n 536870912 --mode $mode --iters 10 >> $O/r2s17_f64.jsonl 2>> $O/r2s17_err.log
  timeout 300 python tools/ncu_case.py --dtype f64 --n 100000000 --mode $mode --iters 40 >> $O/r2s17_f64.jsonl 2>> $O/r2s17_err.log
done
cat $O/r2s17_big_ab.jsonl $O/r2s17_f64.jsonl; tail -n 3 $O/r2s17_err.log
