#!/bin/bash
# round 2, GPU call 33 (1 GPU): in-process A/B (previous commit's library vs working tree) of the call latency
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_bounds.py -x -q ) > $O/r2s33_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 2 $O/r2s33_pytest.log
timeout 600 python tools/exp/latency_ab.py --mode 0 > $O/r2s33_latency_ab_inproc.jsonl 2> $O/r2s33_err.log; echo "ab rc=$?"
timeout 600 python tools/exp/latency_ab.py --mode 1 --dtypes f32,f16,f64 >> $O/r2s33_latency_ab_inproc.jsonl 2>> $O/r2s33_err.log; echo "ab rc=$?"
tail -n 3 $O/r2s33_err.log
