#!/bin/bash
# round 2, GPU call 44 (1 GPU): parity after the launch shape of the latency kernel moved into plan_direct
cd $GRAFT_REPO_ROOT
( timeout 60 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2s44_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2s44_pytest.log
