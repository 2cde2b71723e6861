#!/bin/bash
# round 2, GPU call 25 (1 GPU): full pytest -m gpu on the final state (value-domain compares in the window kernels)
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > $O/r2s25_pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/r2s25_pytest_full.log
tail -n 6 $O/r2s25_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s25_smoke.log 2>&1; echo "smoke rc=$?"
python tools/windows_bench.py --dtypes f32,f64 --windows 17,33,100,129,500,1001,2047 > $O/r2s25_windows_odd.jsonl 2> $O/r2s25_err.log
