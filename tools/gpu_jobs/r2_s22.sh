#!/bin/bash
# round 2, GPU call 22 (1 GPU): the driver's own sequence on the final state: full pytest -m gpu, smoke, both bench arms
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > $O/r2s22_pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/r2s22_pytest_full.log
tail -n 6 $O/r2s22_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s22_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/r2s22_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2s22_bench_ref.json 2> $O/r2s22_bench_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2s22_bench_n1.json 2> $O/r2s22_bench_n1.err; echo "bench rc=$?"
python bench.py > $O/r2s22_bench_n1_default.json 2> $O/r2s22_bench_n1_default.err; echo "bench default rc=$?"
