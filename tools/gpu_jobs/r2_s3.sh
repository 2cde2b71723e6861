#!/bin/bash
# round 2, GPU call 3: parity with the dynamic tail; stamps per dyn pct / geometry; results table
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not large" > $O/r2s3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s3_pytest.log
tail -3 $O/r2s3_pytest.log
python tools/phase_stamps.py --tag auto >> $O/r2s3_stamps.jsonl 2>> $O/r2s3_err.log
for pct in 0 10 20 50 90; do
  AMM_SCAN_DYN_PCT=$pct python tools/phase_stamps.py --tag dyn$pct >> $O/r2s3_stamps.jsonl 2>> $O/r2s3_err.log
done
for cfg in "512 2 1" "256 4 0" "128 8 1"; do
  set -- $cfg
  python tools/phase_stamps.py --tag "t$1_c$2_v$3" --threads $1 --ctas $2 --variant $3 --cases f32:10000000,f32:100000000,u8:100000000 >> $O/r2s3_stamps.jsonl 2>> $O/r2s3_err.log
done
python tools/results_table.py > $O/r2s3_results_table.jsonl 2>> $O/r2s3_err.log
AMM_SCAN_DYN_PCT=0 python tools/results_table.py > $O/r2s3_results_table_static.jsonl 2>> $O/r2s3_err.log
python bench.py --steps 20 --warmup 5 > $O/r2s3_bench_n1.json 2> $O/r2s3_bench_n1.err; echo "bench rc=$?"
tail -n 5 $O/r2s3_err.log; tail -n 5 $O/r2s3_bench_n1.err
