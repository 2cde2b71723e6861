#!/bin/bash
# round 2, GPU call 7 (1 GPU): parity after the counter ring (dynamic tail in pipelined mode), bench, results table,
# latency-kernel geometry sweep, PDL attribute cost on the synchronous call
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not large" > $O/r2s7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s7_pytest.log
tail -n 3 $O/r2s7_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2s7_bench_n1.json 2> $O/r2s7_bench_n1.err; echo "bench rc=$?"
AMM_SCAN_DYN_PIPELINED=0 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras > $O/r2s7_bench_n1_static_pipelined.json 2> $O/r2s7_bench_n1_b.err; echo "bench rc=$?"
python tools/results_table.py > $O/r2s7_results_table.jsonl 2>> $O/r2s7_err.log
L="python tools/latency_c.py --dtypes f32,u8 --calls 2000 --min-log2 12 --max-log2 21"
$L --tag default > $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
AMM_PDL=0 $L --tag nopdl >> $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
for cfg in "256 4" "512 2" "128 8" "512 1" "256 1" "128 4"; do
  set -- $cfg
  AMM_DIRECT_THREADS=$1 AMM_DIRECT_CTAS=$2 AMM_DIRECT_ELEMS32=4194304 $L --tag "t$1_c$2" >> $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
done
AMM_DIRECT_VPT=2 $L --tag vpt2 >> $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
AMM_DIRECT_VPT=4 $L --tag vpt4 >> $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
AMM_DIRECT_ELEMS32=0 $L --tag streaming >> $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
$L --tag default_again >> $O/r2s7_lat.jsonl 2>> $O/r2s7_err.log
tail -n 5 $O/r2s7_err.log
