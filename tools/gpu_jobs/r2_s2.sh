#!/bin/bash
# round 2, GPU call 2: parity of the contiguous-partition / redux / one-trip latency kernels; stamps; latency A/B; bench.py
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not large" > $O/r2s2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s2_pytest.log
tail -3 $O/r2s2_pytest.log
python tools/phase_stamps.py --tag auto >> $O/r2s2_stamps.jsonl 2>> $O/r2s2_err.log
AMM_SCAN_CONTIG=0 python tools/phase_stamps.py --tag gridstride >> $O/r2s2_stamps.jsonl 2>> $O/r2s2_err.log
AMM_SCAN_CONTIG=1 python tools/phase_stamps.py --tag contig --cases f32:1073741824 >> $O/r2s2_stamps.jsonl 2>> $O/r2s2_err.log
for cfg in "256 4 0" "512 2 1" "256 4 1" "512 1 0" "384 2 0"; do
  set -- $cfg
  python tools/phase_stamps.py --tag "t$1_c$2_v$3" --threads $1 --ctas $2 --variant $3 --cases f32:10000000,f32:100000000,u8:100000000 >> $O/r2s2_stamps.jsonl 2>> $O/r2s2_err.log
done
python tools/results_table.py > $O/r2s2_results_table.jsonl 2>> $O/r2s2_err.log
python tools/latency_c.py --tag default --calls 1000 --max-log2 22 > $O/r2s2_latency_default.jsonl 2>> $O/r2s2_err.log
AMM_DIRECT_ELEMS32=0 AMM_DIRECT_ELEMS64=0 python tools/latency_c.py --tag streaming_only --calls 1000 --max-log2 22 > $O/r2s2_latency_streaming.jsonl 2>> $O/r2s2_err.log
AMM_DIRECT_ELEMS32=16777216 AMM_DIRECT_ELEMS64=16777216 python tools/latency_c.py --tag direct_max --calls 1000 --max-log2 22 > $O/r2s2_latency_direct.jsonl 2>> $O/r2s2_err.log
python bench.py --steps 20 --warmup 5 > $O/r2s2_bench_n1.json 2> $O/r2s2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > $O/r2s2_bench_ref.json 2> $O/r2s2_bench_ref.err; echo "ref rc=$?"
tail -5 $O/r2s2_err.log $O/r2s2_bench_n1.err
