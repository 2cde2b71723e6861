#!/bin/bash
# round 2, GPU call 1: parity of the reworked tail + phase stamps (baseline vs packed, geometries)
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not large and not bounds" > $O/r2s1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s1_pytest.log
tail -3 $O/r2s1_pytest.log
for packed in 0 1; do
  AMM_SCAN_PACKED=$packed python tools/phase_stamps.py --tag packed$packed >> $O/r2s1_stamps.jsonl 2>> $O/r2s1_err.log
done
for cfg in "256 4 0" "256 4 4" "256 8 4" "128 8 0" "256 2 1" "256 4 1" "512 2 0"; do
  set -- $cfg
  python tools/phase_stamps.py --tag "t$1_c$2_v$3" --threads $1 --ctas $2 --variant $3 --cases f32:10000000,f32:100000000,u8:100000000 >> $O/r2s1_stamps.jsonl 2>> $O/r2s1_err.log
done
python tools/results_table.py > $O/r2s1_results_table.jsonl 2>> $O/r2s1_err.log
AMM_SCAN_PACKED=0 python tools/results_table.py > $O/r2s1_results_table_unpacked.jsonl 2>> $O/r2s1_err.log
tail -5 $O/r2s1_err.log
