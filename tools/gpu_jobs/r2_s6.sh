#!/bin/bash
# round 2, GPU call 6 (1 GPU): ncu launch list of bench.py and --set full captures of the scan at 10 M / 100 M / 2^30
cd $GRAFT_REPO_ROOT
O=gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$BENCH > $O/r2s6_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2s6_launches.csv $BENCH > $O/r2s6_ncu_launches.log 2>&1
echo "launch list rc=$?"
for c in "f32 10000000" "f32 100000000" "u8 100000000" "f32 1073741824"; do
  set -- $c
  CMD="python tools/ncu_case.py --dtype $1 --n $2"
  $CMD > $O/r2s6_plain_$1_$2.json 2> $O/r2s6_plain_$1_$2.err &&
  ncu --set full --clock-control none --import-source on -k regex:argminmax_scan -s 5 -c 3 -o $O/r2s6_$1_$2 $CMD > $O/r2s6_ncu_$1_$2.log 2>&1
  echo "$c rc=$?"
done
cat $O/r2s6_plain_*.json
