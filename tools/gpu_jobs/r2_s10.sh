#!/bin/bash
# round 2, GPU call 10 (1 GPU): ncu of the final state (launch list of bench.py, --set full of the ring kernel at 4 GiB)
cd $GRAFT_REPO_ROOT
O=gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$BENCH > $O/r2s10_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2s10_launches.csv $BENCH > $O/r2s10_ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD="python tools/ncu_case.py --dtype f32 --n 1073741824"
$CMD > $O/r2s10_plain_f32_1g.json 2> $O/r2s10_plain_f32_1g.err &&
ncu --set full --clock-control none --import-source on -k regex:argminmax_ring -s 5 -c 3 -o $O/r2s10_f32_1g_ring $CMD > $O/r2s10_ncu_f32_1g.log 2>&1
echo "ring rc=$?"
CMD="python tools/ncu_case.py --dtype u8 --n 4294967296"
$CMD > $O/r2s10_plain_u8_4g.json 2> $O/r2s10_plain_u8_4g.err &&
ncu --set full --clock-control none --import-source on -k regex:argminmax_ring -s 5 -c 3 -o $O/r2s10_u8_4g_ring $CMD > $O/r2s10_ncu_u8_4g.log 2>&1
echo "ring u8 rc=$?"
cat $O/r2s10_plain_*.json
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s10_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/r2s10_smoke.log
