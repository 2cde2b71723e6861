#!/bin/bash
# round 2, GPU call 36 (1 GPU): 32-bit keys through the tagged-slot grid step instead of packed atomics? (same library, knobs)
cd $GRAFT_REPO_ROOT
O=gpurun_out
L=$PWD/argminmax_b200/lib/libargminmax_b200.so
timeout 600 python tools/exp/latency_ab.py --a $L --b $L --env-b AMM_DIRECT_PACKED=0,AMM_SCAN_PACKED=0 --dtypes f32,u8,f16 --max-log2 25 > $O/r2s36_packed_vs_slots.jsonl 2> $O/r2s36_err.log; echo "ab rc=$?"
timeout 600 python tools/exp/latency_ab.py --a $L --b $L --env-b AMM_DIRECT_PACKED=0,AMM_SCAN_PACKED=0 --dtypes f32 --mode 1 --max-log2 25 >> $O/r2s36_packed_vs_slots.jsonl 2>> $O/r2s36_err.log; echo "ab rc=$?"
tail -n 3 $O/r2s36_err.log
