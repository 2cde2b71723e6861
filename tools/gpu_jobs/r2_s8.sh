#!/bin/bash
# round 2, GPU call 8 (1 GPU): TMA ring vs register ping-pong (stand-alone harness, tools/exp/tma_ring.cu)
cd $GRAFT_REPO_ROOT
O=gpurun_out
B=tools/exp/bin/tma_ring
run() { timeout 60 $B "$@" >> $O/r2s8_tma_ring.jsonl 2>> $O/r2s8_err.log || echo "{\"failed\": \"$*\", \"rc\": $?}" >> $O/r2s8_tma_ring.jsonl; }
for mib in 4096 400 40; do
  it=20; [ $mib -lt 1000 ] && it=100
  run 0 $mib 4 0 16 $it
  run 0 $mib 5 0 16 $it
  for mode in 1 2; do
    run $mode $mib 4 3 16 $it
    run $mode $mib 3 4 16 $it
    run $mode $mib 2 6 16 $it
    run $mode $mib 2 3 32 $it
    run $mode $mib 1 6 32 $it
    run $mode $mib 2 4 32 $it
    run $mode $mib 1 3 64 $it
    run $mode $mib 6 2 16 $it
  done
done
cat $O/r2s8_tma_ring.jsonl | tail -n 60
