#!/bin/bash
# round 2, GPU call 30 (1 GPU): latency skeleton again, with fence.acq_rel / atom.acq_rel variants of the grid step
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 tools/exp/bin/latency_floor > $O/r2s30_latency_floor.jsonl 2> $O/r2s30_err.log; echo "floor rc=$?"
