#!/bin/bash
# round 2, GPU call 26 (1 GPU): grid-step variants in the latency skeleton (fence-free atomics, tagged
# slots polled by CTA 0, one cluster + DSMEM) and ragged-window throughput
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 tools/exp/bin/latency_floor > $O/r2s26_latency_floor.jsonl 2> $O/r2s26_err.log; echo "floor rc=$?"
timeout 300 python tools/windows_bench.py --ragged --dtypes f32,u8,f16,f64 --windows 16,64,129,256,500,1000,4096,20000 > $O/r2s26_windows_ragged.jsonl 2>> $O/r2s26_err.log; echo "ragged rc=$?"
tail -n 5 $O/r2s26_err.log
