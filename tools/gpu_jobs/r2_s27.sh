#!/bin/bash
# round 2, GPU call 27 (1 GPU): phase stamps of the latency kernel (L2-resident input), 4 Ki .. 4 Mi f32
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python tools/phase_stamps.py --direct --copies 1 --tag direct_warm --cases f32:4096,f32:8192,f32:131072,f32:1048576,f32:4194304,u8:1048576,f64:1048576 > $O/r2s27_stamps_direct.jsonl 2> $O/r2s27_err.log; echo "rc=$?"
tail -n 3 $O/r2s27_err.log
