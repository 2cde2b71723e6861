#!/bin/bash
# round 2, GPU call 13 (8 GPUs): where does the host-slice call lose against plain copies when 8 GPUs pull at once?
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/exp/h2d_multi.py > $O/r2s13_h2d_8.jsonl 2> $O/r2s13_h2d_8.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tools/exp/h2d_multi.py > $O/r2s13_h2d_4.jsonl 2> $O/r2s13_h2d_4.err; echo "rc=$?"
cat $O/r2s13_h2d_8.jsonl $O/r2s13_h2d_4.jsonl
grep -n "Error" $O/r2s13_h2d_8.err | head -n 5
