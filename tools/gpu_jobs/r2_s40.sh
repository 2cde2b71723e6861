#!/bin/bash
# round 2, GPU call 40 (2 GPUs): multi-GPU tests and the driver's N=2 bench launch on the final state
cd $GRAFT_REPO_ROOT
O=gpurun_out
( timeout 240 python -m pytest tests/test_gpu_multi.py -x -q ) > $O/r2s40_pytest_multi.log 2>&1; echo "multi tests rc=$?"; tail -n 3 $O/r2s40_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2s40_bench_n2.json 2> $O/r2s40_bench_n2.err; echo "bench n2 rc=$?"
tail -n 2 $O/r2s40_bench_n2.err
