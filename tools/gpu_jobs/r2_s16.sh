#!/bin/bash
# round 2, GPU call 16 (2 GPUs): final-state validation of the multi-GPU path: tests + both bench arms as the driver runs them
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2s16_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2s16_pytest_multi.log
tail -n 3 $O/r2s16_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/r2s16_bench_ref_n2.json 2> $O/r2s16_bench_ref_n2.err; echo "ref rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2s16_bench_n2.json 2> $O/r2s16_bench_n2.err; echo "bench n2 rc=$?"
grep -n "Error\|rank0\]:  \|SystemExit" $O/r2s16_bench_n2.err | head -n 10
wc -l $O/r2s16_bench_n2.json $O/r2s16_bench_ref_n2.json
