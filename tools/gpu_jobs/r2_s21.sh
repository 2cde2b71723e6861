#!/bin/bash
# round 2, GPU call 21 (8 GPUs): the driver's scaling sequence on the final state (both arms at N = 8, ours at 4 and 2) + multi-GPU tests
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/r2s21_bench_ref_n8.json 2> $O/r2s21_bench_ref_n8.err; echo "ref n8 rc=$?"
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > $O/r2s21_bench_n$N.json 2> $O/r2s21_bench_n$N.err; echo "bench n$N rc=$?"
  grep -n "Error\|rank0\]:  \|SystemExit\|failed" $O/r2s21_bench_n$N.err | head -n 10
done
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2s21_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2s21_pytest_multi.log
tail -n 3 $O/r2s21_pytest_multi.log
wc -l $O/r2s21_bench_n8.json $O/r2s21_bench_n4.json $O/r2s21_bench_n2.json $O/r2s21_bench_ref_n8.json
