#!/bin/bash
# round 2, GPU call 11 (8 GPUs): bench --gpus 8 and 4 (what the driver's scaling run does), multi-GPU tests, sharded latency
cd $GRAFT_REPO_ROOT
O=gpurun_out
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 20 --warmup 5 > $O/r2s11_bench_n$N.json 2> $O/r2s11_bench_n$N.err; echo "bench n$N rc=$?"
  grep -n "Error\|rank0\]:  \|SystemExit\|failed" $O/r2s11_bench_n$N.err | head -n 10
done
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2s11_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2s11_pytest_multi.log
tail -n 4 $O/r2s11_pytest_multi.log
timeout 300 python tools/sharded_latency.py > $O/r2s11_sharded_latency.jsonl 2> $O/r2s11_sharded_latency.err
AMM_SHARDED_THREADS=0 timeout 300 python tools/sharded_latency.py > $O/r2s11_sharded_latency_nothreads.jsonl 2>> $O/r2s11_sharded_latency.err
nvidia-smi topo -m > $O/r2s11_topo.txt 2>&1
lscpu | head -n 25 > $O/r2s11_lscpu.txt 2>&1
wc -c $O/r2s11_bench_n8.json $O/r2s11_bench_n4.json
