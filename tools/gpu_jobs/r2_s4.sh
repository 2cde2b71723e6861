#!/bin/bash
# round 2, GPU call 4 (2 GPUs): multi-GPU parity tests, bench --gpus 2 (peer + nccl), sharded latency
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2s4_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2s4_pytest_multi.log
tail -n 15 $O/r2s4_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2s4_bench_n2.json 2> $O/r2s4_bench_n2.err; echo "bench n2 rc=$?"
tail -n 20 $O/r2s4_bench_n2.err
AMM_SHARDED_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/r2s4_bench_n2_nccl.json 2> $O/r2s4_bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
python tools/sharded_latency.py > $O/r2s4_sharded_latency.jsonl 2> $O/r2s4_sharded_latency.err
AMM_SHARDED_THREADS=0 python tools/sharded_latency.py > $O/r2s4_sharded_latency_nothreads.jsonl 2>> $O/r2s4_sharded_latency.err
tail -n 5 $O/r2s4_sharded_latency.err
