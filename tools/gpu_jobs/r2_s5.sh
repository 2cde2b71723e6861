#!/bin/bash
# round 2, GPU call 5 (2 GPUs): multi-GPU tests again (DeviceGuard fix), bench --gpus 2 with the fused peer exchange
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2s5_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2s5_pytest_multi.log
tail -n 5 $O/r2s5_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2s5_bench_n2.json 2> $O/r2s5_bench_n2.err; echo "bench n2 rc=$?"
grep -n "Error\|rank0\]:  \|SystemExit" $O/r2s5_bench_n2.err | head -n 20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $O/r2s5_bench_ref_n2.json 2> $O/r2s5_bench_ref_n2.err; echo "ref n2 rc=$?"
wc -l $O/r2s5_bench_n2.json $O/r2s5_bench_ref_n2.json
