#!/bin/bash
# round 2, GPU call 37 (1 GPU): latency-kernel limit for 64-bit keys (2^16 vs 2^19 elements), same library, knob
cd $GRAFT_REPO_ROOT
O=gpurun_out
L=$PWD/argminmax_b200/lib/libargminmax_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "latency_kernel_regimes" > $O/r2s37_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/r2s37_pytest.log
timeout 600 python tools/exp/latency_ab.py --a $L --b $L --env-b AMM_DIRECT_ELEMS64=524288 --dtypes i64,f64 --min-log2 15 --max-log2 21 > $O/r2s37_elems64.jsonl 2> $O/r2s37_err.log; echo "ab rc=$?"
timeout 600 python tools/exp/latency_ab.py --a $L --b $L --env-b AMM_DIRECT_ELEMS64=524288 --dtypes f64 --mode 1 --min-log2 15 --max-log2 21 >> $O/r2s37_elems64.jsonl 2>> $O/r2s37_err.log; echo "ab rc=$?"
timeout 600 python tools/exp/latency_ab.py --a $L --b $L --env-b AMM_DIRECT_ELEMS32=2097152,AMM_DIRECT_CTAS=4 --dtypes f32,u8 --min-log2 18 --max-log2 22 >> $O/r2s37_elems64.jsonl 2>> $O/r2s37_err.log; echo "ab rc=$?"
tail -n 3 $O/r2s37_err.log
