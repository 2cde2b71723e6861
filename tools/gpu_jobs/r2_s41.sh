#!/bin/bash
# round 2, GPU call 41 (1 GPU): smoke() with the extended case list + parity subset after moving grid_slot_check into record_slot.h
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s41_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/r2s41_smoke.log
( timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q ) > $O/r2s41_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/r2s41_pytest.log
