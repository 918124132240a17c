#!/bin/bash
# UPO Newton path + the tests that exercise the vector algebra and the Krylov drivers it touches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_upo.py tests/test_gpu_matvec.py tests/test_gpu_restart.py tests/test_gpu_fixtures.py -q -s --durations=8 2>&1 ) > gpurun_out/upo_pytest.log 2>&1
grep -E "UPO|passed|failed|Error|error|^E " gpurun_out/upo_pytest.log | cut -c1-400 | tail -30
tail -14 gpurun_out/upo_pytest.log
