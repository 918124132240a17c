#!/bin/bash
# scalar transport (theta) on the GPU + the matvec parity tests that share the touched stepper code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 280 python -m pytest tests/test_gpu_scalar.py "tests/test_gpu_matvec.py::test_linearized_maps" -q -s --durations=6 2>&1 ) > gpurun_out/scalar_pytest.log 2>&1
grep -E "KAT|passed|failed|^E |Error" gpurun_out/scalar_pytest.log | cut -c1-300 | tail -40
tail -12 gpurun_out/scalar_pytest.log
