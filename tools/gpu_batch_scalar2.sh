#!/bin/bash
# thermosyphon Newton end to end + regression of the orbit / Newton paths with the final library + smoke()
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 170 python -m pytest "tests/test_gpu_scalar.py::test_thermosyphon_newton_end_to_end" tests/test_gpu_upo.py tests/test_gpu_newton.py "tests/test_gpu_floquet.py::test_floquet_map_against_oracle" -q -s --durations=5 2>&1 ) > gpurun_out/scalar2_pytest.log 2>&1
grep -E "thermosyphon|converged fields|KAT|UPO|passed|failed|^E " gpurun_out/scalar2_pytest.log | cut -c1-400 | tail -20
tail -8 gpurun_out/scalar2_pytest.log
