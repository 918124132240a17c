"""BASELINE.json config 2 end to end on the GPU (examples/cylinder/baseflow/newton as shipped), against the reference's own Re = 50 base
flow and the CPU oracle's run of the same case.  The file sorts last in `pytest -m gpu`."""
import os
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_cfg2_cylinder_newton_krylov_base_flow():
    """Config 2 as shipped (examples/cylinder/baseflow/newton: uparam(1) = 2, startFrom BFRe40_1cyl0.f00001, viscosity -50, endTime 1,
    k_dim 100, tolerances 1e-11): Newton-Krylov from the Re = 40 flow to the fixed point at Re = 50.  The CPU oracle's run
    (tools/run_newton_cfg2_oracle.py, profiles/r2_newton_cfg2_oracle.json) takes 4 Newton iterations, residuals 2.589e-03, 3.207e-06,
    1.813e-10, 9.583e-12, and lands on the reference's own Re = 50 base flow (stability/direct/BF_1cyl0.f00001, double precision) to
    1.9e-10 in the energy norm -- the reference's Newton stopped at the same iterate (its residual is 9.6e-12, KAT of test_gpu_newton.py).
    Default: the first two Newton iterations (154 linearised matvecs, 75 s on a B200 -- the run recorded in
    profiles/r2_newton_cfg2_gpu_first2.json); NSB_LONG_TESTS=1 runs all four (about 2.5 minutes)."""
    import run_newton_cfg2
    full = os.environ.get("NSB_LONG_TESTS", "0") == "1"
    s = run_newton_cfg2.run(100, "pmg", maxiter_newton=10 if full else 2)
    print({k: s[k] for k in ("newton_iterations", "residual_history", "wall_s_newton", "linearised_time_steps",
                             "energy_norm_rel_diff_vs_shipped_BF_Re50", "max_abs_diff_vs_shipped", "rel_diff_vs_oracle_run(float32 fixture)")})
    h, ho = s["residual_history"], s["oracle_residual_history"]
    assert np.allclose(np.log10(h[:2]), np.log10(ho[:2]), atol=1e-4), (h, ho)      # 2.5891715e-03, 3.2068673e-06: equal to 8 digits on hardware
    if not full:
        # after two iterations the iterate is already 8.4e-7 from the shipped Re = 50 flow (the Re = 40 start: 6.1e-3)
        assert not s["converged"] and len(h) == 2
        assert s["energy_norm_rel_diff_vs_shipped_BF_Re50"] < 2e-6
        return
    assert s["converged"] and s["final_residual"] < 1e-11
    assert s["newton_iterations"] in (4, 5)                         # the oracle's 4th residual (9.58e-12) is 4 % under the exit test
    assert np.allclose(np.log10(h[:3]), np.log10(ho[:3]), atol=0.01), (h, ho)
    assert s["energy_norm_rel_diff_vs_shipped_BF_Re50"] < 1e-7      # oracle: 1.9e-10
    assert s["rel_diff_vs_oracle_run(float32 fixture)"] < 5e-7
