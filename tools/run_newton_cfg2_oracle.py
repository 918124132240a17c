#!/usr/bin/env python
"""BASELINE config 2 on the CPU oracle: examples/cylinder/baseflow/newton as shipped -- Newton-Krylov (uparam(1) = 2, k_dim 100, endTime 1,
tolerances 1e-11) from the Re = 40 steady flow BFRe40_1cyl0.f00001 to the fixed point at Re = 50, compared with the reference's own Re = 50
base flow (examples/cylinder/stability/direct/BF_1cyl0.f00001 = tests/golden/cyl.npz).  Writes profiles/r2_newton_cfg2_oracle.json and the
converged field tests/golden/cyl_newton_oracle.npz (float32)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nekstab_b200 import cases, restart  # noqa: E402
from oracle import krylov  # noqa: E402
from oracle.stepper import LinearizedStepper, prepare_linearized_solver  # noqa: E402
from util import GOLD, make_oracle  # noqa: E402


def main(k_dim=100, tol=1e-11, maxiter_newton=10):
    g = np.load(os.path.join(GOLD, "cyl.npz"))
    g40 = np.load(os.path.join(GOLD, "cyl_re40.npz"))
    c = cases.cylinder_case(g, sponge=False)                      # baseflow/newton/1cyl.par has no sponge
    s = make_oracle(c)
    lx = int(g40["lx1"])
    U0 = g40["U"].reshape(-1, 2, lx * lx).transpose(1, 0, 2).astype(np.float64).reshape((2,) + s.eshape)
    p0 = restart.pressure_to_mesh2(g40["P"].reshape(c.nel, -1).astype(np.float64), c.lx1, 2).reshape(s.eshape2)
    Uref = c.ubase.reshape((2,) + s.eshape)
    w = s.bm1
    state = {"matvecs": 0, "t0": time.time()}

    def nl(q):
        dt, ns, _ = prepare_linearized_solver(s, q[0], c.end_time)
        st = LinearizedStepper(s, q[0], c.re, None, solver="direct", ifvcor=c.ifvcor)
        fv, fp, _, _ = st.nonlinear_forward_map(q[0], q[1], ns, dt)
        state.update(st=st, dt=dt, ns=ns)
        print("  nonlinear map: nsteps %d, |f|^2 = %.4e, %.0f s" % (ns, krylov.inner((fv, fp), (fv, fp), w), time.time() - state["t0"]), flush=True)
        return (fv, fp)

    def lin(q):
        st, dt, ns = state["st"], state["dt"], state["ns"]

        def mv(x):
            state["matvecs"] += 1
            y = st.linearized_map(x[0], x[1], ns, dt)
            return (y[0] - x[0], y[1] - x[1])
        return mv

    q, it, hist = krylov.newton_krylov(nl, lin, (U0, p0), k_dim, tol, w, maxiter_newton=maxiter_newton, maxiter_gmres=10)
    d = q[0] - Uref
    err = float(np.sqrt(krylov.inner((d,), (d,), w) / krylov.inner((Uref,), (Uref,), w)))
    out = dict(newton_iterations=it, residual_history=[float(h) for h in hist], linearised_matvecs=state["matvecs"], wall_s=time.time() - state["t0"],
               energy_norm_rel_diff_vs_shipped_BF_Re50=err, max_abs_diff=float(np.abs(d).max()),
               start_rel_diff_vs_shipped=float(np.sqrt(krylov.inner((U0 - Uref,), (U0 - Uref,), w) / krylov.inner((Uref,), (Uref,), w))))
    print(json.dumps(out, indent=1))
    with open(os.path.join(ROOT, "profiles", "r2_newton_cfg2_oracle.json"), "w") as f:
        json.dump(out, f, indent=1)
    np.savez_compressed(os.path.join(GOLD, "cyl_newton_oracle.npz"), U=q[0].astype(np.float32), hist=np.array(hist), iters=it, matvecs=state["matvecs"])


if __name__ == "__main__":
    main()
