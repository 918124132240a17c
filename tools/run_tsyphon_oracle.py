#!/usr/bin/env python
"""examples/thersyphon/baseflow on the CPU oracle: Newton-Krylov (uparam(1) = 2) from the shipped Ra = 400 solution to the steady state at
Ra = 500 (tsyphon.par: startfrom BF_Ra400, userparam06 = 500, endTime 0.1, k_dim 100, tolerances 1e-11).  Writes the residual history and
the converged fields to tests/golden/tsyphon_oracle.npz -- the fixture of tests/test_gpu_scalar.py::test_thermosyphon_newton."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nekstab_b200 import cases, restart  # noqa: E402
from oracle import krylov  # noqa: E402
from oracle.scalar import ScalarStepper  # noqa: E402
from oracle.stepper import prepare_linearized_solver  # noqa: E402
from util import GOLD, make_oracle  # noqa: E402


def run(ra=500.0, k_dim=40, tol=1e-11, maxiter_newton=12, maxiter_gmres=10, solver="pcg", verbose=True):
    c = cases.thermosyphon_case(np.load(os.path.join(GOLD, "tsyphon.npz")), ra=ra)
    s = make_oracle(c)
    U, T, tm = c.ubase.reshape((2,) + s.eshape), c.extra["T"].reshape(s.eshape), c.extra["tmask"].reshape(s.eshape)
    p2 = restart.pressure_to_mesh2(c.extra["P"], c.lx1, 2).reshape(s.eshape2)
    w = s.bm1
    kw = dict(cond=1.0, rhocp=1.0, ri=float(c.extra["ri"]), gdir=1, solver=solver, tol_v=1e-13, tol_p=1e-13, max_iter_v=5000, ifvcor=True)
    state = {}

    def inner(a, b):                                   # krylov_inner_product with theta: q = (v, p, theta)
        return float(sum(np.sum(a[0][d] * w * b[0][d]) for d in range(2)) + np.sum(a[2] * w * b[2]))

    def nl(q):
        dt, ns, _ = prepare_linearized_solver(s, q[0], c.end_time)
        st = ScalarStepper(s, q[0], c.re, q[2], tm, **kw)
        u, pr, t = st.map_scalar(q[0], q[1], q[2], ns, dt, mode="nonlinear")
        state.update(st=st, dt=dt, ns=ns)
        return (u - q[0], pr - q[1], t - q[2])

    def lin(q):
        st, dt, ns = state["st"], state["dt"], state["ns"]   # linearised about the current iterate: ubase, tbase <- q (core/newton_krylov.f:374-375)
        def mv(x):
            u, pr, t = st.map_scalar(x[0], x[1], x[2], ns, dt)
            return (u - x[0], pr - x[1], t - x[2])
        return mv

    # the generic drivers of oracle/krylov.py with the theta-aware inner product
    saved = krylov.inner
    krylov.inner = lambda a, b, _w: inner(a, b)
    try:
        t0 = time.time()
        q, it, hist = krylov.newton_krylov(nl, lin, (U, p2, T), k_dim, tol, w, maxiter_newton=maxiter_newton, maxiter_gmres=maxiter_gmres)
    finally:
        krylov.inner = saved
    if verbose:
        print("thermosyphon Newton Ra = %g: %d iterations, residual history %s, %.1f s" % (ra, it, ["%.3e" % h for h in hist], time.time() - t0))
    return c, s, q, it, hist


if __name__ == "__main__":
    c, s, q, it, hist = run()
    np.savez_compressed(os.path.join(GOLD, "tsyphon_oracle.npz"), U=q[0], P=q[1], T=q[2], hist=np.array(hist), iters=it)
