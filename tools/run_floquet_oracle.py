#!/usr/bin/env python
"""KAT-Floquet on the CPU ORACLE (test infrastructure): the shipped Floquet example examples/cylinder/stability/direct_Floquet
(uparam(1) = 3.11: the base flow co-evolves with the full Navier-Stokes stepper from the UPO snapshot BF_1cyl0.f00001, period
7.9213 = 795 steps, sponge 5/5/1.7, orbit stored and replayed) through oracle/stepper.py `floquet_map` + oracle/krylov.py.
Shipped Spectre_Hd.dat: leading multipliers 1.000846 and 0.8117152.  Result of this script with k_dim = 16 (12 min of CPU,
profiles/r2_floquet_oracle.log): 1.00084625 and 0.81171206.
Usage: python tools/run_floquet_oracle.py [k_dim]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nekstab_b200 import cases, restart  # noqa: E402
from oracle import krylov  # noqa: E402
from oracle.ops import SEM  # noqa: E402
from oracle.stepper import LinearizedStepper, prepare_linearized_solver  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    g = np.load(os.path.join(ROOT, "tests", "golden", "cyl.npz"))
    u = np.load(os.path.join(ROOT, "tests", "golden", "cyl_upo.npz"))
    c = cases.cylinder_case(g)
    s = SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)
    U = u["U"].reshape(-1, 2, 36).transpose(1, 0, 2).astype(float)
    T = float(u["time"])
    dt, ns, _ = prepare_linearized_solver(s, U.reshape((2,) + s.eshape), T)
    print("T", T, "nsteps", ns, "(file istep - 1 =", int(u["istep"]) - 1, ") dt", dt)
    st = LinearizedStepper(s, U, c.re, c.spng_fun, solver="direct", ifvcor=False)
    st.spng_str_dns, st.spng_ref = 1.7, st.ub.copy()
    p2 = restart.pressure_to_mesh2(u["P"].reshape(c.nel, -1).astype(float), c.lx1, 2).reshape(s.eshape2)
    w = s.bm1 * (c.spng_fun.reshape(s.eshape) == 0)
    orbit = [None]

    def mv(q):
        v, p, orb = st.floquet_map(q[0], q[1], ns, dt, orbit=orbit[0], pbase=p2)
        orbit[0] = orb
        return (v, p)

    t0 = time.time()
    q0 = (cases.add_noise(c).reshape((2,) + s.eshape), np.zeros(s.eshape2))
    q0 = krylov.scale(q0, 1 / np.sqrt(krylov.inner(q0, q0, w)))
    q0 = mv(q0)
    q0 = krylov.scale(q0, 1 / np.sqrt(krylov.inner(q0, q0, w)))
    print("seed done", time.time() - t0, "s; orbit closure |U(T)-U(0)|/|U| =", np.linalg.norm(orbit[0][-1] - st.ub) / np.linalg.norm(st.ub), flush=True)
    vals, vecs, res, Q, H, cnt, scnt = krylov.krylov_schur(mv, q0, K, 0, w)
    print("time", time.time() - t0, "s")
    print("Floquet multipliers:", vals[:8])
    print("residuals:", res[:8])
    print("shipped Spectre_Hd.dat:", u["Spectre_Hd"][:8, 0])


if __name__ == "__main__":
    main()
