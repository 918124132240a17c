#!/usr/bin/env python
"""Config 2 end to end on the GPU: examples/cylinder/baseflow/newton as shipped -- Newton-Krylov (uparam(1) = 2: nsb_newton_krylov =
newton_krylov + ts_gmres of core/newton_krylov.f:5-297, every nonlinear_forward_map and every newton_linearized_map on the device) from the
Re = 40 steady flow (tests/golden/cyl_re40.npz = BFRe40_1cyl0.f00001) to the fixed point at Re = 50 (`viscosity = -50`, endTime 1, k_dim 100,
tolerances 1e-11, no sponge), compared with the reference's own Re = 50 base flow (stability/direct/BF_1cyl0.f00001 = tests/golden/cyl.npz)
and with the same run on the CPU oracle (tools/run_newton_cfg2_oracle.py -> tests/golden/cyl_newton_oracle.npz).
Usage: python tools/run_newton_cfg2.py [k_dim] [precond: pmg|jacobi] [maxiter_newton]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nekstab_b200 import cases, lib, restart  # noqa: E402


def run(k_dim=100, precond="pmg", tol=1e-11, maxiter_newton=10, maxiter_gmres=10):
    gold = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(gold, "cyl.npz"))
    g40 = np.load(os.path.join(gold, "cyl_re40.npz"))
    c = cases.cylinder_case(g, sponge=False)
    lx = int(g40["lx1"])
    U0 = g40["U"].reshape(-1, 2, lx * lx).transpose(1, 0, 2).astype(np.float64)
    p0 = restart.pressure_to_mesh2(g40["P"].reshape(c.nel, -1).astype(np.float64), c.lx1, 2)
    t0 = time.time()
    ctx = lib.NekStabB200(c)
    try:
        ctx.set_params(1.0 / c.re, 1.0, 1e-11, 1e-11, 2000, 100000)             # baseflow/newton/1cyl.par:31,36
        if precond == "pmg":
            ctx.set_pressure_preconditioner(1, 64)
        ctx.vec_alloc(k_dim + 6)
        ctx.vec_upload(0, U0, p0)
        w0 = ctx.norm(0)
        t1 = time.time()
        ok, it, res, hist, calls = ctx.newton_krylov(0, 1, 2, 3, 4, k_dim, c.end_time, tol, maxiter_newton=maxiter_newton,
                                                     maxiter_gmres=maxiter_gmres)
        wall = time.time() - t1
        st = ctx.stats()
        u, _ = ctx.vec_download(0)
        # distance to the shipped Re = 50 base flow in the energy norm, on the device
        ctx.vec_upload(1, c.ubase, None)
        nref = ctx.norm(1)
        ctx.vec_sub2(1, 0)
        dist = ctx.norm(1) / nref
        out = {"case": "cylinder Newton-Krylov Re 40 -> 50 (cfg 2)", "pressure_preconditioner": precond, "k_dim": k_dim, "converged": bool(ok),
               "newton_iterations": int(it), "residual_history": [float(h) for h in hist], "final_residual": float(res),
               "linearised_time_steps": int(calls), "time_steps": int(st["steps"]), "wall_s_newton": wall, "setup_s": t1 - t0,
               "norm_start": w0, "norm_shipped_BF_Re50": nref, "energy_norm_rel_diff_vs_shipped_BF_Re50": dist,
               "max_abs_diff_vs_shipped": float(np.abs(u - c.ubase.reshape(u.shape)).max())}
        orc = os.path.join(gold, "cyl_newton_oracle.npz")
        if os.path.exists(orc):
            o = np.load(orc)
            uo = o["U"].astype(np.float64).reshape(u.shape)
            out["rel_diff_vs_oracle_run(float32 fixture)"] = float(np.linalg.norm(u - uo) / np.linalg.norm(uo))
            out["oracle_residual_history"] = [float(h) for h in o["hist"]]
        return out
    finally:
        ctx.close()


if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    pc = sys.argv[2] if len(sys.argv) > 2 else "pmg"
    mn = int(sys.argv[3]) if len(sys.argv) > 3 else 10           # maxiter_newton (a short look at the first iterations)
    s = run(k, pc, maxiter_newton=mn)
    print(json.dumps(s, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "newton_cfg2_summary.json"), "w") as f:
        json.dump(s, f, indent=1)
