#!/usr/bin/env python
"""Config 4 end to end on the GPU: backward-facing step, optimal transient growth at T = 1 by Krylov-Schur on the
direct-adjoint map exp(T L^+) exp(T L) (examples/back_fstep/transient_growth as shipped: `transient_growth_map`
core/matvec.f:332-349, k_dim = 64, schur_tgt = 2).  The leading eigenvalue is the optimal energy gain G(T); the shipped
optimal perturbation `pRebfs0.f00001` / response `orebfs0.f00001` pin it to G = 3.2370 (SURVEY 8c KAT-TG).
Usage: python tools/run_tg_cfg4.py [k_dim] [schur_tgt] [precond: pmg|jacobi]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nekstab_b200 import cases, lib, restart  # noqa: E402


def main():
    k_dim = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    schur_tgt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    precond = sys.argv[3] if len(sys.argv) > 3 else "pmg"
    print(json.dumps(run(k_dim, schur_tgt, precond), indent=1))


def run(k_dim=64, schur_tgt=2, precond="pmg"):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bfs.npz"))
    c = cases.bfs_case(g)
    t0 = time.time()
    ctx = lib.NekStabB200(c)
    ctx.set_params(1.0 / c.re, 1.0, c.tol_v, c.tol_p, 2000, 100000)
    if precond == "pmg":
        ctx.set_pressure_preconditioner(1, 64)
    dt, nsteps, ctarg = ctx.prepare_linearized_solver(c.end_time)
    ctx.vec_alloc(k_dim + 3)
    ctx.vec_upload(k_dim + 1, cases.add_noise(c), None)      # core/eigensolvers.f:222-278: noise, normalise, one matvec, normalise
    ctx.normalize(k_dim + 1)
    ctx.matvec(lib.DIRECT_ADJOINT, k_dim + 1, 0)
    ctx.normalize(0)
    t1 = time.time()
    vals, res, V, ncv, scnt = ctx.krylov_schur(lib.DIRECT_ADJOINT, k_dim, schur_tgt, eigen_tol=1e-6, schur_del=0.1, seed_slot=0)
    wall = time.time() - t1
    st = ctx.stats()
    # leading mode = sum_i y_i Q_i; compare with the shipped optimal perturbation (float32) up to sign
    y = V[:, 0]
    ctx.basis_gemv(k_dim, 0, np.ascontiguousarray(y.real), k_dim + 1)
    q = ctx.vec_download(k_dim + 1)[0].reshape(2, c.nel, -1)
    bm1s = ctx.get_field("bm1s").reshape(c.nel, -1)
    pre = g["pRe_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
    ip = lambda a, b: float(sum(np.sum(a[d] * bm1s * b[d]) for d in range(2)))
    cosang = abs(ip(q, pre)) / np.sqrt(ip(q, q) * ip(pre, pre))
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    restart.write_spectrum(os.path.join(out, "Spectre_Hp_cfg4.dat"), vals, res)
    summary = {"case": "backward-facing step transient growth T=1 (cfg 4)", "pressure_preconditioner": precond, "k_dim": k_dim,
               "schur_tgt": schur_tgt, "nsteps": nsteps, "dt": dt, "time_steps": st["steps"], "wall_s_krylov_schur": wall,
               "setup_s": t1 - t0, "schur_restarts": int(scnt), "converged": int(ncv),
               "pres_iters_per_step": st["pres_iters"] / max(st["steps"], 1),
               "leading_gain": float(vals[0].real), "reference_gain(|ore|^2)": 3.23700, "rel_err_gain": abs(vals[0].real - 3.23700) / 3.23700,
               "second_gain": float(vals[1].real), "leading_residual": float(res[0]),
               "cos(optimal perturbation, shipped pRe)": cosang}
    with open(os.path.join(out, "tg_cfg4_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    ctx.close()
    return summary


if __name__ == "__main__":
    main()
