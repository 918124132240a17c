#!/usr/bin/env python
"""Config 1 end to end on the GPU: 2-D cylinder Re=50, direct Arnoldi / Krylov-Schur for the leading eigenpairs
(examples/cylinder/stability/direct as shipped: k_dim=200, schur_tgt=0, endTime 1, tol 1e-7/1e-9, sponge 5/5/1.7), i.e.
`krylov_schur` of core/eigensolvers.f:141-388 with every matvec on the device.  Writes Spectre_Hd.dat / Spectre_NSd.dat /
Spectre_NSd_conv.dat in the reference's format (core/eigensolvers.f:590-604) under gpurun_out/ and compares with the
shipped spectra (tests/golden/cyl.npz).
Usage: python tools/run_arnoldi_cfg1.py [k_dim] [tol_p] [tol_v] [precond: pmg|jacobi] [mxprev] [direct|adjoint]
`adjoint` runs examples/cylinder/stability/adjoint (uparam(1) = 3.2, outflow 'O' -> 'v' masks, 1cyl.usr:126-132) and compares
with the shipped Spectre_Ha.dat."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nekstab_b200 import cases, lib  # noqa: E402


def main():
    k_dim = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    tol_p = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-7
    tol_v = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-9
    precond = sys.argv[4] if len(sys.argv) > 4 else "pmg"
    mxprev = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    which = sys.argv[6] if len(sys.argv) > 6 else "direct"
    print(json.dumps(run(k_dim, tol_p, tol_v, precond, mxprev, which), indent=1))


def run(k_dim=200, tol_p=1e-7, tol_v=1e-9, precond="pmg", mxprev=0, which="direct", tag_suffix=""):
    mode = lib.ADJOINT if which == "adjoint" else lib.DIRECT
    tag = "a" if which == "adjoint" else "d"
    g = np.load(os.path.join(ROOT, "tests", "golden", "cyl.npz"))
    c = cases.cylinder_case(g)
    t0 = time.time()
    ctx = lib.NekStabB200(c)
    ctx.set_params(1.0 / c.re, 1.0, tol_v, tol_p, 2000, 100000)
    if precond == "pmg":
        ctx.set_pressure_preconditioner(1, 64)
    ctx.set_projection(mxprev)
    dt, nsteps, ctarg = ctx.prepare_linearized_solver(c.end_time)
    ctx.vec_alloc(k_dim + 3)
    # seed: noise -> normalise -> one matvec ("smoothing") -> normalise   (core/eigensolvers.f:222-278)
    ctx.vec_upload(k_dim + 1, cases.add_noise(c), None)
    ctx.normalize(k_dim + 1)
    ctx.matvec(mode, k_dim + 1, 0)
    ctx.normalize(0)
    t1 = time.time()
    vals, res, V, ncv, scnt = ctx.krylov_schur(mode, k_dim, 0, eigen_tol=1e-6, schur_del=0.1, seed_slot=0)
    wall = time.time() - t1
    st = ctx.stats()
    tau = dt * nsteps
    lam = np.log(vals.astype(complex)) / tau
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    from nekstab_b200 import restart
    restart.write_spectrum(os.path.join(out, f"Spectre_H{tag}{tag_suffix}.dat"), vals, res)          # '(3E15.7)', core/eigensolvers.f:590-604
    restart.write_spectrum(os.path.join(out, f"Spectre_NS{tag}{tag_suffix}.dat"), lam, res)
    with open(os.path.join(out, f"Spectre_NS{tag}{tag_suffix}_conv.dat"), "w") as f3:
        for i in range(k_dim):
            if res[i] < 1e-6:
                f3.write(restart.fortran_e(lam[i].real) + restart.fortran_e(lam[i].imag) + "\n")
    ref_h = g["Spectre_Ha"] if which == "adjoint" else g["Spectre_Hd"]
    ref_mu = ref_h[:, 0] + 1j * ref_h[:, 1]
    nconv_ref = int((ref_h[:, 2] < 1e-6).sum())
    # match converged reference Ritz values to ours (nearest neighbour)
    errs = []
    for m in ref_mu[:min(nconv_ref, 12)]:
        j = int(np.argmin(np.abs(vals - m)))
        errs.append(float(abs(vals[j] - m) / abs(m)))
    conv = g["Spectre_NSa_conv"] if which == "adjoint" else g["Spectre_NSd_conv"]
    ref_lam = conv[0, 0] + 1j * conv[0, 1]
    jl = int(np.argmin(np.abs(lam - ref_lam)))
    summary = {"case": f"cylinder Re=50 {which} (cfg 1)", "pressure_preconditioner": precond, "residual_projection_mxprev": mxprev, "k_dim": k_dim, "nsteps": nsteps, "dt": dt, "tol_p": tol_p, "tol_v": tol_v,
               "matvecs": k_dim + 1, "time_steps": st["steps"], "wall_s_arnoldi": wall, "setup_s": t1 - t0,
               "pres_iters_per_step": st["pres_iters"] / max(st["steps"], 1), "helm_iters_per_step": st["helm_iters"] / max(st["steps"], 1),
               "converged_ritz_pairs(res<1e-6)": int(ncv), "reference_converged": nconv_ref,
               "leading_mu": [vals[0].real, vals[0].imag], "reference_leading_mu": [ref_mu[0].real, ref_mu[0].imag],
               "leading_lambda": [lam[jl].real, lam[jl].imag], "reference_leading_lambda": [ref_lam.real, ref_lam.imag],
               "rel_err_leading_lambda": float(abs(lam[jl] - ref_lam) / abs(ref_lam)),
               "rel_err_first_converged_ritz_values": errs, "leading_residual": float(res[0]),
               "dof_steps_per_s": c.n * st["steps"] / (st["step_ms"] * 1e-3)}
    summary["ritz_values_first_24"] = [[float(v.real), float(v.imag), float(r)] for v, r in zip(vals[:24], res[:24])]
    with open(os.path.join(out, f"arnoldi_cfg1_{which}{tag_suffix}_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    ctx.close()
    return summary


if __name__ == "__main__":
    main()
