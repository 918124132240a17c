#!/usr/bin/env python
"""Config 3 (lid-driven cavity: direct + adjoint eigenproblem + wavemaker) on the CPU ORACLE: generates the golden vectors
tests/golden/cav_oracle.npz that tests/test_gpu_cavity.py compares the CUDA path with.  The reference ships no spectrum for
this case (SURVEY.md 8 cfg 3), so parity is anchored on the oracle (numpy/scipy, sparse-direct solves = the solver-converged
step), itself pinned to the reference's cylinder / BFS fixtures (tests/test_oracle_fixtures.py).

What runs (core/usr_extra.f mode 3.1 then 3.2 then 4.1 with the shipped cav.par / cav.usr settings): seed = add_noise ->
normalise -> one matvec -> normalise (core/eigensolvers.f:222-278); krylov_schur (k_dim 90, schur_tgt 4, eigen_tol 1e-6,
schur_del 0.1) on exp(T L), T = 0.5 = 348 steps, then on exp(T L+); leading modes Q y (outpost_ks :554-564); bi-orthonormalise
and wavemaker (core/sensitivity.f:7-81, 428-504).
Usage: python tools/make_cav_oracle.py     (~10 min of CPU; run in the build container, commit the .npz)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nekstab_b200 import cases, sensitivity  # noqa: E402
from oracle import krylov  # noqa: E402
from oracle.ops import SEM  # noqa: E402
from oracle.stepper import LinearizedStepper, prepare_linearized_solver  # noqa: E402

K_DIM, SCHUR_TGT = 90, 4


def leading_mode(vals, vecs, Q, k_dim):
    """First Ritz pair with positive imaginary part (or the leading real one): mode = sum_i y_i Q_i."""
    i = next((j for j in range(len(vals)) if vals[j].imag >= 0), 0)
    y = vecs[:, i]
    re = sum(y[j].real * Q[j][0] for j in range(k_dim))
    im = sum(y[j].imag * Q[j][0] for j in range(k_dim))
    return i, re, im


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cav.npz"))
    c = cases.cavity_case(g)
    s = SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)
    dt, nsteps, _ = prepare_linearized_solver(s, c.ubase.reshape((2,) + s.eshape), c.end_time)
    st = LinearizedStepper(s, c.ubase, c.re, None, solver="direct", ifvcor=True)
    w = s.bm1
    out = {"dt": dt, "nsteps": nsteps, "k_dim": K_DIM, "schur_tgt": SCHUR_TGT}
    modes = {}
    for tag, adj in (("d", False), ("a", True)):
        t0 = time.time()
        mv = lambda q, adj=adj: st.linearized_map(q[0], q[1], nsteps, dt, adjoint=adj)
        q0 = (cases.add_noise(c).reshape((2,) + s.eshape), np.zeros(s.eshape2))
        q0 = krylov.scale(q0, 1.0 / np.sqrt(krylov.inner(q0, q0, w)))
        q0 = mv(q0)
        q0 = krylov.scale(q0, 1.0 / np.sqrt(krylov.inner(q0, q0, w)))
        vals, vecs, res, Q, H, cnt, scnt = krylov.krylov_schur(mv, q0, K_DIM, SCHUR_TGT, w, eigen_tol=1e-6, schur_del=0.1)
        i, re, im = leading_mode(vals, vecs, Q, K_DIM)
        modes[tag] = (re, im)
        out[f"vals_{tag}"] = vals[:24]
        out[f"res_{tag}"] = res[:24]
        out[f"cnt_{tag}"], out[f"scnt_{tag}"], out[f"lead_{tag}"] = cnt, scnt, i
        print(f"{tag}: {time.time() - t0:.0f} s, converged {cnt}, restarts {scnt}, leading mu = {vals[i]}, lambda = {np.log(vals[i]) / (dt * nsteps)}")
        print("   first Ritz values:", vals[:8], "residuals", res[:8])
    flat = lambda a: a.reshape(2, c.nel, -1)
    bm1s = w.reshape(c.nel, -1)
    d_re, d_im, a_re, a_im = sensitivity.biorthogonalize(flat(modes["d"][0]), flat(modes["d"][1]), flat(modes["a"][0]), flat(modes["a"][1]), bm1s)
    wm = sensitivity.wave_maker(flat(modes["d"][0]), flat(modes["d"][1]), flat(modes["a"][0]), flat(modes["a"][1]), bm1s)
    out.update(d_re=d_re, d_im=d_im, a_re=a_re, a_im=a_im, wavemaker=wm)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cav_oracle.npz"), **out)
    print("wavemaker max", wm.max(), "at", np.unravel_index(np.argmax(wm), wm.shape))


if __name__ == "__main__":
    main()
