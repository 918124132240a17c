"""ORACLE -- test infrastructure, NOT product code (see oracle/ops.py header).

Newton-Krylov for unstable periodic orbits (uparam(1) = 2.1) restated from the reference:
  newton_krylov             core/newton_krylov.f:5-168   (q%time = the period, updated with the Newton correction, :63-67, :122)
  nonlinear_forward_map     core/newton_krylov.f:336-378 (ic_nwt, fc_nwt, orbit storage uor/vor/wor, f%time = 0)
  newton_linearized_map     core/matvec.f:381-424        (f = (Phi_T - I) q + bvec(fc) q%time ; f%time = <bvec(ic), q>)
  compute_bvec              core/matvec.f:435-475        (one first-order Navier-Stokes step: (q1 - q0) / dt)
  forward_linearized_map    core/matvec.f:187-236        (uparam(1) = 2.1: the base flow of step n is the stored orbit U^{n-1})
A Krylov vector is the tuple (v, p, time) of oracle/krylov.py; the generic newton_krylov / ts_gmres there run unchanged on it.
"""
from __future__ import annotations

import numpy as np

from .stepper import LinearizedStepper, prepare_linearized_solver


def compute_bvec(st: LinearizedStepper, v, p, dt):
    u, pr = st.linearized_map(v, p, 1, dt, adjoint="nonlinear")
    return (u - v.reshape(u.shape)) / dt, (pr - p.reshape(pr.shape)) / dt


class UPOMaps:
    """nonlinear_map / linearized_map_factory for krylov.newton_krylov with the period as an unknown."""

    def __init__(self, sem, re, w, ifvcor=False, solver="direct", cfl_target=0.5, tol_v=1e-13, tol_p=1e-13):
        self.s, self.re, self.w, self.ifvcor, self.solver, self.cfl = sem, re, w, ifvcor, solver, cfl_target
        self.tol_v, self.tol_p = tol_v, tol_p
        self.state = None

    def _stepper(self, v):
        return LinearizedStepper(self.s, v, self.re, None, tol_v=self.tol_v, tol_p=self.tol_p, solver=self.solver, ifvcor=self.ifvcor)

    def nonlinear_map(self, q):
        v, p, T = q
        dt, ns, _ = prepare_linearized_solver(self.s, v, float(T), self.cfl)
        st = self._stepper(v)
        orbit = []
        u, pr = st.linearized_map(v, p, ns, dt, adjoint="nonlinear", record=lambda i, uu, pp: orbit.append(uu.copy()))
        bfc = compute_bvec(st, u, pr, dt)                       # compute_bvec(bvec, fc_nwt)
        bic = compute_bvec(st, v, p, dt)                        # compute_bvec(btvec, ic_nwt)
        self.state = dict(st=st, orbit=orbit, dt=dt, ns=ns, bfc=bfc, bic=bic, fc=(u, pr))
        return (u - v.reshape(u.shape), pr - p.reshape(pr.shape), 0.0)

    def linearized_map_factory(self, q):
        z = self.state                                           # set by the nonlinear map of the same iterate (as in the reference)
        s, w = self.s, self.w

        def mv(x):
            xv, xp, xt = x
            yv, yp, _ = z["st"].floquet_map(xv, xp, z["ns"], z["dt"], orbit=z["orbit"])
            fv = yv - xv.reshape(yv.shape) + z["bfc"][0] * xt
            fp = yp - xp.reshape(yp.shape) + z["bfc"][1] * xt
            ft = float(sum(np.sum(z["bic"][0][d] * w * xv.reshape(yv.shape)[d]) for d in range(s.ldim)))
            return (fv, fp, ft)
        return mv
