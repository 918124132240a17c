"""ORACLE -- test infrastructure, NOT product code (see oracle/ops.py header).

Scalar transport (`ifheat`, ldimt = 1) next to the velocity: the `theta` member of type krylov_vector
(core/krylov_subspace.f:8-15, inner product :41-45) advanced by Nek5000's perturbation / full scalar solvers
[UPSTREAM perturb.f heatp -> cdscalp -> makeqp (makeq_aux, convabp, makeabqp, makebdqp); heat -> cdscal -> makeq (convab)]:

  perturbation:  rhocp (d theta'/dt + U.grad theta' + u'.grad Theta) = cond lap theta' - spng_fun theta'     (nekStab_forcing_temp, core/utils.f:199)
  full:          rhocp (d theta /dt + u.grad theta)                  = cond lap theta
  momentum:      f_g += ri theta  (userf of the shipped Boussinesq cases: ffy = temp * uparam(6), e.g. thersyphon/baseflow/tsyphon.usr)

Same BDF3/EXT3 ramp and residual form as the velocity (cdscalp: H dtheta = bq - H theta^n, theta^{n+1} = theta^n + dtheta); all explicit
terms of a step use level-n fields (fluidp(1) / heatp(1) build the right-hand sides before fluidp(2) / heatp(2) solve).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse.linalg as spla

from .stepper import AB, BD, _LU_CACHE, LinearizedStepper, _h, _mesh_key


class ScalarStepper(LinearizedStepper):
    def __init__(self, sem, ubase, re, tbase, tmask, cond, rhocp=1.0, ri=0.0, gdir=1, **kw):
        super().__init__(sem, ubase, re, **kw)
        self.tb = None if tbase is None else tbase.reshape(sem.eshape)
        self.tmask = tmask.reshape(sem.eshape)
        self.cond, self.rhocp, self.ri, self.gdir = cond, rhocp, ri, gdir
        self._lu_t = {}
        self.iters_t = []

    def scalar_explicit(self, u, th, mode):
        s = self.s
        if mode == "nonlinear":
            return -self.rhocp * s.convop(u, th)
        q = -self.rhocp * (s.convop(self.ub, th) + s.convop(u, self.tb))
        if self.spng is not None:
            q = q - s.bm1 * self.spng * th
        return q

    def _scalar_solve(self, rhs, h2):
        s = self.s
        if self.solver == "direct":
            key = round(h2, 12)
            if key not in self._lu_t:
                gk = ("T", _mesh_key(s), float(self.cond), key, _h(self.tmask))
                if gk not in _LU_CACHE:
                    K, free = s.helm_sparse(self.cond, h2, mask=self.tmask)
                    _LU_CACHE[gk] = (spla.splu(K), free)
                self._lu_t[key] = _LU_CACHE[gk]
            lu, free = self._lu_t[key]
            g = np.zeros(s.nglob)
            g[s.glo.ravel()] = rhs.ravel()
            return s.from_global(lu.solve(g * free))
        # Jacobi-PCG, the velocity's cggo with the scalar's mask and coefficients
        dinv = 1.0 / s.helm_diag(self.cond, h2)
        m = self.tmask
        x = np.zeros(s.eshape); r = rhs.copy(); p = np.zeros(s.eshape)
        rtz1, it = 1.0, 0
        while True:
            z = dinv * r * m
            rtz2, rtz1 = rtz1, s.glsc3(z, r, s.mult)
            rbn2 = np.sqrt(max(s.glsc3(r * r, s.mult, s.binv), 0.0) / s.vol)
            if rbn2 <= self.tol_v or it >= self.max_iter_v:
                break
            beta = 0.0 if it == 0 else rtz1 / rtz2
            p = z + beta * p
            w = m * s.dssum(s.axhelm(p, self.cond, h2))
            alpha = rtz1 / s.glsc3(w, p, s.mult)
            x += alpha * p
            r -= alpha * w
            it += 1
        self.iters_t.append(it)
        return x

    def map_scalar(self, v, p, th, nsteps, dt, mode=False):
        """nsteps of the coupled velocity / scalar stepper from (v, p, th); mode False: direct perturbation, "nonlinear": full equations."""
        s, d = self.s, self.s.ldim
        u = v.reshape((d,) + s.eshape).copy()
        pr = p.reshape(s.eshape2).copy()
        t = th.reshape(s.eshape).copy()
        ulag = [np.zeros_like(u), np.zeros_like(u)]
        flag = [np.zeros_like(u), np.zeros_like(u)]
        tlag = [np.zeros_like(t), np.zeros_like(t)]
        qlag = [np.zeros_like(t), np.zeros_like(t)]
        plag = np.zeros_like(pr)
        for istep in range(1, nsteps + 1):
            k = min(istep, 3)
            bd, ab = BD[k], AB[k]
            h2 = bd[0] / dt
            f = self.explicit_rhs(u, mode)
            f[self.gdir] = f[self.gdir] + s.bm1 * self.ri * t
            q = self.scalar_explicit(u, t, mode)
            # ---- velocity / pressure: identical to LinearizedStepper.linearized_map
            b = ab[0] * f
            for j in range(1, k):
                b = b + ab[j] * flag[j - 1]
            hist = bd[1] * u
            for j in range(2, k + 1):
                hist = hist + bd[j] * ulag[j - 2]
            b = b + s.bm1 * hist / dt
            pt = 2.0 * pr - plag if k == 3 else pr
            r = b + s.opgradt(pt) - np.stack([s.axhelm(u[c], self.h1, h2) for c in range(d)])
            rhs = np.stack([s.mask[c] * s.dssum(r[c]) for c in range(d)])
            du = self._helm_direct(rhs, h2) if self.solver == "direct" else self._helm_pcg(rhs, h2)
            uh = u + du
            g = -s.opdiv(uh)
            if self.ifvcor:
                g = g - g.mean()
            phi = self._press_direct(g) if self.solver == "direct" else self._press_pcg(g)
            unew = uh + s.opbinv(s.opgradt(phi))
            pnew = pt + h2 * phi
            # ---- scalar
            h2t = self.rhocp * bd[0] / dt
            bq = ab[0] * q
            for j in range(1, k):
                bq = bq + ab[j] * qlag[j - 1]
            ht = bd[1] * t
            for j in range(2, k + 1):
                ht = ht + bd[j] * tlag[j - 2]
            bq = bq + s.bm1 * self.rhocp * ht / dt
            rt = self.tmask * s.dssum(bq - s.axhelm(t, self.cond, h2t))
            tnew = t + self._scalar_solve(rt, h2t)
            ulag = [u, ulag[0]]; flag = [f, flag[0]]; plag = pr
            tlag = [t, tlag[0]]; qlag = [q, qlag[0]]
            u, pr, t = unew, pnew, tnew
        return u, pr, t

    def inner_scalar(self, a, b, w=None):
        """krylov_inner_product with theta (core/krylov_subspace.f:37-45): a, b = (v, theta)."""
        w = self.s.bm1 if w is None else w
        return float(sum(np.sum(a[0][c] * w * b[0][c]) for c in range(self.s.ldim)) + np.sum(a[1] * w * b[1]))
