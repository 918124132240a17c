"""ORACLE -- test infrastructure, NOT product code (see oracle/ops.py header).

One linearised Navier-Stokes step in perturbation mode and the maps nekStab builds from it:
  forward_linearized_map   core/matvec.f:163-241      adjoint_linearized_map   core/matvec.f:249-325
  transient_growth_map     core/matvec.f:332-349      newton_linearized_map    core/matvec.f:381-402
  ts_force_sensitivity_map core/matvec.f:357-373      prepare_linearized_solver core/matvec.f:1-52
The step itself is Nek5000's `nek_advance` with ifpert=.true. (P_N-P_{N-2}, BDF3/EXT3 with the order
ramping 1,2,3 because matvec restarts istep at 1, core/matvec.f:216) [UPSTREAM drive1.f, perturb.f
fluidp/perturbv/makefp/advabp/advabp_adjoint/makextp/makebdfp/cresvipp/incomprp], restated as SURVEY.md
App. E items 1-8.  User forcing = nekStab_forcing's perturbation branch (core/utils.f:172-177): -spng_fun*u'.

Two solver modes: 'direct' (sparse LU of the assembled Helmholtz and E operators -- defines the
solver-converged step used for fixture pinning) and 'pcg' (Jacobi-preconditioned CG with Nek-style
stopping norms -- the algorithm the CUDA path runs; used for iteration-level parity).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse.linalg as spla

import hashlib

from .ops import SEM

# Process-wide cache of the sparse-LU factors of the 'direct' mode, keyed by the CONTENT of the mesh (coordinates, numbering, masks) and
# the operator coefficients: the GPU parity tests build many steppers on the same handful of small meshes, and the factorisations are
# what the oracle spends its time on (6.8 of 10.8 s for six steps on the 3-D lx1 = 8 box).
_LU_CACHE: dict = {}


def _h(a) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def _mesh_key(s: SEM):
    return (s.ldim, s.lx1, s.lx2, _h(s.X), _h(s.glo), _h(s.mask))      # recomputed on every (rare) look-up: no stale keys

BD = {1: [1.0, 1.0], 2: [1.5, 2.0, -0.5], 3: [11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0]}   # [UPSTREAM subs1.f setbd], constant dt
AB = {1: [1.0], 2: [2.0, -1.0], 3: [3.0, -3.0, 1.0]}                                   # [UPSTREAM subs1.f setabbd]


def prepare_linearized_solver(sem: SEM, ubase, end_time, cfl_target=0.5):
    """dt and nsteps exactly as core/matvec.f:28-36: ctarg = max sum(|u_i|/dx_i); dt = CFL/ctarg;
    nsteps = ceiling(T/dt); dt = T/nsteps."""
    if cfl_target > 1.0:
        cfl_target = 0.5
    ctarg = sem.cfl_sum(ubase)
    dt = cfl_target / ctarg
    nsteps = int(math.ceil(end_time / dt))
    return end_time / nsteps, nsteps, ctarg


class LinearizedStepper:
    def __init__(self, sem: SEM, ubase, re, spng_fun=None, tol_v=1e-9, tol_p=1e-7, solver="direct",
                 max_iter_v=1000, max_iter_p=20000, ifvcor=None, pressure_precond=None):
        self.s = sem
        d = sem.ldim
        self.ub = ubase.reshape((d,) + sem.eshape)
        self.h1 = 1.0 / re
        self.spng = None if spng_fun is None else spng_fun.reshape(sem.eshape)
        self.tol_v, self.tol_p = tol_v, tol_p
        self.solver = solver
        self.max_iter_v, self.max_iter_p = max_iter_v, max_iter_p
        # all-Dirichlet/periodic velocity => E has the constant null vector [UPSTREAM ifvcor / ortho]
        # Nek decides from the boundary conditions; pass it when known, else test ||E 1|| numerically
        if ifvcor is None:
            e1 = sem.cdabdtp(np.ones(sem.eshape2))
            ifvcor = bool(np.linalg.norm(e1) < 1e-9 * np.linalg.norm(sem.e_diag()))
        self.ifvcor = bool(ifvcor)
        self._lu_h = {}
        self._lu_e = None
        self._ediag = None
        self.pressure_precond = pressure_precond     # None: Jacobi; else an object with .apply(r) (oracle/pmg.py PMG)
        self.spng_str_dns, self.spng_ref = 0.0, None # DNS sponge: jp = 0 branch of nekStab_forcing (core/utils.f:166-171)
        self.iters_v, self.iters_p = [], []
        self.dt = None

    # -------------------------------------------------------------- solvers
    def _helm_direct(self, rhs, h2):
        s = self.s
        key = round(h2, 12)
        out = []
        for c in range(s.ldim):
            kk = (key, c)
            if kk not in self._lu_h:
                same = [k for k in self._lu_h if k[0] == key and np.array_equal(s.mask[k[1]], s.mask[c])]
                if same:
                    self._lu_h[kk] = self._lu_h[same[0]]
                else:
                    gk = ("H", _mesh_key(s), float(self.h1), key, _h(s.mask[c]))
                    if gk not in _LU_CACHE:
                        K, free = s.helm_sparse(self.h1, h2, c)
                        _LU_CACHE[gk] = (spla.splu(K), free)
                    self._lu_h[kk] = _LU_CACHE[gk]
            lu, free = self._lu_h[kk]
            g = np.zeros(s.nglob)
            g[s.glo.ravel()] = rhs[c].ravel()          # rhs is already assembled (consistent across copies)
            out.append(s.from_global(lu.solve(g * free)))
        return np.stack(out)

    def _helm_pcg(self, rhs, h2):
        """Jacobi-PCG [UPSTREAM hmholtz.f cggo]; stop on sqrt(sum r^2 mult binv / vol) <= tol (absolute)."""
        s = self.s
        dinv = 1.0 / s.helm_diag(self.h1, h2)
        out, its = [], []
        for c in range(s.ldim):
            m = s.mask[c]
            x = np.zeros(s.eshape); r = rhs[c].copy(); p = np.zeros(s.eshape)
            rtz1 = 1.0
            it = 0
            while True:
                z = dinv * r * m
                rtz2 = rtz1
                rtz1 = s.glsc3(z, r, s.mult)
                rbn2 = math.sqrt(max(s.glsc3(r * r, s.mult, s.binv), 0.0) / s.vol)
                if rbn2 <= self.tol_v or it >= self.max_iter_v:
                    break
                beta = 0.0 if it == 0 else rtz1 / rtz2
                p = z + beta * p
                w = m * s.dssum(s.axhelm(p, self.h1, h2))
                rho = s.glsc3(w, p, s.mult)
                alpha = rtz1 / rho
                x += alpha * p
                r -= alpha * w
                it += 1
            out.append(x); its.append(it)
        self.iters_v.append(its)
        return np.stack(out)

    def _press_direct(self, g):
        s = self.s
        if self._lu_e is None:
            gk = ("E", _mesh_key(s), bool(self.ifvcor))
            if gk not in _LU_CACHE:
                E = s.e_sparse()
                if self.ifvcor:
                    E = E[1:, 1:]
                _LU_CACHE[gk] = spla.splu(E.tocsc())
            self._lu_e = _LU_CACHE[gk]
        gv = g.ravel()
        if self.ifvcor:
            x = np.concatenate(([0.0], self._lu_e.solve(gv[1:])))
            x -= x.mean()
        else:
            x = self._lu_e.solve(gv)
        return x.reshape(s.eshape2)

    def _press_pcg(self, g):
        """Jacobi-PCG on E (north-star's solver); stop on sqrt(sum r^2/bm2 / vol2) <= tol (absolute)
        [UPSTREAM navier1.f convprn norm]."""
        s = self.s
        if self._ediag is None:
            self._ediag = 1.0 / s.e_diag()
        dinv = self._ediag
        x = np.zeros(s.eshape2); r = g.copy(); p = np.zeros(s.eshape2)
        rtz1 = 1.0
        it = 0
        while True:
            z = dinv * r if self.pressure_precond is None else self.pressure_precond.apply(r)
            rtz2 = rtz1
            rtz1 = float(np.sum(z * r))
            rn = math.sqrt(float(np.sum(r * r / s.bm2)) / s.vol2)
            if rn <= self.tol_p or it >= self.max_iter_p:
                break
            beta = 0.0 if it == 0 else rtz1 / rtz2
            p = z + beta * p
            w = s.cdabdtp(p)
            alpha = rtz1 / float(np.sum(w * p))
            x += alpha * p
            r -= alpha * w
            it += 1
        if self.ifvcor:
            x -= x.mean()
        self.iters_p.append(it)
        return x

    # -------------------------------------------------------------- explicit terms
    def explicit_rhs(self, u, adjoint):
        s = self.s
        if adjoint == "nonlinear":
            # full Navier-Stokes: f = -B (u.grad)u [UPSTREAM navier1.f makef/advab -> convop]; C(u)u = advab_direct(u,u)/2
            f = -0.5 * s.advab_direct(u, u)
            if self.spng_str_dns != 0.0 and self.spng is not None:
                f = f + s.bm1 * self.spng_str_dns * self.spng * (self.spng_ref - u)
            return f
        f = -(s.advab_adjoint(u, self.ub) if adjoint else s.advab_direct(u, self.ub))
        if self.spng is not None:
            f = f - s.bm1 * self.spng * u
        return f

    def nonlinear_forward_map(self, v, p, nsteps, dt):
        """core/newton_krylov.f:336-378: f = phi_T(q) - q with the full (nonlinear) stepper; Dirichlet data live in q."""
        u, pr = self.linearized_map(v, p, nsteps, dt, adjoint="nonlinear")
        return u - v.reshape(u.shape), pr - p.reshape(pr.shape), u, pr

    # -------------------------------------------------------------- the maps
    def linearized_map(self, v, p, nsteps, dt, adjoint=False, record=None):
        """nsteps of the perturbation stepper starting from (v, p) with the order ramp restarted."""
        s = self.s
        d = s.ldim
        u = v.reshape((d,) + s.eshape).copy()
        pr = p.reshape(s.eshape2).copy()
        ulag = [np.zeros_like(u), np.zeros_like(u)]
        flag = [np.zeros_like(u), np.zeros_like(u)]
        plag = np.zeros_like(pr)
        for istep in range(1, nsteps + 1):
            k = min(istep, 3)
            bd, ab = BD[k], AB[k]
            h2 = bd[0] / dt
            f = self.explicit_rhs(u, adjoint)
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
            ulag = [u, ulag[0]]
            flag = [f, flag[0]]
            plag = pr
            u, pr = unew, pnew
            if record is not None:
                record(istep, u, pr)
        return u, pr

    def floquet_map(self, v, p, nsteps, dt, adjoint=False, orbit=None, pbase=None):
        """forward / adjoint_linearized_map with `ifbase` (uparam(1) = 3.11 / 3.21, core/matvec.f:187-236, 277-320): the base flow
        is advanced with the full stepper next to the perturbation; the perturbation's explicit terms of step n see U^{n-1} (the given
        base flow at step 1).  Returns (u, p, orbit); pass the returned orbit (list of U^1..U^nsteps) back in to replay it
        (`ifstorebase`)."""
        s, d = self.s, self.s.ldim
        ub0 = self.ub
        if orbit is None:
            orbit = []
            st = {"u": ub0.copy(), "pr": np.zeros(s.eshape2) if pbase is None else pbase.reshape(s.eshape2).copy()}
            self.linearized_map(st["u"], st["pr"], nsteps, dt, adjoint="nonlinear", record=lambda i, u, pr: orbit.append(u.copy()))
        try:
            def rec(i, u, pr):
                self.ub = orbit[i - 1]              # after step i the base flow is U^i (seen by step i + 1)
            self.ub = ub0
            out = self.linearized_map(v, p, nsteps, dt, adjoint=adjoint, record=rec)
        finally:
            self.ub = ub0
        return out[0], out[1], orbit

    def inner(self, a, b, bm1s=None):
        """krylov_inner_product (core/krylov_subspace.f:24-56) velocity part."""
        w = self.s.bm1 if bm1s is None else bm1s
        return float(sum(np.sum(a[c] * w * b[c]) for c in range(self.s.ldim)))
