"""Config-2 machinery on the GPU: nonlinear_forward_map (core/newton_krylov.f:336-378), the fixed-point KAT of the shipped
Re=50 base flow, and newton_krylov (core/newton_krylov.f:5-168) against the oracle's restatement."""
import os

import numpy as np
import pytest

from nekstab_b200 import cases
from util import GOLD, make_oracle, rel, small_cases, smooth_field

pytestmark = pytest.mark.gpu


def _steady_inflow_guess(c, s, amp=0.0, seed=3):
    """base flow of the test box (non-zero Dirichlet inflow data) + a masked smooth perturbation"""
    u = c.ubase.reshape((c.ldim,) + s.eshape).copy()
    if amp:
        u = u + amp * smooth_field(c, seed).reshape(u.shape)
    return u


@pytest.mark.parametrize("name", ["box2d_n6_outflow", "box3d_n6_dirichlet"])
def test_nonlinear_forward_map(name):
    from nekstab_b200 import lib
    from oracle.stepper import LinearizedStepper
    c = small_cases()[name]
    s = make_oracle(c)
    g = lib.NekStabB200(c)
    try:
        nsteps, dt = 6, 2.0e-3
        g.set_params(1.0 / c.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        g.set_timestep(dt, nsteps)
        g.vec_alloc(3)
        st = LinearizedStepper(s, c.ubase, c.re, None, solver="direct", ifvcor=c.ifvcor)
        q = _steady_inflow_guess(c, s, amp=0.05)
        p0 = np.zeros(s.eshape2)
        g.vec_upload(0, q, p0)
        g.nonlinear_forward_map(0, 1)
        f, fp = g.vec_download(1)
        fo, fpo, uo, po = st.nonlinear_forward_map(q, p0, nsteps, dt)
        assert rel(f, fo) < 1e-9, rel(f, fo)
        # ubase <- q (core/newton_krylov.f:374-375): the Newton-mode matvec now linearises about q
        st2 = LinearizedStepper(s, q, c.re, c.spng_fun, solver="direct", ifvcor=c.ifvcor)   # the perturbation sponge stays on
        v0 = smooth_field(c, 8).reshape(q.shape)
        g.vec_upload(0, v0, p0)
        g.matvec(lib.NEWTON, 0, 2)
        v, _ = g.vec_download(2)
        vo, _ = st2.linearized_map(v0, p0, nsteps, dt)
        assert rel(v, vo - v0) < 1e-8
    finally:
        g.close()


def test_fixed_point_kat_cylinder_baseflow():
    """The shipped Re=50 base flow is a fixed point of the nonlinear map: ||phi_T(U) - U|| = 3.1e-6 (||U|| = 46.15), i.e.
    residual^2 = 9.6e-12 just under Newton's exit test 1e-11 (oracle: tests/test_oracle_fixtures.py)."""
    from nekstab_b200 import lib
    from oracle import sem as osem
    g = np.load(os.path.join(GOLD, "cyl.npz"))
    c = cases.cylinder_case(g, sponge=False)          # the Newton case has no sponge (baseflow/newton/1cyl.par)
    ctx = lib.NekStabB200(c)
    try:
        ctx.set_params(1.0 / c.re, 1.0, 1e-11, 1e-11, 2000, 100000)   # baseflow/newton/1cyl.par:31,36
        ctx.vec_alloc(2)
        z, _ = osem.gll(6); zg, _ = osem.gl(4)
        J12 = osem.interp(zg, z)
        p2 = np.einsum("ai,bj,eji->eba", J12, J12, g["P"].astype(float))
        ctx.vec_upload(0, c.ubase, p2)
        dt, nsteps, ct = C_prepare(ctx, 0, 1.0)
        assert nsteps == 100
        ctx.nonlinear_forward_map(0, 1)
        res = ctx.norm(1)
        assert abs(ctx.norm(0) - 46.1512) < 1e-3
        assert abs(res - 3.096e-6) < 0.1e-6, res
        assert res ** 2 < 1e-11                                      # Newton would stop here (core/newton_krylov.f:109)
    finally:
        ctx.close()


def C_prepare(ctx, slot, end_time):
    import ctypes as C
    dt, ns, ct = C.c_double(), C.c_int(), C.c_double()
    from nekstab_b200.lib import _ck
    _ck(ctx.lib.nsb_prepare_solver_from_slot(slot, end_time, 0.5, C.byref(dt), C.byref(ns), C.byref(ct)))
    return dt.value, ns.value, ct.value


def test_newton_krylov_driver():
    from nekstab_b200 import lib
    from oracle import krylov
    from oracle.stepper import LinearizedStepper, prepare_linearized_solver
    c = small_cases()["box2d_n6_outflow"]
    c.spng_fun = None
    s = make_oracle(c)
    g = lib.NekStabB200(c)
    try:
        T, k, tol = 0.05, 12, 1e-18
        g.set_params(1.0 / c.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        g.vec_alloc(k + 6)
        q0 = _steady_inflow_guess(c, s, amp=0.02)
        p0 = np.zeros(s.eshape2)
        g.vec_upload(0, q0, p0)
        ok, it, res, hist, calls = g.newton_krylov(0, 1, 2, 3, 4, k, T, tol, maxiter_newton=8, maxiter_gmres=6)
        qg, pg = g.vec_download(0)
        # oracle Newton with the same algorithm
        w = s.bm1

        def nl(q):
            dt, ns, _ = prepare_linearized_solver(s, q[0], T)
            st = LinearizedStepper(s, q[0], c.re, None, solver="direct", ifvcor=c.ifvcor)
            fv, fp, _, _ = st.nonlinear_forward_map(q[0], q[1], ns, dt)
            return (fv, fp)

        def lin(q):
            dt, ns, _ = prepare_linearized_solver(s, q[0], T)
            st = LinearizedStepper(s, q[0], c.re, None, solver="direct", ifvcor=c.ifvcor)
            return lambda x: (lambda y: (y[0] - x[0], y[1] - x[1]))(st.linearized_map(x[0], x[1], ns, dt))

        qo, ito, histo = krylov.newton_krylov(nl, lin, (q0, p0), k, tol, w, maxiter_newton=8, maxiter_gmres=6)
        assert ok and res < tol
        assert it == ito
        assert hist[0] > 1e-8 and hist[-1] < tol                     # converged from a genuinely perturbed state
        assert np.allclose(np.log10(hist[:2]), np.log10(histo[:2]), atol=0.05)
        assert rel(qg, qo[0]) < 1e-7
    finally:
        g.close()
