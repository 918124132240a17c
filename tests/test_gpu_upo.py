"""Newton-GMRES for unstable periodic orbits on the GPU (uparam(1) = 2.1): the time component of the Krylov
vectors (core/krylov_subspace.f:14, 47-50), nonlinear_forward_map with orbit storage (core/newton_krylov.f:336-378), the border
vectors of compute_bvec (core/matvec.f:435-475), newton_linearized_map's UPO branch (core/matvec.f:407-419) and newton_krylov with the
period as an unknown (core/newton_krylov.f:63-67, 122) -- against the oracle's restatement (oracle/upo.py)."""
import os

import numpy as np
import pytest

from util import GOLD, make_oracle, rel, small_cases, smooth_field

pytestmark = pytest.mark.gpu


def _guess(c, s, amp, seed=3):
    u = c.ubase.reshape((c.ldim,) + s.eshape).copy()
    return u + amp * smooth_field(c, seed).reshape(u.shape)


def _prepare(ctx, slot, end_time):
    import ctypes as C
    from nekstab_b200.lib import _ck
    dt, ns, ct = C.c_double(), C.c_int(), C.c_double()
    _ck(ctx.lib.nsb_prepare_solver_from_slot(slot, end_time, 0.5, C.byref(dt), C.byref(ns), C.byref(ct)))
    return dt.value, ns.value


@pytest.mark.parametrize("name", ["box2d_n6_outflow", "box3d_n6_dirichlet"])
def test_upo_newton_matvec_against_oracle(name):
    from nekstab_b200 import lib
    from oracle import krylov
    from oracle.upo import UPOMaps
    c = small_cases()[name]
    c.spng_fun = None
    s = make_oracle(c)
    g = lib.NekStabB200(c)
    try:
        T = 0.1
        g.set_params(1.0 / c.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        g.vec_alloc(6)
        g.set_upo(1)
        q0, p0 = _guess(c, s, 0.05), np.zeros(s.eshape2)
        g.vec_upload(0, q0, p0)
        g.vec_set_time(0, T)
        dt, ns = _prepare(g, 0, T)
        m = UPOMaps(s, c.re, s.bm1, ifvcor=c.ifvcor, solver="direct")
        fo = m.nonlinear_map((q0, p0, T))
        assert ns == m.state["ns"] and abs(dt - m.state["dt"]) < 1e-15
        g.nonlinear_forward_map(0, 1)
        f, fp = g.vec_download(1)
        assert rel(f, fo[0]) < 1e-9 and g.vec_get_time(1) == 0.0
        for k in (1, ns):                                            # the stored orbit uor, vor, wor (core/newton_krylov.f:364-368)
            assert rel(g.get_orbit(k), m.state["orbit"][k - 1]) < 1e-11
        # newton_linearized_map, UPO branch: both border terms are exercised by a vector with a time component
        mv = m.linearized_map_factory(None)
        x0, xt = smooth_field(c, 8).reshape(q0.shape), 0.37
        g.vec_upload(2, x0, p0)
        g.vec_set_time(2, xt)
        g.matvec(lib.NEWTON, 2, 3)
        y, yp = g.vec_download(3)
        yo = mv((x0, p0, xt))
        assert rel(y, yo[0]) < 1e-9, rel(y, yo[0])
        assert abs(g.vec_get_time(3) - yo[2]) < 1e-9 * max(1.0, abs(yo[2])), (g.vec_get_time(3), yo[2])
        # ... and they matter: without the time component the answer differs
        g.vec_set_time(2, 0.0)
        g.matvec(lib.NEWTON, 2, 4)
        y2, _ = g.vec_download(4)
        assert rel(y2, yo[0]) > 1e-6
        # krylov_inner_product / norm / cmult / add2 / copy with the time component (core/krylov_subspace.f:47-50, 107, 117, 147)
        w = s.bm1
        a = (y.reshape(q0.shape), yp.reshape(p0.shape), g.vec_get_time(3))
        assert abs(g.inner_product(3, 3) - krylov.inner(a, a, w)) < 1e-12 * krylov.inner(a, a, w)
        g.vec_set_time(2, xt)
        b = (x0, p0, xt)
        assert abs(g.inner_product(3, 2) - krylov.inner(a, b, w)) < 1e-11 * abs(krylov.inner(a, a, w))
        g.vec_copy(5, 3); g.vec_cmult(5, 2.0); g.vec_add2(5, 2)
        assert abs(g.vec_get_time(5) - (2.0 * a[2] + xt)) < 1e-14 * max(1.0, abs(a[2]))
        nrm = g.norm(5)
        ref = krylov.axpy(krylov.scale(a, 2.0), 1.0, b)
        assert abs(nrm - np.sqrt(krylov.inner(ref, ref, w))) < 1e-11 * nrm
        # orthonormalisation carries the time component: f <- f - Q (Q^T f) in the extended inner product
        g.normalize(2)
        h = g.orthonormalize(1, 2, 3)
        assert abs(g.inner_product(2, 3)) < 1e-12 and abs(g.inner_product(3, 3) - 1.0) < 1e-12
    finally:
        g.close()


def test_upo_newton_krylov_iteration_against_oracle():
    """One full Newton iteration with the period as an unknown: nonlinear map, GMRES on the bordered operator, update of q and q%time.
    (No small closed box has a periodic orbit to converge to; the conditioning of this single iteration is 1e-12 -> 2e-12 in the period,
    measured with the oracle's direct vs PCG solvers.)"""
    from nekstab_b200 import lib
    from oracle import krylov
    from oracle.upo import UPOMaps
    c = small_cases()["box2d_n6_outflow"]
    c.spng_fun = None
    s = make_oracle(c)
    g = lib.NekStabB200(c)
    try:
        T, k, tol = 0.2, 10, 1e-18
        g.set_params(1.0 / c.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        g.vec_alloc(k + 6)
        g.set_upo(1)
        q0, p0 = _guess(c, s, 0.02), np.zeros(s.eshape2)
        g.vec_upload(0, q0, p0)
        ok, it, res, hist, calls = g.newton_krylov(0, 1, 2, 3, 4, k, T, tol, maxiter_newton=1, maxiter_gmres=4)
        qg, _ = g.vec_download(0)
        period = g.vec_get_time(0)
        m = UPOMaps(s, c.re, s.bm1, ifvcor=c.ifvcor, solver="direct")
        qo, ito, histo = krylov.newton_krylov(m.nonlinear_map, m.linearized_map_factory, (q0, p0, T), k, tol, s.bm1,
                                              maxiter_newton=1, maxiter_gmres=4)
        print("UPO Newton iteration: residual", hist, histo, "period", period, qo[2])
        assert not ok
        assert abs(hist[0] - histo[0]) < 1e-9 * histo[0]
        assert abs(period - qo[2]) < 1e-8 and abs(period - T) > 1e-3          # 0.2625286532 ; the period moved
        assert rel(qg, qo[0]) < 1e-8
    finally:
        g.close()


def test_upo_residual_of_the_shipped_periodic_base_flow():
    """examples/cylinder/stability/direct_Floquet/BF_1cyl0.f00001 is the product of the reference's Newton-UPO solver (period 7.9213 in
    the file header, 795 steps): under the GPU's full Navier-Stokes stepper (no sponge, baseflow/newton_upo/1cyl.par) it is periodic to
    |phi_T(U) - U|^2 = 9.3e-12 < 1e-11, i.e. the reference's own Newton iteration would stop on it -- KAT-UPO."""
    from nekstab_b200 import cases, lib, restart
    g = np.load(os.path.join(GOLD, "cyl.npz"))
    u = np.load(os.path.join(GOLD, "cyl_upo.npz"))
    c = cases.cylinder_case(g, sponge=False)
    lx = int(u["lx1"])
    U = u["U"].reshape(-1, 2, lx * lx).transpose(1, 0, 2).astype(np.float64)
    T = float(u["time"])
    ctx = lib.NekStabB200(c)
    try:
        ctx.set_params(1.0 / c.re, 1.0, 1e-11, 1e-11, 2000, 100000)
        ctx.set_pressure_preconditioner(1, 64)
        ctx.vec_alloc(3)
        ctx.set_upo(1)
        p2 = restart.pressure_to_mesh2(u["P"].reshape(c.nel, -1).astype(np.float64), c.lx1, 2)
        ctx.vec_upload(0, U, p2)
        ctx.vec_set_time(0, T)
        dt, ns = _prepare(ctx, 0, T)
        assert ns == 795
        ctx.nonlinear_forward_map(0, 1)
        res, nrm = ctx.norm(1), ctx.norm(0)
        print("UPO residual of the shipped periodic base flow: |phi_T(U) - U| = %.3e, |U| = %.4f, period %.4f, %d steps" % (res, nrm, T, ns))
        assert abs(res - 3.05e-6) < 0.1e-6, res                      # measured 3.051e-06 (|U| = 46.8123)
        assert res ** 2 < 1e-11                                      # the reference's Newton exit test (core/newton_krylov.f:109, 1cyl.par tolerances)
        # the bordered Newton operator runs on the stored orbit
        ctx.vec_upload(2, cases.add_noise(c), None)
        ctx.vec_set_time(2, 0.1)
        ctx.matvec(lib.NEWTON, 2, 1)
        assert np.isfinite(ctx.norm(1)) and np.isfinite(ctx.vec_get_time(1))
    finally:
        ctx.close()
