"""GPU path against the reference's shipped fixtures (the same KATs that pin the oracle), at the
reference's full sizes: BFS transient-growth pair (cfg 4: 1670 elements, 172 steps) and the cylinder leading eigenpair
(cfg 1: 1996 elements, 100 steps).  The fixtures are float32 fields produced at tolerances 1e-7..1e-9."""
import os

import numpy as np
import pytest

from nekstab_b200 import cases
from util import GOLD

pytestmark = pytest.mark.gpu


def _inner(a, b, w):
    return float(sum(np.sum(a[d] * w * b[d]) for d in range(a.shape[0])))


def test_kat_tg_bfs_on_gpu():
    from nekstab_b200 import lib
    from oracle import sem as osem
    g = np.load(os.path.join(GOLD, "bfs.npz"))
    c = cases.bfs_case(g)
    ctx = lib.NekStabB200(c)
    try:
        ctx.set_params(1.0 / c.re, 1.0, 1e-11, 1e-10, 2000, 100000)
        dt, nsteps, ctarg = ctx.prepare_linearized_solver(c.end_time)
        assert nsteps == 172
        bm1 = ctx.get_field("bm1").reshape(c.nel, -1)
        bm1s = ctx.get_field("bm1s").reshape(c.nel, -1)
        z, _ = osem.gll(6); zg, _ = osem.gl(4)
        J12 = osem.interp(zg, z)
        to_m2 = lambda p: np.einsum("ai,bj,eji->eba", J12, J12, p)   # file pressures live on mesh 1
        pre = g["pRe_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
        ore = g["ore_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
        ctx.vec_alloc(3)
        ctx.vec_upload(0, pre, to_m2(g["pRe_P"].astype(float)))
        ctx.matvec(lib.DIRECT, 0, 1)
        u, p = ctx.vec_download(1)
        u = u.reshape(2, c.nel, -1)
        err = np.sqrt(_inner(u - ore, u - ore, bm1) / _inner(ore, ore, bm1))
        assert err < 3e-7, err                                     # oracle: 1.3e-7 (float32 fixture)
        assert abs(_inner(u, u, bm1s) - 3.23700) < 1e-5            # optimal energy gain G(T=1)
        # M^T M p = lambda p (core/matvec.f:332-349)
        ctx.matvec(lib.ADJOINT, 1, 2)        # = transient_growth_map's second half (core/matvec.f:344)
        ua, _ = ctx.vec_download(2)
        ua = ua.reshape(2, c.nel, -1)
        lam = _inner(ua, pre, bm1s) / _inner(pre, pre, bm1s)
        res = ua - lam * pre
        assert abs(lam - 3.2370) < 1e-4
        assert np.sqrt(_inner(res, res, bm1s) / _inner(ua, ua, bm1s)) < 3e-7
        st = ctx.stats()
        assert st["steps"] == 2 * 172
    finally:
        ctx.close()


def test_kat_eig_cylinder_on_gpu():
    from nekstab_b200 import lib
    from oracle import sem as osem
    g = np.load(os.path.join(GOLD, "cyl.npz"))
    c = cases.cylinder_case(g)
    ctx = lib.NekStabB200(c)
    try:
        ctx.set_params(1.0 / c.re, 1.0, 1e-11, 1e-10, 2000, 100000)
        dt, nsteps, ctarg = ctx.prepare_linearized_solver(c.end_time)
        assert nsteps == 100 and abs(dt - 0.01) < 1e-15 and abs(ctarg - 49.72) < 0.01
        bm1s = ctx.get_field("bm1s").reshape(c.nel, -1)
        z, _ = osem.gll(6); zg, _ = osem.gl(4)
        J12 = osem.interp(zg, z)
        to_m2 = lambda p: np.einsum("ai,bj,eji->eba", J12, J12, p)
        dre = g["dRe_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
        dim = g["dIm_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
        ctx.vec_alloc(4)
        ctx.vec_upload(0, dre, to_m2(g["dRe_P"].astype(float)))
        ctx.vec_upload(1, dim, to_m2(g["dIm_P"].astype(float)))
        assert abs(ctx.inner_product(0, 0) + ctx.inner_product(1, 1) - 1.0) < 5e-9     # KAT-norm on the device
        ctx.matvec(lib.DIRECT, 0, 2)
        ctx.matvec(lib.DIRECT, 1, 3)
        ur = ctx.vec_download(2)[0].reshape(2, c.nel, -1)
        ui = ctx.vec_download(3)[0].reshape(2, c.nel, -1)
        mu = g["Spectre_Hd"][0, 0] + 1j * g["Spectre_Hd"][0, 1]
        q, Mq = dre + 1j * dim, ur + 1j * ui
        ip = lambda a, b: sum(np.sum(np.conj(a[d]) * bm1s * b[d]) for d in range(2))
        ray = ip(q, Mq) / ip(q, q)
        assert abs(ray - mu) < 2e-7                                # all 7 printed digits of Spectre_Hd.dat:1
        r = Mq - mu * q
        assert np.sqrt(abs(ip(r, r)) / abs(ip(Mq, Mq))) < 3e-6
    finally:
        ctx.close()


def test_krylov_schur_and_gmres_drivers_on_gpu():
    """nsb_krylov_schur (Arnoldi + eig + Schur condensation + basis rotation) and nsb_ts_gmres against the oracle's
    restatement of the same drivers run on the oracle's own matvec."""
    from nekstab_b200 import lib
    from oracle import krylov
    from oracle.stepper import LinearizedStepper
    from util import make_oracle, small_cases, smooth_field
    c = small_cases()["box2d_n6_outflow"]
    s = make_oracle(c)
    ctx = lib.NekStabB200(c)
    try:
        nsteps, dt, k = 5, 4.0e-3, 10
        ctx.set_params(1.0 / c.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        ctx.set_timestep(dt, nsteps)
        ctx.vec_alloc(k + 6)
        bm1s = s.bm1 * (c.spng_fun.reshape(s.eshape) == 0)
        st = LinearizedStepper(s, c.ubase, c.re, c.spng_fun, solver="direct", ifvcor=c.ifvcor)
        v0 = smooth_field(c, 44).reshape((c.ldim,) + s.eshape)
        p0 = np.zeros(s.eshape2)
        nrm = np.sqrt(sum(np.sum(v0[d] ** 2 * bm1s) for d in range(c.ldim)))
        ctx.vec_upload(0, v0, p0)
        ctx.normalize(0)
        vals, res, V, ncv, scnt = ctx.krylov_schur(lib.DIRECT, k, 2, eigen_tol=1e-9, schur_del=0.1, seed_slot=0, max_restarts=6)
        mv = lambda q: st.linearized_map(q[0], q[1], nsteps, dt)
        vo, veco, reso, Qo, Ho, cnto, scnto = krylov.krylov_schur(mv, (v0 / nrm, p0), k, 2, bm1s, eigen_tol=1e-9, max_restarts=6)
        assert scnt == scnto and ncv == cnto
        nconv = max(cnto, 2)
        assert np.abs(np.sort_complex(vals[:nconv]) - np.sort_complex(vo[:nconv])).max() < 1e-8
        # GMRES on (exp(TL) - I) x = b
        rhs = smooth_field(c, 45).reshape((c.ldim,) + s.eshape)
        ctx.vec_upload(k + 2, rhs, p0)
        calls, r2 = ctx.ts_gmres(lib.NEWTON, k + 2, k + 3, 0, k + 4, 2, k, 1e-16)
        x, _ = ctx.vec_download(k + 3)
        jac = lambda q: (lambda f: (f[0] - q[0], f[1] - q[1]))(st.linearized_map(q[0], q[1], nsteps, dt))
        solo, callso, r2o = krylov.ts_gmres(jac, (rhs, p0), 2, k, 1e-16, bm1s)
        # same restarted-GMRES trajectory (2 cycles of k steps): same residual history end point and same iterate
        assert abs(r2 - r2o) < 1e-6 * r2o and calls == callso
        num = np.sqrt(sum(np.sum((x.reshape(solo[0].shape)[d] - solo[0][d]) ** 2 * bm1s) for d in range(c.ldim)))
        den = np.sqrt(sum(np.sum(solo[0][d] ** 2 * bm1s) for d in range(c.ldim)))
        assert num / den < 1e-6, num / den
    finally:
        ctx.close()
