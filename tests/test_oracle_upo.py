"""The oracle's UPO-Newton restatement (oracle/upo.py) checked against what it must be mathematically: the bordered operator of
newton_linearized_map (core/matvec.f:397-419) is the Jacobian of F(q, T) = phi_T(q) - q, with compute_bvec's first-order time
derivative (core/matvec.f:435-475) as its T column."""
import numpy as np

from util import make_oracle, small_cases, smooth_field


def _setup():
    from oracle.upo import UPOMaps
    c = small_cases()["box2d_n6_outflow"]
    c.spng_fun = None
    s = make_oracle(c)
    u = c.ubase.reshape((c.ldim,) + s.eshape).copy()
    q0 = u + 0.05 * smooth_field(c, 3).reshape(u.shape)
    return c, s, q0, np.zeros(s.eshape2), UPOMaps(s, c.re, s.bm1, ifvcor=c.ifvcor, solver="direct")


def test_bordered_operator_is_the_jacobian_of_the_upo_map():
    from oracle import krylov
    c, s, q0, p0, m = _setup()
    T, eps = 0.1, 1e-6
    f0 = m.nonlinear_map((q0, p0, T))
    mv = m.linearized_map_factory(None)
    x = s.mask * smooth_field(c, 8).reshape(q0.shape)                  # a perturbation that keeps the Dirichlet data
    w = s.bm1
    # velocity block: (Phi_T - I) x by finite differences of the nonlinear map (same T, hence same dt and nsteps)
    y = mv((x, p0, 0.0))
    m2 = type(m)(s, c.re, w, ifvcor=c.ifvcor, solver="direct")
    f1 = m2.nonlinear_map((q0 + eps * x, p0, T))
    fd = (f1[0] - f0[0]) / eps
    err = np.sqrt(krylov.inner((fd - y[0],), (fd - y[0],), w) / krylov.inner((y[0],), (y[0],), w))
    assert err < 1e-4, err
    # T column: d phi_T(q) / dT = the time derivative at the end of the orbit, approximated by bvec(fc) to first order in dt
    f2 = m2.nonlinear_map((q0, p0, T * (1 + 1e-4)))
    assert m2.state["ns"] == m.state["ns"]
    dfdT = (f2[0] - f0[0]) / (T * 1e-4)
    yt = mv((np.zeros_like(x), p0, 1.0))[0]                            # = bvec(fc)
    cos = krylov.inner((dfdT,), (yt,), w) / np.sqrt(krylov.inner((dfdT,), (dfdT,), w) * krylov.inner((yt,), (yt,), w))
    ratio = np.sqrt(krylov.inner((yt,), (yt,), w) / krylov.inner((dfdT,), (dfdT,), w))
    assert cos > 0.98 and abs(ratio - 1.0) < 0.15, (cos, ratio)
    # time row: f%time = <bvec(ic), x>; bvec(ic) is the time derivative at the start of the orbit
    bic = m.state["bic"][0]
    assert abs(y[2] - krylov.inner((bic,), (x,), w)) < 1e-12 * abs(y[2])


def test_krylov_algebra_with_a_time_component():
    from oracle import krylov
    c, s, q0, p0, m = _setup()
    w = s.bm1
    a, b = (q0, p0, 0.3), (2.0 * q0, p0, -0.5)
    assert abs(krylov.inner(a, b, w) - (2.0 * krylov.inner((q0,), (q0,), w) - 0.15)) < 1e-12 * krylov.inner((q0,), (q0,), w)
    r = krylov.axpy(krylov.scale(a, 2.0), -1.0, b)
    assert abs(r[2] - 1.1) < 1e-15 and np.allclose(r[0], 0.0)
    H = np.zeros((2, 1))
    qn = krylov.scale(a, 1.0 / np.sqrt(krylov.inner(a, a, w)))
    f = krylov.update_hessenberg_matrix(H, (smooth_field(c, 5).reshape(q0.shape), p0, 0.7), [qn], 1, w)
    assert abs(krylov.inner(f, qn, w)) < 1e-13 and abs(krylov.inner(f, f, w) - 1.0) < 1e-13
