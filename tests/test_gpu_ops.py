"""GPU parity tests, operator level: every CUDA kernel of the path against the CPU oracle on the same seeded
inputs, through the C ABI (ctypes).  Tolerances are for float64 arithmetic with different summation orders."""
import numpy as np
import pytest

from util import make_oracle, random_nodal, rel, small_cases, smooth_field

pytestmark = pytest.mark.gpu

CASES = small_cases()
TOL = 2e-12


@pytest.fixture(scope="module", params=list(CASES))
def ctx(request):
    from nekstab_b200.lib import NekStabB200
    c = CASES[request.param]
    s = make_oracle(c)
    g = NekStabB200(c)
    yield c, s, g
    g.close()


def test_geometry_fields(ctx):
    c, s, g = ctx
    assert rel(g.get_field("bm1"), s.bm1) < TOL
    assert rel(g.get_field("jacm1"), s.jac) < TOL
    assert rel(g.get_field("binvm1"), s.binv) < TOL
    assert rel(g.get_field("vmult"), s.mult) < TOL
    assert rel(g.get_field("bm2"), s.bm2) < TOL
    d = c.ldim
    order = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)] if d == 3 else [(0, 0), (1, 1), (0, 1)]
    gmax = np.abs(s.G).max()
    for q, (i, j) in enumerate(order):
        assert np.abs(g.get_field(f"g{q+1}") - s.G[i, j].ravel()).max() < 5e-13 * gmax, (i, j)
    assert abs(g.get_field("vol")[0] - s.vol) < 1e-12 * s.vol


def test_diagonals(ctx):
    c, s, g = ctx
    h1 = 1.0 / c.re
    assert rel(g.get_field("hdiagA"), (s.helm_diag(1.0, 0.0))) < 1e-11
    assert rel(g.get_field("ediag"), s.e_diag()) < 1e-11


def test_ifvcor(ctx):
    c, s, g = ctx
    assert bool(g.get_field("ifvcor")[0]) == bool(c.ifvcor)
    # the numerical test agrees with the boundary-condition rule on these meshes (affine when all-Dirichlet)
    e1 = s.cdabdtp(np.ones(s.eshape2))
    assert bool(np.linalg.norm(e1) < 1e-9 * np.linalg.norm(s.e_diag())) == bool(c.ifvcor)


def test_axhelm(ctx):
    c, s, g = ctx
    u = random_nodal(c, 1, masked=False)[0]
    for h1, h2 in ((1.0, 0.0), (0.02, 150.0)):
        w = g.op_axhelm(u, h1, h2)
        assert rel(w, s.axhelm(u.reshape(s.eshape), h1, h2)) < TOL


def test_dssum_and_glsc3(ctx):
    c, s, g = ctx
    rng = np.random.default_rng(3)
    u = rng.standard_normal(c.n)
    assert rel(g.op_dssum(u), s.dssum(u.reshape(s.eshape))) < TOL
    a, b = rng.standard_normal(c.n), rng.standard_normal(c.n)
    ref = s.glsc3(a.reshape(s.eshape), s.bm1, b.reshape(s.eshape))
    assert abs(g.op_glsc3(a, s.bm1.ravel(), b) - ref) < 1e-12 * max(abs(ref), np.sqrt(c.n))


def test_gradt_div_E(ctx):
    c, s, g = ctx
    rng = np.random.default_rng(4)
    p = rng.standard_normal(c.nel * s.lx2 ** c.ldim)
    assert rel(g.op_opgradt(p), s.opgradt(p.reshape(s.eshape2))) < TOL
    u = random_nodal(c, 5, masked=False)
    assert rel(g.op_opdiv(u), s.opdiv(u.reshape((c.ldim,) + s.eshape))) < TOL
    assert rel(g.op_cdabdtp(p), s.cdabdtp(p.reshape(s.eshape2))) < 5e-12
    # adjointness: <D u, p> == <u, D^T p>
    lhs = float(np.dot(g.op_opdiv(u), p)); rhs = float(np.sum(u.reshape(c.ldim, -1) * g.op_opgradt(p)))
    assert abs(lhs - rhs) < 1e-11 * max(abs(lhs), 1.0)


def test_advection(ctx):
    c, s, g = ctx
    up = smooth_field(c, 7)
    ub = c.ubase.reshape((c.ldim,) + s.eshape)
    upr = up.reshape((c.ldim,) + s.eshape)
    sp = s.bm1 * c.spng_fun.reshape(s.eshape)
    ref_d = s.advab_direct(upr, ub) + sp * upr
    ref_a = s.advab_adjoint(upr, ub) + sp * upr
    assert rel(g.op_advab(0, up), ref_d) < TOL
    assert rel(g.op_advab(1, up), ref_a) < TOL


def test_cfl(ctx):
    c, s, g = ctx
    ub = c.ubase.reshape((c.ldim,) + s.eshape)
    assert abs(g.op_cfl(c.ubase, 1.0) - s.cfl_sum(ub)) < 1e-12 * s.cfl_sum(ub)
    from oracle.stepper import prepare_linearized_solver
    dt, ns, ct = g.prepare_linearized_solver(c.end_time)
    dto, nso, cto = prepare_linearized_solver(s, ub, c.end_time)
    assert ns == nso and abs(dt - dto) < 1e-15 and abs(ct - cto) < 1e-12 * cto


def test_helmholtz_solve(ctx):
    c, s, g = ctx
    from oracle.stepper import LinearizedStepper
    st = LinearizedStepper(s, c.ubase, c.re, None, solver="direct")
    h1, h2 = 1.0 / c.re, 11.0 / 6.0 / 0.01
    r = smooth_field(c, 9, masked=False).reshape((c.ldim,) + s.eshape) * s.bm1
    g.set_params(h1, 1.0, 1e-13, 1e-13, 2000, 50000)
    x, it = g.op_hmholtz(r, h1, h2)
    rhs = np.stack([s.mask[k] * s.dssum(r[k]) for k in range(c.ldim)])
    ref = st._helm_direct(rhs, h2)
    assert rel(x, ref) < 1e-10
    # same algorithm, same stopping rule => same iteration count as the oracle's PCG
    st2 = LinearizedStepper(s, c.ubase, c.re, None, tol_v=1e-13, solver="pcg", max_iter_v=2000)
    xo = st2._helm_pcg(rhs, h2)
    assert rel(x, xo) < 1e-10
    assert abs(it - sum(st2.iters_v[-1])) <= 2 * c.ldim


def test_pressure_solve(ctx):
    c, s, g = ctx
    from oracle.stepper import LinearizedStepper
    st = LinearizedStepper(s, c.ubase, c.re, None, solver="direct", ifvcor=c.ifvcor)
    u = smooth_field(c, 11)
    gg = -s.opdiv(u.reshape((c.ldim,) + s.eshape))
    g.set_params(1.0 / c.re, 1.0, 1e-13, 1e-12, 2000, 50000)
    phi, it = g.op_esolver(gg)
    if st.ifvcor:
        gg = gg - gg.mean()
    ref = st._press_direct(gg)
    assert rel(phi, ref) < 1e-8, it
    st2 = LinearizedStepper(s, c.ubase, c.re, None, tol_p=1e-12, solver="pcg", ifvcor=c.ifvcor)
    xo = st2._press_pcg(gg.copy())
    assert rel(phi, xo) < 1e-8
    assert abs(it - st2.iters_p[-1]) <= max(3, 0.02 * it)
