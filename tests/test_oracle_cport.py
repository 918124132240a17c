"""The C / OpenMP restatement of the 3-D step (oracle/cport.c, used as the CPU baseline of bench.py) against the numpy
oracle it restates: every operator to rounding, the solvers to the same iteration counts, whole steps to 1e-10."""
import numpy as np
import pytest

from util import make_oracle, rel, small_cases, smooth_field

CASES = small_cases()
NAMES = ["box3d_n8_outflow", "box3d_n6_dirichlet", "box3d_n4_outflow"]


@pytest.fixture(scope="module", params=NAMES)
def ctx(request):
    from oracle import cport, pmg
    c = CASES[request.param]
    s = make_oracle(c)
    M = pmg.PMG(s, nagg=3, ifvcor=bool(c.ifvcor))
    return c, s, M, cport.CPort(s, M)


def test_operators(ctx):
    c, s, M, cp = ctx
    rng = np.random.default_rng(0)
    u = rng.standard_normal(s.eshape)
    assert rel(cp.axhelm(u, 0.02, 150.0), s.axhelm(u, 0.02, 150.0)) < 1e-13
    assert rel(cp.dssum(u), s.dssum(u)) < 1e-14
    p = rng.standard_normal(s.eshape2)
    assert rel(cp.opgradt(p), s.opgradt(p)) < 1e-13
    v = rng.standard_normal((3,) + s.eshape)
    assert rel(cp.opdiv(v), s.opdiv(v)) < 1e-13
    assert rel(cp.cdabdtp(p), s.cdabdtp(p)) < 1e-12
    ub = c.ubase.reshape((3,) + s.eshape)
    spng = c.spng_fun.reshape(s.eshape)
    ref = -s.advab_direct(v, ub) - s.bm1 * spng * v
    assert rel(cp.advab_direct(v, ub, spng), ref) < 1e-12
    assert rel(cp.pmg_apply(p), M.apply(p)) < 1e-12


def test_solvers_match_oracle_iteration_for_iteration(ctx):
    c, s, M, cp = ctx
    from oracle import pmg
    from oracle.stepper import LinearizedStepper
    g = -s.opdiv(smooth_field(c, 11).reshape((3,) + s.eshape))
    if c.ifvcor:
        g = g - g.mean()
    nrm = lambda r: float(np.sqrt(np.sum(r * r / s.bm2) / s.vol2))
    xo, ito = pmg.pcg(s.cdabdtp, M.apply, g.copy(), 1e-11, norm=nrm)
    x, it = cp.pressure_pcg(g, 1e-11, 20000)
    assert abs(it - ito) <= 2 and rel(x, xo) < 1e-7
    h1, h2 = 1.0 / c.re, 11.0 / 6.0 / 0.01
    st = LinearizedStepper(s, c.ubase, c.re, None, tol_v=1e-12, solver="pcg", max_iter_v=2000)
    r = smooth_field(c, 9, masked=False).reshape((3,) + s.eshape) * s.bm1
    rhs = np.stack([s.mask[k] * s.dssum(r[k]) for k in range(3)])
    ref = st._helm_pcg(rhs, h2)
    out, its = cp.helmholtz_pcg(rhs, h1, h2, 1e-12, 2000)
    assert rel(out, ref) < 1e-9 and max(abs(a - b) for a, b in zip(its, st.iters_v[-1])) <= 2


@pytest.mark.parametrize("use_pmg", [False, True])
def test_whole_step_matches_numpy_stepper(ctx, use_pmg):
    c, s, M, cp = ctx
    from oracle import cport
    from oracle.stepper import LinearizedStepper
    nsteps, dt = 4, 2.0e-3
    st = LinearizedStepper(s, c.ubase, c.re, c.spng_fun, tol_v=1e-12, tol_p=1e-12, solver="pcg", ifvcor=c.ifvcor,
                           pressure_precond=M if use_pmg else None)
    cs = cport.CStepper(s, c.ubase, c.re, c.spng_fun, tol_v=1e-12, tol_p=1e-12, ifvcor=c.ifvcor, pmg=M if use_pmg else None)
    v0 = smooth_field(c, 21).reshape((3,) + s.eshape)
    p0 = 0.1 * np.random.default_rng(5).standard_normal(s.eshape2)
    vo, po = st.linearized_map(v0, p0, nsteps, dt)
    v, p = cs.linearized_map(v0, p0, nsteps, dt)
    assert rel(v, vo) < 1e-10
    assert rel(p, po) < 1e-7
