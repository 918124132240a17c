"""CPU tests of the oracle restatement of the pressure preconditioner (oracle/pmg.py): what CG needs (symmetric,
positive definite), that it leaves the converged solution alone, and that it cuts the iteration count."""
import numpy as np
import pytest

from util import GOLD, make_oracle, small_cases

CASES = small_cases()


@pytest.mark.parametrize("name", ["box2d_n6_outflow", "box3d_n6_dirichlet", "box2d_n4_periodic"])
def test_spd_and_solution(name):
    from oracle import pmg
    c = CASES[name]
    s = make_oracle(c)
    M = pmg.PMG(s, nagg=3, ifvcor=bool(c.ifvcor))
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(s.eshape2), rng.standard_normal(s.eshape2)
    if c.ifvcor:
        a -= a.mean(); b -= b.mean()
    za, zb = M.apply(a), M.apply(b)
    assert abs(np.sum(a * zb) - np.sum(b * za)) < 1e-11 * np.linalg.norm(a) * np.linalg.norm(zb)
    assert np.sum(a * za) > 0 and np.sum(b * zb) > 0
    rhs = s.cdabdtp(a)
    dinv = 1.0 / s.e_diag()
    xj, itj = pmg.pcg(s.cdabdtp, lambda r: dinv * r, rhs, 1e-11)
    xm, itm = pmg.pcg(s.cdabdtp, M.apply, rhs, 1e-11)
    d = xj - xm
    if c.ifvcor:
        d -= d.mean()
    assert np.linalg.norm(d) < 1e-7 * np.linalg.norm(xj)
    assert itm < itj


def test_q1_diagonal_is_galerkin_diagonal():
    """The probed diagonal equals diag(P^T E P) computed column by column."""
    from oracle import pmg
    c = CASES["box2d_n6_outflow"]
    s = make_oracle(c)
    M = pmg.PMG(s, nagg=2)
    for v in range(0, M.nv, 3):
        xv = np.zeros(M.nv); xv[v] = 1.0
        p = M.prolong_q1(xv[M.vid])
        assert abs(np.sum(p * s.cdabdtp(p)) - M.d1[v]) < 1e-10 * M.d1[v]


def test_rcb_aggregates_partition():
    from oracle import pmg
    rng = np.random.default_rng(0)
    cent = rng.random((103, 3))
    for nagg in (1, 2, 7, 16):
        a = pmg.rcb_aggregates(cent, nagg)
        cnt = np.bincount(a, minlength=nagg)
        assert cnt.sum() == 103 and cnt.min() >= 103 // nagg - 1 and cnt.max() <= 103 // nagg + 2


def test_cylinder_iteration_count():
    """Shipped cylinder mesh, lx1=6: Jacobi-PCG needs thousands of iterations, the three-level operator ~2e2."""
    from nekstab_b200 import cases
    from oracle import pmg
    c = cases.cylinder_case(np.load(GOLD + "/cyl.npz"), sponge=False)
    s = make_oracle(c)
    E = s.e_sparse().tocsr()
    ae = lambda p: (E @ p.ravel()).reshape(p.shape)
    M = pmg.PMG(s, nagg=64, apply_e=ae)
    rng = np.random.default_rng(0)
    u = rng.standard_normal((2,) + s.eshape)
    u = np.stack([s.dssum(u[k]) * s.mult * s.mask[k] for k in range(2)])
    b = -s.opdiv(u)
    x, it = pmg.pcg(ae, M.apply, b, 1e-8)
    assert it < 260, it
    assert np.linalg.norm((ae(x) - b).ravel()) <= 1.01e-8 * np.linalg.norm(b.ravel())


def test_q1_vcycle_prototype():
    """Prototype of the next coarse-level design: assembled P^T E P by distance-4 probing equals the Galerkin
    matrix, the resulting operator is symmetric positive definite and needs fewer iterations than the additive coarse levels."""
    from nekstab_b200 import cases
    from oracle import pmg
    c = CASES["box2d_n6_outflow"]
    s = make_oracle(c)
    M = pmg.PMG(s, nagg=3, q1_cycle=(1, 0.7))
    for v in (0, M.nv // 2, M.nv - 1):                       # columns of the probed matrix = E applied to a single hat
        xv = np.zeros(M.nv); xv[v] = 1.0
        col = M._assemble_v(M.restrict_q1(s.cdabdtp(M.prolong_q1(xv[M.vid]))))
        assert np.abs(M.Ac[:, v].toarray().ravel() - col).max() < 1e-10 * np.abs(col).max()
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal(s.eshape2), rng.standard_normal(s.eshape2)
    za, zb = M.apply(a), M.apply(b)
    assert abs(np.sum(a * zb) - np.sum(b * za)) < 1e-10 * np.linalg.norm(a) * np.linalg.norm(zb)
    assert np.sum(a * za) > 0
    cyl = cases.cylinder_case(np.load(GOLD + "/cyl.npz"), sponge=False)
    sc = make_oracle(cyl)
    E = sc.e_sparse().tocsr()
    ae = lambda p: (E @ p.ravel()).reshape(p.shape)
    u = rng.standard_normal((2,) + sc.eshape)
    u = np.stack([sc.dssum(u[k]) * sc.mult * sc.mask[k] for k in range(2)])
    rhs = -sc.opdiv(u)
    it_add = pmg.pcg(ae, pmg.PMG(sc, nagg=64, apply_e=ae).apply, rhs, 1e-8)[1]
    it_cyc = pmg.pcg(ae, pmg.PMG(sc, nagg=64, apply_e=ae, q1_cycle=(1, 0.7)).apply, rhs, 1e-8)[1]
    assert it_cyc < 0.9 * it_add, (it_cyc, it_add)
