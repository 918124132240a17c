"""Size-independent properties at BASELINE.json's full single-GPU size (config 5: 19 960 hexahedra, lx1 = 8, 1.02e7 points
per field), where the oracle is too slow to serve as the checker: averaging projector of the gather-scatter, symmetry and
positivity of E and of the pressure preconditioner, the converged pressure solve, linearity of the matvec, orthonormality
of the Gram-Schmidt step.  Everything goes through the C ABI.  (The oracle comparison of the cfg-5 matvec itself lives in
tests/test_gpu_cfg5_oracle.py.)"""
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    sys.path.insert(0, ROOT)
    import bench
    from nekstab_b200 import lib
    case, n_glob = bench.build_workload(10)
    assert case.nel == 19960 and n_glob == 19960 * 512
    g = lib.NekStabB200(case)
    g.set_params(1.0 / case.re, 1.0, 1e-10, 1e-10, 2000, 100000)
    g.set_pressure_preconditioner(1, 0)
    yield case, g
    g.close()


def test_dssum_average_is_a_projector(full):
    c, g = full
    rng = np.random.default_rng(0)
    u = rng.standard_normal(c.n)
    mult = 1.0 / g.op_dssum(np.ones(c.n))
    a = g.op_dssum(u) * mult                       # direct-stiffness average
    b = g.op_dssum(a) * mult
    assert np.abs(b - a).max() <= 1e-13 * np.abs(a).max()
    assert abs(np.sum(a * g.op_dssum(u)) - np.sum(g.op_dssum(a) * u)) <= 1e-12 * abs(np.sum(a * u)) + 1e-9   # QQ^T symmetric


def test_E_and_preconditioner_symmetric_positive(full):
    c, g = full
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(g.n2), rng.standard_normal(g.n2)
    ea, eb = g.op_cdabdtp(a), g.op_cdabdtp(b)
    assert abs(a @ eb - b @ ea) <= 1e-11 * np.linalg.norm(a) * np.linalg.norm(eb)
    assert a @ ea > 0 and b @ eb > 0
    za, zb = g.op_pc_apply(a), g.op_pc_apply(b)
    assert abs(a @ zb - b @ za) <= 1e-11 * np.linalg.norm(a) * np.linalg.norm(zb)
    assert a @ za > 0 and b @ zb > 0
    info = g.pc_get(3)
    assert int(info[1]) == 512 and int(info[2]) <= 64          # aggregates; colours used by the probing


def test_pressure_solve_residual_and_iterations(full):
    c, g = full
    rng = np.random.default_rng(2)
    x_true = rng.standard_normal(g.n2)
    rhs = g.op_cdabdtp(x_true)
    bm2 = g.get_field("bm2")
    vol2 = bm2.sum()
    tol = 1e-9 * float(np.sqrt(np.sum(rhs * rhs / bm2) / vol2))
    g.set_params(1.0 / c.re, 1.0, 1e-10, tol, 2000, 100000)
    x, it = g.op_esolver(rhs)
    r = rhs - g.op_cdabdtp(x)
    assert float(np.sqrt(np.sum(r * r / bm2) / vol2)) <= 1.5 * tol      # the recursive residual is the true residual
    assert it < 1500                                                      # Jacobi needs > 2e4 for this reduction on white noise
    g.set_params(1.0 / c.re, 1.0, 1e-10, 1e-10, 2000, 100000)


def test_matvec_is_linear_and_gram_schmidt_orthonormal(full):
    c, g = full
    from nekstab_b200 import cases, lib
    g.vec_alloc(8)
    dt, _, _ = g.prepare_linearized_solver(1.0, 0.5)
    g.set_timestep(dt, 1)
    rng = np.random.default_rng(3)
    seed = cases.add_noise(c).reshape(3, -1)
    mult = 1.0 / g.op_dssum(np.ones(c.n))
    other = np.stack([g.op_dssum(rng.standard_normal(c.n)) * mult for _ in range(3)]) * c.mask.reshape(3, -1)
    g.vec_upload(0, seed, None)
    g.vec_upload(1, other, None)
    g.normalize(0); g.normalize(1)
    al, be = 0.7, -1.3
    g.vec_copy(2, 0); g.vec_cmult(2, al)
    g.vec_copy(3, 1); g.vec_cmult(3, be)
    g.vec_add2(2, 3)                                   # slot 2 = al x + be y
    g.matvec(lib.DIRECT, 0, 4)
    g.matvec(lib.DIRECT, 1, 5)
    g.matvec(lib.DIRECT, 2, 6)
    g.vec_cmult(4, al); g.vec_cmult(5, be); g.vec_add2(4, 5)
    g.vec_sub2(6, 4)
    assert g.norm(6) <= 1e-7 * g.norm(4)               # linear to the solver tolerance (1e-10 absolute residuals)
    # Gram-Schmidt (CGS2/DGKS) on the device-resident basis
    g.vec_upload(0, seed, None); g.normalize(0)
    for k in range(1, 4):
        g.vec_upload(k, np.stack([g.op_dssum(rng.standard_normal(c.n)) * mult for _ in range(3)]), None)
        g.orthonormalize(k, 0, k)
    for i in range(4):
        for j in range(i + 1):
            ip = g.inner_product(i, j)
            assert abs(ip - (1.0 if i == j else 0.0)) < 1e-12, (i, j, ip)

