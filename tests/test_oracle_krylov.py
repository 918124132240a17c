"""The oracle's Krylov layer (restated from core/krylov_decomposition.f, core/eigensolvers.f, core/newton_krylov.f) on a
small dense operator with a known spectrum: exp(tau*A).  Validates the restatement itself before it is used as the
checker for the CUDA-side Arnoldi / Krylov-Schur / GMRES drivers."""
import numpy as np
import scipy.linalg as sla

from oracle import krylov


def _setup(n=60, seed=1):
    rng = np.random.default_rng(seed)
    A = -np.diag(np.linspace(0.05, 3.0, n)) + 0.4 * np.triu(rng.standard_normal((n, n)), 1) / np.sqrt(n)
    A[0, 0], A[1, 1], A[0, 1], A[1, 0] = 0.02, 0.02, 0.7, -0.7      # one unstable oscillatory pair
    M = sla.expm(1.0 * A)
    w = np.ones(n)

    def mv(q):
        return ((M @ q[0][0])[None, :], q[1])
    q0 = (rng.standard_normal((1, n)), np.zeros(1))
    q0 = krylov.scale(q0, 1.0 / np.sqrt(krylov.inner(q0, q0, w)))
    return A, M, mv, q0, w


def test_arnoldi_relation_and_orthonormality():
    A, M, mv, q0, w = _setup()
    k = 20
    Q, H = krylov.arnoldi_factorization(mv, q0, k, w)
    Qm = np.stack([q[0][0] for q in Q], axis=1)
    assert np.abs(Qm.T @ Qm - np.eye(k + 1)).max() < 1e-13
    assert np.abs(M @ Qm[:, :k] - Qm @ H).max() < 1e-12
    assert np.abs(np.tril(H, -2)).max() == 0.0


def test_krylov_schur_converges_to_leading_pair():
    A, M, mv, q0, w = _setup()
    vals, vecs, res, Q, H, cnt, scnt = krylov.krylov_schur(mv, q0, 16, 2, w, eigen_tol=1e-10)
    assert cnt >= 2 and scnt >= 1                                     # needed at least one Schur restart
    exact = np.linalg.eigvals(M)
    lead = exact[np.argsort(-np.abs(exact))][:2]
    assert np.abs(np.sort_complex(vals[:2]) - np.sort_complex(lead)).max() < 1e-9
    lam = np.log(vals[0]) / 1.0
    assert abs(lam.real - 0.02) < 1e-8 and abs(abs(lam.imag) - 0.7) < 1e-8


def test_ts_gmres_solves_newton_system():
    A, M, mv, q0, w = _setup()
    n = M.shape[0]

    def jac(q):                                                        # (exp(TL) - I) q, core/matvec.f:397-400
        return ((M @ q[0][0] - q[0][0])[None, :], q[1])
    rng = np.random.default_rng(5)
    rhs = (rng.standard_normal((1, n)), np.zeros(1))
    sol, calls, res = krylov.ts_gmres(jac, rhs, 10, 30, 1e-20, w)
    assert res < 1e-20
    assert np.abs((M - np.eye(n)) @ sol[0][0] - rhs[0][0]).max() < 1e-9
