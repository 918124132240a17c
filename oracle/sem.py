"""ORACLE (test infrastructure, not product code).  Quadrature nodes/weights and the small dense
matrices of the spectral-element method, built on numpy.polynomial.legendre so that they are
independent of the product's Newton-iteration copies.
Follows [UPSTREAM Nek5000 speclib.f: zwgll, zwgl, dgll, igllm/iglm] as used by the time stepper that
nekStab drives through `nek_advance` (core/matvec.f:222); conventions of SURVEY.md App. E.1.
"""
import numpy as np
from numpy.polynomial import legendre as L


def gll(n):
    N = n - 1
    cN = np.zeros(N + 1); cN[N] = 1.0
    xi = np.sort(L.legroots(L.legder(cN))) if N > 1 else np.array([])
    x = np.concatenate(([-1.0], xi, [1.0]))
    # polish interior roots (Newton on P_N')
    for _ in range(3):
        d1 = L.legval(x[1:-1], L.legder(cN)); d2 = L.legval(x[1:-1], L.legder(cN, 2))
        x[1:-1] -= d1 / d2
    x = 0.5 * (x - x[::-1])
    w = 2.0 / (N * (N + 1) * L.legval(x, cN) ** 2)
    return x, w


def gl(n):
    x, w = L.leggauss(n)
    return x, w


def interp(xto, xfrom):
    """Lagrange interpolation matrix (len(xto), len(xfrom)) via Vandermonde solve in Legendre basis."""
    n = len(xfrom)
    V = L.legvander(xfrom, n - 1)
    Vt = L.legvander(xto, n - 1)
    return np.linalg.solve(V.T, Vt.T).T


def deriv(x):
    """Derivative matrix on nodes x: D[i,l] = l_l'(x_i)."""
    n = len(x)
    V = L.legvander(x, n - 1)
    dV = np.stack([L.legval(x, L.legder(np.eye(n)[k])) for k in range(n)], axis=1)
    return np.linalg.solve(V.T, dV.T).T
