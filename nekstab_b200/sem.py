"""Host-side spectral-element quadrature helpers (GLL / GL nodes, weights, derivative and
interpolation matrices).  Stands in for the few speclib.f routines [UPSTREAM Nek5000: zwgll, zwgl,
dgll, igllm] that the case builder needs when it prepares the arrays a Nek5000 host would pass to
``nsb_init``.  The CUDA library has its own C++ copy (csrc/sem_host.cpp); the oracle has an
independent one built on numpy.polynomial (oracle/sem.py) -- tests cross-check the three.
"""
from __future__ import annotations

import numpy as np


def _legendre(n: int, x: np.ndarray):
    """P_n(x) and P_n'(x) by the three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    d0 = np.zeros_like(x)
    d1 = np.ones_like(x)
    for k in range(1, n):
        p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1)
        d2 = d0 + (2 * k + 1) * p1
        p0, p1, d0, d1 = p1, p2, d1, d2
    return p1, d1


def zwgll(n: int):
    """n Gauss-Lobatto-Legendre nodes and weights on [-1,1]."""
    N = n - 1
    x = -np.cos(np.pi * np.arange(n) / N)
    for _ in range(100):
        p, dp = _legendre(N, x)
        # interior nodes are roots of P_N'; Newton on q = P_N' with q' from the Legendre ODE
        xi = x[1:-1]
        q = dp[1:-1]
        dq = (2 * xi * dp[1:-1] - N * (N + 1) * p[1:-1]) / (1 - xi * xi)
        dx = q / dq
        x[1:-1] = xi - dx
        if np.max(np.abs(dx), initial=0.0) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])            # enforce symmetry
    p, _ = _legendre(N, x)
    w = 2.0 / (N * (N + 1) * p * p)
    return x, w


def zwgl(n: int):
    """n Gauss-Legendre nodes and weights on [-1,1]."""
    x = -np.cos(np.pi * (np.arange(n) + 0.75) / (n + 0.5))
    for _ in range(100):
        p, dp = _legendre(n, x)
        dx = p / dp
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x = 0.5 * (x - x[::-1])
    _, dp = _legendre(n, x)
    w = 2.0 / ((1 - x * x) * dp * dp)
    return x, w


def lagrange_interp_matrix(xto: np.ndarray, xfrom: np.ndarray) -> np.ndarray:
    """J[i,l] = l-th Lagrange cardinal function on `xfrom` evaluated at xto[i] (barycentric form)."""
    xfrom = np.asarray(xfrom, float)
    xto = np.asarray(xto, float)
    n = len(xfrom)
    bw = np.array([1.0 / np.prod(xfrom[l] - np.delete(xfrom, l)) for l in range(n)])
    J = np.zeros((len(xto), n))
    for i, x in enumerate(xto):
        d = x - xfrom
        hit = np.nonzero(np.abs(d) < 1e-15)[0]
        if hit.size:
            J[i, hit[0]] = 1.0
        else:
            t = bw / d
            J[i] = t / t.sum()
    return J


def deriv_matrix(x: np.ndarray) -> np.ndarray:
    """D[i,l] = d/dx of the l-th Lagrange cardinal function on nodes x, evaluated at x[i]."""
    x = np.asarray(x, float)
    n = len(x)
    bw = np.array([1.0 / np.prod(x[l] - np.delete(x, l)) for l in range(n)])
    D = np.zeros((n, n))
    for i in range(n):
        for l in range(n):
            if i != l:
                D[i, l] = (bw[l] / bw[i]) / (x[i] - x[l])
        D[i, i] = -np.sum(D[i])
    return D
