"""Quadrature nodes/weights and SEM matrices: product (Newton iteration) vs oracle (numpy.polynomial) vs closed forms."""
import numpy as np
import pytest

from nekstab_b200 import sem
from oracle import sem as osem


@pytest.mark.parametrize("n", [4, 6, 8, 9, 12])
def test_gll_gl_and_matrices(n):
    x, w = sem.zwgll(n)
    xo, wo = osem.gll(n)
    assert np.abs(x - xo).max() < 1e-14 and np.abs(w - wo).max() < 1e-14
    assert abs(w.sum() - 2.0) < 1e-14
    xg, wg = sem.zwgl(n)
    xgo, wgo = osem.gl(n)
    assert np.abs(xg - xgo).max() < 1e-14 and np.abs(wg - wgo).max() < 1e-14
    D = sem.deriv_matrix(x)
    assert np.abs(D - osem.deriv(xo)).max() < 1e-12
    N = n - 1
    assert abs(D[0, 0] + N * (N + 1) / 4) < 1e-12          # SURVEY App. E.1 corner value
    assert np.abs(D @ np.ones(n)).max() < 1e-12
    J = sem.lagrange_interp_matrix(xg, x)
    assert np.abs(J - osem.interp(xgo, xo)).max() < 1e-13
    # exactness: quadrature integrates degree 2n-3 (GLL) / 2n-1 (GL); D differentiates degree n-1 exactly
    k = 2 * n - 3
    assert abs(np.sum(w * x ** (k - 1)) - (2.0 / k if (k - 1) % 2 == 0 else 0.0)) < 1e-13
    assert np.abs(D @ x ** (n - 1) - (n - 1) * x ** (n - 2)).max() < 1e-11
