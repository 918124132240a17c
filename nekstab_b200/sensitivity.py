"""Host-side post-processing of converged direct / adjoint modes: bi-orthonormalisation and the wavemaker (structural
sensitivity) of Giannetti & Luchini, as nekStab computes them after the two eigenproblems of a direct + adjoint study
(BASELINE config 3: "direct + adjoint eigenproblem with wavemaker sensitivity").

Reference: `wave_maker` core/sensitivity.f:7-81, `biorthogonalize` core/sensitivity.f:428-504, the `bm1s`-weighted
`inner_product` / `norm` core/eigensolvers.f:7-85.  Pointwise work on four mode fields plus four inner products -- outside the
Krylov hot path, so it stays on the host (numpy); the modes come from `nsb_basis_gemv_complex` + `nsb_vec_download`.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def inner_product(p: np.ndarray, q: np.ndarray, bm1s: np.ndarray) -> float:
    """core/eigensolvers.f:7-58: sum_c glsc3(p_c, bm1s, q_c) over the velocity components (ldim, nel, npts)."""
    return float(sum(np.sum(p[c] * bm1s * q[c]) for c in range(p.shape[0])))


def biorthogonalize(d_re, d_im, a_re, a_im, bm1s) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """core/sensitivity.f:428-504: scale the direct mode to unit norm, then rotate / scale the adjoint mode so that
    <a, d> = 1 (complex inner product <a, d> = sum conj(a) . d with the bm1s weight).  Returns (d_re, d_im, a_re, a_im)."""
    gamma = 1.0 / np.sqrt(inner_product(d_re, d_re, bm1s) + inner_product(d_im, d_im, bm1s))
    d_re, d_im = d_re * gamma, d_im * gamma
    gam = inner_product(a_re, d_re, bm1s) + inner_product(a_im, d_im, bm1s)          # real part      (:477-479)
    dlt = inner_product(a_re, d_im, bm1s) - inner_product(a_im, d_re, bm1s)          # imaginary part (:481-483)
    den = gam ** 2 + dlt ** 2
    return d_re, d_im, (gam * a_re - dlt * a_im) / den, (gam * a_im + dlt * a_re) / den


def wave_maker(d_re, d_im, a_re, a_im, bm1s) -> np.ndarray:
    """core/sensitivity.f:63-75: |u_direct(x)| * |u_adjoint(x)| after bi-orthonormalisation; (nel, npts)."""
    d_re, d_im, a_re, a_im = biorthogonalize(d_re, d_im, a_re, a_im, bm1s)
    w1 = np.sqrt(np.sum(d_re ** 2 + d_im ** 2, axis=0))
    w2 = np.sqrt(np.sum(a_re ** 2 + a_im ** 2, axis=0))
    return w1 * w2
