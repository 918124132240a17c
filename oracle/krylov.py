"""ORACLE -- test infrastructure, NOT product code (see oracle/ops.py header).

The Krylov layer of nekStab restated 1:1 in numpy/scipy:
  update_hessenberg_matrix  core/krylov_decomposition.f:116-202  (modified Gram-Schmidt applied TWICE)
  arnoldi_factorization     core/krylov_decomposition.f:7-104
  krylov_schur loop         core/eigensolvers.f:335-373
  schur_condensation        core/eigensolvers.f:395-499
  select_eigenvalues        core/eigensolvers.f:729-795
  eig/schur/ordschur/lstsq  core/lapack_wrapper.f:7-339 (scipy's dgeev/dgees/dtrsen/dgels)
  ts_gmres                  core/newton_krylov.f:175-297
A Krylov vector is a tuple (v, p) or (v, p, time): v (ldim, ...) velocity, p pressure, time the period unknown of the UPO Newton
(core/krylov_subspace.f:8-15).  The inner product is the reference's semi-norm: velocity only, weight bm1s (:37-45), plus
time * time in the UPO case (:47-50).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
from scipy.linalg import lapack


def inner(a, b, w):
    """krylov_inner_product (core/krylov_subspace.f:24-56): velocity under bm1s, plus the product of the time components when the
    vectors carry one (third tuple entry; uparam(1) = 2.1, :47-50)."""
    v = float(sum(np.sum(a[0][d] * w * b[0][d]) for d in range(a[0].shape[0])))
    return v + float(a[2]) * float(b[2]) if len(a) > 2 else v


def axpy(a, alpha, b):
    return tuple(x + alpha * y for x, y in zip(a, b))


def scale(a, alpha):
    return tuple(x * alpha for x in a)


def update_hessenberg_matrix(H, f, Q, k, w):
    """Two sweeps of modified Gram-Schmidt of f against Q[0..k-1]; H[:k, k-1] accumulates both; normalise."""
    for sweep in range(2):
        for i in range(k):
            alpha = inner(f, Q[i], w)
            f = axpy(f, -alpha, Q[i])
            H[i, k - 1] = alpha if sweep == 0 else H[i, k - 1] + alpha
    beta = np.sqrt(inner(f, f, w))
    H[k, k - 1] = beta
    return scale(f, 1.0 / beta)


def arnoldi_factorization(matvec, q0, k, w, Q=None, H=None, mstart=1):
    """Q: list of k+1 vectors, H: (k+1, k).  matvec(q) -> (v, p)."""
    if Q is None:
        Q = [q0] + [None] * k
        H = np.zeros((k + 1, k))
    for m in range(mstart, k + 1):
        f = matvec(Q[m - 1])
        Q[m] = update_hessenberg_matrix(H, f, Q, m, w)
    return Q, H


def eig(A):
    """core/lapack_wrapper.f:129-256: dgeev + the reference's exchange sort by decreasing magnitude."""
    wr, wi, _, vr, info = lapack.dgeev(np.asfortranarray(A), compute_vl=0, compute_vr=1)
    n = A.shape[0]
    vals = wr + 1j * wi
    vecs = vr.astype(complex)
    for i in range(n - 1):
        if wi[i] > 0:
            vecs[:, i] = vr[:, i] + 1j * vr[:, i + 1]
            vecs[:, i + 1] = vr[:, i] - 1j * vr[:, i + 1]
    nrm = np.abs(vals).copy()
    for k in range(n - 1):
        for l in range(k + 1, n):
            if nrm[k] < nrm[l]:
                nrm[[k, l]] = nrm[[l, k]]
                vals[[k, l]] = vals[[l, k]]
                vecs[:, [k, l]] = vecs[:, [l, k]]
    return vals, vecs


def select_eigenvalues(vals, delta, nev):
    n = len(vals)
    mag = np.abs(vals)
    idx = np.argsort(mag, kind="stable")
    sel = mag >= (1.0 - delta)
    sel[idx[max(0, n - (nev + 4)):]] = True
    if n - (nev + 5) >= 0 and vals[idx[n - (nev + 4)]].imag == -vals[idx[n - (nev + 5)]].imag:
        sel[idx[n - (nev + 5)]] = True
    return sel, int(sel.sum())


def schur_condensation(H, Q, ksize, schur_tgt, schur_del):
    """Returns new mstart; H and Q modified in place (core/eigensolvers.f:395-499)."""
    k = ksize
    b = np.zeros(k); b[k - 1] = H[k, k - 1]
    T, Z, sdim = sla.schur(H[:k, :k], output="real", sort=lambda re, im: np.hypot(re, im) > 0.9)
    vals = np.array(sla.eigvals(T))
    # eigenvalues in the order they sit on the diagonal of T (what dgees returns in wr, wi)
    wr, wi = _diag_eigs(T)
    vals = wr + 1j * wi
    sel, ms = select_eigenvalues(vals, schur_del, schur_tgt)
    res = lapack.dtrsen(sel.astype(np.int32), np.asfortranarray(T), np.asfortranarray(Z), job="N", wantq=1)
    T, Z = res[0], res[1]
    H[:k, :k] = T
    H[:ms, ms:k] = 0.0
    H[ms:k + 1, :] = 0.0
    Zm = np.asarray(Z)
    newQ = []
    for j in range(k):
        v = sum(Zm[i, j] * Q[i][0] for i in range(k))
        p = sum(Zm[i, j] * Q[i][1] for i in range(k))
        newQ.append((v, p))
    Q[:k] = newQ
    H[ms, :] = b @ Zm
    ms += 1
    Q[ms - 1] = Q[k]
    return ms


def _diag_eigs(T):
    n = T.shape[0]
    wr, wi = np.zeros(n), np.zeros(n)
    i = 0
    while i < n:
        if i + 1 < n and T[i + 1, i] != 0.0:
            ev = np.linalg.eigvals(T[i:i + 2, i:i + 2])
            ev = sorted(ev, key=lambda z: -z.imag)
            wr[i], wi[i] = ev[0].real, ev[0].imag
            wr[i + 1], wi[i + 1] = ev[1].real, ev[1].imag
            i += 2
        else:
            wr[i] = T[i, i]
            i += 1
    return wr, wi


def krylov_schur(matvec, q0, k_dim, schur_tgt, w, eigen_tol=1e-6, schur_del=0.1, max_restarts=50):
    Q = [q0] + [None] * k_dim
    H = np.zeros((k_dim + 1, k_dim))
    mstart, scnt = 1, 0
    while True:
        arnoldi_factorization(matvec, None, k_dim, w, Q, H, mstart)
        vals, vecs = eig(H[:k_dim, :k_dim])
        residual = np.abs(H[k_dim, k_dim - 1] * vecs[k_dim - 1, :])
        cnt = int(np.sum(residual < eigen_tol))
        if schur_tgt <= 0 or cnt >= schur_tgt or scnt >= max_restarts:
            break
        scnt += 1
        mstart = schur_condensation(H, Q, k_dim, schur_tgt, schur_del)
    return vals, vecs, residual, Q, H, cnt, scnt


def ts_gmres(matvec, rhs, maxiter, ksize, tol, w):
    """core/newton_krylov.f:175-297; squared residual norms are compared with tol (:268,:288)."""
    sol = tuple(np.zeros_like(x) if isinstance(x, np.ndarray) else 0.0 for x in rhs)
    beta = np.sqrt(inner(rhs, rhs, w))
    q1 = scale(rhs, 1.0 / beta)
    calls = 0
    for _ in range(maxiter):
        Q = [q1] + [None] * ksize
        H = np.zeros((ksize + 1, ksize))
        e = np.zeros(ksize + 1); e[0] = beta
        kk = ksize
        for k in range(1, ksize + 1):
            arnoldi_factorization(matvec, None, k, w, Q, H, mstart=k)
            y = np.linalg.lstsq(H[:k + 1, :k], e[:k + 1], rcond=None)[0]
            beta = np.linalg.norm(e[:k + 1] - H[:k + 1, :k] @ y)
            if beta ** 2 < tol:
                calls += k
                kk = k
                break
        dq = tuple(sum(y[j] * Q[j][m] for j in range(kk)) for m in range(len(rhs)))          # krylov_matmul (:275)
        sol = axpy(sol, 1.0, dq)
        f = matvec(sol)
        f = scale(axpy(f, -1.0, rhs), -1.0)
        beta = np.sqrt(inner(f, f, w))
        q1 = scale(f, 1.0 / beta)
        if beta ** 2 < tol:
            break
    return sol, calls, beta ** 2


def newton_krylov(nonlinear_map, linearized_map_factory, q0, k_dim, tol, w, maxiter_newton=100, maxiter_gmres=100):
    """core/newton_krylov.f:5-168, fixed-point branch.  nonlinear_map(q) -> phi_T(q) - q ; linearized_map_factory(q) returns
    the matvec q' -> (exp(TL(q)) - I) q' about the current iterate (newton_linearized_map, core/matvec.f:381-402)."""
    q = q0
    hist = []
    for it in range(1, maxiter_newton + 1):
        f = nonlinear_map(q)
        residual = inner(f, f, w)                      # squared norm (:99)
        hist.append(residual)
        if residual < tol:
            break
        dq, calls, _ = ts_gmres(linearized_map_factory(q), f, maxiter_gmres, k_dim, tol, w)
        q = axpy(q, -1.0, dq)
    return q, it, hist
