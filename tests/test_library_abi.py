"""The C-ABI shared library: loads without a GPU, exports every symbol include/nekstab_b200.h declares, fails
loudly (no CPU fallback) when no CUDA device exists, and its host-only pieces (LAPACK wrappers, eigenvalue
selection) agree with the oracle / scipy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nekstab_b200 import lib
from util import ROOT


@pytest.fixture(scope="module")
def L():
    l = lib.load_library()
    lp = lib.find_lapack()
    assert lp, "no LAPACK library found"
    assert l.nsb_lapack_load(lp.encode()) == 0
    return l


def test_exports_match_header(L):
    hdr = open(os.path.join(ROOT, "include", "nekstab_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(nsb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(set(lib.EXPORTED)) == names                # the ctypes binding covers the whole header


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure path")
def test_no_cpu_fallback(L):
    x = np.zeros(16)
    glo = np.arange(16, dtype=np.int64)
    rc = L.nsb_init(2, 4, 6, 2, 1, 1, lib._p(x), lib._p(x), None, lib._p(x), lib._p(x), None, lib._p(glo), 0)
    assert rc != 0
    assert b"no CUDA device" in L.nsb_last_error() or b"CUDA" in L.nsb_last_error()
    assert L.nsb_matvec(1, 0, 1) != 0                        # nothing works without nsb_init
    assert b"nsb_init" in L.nsb_last_error()


def test_init_argument_validation(L):
    x = np.zeros(16)
    glo = np.arange(16, dtype=np.int64)
    assert L.nsb_init(4, 4, 6, 2, 1, 1, lib._p(x), lib._p(x), None, lib._p(x), lib._p(x), None, lib._p(glo), 0) != 0
    assert L.nsb_init(2, 6, 9, 6, 1, 1, lib._p(x), lib._p(x), None, lib._p(x), lib._p(x), None, lib._p(glo), 0) != 0   # lx2 != lx1-2
    assert b"lx2" in L.nsb_last_error()


def test_lapack_wrappers(L):
    rng = np.random.default_rng(0)
    n = 12
    A = np.asfortranarray(rng.standard_normal((n, n)) / np.sqrt(n))
    vr, vi, V = np.zeros(n), np.zeros(n), np.zeros(2 * n * n)
    assert L.nsb_lapack_eig(lib._p(A), n, lib._p(vr), lib._p(vi), lib._p(V)) == 0
    vals = vr + 1j * vi
    assert np.all(np.diff(np.abs(vals)) <= 1e-15)                                   # decreasing magnitude
    Vc = (V[0::2] + 1j * V[1::2]).reshape(n, n).T
    assert np.abs(A @ Vc - Vc * vals).max() < 1e-12
    from oracle.krylov import eig
    vo, _ = eig(A)
    assert np.abs(np.sort_complex(vals) - np.sort_complex(vo)).max() < 1e-13
    # schur + ordschur: T stays quasi-triangular, Q orthogonal, A = Q T Q^T, selected eigenvalues lead
    T = A.copy(order="F"); Q = np.zeros((n, n), order="F"); wr, wi = np.zeros(n), np.zeros(n)
    assert L.nsb_lapack_schur(lib._p(T), n, lib._p(Q), lib._p(wr), lib._p(wi)) == 0
    assert np.abs(Q @ T @ Q.T - A).max() < 1e-13 and np.abs(Q.T @ Q - np.eye(n)).max() < 1e-13
    sel = (np.hypot(wr, wi) > np.median(np.hypot(wr, wi))).astype(np.int32)
    # keep conjugate pairs together
    for i in range(n - 1):
        if wi[i] > 0:
            sel[i + 1] = sel[i]
    m = int(sel.sum())
    assert L.nsb_lapack_ordschur(lib._p(T), lib._p(Q), lib._p(sel), n) == 0
    assert np.abs(Q @ T @ Q.T - A).max() < 1e-12
    lead = np.linalg.eigvals(T[:m, :m])
    want = (wr + 1j * wi)[sel.astype(bool)]
    assert np.abs(np.sort_complex(lead) - np.sort_complex(want)).max() < 1e-10
    # lstsq
    M = np.asfortranarray(rng.standard_normal((9, 8))); b = rng.standard_normal(9); x = np.zeros(8)
    assert L.nsb_lapack_lstsq(lib._p(M), lib._p(b), lib._p(x), 9, 8) == 0
    assert np.abs(x - np.linalg.lstsq(M, b, rcond=None)[0]).max() < 1e-12


def test_select_eigenvalues_matches_reference_rule(L):
    from oracle.krylov import select_eigenvalues
    rng = np.random.default_rng(2)
    for trial in range(20):
        n = 30
        re_, im_ = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        for i in range(0, n - 1, 2):                       # conjugate pairs
            re_[i + 1], im_[i + 1] = re_[i], -im_[i]
        sel = np.zeros(n, dtype=np.int32); cnt = C.c_int()
        assert L.nsb_select_eigenvalues(lib._p(sel), C.byref(cnt), lib._p(re_), lib._p(im_), 0.1, 2, n) == 0
        so, co = select_eigenvalues(re_ + 1j * im_, 0.1, 2)
        assert cnt.value == co and np.array_equal(sel.astype(bool), so)
        assert cnt.value >= 2 + 4
