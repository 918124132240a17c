"""CPU tests of the host-side set-up logic of the pressure preconditioner (csrc/pmg.cu), through the C ABI's host-only entry
points: aggregates, vertex colouring, FDM factors and the dense SPD inverse, against oracle/pmg.py / scipy."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla

from nekstab_b200 import cases, lib
from util import GOLD, small_cases


@pytest.fixture(scope="module")
def L():
    return lib.load_library()


def test_aggregates_match_oracle_bisection(L):
    from oracle import pmg
    rng = np.random.default_rng(0)
    for nel, ldim, nagg in ((103, 3, 7), (64, 2, 16), (50, 3, 1), (9, 2, 20)):
        cent = rng.random((nel, ldim))
        cent[::5] = cent[0]                                    # ties are broken by the element index
        out = np.zeros(nel, dtype=np.int32)
        assert L.nsb_pm_host_aggregates(ldim, nel, lib._p(np.ascontiguousarray(cent)), nagg, out.ctypes.data_as(lib._ip)) == 0
        ref = pmg.rcb_aggregates(cent, nagg)
        assert np.array_equal(out, ref)
        cnt = np.bincount(out, minlength=min(nagg, nel))
        assert cnt.min() >= 1 and cnt.max() - cnt.min() <= 1 + nel // max(1, min(nagg, nel)) // 2


def test_colouring_is_a_valid_distance2_colouring(L):
    g = np.load(GOLD + "/cyl.npz")
    for case in (cases.cylinder_case(g, sponge=False), small_cases()["box3d_n6_dirichlet"]):
        nk = 2 ** case.ldim
        N = case.lx1 - 1
        G = case.glo.reshape((case.nel,) + (case.lx1,) * case.ldim)
        vg = np.stack([G[(slice(None),) + tuple(N * ((k >> (case.ldim - 1 - a)) & 1) for a in range(case.ldim))] for k in range(nk)], axis=1)
        vg = np.ascontiguousarray(vg, dtype=np.int64)
        col = np.zeros((case.nel, nk), dtype=np.int32)
        ncol = C.c_int()
        assert L.nsb_pm_host_colouring(case.nel, nk, vg.ctypes.data_as(lib._lp), col.ctypes.data_as(lib._ip), C.byref(ncol)) == 0
        assert col.min() == 0 and col.max() == ncol.value - 1
        # one colour per global vertex
        cv = {}
        for gid, cc in zip(vg.ravel(), col.ravel()):
            assert cv.setdefault(int(gid), int(cc)) == int(cc)
        # two vertices are within distance 2 iff both lie in elements touching a common vertex u: all vertices of the elements
        # around u must carry different colours (then hats of one colour do not see each other through E)
        elems_of = {}
        for e in range(case.nel):
            for gid in vg[e]:
                elems_of.setdefault(int(gid), []).append(e)
        for u, els in elems_of.items():
            near = set()
            for e2 in els:
                near.update(int(x) for x in vg[e2])
            cols = [cv[x] for x in near]
            assert len(cols) == len(set(cols)), "two vertices within distance 2 share a colour"
        assert ncol.value <= (40 if case.ldim == 2 else 80)


def test_fdm_factors_match_scipy(L):
    from oracle import sem
    for lx1 in (4, 6, 8):
        lx2 = lx1 - 2
        z, w = sem.gll(lx1)
        zg, wg = sem.gl(lx2)
        J12 = sem.interp(zg, z)
        D12 = J12 @ sem.deriv(z)
        for wf, wl in ((0.5 / w[0], 0.5 / w[-1]), (0.0, 1.0 / w[-1]), (0.0, 0.0), (0.3 / w[0], 0.0)):
            wi = 1.0 / w
            wi[0], wi[-1] = wf, wl
            A = ((wg[:, None] * D12) * wi) @ (wg[:, None] * D12).T
            M = ((wg[:, None] * J12) * wi) @ (wg[:, None] * J12).T
            lam_ref = sla.eigh(A, M, eigvals_only=True)
            S = np.zeros((lx2, lx2)); lam = np.zeros(lx2)
            assert L.nsb_pm_host_fdm_1d(lx1, wf, wl, lib._p(S), lib._p(lam)) == 0
            assert np.abs(np.sort(lam) - lam_ref).max() < 1e-10 * max(1.0, np.abs(lam_ref).max())
            assert np.abs(S.T @ M @ S - np.eye(lx2)).max() < 1e-11                      # S^T M S = I
            assert np.abs(S.T @ A @ S - np.diag(lam)).max() < 1e-9 * max(1.0, np.abs(lam).max())


def test_spd_inverse(L):
    rng = np.random.default_rng(3)
    for n in (1, 5, 64, 300):
        B = rng.standard_normal((n, n))
        A = B @ B.T + n * np.eye(n)
        X = np.ascontiguousarray(A.copy())
        assert L.nsb_pm_host_spd_inverse(n, lib._p(X)) == 0
        assert np.abs(X @ A - np.eye(n)).max() < 1e-10
    X = np.ascontiguousarray(-np.eye(3))
    assert L.nsb_pm_host_spd_inverse(3, lib._p(X)) != 0 and b"positive definite" in L.nsb_last_error()
