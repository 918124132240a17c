"""ORACLE -- test infrastructure, NOT product code (see oracle/ops.py header for who may import this package).

CPU restatement (numpy/scipy) of the three-level additive pressure preconditioner of csrc/pmg.cu:

    M^-1 r = sum_e R_e^T Etilde_e^-1 R_e r          element blocks by fast diagonalisation (FDM)
           + P diag(P^T E P)^-1 P^T r                Jacobi sweep on the Q1 space of the element-vertex mesh
           + Pa (Pa^T E Pa)^-1 Pa^T r                piecewise constants on element aggregates, solved exactly

The reference's pressure preconditioner is Nek5000's hybrid Schwarz multigrid (`preconditioner = semg_xxt`, 1cyl.par:28;
[UPSTREAM] hsmg.f hsmg_solve, fasts.f / fast3d.f gen_fast, crs_solve): element-local FDM solves + a vertex-mesh coarse
problem.  Its source is not vendored and a preconditioner does not change the converged pressure, so this is a
preconditioner of the same class, not a restatement of hsmg: parity is asserted on (i) the operator M^-1 itself
(GPU vs this file, same aggregates), (ii) the converged solution against the sparse-direct solve, (iii) symmetry and
positive definiteness of M^-1 (what CG needs).  Parity status of iteration counts vs the reference: unpinned.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .ops import SEM


def rcb_aggregates(cent: np.ndarray, nagg: int) -> np.ndarray:
    """Recursive coordinate bisection of element centroids (nel, ldim) into `nagg` groups; ties by element index."""
    nel = cent.shape[0]
    out = np.zeros(nel, dtype=np.int64)

    def rec(idx, first, ng):
        if ng == 1 or idx.size <= 1:
            out[idx] = first
            return
        ext = cent[idx].max(0) - cent[idx].min(0)
        d = int(np.argmax(ext))
        order = idx[np.lexsort((idx, cent[idx, d]))]
        nl = ng // 2
        cut = (idx.size * nl) // ng
        rec(order[:cut], first, nl)
        rec(order[cut:], first + nl, ng - nl)

    rec(np.arange(nel), 0, max(1, min(nagg, nel)))
    return out


class PMG:
    def __init__(self, s: SEM, agg: np.ndarray | None = None, nagg: int = 0, ifvcor: bool = False, apply_e=None, q1_cycle=None):
        """apply_e: callable p -> E p on element-shaped arrays (default: s.cdabdtp).
        q1_cycle = (nu, omega): PROTOTYPE of the next coarse-level design (tools/precond_study5.py; not in the CUDA
        path yet): the Q1 level becomes a symmetric V-cycle on the assembled A_c = P^T E P (nu damped-Jacobi sweeps before and
        after an exact solve on vertex aggregates) instead of one Jacobi sweep + the element-aggregate level."""
        self.s = s
        self.q1_cycle = q1_cycle
        self.ifvcor = bool(ifvcor)
        d, nel, L1, L2 = s.ldim, s.nel, s.lx1, s.lx2
        self.apply_e = apply_e or s.cdabdtp
        X = s.X                                              # (d, nel, [k,] j, i)
        N = L1 - 1
        mid = L1 // 2
        w, w2 = s.w, s.wg
        BD = w2[:, None] * s.D12
        BJ = w2[:, None] * s.J12
        mloc = 1.0 / (s.binv * s.bm1)
        mall = np.prod(s.mask, axis=0)
        # ---- FDM
        self.S = np.zeros((nel, d, L2, L2))
        lam = np.zeros((nel, d, L2))
        h = np.zeros((nel, d))
        for dr in range(d):
            ax = -1 - dr                                     # array axis of direction dr (relative to the element axes)
            lo = np.take(X, 0, axis=ax).reshape(d, nel, -1).mean(-1)
            hi = np.take(X, N, axis=ax).reshape(d, nel, -1).mean(-1)
            h[:, dr] = np.linalg.norm(hi - lo, axis=0)
        for dr in range(d):
            ax = -1 - dr
            sel = [slice(None)] + [mid] * d                  # probe node: tangential indices = lx1/2

            def probe(a, end):
                ss = list(sel)
                ss[d + 1 + ax if ax < 0 else ax] = end
                return a[tuple(ss)]
            ml, mr = probe(mloc, 0), probe(mloc, N)
            kl, kr = probe(mall, 0), probe(mall, N)
            for e in range(nel):
                wi = 1.0 / w
                wi[0] = kl[e] / (w[0] * ml[e])
                wi[-1] = kr[e] / (w[-1] * mr[e])
                A = (BD * wi) @ BD.T
                M = (BJ * wi) @ BJ.T
                lm, S = sla.eigh(A, M)
                self.S[e, dr] = S
                lam[e, dr] = lm
        cd = np.prod(h, axis=1, keepdims=True) / h ** 2 * (0.5 if d == 3 else 1.0)
        lam = lam * cd[:, :, None]
        den = 0.0
        for dr in range(d):
            shp = [nel] + [1] * d
            shp[d - dr] = L2
            den = den + lam[:, dr].reshape(shp)
        mx = den.reshape(nel, -1).max(1).reshape([nel] + [1] * d)
        self.deninv = np.where(den > 1e-12 * mx, 1.0 / np.where(den > 1e-12 * mx, den, 1.0), 0.0)
        self.h = h
        # ---- Q1 level
        corner = [0, N]
        G = s.glo.reshape((nel,) + (L1,) * d)
        zg = s.zg
        l = [(1 - zg) / 2, (1 + zg) / 2]
        nk = 2 ** d
        vg = np.zeros((nel, nk), dtype=np.int64)
        self.phi = np.zeros((nk,) + (L2,) * d)
        for k in range(nk):
            bits = [(k >> dr) & 1 for dr in range(d)]        # bit dr = direction dr (r fastest)
            idx = tuple(corner[bits[d - 1 - a]] for a in range(d))      # array axes ordered (t,) s, r
            vg[:, k] = G[(slice(None),) + idx]
            f = 1.0
            for a in range(d):
                f = np.multiply.outer(f, l[bits[d - 1 - a]]) if a else l[bits[d - 1]]
            self.phi[k] = f
        uv, vid = np.unique(vg, return_inverse=True)
        self.vid = vid.reshape(nel, nk)
        self.nv = uv.size
        # ---- aggregates
        cent = X.reshape(d, nel, -1).mean(-1).T
        if agg is None:
            agg = rcb_aggregates(cent, nagg if nagg > 0 else max(1, nel // 32))
        self.agg = np.asarray(agg, dtype=np.int64)
        self.nagg = int(self.agg.max()) + 1
        # ---- Galerkin pieces by probing with E
        self.d1 = self._q1_diagonal()
        A2 = np.zeros((self.nagg, self.nagg))
        for a in range(self.nagg):
            p = np.zeros(s.eshape2)
            p[self.agg == a] = 1.0
            A2[:, a] = self._agg_sums(self.apply_e(p))
        A2 = 0.5 * (A2 + A2.T)
        if ifvcor:                                           # E 1 = 0 => A2 1 = 0: shift the null vector (r is kept orthogonal to 1)
            A2 = A2 + np.trace(A2) / self.nagg ** 2
        self.A2 = A2
        self.A2inv = np.zeros_like(A2) if (ifvcor and self.nagg == 1) else np.linalg.inv(A2)
        if q1_cycle is not None:
            self._setup_q1_cycle()

    # ------------------------------------------------------------------ pieces
    def restrict_q1(self, r):
        nk = self.phi.shape[0]
        return r.reshape(self.s.nel, -1) @ self.phi.reshape(nk, -1).T

    def prolong_q1(self, xek):
        nk = self.phi.shape[0]
        return (xek @ self.phi.reshape(nk, -1)).reshape(self.s.eshape2)

    def _assemble_v(self, rc):
        return np.bincount(self.vid.ravel(), weights=rc.ravel(), minlength=self.nv)

    def _agg_sums(self, r):
        return np.bincount(self.agg, weights=r.reshape(self.s.nel, -1).sum(1), minlength=self.nagg)

    def _colouring(self):
        """Greedy distance-2 colouring of the vertex graph (vertices adjacent when they share an element)."""
        nv = self.nv
        adj = [set() for _ in range(nv)]
        for row in self.vid:
            for a in row:
                adj[a].update(row)
        col = -np.ones(nv, dtype=np.int64)
        for v in range(nv):
            forb = set()
            for u in adj[v]:
                for t in adj[u]:
                    if col[t] >= 0:
                        forb.add(col[t])
            c = 0
            while c in forb:
                c += 1
            col[v] = c
        return col

    def _q1_diagonal(self):
        col = self._colouring()
        d1 = np.zeros(self.nv)
        for c in range(int(col.max()) + 1):
            xv = (col == c).astype(float)
            p = self.prolong_q1(xv[self.vid])
            rv = self._assemble_v(self.restrict_q1(self.apply_e(p)))
            d1[col == c] = rv[col == c]
        self.ncolours = int(col.max()) + 1
        return d1

    # ------------------------------------------------------------------ prototype: Q1 V-cycle (see __init__)
    def _neighbourhoods(self):
        nv = self.nv
        adj = [set() for _ in range(nv)]
        for row in self.vid:
            for a in row:
                adj[a].update(int(x) for x in row)
        n2 = [set().union(*(adj[u] for u in adj[v])) for v in range(nv)]          # graph distance <= 2
        return adj, n2

    def _setup_q1_cycle(self):
        """Assemble A_c = P^T E P column by column: vertices of one colour are more than 4 apart (distance-4 colouring), so the
        columns probed together (support: distance <= 2) do not overlap; vertex aggregates = aggregate of the first element
        holding the vertex; A2v = P2^T A_c P2 solved exactly (constant shifted when E is singular)."""
        import scipy.sparse as sp
        nv = self.nv
        _, n2 = self._neighbourhoods()
        col = -np.ones(nv, dtype=np.int64)
        for v in range(nv):
            forb = set()
            for u in n2[v]:
                for t in n2[u]:
                    if col[t] >= 0:
                        forb.add(int(col[t]))
            c = 0
            while c in forb:
                c += 1
            col[v] = c
        self.ncolours4 = int(col.max()) + 1
        rows, cols, vals = [], [], []
        for c in range(self.ncolours4):
            xv = (col == c).astype(float)
            y = self._assemble_v(self.restrict_q1(self.apply_e(self.prolong_q1(xv[self.vid]))))
            for v in np.flatnonzero(col == c):
                for w in n2[v]:
                    rows.append(w); cols.append(v); vals.append(y[w])
        A = sp.coo_matrix((vals, (rows, cols)), shape=(nv, nv)).tocsr()
        self.Ac = (0.5 * (A + A.T)).tocsr()
        self.dAc = self.Ac.diagonal()
        vagg = np.zeros(nv, dtype=np.int64)
        for e in range(self.s.nel - 1, -1, -1):
            vagg[self.vid[e]] = self.agg[e]
        self.P2 = sp.coo_matrix((np.ones(nv), (np.arange(nv), vagg)), shape=(nv, self.nagg)).tocsr()
        A2v = (self.P2.T @ self.Ac @ self.P2).toarray()
        if self.ifvcor:
            A2v = A2v + np.trace(A2v) / self.nagg ** 2
        self.A2vinv = np.zeros_like(A2v) if (self.ifvcor and self.nagg == 1) else np.linalg.inv(A2v)

    def _q1_vcycle(self, rc):
        nu, om = self.q1_cycle
        x = np.zeros_like(rc)
        for _ in range(nu):
            x = x + om * (rc - self.Ac @ x) / self.dAc
        x = x + self.P2 @ (self.A2vinv @ (self.P2.T @ (rc - self.Ac @ x)))
        for _ in range(nu):
            x = x + om * (rc - self.Ac @ x) / self.dAc
        return x

    def fdm(self, r):
        d = self.s.ldim
        t = r
        for dr in range(d):                                  # S^T along every direction
            t = np.moveaxis(np.einsum('eai,e...a->e...i', self.S[:, dr], np.moveaxis(t, d - dr, -1)), -1, d - dr)
        t = t * self.deninv
        for dr in range(d):
            t = np.moveaxis(np.einsum('eai,e...i->e...a', self.S[:, dr], np.moveaxis(t, d - dr, -1)), -1, d - dr)
        return t

    def apply(self, r):
        """z = M^-1 r on element-shaped mesh-2 arrays."""
        rc = self.restrict_q1(r)
        if self.q1_cycle is not None:
            xv = self._q1_vcycle(self._assemble_v(rc))
            return self.fdm(r) + self.prolong_q1(xv[self.vid])
        xv = self._assemble_v(rc) / self.d1
        x2 = self.A2inv @ np.bincount(self.agg, weights=rc.sum(1), minlength=self.nagg)
        shp = [self.s.nel] + [1] * self.s.ldim
        return self.fdm(r) + self.prolong_q1(xv[self.vid]) + x2[self.agg].reshape(shp)


def pcg(apply_e, minv, b, tol, maxit=20000, norm=None):
    """Preconditioned CG; stops on norm(r) <= tol (norm defaults to the 2-norm relative to b)."""
    x = np.zeros_like(b)
    r = b.copy()
    p = np.zeros_like(b)
    rtz1 = 1.0
    nb = np.linalg.norm(b.ravel())
    it = 0
    while True:
        rn = norm(r) if norm else np.linalg.norm(r.ravel()) / nb
        if rn <= tol or it >= maxit:
            return x, it
        z = minv(r)
        rtz2, rtz1 = rtz1, float(np.sum(z * r))
        beta = 0.0 if it == 0 else rtz1 / rtz2
        p = z + beta * p
        w = apply_e(p)
        alpha = rtz1 / float(np.sum(w * p))
        x += alpha * p
        r -= alpha * w
        it += 1
