"""ORACLE -- test infrastructure, NOT product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

CPU restatement (numpy/scipy, float64) of the spectral-element operators that Nek5000's perturbation
time stepper applies when nekStab calls `nek_advance` (reference call sites: core/matvec.f:222,305;
core/newton_krylov.f:359).  The Nek5000 sources are NOT vendored in /root/reference (fork
github.com/nekStab/Nek5000, branch master, unpinned -- Nek5000clone.sh:3-7), so the routines below
restate the published algorithms [UPSTREAM]: coef.f geom1/geom2 (metrics), hmholtz.f axhelm/cggo,
navier1.f opgradt/cdtp, opdiv/multd, opbinv, cdabdtp, convect.f convect_new / convect_adj (dealiased
advection on GL(lxd)), dssum.f dssum, math.f glsc3, subs1.f compute_cfl.  Conventions: SURVEY.md App. E.1.

Parity status: pinned against the reference's shipped fixtures (tests/test_oracle_fixtures.py):
KAT-TG (BFS optimal perturbation -> optimal response, direct and adjoint), KAT-eig (cylinder leading
eigenpair), KAT-norm, KAT-steps, KAT-part.  Stopping norms of the iterative solvers are unpinned
(they do not affect the converged step).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import sem


def apply_1d(M, a, axis):
    """Contract matrix M (m, n) with `a` along `axis` (length n)."""
    out = np.tensordot(a, M, axes=([axis], [1]))
    return np.moveaxis(out, -1, axis)


class SEM:
    """Discrete operators on one (global or rank-local) set of elements.

    xyz  : (ldim, nel, lx1^ldim) GLL coordinates      glo : (nel, lx1^ldim) global node ids
    mask : (ldim, nel, lx1^ldim) Dirichlet masks      nglob_hint: number of global ids (optional)
    """

    def __init__(self, ldim, lx1, xyz, glo, mask, lxd=None, lx2=None):
        self.ldim, self.lx1 = ldim, lx1
        self.lx2 = lx2 or lx1 - 2
        self.lxd = lxd or 3 * lx1 // 2
        self.nel = xyz.shape[1]
        d = ldim
        self.eshape = (self.nel,) + (lx1,) * d
        self.eshape2 = (self.nel,) + (self.lx2,) * d
        self.eshaped = (self.nel,) + (self.lxd,) * d
        self.z, self.w = sem.gll(lx1)
        self.D = sem.deriv(self.z)
        self.zg, self.wg = sem.gl(self.lx2)
        self.J12 = sem.interp(self.zg, self.z)
        self.D12 = self.J12 @ self.D
        self.zd, self.wd = sem.gl(self.lxd)
        self.Jd = sem.interp(self.zd, self.z)
        self.Dd = self.Jd @ self.D
        self.glo = glo.reshape(self.nel, -1)
        self.nglob = int(self.glo.max()) + 1
        self.mask = mask.reshape((d,) + self.eshape)
        X = xyz.reshape((d,) + self.eshape)
        self.X = X
        # --- geometry on mesh 1 [UPSTREAM coef.f geom1: xrm1.., rxm1.. (Jacobian-included), jacm1, bm1, g1m1..]
        # axes: element data [e, (k,) j, i]; direction 0 = r acts on the LAST axis.
        Jm = np.empty(self.eshape + (d, d))            # Jm[..., c, i] = d x_c / d r_i
        for c in range(d):
            for i in range(d):
                Jm[..., c, i] = apply_1d(self.D, X[c], axis=-1 - i)
        self.jac = np.linalg.det(Jm)
        assert (self.jac > 0).all(), "non-positive Jacobian"
        Rm = self.jac[..., None, None] * np.linalg.inv(Jm)   # Rm[..., i, c] = J d r_i / d x_c
        self.R = np.moveaxis(Rm, (-2, -1), (0, 1))           # (d, d, nel, ...)
        self.w3 = self._tensor_w(self.w)
        self.bm1 = self.jac * self.w3
        self.G = np.einsum('ic...,jc...->ij...', self.R, self.R) * (self.w3 / self.jac)
        self.vol = self.bm1.sum()
        # assembled mass, multiplicity
        self.mult = 1.0 / self.dssum(np.ones(self.eshape))
        self.binv = 1.0 / self.dssum(self.bm1)               # binvm1 (unmasked, as Nek stores it)
        # --- mesh 2 (GL(lx1-2)) [UPSTREAM coef.f geom2: rxm2.. = interpolated rxm1.., bm2 = w3m2*jacm2]
        self.w32 = self._tensor_w(self.wg)
        self.R2 = np.stack([np.stack([self.to_m2(self.R[i, c]) for c in range(d)]) for i in range(d)])
        self.jac2 = self.to_m2(self.jac)
        self.bm2 = self.jac2 * self.w32
        self.vol2 = self.bm2.sum()
        # --- dealiasing mesh GL(lxd) [UPSTREAM convect.f set_dealias_rx: rx = w * interpolated rxm1]
        self.w3d = self._tensor_w(self.wd)
        self.Rd = np.stack([np.stack([self.to_fine(self.R[i, c]) * self.w3d for c in range(d)]) for i in range(d)])

    # ------------------------------------------------------------------ helpers
    def _tensor_w(self, w):
        out = w
        for _ in range(self.ldim - 1):
            out = np.multiply.outer(out, w)
        return out

    def _tensor_apply(self, mats, a):
        """mats[i] acts along direction i (i=0 is r = last axis)."""
        for i, M in enumerate(mats):
            a = apply_1d(M, a, axis=-1 - i)
        return a

    def to_m2(self, a):
        return self._tensor_apply([self.J12] * self.ldim, a)

    def to_fine(self, a):
        return self._tensor_apply([self.Jd] * self.ldim, a)

    def from_fine_T(self, a):
        return self._tensor_apply([self.Jd.T] * self.ldim, a)

    def grad_rst(self, a):
        return [apply_1d(self.D, a, axis=-1 - i) for i in range(self.ldim)]

    def fine_grad_rst(self, a):
        """r,s,t derivatives of the degree-(lx1-1) interpolant evaluated on GL(lxd) (= grad_rst on the fine
        mesh of the interpolated field [UPSTREAM convect.f convect_new], SURVEY App. D identity)."""
        out = []
        for i in range(self.ldim):
            mats = [self.Jd] * self.ldim
            mats[i] = self.Dd
            out.append(self._tensor_apply(mats, a))
        return out

    # ------------------------------------------------------------------ gather-scatter & reductions
    def dssum(self, a):
        """Direct-stiffness sum: every copy of a global node receives the sum over copies
        [UPSTREAM dssum.f dssum -> gslib gs_op(add)]."""
        g = self.glo.ravel()
        s = np.bincount(g, weights=a.ravel(), minlength=self.nglob)
        return s[g].reshape(a.shape)

    def glsc3(self, a, b, c):
        return float(np.sum(a * b * c))

    # ------------------------------------------------------------------ Helmholtz
    def axhelm(self, u, h1, h2):
        """w = (h1*A + h2*B) u, element-local [UPSTREAM hmholtz.f axhelm]."""
        d = self.ldim
        ur = self.grad_rst(u)
        out = h2 * self.bm1 * u
        for i in range(d):
            t = sum(self.G[i, j] * ur[j] for j in range(d))
            out = out + h1 * apply_1d(self.D.T, t, axis=-1 - i)
        return out

    def helm_diag(self, h1, h2):
        """Assembled diagonal of h1*A + h2*B (Jacobi preconditioner) [UPSTREAM hmholtz.f setprec].
        Exact diagonal including the mixed G_ij terms (Nek's setprec adds them only at element
        edges for deformed elements; ours is the true diagonal)."""
        d = self.ldim
        D2 = self.D ** 2
        dg = h2 * self.bm1.copy()
        for i in range(d):
            # sum_l D[l,p]^2 G_ii(l along direction i)
            dg = dg + h1 * apply_1d(D2.T, self.G[i, i], axis=-1 - i)
        # mixed terms: 2 * G_ij(p) D[p_i,p_i] D[p_j,p_j]
        dd = np.diag(self.D)
        for i in range(d):
            for j in range(i + 1, d):
                shp_i = [1] * (d + 1); shp_i[-1 - i] = self.lx1
                shp_j = [1] * (d + 1); shp_j[-1 - j] = self.lx1
                dg = dg + h1 * 2.0 * self.G[i, j] * dd.reshape(shp_i) * dd.reshape(shp_j)
        return self.dssum(dg)

    # ------------------------------------------------------------------ mesh-1 <-> mesh-2 operators
    def _d12(self, i):
        mats = [self.J12] * self.ldim
        mats[i] = self.D12
        return mats

    def opdiv(self, u):
        """Weak divergence M1 -> M2: sum_c D_c u_c [UPSTREAM navier1.f opdiv/multd]."""
        d = self.ldim
        out = 0.0
        for c in range(d):
            for i in range(d):
                out = out + self.R2[i, c] * self._tensor_apply(self._d12(i), u[c])
        return out * self.w32

    def opgradt(self, p):
        """Transpose of opdiv, M2 -> M1, un-assembled [UPSTREAM navier1.f opgradt/cdtp]."""
        d = self.ldim
        out = []
        wp = p * self.w32
        for c in range(d):
            acc = 0.0
            for i in range(d):
                acc = acc + self._tensor_apply([M.T for M in self._d12(i)], self.R2[i, c] * wp)
            out.append(acc)
        return np.stack(out)

    def opbinv(self, w):
        """mask, dssum, multiply by the inverse assembled mass [UPSTREAM navier1.f opbinv, h2inv=1]."""
        return np.stack([self.dssum(self.mask[c] * w[c]) * self.binv for c in range(self.ldim)])

    def cdabdtp(self, p):
        """E p = D B^-1 QQ^T D^T p [UPSTREAM navier1.f cdabdtp]."""
        return self.opdiv(self.opbinv(self.opgradt(p)))

    def e_diag(self):
        """Exact diagonal of E (Jacobi preconditioner for the pressure CG the north-star prescribes)."""
        d = self.ldim
        wb = np.stack([self.mask[c] * self.binv for c in range(d)])   # includes assembled-mass inverse
        out = np.zeros(self.eshape2)
        mats = {0: self.J12, 1: self.D12}
        # E_ii = sum_c sum_k D_c[i,k]^2 wb_c[k] ; D_c[i,k] = w2_i sum_dir R2[dir,c](i) prod_axis M_axis[i_axis,k_axis]
        # expand the square: sum over (dir, dir') of R2[dir,c] R2[dir',c] * (prod_axis M^dir_axis * M^dir'_axis) applied to wb_c
        for c in range(d):
            for a in range(d):
                for b in range(d):
                    ms = []
                    for ax in range(d):
                        Ma = self.D12 if ax == a else self.J12
                        Mb = self.D12 if ax == b else self.J12
                        ms.append(Ma * Mb)
                    out = out + self.R2[a, c] * self.R2[b, c] * self._tensor_apply(ms, wb[c])
        return out * self.w32 ** 2

    # ------------------------------------------------------------------ dealiased advection
    def _contravariant_fine(self, c):
        """(Rd . c_fine): the convecting field's contravariant components times fine weights."""
        d = self.ldim
        cf = [self.to_fine(c[k]) for k in range(d)]
        return [sum(self.Rd[i, k] * cf[k] for k in range(d)) for i in range(d)]

    def convop(self, c, u):
        """Mass-weighted dealiased (c . grad) u for one scalar u [UPSTREAM convect.f convect_new]."""
        cr = self._contravariant_fine(c)
        du = self.fine_grad_rst(u)
        return self.from_fine_T(sum(cr[i] * du[i] for i in range(self.ldim)))

    def advab_direct(self, up, ub):
        """B[(u'.grad)U + (U.grad)u'] per component [UPSTREAM perturb.f advabp]."""
        d = self.ldim
        crp = self._contravariant_fine(up)
        crb = self._contravariant_fine(ub)
        out = []
        for k in range(d):
            dU = self.fine_grad_rst(ub[k])
            du = self.fine_grad_rst(up[k])
            out.append(self.from_fine_T(sum(crp[i] * dU[i] + crb[i] * du[i] for i in range(d))))
        return np.stack(out)

    def advab_adjoint(self, up, ub):
        """B[(grad U)^T u' - (U.grad)u'] per component [UPSTREAM perturb.f advabp_adjoint / convect_adj]."""
        d = self.ldim
        crb = self._contravariant_fine(ub)
        upf = [self.to_fine(up[j]) for j in range(d)]
        dU = [self.fine_grad_rst(ub[j]) for j in range(d)]
        out = []
        for i in range(d):
            du = self.fine_grad_rst(up[i])
            conv = sum(crb[k] * du[k] for k in range(d))
            # sum_j u'_j dU_j/dx_i * (w J) = sum_j u'_j sum_k Rd[k,i] dU_j/dr_k
            gt = sum(upf[j] * sum(self.Rd[k, i] * dU[j][k] for k in range(d)) for j in range(d))
            out.append(self.from_fine_T(gt - conv))
        return np.stack(out)

    # ------------------------------------------------------------------ CFL
    def cfl_sum(self, u):
        """max over points of sum_i |u . grad r_i| / dr_i   (dt = 1) [UPSTREAM subs1.f compute_cfl]."""
        d = self.ldim
        z = self.z
        dr = np.empty(self.lx1)
        dr[0] = z[1] - z[0]; dr[-1] = z[-1] - z[-2]; dr[1:-1] = 0.5 * (z[2:] - z[:-2])
        tot = 0.0
        for i in range(d):
            ur = sum(u[c] * self.R[i, c] for c in range(d)) / self.jac
            shp = [1] * (d + 1); shp[-1 - i] = self.lx1
            tot = tot + np.abs(ur / dr.reshape(shp))
        return float(tot.max())

    # ------------------------------------------------------------------ assembled sparse operators (direct solves)
    def _elem_matrix(self, fn, nin, nout, chunk=64):
        """Dense per-element matrices of an element-local linear map (nel, nout, nin)."""
        out = np.empty((self.nel, nout, nin))
        eye = np.eye(nin)
        return out, eye

    def helm_sparse(self, h1, h2, comp=0, mask=None):
        """Assembled, masked Helmholtz matrix on global nodes (Dirichlet rows/cols replaced by identity).  mask: an explicit
        Dirichlet mask (the scalar's tmask) instead of the velocity component's."""
        d = self.ldim
        npt = self.lx1 ** d
        # element matrices via tensor structure: apply axhelm to unit vectors, one basis fn at a time
        Ke = np.empty((self.nel, npt, npt))
        for k in range(npt):
            u = np.zeros((self.nel, npt)); u[:, k] = 1.0
            Ke[:, :, k] = self.axhelm(u.reshape(self.eshape), h1, h2).reshape(self.nel, npt)
        rows = np.repeat(self.glo[:, :, None], npt, axis=2).ravel()
        cols = np.repeat(self.glo[:, None, :], npt, axis=1).ravel()
        K = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(self.nglob, self.nglob)).tocsr()
        free = np.zeros(self.nglob)
        np.maximum.at(free, self.glo.ravel(), (self.mask[comp] if mask is None else mask).ravel())
        Fm = sp.diags(free)
        K = Fm @ K @ Fm + sp.diags(1.0 - free)
        return K.tocsc(), free

    def div_sparse(self):
        """Sparse D_c : global velocity nodes -> mesh-2 points (list over c)."""
        d = self.ldim
        npt, np2 = self.lx1 ** d, self.lx2 ** d
        n2 = self.nel * np2
        out = []
        for c in range(d):
            De = np.empty((self.nel, np2, npt))
            for k in range(npt):
                u = np.zeros((d, self.nel, npt)); u[c, :, k] = 1.0
                De[:, :, k] = self.opdiv(u.reshape((d,) + self.eshape)).reshape(self.nel, np2)
            rows = np.repeat((np.arange(self.nel)[:, None] * np2 + np.arange(np2)[None, :])[:, :, None], npt, axis=2).ravel()
            cols = np.repeat(self.glo[:, None, :], np2, axis=1).ravel()
            out.append(sp.coo_matrix((De.ravel(), (rows, cols)), shape=(n2, self.nglob)).tocsr())
        return out

    def e_sparse(self):
        d = self.ldim
        Ds = self.div_sparse()
        E = None
        for c in range(d):
            free = np.zeros(self.nglob)
            np.maximum.at(free, self.glo.ravel(), self.mask[c].ravel())
            bg = np.zeros(self.nglob)
            bg[self.glo.ravel()] = self.binv.ravel()
            W = sp.diags(free * bg)
            T = Ds[c] @ W @ Ds[c].T
            E = T if E is None else E + T
        return E.tocsc()

    def to_global(self, a):
        out = np.zeros(self.nglob)
        out[self.glo.ravel()] = a.ravel()
        return out

    def from_global(self, g):
        return g[self.glo.ravel()].reshape(self.eshape)
