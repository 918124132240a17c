"""Scratch study 4 (CPU, oracle): is the extended-grid FDM Schwarz of study 3 limited by non-separability of the mesh or by
the zero-padded corners?  Uniform Cartesian box (separable by construction) vs deformed box; exact cross-set solves for reference."""
import sys, os
import numpy as np, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekstab_b200 import cases
from oracle.ops import SEM
from oracle import pmg

def study(c, label):
    lx1 = c.lx1
    s = SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)
    E = s.e_sparse().tocsr(); ae = lambda p: (E @ p.ravel()).reshape(p.shape)
    n2 = E.shape[0]; nel = s.nel; L2 = s.lx2; LE = L2 + 2
    M3 = pmg.PMG(s, nagg=max(1, nel // 16), apply_e=ae)
    rng = np.random.default_rng(0)
    u = rng.standard_normal((2,) + s.eshape); u = np.stack([s.dssum(u[k]) * s.mult * s.mask[k] for k in range(2)])
    b = -s.opdiv(u)
    def coarse(r):
        rc = M3.restrict_q1(r); xv = M3._assemble_v(rc) / M3.d1
        x2 = M3.A2inv @ np.bincount(M3.agg, weights=rc.sum(1), minlength=M3.nagg)
        return M3.prolong_q1(xv[M3.vid]) + x2[M3.agg].reshape(-1, 1, 1)
    print(label, "nel", nel, "| non-overlapping FDM 3-level:", pmg.pcg(ae, M3.apply, b, 1e-8)[1], end=" | ")
    G = c.glo.reshape(nel, lx1, lx1); idx2 = np.arange(n2).reshape(nel, L2, L2)
    fn = lambda e, f: [G[e, :, 0], G[e, :, -1], G[e, 0, :], G[e, -1, :]][f]
    fl = lambda e, f: [idx2[e, :, 0], idx2[e, :, -1], idx2[e, 0, :], idx2[e, -1, :]][f]
    fmap = {}
    for e in range(nel):
        for f in range(4):
            ids = fn(e, f); fmap.setdefault((min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2])), []).append((e, f))
    X = c.xyz.reshape(2, nel, lx1, lx1); mid = lx1 // 2
    hh = np.stack([np.linalg.norm(X[:, :, :, -1].mean(2) - X[:, :, :, 0].mean(2), axis=0), np.linalg.norm(X[:, :, -1, :].mean(2) - X[:, :, 0, :].mean(2), axis=0)], 1)
    ext = -np.ones((nel, LE, LE), dtype=np.int64); nb_h = np.zeros((nel, 4)); nb_mask = np.zeros((nel, 4)); nb_e = -np.ones((nel, 4), dtype=int)
    m0 = s.mask[0].reshape(nel, lx1, lx1)
    for e in range(nel):
        ext[e, 1:-1, 1:-1] = idx2[e]
        for f in range(4):
            ids = fn(e, f); other = [t for t in fmap[(min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2]))] if t[0] != e]
            if not other:
                nb_mask[e, f] = [m0[e, mid, 0], m0[e, mid, -1], m0[e, 0, mid], m0[e, -1, mid]][f]; continue
            e2, f2 = other[0]; lay = fl(e2, f2)
            if fn(e2, f2)[0] != ids[0]: lay = lay[::-1]
            if f == 0: ext[e, 1:-1, 0] = lay
            elif f == 1: ext[e, 1:-1, -1] = lay
            elif f == 2: ext[e, 0, 1:-1] = lay
            else: ext[e, -1, 1:-1] = lay
            nb_h[e, f] = hh[e2, 0 if f2 < 2 else 1]; nb_e[e, f] = e2
    have = ext >= 0
    # exact inverses on the cross-shaped sets
    invs = []
    for e in range(nel):
        ss = ext[e][have[e]]; invs.append(np.linalg.inv(E[ss][:, ss].toarray()))
    def exact_cross(r):
        rr = r.ravel(); z = np.zeros(n2)
        for e in range(nel):
            ss = ext[e][have[e]]; z[ss] += invs[e] @ rr[ss]
        return z.reshape(r.shape)
    print("exact cross-set overlap:", pmg.pcg(ae, lambda r: exact_cross(r) + coarse(r), b, 1e-8)[1], end=" | ")
    w, w2 = s.w, s.wg; D12, J12 = s.D12, s.J12
    def ext_1d(h, hl, hr, ml_mask, mr_mask, far=2.0):
        els = [hl, h, hr]; nvel = 3 * (lx1 - 1) + 1; mass = np.zeros(nvel)
        for k, he in enumerate(els):
            if he > 0: mass[k * (lx1 - 1): k * (lx1 - 1) + lx1] += w * he / 2
        if hl > 0: mass[0] *= far
        if hr > 0: mass[-1] *= far
        W = np.where(mass > 0, 1.0 / np.where(mass > 0, mass, 1), 0.0)
        if hl == 0: W[lx1 - 1] = ml_mask / (w[0] * h / 2)
        if hr == 0: W[2 * (lx1 - 1)] = mr_mask / (w[-1] * h / 2)
        BD = np.zeros((3 * L2, nvel)); BJ = np.zeros((3 * L2, nvel))
        for k, he in enumerate(els):
            if he > 0:
                sl = slice(k * (lx1 - 1), k * (lx1 - 1) + lx1)
                BD[k * L2:(k + 1) * L2, sl] = (w2 * he / 2)[:, None] * D12 * (2 / he); BJ[k * L2:(k + 1) * L2, sl] = (w2 * he / 2)[:, None] * J12
        A = (BD * W) @ BD.T; M = (BJ * W) @ BJ.T
        sel = np.arange(L2 - 1, 2 * L2 + 1); A = A[np.ix_(sel, sel)]; M = M[np.ix_(sel, sel)]
        for k, he in ((0, hl), (LE - 1, hr)):
            if he == 0:
                A[k, :] = 0; A[:, k] = 0; M[k, :] = 0; M[:, k] = 0; M[k, k] = 1.0; A[k, k] = 1e30
        return A, M
    Sx = np.zeros((nel, LE, LE)); Sy = np.zeros((nel, LE, LE)); lx = np.zeros((nel, LE)); ly = np.zeros((nel, LE))
    for e in range(nel):
        A, M = ext_1d(hh[e, 0], nb_h[e, 0], nb_h[e, 1], nb_mask[e, 0], nb_mask[e, 1]); lam, S = sla.eigh(A, M); Sx[e] = S; lx[e] = lam
        A, M = ext_1d(hh[e, 1], nb_h[e, 2], nb_h[e, 3], nb_mask[e, 2], nb_mask[e, 3]); lam, S = sla.eigh(A, M); Sy[e] = S; ly[e] = lam
    den = lx[:, None, :] + ly[:, :, None]; deninv = np.where(den < 1e20, 1.0 / den, 0.0)
    def fdm_ext(r):
        rr = r.ravel(); re = np.where(have, rr[np.maximum(ext, 0)], 0.0)
        t = np.einsum('eIi,eJj,eJI->eji', Sx, Sy, re) * deninv
        ze = np.einsum('eIi,eJj,eji->eJI', Sx, Sy, t)
        z = np.zeros(n2); np.add.at(z, ext[have], ze[have]); return z.reshape(r.shape)
    print("ext-FDM (corners zero-padded):", pmg.pcg(ae, lambda r: fdm_ext(r) + coarse(r), b, 1e-8)[1])

study(cases.box_case(12, 12, 6, lxy=(2.0, 2.0), outflow=True, deform=0.0), "uniform box ")
study(cases.box_case(12, 12, 6, lxy=(2.0, 2.0), outflow=True, deform=0.08), "deformed box")
