"""Scratch study 5 (CPU, oracle): extended-grid FDM Schwarz with overlap ONLY in each element's thin direction (1-D extension:
exact tensor structure, no corner nodes), on the stretched cylinder mesh.  Follow-up of studies 3/4."""
import sys, os
import numpy as np, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekstab_b200 import cases
from oracle.ops import SEM
from oracle import pmg

lx1 = int(sys.argv[1]) if len(sys.argv) > 1 else 6
g = np.load("tests/golden/cyl.npz")
c = cases.cylinder_case(g, lx1=lx1, sponge=False)
s = SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)
E = s.e_sparse().tocsr(); ae = lambda p: (E @ p.ravel()).reshape(p.shape)
n2 = E.shape[0]; nel = s.nel; L2 = s.lx2; LE = L2 + 2
M3 = pmg.PMG(s, nagg=64, apply_e=ae)
rng = np.random.default_rng(0)
u = rng.standard_normal((2,) + s.eshape); u = np.stack([s.dssum(u[k]) * s.mult * s.mask[k] for k in range(2)])
b = -s.opdiv(u)
def coarse(r):
    rc = M3.restrict_q1(r); xv = M3._assemble_v(rc) / M3.d1
    x2 = M3.A2inv @ np.bincount(M3.agg, weights=rc.sum(1), minlength=M3.nagg)
    return M3.prolong_q1(xv[M3.vid]) + x2[M3.agg].reshape(-1, 1, 1)
print("non-overlapping FDM 3-level:", pmg.pcg(ae, M3.apply, b, 1e-8)[1])
G = c.glo.reshape(nel, lx1, lx1); idx2 = np.arange(n2).reshape(nel, L2, L2)
fn = lambda e, f: [G[e, :, 0], G[e, :, -1], G[e, 0, :], G[e, -1, :]][f]
fl = lambda e, f: [idx2[e, :, 0], idx2[e, :, -1], idx2[e, 0, :], idx2[e, -1, :]][f]
fmap = {}
for e in range(nel):
    for f in range(4):
        ids = fn(e, f); fmap.setdefault((min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2])), []).append((e, f))
X = c.xyz.reshape(2, nel, lx1, lx1); mid = lx1 // 2
hh = np.stack([np.linalg.norm(X[:, :, :, -1].mean(2) - X[:, :, :, 0].mean(2), axis=0), np.linalg.norm(X[:, :, -1, :].mean(2) - X[:, :, 0, :].mean(2), axis=0)], 1)
print("aspect ratio h_max/h_min: median %.2f, 90%% %.2f, max %.2f" % tuple(np.percentile(hh.max(1) / hh.min(1), [50, 90, 100])))
m0 = s.mask[0].reshape(nel, lx1, lx1); mloc = (1.0 / (s.binv * s.bm1)).reshape(nel, lx1, lx1)
w, w2 = s.w, s.wg; D12, J12 = s.D12, s.J12
def ext_1d(h, hl, hr, wl_end, wr_end):
    els = [hl, h, hr]; nvel = 3 * (lx1 - 1) + 1; mass = np.zeros(nvel)
    for k, he in enumerate(els):
        if he > 0: mass[k * (lx1 - 1): k * (lx1 - 1) + lx1] += w * he / 2
    if hl > 0: mass[0] *= 2
    if hr > 0: mass[-1] *= 2
    W = np.where(mass > 0, 1.0 / np.where(mass > 0, mass, 1), 0.0)
    if hl == 0: W[lx1 - 1] = wl_end
    if hr == 0: W[2 * (lx1 - 1)] = wr_end
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
def build(thresh):
    """extend element e in direction d only if h_other/h_d >= thresh (thin direction)"""
    ext = -np.ones((nel, LE, LE), dtype=np.int64); Sx = np.zeros((nel, LE, LE)); Sy = np.zeros((nel, LE, LE)); lx = np.zeros((nel, LE)); ly = np.zeros((nel, LE))
    next = 0
    for e in range(nel):
        ext[e, 1:-1, 1:-1] = idx2[e]
        for d in range(2):
            do_ext = hh[e, 1 - d] / hh[e, d] >= thresh
            hn = [0.0, 0.0]; wend = [0.0, 0.0]
            for side in range(2):
                f = 2 * d + side
                ids = fn(e, f); other = [t for t in fmap[(min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2]))] if t[0] != e]
                node = (mid, 0 if side == 0 else lx1 - 1) if d == 0 else (0 if side == 0 else lx1 - 1, mid)
                # block-type end weight: mask / (w * h/2 * assembled/local mass)
                wend[side] = m0[e][node] / (w[0] * hh[e, d] / 2 * mloc[e][node])
                if other and do_ext:
                    e2, f2 = other[0]; lay = fl(e2, f2)
                    if fn(e2, f2)[0] != ids[0]: lay = lay[::-1]
                    if f == 0: ext[e, 1:-1, 0] = lay
                    elif f == 1: ext[e, 1:-1, -1] = lay
                    elif f == 2: ext[e, 0, 1:-1] = lay
                    else: ext[e, -1, 1:-1] = lay
                    hn[side] = hh[e2, 0 if f2 < 2 else 1]
            next += do_ext
            A, M = ext_1d(hh[e, d], hn[0], hn[1], wend[0], wend[1]); lam, S = sla.eigh(A, M)
            if d == 0: Sx[e] = S; lx[e] = lam
            else: Sy[e] = S; ly[e] = lam
    den = lx[:, None, :] + ly[:, :, None]; deninv = np.where(den < 1e20, 1.0 / den, 0.0); have = ext >= 0
    def fdm_ext(r):
        rr = r.ravel(); re = np.where(have, rr[np.maximum(ext, 0)], 0.0)
        t = np.einsum('eIi,eJj,eJI->eji', Sx, Sy, re) * deninv
        ze = np.einsum('eIi,eJj,eji->eJI', Sx, Sy, t)
        z = np.zeros(n2); np.add.at(z, ext[have], ze[have]); return z.reshape(r.shape)
    return fdm_ext, next
for thresh in (1e9, 3.0, 2.0, 1.5, 1.0, 0.0):
    f, nx = build(thresh)
    print("extend when aspect >= %-5g: %4d extended directions, iterations %d" % (thresh, nx, pmg.pcg(ae, lambda r: f(r) + coarse(r), b, 1e-8)[1]))

# exact local inverses on the same 1-D extended index sets (thin direction only, aspect >= 1) and on both directions
def exact_sets(thresh):
    sets = []
    for e in range(nel):
        ids_all = [idx2[e].ravel()]
        for d in range(2):
            if hh[e, 1 - d] / hh[e, d] < thresh: continue
            for side in range(2):
                f = 2 * d + side
                ids = fn(e, f); other = [t for t in fmap[(min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2]))] if t[0] != e]
                if other: ids_all.append(fl(*other[0]))
        sets.append(np.concatenate(ids_all))
    invs = [np.linalg.inv(E[ss][:, ss].toarray()) for ss in sets]
    def ap(r):
        rr = r.ravel(); z = np.zeros(n2)
        for ss, Ai in zip(sets, invs): z[ss] += Ai @ rr[ss]
        return z.reshape(r.shape)
    return ap
for thresh in (1.0, 0.0):
    ap = exact_sets(thresh)
    print("EXACT local inverses, extend when aspect >= %g: iterations %d" % (thresh, pmg.pcg(ae, lambda r: ap(r) + coarse(r), b, 1e-8)[1]))

# ---- stronger Q1-level solve: symmetric V-cycle on A_c = P^T E P (damped-Jacobi smoothing + exact aggregate correction)
import scipy.sparse as sp, scipy.sparse.linalg as spla
np2 = L2 * L2
rows = (np.arange(nel)[:, None, None] * np2 + np.arange(np2)[None, None, :]).repeat(4, 1).ravel()
cols = M3.vid[:, :, None].repeat(np2, 2).ravel()
P = sp.coo_matrix((np.tile(M3.phi.reshape(4, -1), (nel, 1, 1)).ravel(), (rows, cols)), shape=(n2, M3.nv)).tocsr()
Ac = (P.T @ E @ P).tocsr(); dA = Ac.diagonal()
vagg = np.zeros(M3.nv, dtype=np.int64); vagg[M3.vid.ravel()] = np.repeat(M3.agg, 4)
P2 = sp.coo_matrix((np.ones(M3.nv), (np.arange(M3.nv), vagg)), shape=(M3.nv, M3.nagg)).tocsr()
A2 = (P2.T @ Ac @ P2).toarray(); A2i = np.linalg.inv(A2)
lu = spla.splu(Ac.tocsc())
def vcycle(nu, omega=0.7):
    def B(rc):
        x = np.zeros_like(rc)
        for _ in range(nu): x = x + omega * (rc - Ac @ x) / dA
        x = x + P2 @ (A2i @ (P2.T @ (rc - Ac @ x)))
        for _ in range(nu): x = x + omega * (rc - Ac @ x) / dA
        return x
    return B
ap_ex = exact_sets(0.0)
for name, loc in (("non-overlapping FDM blocks", M3.fdm), ("exact overlapping cross sets", ap_ex)):
    print(name)
    print("   exact Q1 solve:", pmg.pcg(ae, lambda r: loc(r) + (P @ lu.solve(P.T @ r.ravel())).reshape(r.shape), b, 1e-8)[1])
    for nu in (1, 2, 4):
        B = vcycle(nu)
        print("   Q1 V-cycle, %d+%d damped-Jacobi sweeps + aggregate solve:" % (nu, nu), pmg.pcg(ae, lambda r: loc(r) + (P @ B(P.T @ r.ravel())).reshape(r.shape), b, 1e-8)[1])
