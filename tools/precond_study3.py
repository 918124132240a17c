"""Scratch study 3 (CPU, oracle): overlapping (depth-1) FDM Schwarz on the extended (lx2+2)^2 grid vs the non-overlapping blocks."""
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
E = s.e_sparse().tocsr()
n2 = E.shape[0]; nel = s.nel; L2 = s.lx2; np2 = L2 ** 2; LE = L2 + 2
ae = lambda p: (E @ p.ravel()).reshape(p.shape)
M3 = pmg.PMG(s, nagg=64, apply_e=ae)
rng = np.random.default_rng(0)
u = rng.standard_normal((2,) + s.eshape)
u = np.stack([s.dssum(u[k]) * s.mult * s.mask[k] for k in range(2)])
b = -s.opdiv(u)
x, it = pmg.pcg(ae, M3.apply, b, 1e-8); print("non-overlapping 3-level:", it)

# ---- neighbours
G = c.glo.reshape(nel, lx1, lx1)
idx2 = np.arange(n2).reshape(nel, L2, L2)
def face_nodes(e, f):   # global ids along the face, in my tangential order; f: 0 i=0 (west), 1 i=N (east), 2 j=0 (south), 3 j=N (north)
    return [G[e, :, 0], G[e, :, -1], G[e, 0, :], G[e, -1, :]][f]
def face_layer(e, f):   # pressure indices of the layer adjacent to face f, in tangential order
    return [idx2[e, :, 0], idx2[e, :, -1], idx2[e, 0, :], idx2[e, -1, :]][f]
fmap = {}
for e in range(nel):
    for f in range(4):
        ids = face_nodes(e, f); key = (min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2]))
        fmap.setdefault(key, []).append((e, f))
X = c.xyz.reshape(2, nel, lx1, lx1); mid = lx1 // 2
hh = np.zeros((nel, 2))
for e in range(nel):
    hh[e, 0] = np.linalg.norm(X[:, e, :, -1].mean(1) - X[:, e, :, 0].mean(1))
    hh[e, 1] = np.linalg.norm(X[:, e, -1, :].mean(1) - X[:, e, 0, :].mean(1))
ext = -np.ones((nel, LE, LE), dtype=np.int64)       # [j, i] extended index sets
nb_h = np.zeros((nel, 4))                            # neighbour length normal to face f (0 = none)
nb_mask = np.zeros((nel, 4))                         # Dirichlet (0) / free (1) at a face without neighbour
m0 = s.mask[0].reshape(nel, lx1, lx1)
for e in range(nel):
    ext[e, 1:-1, 1:-1] = idx2[e]
    for f in range(4):
        ids = face_nodes(e, f); key = (min(ids[0], ids[-1]), max(ids[0], ids[-1]), min(ids[1], ids[-2]))
        other = [t for t in fmap[key] if t[0] != e]
        if not other:
            nb_mask[e, f] = [m0[e, mid, 0], m0[e, mid, -1], m0[e, 0, mid], m0[e, -1, mid]][f]
            continue
        e2, f2 = other[0]
        lay = face_layer(e2, f2)
        if face_nodes(e2, f2)[0] != ids[0]: lay = lay[::-1]
        if f == 0: ext[e, 1:-1, 0] = lay
        elif f == 1: ext[e, 1:-1, -1] = lay
        elif f == 2: ext[e, 0, 1:-1] = lay
        else: ext[e, -1, 1:-1] = lay
        nb_h[e, f] = hh[e2, 0 if f2 < 2 else 1]
# ---- 1-D extended operators
w, w2 = s.w, s.wg; D12, J12 = s.D12, s.J12
def ext_1d(h, hl, hr, ml_mask, mr_mask):
    """A, M (LE x LE) on [last GL of left | GL of centre | first GL of right]; absent neighbour -> decoupled row."""
    els = [hl, h, hr]
    nvel = 3 * (lx1 - 1) + 1
    mass = np.zeros(nvel)
    for k, he in enumerate(els):
        if he > 0: mass[k * (lx1 - 1): k * (lx1 - 1) + lx1] += w * he / 2
    # patch ends: assume a further neighbour of the same size
    if hl > 0: mass[0] *= 2
    if hr > 0: mass[-1] *= 2
    W = np.where(mass > 0, 1.0 / np.where(mass > 0, mass, 1), 0.0)
    # domain boundary at the centre element's end when there is no neighbour: mask decides
    if hl == 0: W[lx1 - 1] = ml_mask / (w[0] * h / 2)
    if hr == 0: W[2 * (lx1 - 1)] = mr_mask / (w[-1] * h / 2)
    A = np.zeros((3 * L2, 3 * L2)); M = np.zeros((3 * L2, 3 * L2))
    BD = np.zeros((3 * L2, nvel)); BJ = np.zeros((3 * L2, nvel))
    for k, he in enumerate(els):
        if he > 0:
            sl = slice(k * (lx1 - 1), k * (lx1 - 1) + lx1)
            BD[k * L2:(k + 1) * L2, sl] = (w2 * he / 2)[:, None] * D12 * (2 / he)
            BJ[k * L2:(k + 1) * L2, sl] = (w2 * he / 2)[:, None] * J12
    A = (BD * W) @ BD.T; M = (BJ * W) @ BJ.T
    sel = np.arange(L2 - 1, 2 * L2 + 1)
    A = A[np.ix_(sel, sel)]; M = M[np.ix_(sel, sel)]
    for k, he in ((0, hl), (LE - 1, hr)):
        if he == 0:
            A[k, :] = 0; A[:, k] = 0; M[k, :] = 0; M[:, k] = 0; M[k, k] = 1.0; A[k, k] = 1e30   # decoupled, infinitely stiff
    return A, M
Sx = np.zeros((nel, LE, LE)); Sy = np.zeros((nel, LE, LE)); lx = np.zeros((nel, LE)); ly = np.zeros((nel, LE))
for e in range(nel):
    A, M = ext_1d(hh[e, 0], nb_h[e, 0], nb_h[e, 1], nb_mask[e, 0], nb_mask[e, 1]); lam, S = sla.eigh(A, M); Sx[e] = S; lx[e] = lam
    A, M = ext_1d(hh[e, 1], nb_h[e, 2], nb_h[e, 3], nb_mask[e, 2], nb_mask[e, 3]); lam, S = sla.eigh(A, M); Sy[e] = S; ly[e] = lam
den = lx[:, None, :] + ly[:, :, None]       # E = Ax (x) My + Mx (x) Ay in physical units -> eigenvalues add
deninv = np.where(den < 1e20, 1.0 / den, 0.0)
have = ext >= 0
def fdm_ext(r):
    rr = r.ravel()
    re = np.where(have, rr[np.maximum(ext, 0)], 0.0)          # [e, j, i]
    t = np.einsum('eIi,eJj,eJI->eji', Sx, Sy, re) * deninv
    ze = np.einsum('eIi,eJj,eji->eJI', Sx, Sy, t)
    z = np.zeros(n2)
    np.add.at(z, ext[have], ze[have])
    return z.reshape(r.shape)
a_ = rng.standard_normal(s.eshape2); b_ = rng.standard_normal(s.eshape2)
print("sym check", np.sum(a_ * fdm_ext(b_)), np.sum(b_ * fdm_ext(a_)), "pos", np.sum(a_ * fdm_ext(a_)))
def three_level(loc):
    def M(r):
        rc = M3.restrict_q1(r)
        xv = M3._assemble_v(rc) / M3.d1
        x2 = M3.A2inv @ np.bincount(M3.agg, weights=rc.sum(1), minlength=M3.nagg)
        return loc(r) + M3.prolong_q1(xv[M3.vid]) + x2[M3.agg].reshape(-1, 1, 1)
    return M
x, it = pmg.pcg(ae, three_level(fdm_ext), b, 1e-8); print("overlapping ext-FDM 3-level:", it)
x, it = pmg.pcg(ae, three_level(lambda r: 0.5 * fdm_ext(r)), b, 1e-8); print("overlapping ext-FDM x0.5, 3-level:", it)
x, it = pmg.pcg(ae, three_level(lambda r: 0.5 * (fdm_ext(r) + M3.fdm(r))), b, 1e-8); print("avg(ext, block) 3-level:", it)
