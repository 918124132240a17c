"""Scratch study (CPU, oracle): PCG iteration counts on E for candidate pressure preconditioners."""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekstab_b200 import cases
from oracle.ops import SEM

lx1 = int(sys.argv[1]) if len(sys.argv) > 1 else 6
g = np.load("tests/golden/cyl.npz")
c = cases.cylinder_case(g, lx1=lx1, sponge=False)
s = SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)
t = time.time(); E = s.e_sparse().tocsr(); print("E", E.shape, E.nnz, time.time() - t)
n2 = E.shape[0]; nel = s.nel; np2 = s.lx2 ** 2
rng = np.random.default_rng(0)
# rhs: divergence of a random smooth-ish velocity
u = rng.standard_normal((2,) + s.eshape)
u = np.stack([s.dssum(u[k]) * s.mult * s.mask[k] for k in range(2)])
b = -s.opdiv(u).ravel()

def pcg(M, tol=1e-8, maxit=20000):
    x = np.zeros(n2); r = b.copy(); p = np.zeros(n2); rtz1 = 1.0
    r0 = np.linalg.norm(r)
    for it in range(maxit):
        z = M(r); rtz2 = rtz1; rtz1 = z @ r
        if np.linalg.norm(r) <= tol * r0: return it
        beta = 0 if it == 0 else rtz1 / rtz2
        p = z + beta * p; w = E @ p; alpha = rtz1 / (w @ p); x += alpha * p; r -= alpha * w
    return maxit

dinv = 1.0 / E.diagonal()
print("jacobi", pcg(lambda r: dinv * r))

# element blocks (exact)
blocks = np.stack([np.linalg.inv(E[e*np2:(e+1)*np2, e*np2:(e+1)*np2].toarray()) for e in range(nel)])
def blk(r): return np.einsum('eij,ej->ei', blocks, r.reshape(nel, np2)).ravel()
print("block-jacobi exact", pcg(blk))

# coarse space: Q1 vertex functions on the element vertex mesh, evaluated at GL points
vert = c.glo.reshape(nel, lx1, lx1)[:, [0, 0, -1, -1], [0, -1, 0, -1]]     # (nel,4) global ids of corners
uv, vid = np.unique(vert, return_inverse=True); vid = vid.reshape(nel, 4); nv = uv.size
zg = s.zg; l0 = (1 - zg) / 2; l1 = (1 + zg) / 2
phi = np.stack([np.outer(a, b_).ravel() for a in (l0, l1) for b_ in (l0, l1)])   # (4, np2) index [j(y), i(x)] corner order (j,i)
rows = (np.arange(nel)[:, None, None] * np2 + np.arange(np2)[None, None, :]).repeat(4, 1).ravel()
cols = vid[:, :, None].repeat(np2, 2).ravel()
P = sp.coo_matrix((np.tile(phi, (nel, 1, 1)).ravel(), (rows, cols)), shape=(n2, nv)).tocsr()
Ac = (P.T @ E @ P).tocsc(); print("coarse", nv, Ac.nnz)
lu = spla.splu(Ac)
def crs(r): return P @ lu.solve(P.T @ r)
print("jacobi + coarse (additive)", pcg(lambda r: dinv * r + crs(r)))
print("block exact + coarse (additive)", pcg(lambda r: blk(r) + crs(r)))
def hybrid(loc):
    def M(r):
        z = loc(r); r1 = r - E @ z; z = z + crs(r1); r2 = r - E @ z; return z + loc(r2)
    return M
print("block exact, coarse, block (multiplicative symmetric)", pcg(hybrid(blk)))
# piecewise-constant coarse space (P0 per element)
P0 = sp.coo_matrix((np.ones(n2), (np.arange(n2), np.arange(n2) // np2)), shape=(n2, nel)).tocsr()
A0 = (P0.T @ E @ P0).tocsc(); lu0 = spla.splu(A0)
def crs0(r): return P0 @ lu0.solve(P0.T @ r)
print("block exact + P0 coarse", pcg(lambda r: blk(r) + crs0(r)))
print("block exact + P0 + Q1 coarse", pcg(lambda r: blk(r) + crs0(r) + crs(r)))

# ---- overlapping Schwarz, exact local solves: element + one layer of pressure nodes from each face neighbour
L2 = s.lx2
G = c.glo.reshape(nel, lx1, lx1)
def face_ids(e, f):  # f: 0 = j=0 (south), 1 = j=-1, 2 = i=0, 3 = i=-1 ; return sorted tuple of end ids + interior
    a = [G[e, 0, :], G[e, -1, :], G[e, :, 0], G[e, :, -1]][f]
    return (min(a[0], a[-1]), max(a[0], a[-1]), a[lx1 // 2] if lx1 % 2 else min(a[1], a[-2]))
fmap = {}
for e in range(nel):
    for f in range(4):
        fmap.setdefault(face_ids(e, f), []).append((e, f))
idx2 = np.arange(n2).reshape(nel, L2, L2)
def layer(e, f, depth=1):
    if f == 0: return idx2[e, :depth, :].ravel()
    if f == 1: return idx2[e, L2 - depth:, :].ravel()
    if f == 2: return idx2[e, :, :depth].ravel()
    return idx2[e, :, L2 - depth:].ravel()
for depth in (1, 2):
    sets = []
    for e in range(nel):
        ids = [idx2[e].ravel()]
        for f in range(4):
            for (e2, f2) in fmap[face_ids(e, f)]:
                if e2 != e: ids.append(layer(e2, f2, depth))
        sets.append(np.concatenate(ids))
    Ec = E.tocsc()
    invs = [np.linalg.inv(E[ss][:, ss].toarray()) for ss in sets]
    cnt = np.zeros(n2)
    for ss in sets: cnt[ss] += 1
    def osch(r, w=None):
        z = np.zeros(n2)
        for ss, Ai in zip(sets, invs): z[ss] += Ai @ r[ss]
        return z
    sq = 1 / np.sqrt(cnt)
    print("depth", depth, "overlap exact (additive)", pcg(osch))
    print("depth", depth, "overlap exact + Q1 coarse", pcg(lambda r: osch(r) + crs(r)))
    print("depth", depth, "overlap exact sym-weighted + Q1 coarse", pcg(lambda r: sq * osch(sq * r) + crs(r)))

# ---- FDM approximation of the element blocks from geometry
import scipy.linalg as sla
X = c.xyz.reshape(2, nel, lx1, lx1)
w = s.w; w2 = s.wg
BD = (w2[:, None] * s.D12); BJ = (w2[:, None] * s.J12)
mloc = 1.0 / (s.binv * s.bm1)             # assembled/local mass ratio (nel,lx1,lx1)
m0 = s.mask[0].reshape(nel, lx1, lx1)     # all comps share masks here
fdm_S = np.zeros((nel, 2, L2, L2)); fdm_lam = np.zeros((nel, 2, L2))
hh = np.zeros((nel, 2))
mid = lx1 // 2
for e in range(nel):
    # lengths: distance between opposite face mid-points
    pxm = X[:, e, mid, 0]; pxp = X[:, e, mid, -1]; pym = X[:, e, 0, mid]; pyp = X[:, e, -1, mid]
    hx = np.linalg.norm(pxp - pxm); hy = np.linalg.norm(pyp - pym); hh[e] = hx, hy
    for d in range(2):
        if d == 0: ml, mr, kl, kr = mloc[e, mid, 0], mloc[e, mid, -1], m0[e, mid, 0], m0[e, mid, -1]
        else:      ml, mr, kl, kr = mloc[e, 0, mid], mloc[e, -1, mid], m0[e, 0, mid], m0[e, -1, mid]
        wi = 1.0 / w.copy(); wi[0] = kl / (w[0] * ml); wi[-1] = kr / (w[-1] * mr)
        A = BD @ np.diag(wi) @ BD.T; M = BJ @ np.diag(wi) @ BJ.T
        lam, S = sla.eigh(A, M)
        fdm_S[e, d] = S; fdm_lam[e, d] = lam
rx = hh[:, 1] / hh[:, 0]
den = rx[:, None, None] * fdm_lam[:, 0][:, None, :] + (1 / rx)[:, None, None] * fdm_lam[:, 1][:, :, None]   # [e, j(y), i(x)]
print("min den", den.min(), (den <= 1e-12).sum())
deninv = np.where(den > 1e-10 * den.max(), 1.0 / np.maximum(den, 1e-300), 0.0)
def fdm(r):
    r = r.reshape(nel, L2, L2)
    t = np.einsum('eIi,eJj,eIJ->eji', fdm_S[:, 0], fdm_S[:, 1], r.transpose(0, 2, 1))  # S^T r : r[e,J(y),I(x)] -> t[e,j,i]
    t = t * deninv
    z = np.einsum('eIi,eJj,eji->eJI', fdm_S[:, 0], fdm_S[:, 1], t)
    return z.ravel()
# quality vs exact block
rr = rng.standard_normal(n2)
print("fdm vs exact block rel diff", np.linalg.norm(fdm(rr) - blk(rr)) / np.linalg.norm(blk(rr)))
print("FDM block-jacobi", pcg(fdm))
print("FDM + Q1 coarse", pcg(lambda r: fdm(r) + crs(r)))
print("FDM + P0 + Q1 coarse", pcg(lambda r: fdm(r) + crs0(r) + crs(r)))
# ---- 3-level additive: FDM + P1 (D1^-1 + P2 A2^-1 P2^T) P1^T with P2 = aggregates of vertices
for nagg in (64, 256):
    # aggregates of elements by partition key chunk; vertex -> aggregate of the first element holding it
    ekey = (c.key.astype(np.int64) * nagg) // c.d2
    vagg = np.zeros(nv, dtype=np.int64); vagg[vid.ravel()] = np.repeat(ekey, 4)
    P2 = sp.coo_matrix((np.ones(nv), (np.arange(nv), vagg)), shape=(nv, nagg)).tocsr()
    A2 = (P2.T @ Ac @ P2).toarray(); A2i = np.linalg.pinv(A2)
    d1 = 1.0 / Ac.diagonal()
    def crs3(r):
        rc = P.T @ r
        return P @ (d1 * rc + P2 @ (A2i @ (P2.T @ rc)))
    print("nagg", nagg, "FDM + 3-level (Jacobi Q1 + P0 aggregates)", pcg(lambda r: fdm(r) + crs3(r)))
    print("nagg", nagg, "exact block + 3-level", pcg(lambda r: blk(r) + crs3(r)))
