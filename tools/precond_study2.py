"""Scratch study 2 (CPU, oracle): cheap multilevel additive variants for the pressure preconditioner."""
import sys, os
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekstab_b200 import cases
from oracle.ops import SEM

lx1 = int(sys.argv[1]) if len(sys.argv) > 1 else 6
g = np.load("tests/golden/cyl.npz")
c = cases.cylinder_case(g, lx1=lx1, sponge=False)
s = SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)
E = s.e_sparse().tocsr()
n2 = E.shape[0]; nel = s.nel; L2 = s.lx2; np2 = L2 ** 2
rng = np.random.default_rng(0)
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

# FDM
X = c.xyz.reshape(2, nel, lx1, lx1); w = s.w; w2 = s.wg
BD = (w2[:, None] * s.D12); BJ = (w2[:, None] * s.J12)
mloc = 1.0 / (s.binv * s.bm1); m0 = s.mask[0].reshape(nel, lx1, lx1)
fdm_S = np.zeros((nel, 2, L2, L2)); fdm_lam = np.zeros((nel, 2, L2)); hh = np.zeros((nel, 2)); mid = lx1 // 2
for e in range(nel):
    hx = np.linalg.norm(X[:, e, mid, -1] - X[:, e, mid, 0]); hy = np.linalg.norm(X[:, e, -1, mid] - X[:, e, 0, mid]); hh[e] = hx, hy
    for d in range(2):
        if d == 0: ml, mr, kl, kr = mloc[e, mid, 0], mloc[e, mid, -1], m0[e, mid, 0], m0[e, mid, -1]
        else:      ml, mr, kl, kr = mloc[e, 0, mid], mloc[e, -1, mid], m0[e, 0, mid], m0[e, -1, mid]
        wi = 1.0 / w.copy(); wi[0] = kl / (w[0] * ml); wi[-1] = kr / (w[-1] * mr)
        lam, S = sla.eigh(BD @ np.diag(wi) @ BD.T, BJ @ np.diag(wi) @ BJ.T)
        fdm_S[e, d] = S; fdm_lam[e, d] = lam
rx = hh[:, 1] / hh[:, 0]
den = rx[:, None, None] * fdm_lam[:, 0][:, None, :] + (1 / rx)[:, None, None] * fdm_lam[:, 1][:, :, None]
deninv = 1.0 / den
def fdm(r):
    r = r.reshape(nel, L2, L2)
    t = np.einsum('eIi,eJj,eJI->eji', fdm_S[:, 0], fdm_S[:, 1], r) * deninv
    return np.einsum('eIi,eJj,eji->eJI', fdm_S[:, 0], fdm_S[:, 1], t).ravel()

vert = c.glo.reshape(nel, lx1, lx1)[:, [0, 0, -1, -1], [0, -1, 0, -1]]
uv, vid = np.unique(vert, return_inverse=True); vid = vid.reshape(nel, 4); nv = uv.size
zg = s.zg; l0 = (1 - zg) / 2; l1 = (1 + zg) / 2
phi = np.stack([np.outer(a, b_).ravel() for a in (l0, l1) for b_ in (l0, l1)])
rows = (np.arange(nel)[:, None, None] * np2 + np.arange(np2)[None, None, :]).repeat(4, 1).ravel()
def makeP(vid_, nv_):
    cols = vid_[:, :, None].repeat(np2, 2).ravel()
    return sp.coo_matrix((np.tile(phi, (nel, 1, 1)).ravel(), (rows, cols)), shape=(n2, nv_)).tocsr()
P = makeP(vid, nv)
Ac = (P.T @ E @ P).tocsc(); lu = spla.splu(Ac)
d1 = 1.0 / Ac.diagonal()
print("FDM + Q1 exact", pcg(lambda r: fdm(r) + P @ lu.solve(P.T @ r)))
# approximate diag from element-block parts only: sum over (e,k) of phi_k^T E_ee phi_k
Pd = makeP(np.arange(nel * 4).reshape(nel, 4), nel * 4)      # discontinuous vertex functions
Ebd = sp.block_diag([E[e*np2:(e+1)*np2, e*np2:(e+1)*np2] for e in range(nel)]).tocsr()
dd = np.asarray((Pd.multiply(Ebd @ Pd)).sum(axis=0)).ravel()
d1a = np.zeros(nv); np.add.at(d1a, vid.ravel(), dd); d1a = 1.0 / d1a
print("ratio approx/true diag: min %.3f max %.3f" % ((Ac.diagonal() * d1a).min(), (Ac.diagonal() * d1a).max()))
P0 = sp.coo_matrix((np.ones(n2), (np.arange(n2), np.arange(n2) // np2)), shape=(n2, nel)).tocsr()
A0 = (P0.T @ E @ P0).tocsc(); d0 = 1.0 / A0.diagonal()
def agg_ops(nagg):
    ekey = (c.key.astype(np.int64) * nagg) // c.d2
    Pa = sp.coo_matrix((np.ones(n2), (np.arange(n2), np.repeat(ekey, np2))), shape=(n2, nagg)).tocsr()
    A2 = (Pa.T @ E @ Pa).toarray(); return Pa, np.linalg.pinv(A2)
for nagg in (64, 256, 1024):
    Pa, A2i = agg_ops(nagg)
    ca = lambda r: Pa @ (A2i @ (Pa.T @ r))
    print("nagg", nagg)
    print("  FDM + Q1jac(true) + P0agg      ", pcg(lambda r: fdm(r) + P @ (d1 * (P.T @ r)) + ca(r)))
    print("  FDM + Q1jac(approx) + P0agg    ", pcg(lambda r: fdm(r) + P @ (d1a * (P.T @ r)) + ca(r)))
    print("  FDM + Q1jac(2x approx) + P0agg ", pcg(lambda r: fdm(r) + P @ (2 * d1a * (P.T @ r)) + ca(r)))
    print("  FDM + P0elem-jac + P0agg       ", pcg(lambda r: fdm(r) + P0 @ (d0 * (P0.T @ r)) + ca(r)))
    print("  FDM + P0agg                    ", pcg(lambda r: fdm(r) + ca(r)))
    print("  FDM + Q1jac + P0elem-jac + P0agg", pcg(lambda r: fdm(r) + P @ (d1 * (P.T @ r)) + P0 @ (d0 * (P0.T @ r)) + ca(r)))
# split interface vertices: 8 ranks
from nekstab_b200.cases import partition
rk = partition(c.key, 8, c.d2)
vkey = vid * 8 + rk[:, None]
uv2, vid2 = np.unique(vkey, return_inverse=True); vid2 = vid2.reshape(nel, 4)
Ps = makeP(vid2, uv2.size); ds = 1.0 / (Ps.T @ E @ Ps).diagonal()
Pa, A2i = agg_ops(256)
print("split vertices (8 ranks): FDM + Q1jac + P0agg(256)", pcg(lambda r: fdm(r) + Ps @ (ds * (Ps.T @ r)) + Pa @ (A2i @ (Pa.T @ r))))

# ---- Q1-level Jacobi with the diagonal of the H1 (SEM Laplacian) energy of the hat functions: additive over elements
z1 = s.z; a0 = (1 - z1) / 2; a1 = (1 + z1) / 2
phi1 = np.stack([np.outer(a, b_) for a in (a0, a1) for b_ in (a0, a1)])          # (4, lx1, lx1) on GLL nodes
dl = np.zeros((nel, 4))
for k in range(4):
    f = np.broadcast_to(phi1[k], s.eshape)
    dl[:, k] = (f * s.axhelm(f, 1.0, 0.0)).reshape(nel, -1).sum(1)
d1l = np.zeros(nv); np.add.at(d1l, vid.ravel(), dl.ravel()); 
rat = Ac.diagonal() / d1l
print("true/laplace diag ratio: min %.3f max %.3f median %.3f" % (rat.min(), rat.max(), np.median(rat)))
d1l = 1.0 / d1l
for nagg in (256,):
    Pa, A2i = agg_ops(nagg)
    ca = lambda r: Pa @ (A2i @ (Pa.T @ r))
    for sc in (0.5, 1.0, 1.5, 2.0):
        print("  FDM + Q1jac(laplace diag x%.1f) + P0agg" % sc, pcg(lambda r: fdm(r) + P @ (sc * d1l * (P.T @ r)) + ca(r)))
    for sc in (0.7, 1.5, 2.0):
        print("  FDM + Q1jac(true diag x%.1f) + P0agg" % sc, pcg(lambda r: fdm(r) + P @ (sc * d1 * (P.T @ r)) + ca(r)))
