"""Case builder: prepares, on the host, the arrays a Nek5000 executable holds in its COMMON blocks and
would hand to ``nsb_init`` (include/nekstab_b200.h): GLL coordinates, Dirichlet masks, global node
numbers, the element partition, base flow, sponge function.  It stands in for the Nek5000 *setup* phase
[UPSTREAM: connect2.f/map2.f get_vert_map + assign_gllnid, navier8.f setvert2d, bdry.f bcmask,
coef.f] which is outside the hot path; the hot path itself never runs here.

Conventions: fields are ``a[e, k, j, i]`` with i (r) fastest; `.ma2` vertex
order is lexicographic in (r,s); `.re2` faces are 1:s=-1, 2:r=+1, 3:s=+1, 4:r=-1, 5:t=-1, 6:t=+1.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import sem


# ----------------------------------------------------------------------------- numbering
def global_numbering_2d(vert: np.ndarray, lx1: int) -> np.ndarray:
    """Global GLL node ids (nel, lx1*lx1) from `.ma2` vertex ids.

    Vertex nodes take the vertex id; edge-interior nodes are identified by the sorted vertex pair and
    the position counted from the smaller vertex id; element-interior nodes are private.  Yields the
    same equivalence classes as gslib's setup on Nek5000's `setvert2d` numbers [UPSTREAM].
    """
    vert = np.asarray(vert, dtype=np.int64)
    nel = vert.shape[0]
    N = lx1 - 1
    nv = int(vert.max()) + 1
    glo = np.full((nel, lx1, lx1), -1, dtype=np.int64)        # [e, j, i]
    glo[:, 0, 0], glo[:, 0, N], glo[:, N, 0], glo[:, N, N] = vert[:, 0], vert[:, 1], vert[:, 2], vert[:, 3]
    # edges: (va, vb, slice of nodes ordered from va to vb)
    edges = [(0, 1, (0, slice(1, N))), (2, 3, (N, slice(1, N))),          # s=-1, s=+1 (run along i)
             (0, 2, (slice(1, N), 0)), (1, 3, (slice(1, N), N))]          # r=-1, r=+1 (run along j)
    pair = []
    for a, b, _ in edges:
        lo = np.minimum(vert[:, a], vert[:, b]); hi = np.maximum(vert[:, a], vert[:, b])
        pair.append(lo * nv + hi)
    pair = np.stack(pair, axis=1)                                         # (nel, 4)
    uniq, eid = np.unique(pair.ravel(), return_inverse=True)
    eid = eid.reshape(nel, 4)
    base_e = nv
    pos = np.arange(N - 1, dtype=np.int64)
    for k, (a, b, sl) in enumerate(edges):
        fwd = vert[:, a] < vert[:, b]
        p = np.where(fwd[:, None], pos[None, :], pos[None, ::-1])
        ids = base_e + eid[:, k:k + 1] * (N - 1) + p
        if isinstance(sl[0], int):
            glo[:, sl[0], sl[1]] = ids
        else:
            glo[:, sl[0], sl[1]] = ids
    base_i = base_e + len(uniq) * (N - 1)
    if N > 1:
        ii = np.arange((N - 1) * (N - 1), dtype=np.int64).reshape(N - 1, N - 1)
        glo[:, 1:N, 1:N] = base_i + np.arange(nel, dtype=np.int64)[:, None, None] * (N - 1) ** 2 + ii
    assert glo.min() >= 0
    return _compress(glo.reshape(nel, -1))


def _compress(glo: np.ndarray) -> np.ndarray:
    """Renumber ids densely 0..nuniq-1 preserving order."""
    _, inv = np.unique(glo.ravel(), return_inverse=True)
    return inv.reshape(glo.shape).astype(np.int64)


def extrude_numbering(glo2d: np.ndarray, lx1: int, nz: int, periodic: bool = True, compress: bool = True) -> np.ndarray:
    """Global ids for the z-extrusion of a 2-D numbering: node = (2-D node, global z level).  With compress=False the
    ids are `glo2d*nlev + level`, i.e. consistent across ranks that each extrude their own share of the 2-D mesh."""
    nel2, npt2 = glo2d.shape
    N = lx1 - 1
    nlev = nz * N if periodic else nz * N + 1
    k = np.arange(lx1)
    out = np.empty((nz, nel2, lx1, npt2), dtype=np.int64)
    for L in range(nz):
        lev = (L * N + k) % nlev if periodic else (L * N + k)
        out[L] = glo2d[:, None, :] * nlev + lev[None, :, None]
    out = out.reshape(nz * nel2, lx1 * npt2)
    return _compress(out) if compress else out


# ----------------------------------------------------------------------------- partition
def partition(key: np.ndarray, nranks: int, d2: Optional[int] = None) -> np.ndarray:
    """Element -> rank map, Nek5000's rule recovered from the shipped field files
    [UPSTREAM map2.f assign_gllnid].  Power-of-two rank counts: ``rank = key // (npstar/P)`` with
    npstar the power-of-two key range; otherwise sort by key (Nek's unstable heap sort,
    `_nek_isort`) and deal out floor(nel/P) elements to the first ``P - nel % P`` ranks and one more to the rest.  Local order = ascending global id."""
    key = np.asarray(key, dtype=np.int64)
    nel = key.size
    if nranks == 1:
        return np.zeros(nel, dtype=np.int32)
    if nranks & (nranks - 1) == 0:
        if d2 is None:
            d2 = 1
            while d2 < key.max() + 1:
                d2 *= 2
        if d2 % nranks == 0 and d2 >= nranks:
            return (key // (d2 // nranks)).astype(np.int32)
    order = _nek_isort(key)
    rank = np.empty(nel, dtype=np.int32)
    small, nbig = nel // nranks, nel % nranks
    pos = 0
    for r in range(nranks):
        cnt = small if r < nranks - nbig else small + 1
        rank[order[pos:pos + cnt]] = r
        pos += cnt
    return rank


def _nek_isort(key: np.ndarray) -> np.ndarray:
    """Permutation produced by Nek5000's `isort` [UPSTREAM math.f: heap sort, Numerical Recipes 1st ed.
    p.231].  It is *not* stable; the order of equal keys decides which elements straddle a rank border
    for non-power-of-two rank counts.  Reproduces the 6-rank element maps of the shipped cylinder and
    BFS field files exactly (tests/test_partition.py)."""
    a = [0] + [int(k) for k in key]
    n = len(key)
    ind = [0] + list(range(n))
    if n <= 1:
        return np.arange(n)
    l, ir = n // 2 + 1, n
    while True:
        if l > 1:
            l -= 1
            aa, ii = a[l], ind[l]
        else:
            aa, ii = a[ir], ind[ir]
            a[ir], ind[ir] = a[1], ind[1]
            ir -= 1
            if ir == 1:
                a[1], ind[1] = aa, ii
                break
        i, j = l, l + l
        while j <= ir:
            if j < ir and a[j] < a[j + 1]:
                j += 1
            if aa < a[j]:
                a[i], ind[i] = a[j], ind[j]
                i, j = j, j + j
            else:
                j = ir + 1
        a[i], ind[i] = aa, ii
    return np.array(ind[1:], dtype=np.int64)


# ----------------------------------------------------------------------------- masks
_FACE_SLICES_2D = {1: (0, slice(None)), 2: (slice(None), -1), 3: (-1, slice(None)), 4: (slice(None), 0)}


def dirichlet_mask(is_dirichlet_face: np.ndarray, lx1: int, ldim: int) -> np.ndarray:
    """(nel, lx1^ldim) array, 0 on every node of a Dirichlet face, 1 elsewhere.  `is_dirichlet_face`
    is (nel, 2*ldim) bool in `.re2` face order.  The zero is then spread to every copy of the node by the
    caller through a min-gather (Nek's bcmask + dsop 'MUL' [UPSTREAM bdry.f])."""
    nel = is_dirichlet_face.shape[0]
    shp = (nel,) + (lx1,) * ldim
    m = np.ones(shp)
    for f in range(2 * ldim):
        sel = is_dirichlet_face[:, f]
        if not sel.any():
            continue
        if ldim == 2:
            sl = _FACE_SLICES_2D[f + 1]
            sub = m[sel]; sub[(slice(None),) + sl] = 0.0; m[sel] = sub
        else:
            sub = m[sel]
            if f < 4:
                sl = _FACE_SLICES_2D[f + 1]
                sub[(slice(None), slice(None)) + sl] = 0.0
            elif f == 4:
                sub[:, 0] = 0.0
            else:
                sub[:, -1] = 0.0
            m[sel] = sub
    return m.reshape(nel, -1)


def spread_min(a: np.ndarray, glo: np.ndarray) -> np.ndarray:
    """Every copy of a global node receives the minimum over its copies."""
    g = glo.ravel()
    mn = np.full(int(g.max()) + 1, np.inf)
    np.minimum.at(mn, g, a.ravel())
    return mn[g].reshape(a.shape)


# ----------------------------------------------------------------------------- sponge
def mth_stepf(x):
    """Smooth step of core/utils.f:330-342."""
    x = np.asarray(x, dtype=np.float64)
    out = np.ones_like(x)
    out[x <= 0.0010] = 0.0
    mid = (x > 0.0010) & (x <= 0.9990)
    xm = x[mid]
    out[mid] = 1.0 / (1.0 + np.exp(1.0 / (xm - 1.0) + 1.0 / xm))
    return out


def sponge_function(coords, lspg, rspg, acc_spg=0.333):
    """spng_fun of core/utils.f:205-328 (spng_init + spng_set).  `coords` is a list of per-direction
    coordinate arrays, `lspg`/`rspg` the per-direction left/right sponge lengths (xLspg.., xRspg..).
    Note the ramp argument is divided by the *section width* spng_wl/wr, as the reference does
    (:306,:311)."""
    fun = np.zeros_like(coords[0])
    for x, L, R in zip(coords, lspg, rspg):
        wl, wr = (1.0 - acc_spg) * L, (1.0 - acc_spg) * R
        dl, dr = acc_spg * L, acc_spg * R
        if not (wl > 0.0 or wr > 0.0):
            continue
        bmin, bmax = x.min(), x.max()
        xxmax, xxmin = bmax - wr, bmin + wl
        xxmax_c, xxmin_c = xxmax + dr, xxmin - dl
        r = np.zeros_like(x)
        r[x <= xxmin_c] = 1.0
        sel = (x > xxmin_c) & (x < xxmin)
        if wl > 0:
            r[sel] = mth_stepf((xxmin - x[sel]) / wl)
        sel = (x > xxmax) & (x < xxmax_c)
        if wr > 0:
            r[sel] = mth_stepf((x[sel] - xxmax) / wr)
        r[x >= xxmax_c] = 1.0
        r[(x >= xxmin) & (x <= xxmax)] = 0.0
        fun = np.maximum(fun, r)
    return fun


# ----------------------------------------------------------------------------- noise seed
def mth_rand(ix, iy, iz, ieg, xl, fcoeff, if3d):
    """core/utils.f:457-469, vectorised.  ix,iy,iz,ieg 1-based."""
    r = fcoeff[0] * (ieg + xl[0] * np.sin(xl[1])) + fcoeff[1] * ix * iy + fcoeff[2] * ix
    if if3d:
        r = fcoeff[0] * (ieg + xl[2] * np.sin(r)) + fcoeff[1] * iz * ix + fcoeff[2] * iz
    r = 1.0e3 * np.sin(r)
    r = 1.0e3 * np.sin(r)
    return np.cos(r)


def raw_noise(case: "Case") -> np.ndarray:
    """mth_rand per GLL point and component (core/utils.f:344-383), before the direct-stiffness average."""
    d, lx1, nel = case.ldim, case.lx1, case.nel
    ieg = (np.arange(1, nel + 1) if case.lglel is None else np.asarray(case.lglel)).astype(np.float64)[:, None]
    p = np.arange(lx1 ** d)
    ix = (p % lx1 + 1).astype(np.float64)[None, :]
    iy = ((p // lx1) % lx1 + 1).astype(np.float64)[None, :]
    iz = (p // (lx1 * lx1) + 1).astype(np.float64)[None, :] if d == 3 else np.ones((1, lx1 ** d))
    xl = [case.xyz[k] for k in range(d)]
    coeffs = [(3.0e4, -1.5e3, 0.5e5), (2.3e4, 2.3e3, -2.0e5), (2.0e4, 1.0e3, 1.0e5)]
    return np.stack([mth_rand(ix, iy, iz, ieg, xl, coeffs[k], d == 3) for k in range(d)])


def add_noise(case: "Case", lglel=None) -> np.ndarray:
    """Deterministic noise seed of core/utils.f:344-408 (add_noise on zero fields): mth_rand per GLL point and
    component, then direct-stiffness average (opdssum * vmult, dsavg) and Dirichlet masking (bcdirvc)."""
    d, lx1, nel = case.ldim, case.lx1, case.nel
    ieg = (np.arange(1, nel + 1) if lglel is None else np.asarray(lglel)).astype(np.float64)[:, None]
    p = np.arange(lx1 ** d)
    ix = (p % lx1 + 1).astype(np.float64)[None, :]
    iy = ((p // lx1) % lx1 + 1).astype(np.float64)[None, :]
    iz = (p // (lx1 * lx1) + 1).astype(np.float64)[None, :] if d == 3 else np.ones((1, lx1 ** d))
    xl = [case.xyz[k] for k in range(d)]
    coeffs = [(3.0e4, -1.5e3, 0.5e5), (2.3e4, 2.3e3, -2.0e5), (2.0e4, 1.0e3, 1.0e5)]
    g = case.glo.ravel()
    ng = int(g.max()) + 1
    cnt = np.bincount(g, minlength=ng)
    out = []
    for k in range(d):
        q = mth_rand(ix, iy, iz, ieg, xl, coeffs[k], d == 3)
        avg = np.bincount(g, weights=q.ravel(), minlength=ng) / cnt
        out.append(avg[g].reshape(nel, -1) * case.mask[k])
    return np.stack(out)


# ----------------------------------------------------------------------------- the case
@dataclass
class Case:
    """Everything ``nsb_init`` needs, for the *global* mesh (use `local_part` for one rank's share)."""
    name: str
    ldim: int
    lx1: int
    nel: int
    xyz: np.ndarray              # (ldim, nel, lx1^ldim)
    glo: np.ndarray              # (nel, lx1^ldim) int64
    mask: np.ndarray             # (ldim, nel, lx1^ldim) 1 = free, 0 = Dirichlet
    key: np.ndarray              # (nel,) partition key
    d2: int
    ubase: np.ndarray            # (ldim, nel, lx1^ldim)
    re: float
    end_time: float
    cfl_target: float = 0.5
    tol_p: float = 1e-7
    tol_v: float = 1e-9
    spng_fun: Optional[np.ndarray] = None      # (nel, lx1^ldim) or None
    lglel: Optional[np.ndarray] = None         # global element ids (1-based) of the local elements
    nelg: Optional[int] = None                 # global element count (None: this is the global mesh)
    # Nek's ifvcor: no outflow-type boundary => pressure defined up to a constant (E singular) => `ortho`.
    # None lets the library decide numerically from ||E 1||.
    ifvcor: Optional[bool] = None
    ifvcor_adjoint: Optional[bool] = None
    extra: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def lx2(self):
        return self.lx1 - 2

    @property
    def lxd(self):
        return 3 * self.lx1 // 2

    @property
    def npts(self):
        return self.lx1 ** self.ldim

    @property
    def n(self):
        return self.nel * self.npts

    def with_adjoint_bcs(self) -> "Case":
        """Outflow 'O' -> Dirichlet for the adjoint problem (examples/cylinder/stability/direct/1cyl.usr:126-132)."""
        if "mask_adjoint" not in self.extra:
            return self
        import copy
        c = copy.copy(self)
        c.mask = self.extra["mask_adjoint"]
        return c

    def local_part(self, rank: int, nranks: int) -> "Case":
        """The elements of `rank` under Nek5000's partition rule, ascending global element id."""
        import copy
        r = partition(self.key, nranks, self.d2)
        sel = np.nonzero(r == rank)[0]
        c = copy.copy(self)
        c.nel = sel.size
        c.nelg = self.nelg or self.nel
        c.xyz = self.xyz[:, sel]
        c.glo = self.glo[sel]
        c.mask = self.mask[:, sel]
        c.key = self.key[sel]
        c.ubase = self.ubase[:, sel]
        c.spng_fun = None if self.spng_fun is None else self.spng_fun[sel]
        c.lglel = (sel + 1).astype(np.int64) if self.lglel is None else self.lglel[sel]
        c.extra = {k: (v[:, sel] if v.ndim == 3 else v[sel]) for k, v in self.extra.items()
                   if isinstance(v, np.ndarray) and v.ndim >= 2 and v.shape[-2] == self.nel}
        return c


def _mask_from_faces(dirich_faces, glo, lx1, ldim):
    m = dirichlet_mask(dirich_faces, lx1, ldim)
    m = spread_min(m, glo)
    return np.repeat(m[None], ldim, axis=0)


def _interp_elements(a: np.ndarray, lx_from: int, lx_to: int, ldim: int) -> np.ndarray:
    """Tensor-product interpolation of (..., nel, lx_from^ldim) element data to lx_to GLL points."""
    if lx_from == lx_to:
        return a
    J = sem.lagrange_interp_matrix(sem.zwgll(lx_to)[0], sem.zwgll(lx_from)[0])
    lead = a.shape[:-1]
    b = a.reshape(lead + (lx_from,) * ldim)
    for ax in range(ldim):
        b = np.moveaxis(np.tensordot(b, J, axes=([b.ndim - 1 - ax], [1])), -1, b.ndim - 1 - ax)
    return b.reshape(lead + (lx_to ** ldim,))


def cylinder_case(g: dict, lx1: Optional[int] = None, sponge: bool = True) -> Case:
    """Config 1: examples/cylinder/stability/direct (Re=50, endTime 1, sponge 5/5, tol 1e-7/1e-9;
    1cyl.par:1-36).  `g` = tests/golden/cyl.npz."""
    lx_file = int(g["lx1"])
    lx1 = lx1 or lx_file
    X = _interp_elements(g["X"].reshape(-1, 2, lx_file ** 2).transpose(1, 0, 2), lx_file, lx1, 2)
    U = _interp_elements(g["U"].reshape(-1, 2, lx_file ** 2).transpose(1, 0, 2), lx_file, lx1, 2)
    glo = global_numbering_2d(g["vert"], lx1)
    bc = g["bc"]                                      # (nel,4) uint8 codes: see tools/make_golden.py
    codes = {c: i for i, c in enumerate([s.decode() if isinstance(s, bytes) else str(s) for s in g["bc_names"]])}
    dir_d = (bc == codes["v  "]) | (bc == codes["W  "])
    dir_a = dir_d | (bc == codes["O  "])
    case = Case("cylinder_re50", 2, lx1, X.shape[1], X, glo, _mask_from_faces(dir_d, glo, lx1, 2),
                g["key"].astype(np.int64), int(g["d2"]), U, re=50.0, end_time=1.0, tol_p=1e-7, tol_v=1e-9)
    case.extra["mask_adjoint"] = _mask_from_faces(dir_a, glo, lx1, 2)
    case.ifvcor, case.ifvcor_adjoint = False, True
    if sponge:
        case.spng_fun = sponge_function([X[0], X[1]], [5.0, 0.0], [5.0, 0.0])
    return case


def bfs_case(g: dict, sponge: bool = True) -> Case:
    """Config 4: examples/back_fstep/transient_growth (Re=500, endTime 1, sponge 5/10; bfs.par).  Geometry is
    rebuilt in double precision from the straight-sided `.re2` vertices; all tagged faces are Dirichlet
    (bfs.usr:101-103 setbc 4->v, 2->v, 3->W)."""
    lx1 = int(g["lx1"])
    z, _ = sem.zwgll(lx1)
    V = g["re2_xyz"]                                   # (nel, 2, 4) preprocessor vertex order (ccw)
    r = z[None, None, :]; s = z[None, :, None]
    h = [(1 - r) * (1 - s) / 4, (1 + r) * (1 - s) / 4, (1 + r) * (1 + s) / 4, (1 - r) * (1 + s) / 4]
    X = np.stack([sum(V[:, c, k][:, None, None] * h[k] for k in range(4)) for c in range(2)])
    X = X.reshape(2, V.shape[0], lx1 * lx1)
    glo = global_numbering_2d(g["vert"], lx1)
    dirf = g["bc_id"] > 0
    U = g["U"].reshape(-1, 2, lx1 * lx1).transpose(1, 0, 2).astype(np.float64)
    case = Case("bfs_re500", 2, lx1, X.shape[1], X, glo, _mask_from_faces(dirf, glo, lx1, 2),
                g["key"].astype(np.int64), int(g["d2"]), U, re=500.0, end_time=1.0, tol_p=1e-8, tol_v=1e-8)
    case.ifvcor = case.ifvcor_adjoint = True
    if sponge:
        case.spng_fun = sponge_function([X[0], X[1]], [5.0, 0.0], [10.0, 0.0])
    return case


def cavity_case(g: dict) -> Case:
    """Config 3: examples/lid_driven (Re = 3600 `viscosity = -3600` cav.par:28, endTime 0.5 cav.par:4, tol 1e-9/1e-9, k_dim 90 cav.par:8,
    schur_tgt 4 cav.usr:22; lid 'v' at the top face, walls 'W' elsewhere => all-Dirichlet velocity, singular E).  `g` =
    tests/golden/cav.npz.  The coordinates are those of the shipped base flow (y in [0, 1.2]; cav.par:9 / cav.usr:107-109 would
    rescale to 1.5 -- SURVEY.md cfg-3 caveat: the fixture is kept as shipped so that U is a solution ON ITS OWN mesh).  No sponge."""
    lx1 = int(g["lx1"])
    X = g["X"].reshape(-1, 2, lx1 * lx1).transpose(1, 0, 2).astype(np.float64)
    U = g["U"].reshape(-1, 2, lx1 * lx1).transpose(1, 0, 2).astype(np.float64)
    glo = global_numbering_2d(g["vert"], lx1)
    codes = {c: i for i, c in enumerate([s.decode() if isinstance(s, bytes) else str(s) for s in g["bc_names"]])}
    dirf = (g["bc"] == codes["v  "]) | (g["bc"] == codes["W  "])
    case = Case("cavity_re3600", 2, lx1, X.shape[1], X, glo, _mask_from_faces(dirf, glo, lx1, 2), g["key"].astype(np.int64),
                int(g["d2"]), U, re=3600.0, end_time=0.5, tol_p=1e-9, tol_v=1e-9)
    case.ifvcor = case.ifvcor_adjoint = True
    return case


def thermosyphon_case(g: dict, ra: float = 400.0) -> Case:
    """Scalar-transport fixture: examples/thersyphon/baseflow (2-D annulus 1 <= r <= 2, periodic in theta, 256 elements, lx1 = 6).
    tsyphon.par: viscosity = 5 (the Prandtl number, positive => taken as is), conductivity = 1, rhocp = 1, endTime 0.1, tolerances 1e-11;
    tsyphon.usr userf: ffy = T * Pr * Ra.  Walls 'W' (velocity) / 't' (temperature, 0.5 (1 + tanh(-20 y)), carried by the field itself)
    at both radii => all-Dirichlet velocity, singular E.  `g` = tests/golden/tsyphon.npz; the shipped field is the Newton solution at
    Ra = 400.  case.extra: "T" the base temperature, "tmask" its Dirichlet mask, "ri" = Pr * Ra, "cond", "rhocp", "P"."""
    lx1 = int(g["lx1"])
    X = g["X"].reshape(-1, 2, lx1 * lx1).transpose(1, 0, 2).astype(np.float64)
    U = g["U"].reshape(-1, 2, lx1 * lx1).transpose(1, 0, 2).astype(np.float64)
    glo = global_numbering_2d(g["vert"], lx1)
    codes = {c: i for i, c in enumerate([s.decode() if isinstance(s, bytes) else str(s) for s in g["bc_names"]])}
    pr = 5.0
    case = Case("thermosyphon_ra%g" % ra, 2, lx1, X.shape[1], X, glo, _mask_from_faces(g["bc"] == codes["W  "], glo, lx1, 2),
                g["key"].astype(np.int64), int(g["d2"]), U, re=1.0 / pr, end_time=0.1, tol_p=1e-11, tol_v=1e-11)
    case.ifvcor = case.ifvcor_adjoint = True
    case.extra["T"] = g["T"].reshape(-1, lx1 * lx1).astype(np.float64)
    case.extra["P"] = g["P"].reshape(-1, lx1 * lx1).astype(np.float64)
    case.extra["tmask"] = _mask_from_faces(g["bct"] == codes["t  "], glo, lx1, 2)[0]
    case.extra["ri"], case.extra["cond"], case.extra["rhocp"] = np.float64(pr * ra), np.float64(1.0), np.float64(1.0)
    return case


def extrude(c2: Case, nz: int, lz: float, name: Optional[str] = None, compress_ids: bool = True) -> Case:
    """Config 5 recipe: extrude a 2-D case into nz uniform periodic layers over [0,lz];
    element eg3 = layer*nel2 + eg2, key3 = key2, z-invariant base flow with W=0."""
    lx1 = c2.lx1
    z, _ = sem.zwgll(lx1)
    nel2, np2 = c2.nel, c2.npts
    dz = lz / nz

    def lift(a):                                       # (nel2, np2) -> (nz*nel2, lx1*np2)
        return np.broadcast_to(a[None, :, None, :], (nz, nel2, lx1, np2)).reshape(nz * nel2, lx1 * np2)

    zc = (np.arange(nz)[:, None] + (z[None, :] + 1) / 2) * dz            # (nz, lx1)
    Z = np.broadcast_to(zc[:, None, :, None], (nz, nel2, lx1, np2)).reshape(nz * nel2, lx1 * np2)
    xyz = np.stack([lift(c2.xyz[0]), lift(c2.xyz[1]), Z])
    glo = extrude_numbering(c2.glo, lx1, nz, periodic=True, compress=compress_ids)
    m2 = lift(c2.mask[0])
    mask = np.stack([m2, m2, m2])
    ub = np.stack([lift(c2.ubase[0]), lift(c2.ubase[1]), np.zeros_like(Z)])
    # partition key: every z-layer of a 2-D element inherits the element's RSB key, so the power-of-two rule
    # rank = key // (d2/P) splits the extruded mesh into the same balanced 2-D sub-domains as the shipped mesh
    # (genmap is not available to produce a true 3-D RSB key: "parity unpinned" for this synthetic mesh)
    key = np.broadcast_to(c2.key[None, :], (nz, nel2)).reshape(-1).copy()
    d2 = c2.d2
    case = Case(name or (c2.name + f"_x{nz}"), 3, lx1, nz * nel2, np.ascontiguousarray(xyz), glo,
                np.ascontiguousarray(mask), key, d2, np.ascontiguousarray(ub), re=c2.re,
                end_time=c2.end_time, tol_p=c2.tol_p, tol_v=c2.tol_v)
    case.ifvcor, case.ifvcor_adjoint = c2.ifvcor, c2.ifvcor_adjoint
    if c2.nelg is not None:
        case.nelg = c2.nelg * nz
    if c2.lglel is not None:                      # eg3 = layer*nel2_global + eg2
        case.lglel = (np.arange(nz)[:, None] * c2.nelg + c2.lglel[None, :]).reshape(-1)
    if c2.spng_fun is not None:
        case.spng_fun = np.ascontiguousarray(lift(c2.spng_fun))
    if "mask_adjoint" in c2.extra:
        ma = lift(c2.extra["mask_adjoint"][0])
        case.extra["mask_adjoint"] = np.stack([ma, ma, ma])
    return case


def box_case(nex: int, ney: int, lx1: int, *, lxy=(2.0, 1.0), periodic_y=False, outflow=True, deform=0.05,
             shear=0.0, re=40.0, end_time=0.1, seed=0) -> Case:
    """Small synthetic 2-D channel-like box for unit tests: inflow 'v' at x=0, 'O' (or 'v') at x=L, walls 'W'
    (or periodic) in y, smoothly deformed interior so that all metric terms are exercised; base flow =
    a smooth, not divergence-free, field (parity tests only need a deterministic input)."""
    nvx, nvy = nex + 1, (ney if periodic_y else ney + 1)
    vert = np.empty((nex * ney, 4), dtype=np.int64)
    z, _ = sem.zwgll(lx1)
    X = np.empty((2, nex * ney, lx1, lx1))
    dirf = np.zeros((nex * ney, 4), dtype=bool)
    for ey in range(ney):
        for ex in range(nex):
            e = ey * nex + ex
            v = lambda ix, iy: (iy % nvy) * nvx + ix
            vert[e] = [v(ex, ey), v(ex + 1, ey), v(ex, ey + 1), v(ex + 1, ey + 1)]
            x0 = lxy[0] * (ex + (z[None, :] + 1) / 2) / nex
            y0 = lxy[1] * (ey + (z[:, None] + 1) / 2) / ney
            x0, y0 = np.broadcast_arrays(x0, y0)
            sx = np.sin(np.pi * x0 / lxy[0]); sy = np.sin(2 * np.pi * y0 / lxy[1])
            X[0, e] = x0 + deform * sx * sy * lxy[0] / nex + shear * y0
            X[1, e] = y0 + (0.0 if periodic_y else deform * sx * np.sin(np.pi * y0 / lxy[1]) * lxy[1] / ney)
            dirf[e, 3] = ex == 0
            dirf[e, 1] = (ex == nex - 1) and not outflow
            if not periodic_y:
                dirf[e, 0] = ey == 0
                dirf[e, 2] = ey == ney - 1
    X = X.reshape(2, nex * ney, lx1 * lx1)
    glo = global_numbering_2d(vert, lx1)
    mask = _mask_from_faces(dirf, glo, lx1, 2)
    yy = X[1] / lxy[1]; xx = X[0] / lxy[0]
    U = np.stack([1.0 + 0.3 * np.sin(2 * np.pi * yy) * np.cos(np.pi * xx) if periodic_y else 4 * yy * (1 - yy) * (1 + 0.2 * np.sin(np.pi * xx)),
                  0.1 * np.sin(2 * np.pi * xx) * np.sin(2 * np.pi * yy)])
    key = np.arange(nex * ney, dtype=np.int64)
    d2 = 1
    while d2 < nex * ney:
        d2 *= 2
    case = Case(f"box{nex}x{ney}", 2, lx1, nex * ney, X, glo, mask, key, d2, U, re=re, end_time=end_time,
                tol_p=1e-10, tol_v=1e-10)
    case.ifvcor = not outflow
    return case
