"""Checkpoint / restart protocol of the Arnoldi factorisation (KRY* field files + HES* Hessenberg files) and the
`Spectre_*.dat` spectrum files, as nekStab writes and reads them -- host-side I/O either side of the hot path.

Reference: `arnoldi_checkpoint` core/eigensolvers.f:802-905 (KRY<session>0.f%05d with index k+1, `HES<session>%04d` written
list-directed row by row, `Spectre_H<evop>%04d.dat` / `Spectre_NS<evop>%04d.dat` in '(3E15.7)'), the restart branch of
`krylov_schur` core/eigensolvers.f:284-325 (reads HES, `mstart = mstart+1`, `load_files(Q, mstart, k_dim+1, 'KRY')`),
`load_files` core/IO.f:15-60.  Nek5000's `outpost` writes the P_N-P_{N-2} pressure on mesh 1 and `load_fld` maps it back
([UPSTREAM] prepost.f prepost_map / map_pm1_to_pr): both are spectral interpolations GL(lx2) <-> GLL(lx1), exact for the
degree-(lx2-1) pressure, restated in `pressure_to_mesh1` / `pressure_to_mesh2`.

In a drop-in the reference's Fortran routines keep doing this I/O (they call nsb_vec_download / nsb_vec_upload at the field
access sites); this module is the same protocol for the Python stand-in host, so that checkpoints
written by either side can be read by the other.
"""
from __future__ import annotations

import math
import os
from typing import Tuple

import numpy as np

from . import nekio, sem


def fortran_e(x: float, w: int = 15, d: int = 7) -> str:
    """Fortran `Ew.d` edit descriptor (0.dddddddE+ee), right-justified in w columns."""
    if x == 0.0 or not math.isfinite(x):
        body = "0." + "0" * d + "E+00" if x == 0.0 else str(x)
        return body.rjust(w)
    s = "%.*E" % (d - 1, abs(x))                    # d.dddddE+ee with d significant digits
    mant, ex = s.split("E")
    digits = mant.replace(".", "")
    e = int(ex) + 1
    body = "0." + digits + "E" + ("+" if e >= 0 else "-") + "%02d" % abs(e)
    if x < 0:
        body = "-" + body
    return body.rjust(w)


def write_spectrum(path: str, vals: np.ndarray, residual: np.ndarray) -> None:
    """`write(67,'(3E15.7)') real(vals(i)), aimag(vals(i)), residual(i)` (core/eigensolvers.f:873-875, 590-604)."""
    with open(path, "w") as f:
        for v, r in zip(np.asarray(vals), np.asarray(residual)):
            f.write(fortran_e(v.real) + fortran_e(v.imag) + fortran_e(float(r)) + "\n")


def log_transform(vals: np.ndarray, tau: float) -> np.ndarray:
    """Eigenvalues of the linearised Navier-Stokes operator from those of exp(tau L) (core/eigensolvers.f:908-915)."""
    vals = np.asarray(vals, dtype=complex)
    out = (np.log(np.abs(vals)) + 1j * np.arctan2(vals.imag, vals.real)) / tau
    # `if (aimag(x) .eq. 0) logx = cmplx(real(logx), 0)` (:912-913): a negative real Ritz value gets imaginary part 0, not pi/tau
    out.imag[vals.imag == 0.0] = 0.0
    return out


def kry_filename(session: str, i: int, prefix: str = "KRY") -> str:
    return "%s%s0.f%05d" % (prefix, session, i)


def hes_filename(session: str, k: int) -> str:
    return "HES%s%04d" % (session, k)


def write_hessenberg(path: str, H: np.ndarray, k: int) -> None:
    """`write(67,*) ((H(i,j), j=1,k), i=1,k+1)` (core/eigensolvers.f:885-889): row by row, full double precision."""
    with open(path, "w") as f:
        for i in range(k + 1):
            f.write(" ".join("%.17E" % H[i, j] for j in range(k)) + "\n")


def read_hessenberg(path: str, mstart: int) -> np.ndarray:
    """`read(67,*) ((H(i,j), j=1,mstart), i=1,mstart+1)` (core/eigensolvers.f:301); list-directed: any white space."""
    a = np.array(open(path).read().replace("D", "E").split(), dtype=float)
    if a.size != (mstart + 1) * mstart:
        raise ValueError(f"{path}: expected {(mstart + 1) * mstart} numbers for mstart={mstart}, found {a.size}")
    return a.reshape(mstart + 1, mstart)


def _interp_tensor(a: np.ndarray, J: np.ndarray, ldim: int) -> np.ndarray:
    for ax in range(ldim):
        a = np.moveaxis(np.tensordot(a, J, axes=([-1 - ax], [1])), -1, -1 - ax)
    return a


def pressure_to_mesh1(p2: np.ndarray, lx1: int, ldim: int) -> np.ndarray:
    """(nel, lx2^ldim) -> (nel, lx1^ldim): GL(lx2) -> GLL(lx1) interpolation [UPSTREAM prepost.f prepost_map]."""
    lx2 = lx1 - 2
    J = sem.lagrange_interp_matrix(sem.zwgll(lx1)[0], sem.zwgl(lx2)[0])
    a = np.asarray(p2, float).reshape((-1,) + (lx2,) * ldim)
    return _interp_tensor(a, J, ldim).reshape(a.shape[0], -1)


def pressure_to_mesh2(p1: np.ndarray, lx1: int, ldim: int) -> np.ndarray:
    """(nel, lx1^ldim) -> (nel, lx2^ldim): GLL(lx1) -> GL(lx2) interpolation [UPSTREAM map_pm1_to_pr]."""
    lx2 = lx1 - 2
    J = sem.lagrange_interp_matrix(sem.zwgl(lx2)[0], sem.zwgll(lx1)[0])
    a = np.asarray(p1, float).reshape((-1,) + (lx1,) * ldim)
    return _interp_tensor(a, J, ldim).reshape(a.shape[0], -1)


class GatherComm:
    """Host plumbing for multi-rank checkpoints: `rank`, `world` and `gather(obj)` -> list of every rank's object on rank 0
    (None elsewhere).  `from_torch()` wraps an initialised torch.distributed group; in a drop-in the reference's own `outpost`
    does this gather over MPI [UPSTREAM prepost.f]."""

    def __init__(self, rank: int, world: int, gather):
        self.rank, self.world, self.gather = rank, world, gather

    @staticmethod
    def from_torch():
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()

        def gather(obj):
            out = [None] * world if rank == 0 else None
            dist.gather_object(obj, out, dst=0)
            return out
        return GatherComm(rank, world, gather)


def _is_partial(case) -> bool:
    return case.nelg is not None and int(case.nelg) != int(case.nel)


def write_krylov_vector(path: str, case, v: np.ndarray, p: np.ndarray, *, time: float = 0.0, istep: int = 0, wdsize: int = 8,
                        comm: "GatherComm | None" = None, theta: "np.ndarray | None" = None) -> None:
    """One Krylov vector as ONE global Nek field file (velocity + pressure on mesh 1), like the reference's `outpost`: element
    order = rank-major, ascending global id within a rank, `nelg` = the global element count.  A case that holds
    only one rank's share of the mesh needs `comm` (every rank calls; rank 0 writes); without it this raises instead of writing
    a truncated file that `read_krylov_vector` would later accept."""
    d, nel, L = case.ldim, case.nel, case.lx1
    shp = (nel,) + ((L,) * 3 if d == 3 else (1, L, L))
    U = np.asarray(v, float).reshape(d, nel, -1).transpose(1, 0, 2).reshape((nel, d) + shp[1:])
    P = pressure_to_mesh1(np.asarray(p, float).reshape(nel, -1), L, d).reshape(shp)
    T = None if theta is None else np.asarray(theta, float).reshape(shp)          # the vector's scalar (`ifheat`): field T of the file
    elmap = np.asarray(case.lglel if case.lglel is not None else np.arange(1, nel + 1), dtype=np.int32)
    if _is_partial(case):
        if comm is None:
            raise ValueError(f"write_krylov_vector: the case holds {nel} of {case.nelg} elements (one rank's share); pass comm= "
                             "(restart.GatherComm) so that rank 0 can write one global file")
        parts = comm.gather((elmap, U, P, T))
        if comm.rank != 0:
            return
        elmap = np.concatenate([q[0] for q in parts])
        U = np.concatenate([q[1] for q in parts])
        P = np.concatenate([q[2] for q in parts])
        T = None if T is None else np.concatenate([q[3] for q in parts])
        if elmap.size != int(case.nelg) or np.unique(elmap).size != elmap.size:
            raise ValueError(f"write_krylov_vector: gathered {elmap.size} elements, expected {case.nelg} distinct ones")
    nekio.write_field(path, U=U, P=P, T=T, time=time, istep=istep, wdsize=wdsize, elmap=elmap)


def read_krylov_vector(path: str, case, with_theta: bool = False):
    """Inverse of `write_krylov_vector`: (ldim, nel, npts) velocity and (nel, lx2^ldim) pressure in the case's element order;
    with_theta: also the scalar (nel, npts) (zeros when the file has no T field)."""
    ff = nekio.read_field(path)
    d, nel, L = case.ldim, case.nel, case.lx1
    if ff.nx != L or ff.ldim != d or ff.nel < nel:
        raise ValueError(f"{path}: {ff.nel} elements of {ff.nx}^{ff.ldim} points, expected >= {nel} of {L}^{d}")
    want = np.asarray(case.lglel if case.lglel is not None else np.arange(1, nel + 1))
    pos = {int(g): i for i, g in enumerate(ff.elmap)}
    try:
        order = np.array([pos[int(g)] for g in want])       # a rank's share picks its own elements out of a global file
    except KeyError as e:
        raise ValueError(f"{path}: global element {e} not in the file") from None
    v = ff.data["U"][order].reshape(nel, d, -1).transpose(1, 0, 2).copy()
    p = pressure_to_mesh2(ff.data["P"][order].reshape(nel, -1), L, d) if "P" in ff.data else np.zeros((nel, (L - 2) ** d))
    if with_theta:
        t = ff.data["T"][order].reshape(nel, -1).copy() if "T" in ff.data else np.zeros((nel, L ** d))
        return v, p, t
    return v, p


def arnoldi_checkpoint(ctx, case, session: str, H: np.ndarray, k: int, slot: int, *, outdir: str = ".", evop: str = "d",
                       tau: float = 1.0, eigen_tol: float = 1e-6, wdsize: int = 8, comm: "GatherComm | None" = None) -> int:
    """core/eigensolvers.f:802-905: write Krylov vector k+1 (device slot `slot`), the spectra of H(1:k,1:k) and H itself.
    Returns the number of Ritz pairs with residual |H(k+1,k) y_k| below eigen_tol (the count the reference logs).
    Multi-rank: every rank calls with `comm`; the KRY file is gathered into one global file and, like the reference
    (`if (nid .eq. 0)` :866-889), only rank 0 writes the HES and Spectre files."""
    v, p = ctx.vec_download(slot)
    theta = ctx.vec_download_scalar(slot) if getattr(ctx, "scalar_on", False) else None
    write_krylov_vector(os.path.join(outdir, kry_filename(session, k + 1)), case, v, p, time=tau * k, istep=k, wdsize=wdsize, comm=comm,
                        theta=theta)
    vals, vecs = np.linalg.eig(H[:k, :k])
    order = np.argsort(-np.abs(vals), kind="stable")          # `eig` sorts by decreasing magnitude (core/lapack_wrapper.f:129)
    vals, vecs = vals[order], vecs[:, order]
    residual = np.abs(H[k, k - 1] * vecs[k - 1, :])
    if comm is not None and comm.rank != 0:
        return int(np.count_nonzero(residual < eigen_tol))
    write_spectrum(os.path.join(outdir, "Spectre_H%s%04d.dat" % (evop, k)), vals, residual)
    write_spectrum(os.path.join(outdir, "Spectre_NS%s%04d.dat" % (evop, k)), log_transform(vals, tau), residual)
    write_hessenberg(os.path.join(outdir, hes_filename(session, k)), H, k)
    return int(np.count_nonzero(residual < eigen_tol))


def load_restart(ctx, case, session: str, mstart: int, k_dim: int, *, indir: str = ".", first_slot: int = 0) -> Tuple[np.ndarray, int]:
    """Restart branch of krylov_schur (core/eigensolvers.f:284-325): H from HES<session><mstart>, Krylov vectors 1..mstart+1
    from the KRY files into device slots first_slot.., returns (H padded to (k_dim+1, k_dim), the next Arnoldi step)."""
    Hs = read_hessenberg(os.path.join(indir, hes_filename(session, mstart)), mstart)
    H = np.zeros((k_dim + 1, k_dim), order="F")
    H[:mstart + 1, :mstart] = Hs
    for i in range(1, mstart + 2):
        v, p = read_krylov_vector(os.path.join(indir, kry_filename(session, i)), case)
        ctx.vec_upload(first_slot + i - 1, v, p)
    return H, mstart + 1
