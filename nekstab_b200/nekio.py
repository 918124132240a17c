"""Readers/writers for the Nek5000 on-disk formats that sit either side of the hot path.

These are the data formats nekStab feeds the Krylov loop from:
  * ``<prefix><session>0.f%05d`` field files  (load_fld / outpost; core/eigensolvers.f:183,282)
  * ``.re2`` binary mesh (vertices, curved sides, boundary conditions)
  * ``.ma2`` binary map (RSB partition key + global vertex ids per element)
Only numpy is used.  Nothing here touches the GPU.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

_TAG = 6.54321


def _endian_from_tag(raw: bytes) -> str:
    for e in ("<", ">"):
        if abs(struct.unpack(e + "f", raw)[0] - _TAG) < 1e-4:
            return e
    raise ValueError("endian tag 6.54321 not found")


@dataclass
class FieldFile:
    """Contents of one Nek5000 ``.f%05d`` file (elements in *global id* order after `sort_global`)."""
    wdsize: int
    nx: int
    ny: int
    nz: int
    nel: int
    nelg: int
    time: float
    istep: int
    rdcode: str
    elmap: np.ndarray                      # (nel,) global element ids, 1-based, file order
    data: Dict[str, np.ndarray] = field(default_factory=dict)   # 'X','U': (nel, ldim, nz, ny, nx); 'P': (nel,nz,ny,nx); 'T'

    @property
    def ldim(self) -> int:
        return 3 if self.nz > 1 else 2

    def sort_global(self) -> "FieldFile":
        order = np.argsort(self.elmap, kind="stable")
        out = FieldFile(self.wdsize, self.nx, self.ny, self.nz, self.nel, self.nelg, self.time,
                        self.istep, self.rdcode, self.elmap[order])
        out.data = {k: v[order] for k, v in self.data.items()}
        return out

    def rank_runs(self) -> List[np.ndarray]:
        """Split the file's element map into ascending runs = writer MPI ranks."""
        brk = np.nonzero(np.diff(self.elmap) < 0)[0] + 1
        return np.split(self.elmap, brk)


def read_field(path: str) -> FieldFile:
    with open(path, "rb") as f:
        hdr = f.read(132).decode("ascii", errors="replace")
        tok = hdr.split()
        if tok[0] != "#std":
            raise ValueError(f"{path}: not a #std field file")
        wdsize, nx, ny, nz, nel, nelg = (int(t) for t in tok[1:7])
        time = float(tok[7])
        istep = int(tok[8])
        rdcode = tok[11]
        e = _endian_from_tag(f.read(4))
        elmap = np.fromfile(f, dtype=e + "i4", count=nel)
        ldim = 3 if nz > 1 else 2
        npt = nx * ny * nz
        ft = e + ("f4" if wdsize == 4 else "f8")
        ff = FieldFile(wdsize, nx, ny, nz, nel, nelg, time, istep, rdcode, elmap)
        i = 0
        while i < len(rdcode):
            c = rdcode[i]
            if c in "XU":
                a = np.fromfile(f, dtype=ft, count=nel * ldim * npt).astype(np.float64)
                ff.data[c] = a.reshape(nel, ldim, nz, ny, nx)
            elif c == "P":
                a = np.fromfile(f, dtype=ft, count=nel * npt).astype(np.float64)
                ff.data[c] = a.reshape(nel, nz, ny, nx)
            elif c == "T":
                a = np.fromfile(f, dtype=ft, count=nel * npt).astype(np.float64)
                ff.data[c] = a.reshape(nel, nz, ny, nx)
            elif c == "S":
                ns = int(rdcode[i + 1:i + 3])
                i += 2
                a = np.fromfile(f, dtype=ft, count=ns * nel * npt).astype(np.float64)
                ff.data[c] = a.reshape(ns, nel, nz, ny, nx)
            i += 1
    return ff


def write_field(path: str, X=None, U=None, P=None, T=None, *, time=0.0, istep=0, wdsize=8,
                elmap=None) -> None:
    """Write a single-file ``#std`` field file (layout of SURVEY App. A).  Arrays as in `FieldFile.data`."""
    ref = next(a for a in (X, U, P, T) if a is not None)
    nel = ref.shape[0]
    nz, ny, nx = ref.shape[-3:]
    if elmap is None:
        elmap = np.arange(1, nel + 1, dtype=np.int32)
    code = ("X" if X is not None else "") + ("U" if U is not None else "") + \
           ("P" if P is not None else "") + ("T" if T is not None else "")
    hdr = "#std %1d %2d %2d %2d %10d %10d %20.13E %9d %6d %6d %s" % (
        wdsize, nx, ny, nz, nel, nel, time, istep, 0, 1, code)
    hdr = hdr.ljust(132)[:132]
    ft = "<f4" if wdsize == 4 else "<f8"
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(struct.pack("<f", _TAG))
        np.asarray(elmap, dtype="<i4").tofile(f)
        for a in (X, U, P, T):
            if a is not None:
                np.ascontiguousarray(a, dtype=ft).tofile(f)


@dataclass
class Re2:
    ldim: int
    nel: int
    xyz: np.ndarray        # (nel, ldim, 2**ldim) vertex coordinates, Nek "preprocessor" vertex order
    curves: List[tuple]    # (elem(1-based), face(1-based), p1..p5, type)
    bcs: List[tuple]       # (elem, face, p1..p5, code)  -- velocity
    bcs_more: List[list] = field(default_factory=list)   # the same for the temperature and every passive scalar the file carries

    def bc_codes(self, ifield: int = 0) -> np.ndarray:
        """(nel, 2*ldim) array of 3-char codes, 'E  ' where no record exists.  ifield 0: velocity, 1: temperature, ..."""
        out = np.full((self.nel, 2 * self.ldim), "E  ", dtype="<U3")
        for (e, f, *_p, code) in (self.bcs if ifield == 0 else self.bcs_more[ifield - 1]):
            out[e - 1, f - 1] = code
        return out

    def bc_param(self, k: int = 4) -> np.ndarray:
        """(nel, 2*ldim) k-th real of each BC record (gmsh '#v003' meshes carry the boundary id in p5)."""
        out = np.zeros((self.nel, 2 * self.ldim))
        for rec in self.bcs:
            out[rec[0] - 1, rec[1] - 1] = rec[2 + k]
        return out


def read_re2(path: str) -> Re2:
    with open(path, "rb") as f:
        hdr = f.read(80).decode("ascii", errors="replace")
        tok = hdr.split()
        nel, ldim = int(tok[1]), int(tok[2])
        e = _endian_from_tag(f.read(4))
        nv = 2 ** ldim
        rec = 1 + ldim * nv
        a = np.fromfile(f, dtype=e + "f8", count=nel * rec).reshape(nel, rec)
        xyz = a[:, 1:].reshape(nel, ldim, nv)
        ncurve = int(np.fromfile(f, dtype=e + "f8", count=1)[0])
        curves = []
        for _ in range(ncurve):
            raw = f.read(64)
            v = struct.unpack(e + "7d", raw[:56])
            curves.append((int(v[0]), int(v[1]), *v[2:7], raw[56:57].decode("ascii", errors="replace")))   # 1 character, the rest of the word is padding
        nbc = int(np.fromfile(f, dtype=e + "f8", count=1)[0])
        bcs = []
        for _ in range(nbc):
            raw = f.read(64)
            v = struct.unpack(e + "7d", raw[:56])
            code = raw[56:64].decode("ascii")[:3]
            bcs.append((int(v[0]), int(v[1]), *v[2:7], code))
        # further fields (temperature, passive scalars): one more block each, same record format
        more = []
        while True:
            cnt = np.fromfile(f, dtype=e + "f8", count=1)
            if cnt.size == 0:
                break
            blk = []
            for _ in range(int(cnt[0])):
                raw = f.read(64)
                v = struct.unpack(e + "7d", raw[:56])
                blk.append((int(v[0]), int(v[1]), *v[2:7], raw[56:64].decode("ascii", errors="replace")[:3]))
            more.append(blk)
    r = Re2(ldim, nel, xyz, curves, bcs)
    r.bcs_more = more
    return r


@dataclass
class Ma2:
    nel: int
    depth: int
    d2: int
    key: np.ndarray        # (nel,) RSB leaf key in [0, d2)
    vert: np.ndarray       # (nel, 2**ldim) global vertex ids, lexicographic (r,s[,t]) order


def read_ma2(path: str) -> Ma2:
    with open(path, "rb") as f:
        hdr = f.read(132).decode("ascii", errors="replace")
        tok = hdr.split()
        nel, _nact, depth, d2, npts, _nrank, _nout = (int(t) for t in tok[1:8])
        e = _endian_from_tag(f.read(4))
        nv = npts // nel
        a = np.fromfile(f, dtype=e + "i4", count=nel * (nv + 1)).reshape(nel, nv + 1)
    return Ma2(nel, depth, d2, a[:, 0].copy(), a[:, 1:].copy())


def read_spectre(path: str) -> np.ndarray:
    """Rows of a ``Spectre_*.dat`` file (core/eigensolvers.f:590-604) as a float array."""
    return np.loadtxt(path, ndmin=2)
