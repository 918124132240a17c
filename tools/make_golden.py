#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's shipped fixtures (run in the build container only;
/root/reference does not exist on the GPU box).  Usage: python tools/make_golden.py [/root/reference]

What is extracted (SURVEY.md 8c "golden vectors"):
  cyl.npz : examples/cylinder/stability/direct/{BF_1cyl0.f00001, 1cyl.ma2, 1cyl.re2, Spectre_*.dat},
            .../adjoint/Spectre_*.dat, .../postproc/sensitivity_budget_wavemaker/{dRe,dIm}1cyl0.f00001,
            element->rank maps of the field files (partition KAT)
  bfs.npz : examples/back_fstep/transient_growth/{BF_bfs0,pRebfs0,orebfs0}.f00001, bfs.ma2, bfs.re2
  cyl_upo.npz : examples/cylinder/stability/direct_Floquet/{BF_1cyl0.f00001 (the UPO snapshot), Spectre_Hd.dat, Spectre_NSd_conv.dat}
  cav.npz : examples/lid_driven/{BF_cav0.f00001, cav.ma2, cav.re2} (config 3; the shipped base flow lives on y in [0, 1.2]
            although cav.par:9 sets the aspect ratio 1.5 that cav.usr:107-109 rescales the mesh to: the fixture's own
            coordinates are kept, SURVEY.md 8 cfg-3 caveat)
  cyl_re40.npz : examples/cylinder/baseflow/newton/BFRe40_1cyl0.f00001 (config 2: the Newton-Krylov case's initial condition)
  tsyphon.npz : examples/thersyphon/baseflow/{BF_Ra400_tsyphon0.f00001, tsyphon.ma2, tsyphon.re2} (scalar transport: the shipped
            Boussinesq base flow with temperature)
All element data are re-ordered to ascending global element id.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekstab_b200 import nekio  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
REF = _pos[0] if _pos else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def runs_to_rank(ff):
    rank = np.empty(ff.nelg, dtype=np.int16)
    for r, run in enumerate(ff.rank_runs()):
        rank[run - 1] = r
    return rank


def cyl():
    d = f"{REF}/examples/cylinder/stability/direct"
    a = f"{REF}/examples/cylinder/stability/adjoint"
    p = f"{REF}/examples/cylinder/postproc/sensitivity_budget_wavemaker"
    bf_raw = nekio.read_field(f"{d}/BF_1cyl0.f00001")
    bf = bf_raw.sort_global()
    re2 = nekio.read_re2(f"{d}/1cyl.re2")
    ma2 = nekio.read_ma2(f"{d}/1cyl.ma2")
    names = np.array(["E  ", "P  ", "v  ", "O  ", "W  "])
    codes = re2.bc_codes()
    bc = np.zeros(codes.shape, dtype=np.uint8)
    for i, nme in enumerate(names):
        bc[codes == nme] = i
    dre = nekio.read_field(f"{p}/dRe1cyl0.f00001").sort_global()
    dim = nekio.read_field(f"{p}/dIm1cyl0.f00001").sort_global()
    bf40 = nekio.read_field(f"{REF}/examples/cylinder/baseflow/newton/BFRe40_1cyl0.f00001") \
        if os.path.exists(f"{REF}/examples/cylinder/baseflow/newton/BFRe40_1cyl0.f00001") else None
    out = dict(
        lx1=bf.nx, X=bf.data["X"][:, :, 0], U=bf.data["U"][:, :, 0], P=bf.data["P"][:, 0],
        vert=ma2.vert.astype(np.int32), key=ma2.key.astype(np.int32), d2=ma2.d2,
        bc=bc, bc_names=names.astype("S3"),
        dRe_U=dre.data["U"][:, :, 0].astype(np.float32), dRe_P=dre.data["P"][:, 0].astype(np.float32),
        dIm_U=dim.data["U"][:, :, 0].astype(np.float32), dIm_P=dim.data["P"][:, 0].astype(np.float32),
        mode_istep=dre.istep,
        Spectre_Hd=nekio.read_spectre(f"{d}/Spectre_Hd.dat"),
        Spectre_NSd=nekio.read_spectre(f"{d}/Spectre_NSd.dat"),
        Spectre_NSd_conv=nekio.read_spectre(f"{d}/Spectre_NSd_conv.dat"),
        Spectre_Ha=nekio.read_spectre(f"{a}/Spectre_Ha.dat"),
        Spectre_NSa_conv=nekio.read_spectre(f"{a}/Spectre_NSa_conv.dat"),
        rank_p6=runs_to_rank(bf_raw),
    )
    if bf40 is not None:
        out["rank_p4"] = runs_to_rank(bf40)
    np.savez_compressed(f"{OUT}/cyl.npz", **out)
    print("cyl.npz", os.path.getsize(f"{OUT}/cyl.npz") / 1e6, "MB")


def bfs():
    d = f"{REF}/examples/back_fstep/transient_growth"
    bf_raw = nekio.read_field(f"{d}/BF_bfs0.f00001")
    bf = bf_raw.sort_global()
    pre = nekio.read_field(f"{d}/pRebfs0.f00001").sort_global()
    pim = nekio.read_field(f"{d}/pImbfs0.f00001").sort_global()
    ore = nekio.read_field(f"{d}/orebfs0.f00001").sort_global()
    re2 = nekio.read_re2(f"{d}/bfs.re2")
    ma2 = nekio.read_ma2(f"{d}/bfs.ma2")
    f32 = np.float32
    out = dict(
        lx1=bf.nx, re2_xyz=re2.xyz, X=bf.data["X"][:, :, 0].astype(f32), U=bf.data["U"][:, :, 0].astype(f32),
        vert=ma2.vert.astype(np.int32), key=ma2.key.astype(np.int32), d2=ma2.d2,
        bc_id=re2.bc_param(4).astype(np.uint8),
        pRe_U=pre.data["U"][:, :, 0].astype(f32), pRe_P=pre.data["P"][:, 0].astype(f32),
        pIm_absmax=np.abs(pim.data["U"]).max(),
        ore_U=ore.data["U"][:, :, 0].astype(f32), ore_P=ore.data["P"][:, 0].astype(f32),
        mode_istep=pre.istep, rank_p4=runs_to_rank(bf_raw),
        rank_p6=runs_to_rank(nekio.read_field(f"{d}/pRebfs0.f00001")),
        barkley=np.loadtxt(f"{REF}/examples/back_fstep/barkley2008_fig5.ref", ndmin=2)
        if os.path.exists(f"{REF}/examples/back_fstep/barkley2008_fig5.ref") else np.zeros((0, 2)),
    )
    np.savez_compressed(f"{OUT}/bfs.npz", **out)
    print("bfs.npz", os.path.getsize(f"{OUT}/bfs.npz") / 1e6, "MB")


def cav():
    d = f"{REF}/examples/lid_driven"
    bf_raw = nekio.read_field(f"{d}/BF_cav0.f00001")
    bf = bf_raw.sort_global()
    re2 = nekio.read_re2(f"{d}/cav.re2")
    ma2 = nekio.read_ma2(f"{d}/cav.ma2")
    names = np.array(["E  ", "P  ", "v  ", "O  ", "W  "])
    codes = re2.bc_codes()
    bc = np.zeros(codes.shape, dtype=np.uint8)
    for i, nme in enumerate(names):
        bc[codes == nme] = i
    out = dict(lx1=bf.nx, X=bf.data["X"][:, :, 0], U=bf.data["U"][:, :, 0], P=bf.data["P"][:, 0], time=bf.time, istep=bf.istep,
               vert=ma2.vert.astype(np.int32), key=ma2.key.astype(np.int32), d2=ma2.d2, bc=bc, bc_names=names.astype("S3"),
               re2_xyz=re2.xyz, rank_file=runs_to_rank(bf_raw))
    np.savez_compressed(f"{OUT}/cav.npz", **out)
    print("cav.npz", os.path.getsize(f"{OUT}/cav.npz") / 1e6, "MB")


def tsyphon():
    """examples/thersyphon/baseflow: the Newton-converged steady thermosyphon at Ra = 400 (annulus r in [1, 2], periodic in theta,
    walls 'W' / 't' at both radii; Pr = 5: tsyphon.par viscosity = 5, conductivity = 1, rhocp = 1; buoyancy ffy = T * Pr * Ra,
    tsyphon.usr userf).  A double-precision field file with coordinates: the curved-element GLL points are taken from it."""
    d = f"{REF}/examples/thersyphon/baseflow"
    bf = nekio.read_field(f"{d}/BF_Ra400_tsyphon0.f00001").sort_global()
    re2 = nekio.read_re2(f"{d}/tsyphon.re2")
    ma2 = nekio.read_ma2(f"{d}/tsyphon.ma2")
    names = np.array(["E  ", "P  ", "W  ", "t  "])
    out = dict(lx1=bf.nx, X=bf.data["X"][:, :, 0], U=bf.data["U"][:, :, 0], P=bf.data["P"][:, 0], T=bf.data["T"][:, 0], time=bf.time,
               istep=bf.istep, vert=ma2.vert.astype(np.int32), key=ma2.key.astype(np.int32), d2=ma2.d2, bc_names=names.astype("S3"))
    for nm, ifield in (("bc", 0), ("bct", 1)):
        codes = re2.bc_codes(ifield)
        bc = np.zeros(codes.shape, dtype=np.uint8)
        for i, nme in enumerate(names):
            bc[codes == nme] = i
        out[nm] = bc
    np.savez_compressed(f"{OUT}/tsyphon.npz", **out)
    print("tsyphon.npz", os.path.getsize(f"{OUT}/tsyphon.npz") / 1e6, "MB")


def cyl_re40():
    """examples/cylinder/baseflow/newton: the Re = 40 steady flow the shipped Newton-Krylov case (config 2) starts from
    (`startFrom = BFRe40_1cyl0.f00001`, 1cyl.par:2; target Re = 50, endTime 1, k_dim 100, tolerances 1e-11)."""
    d = f"{REF}/examples/cylinder/baseflow/newton"
    bf = nekio.read_field(f"{d}/BFRe40_1cyl0.f00001").sort_global()
    out = dict(lx1=bf.nx, U=bf.data["U"][:, :, 0].astype(np.float32), P=bf.data["P"][:, 0].astype(np.float32), time=bf.time, istep=bf.istep)
    np.savez_compressed(f"{OUT}/cyl_re40.npz", **out)
    print("cyl_re40.npz", os.path.getsize(f"{OUT}/cyl_re40.npz") / 1e6, "MB")


def cyl_upo():
    """examples/cylinder/stability/direct_Floquet: the periodic orbit snapshot the Floquet example starts from (`startFrom =
    BF_1cyl0.f00001 # here UPO file`, 1cyl.par:2; time = the period 7.9213, istep = 796) and its shipped Floquet spectrum."""
    d = f"{REF}/examples/cylinder/stability/direct_Floquet"
    bf = nekio.read_field(f"{d}/BF_1cyl0.f00001").sort_global()
    out = dict(lx1=bf.nx, U=bf.data["U"][:, :, 0], P=bf.data["P"][:, 0], time=bf.time, istep=bf.istep,
               Spectre_Hd=nekio.read_spectre(f"{d}/Spectre_Hd.dat"), Spectre_NSd_conv=nekio.read_spectre(f"{d}/Spectre_NSd_conv.dat"))
    np.savez_compressed(f"{OUT}/cyl_upo.npz", **out)
    print("cyl_upo.npz", os.path.getsize(f"{OUT}/cyl_upo.npz") / 1e6, "MB")


def spectrum_text():
    """First lines of the shipped spectrum files as TEXT: pins the '(3E15.7)' writer of nekstab_b200/restart.py byte for byte."""
    d = f"{REF}/examples/cylinder/stability/direct"
    for name in ("Spectre_Hd.dat", "Spectre_NSd.dat"):
        lines = open(f"{d}/{name}").read().splitlines()[:24]
        with open(f"{OUT}/{name.replace('.dat', '_head.dat')}", "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if "--cav-only" in sys.argv:
        cav()
        sys.exit(0)
    if "--upo-only" in sys.argv:
        cyl_upo()
        sys.exit(0)
    if "--re40-only" in sys.argv:
        cyl_re40()
        sys.exit(0)
    if "--tsyphon-only" in sys.argv:
        tsyphon()
        sys.exit(0)
    if "--spectrum-only" not in sys.argv:
        cyl()
        bfs()
        cav()
        cyl_upo()
        tsyphon()
        cyl_re40()
    spectrum_text()
