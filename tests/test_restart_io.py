"""CPU tests of the checkpoint/restart file protocol (nekstab_b200/restart.py; reference: core/eigensolvers.f:284-325,
802-905, core/IO.f:15-60): text formats pinned byte for byte to the shipped spectrum files, binary round trips exact."""
import os

import numpy as np
import pytest

from nekstab_b200 import restart
from util import GOLD, small_cases


def test_spectrum_writer_matches_shipped_text(tmp_path):
    for name in ("Spectre_Hd_head.dat", "Spectre_NSd_head.dat"):
        ref = open(os.path.join(GOLD, name)).read()
        rows = np.loadtxt(os.path.join(GOLD, name), ndmin=2)
        out = tmp_path / name
        restart.write_spectrum(str(out), rows[:, 0] + 1j * rows[:, 1], rows[:, 2])
        assert open(out).read() == ref                      # '(3E15.7)': 0.dddddddE+ee, byte for byte


def test_log_transform_reproduces_shipped_ns_spectrum():
    g = np.load(os.path.join(GOLD, "cyl.npz"))
    h, ns = g["Spectre_Hd"], g["Spectre_NSd"]
    lam = restart.log_transform(h[:, 0] + 1j * h[:, 1], 1.0)    # tau = dt*nsteps = 1 (SURVEY 8c relation check)
    big = np.abs(h[:, 0] + 1j * h[:, 1]) > 1e-3
    assert np.abs(lam.real - ns[:, 0])[big].max() < 2e-6 and np.abs(np.abs(lam.imag) - np.abs(ns[:, 1]))[big].max() < 2e-6


def test_fortran_e_edge_cases():
    assert restart.fortran_e(0.0) == "  0.0000000E+00"
    assert restart.fortran_e(-0.6972442) == " -0.6972442E+00"
    assert restart.fortran_e(9.99999999e-4) == "  0.1000000E-02"      # rounding carries into the exponent
    assert restart.fortran_e(5.095353e-19) == "  0.5095353E-18"


def test_hessenberg_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    k = 7
    H = np.triu(rng.standard_normal((k + 1, k)), -1)
    path = str(tmp_path / restart.hes_filename("1cyl", k))
    assert os.path.basename(path) == "HES1cyl0007"
    restart.write_hessenberg(path, H, k)
    assert np.array_equal(restart.read_hessenberg(path, k), H)          # %.17E is exact for float64
    # Fortran list-directed output may wrap lines and use D exponents
    with open(path, "w") as f:
        f.write("\n".join("  %.15E" % v for v in H.ravel()).replace("E", "D"))
    assert np.abs(restart.read_hessenberg(path, k) - H).max() < 1e-14


def test_pressure_mesh_maps_are_inverse():
    rng = np.random.default_rng(1)
    for ldim, lx1 in ((2, 6), (3, 8), (3, 4)):
        p2 = rng.standard_normal((5, (lx1 - 2) ** ldim))
        p1 = restart.pressure_to_mesh1(p2, lx1, ldim)
        assert p1.shape == (5, lx1 ** ldim)
        assert np.abs(restart.pressure_to_mesh2(p1, lx1, ldim) - p2).max() < 1e-12


def test_krylov_vector_file_round_trip(tmp_path):
    assert restart.kry_filename("1cyl", 12) == "KRY1cyl0.f00012"
    for name in ("box2d_n6_outflow", "box3d_n8_outflow"):
        c = small_cases()[name]
        rng = np.random.default_rng(2)
        v = rng.standard_normal((c.ldim, c.nel, c.npts))
        p = rng.standard_normal((c.nel, c.lx2 ** c.ldim))
        path = str(tmp_path / restart.kry_filename(name, 3))
        restart.write_krylov_vector(path, c, v, p, time=2.0, istep=2, wdsize=8)
        v2, p2 = restart.read_krylov_vector(path, c)
        assert np.array_equal(v2, v) and np.abs(p2 - p).max() < 1e-12
        # a rank's share reads its own elements out of a file written in another element order
        part = c.local_part(1, 2)
        v3, p3 = restart.read_krylov_vector(path, part)
        sel = part.lglel - 1
        assert np.array_equal(v3, v[:, sel]) and np.abs(p3 - p[sel]).max() < 1e-12
        # the vector's scalar (`ifheat`) travels as the T field; files without one read back zeros
        th = rng.standard_normal((c.nel, c.npts))
        restart.write_krylov_vector(path, c, v, p, theta=th)
        v5, p5, t5 = restart.read_krylov_vector(path, c, with_theta=True)
        assert np.array_equal(v5, v) and np.array_equal(t5, th) and np.abs(p5 - p).max() < 1e-12
        _, _, t6 = restart.read_krylov_vector(path, part, with_theta=True)
        assert np.array_equal(t6, th[sel])
        restart.write_krylov_vector(path, c, v, p)
        assert not restart.read_krylov_vector(path, c, with_theta=True)[2].any()
        # single precision files (writeDoublePrecision = no, 1cyl.par:20) lose digits but keep the layout
        restart.write_krylov_vector(path, c, v, p, wdsize=4)
        v4, _ = restart.read_krylov_vector(path, c)
        assert np.abs(v4 - v).max() < 1e-6 * np.abs(v).max()


def test_log_transform_negative_real_ritz_value_has_zero_imaginary_part():
    """core/eigensolvers.f:908-915: `if (aimag(x) .eq. 0) log_transform = real(log_transform)` -- not pi/tau."""
    lam = restart.log_transform(np.array([-0.5 + 0j, 0.5 + 0j, -0.5 + 1e-30j, 0.3 - 0.4j]), 2.0)
    assert lam[0].imag == 0.0 and abs(lam[0].real - np.log(0.5) / 2.0) < 1e-15
    assert lam[1].imag == 0.0
    assert abs(lam[2].imag - np.pi / 2.0) < 1e-12            # a (tiny) non-zero imaginary part keeps the principal value
    assert abs(lam[3] - np.log(0.3 - 0.4j) / 2.0) < 1e-15


def test_partial_case_needs_a_communicator(tmp_path):
    """a rank's share must not be written as if it were the whole mesh (the last writer used to win)."""
    c = small_cases()["box2d_n6_outflow"]
    part = c.local_part(0, 2)
    v = np.zeros((c.ldim, part.nel, c.npts)); p = np.zeros((part.nel, c.lx2 ** c.ldim))
    with pytest.raises(ValueError, match="comm="):
        restart.write_krylov_vector(str(tmp_path / "KRYx0.f00001"), part, v, p)


def test_multirank_checkpoint_gathers_one_global_file(tmp_path):
    """Two ranks' shares gathered through restart.GatherComm into ONE file with nelg = the global count, rank-major element
    order (the reference's `outpost` layout); each share and the global case read their elements back."""
    c = small_cases()["box3d_n8_outflow"]
    rng = np.random.default_rng(7)
    v = rng.standard_normal((c.ldim, c.nel, c.npts))
    p = rng.standard_normal((c.nel, c.lx2 ** c.ldim))
    parts = [c.local_part(r, 2) for r in range(2)]
    sels = [q.lglel - 1 for q in parts]
    box = []                                                     # in-process stand-in for gather_object: rank 1 calls first
    path = str(tmp_path / restart.kry_filename("box", 2))
    for r in (1, 0):
        def gather(obj, r=r):
            box.append((r, obj))
            return [o for _, o in sorted(box, key=lambda t: t[0])] if r == 0 else None
        restart.write_krylov_vector(path, parts[r], v[:, sels[r]], p[sels[r]], comm=restart.GatherComm(r, 2, gather))
    from nekstab_b200 import nekio
    ff = nekio.read_field(path)
    assert ff.nel == c.nel and list(ff.elmap) == list(np.concatenate([q.lglel for q in parts]))
    v2, p2 = restart.read_krylov_vector(path, c)
    assert np.array_equal(v2, v) and np.abs(p2 - p).max() < 1e-12
    for r in range(2):
        v3, p3 = restart.read_krylov_vector(path, parts[r])
        assert np.array_equal(v3, v[:, sels[r]])
    # ... and the vector's scalar (`ifheat`) travels through the same gather
    th = rng.standard_normal((c.nel, c.npts))
    box.clear()
    for r in (1, 0):
        def gather(obj, r=r):
            box.append((r, obj))
            return [o for _, o in sorted(box, key=lambda t: t[0])] if r == 0 else None
        restart.write_krylov_vector(path, parts[r], v[:, sels[r]], p[sels[r]], comm=restart.GatherComm(r, 2, gather), theta=th[sels[r]])
    assert np.array_equal(restart.read_krylov_vector(path, c, with_theta=True)[2], th)
    assert np.array_equal(restart.read_krylov_vector(path, parts[1], with_theta=True)[2], th[sels[1]])
