"""GPU test of the Arnoldi checkpoint / restart protocol (core/eigensolvers.f:284-325, 802-905): an interrupted factorisation
resumed from its KRY / HES files in a fresh context reproduces the uninterrupted one."""
import numpy as np
import pytest

from util import small_cases, smooth_field

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["box2d_n6_outflow", "box3d_n6_dirichlet"])
def test_arnoldi_restart_reproduces_uninterrupted_run(name, tmp_path):
    from nekstab_b200 import lib, restart
    c = small_cases()[name]
    k_dim, k_stop = 6, 3

    def fresh():
        g = lib.NekStabB200(c)
        g.set_params(1.0 / c.re, 1.0, 1e-12, 1e-12, 3000, 100000)
        g.set_pressure_preconditioner(1, 2)
        g.set_timestep(2.0e-3, 4)
        g.vec_alloc(k_dim + 2)
        return g

    g = fresh()
    try:
        g.vec_upload(0, smooth_field(c, 31), None)
        g.normalize(0)
        H_ref = np.zeros((k_dim + 1, k_dim), order="F")
        g.arnoldi_factorization(lib.DIRECT, 0, H_ref, 1, k_dim, k_dim)
        q_ref = [g.vec_download(i)[0].copy() for i in range(k_dim + 1)]
    finally:
        g.close()
    # interrupted run: checkpoint after every step like `ifres = .true.` (core/krylov_decomposition.f:89), stop after k_stop
    g = fresh()
    try:
        g.vec_upload(0, smooth_field(c, 31), None)
        g.normalize(0)
        v, p = g.vec_download(0)
        restart.write_krylov_vector(str(tmp_path / restart.kry_filename("t", 1)), c, v, p)      # core/eigensolvers.f:280-282
        H = np.zeros((k_dim + 1, k_dim), order="F")
        for k in range(1, k_stop + 1):
            g.arnoldi_factorization(lib.DIRECT, 0, H, k, k, k_dim)
            cnt = restart.arnoldi_checkpoint(g, c, "t", H, k, k, outdir=str(tmp_path), tau=8.0e-3)
            assert 0 <= cnt <= k
    finally:
        g.close()
    for fn in ("HESt0003", "KRYt0.f00004", "Spectre_Hd0003.dat", "Spectre_NSd0003.dat"):
        assert (tmp_path / fn).exists(), fn
    # resumed run in a fresh context
    g = fresh()
    try:
        H2, mstart = restart.load_restart(g, c, "t", k_stop, k_dim, indir=str(tmp_path))
        assert mstart == k_stop + 1
        assert np.array_equal(H2[:k_stop + 1, :k_stop], H[:k_stop + 1, :k_stop])
        g.arnoldi_factorization(lib.DIRECT, 0, H2, mstart, k_dim, k_dim)
        assert np.abs(H2 - H_ref).max() < 1e-9 * np.abs(H_ref).max()
        for i in (k_stop + 1, k_dim):
            q = g.vec_download(i)[0]
            assert np.linalg.norm(q - q_ref[i]) < 1e-8 * np.linalg.norm(q_ref[i])
    finally:
        g.close()
