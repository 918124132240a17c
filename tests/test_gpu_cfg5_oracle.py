"""The cfg-5 matvec against the C / OpenMP oracle (oracle/cport.c, pinned to the numpy oracle in tests/test_oracle_cport.py and
through it to the reference's fixtures): 2 steps of the direct map at tolerances 1e-12 on the full 1996-element 2-D cylinder mesh x
NSB_FULLSIZE_NZ periodic layers (default 2 = 3 992 hexahedra, 2.04e6 points; NSB_FULLSIZE_NZ=10 is the full 19 960-element
workload of BASELINE.json configs[4], ~4 min of CPU: run once per round under gpurun, log in profiles/).  North-star tolerance:
1e-10 relative in the energy norm per matvec.  Through the C ABI."""
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)


def test_cfg5_matvec_against_the_c_oracle():
    """VERDICT r1 parity gap (a): the 3-D cfg-5 matvec checked against the oracle, not only through properties."""
    import os
    import bench
    from nekstab_b200 import cases, lib
    from oracle import cport
    from oracle.ops import SEM
    from oracle.pmg import PMG
    nz = int(os.environ.get("NSB_FULLSIZE_NZ", "2"))
    nsteps, tol = 2, 1e-12
    c, _ = bench.build_workload(nz)
    cport.set_threads(0)
    s = SEM(3, 8, c.xyz, c.glo, c.mask)
    cp0 = cport.CPort(s, None)
    pc = PMG(s, nagg=max(1, min(512, c.nel // 32)), apply_e=lambda p: cp0.cdabdtp(p).reshape(s.eshape2))
    st = cport.CStepper(s, c.ubase, c.re, None, tol_v=tol, tol_p=tol, max_iter_v=3000, max_iter_p=100000, ifvcor=False, pmg=pc)
    v0 = cases.add_noise(c).reshape((3,) + s.eshape)
    v0 = v0 / np.sqrt(sum(float(np.sum(v0[k] * s.bm1 * v0[k])) for k in range(3)))
    p0 = np.zeros(s.eshape2)
    g = lib.NekStabB200(c)
    try:
        g.set_params(1.0 / c.re, 1.0, tol, tol, 3000, 100000)
        g.set_pressure_preconditioner(1, 0)
        dt, _, _ = g.prepare_linearized_solver(1.0, 0.5)
        g.set_timestep(dt, nsteps)
        g.vec_alloc(2)
        g.vec_upload(0, v0, p0)
        g.matvec(lib.DIRECT, 0, 1)
        v, p = g.vec_download(1)
        stats = g.stats()
    finally:
        g.close()
    vo, po = st.linearized_map(v0, p0, nsteps, dt)
    vo = vo.reshape(3, -1)
    bm1 = s.bm1.reshape(-1)
    err = float(np.sqrt(np.sum((v - vo) ** 2 * bm1[None]) / np.sum(vo ** 2 * bm1[None])))
    perr = float(np.linalg.norm(p - po.ravel()) / np.linalg.norm(po.ravel()))
    print(f"cfg5 nz={nz} ({c.nel} hexahedra, n={c.n}): matvec({nsteps} steps) vs oracle/cport: energy-norm rel err {err:.3e}, "
          f"pressure rel err {perr:.3e}; GPU iterations {stats['pres_iters']} / {stats['helm_iters']}, CPU {st.iters_p} / {st.iters_v}")
    assert err < 1e-10, err
    assert perr < 1e-6, perr
