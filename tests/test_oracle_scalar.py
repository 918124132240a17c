"""The oracle's scalar-transport restatement (oracle/scalar.py) pinned (1) to the reference's shipped Boussinesq base flow
(examples/thersyphon/baseflow/BF_Ra400_tsyphon0.f00001: a fixed point of the full velocity / temperature stepper at Ra = 400 and at no other
Rayleigh number) and (2) to itself: the direct perturbation map is the Jacobian of the full map."""
import os

import numpy as np

from util import GOLD, make_oracle, small_cases, smooth_field


def test_kat_thermal_the_shipped_thermosyphon_is_a_fixed_point_at_ra_400():
    from nekstab_b200 import cases, restart
    from oracle.scalar import ScalarStepper
    from oracle.stepper import prepare_linearized_solver
    c = cases.thermosyphon_case(np.load(os.path.join(GOLD, "tsyphon.npz")))
    s = make_oracle(c)
    assert abs(s.vol - 3.0 * np.pi) < 1e-7                                       # the annulus 1 <= r <= 2 from the file's curved-element GLL points
    U, T, tm = c.ubase.reshape((2,) + s.eshape), c.extra["T"].reshape(s.eshape), c.extra["tmask"].reshape(s.eshape)
    assert np.abs((1 - tm) * (T - 0.5 * (1 + np.tanh(-20 * s.X[1])))).max() < 1e-10   # userbc: temp = 0.5 (1 + tanh(-20 y)) on the 't' walls
    p2 = restart.pressure_to_mesh2(c.extra["P"], c.lx1, 2).reshape(s.eshape2)
    dt, ns, _ = prepare_linearized_solver(s, U, c.end_time)
    assert ns == 4
    res = {}
    for ra in (400.0, 500.0):
        st = ScalarStepper(s, U, c.re, T, tm, cond=1.0, rhocp=1.0, ri=5.0 * ra, gdir=1, solver="direct", ifvcor=True)
        u, _, t = st.map_scalar(U, p2, T, ns, dt, mode="nonlinear")
        res[ra] = st.inner_scalar((u - U, t - T), (u - U, t - T))
    assert res[400.0] < 2e-10, res                                              # measured 1.16e-10 (|U, T|^2 = 4.39)
    assert res[500.0] > 0.5, res                                                # 0.85: the buoyancy term is pinned, not just present


def test_direct_map_is_the_jacobian_of_the_full_map():
    from oracle.scalar import ScalarStepper
    c = small_cases()["box2d_n6_outflow"]
    s = make_oracle(c)
    U = c.ubase.reshape((2,) + s.eshape)
    tm = s.mask[0].copy()
    Tb = smooth_field(c, 11).reshape(U.shape)[0]
    kw = dict(cond=0.7 / c.re, rhocp=1.3, ri=0.4, gdir=1, solver="direct", ifvcor=c.ifvcor)
    st = ScalarStepper(s, U, c.re, Tb, tm, **kw)                                  # no sponge: the full equations carry none
    v = s.mask * smooth_field(c, 8).reshape(U.shape)
    th = tm * smooth_field(c, 12).reshape(U.shape)[1]
    p0 = np.zeros(s.eshape2)
    ns, dt, eps = 1, 2e-3, 1e-6                                                 # one step: the frozen base flow IS the trajectory the full map is differentiated about
    lin = st.map_scalar(v, p0, th, ns, dt)
    a = st.map_scalar(U, p0, Tb, ns, dt, mode="nonlinear")
    b = st.map_scalar(U + eps * v, p0, Tb + eps * th, ns, dt, mode="nonlinear")
    for k in (0, 2):
        fd = (b[k] - a[k]) / eps
        assert np.linalg.norm(fd - lin[k]) < 1e-4 * np.linalg.norm(lin[k]), k
    # the PCG path (what the GPU runs) agrees with the sparse-direct one
    st2 = ScalarStepper(s, U, c.re, Tb, tm, tol_v=1e-13, tol_p=1e-13, **{**kw, "solver": "pcg"})
    lin2 = st2.map_scalar(v, p0, th, ns, dt)
    assert np.linalg.norm(lin2[0] - lin[0]) < 1e-11 * np.linalg.norm(lin[0]) and np.linalg.norm(lin2[2] - lin[2]) < 1e-11 * np.linalg.norm(lin[2])
