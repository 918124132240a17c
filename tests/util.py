"""Shared helpers for the parity tests: small deterministic cases and oracle construction."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from nekstab_b200 import cases  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def small_cases():
    """name -> Case; covers 2-D/3-D, lx1 in {4,6,8}, outflow vs all-Dirichlet (singular E), periodic, deformed."""
    out = {}
    out["box2d_n6_outflow"] = cases.box_case(4, 3, 6, outflow=True, deform=0.08)
    out["box2d_n8_dirichlet"] = cases.box_case(3, 3, 8, outflow=False, deform=0.0, shear=0.3)
    out["box2d_n4_periodic"] = cases.box_case(4, 4, 4, outflow=True, periodic_y=True, deform=0.06)
    c2 = cases.box_case(3, 2, 8, outflow=True, deform=0.07)
    out["box3d_n8_outflow"] = cases.extrude(c2, 3, 1.5)
    c2 = cases.box_case(3, 3, 6, outflow=False, deform=0.0, shear=0.25)
    out["box3d_n6_dirichlet"] = cases.extrude(c2, 3, 1.0)
    c2 = cases.box_case(2, 2, 4, outflow=True, deform=0.05)
    out["box3d_n4_outflow"] = cases.extrude(c2, 4, 1.0)
    for c in out.values():
        add_sponge_and_3d_flow(c)
    return out


def add_sponge_and_3d_flow(c):
    x = c.xyz[0]
    L = x.max() - x.min()
    c.spng_fun = cases.sponge_function([c.xyz[d] for d in range(c.ldim)], [0.0] * c.ldim, [0.25 * L] + [0.0] * (c.ldim - 1))
    if c.ldim == 3:
        z = c.xyz[2]
        lz = z.max() - z.min()
        c.ubase = c.ubase.copy()
        c.ubase[2] = 0.05 * np.sin(2 * np.pi * z / lz) * np.sin(np.pi * c.xyz[1] / (c.xyz[1].max() - c.xyz[1].min() + 1e-30))
        c.ubase[0] = c.ubase[0] * (1.0 + 0.1 * np.cos(2 * np.pi * z / lz))


def smooth_field(c, seed=0, masked=True):
    """Deterministic smooth, C0-continuous, masked velocity field (ldim, nel, npts)."""
    rng = np.random.default_rng(seed)
    out = []
    for d in range(c.ldim):
        a = rng.standard_normal(6)
        f = 0.0
        for k in range(c.ldim):
            xk = c.xyz[k]
            s = (xk - xk.min()) / (xk.max() - xk.min() + 1e-30)
            f = f + a[k] * np.sin(2 * np.pi * s + a[k + 3]) + 0.3 * a[k] * s * s
        out.append(f * (c.mask[d] if masked else 1.0))
    return np.stack(out)


def random_nodal(c, seed=0, masked=True):
    """i.i.d. N(0,1) per unique global node scattered to elements, masked."""
    rng = np.random.default_rng(seed)
    ng = int(c.glo.max()) + 1
    out = []
    for d in range(c.ldim):
        g = rng.standard_normal(ng)
        out.append(g[c.glo] * (c.mask[d] if masked else 1.0))
    return np.stack(out)


def make_oracle(c):
    from oracle.ops import SEM
    return SEM(c.ldim, c.lx1, c.xyz, c.glo, c.mask)


def rel(a, b):
    a = np.asarray(a, float).ravel(); b = np.asarray(b, float).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
