"""KAT-part and numbering KATs: Nek5000's element->rank rule and the GLL node equivalence
classes, checked against the element maps of the reference's shipped field files (tests/golden/*.npz)."""
import os

import numpy as np
import pytest

from nekstab_b200 import cases
from util import GOLD


@pytest.fixture(scope="module")
def gold():
    return {k: np.load(os.path.join(GOLD, f"{k}.npz")) for k in ("cyl", "bfs")}


@pytest.mark.parametrize("name", ["cyl", "bfs"])
@pytest.mark.parametrize("P", [4, 6])
def test_partition_bit_exact(gold, name, P):
    g = gold[name]
    r = cases.partition(g["key"], P, int(g["d2"]))
    assert np.array_equal(r, g[f"rank_p{P}"])          # includes the unstable heap-sort tie order for P=6


def test_partition_balance_and_power_of_two(gold):
    g = gold["cyl"]
    for P, expect in ((2, [998, 998]), (4, [499] * 4), (8, None)):
        r = cases.partition(g["key"], P, int(g["d2"]))
        cnt = np.bincount(r, minlength=P)
        if expect:
            assert cnt.tolist() == expect
        else:
            assert set(cnt.tolist()) <= {249, 250}
        assert np.array_equal(r, g["key"] // (int(g["d2"]) // P))
    assert np.all(cases.partition(g["key"], 1) == 0)


def test_numbering_counts(gold):
    c = cases.cylinder_case(gold["cyl"])
    assert c.n == 71856 and int(c.glo.max()) + 1 == 50089                    # 21 767 redundant copies
    b = cases.bfs_case(gold["bfs"])
    assert b.n == 60120 and int(b.glo.max()) + 1 == 42341
    # coincident nodes share coordinates (periodic images differ by the period in y only)
    for case in (c, b):
        g = case.glo.ravel()
        for d in range(2):
            x = case.xyz[d].ravel()
            mx = np.full(g.max() + 1, -np.inf); mn = np.full(g.max() + 1, np.inf)
            np.maximum.at(mx, g, x); np.minimum.at(mn, g, x)
            spread = mx - mn
            ok = (spread < 1e-6) | (np.abs(spread - 32.0) < 1e-6)
            assert ok.all()


def test_extrusion_numbering():
    c2 = cases.box_case(3, 2, 4)
    c3 = cases.extrude(c2, 3, 1.0)
    n2 = int(c2.glo.max()) + 1
    assert int(c3.glo.max()) + 1 == n2 * 3 * 3            # periodic: nz*(lx1-1) levels
    assert c3.nel == 18 and c3.xyz.shape == (3, 18, 64)
    z = c3.xyz[2].reshape(18, 4, 16)
    assert np.allclose(z[:, 0, :].min(), 0.0) and np.allclose(z[-1, -1, :], 1.0)


def test_sponge_and_mask_kat(gold):
    c = cases.cylinder_case(gold["cyl"])
    nz = c.spng_fun != 0
    x = c.xyz[0]
    assert np.all((x[nz] < -12.66) | (x[nz] > 46.66))      # nonzero for x<-12.67, x>46.67
    assert (c.mask[0] == 0).sum() == 276 and (c.extra["mask_adjoint"][0] == 0).sum() == 456
