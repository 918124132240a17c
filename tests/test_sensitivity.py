"""CPU tests of the host-side wavemaker post-processing (nekstab_b200/sensitivity.py; core/sensitivity.f:7-81, 428-504) on the
shipped cylinder direct mode (tests/golden/cyl.npz) and a synthetic adjoint mode."""
import os

import numpy as np

from nekstab_b200 import cases, sensitivity
from util import GOLD, make_oracle


def _modes():
    g = np.load(os.path.join(GOLD, "cyl.npz"))
    c = cases.cylinder_case(g)
    s = make_oracle(c)
    bm1s = (s.bm1 * (c.spng_fun.reshape(s.eshape) == 0)).reshape(c.nel, -1)
    dre = g["dRe_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
    dim = g["dIm_U"].astype(float).transpose(1, 0, 2, 3).reshape(2, c.nel, -1)
    return c, bm1s, dre, dim


def test_biorthogonalize_normalises_direct_and_pairs_adjoint():
    c, bm1s, dre, dim = _modes()
    rng = np.random.default_rng(0)
    # synthetic adjoint: the direct mode rotated by a complex factor plus a smooth perturbation, arbitrary scale
    z = 3.7 * np.exp(0.9j)
    a = z * (dre + 1j * dim) + 0.2 * rng.standard_normal(dre.shape) * np.abs(dre).max()
    d_re, d_im, a_re, a_im = sensitivity.biorthogonalize(2.5 * dre, 2.5 * dim, a.real, a.imag, bm1s)
    ip = sensitivity.inner_product
    assert abs(ip(d_re, d_re, bm1s) + ip(d_im, d_im, bm1s) - 1.0) < 1e-12
    re = ip(a_re, d_re, bm1s) + ip(a_im, d_im, bm1s)           # <a, d> = sum conj(a) d
    im = ip(a_re, d_im, bm1s) - ip(a_im, d_re, bm1s)
    assert abs(re - 1.0) < 1e-12 and abs(im) < 1e-12


def test_wave_maker_is_pointwise_product_and_scale_invariant():
    c, bm1s, dre, dim = _modes()
    rng = np.random.default_rng(1)
    a = (dre + 1j * dim) * np.exp(0.3j) + 0.1 * rng.standard_normal(dre.shape) * np.abs(dre).max()
    wm = sensitivity.wave_maker(dre, dim, a.real, a.imag, bm1s)
    assert wm.shape == (c.nel, c.lx1 ** 2) and wm.min() >= 0
    d_re, d_im, a_re, a_im = sensitivity.biorthogonalize(dre, dim, a.real, a.imag, bm1s)
    ref = np.sqrt((d_re ** 2 + d_im ** 2).sum(0)) * np.sqrt((a_re ** 2 + a_im ** 2).sum(0))
    assert np.array_equal(wm, ref)
    # the result does not depend on the scale or phase of either input mode
    f, h = 7.0 * np.exp(1.1j), 0.03 * np.exp(-2.0j)
    d2, a2 = f * (dre + 1j * dim), h * a
    wm2 = sensitivity.wave_maker(d2.real, d2.imag, a2.real, a2.imag, bm1s)
    assert np.abs(wm2 - wm).max() < 1e-10 * wm.max()
