"""CPU checks of the rest-frame optical-depth table the fused kernels look tau / N up in (built on the host by
libgpdla.so, csrc/gpdla_rest_table.h; exposed for verification as ``gpdla_rest_table``): against the oracle's line sum
(voigt.c:282-290 through the Faddeeva function) and against mpmath."""
import ctypes

import numpy as np
import pytest

from gp_dla_detection_b200 import _lib
from oracle import process_qsos_oracle as O


def build(num_lines, pixel_spacing=1e-4):
    lib = _lib.load()
    ncell, deg = ctypes.c_int32(), ctypes.c_int32()
    h, lo = ctypes.c_double(), ctypes.c_double()
    assert lib.gpdla_rest_table(num_lines, pixel_spacing, None, ctypes.byref(ncell), ctypes.byref(deg), ctypes.byref(h),
                                ctypes.byref(lo)) == 0
    coef = np.empty((deg.value + 1, ncell.value))
    n2 = ctypes.c_int32()
    assert lib.gpdla_rest_table(num_lines, pixel_spacing, coef.ctypes.data_as(_lib.c_double_p), ctypes.byref(n2), None, None,
                                None) == 0
    assert n2.value == ncell.value
    return coef, h.value, lo.value


def tau_over_n(w, num_lines):
    """sum_j leading_constant_j voigt(c (w / lambda_j - 1), sigma, gamma_j) at rest wavelength w (Angstrom); the velocity
    through expm1 so that the float64 evaluation is good to ~1e-15 relative (voigt.c:287 itself cancels to ~1e-13)."""
    tot = np.zeros_like(w)
    for j in range(num_lines):
        v = O.C_CGS * np.expm1(np.log(w) - np.log(O.TRANSITION_WAVELENGTHS[j] * 1e8))
        tot += O.LEADING_CONSTANTS[j] * O.cerf_voigt(v, O.SIGMA, O.GAMMAS[j])
    return tot


def lookup(coef, h, lo, w, shift=0.0):
    """the device's arithmetic: u = ln(w / lo) / h, cell = round(u + shift) (the cell is chosen for the mean position
    of a group of samples, up to half a cell away from this sample's), Horner in s = u - cell, |s| <= 1"""
    u = np.log(w / lo) / h
    c = np.rint(u + shift)
    s = u - c
    c = c.astype(np.int64)
    t = coef[-1, c]
    for p in range(coef.shape[0] - 2, -1, -1):
        t = t * s + coef[p, c]
    return t


@pytest.mark.parametrize("num_lines", [3, 1, 31])
def test_table_reproduces_the_line_sum(num_lines):
    coef, h, lo = build(num_lines)
    assert abs(h - 1e-4 * np.log(10.0)) < 1e-18 and lo == 880.0
    rng = np.random.default_rng(num_lines)
    w = np.exp(rng.uniform(np.log(882.0), np.log(1715.0), 200000))
    t = lookup(coef, h, lo, w, rng.uniform(-0.5, 0.5, w.size))
    ok = ~np.isnan(t)
    exact = tau_over_n(w[ok], num_lines)
    rel = np.abs(t[ok] - exact) / exact
    assert rel.max() < 2e-12, rel.max()           # design: 4.5e-13 from the polynomials + the Faddeeva code's own 1e-13
    assert coef.shape[0] == 9                     # degree 8
    # the cells left to direct evaluation are exactly those within 16.5 cells of a line centre (and the two end cells)
    centres = np.log(O.TRANSITION_WAVELENGTHS[:num_lines] * 1e8 / lo) / h
    cells = np.arange(coef.shape[1])
    near = np.min(np.abs(cells[:, None] - centres[None, :]), axis=1) < 16.5
    near[[0, -1]] = True
    assert np.array_equal(np.isnan(coef[0]), near)
    assert np.array_equal(np.isnan(coef).any(axis=0), near)
    if num_lines == 3:
        assert ok.mean() > 0.96


def test_table_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    coef, h, lo = build(3)
    rng = np.random.default_rng(0)
    lam = O.TRANSITION_WAVELENGTHS[:3] * 1e8
    # points just outside the direct-evaluation zones (where the polynomials are worst) and far out in the wings
    w = np.concatenate([lam[j] * np.exp(sgn * h * rng.uniform(18.0, 20.0, 6)) for j in range(3) for sgn in (-1, 1)]
                       + [np.exp(rng.uniform(np.log(1250.0), np.log(1700.0), 10))])
    t = lookup(coef, h, lo, w, rng.uniform(-0.5, 0.5, w.size))
    assert not np.isnan(t).any()
    for wi, ti in zip(w, t):
        tot = mp.mpf(0)
        for j in range(3):
            v = mp.mpf(O.C_CGS) * mp.expm1(mp.log(mp.mpf(float(wi))) - mp.log(mp.mpf(float(lam[j]))))
            z = mp.mpc(v, O.GAMMAS[j]) / (mp.sqrt(2) * O.SIGMA)
            tot += mp.mpf(float(O.LEADING_CONSTANTS[j])) * (mp.exp(-z * z) * mp.erfc(-1j * z)).real / (mp.sqrt(2 * mp.pi) * O.SIGMA)
        assert abs(ti - float(tot)) / float(tot) < 1e-12, (wi, ti, float(tot))


def test_other_pixel_spacing_and_errors():
    coef, h, lo = build(3, 2e-4)
    assert abs(h - 2e-4 * np.log(10.0)) < 1e-18
    w = np.exp(np.random.default_rng(5).uniform(np.log(1300.0), np.log(1600.0), 1000))
    t = lookup(coef, h, lo, w)
    assert np.max(np.abs(t - tau_over_n(w, 3)) / tau_over_n(w, 3)) < 1e-11
    lib = _lib.load()
    n = ctypes.c_int32()
    assert lib.gpdla_rest_table(0, 1e-4, None, ctypes.byref(n), None, None, None) == _lib.GPDLA_ERR_INVALID
    assert lib.gpdla_rest_table(32, 1e-4, None, ctypes.byref(n), None, None, None) == _lib.GPDLA_ERR_INVALID
    assert lib.gpdla_rest_table(3, 0.0, None, ctypes.byref(n), None, None, None) == _lib.GPDLA_ERR_INVALID
