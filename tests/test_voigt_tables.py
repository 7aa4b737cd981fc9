"""CPU pin of the device's Voigt tables (gp_dla_detection_b200/csrc/voigt_tables.h): the header is parsed and the
evaluation order of csrc/gpdla_math.cuh (tau_core: H1 as even + odd halves in t^2, H3 by Horner, Cody-Waite exp with the
degree-11 polynomial; tau_wing: A(u), B(u)) is replayed in float64 NumPy against the Faddeeva function (the code behind
libcerf's voigt(), voigt.c:288) for the three Lyman-series damping parameters the hot path uses."""
import os
import re

import numpy as np
from scipy.special import wofz

from oracle import process_qsos_oracle as O

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gp_dla_detection_b200", "csrc",
                   "voigt_tables.h")


def _tables():
    txt = open(HDR).read().replace("\\\n", " ")
    out = {}
    for m in re.finditer(r"#define\s+(GPDLA_\w+)\s+(.*)", txt):
        name, val = m.group(1), m.group(2).strip()
        if val.startswith("{"):
            out[name] = np.array([float(v) for v in val.strip("{} ").split(",") if v.strip()])
        else:
            out[name] = float(val)
    return out


def _exp_nonpos(x, poly):
    kd = np.rint(x * 1.4426950408889634074)
    r = kd * -6.93147180369123816490e-01 + x
    r = kd * -1.90821492927058770002e-10 + r
    p = np.full_like(x, poly[11])
    for i in range(10, -1, -1):
        p = p * r + poly[i]
    return np.ldexp(p, kd.astype(np.int64))


def _rew_core(x, y, T):
    d1, d3, stride = int(T["GPDLA_VOIGT_DEG_H1"]), int(T["GPDLA_VOIGT_DEG_H3"]), int(T["GPDLA_VOIGT_CORE_STRIDE"])
    tab = T["GPDLA_VOIGT_CORE_TABLE"].reshape(-1, stride)
    ax = np.abs(x)
    idx = np.minimum((ax * T["GPDLA_VOIGT_INV_H"]).astype(np.int64), int(T["GPDLA_VOIGT_NINT"]) - 1)
    t = ax * (2.0 * T["GPDLA_VOIGT_INV_H"]) - (2.0 * idx + 1.0)
    t2 = t * t
    c = tab[idx]
    h1e, h1o = c[:, d1 - 1].copy(), c[:, d1].copy()
    for i in range(d1 - 3, -1, -2):
        h1e = h1e * t2 + c[:, i]
        h1o = h1o * t2 + c[:, i + 1]
    h1 = h1o * t + h1e
    c3 = c[:, d1 + 1:]
    h3 = c3[:, d3].copy()
    for i in range(d3 - 1, -1, -1):
        h3 = h3 * t + c3[:, i]
    x2 = x * x
    e = _exp_nonpos(-x2, T["GPDLA_EXP_POLY"])
    y2 = y * y
    p4 = (x2 * (x2 * 4.0 - 12.0) + 3.0) * (1.0 / 6.0)
    even = y2 * (y2 * p4 + (-2.0 * x2 + 1.0)) + 1.0
    return e * even + y * (y2 * h3 + h1)


def _rew_wing(x, y, T):
    u = 1.0 / (x * x)
    a = np.polyval(T["GPDLA_VOIGT_WING_A"][::-1], u)
    b = np.polyval(T["GPDLA_VOIGT_WING_B"][::-1], u)
    return y / np.sqrt(np.pi) * u * (a - y * y * u * b)


def test_core_and_wing_tables_reproduce_the_faddeeva_function():
    T = _tables()
    assert T["GPDLA_VOIGT_CORE_TABLE"].size == int(T["GPDLA_VOIGT_NINT"]) * int(T["GPDLA_VOIGT_CORE_STRIDE"])
    assert int(T["GPDLA_VOIGT_NINT"]) / T["GPDLA_VOIGT_INV_H"] == T["GPDLA_VOIGT_X0"]
    sigma = O.SIGMA if hasattr(O, "SIGMA") else O.sigma
    gammas = np.asarray(O.GAMMAS if hasattr(O, "GAMMAS") else O.gammas)[:3]
    rng = np.random.default_rng(5)
    xc = np.concatenate([rng.uniform(0, T["GPDLA_VOIGT_X0"], 20000), np.arange(64) / 4.0, np.arange(1, 65) / 4.0 - 1e-12])
    xw = np.concatenate([T["GPDLA_VOIGT_X0"] * 10 ** rng.uniform(0, 3, 20000), [T["GPDLA_VOIGT_X0"]]])
    for g in gammas:
        y = g / (np.sqrt(2.0) * sigma)
        ref = wofz(xc + 1j * y).real
        got = _rew_core(xc, y, T)
        # wofz itself is good to ~1e-14 here; the tables were fitted to 60-digit mpmath values
        assert np.max(np.abs(got - ref) / ref) < 3e-14
        assert np.array_equal(_rew_core(-xc, y, T), got)
        refw = wofz(xw + 1j * y).real
        assert np.max(np.abs(_rew_wing(xw, y, T) - refw) / refw) < 3e-14
