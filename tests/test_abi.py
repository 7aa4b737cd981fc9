"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gpdla.h declares, and fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from gp_dla_detection_b200 import _lib, api
from oracle import process_qsos_oracle as O


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpdla.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpdla_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SYMBOLS, "binding missing for " + n
    assert set(_lib.SYMBOLS) == set(names)


def test_line_constants_bit_equal_to_oracle_tables():
    lib = _lib.load()
    tw, lc, gam, ip = np.zeros(31), np.zeros(31), np.zeros(31), np.zeros(7)
    lib.gpdla_line_constants(api._dp(tw), api._dp(lc), api._dp(gam), api._dp(ip))
    assert np.array_equal(tw, O.TRANSITION_WAVELENGTHS)
    assert np.array_equal(lc, O.LEADING_CONSTANTS)
    assert np.array_equal(gam, O.GAMMAS)
    assert np.array_equal(ip, O.INSTRUMENT_PROFILE)


def test_default_parameters_match_set_parameters():
    lib = _lib.load()
    p = _lib.GpdlaParams()
    lib.gpdla_default_parameters(ctypes.byref(p))
    assert (p.min_lambda, p.max_lambda, p.num_lines, p.pixel_spacing) == (911.75, 1215.75, 3, 1e-4)
    assert p.prior_z_qso_increase == O.prior_z_qso_increase and p.max_z_cut == O.max_z_cut
    assert p.lya_wavelength == O.lya_wavelength and p.lyman_limit == O.lyman_limit


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GpdlaError):
        api.voigt(np.linspace(3600, 4000, 50), 2.0, 1e20, 3)
    ctx = ctypes.c_void_p()
    assert _lib.load().gpdla_create(ctypes.byref(ctx), 0) == _lib.GPDLA_ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gp_dla_detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libgpdla_oracle" not in src and "voigt_ref" not in src, f


def test_ascii_catalog_writer_formats(tmp_path):
    """generate_ascii_catalog.m:9-20,49-81: same printf formats, three-digit exponents, THING_ID only."""
    from gp_dla_detection_b200 import catalog_writer as cw
    res = dict(min_z_dlas=np.array([2.01234, 1.9]), max_z_dlas=np.array([2.98768, 3.1]),
               log_priors_no_dla=np.array([-0.10382, -0.2]), log_priors_dla=np.array([-2.31655, -1.7]),
               log_likelihoods_no_dla=np.array([-190.48437, -1.5e3]), log_likelihoods_dla=np.array([-221.16647, -1.4e3]),
               model_posteriors=np.array([[1.0, 5.107e-15], [3.2e-101, 1.0]]), p_dlas=np.array([5.107e-15, 1.0]),
               map_z_dlas=np.array([2.5, 2.25]), map_log_nhis=np.array([20.0092, 21.5]))
    p = tmp_path / "results.dat"
    cw.write_results(str(p), res, [12345, 987654321])
    lines = p.read_text().splitlines()
    assert lines[0] == "000012345 2.0123 2.9877 -0.10382 -2.31655 -1.90484e+02 -2.21166e+02 1.00000e+000 5.10700e-015 2.5000 20.0092"
    assert lines[1] == "987654321 1.9000 3.1000 -0.20000 -1.70000 -1.50000e+03 -1.40000e+03 3.20000e-101 1.00000e+000 2.2500 21.5000"
    s = tmp_path / "samples.dat"
    cw.write_dla_samples(str(s), [0.5, 0.25], [20.2966751, 21.0])
    assert s.read_text() == "0.500000 20.296675\n0.250000 21.000000\n"
