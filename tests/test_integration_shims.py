"""The reference-side bindings INTEGRATION.md shows must at least compile and link against include/gpdla.h and
libgpdla.so: the MEX shim is extracted from the document, built against the oracle's stand-in mex.h and driven through a
small harness.  Without a GPU the call ends in the shim's own error path (libgpdla has no CPU fallback); on a B200 it
returns the profile of the reference's voigt.c."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_mex_shim(tmp_path):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```c\n(.*?)```", text, flags=re.S)
    shim = [b for b in blocks if "mexFunction" in b and "gpdla_voigt" in b]
    assert len(shim) == 1, "INTEGRATION.md must hold exactly one MEX shim for voigt"
    src = tmp_path / "voigt_gpdla.c"
    src.write_text(shim[0])
    so = tmp_path / "voigt_gpdla_mex.so"
    libdir = os.path.join(ROOT, "gp_dla_detection_b200")
    cmd = ["gcc", "-O1", "-shared", "-fPIC", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
           "-I" + os.path.join(ROOT, "include"), str(src), os.path.join(ROOT, "oracle", "ref_shim", "mex_harness.c"),
           "-L" + libdir, "-l:libgpdla.so", "-Wl,-rpath," + libdir, "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lib = ctypes.CDLL(str(so))
    lib.harness_voigt.restype = ctypes.c_int
    lib.harness_voigt.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_int]
    return lib


def call(lib, lam, z, N, nl):
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    out = np.zeros(max(lam.size - 6, 1))
    msg = ctypes.create_string_buffer(512)
    rc = lib.harness_voigt(lam.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), lam.size, z, N, nl,
                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), msg, 512)
    return rc, out[:max(lam.size - 6, 0)], msg.value.decode()


def test_mex_shim_compiles_links_and_validates(tmp_path):
    lib = build_mex_shim(tmp_path)
    rc, _, msg = call(lib, 10.0 ** (3.6 + 1e-4 * np.arange(5)), 2.2, 1e20, 3)      # fewer than 7 wavelengths: the shim's check
    assert rc == 1 and "at least 7 wavelengths" in msg
    import torch
    if not torch.cuda.is_available():
        rc, _, msg = call(lib, 10.0 ** (3.6 + 1e-4 * np.arange(50)), 2.2, 1e20, 3)
        assert rc == 1 and "gpdla:voigt" in msg                                     # GPDLA_ERR_CUDA surfaces as a MEX error


@pytest.mark.gpu
def test_mex_shim_returns_the_reference_profile(tmp_path):
    from oracle import process_qsos_oracle as O
    lib = build_mex_shim(tmp_path)
    lam = 10.0 ** (3.5563 + 1e-4 * np.arange(400))
    for z, N, nl in ((2.05, 1e21, 3), (2.1, 10 ** 20.3, 0)):                        # 0: three-argument call, 31 lines
        rc, prof, msg = call(lib, lam, z, N, nl)
        assert rc == 0, msg
        assert np.max(np.abs(prof - O.voigt(lam, z, N, nl if nl else 31))) <= 5e-15
