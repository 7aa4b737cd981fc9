"""ctypes loaders for the oracle's native pieces -- TEST INFRASTRUCTURE ONLY.

``ref_voigt``      the reference's own ``voigt.c`` (compiled by ``oracle/Makefile`` into
                   ``oracle/_ref/voigt_ref.so``) with libcerf replaced by SciPy's Faddeeva ``wofz``.
``wofz_pointer``   address of ``scipy.special.cython_special``'s C-level ``wofz``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)


def wofz_pointer() -> int:
    import scipy.special.cython_special as cs
    cap = cs.__pyx_capi__["wofz"]
    get_name = ctypes.pythonapi.PyCapsule_GetName
    get_name.restype, get_name.argtypes = ctypes.c_char_p, [ctypes.py_object]
    get_ptr = ctypes.pythonapi.PyCapsule_GetPointer
    get_ptr.restype, get_ptr.argtypes = ctypes.c_void_p, [ctypes.py_object, ctypes.c_char_p]
    name = get_name(cap)
    assert b"double_complex (__pyx_t_double_complex, int" in name, name
    return get_ptr(cap, name)


_ref = None


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "voigt_ref.so"))


def _load_ref():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "voigt_ref.so"))
        lib.ref_set_wofz.argtypes = [ctypes.c_void_p]
        lib.ref_voigt.argtypes = [_dp, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp]
        lib.ref_voigt.restype = ctypes.c_int
        lib.ref_set_wofz(wofz_pointer())
        _ref = lib
    return _ref


def ref_voigt(lambdas, z, N, num_lines=31):
    """Run the reference's mexFunction (voigt.c:253-304) on plain arrays."""
    lib = _load_ref()
    lam = np.ascontiguousarray(lambdas, dtype=np.float64).ravel()
    out = np.empty(lam.size - 6)
    rc = lib.ref_voigt(lam.ctypes.data_as(_dp), lam.size, float(z), float(N), int(num_lines),
                       out.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("ref_voigt failed (rc=%d)" % rc)
    return out


# ------------------------------------------------------------------ C restatement (oracle/c)
_c = None
_u8p = ctypes.POINTER(ctypes.c_ubyte)
_ip = ctypes.POINTER(ctypes.c_int)


def have_c_oracle() -> bool:
    return os.path.exists(os.path.join(_HERE, "c", "libgpdla_oracle.so"))


def _load_c():
    global _c
    if _c is None:
        from . import process_qsos_oracle as O
        lib = ctypes.CDLL(os.path.join(_HERE, "c", "libgpdla_oracle.so"))
        lib.gpdla_oracle_set_wofz.argtypes = [ctypes.c_void_p]
        lib.gpdla_oracle_set_tables.argtypes = [_dp, _dp, _dp, _dp]
        lib.gpdla_oracle_voigt.argtypes = [_dp, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp]
        lib.gpdla_oracle_log_mvnpdf_low_rank.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_long, ctypes.c_int]
        lib.gpdla_oracle_log_mvnpdf_low_rank.restype = ctypes.c_double
        lib.gpdla_oracle_sample_loglik.argtypes = [
            _dp, ctypes.c_long, _u8p, _dp, _dp, _dp, _dp, _dp, ctypes.c_long, ctypes.c_int,
            _dp, _dp, ctypes.c_long, ctypes.c_int, _ip, ctypes.c_int, _dp, ctypes.c_int]
        lib.gpdla_oracle_set_wofz(wofz_pointer())
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_dp)
        lib.gpdla_oracle_set_tables(f(O.TRANSITION_WAVELENGTHS), f(O.LEADING_CONSTANTS), f(O.GAMMAS),
                                    f(O.INSTRUMENT_PROFILE))
        _c = lib
    return _c


def c_voigt(lambdas, z, N, num_lines=31):
    lib = _load_c()
    lam = np.ascontiguousarray(lambdas, dtype=np.float64).ravel()
    out = np.empty(lam.size - 6)
    if lib.gpdla_oracle_voigt(lam.ctypes.data_as(_dp), lam.size, float(z), float(N), int(num_lines),
                              out.ctypes.data_as(_dp)):
        raise RuntimeError("gpdla_oracle_voigt failed")
    return out


def c_log_mvnpdf_low_rank(y, mu, M, d):
    lib = _load_c()
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (y, mu, M, d)]
    n, k = a[2].shape
    return lib.gpdla_oracle_log_mvnpdf_low_rank(*[x.ctypes.data_as(_dp) for x in a], n, k)


def c_sample_loglik(padded, keep, y, mu, M, omega2, v, sample_z, nhi, num_lines, partners=None, nthreads=0):
    """process_qsos.m:185-199 for one quasar, threaded over samples (the parfor analogue)."""
    lib = _load_c()
    padded, y, mu, M, omega2, v, sample_z, nhi = [
        np.ascontiguousarray(x, dtype=np.float64) for x in (padded, y, mu, M, omega2, v, sample_z, nhi)]
    keep8 = np.ascontiguousarray(keep, dtype=np.uint8)
    n, k = M.shape
    S = sample_z.size
    sll = np.empty(S)
    if partners is None:
        pp, npart = None, 0
    else:
        partners = np.ascontiguousarray(partners, dtype=np.int32).reshape(-1, S)
        pp, npart = partners.ctypes.data_as(_ip), partners.shape[0]
    rc = lib.gpdla_oracle_sample_loglik(
        padded.ctypes.data_as(_dp), keep8.size, keep8.ctypes.data_as(_u8p), y.ctypes.data_as(_dp),
        mu.ctypes.data_as(_dp), M.ctypes.data_as(_dp), omega2.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
        n, k, sample_z.ctypes.data_as(_dp), nhi.ctypes.data_as(_dp), S, int(num_lines), pp, npart,
        sll.ctypes.data_as(_dp), int(nthreads))
    if rc:
        raise RuntimeError("gpdla_oracle_sample_loglik failed")
    return sll
