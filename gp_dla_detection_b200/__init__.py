"""gp_dla_detection_b200 -- B200-native per-quasar DLA model-selection hot path.

Drop-in for the reference's ``process_qsos`` script (process_qsos.m) and its native
``voigt(lambdas, z, N, num_lines)`` MEX function (voigt.c:253-304); everything runs in
hand-written sm_100a CUDA kernels behind the C ABI declared in ``include/gpdla.h``.
There is no CPU fallback: importing :mod:`gp_dla_detection_b200.api` without the built
``libgpdla.so`` raises.
"""
from .params import Parameters, DEFAULT as DEFAULT_PARAMETERS  # noqa: F401

__all__ = ["Parameters", "DEFAULT_PARAMETERS"]
