"""CPU oracle for the gp_dla_detection hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / CPU baseline.  The product
(``gp_dla_detection_b200``) never imports this package and has no CPU fallback.

Parity status: **parity unpinned** against the reference's own fixtures -- the
reference ships no tests, golden vectors or fixtures, and neither
MATLAB/Octave nor libcerf exist in this image, so ``process_qsos.m`` cannot be
run here.  What *is* pinned:

* ``oracle/_ref/voigt_ref.so`` compiles the reference's own ``voigt.c`` (from
  ``/root/reference`` where it lies) against a tiny ``mex.h``/``cerf.h`` shim
  whose ``voigt(x, sigma, gamma)`` calls the Faddeeva-package ``wofz`` inside
  SciPy (the same S. G. Johnson code libcerf wraps); the Python restatement in
  ``oracle/process_qsos_oracle.py`` is checked bit-for-bit against it.
* The Faddeeva function is refereed against mpmath at 40 digits.
* ``log_mvnpdf_low_rank`` is checked against the dense multivariate normal.
"""
