"""Planning data for the next kernel step: can the three-line optical depth per unit column density,
tau/N = T(w) with w = lambda_obs / (1 + z_dla) the absorber rest-frame wavelength, come from a table in log10(w)
(cells of 1e-4 dex / R, a degree-p polynomial per cell) instead of 45 FP64 instructions per (sample, pixel)?
For each (R, p) this prints the distance from a line centre (in 1e-4-dex pixels) beyond which the per-cell Chebyshev
interpolant reproduces T to 1e-13 relative, and the table size.  CPU only (mpmath-free: T is evaluated with the oracle's
Faddeeva-based voigt in float64, whose own error, ~1e-14, bounds what can be resolved).
  python tools/rest_table_study.py > profiles/r01_rest_table_study.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from numpy.polynomial import chebyshev as C
from oracle import process_qsos_oracle as O

NL = 3
lam0 = O.TRANSITION_WAVELENGTHS[:NL] * 1e8          # Angstrom


def T(w):
    """tau / N at rest wavelength w (Angstrom): sum_j lc_j voigt(c (w / lambda_j - 1), sigma, gamma_j)"""
    tot = np.zeros_like(w)
    for j in range(NL):
        v = O.C_CGS * (w / lam0[j] - 1.0)
        tot += O.LEADING_CONSTANTS[j] * O.cerf_voigt(v, O.SIGMA, O.GAMMAS[j])
    return tot


lo, hi = np.log10(880.0), np.log10(1225.0)           # rest range a DLA at min..max z can put under the spectrum
centres = np.log10(lam0)
print("# rest-frame table of tau/N for %d Lyman lines, log10(w) in [%.4f, %.4f]" % (NL, lo, hi))
print("# R = cells per 1e-4-dex pixel, p = polynomial degree per cell; D = pixels from the nearest line centre beyond which")
print("# the interpolant is within 1e-13 (1e-12) relative of T; KB = table size for the whole range")
print("R  p   D(1e-13)  D(1e-12)   KB")
for R in (1, 2, 4):
    h = 1e-4 / R
    ncell = int(np.ceil((hi - lo) / h))
    edges = lo + h * np.arange(ncell + 1)
    mid = 0.5 * (edges[:-1] + edges[1:])
    dist_pix = np.min(np.abs(mid[:, None] - centres[None, :]), axis=1) / 1e-4
    for p in (3, 4, 5, 6, 7):
        nodes = np.cos(np.pi * (np.arange(p + 1) + 0.5) / (p + 1))          # Chebyshev nodes on [-1, 1]
        test = np.linspace(-1, 1, 33)
        x_nodes = mid[:, None] + 0.5 * h * nodes[None, :]
        f_nodes = T(10.0 ** x_nodes)
        # interpolate each cell (vectorised: same Vandermonde for all cells)
        V = C.chebvander(nodes, p)
        coef = np.linalg.solve(V, f_nodes.T)                                  # [p+1, ncell]
        approx = C.chebvander(test, p) @ coef                                 # [33, ncell]
        exact = T(10.0 ** (mid[None, :] + 0.5 * h * test[:, None]))
        err = np.max(np.abs(approx - exact) / exact, axis=0)                  # per cell
        def dmin(tol):
            bad = dist_pix[err > tol]
            return float(bad.max()) if bad.size else 0.0
        print("%d  %d  %8.1f  %8.1f  %6.0f" % (R, p, dmin(1e-13), dmin(1e-12), ncell * (p + 1) * 8 / 1024))
