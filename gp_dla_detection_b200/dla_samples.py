"""DLA parameter samples -- the producer of the hot path's `offset_samples` / `log_nhi_samples` /
`nhi_samples` inputs (``generate_dla_samples.m``, ``multi_dlas/set_lls_parameters.m``; SURVEY.md section
8(f) rank 1).  Host-side, NumPy/SciPy only.

The reference draws a 2-D (3-D for the sub-DLA model) Halton sequence scrambled with MATLAB's ``'rr2'``
(reverse-radix) permutation, uses dimension 1 as the uniform redshift offset and pushes dimension 2 through the
inverse CDF of a mixture: ``alpha`` x (exponential of a quadratic fitted to the KDE of the catalogue's
log10 N_HI values on [20, 22], normalised on [20, 25]) + ``(1 - alpha)`` x uniform on [20, 23].

Parity status: **unpinned** -- MATLAB's Statistics Toolbox (``haltonset``, ``scramble``, ``ksdensity``) is not
in the reference tree and cannot run here, so this restates their documented algorithms: Halton radical inverse
starting at index 0, RR2 digit permutation (bit-reversed order of 0..2^m-1 with out-of-range values dropped),
normal-kernel KDE with the robust normal-reference bandwidth.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

_PRIMES = (2, 3, 5, 7, 11, 13)


def rr2_permutation(base: int) -> np.ndarray:
    """Kocis-Whiten reverse-radix permutation of the digits 0..base-1 (MATLAB ``scramble(p, 'rr2')``)."""
    m = max(int(np.ceil(np.log2(base))), 1)
    rev = [int(format(i, "0%db" % m)[::-1], 2) for i in range(1 << m)]
    return np.array([r for r in rev if r < base], dtype=np.int64)


def halton_rr2(num_points: int, dims: int, skip: int = 0) -> np.ndarray:
    """``net(scramble(haltonset(dims), 'rr2'), num_points)``: rows = points (index ``skip`` .. ), columns = dims."""
    idx = np.arange(skip, skip + num_points, dtype=np.int64)
    out = np.zeros((num_points, dims))
    for d in range(dims):
        b = _PRIMES[d]
        perm = rr2_permutation(b)
        i = idx.copy()
        f = 1.0 / b
        while np.any(i > 0):
            out[:, d] += f * perm[i % b]
            i //= b
            f /= b
    return out


def ksdensity_normal(data: Sequence[float], x: np.ndarray) -> np.ndarray:
    """MATLAB ``ksdensity(data, x)`` defaults: normal kernel, bandwidth sigma (4 / (3 n))^(1/5) with the robust
    sigma = median(|data - median|) / 0.6745."""
    data = np.asarray(data, dtype=np.float64)
    n = data.size
    sig = np.median(np.abs(data - np.median(data))) / 0.6745
    if sig <= 0:
        sig = max(data.max() - data.min(), 1e-12)
    h = sig * (4.0 / (3.0 * n)) ** 0.2
    z = (x[:, None] - data[None, :]) / h
    return np.exp(-0.5 * z * z).sum(axis=1) / (n * h * np.sqrt(2 * np.pi))


def generate_dla_samples(num_dla_samples: int = 10000, catalog_log_nhis: Optional[Sequence[float]] = None,
                         log_pdf_poly: Optional[Sequence[float]] = None, alpha: float = 0.9,
                         uniform_min_log_nhi: float = 20.0, uniform_max_log_nhi: float = 23.0,
                         fit_min_log_nhi: float = 20.0, fit_max_log_nhi: float = 22.0) -> Dict[str, np.ndarray]:
    """``generate_dla_samples.m:8-57``.  Give either the catalogue's observed ``log10 N_HI`` values (the KDE +
    quadratic fit of :30-38 is then done here) or the fitted quadratic's coefficients ``log_pdf_poly`` (highest
    power first, as ``polyfit`` returns them)."""
    from scipy.integrate import quad
    from scipy.optimize import brentq
    seq = halton_rr2(num_dla_samples, 2)
    offset_samples = seq[:, 0]                                                      # :13
    if log_pdf_poly is None:
        if catalog_log_nhis is None:
            raise ValueError("need catalog_log_nhis or log_pdf_poly")
        x = np.linspace(fit_min_log_nhi, fit_max_log_nhi, 1000)                      # :32
        log_pdf_poly = np.polyfit(x, np.log(ksdensity_normal(catalog_log_nhis, x)), 2)   # :33-34
    f = np.asarray(log_pdf_poly, dtype=np.float64)
    unnormalized_pdf = lambda nhi: np.exp(np.polyval(f, nhi))                        # :37
    Z = quad(unnormalized_pdf, fit_min_log_nhi, 25.0)[0]                             # :38
    width = uniform_max_log_nhi - uniform_min_log_nhi

    def normalized_pdf(nhi):                                                         # :42-44
        u = ((nhi >= uniform_min_log_nhi) & (nhi <= uniform_max_log_nhi)) / width
        return alpha * unnormalized_pdf(nhi) / Z + (1 - alpha) * u

    # cdf on a fine grid (trapezoid refined by quad at the bracket ends would be overkill: 1e-9 is plenty)
    grid = np.linspace(fit_min_log_nhi, 25.0, 500001)
    pdf = normalized_pdf(grid)
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(grid))])
    log_nhi_samples = np.empty(num_dla_samples)
    for i, u in enumerate(seq[:, 1]):                                                # :50-54 (fzero from 20.5)
        j = min(max(int(np.searchsorted(cdf, u)), 1), grid.size - 1)
        lo, hi = grid[j - 1], grid[j]
        g = lambda t: cdf[j - 1] + quad(normalized_pdf, lo, t)[0] - u
        log_nhi_samples[i] = brentq(g, lo, hi) if g(lo) * g(hi) < 0 else (lo if abs(g(lo)) < abs(g(hi)) else hi)
    return dict(offset_samples=offset_samples, log_nhi_samples=log_nhi_samples,
                nhi_samples=10.0 ** log_nhi_samples, log_pdf_poly=f, alpha=alpha)
