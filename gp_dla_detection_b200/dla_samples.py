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


def generate_lls_samples(num_dla_samples: int = 10000, catalog_log_nhis: Optional[Sequence[float]] = None,
                         log_pdf_poly: Optional[Sequence[float]] = None, alpha: float = 0.97,
                         min_lls_log_nhi: float = 19.5, uniform_min_log_nhi: float = 19.5,
                         uniform_max_log_nhi: float = 23.0, fit_min_log_nhi: float = 20.0,
                         fit_max_log_nhi: float = 22.0, extrapolate_min_log_nhi: float = 19.5,
                         peak_log_nhi: float = 20.03269) -> Dict[str, np.ndarray]:
    """``multi_dlas/set_lls_parameters.m:5-71``: the sub-DLA (Lyman-limit system) column-density samples and the
    partition functions ``Z_lls`` / ``Z_dla`` of the multi-DLA model (``gpdla_set_lls_samples``).

    Dimension 3 of the 3-D RR2-scrambled Halton sequence is mapped uniformly onto
    ``[min_lls_log_nhi, fit_min_log_nhi)`` (:55-63).  The column-density density is the quadratic fit of
    ``generate_dla_samples``, held constant at its value at ``peak_log_nhi`` below that point (:42-45), normalised
    on ``[extrapolate_min_log_nhi, 25]`` (:46) and mixed with a uniform on ``[uniform_min, uniform_max]`` (:50-52);
    ``Z_lls`` integrates the mixture over ``[min_lls_log_nhi, fit_min_log_nhi]``, ``Z_dla`` over
    ``[fit_min_log_nhi, uniform_max_log_nhi]`` (:68-71).  Parity status: unpinned, like ``generate_dla_samples``.
    """
    from scipy.integrate import quad
    seq = halton_rr2(num_dla_samples, 3)                                             # :15
    if log_pdf_poly is None:
        if catalog_log_nhis is None:
            raise ValueError("need catalog_log_nhis or log_pdf_poly")
        x = np.linspace(fit_min_log_nhi, fit_max_log_nhi, 1000)                      # :38
        log_pdf_poly = np.polyfit(x, np.log(ksdensity_normal(catalog_log_nhis, x)), 2)   # :39-40
    f = np.asarray(log_pdf_poly, dtype=np.float64)
    plateau = float(np.exp(np.polyval(f, peak_log_nhi)))

    def unnormalized_pdf(nhi):                                                       # :43-45 (heaviside(0) = 1/2)
        nhi = np.asarray(nhi, dtype=np.float64)
        h = np.where(nhi > peak_log_nhi, 1.0, np.where(nhi < peak_log_nhi, 0.0, 0.5))
        return np.exp(np.polyval(f, nhi)) * h + plateau * (1.0 - h)

    Z = quad(unnormalized_pdf, extrapolate_min_log_nhi, 25.0, points=[peak_log_nhi], limit=200)[0]   # :46
    width = uniform_max_log_nhi - uniform_min_log_nhi

    def normalized_pdf(nhi):                                                         # :50-52
        u = ((nhi >= uniform_min_log_nhi) & (nhi <= uniform_max_log_nhi)) / width
        return alpha * unnormalized_pdf(nhi) / Z + (1 - alpha) * u

    lls_offset_samples = seq[:, 2]                                                   # :55
    lls_log_nhi_samples = min_lls_log_nhi + (fit_min_log_nhi - min_lls_log_nhi) * lls_offset_samples   # :58-60
    Z_lls = quad(normalized_pdf, min_lls_log_nhi, fit_min_log_nhi, limit=200)[0]     # :70
    Z_dla = quad(normalized_pdf, fit_min_log_nhi, uniform_max_log_nhi, points=[peak_log_nhi], limit=200)[0]   # :71
    return dict(offset_samples=seq[:, 0], lls_offset_samples=lls_offset_samples,
                lls_log_nhi_samples=lls_log_nhi_samples, lls_nhi_samples=10.0 ** lls_log_nhi_samples,
                Z_lls=float(Z_lls), Z_dla=float(Z_dla), log_pdf_poly=f, alpha=alpha)
