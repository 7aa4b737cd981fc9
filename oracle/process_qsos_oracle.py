"""Literal CPU restatement of the reference's single-DLA hot path (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  parity unpinned (the
reference has no fixtures); anchored instead on the compiled ``voigt.c``
(``oracle/_ref``), mpmath, and the dense multivariate normal.

Follows, operation by operation and in the same order:

* ``voigt.c:22-251`` (constant tables, recomputed from the commented formulas and
  asserted equal to the printed literals in ``tests/``) and ``voigt.c:277-299``
  (multipliers, raw profile, 7-tap instrument convolution);
* ``log_mvnpdf_low_rank.m:5-34`` (Woodbury with upper Cholesky, ``L\\(L'\\...)``);
* ``process_qsos.m:88-233`` (per-quasar loop, priors, interpolation, padding,
  per-sample loop, log-sum-exp, model posteriors);
* ``set_parameters.m:5-73`` (constants, ``min_z_dla``/``max_z_dla``);
* ``generate_ascii_catalog.m:73-80`` (MAP sample = first nan-ignoring arg-max).

libcerf (``#include <cerf.h>``, ``voigt.c:5,288``; unpinned, absent from the
reference tree) is replaced by ``scipy.special.wofz`` -- the same Faddeeva
package libcerf wraps and the substitution the reference itself makes in
``CDDF_analysis/voigt.py:221-228``:
``voigt(x, sigma, gamma) = Re w((x + i gamma)/(sqrt2 sigma)) / (sqrt(2 pi) sigma)``.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import wofz

# ---------------------------------------------------------------- set_parameters.m
lya_wavelength = 1215.6701          # set_parameters.m:5
lyb_wavelength = 1025.7223          # :6
lyman_limit = 911.7633              # :7
speed_of_light = 299792458.0        # :8


def kms_to_z(kms):                  # set_parameters.m:11
    return (kms * 1000.0) / speed_of_light


min_lambda = 911.75                 # :33
max_lambda = 1215.75                # :34
dlambda = 0.25                      # :35
k_default = 20                      # :36
num_dla_samples_default = 10000     # :48
prior_z_qso_increase = kms_to_z(30000.0)   # :56
width = 3                           # :59
pixel_spacing = 1e-4                # :60
num_lines_default = 3               # :63
max_z_cut = kms_to_z(3000.0)        # :65
min_z_cut = kms_to_z(3000.0)        # :69


def max_z_dla(wavelengths, z_qso):  # set_parameters.m:66-67
    return (np.max(wavelengths) / lya_wavelength - 1.0) - max_z_cut


def min_z_dla(wavelengths, z_qso):  # set_parameters.m:70-73
    return max(np.min(wavelengths) / lya_wavelength - 1.0,
               lyman_limit * (1.0 + z_qso) / lya_wavelength - 1.0 + min_z_cut)


# ---------------------------------------------------------------- voigt.c tables
C_CGS = 2.99792458e10               # voigt.c:22
SIGMA = 9.08537121627923800e5       # voigt.c:41
_E_CGS = 4.803204672997660e-10      # voigt.c:27
_M_E = 9.10938356e-28               # voigt.c:25

TRANSITION_WAVELENGTHS = np.array([  # voigt.c:31-64 (cm)
    1.2156701e-05, 1.0257223e-05, 9.725368e-06, 9.497431e-06, 9.378035e-06, 9.307483e-06,
    9.262257e-06, 9.231504e-06, 9.209631e-06, 9.193514e-06, 9.181294e-06, 9.171806e-06,
    9.16429e-06, 9.15824e-06, 9.15329e-06, 9.14919e-06, 9.14576e-06, 9.14286e-06,
    9.14039e-06, 9.13826e-06, 9.13641e-06, 9.13480e-06, 9.13339e-06, 9.13215e-06,
    9.13104e-06, 9.13006e-06, 9.12918e-06, 9.12839e-06, 9.12768e-06, 9.12703e-06,
    9.12645e-06])
OSCILLATOR_STRENGTHS = np.array([    # voigt.c:66-99
    0.416400, 0.079120, 0.029000, 0.013940, 0.007799, 0.004814, 0.003183, 0.002216,
    0.001605, 0.00120, 0.000921, 0.0007226, 0.000577, 0.000469, 0.000386, 0.000321,
    0.000270, 0.000230, 0.000197, 0.000170, 0.000148, 0.000129, 0.000114, 0.000101,
    0.000089, 0.000080, 0.000071, 0.000064, 0.000058, 0.000053, 0.000048])
GAMMAS_RATE = np.array([             # voigt.c:101-134 (s^-1)
    6.265e+08, 1.897e+08, 8.127e+07, 4.204e+07, 2.450e+07, 1.236e+07, 8.255e+06, 5.785e+06,
    4.210e+06, 3.160e+06, 2.432e+06, 1.911e+06, 1.529e+06, 1.243e+06, 1.024e+06, 8.533e+05,
    7.186e+05, 6.109e+05, 5.237e+05, 4.523e+05, 3.933e+05, 3.443e+05, 3.030e+05, 2.679e+05,
    2.382e+05, 2.127e+05, 1.907e+05, 1.716e+05, 1.550e+05, 1.405e+05, 1.277e+05])
# voigt.c:141-143 / :186: the printed 17-digit literals are these formulas evaluated in
# double precision (tests/test_oracle_tables.py asserts equality with the compiled voigt.c
# through oracle/_ref and with the leading printed entries).
LEADING_CONSTANTS = math.pi * _E_CGS * _E_CGS * OSCILLATOR_STRENGTHS * TRANSITION_WAVELENGTHS / (_M_E * C_CGS)
GAMMAS = GAMMAS_RATE * TRANSITION_WAVELENGTHS / (4.0 * math.pi)

INSTRUMENT_PROFILE = np.array([      # voigt.c:242-251
    2.17460992138080811e-03, 4.11623059580451742e-02, 2.40309364651846963e-01,
    4.32707438937454059e-01,
    2.40309364651846963e-01, 4.11623059580451742e-02, 2.17460992138080811e-03])

NUM_LINES_MAX = 31                   # voigt.c:16


def cerf_voigt(x, sigma, gamma):
    """libcerf ``voigt(x, sigma, gamma)`` (call site voigt.c:288) via the Faddeeva package."""
    z = (x + 1j * abs(gamma)) / math.sqrt(2.0) / abs(sigma)
    return wofz(z).real / (math.sqrt(2.0 * math.pi) * abs(sigma))


def voigt(lambdas, z, N, num_lines=NUM_LINES_MAX):
    """``profile = voigt(lambdas, z, N [, num_lines])`` -- voigt.c:253-304.

    Returns ``len(lambdas) - 6`` values (instrument-convolved absorption)."""
    lambdas = np.asarray(lambdas, dtype=np.float64).ravel()
    num_points = lambdas.size
    multipliers = C_CGS / (TRANSITION_WAVELENGTHS[:num_lines] * (1.0 + z)) / 1e8   # voigt.c:279
    total = np.zeros(num_points)
    for j in range(num_lines):                                                     # voigt.c:285-290
        velocity = lambdas * multipliers[j] - C_CGS
        total += -LEADING_CONSTANTS[j] * cerf_voigt(velocity, SIGMA, GAMMAS[j])
    raw_profile = np.exp(N * total)                                                # voigt.c:291
    n_out = num_points - 2 * width
    profile = np.zeros(n_out)
    for kk in range(2 * width + 1):                                                # voigt.c:297-299
        profile += raw_profile[kk:kk + n_out] * INSTRUMENT_PROFILE[kk]
    return profile


# ---------------------------------------------------------------- log_mvnpdf_low_rank.m
LOG_2PI = 1.83787706640934534        # log_mvnpdf_low_rank.m:7


def log_mvnpdf_low_rank(y, mu, M, d):
    """``log N(y; mu, M M' + diag(d))`` -- log_mvnpdf_low_rank.m:5-34, same operation order."""
    from scipy.linalg import cholesky, solve_triangular
    n, k = M.shape
    y = y - mu                                           # :11
    d_inv = 1.0 / d                                      # :13
    D_inv_y = d_inv * y                                  # :14
    D_inv_M = d_inv[:, None] * M                         # :15
    B = M.T @ D_inv_M                                    # :22
    B[np.diag_indices(k)] += 1.0                         # :23
    L = cholesky(B, lower=False)                         # :24  (MATLAB chol = upper, L'L = B)
    C = solve_triangular(L, solve_triangular(L.T, D_inv_M.T, lower=True), lower=False)   # :26
    K_inv_y = D_inv_y - D_inv_M @ (C @ y)                # :28
    log_det_K = np.sum(np.log(d)) + 2.0 * np.sum(np.log(np.diag(L)))   # :30
    return -0.5 * (y @ K_inv_y + log_det_K + n * LOG_2PI)              # :32


# ---------------------------------------------------------------- process_qsos.m
def matlab_logspace(a, b, n):
    """MATLAB ``logspace(a, b, n)`` = ``10.^linspace(a, b, n)``."""
    return 10.0 ** np.linspace(a, b, n)


def padded_wavelengths(this_unmasked_wavelengths):
    """process_qsos.m:168-176."""
    lo = math.log10(np.min(this_unmasked_wavelengths))
    hi = math.log10(np.max(this_unmasked_wavelengths))
    return np.concatenate([
        matlab_logspace(lo - width * pixel_spacing, lo - pixel_spacing, width),
        this_unmasked_wavelengths,
        matlab_logspace(hi + pixel_spacing, hi + width * pixel_spacing, width)])


def log_priors(prior_z_qsos, prior_dla_ind, z_qso):
    """process_qsos.m:122-131."""
    less_ind = prior_z_qsos < (z_qso + prior_z_qso_increase)
    this_num_dlas = int(np.count_nonzero(prior_dla_ind[less_ind]))
    this_num_quasars = int(np.count_nonzero(less_ind))
    with np.errstate(divide="ignore", invalid="ignore"):
        lp_dla = np.log(np.float64(this_num_dlas)) - np.log(np.float64(this_num_quasars))
        lp_no = np.log(np.float64(this_num_quasars - this_num_dlas)) - np.log(np.float64(this_num_quasars))
    return lp_no, lp_dla


def process_one_quasar(model, samples, wavelengths, flux, noise_variance, pixel_mask, z_qso,
                       num_lines=num_lines_default, sample_subset=None, engine="numpy", nthreads=0):
    """Body of the per-quasar loop, process_qsos.m:96-213 (priors excluded).

    ``engine="c"`` runs the per-sample loop (:185-199) in ``oracle/c/gpdla_oracle.c`` (same
    arithmetic, threaded over samples); ``"numpy"`` is the literal pure-Python loop.

    ``sample_subset`` (optional index array) evaluates only those samples -- used to keep
    the pure-Python loop short in tests; the LSE then runs over that subset."""
    rest_wavelengths = model["rest_wavelengths"]
    mu, M, log_omega = model["mu"], model["M"], model["log_omega"]
    c_0, tau_0, beta = math.exp(model["log_c_0"]), math.exp(model["log_tau_0"]), math.exp(model["log_beta"])
    offset_samples = np.asarray(samples["offset_samples"], dtype=np.float64)
    nhi_samples = np.asarray(samples["nhi_samples"], dtype=np.float64)
    if sample_subset is not None:
        offset_samples = offset_samples[sample_subset]
        nhi_samples = nhi_samples[sample_subset]

    this_wavelengths = np.asarray(wavelengths, dtype=np.float64)
    this_pixel_mask = np.asarray(pixel_mask).astype(bool)
    this_rest_wavelengths = this_wavelengths / (1.0 + z_qso)                       # :102
    unmasked_ind = (this_rest_wavelengths >= min_lambda) & (this_rest_wavelengths <= max_lambda)   # :104-105
    this_unmasked_wavelengths = this_wavelengths[unmasked_ind]                     # :108
    ind = unmasked_ind & (~this_pixel_mask)                                        # :110
    out = dict(n_pixels=int(np.count_nonzero(ind)))
    if out["n_pixels"] == 0:
        return None
    this_wavelengths = this_wavelengths[ind]
    this_rest_wavelengths = this_rest_wavelengths[ind]
    this_flux = np.asarray(flux, dtype=np.float64)[ind]
    this_noise_variance = np.asarray(noise_variance, dtype=np.float64)[ind]
    this_lya_zs = (this_wavelengths - lya_wavelength) / lya_wavelength            # :117-119

    this_mu = np.interp(this_rest_wavelengths, rest_wavelengths, mu)              # :138
    this_M = np.stack([np.interp(this_rest_wavelengths, rest_wavelengths, M[:, j])
                       for j in range(M.shape[1])], axis=1)                        # :139
    this_log_omega = np.interp(this_rest_wavelengths, rest_wavelengths, log_omega)   # :141
    this_omega2 = np.exp(2.0 * this_log_omega)                                     # :142
    this_scaling_factor = 1.0 - np.exp(-tau_0 * (1.0 + this_lya_zs) ** beta) + c_0   # :144
    this_omega2 = this_omega2 * this_scaling_factor ** 2                           # :146

    out["log_likelihoods_no_dla"] = log_mvnpdf_low_rank(
        this_flux, this_mu, this_M, this_omega2 + this_noise_variance)             # :149-151
    out["min_z_dlas"] = min_z_dla(this_wavelengths, z_qso)                         # :159
    out["max_z_dlas"] = max_z_dla(this_wavelengths, z_qso)                         # :160
    sample_z_dlas = out["min_z_dlas"] + (out["max_z_dlas"] - out["min_z_dlas"]) * offset_samples   # :162-164
    padded = padded_wavelengths(this_unmasked_wavelengths)                         # :168-176
    ind2 = ~this_pixel_mask[unmasked_ind]                                          # :181

    S = offset_samples.size
    sll = np.empty(S)
    if engine == "c":
        from .ref import c_sample_loglik
        sll = c_sample_loglik(padded, ind2, this_flux, this_mu, this_M, this_omega2, this_noise_variance,
                              sample_z_dlas, nhi_samples, num_lines, nthreads=nthreads)
    for i in range(S if engine == "numpy" else 0):                                 # :185-199
        absorption = voigt(padded, sample_z_dlas[i], nhi_samples[i], num_lines)
        absorption = absorption[ind2]
        dla_mu = this_mu * absorption
        dla_M = this_M * absorption[:, None]
        dla_omega2 = this_omega2 * absorption ** 2
        sll[i] = log_mvnpdf_low_rank(this_flux, dla_mu, dla_M, dla_omega2 + this_noise_variance)
    out["sample_log_likelihoods_dla"] = sll
    max_log_likelihood = np.max(sll)                                               # :203
    sample_probabilities = np.exp(sll - max_log_likelihood)                        # :205-207
    out["log_likelihoods_dla"] = max_log_likelihood + math.log(np.mean(sample_probabilities))   # :209-210
    out["sample_z_dlas"] = sample_z_dlas
    return out


def model_posteriors(log_posteriors_no_dla, log_posteriors_dla):
    """process_qsos.m:224-233."""
    lp = np.stack([log_posteriors_no_dla, log_posteriors_dla], axis=1)
    mx = np.max(lp, axis=1, keepdims=True)
    mp = np.exp(lp - mx)
    mp = mp / np.sum(mp, axis=1, keepdims=True)
    p_no_dlas = mp[:, 0]
    return mp, p_no_dlas, 1.0 - p_no_dlas


def process_qsos(model, samples, spectra, prior, num_lines=num_lines_default, sample_subset=None,
                 engine="numpy", nthreads=0):
    """Whole-script restatement: process_qsos.m:63-233 plus the MAP of
    generate_ascii_catalog.m:73-80.  ``spectra`` holds ragged lists
    (``all_wavelengths`` ...) and ``z_qsos``; ``prior`` holds ``z_qsos`` and ``dla_ind``."""
    Q = len(spectra["z_qsos"])
    S = len(samples["offset_samples"]) if sample_subset is None else len(sample_subset)
    names = ["min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla",
             "log_likelihoods_no_dla", "log_likelihoods_dla", "log_posteriors_no_dla",
             "log_posteriors_dla", "map_z_dlas", "map_log_nhis"]
    res = {nm: np.full(Q, np.nan) for nm in names}                                 # :74-82
    res["sample_log_likelihoods_dla"] = np.full((Q, S), np.nan)
    res["map_inds"] = np.full(Q, -1, dtype=np.int64)
    offs = np.asarray(samples["offset_samples"], dtype=np.float64)
    lognhi = np.asarray(samples["log_nhi_samples"], dtype=np.float64)
    if sample_subset is not None:
        offs, lognhi = offs[sample_subset], lognhi[sample_subset]
    for q in range(Q):
        z_qso = float(spectra["z_qsos"][q])
        lp_no, lp_dla = log_priors(prior["z_qsos"], prior["dla_ind"], z_qso)
        res["log_priors_no_dla"][q], res["log_priors_dla"][q] = lp_no, lp_dla
        o = process_one_quasar(model, samples, spectra["all_wavelengths"][q], spectra["all_flux"][q],
                               spectra["all_noise_variance"][q], spectra["all_pixel_mask"][q], z_qso,
                               num_lines=num_lines, sample_subset=sample_subset, engine=engine,
                               nthreads=nthreads)
        if o is None:
            continue
        for nm in ("min_z_dlas", "max_z_dlas", "log_likelihoods_no_dla", "log_likelihoods_dla"):
            res[nm][q] = o[nm]
        res["sample_log_likelihoods_dla"][q] = o["sample_log_likelihoods_dla"]
        res["log_posteriors_no_dla"][q] = lp_no + o["log_likelihoods_no_dla"]      # :153-154
        res["log_posteriors_dla"][q] = lp_dla + o["log_likelihoods_dla"]           # :212-213
        sll = o["sample_log_likelihoods_dla"]
        if not np.all(np.isnan(sll)):
            map_ind = int(np.nanargmax(sll))                                       # generate_ascii_catalog.m:73
            res["map_inds"][q] = map_ind
            res["map_z_dlas"][q] = o["min_z_dlas"] + (o["max_z_dlas"] - o["min_z_dlas"]) * offs[map_ind]   # :75-76
            res["map_log_nhis"][q] = lognhi[map_ind]                               # :80
    mp, p_no, p_dla = model_posteriors(res["log_posteriors_no_dla"], res["log_posteriors_dla"])
    res["model_posteriors"], res["p_no_dlas"], res["p_dlas"] = mp, p_no, p_dla
    return res
