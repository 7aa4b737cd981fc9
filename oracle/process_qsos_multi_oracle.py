"""Literal CPU restatement of the reference's multi-DLA / sub-DLA / mean-flux path (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  parity unpinned (the reference has no
fixtures and MATLAB cannot run here).  Follows ``multi_dlas/process_qsos_multiple_dlas_meanflux.m``
line by line (cited below) with the parameters of ``multi_dlas/set_parameters_multi.m`` and
``multi_dlas/set_lls_parameters.m``; ``voigt`` and ``log_mvnpdf_low_rank`` are the single-DLA
oracle's.  Two MATLAB built-ins are restated from their documented behaviour:

* ``rng('default')`` + ``rand`` = MT19937 seeded 5489 with 53-bit doubles = ``numpy.random.RandomState(5489)``
  (first values 0.8147, 0.9058, 0.1270 checked in tests);
* ``randsample(n, k, true, w)`` (Statistics Toolbox, not in the reference tree) = inverse-CDF lookup of
  ``rand(k, 1)`` in ``edges = min([0 cumsum(w / sum(w))], 1); edges(end) = 1`` with ``histc`` binning
  (``edges(i) <= x < edges(i+1)``).
"""
from __future__ import annotations

import math

import numpy as np

from . import process_qsos_oracle as O

max_dlas_default = 4                                   # ...meanflux.m:32
min_z_separation = O.kms_to_z(3000.0)                  # :33
prev_tau_0 = 0.0023                                    # :36
prev_beta = 3.65                                       # :37
num_forest_lines = 31                                  # set_parameters_multi.m:76
all_transition_wavelengths = O.TRANSITION_WAVELENGTHS * 1e8   # set_parameters_multi.m:77-109 (Angstrom)
all_oscillator_strengths = O.OSCILLATOR_STRENGTHS             # :111-143
lya_oscillator_strength = 0.416400                            # :144


def matlab_randsample_weighted(rs: np.random.RandomState, n: int, k: int, w: np.ndarray) -> np.ndarray:
    """``randsample(n, k, true, w)``; returns 0-based indices."""
    p = w / np.sum(w)
    edges = np.minimum(np.concatenate([[0.0], np.cumsum(p)]), 1.0)
    edges[-1] = 1.0
    u = rs.random_sample(k)
    return np.searchsorted(edges, u, side="right") - 1


def multi_priors(prior_z_qsos, prior_dla_ind, z_qso, max_dlas, Z_lls, Z_dla):
    """...meanflux.m:190-216."""
    less_ind = prior_z_qsos < (z_qso + O.prior_z_qso_increase)
    this_num_dlas = float(np.count_nonzero(prior_dla_ind[less_ind]))
    this_num_quasars = float(np.count_nonzero(less_ind))
    with np.errstate(divide="ignore", invalid="ignore"):
        this_p_dlas = (np.float64(this_num_dlas) / np.float64(this_num_quasars)) ** np.arange(1, max_dlas + 1)
        for i in range(max_dlas - 1):                                              # :197-199
            this_p_dlas[i] = this_p_dlas[i] - this_p_dlas[i + 1]
        log_priors_dla = np.log(this_p_dlas)                                       # :204
        log_priors_lls = (np.log(np.float64(this_num_dlas)) - np.log(np.float64(this_num_quasars))
                          + math.log(Z_lls) - math.log(Z_dla))                     # :208-210
        log_priors_no_dla = (np.log(np.float64(this_num_quasars - this_num_dlas - Z_lls * this_num_dlas / Z_dla))
                             - np.log(np.float64(this_num_quasars)))               # :214-216
    return log_priors_no_dla, log_priors_lls, log_priors_dla


def suppressed_model(model, this_wavelengths, this_rest_wavelengths, z_qso):
    """Interpolated mu, M, omega2 with Lyman-series mean-flux suppression and noise scaling (:228-293)."""
    rest_wavelengths, mu, M, log_omega = model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"]
    c_0, tau_0, beta = math.exp(model["log_c_0"]), math.exp(model["log_tau_0"]), math.exp(model["log_beta"])
    this_lya_zs = (this_wavelengths - O.lya_wavelength) / O.lya_wavelength        # :175-177
    this_mu = np.interp(this_rest_wavelengths, rest_wavelengths, mu)
    this_M = np.stack([np.interp(this_rest_wavelengths, rest_wavelengths, M[:, j]) for j in range(M.shape[1])], axis=1)
    this_omega2 = np.exp(2.0 * np.interp(this_rest_wavelengths, rest_wavelengths, log_omega))   # :240-241
    lya_optical_depth = tau_0 * (1.0 + this_lya_zs) ** beta                       # :245
    for l in range(1, num_forest_lines):                                          # :247-259
        lyman_1pz = all_transition_wavelengths[0] * (1.0 + this_lya_zs) / all_transition_wavelengths[l]
        indicator = lyman_1pz <= (1.0 + z_qso)
        lyman_1pz = lyman_1pz * indicator
        tau = tau_0 * all_transition_wavelengths[l] * all_oscillator_strengths[l] / (
            all_transition_wavelengths[0] * all_oscillator_strengths[0])
        lya_optical_depth = lya_optical_depth + tau * lyman_1pz ** beta
    this_scaling_factor = 1.0 - np.exp(-lya_optical_depth) + c_0                  # :261
    this_omega2 = this_omega2 * this_scaling_factor ** 2                          # :263
    total = np.zeros(this_wavelengths.size)
    for l in range(num_forest_lines):                                             # :269-283
        this_lyseries_zs = (this_wavelengths - all_transition_wavelengths[l]) / all_transition_wavelengths[l]   # :183-187
        this_tau_0 = prev_tau_0 * all_oscillator_strengths[l] / lya_oscillator_strength * \
            all_transition_wavelengths[l] / O.lya_wavelength
        depth = this_tau_0 * ((1.0 + this_lyseries_zs) ** prev_beta)
        if l > 0:
            depth = np.where(this_lyseries_zs > z_qso, 0.0, depth)                # nan + nansum == skip
        total = total + depth
    lya_absorption = np.exp(-total)                                               # :285
    return this_mu * lya_absorption, this_M * lya_absorption[:, None], this_omega2 * lya_absorption ** 2   # :287-293


def process_one_quasar_multi(model, samples, wavelengths, flux, noise_variance, pixel_mask, z_qso, max_dlas=4,
                             num_lines=3, base_sample_inds=None, engine="c", nthreads=0):
    """Per-quasar body, ...meanflux.m:141-479 (priors excluded).  ``base_sample_inds`` (optional,
    ``[max_dlas-1, S]`` 0-based) overrides the resampling of :466-472 for parity runs."""
    offset_samples = np.asarray(samples["offset_samples"], dtype=np.float64)
    log_nhi_samples = np.asarray(samples["log_nhi_samples"], dtype=np.float64)
    nhi_samples = np.asarray(samples["nhi_samples"], dtype=np.float64)
    lls_nhi_samples = np.asarray(samples["lls_nhi_samples"], dtype=np.float64)
    S = offset_samples.size
    rs = np.random.RandomState(5489)                                              # :143 rng('default')

    this_wavelengths = np.asarray(wavelengths, dtype=np.float64)
    this_pixel_mask = np.asarray(pixel_mask).astype(bool)
    this_rest_wavelengths = this_wavelengths / (1.0 + z_qso)                      # :159
    unmasked_ind = (this_rest_wavelengths >= O.min_lambda) & (this_rest_wavelengths <= O.max_lambda)
    this_unmasked_wavelengths = this_wavelengths[unmasked_ind]                    # :166
    ind = unmasked_ind & (~this_pixel_mask)
    if np.count_nonzero(ind) == 0:                                                # :227-238 (empty spectrum)
        return None
    this_wavelengths = this_wavelengths[ind]
    this_rest_wavelengths = this_rest_wavelengths[ind]
    this_flux = np.asarray(flux, dtype=np.float64)[ind]
    this_noise_variance = np.asarray(noise_variance, dtype=np.float64)[ind]
    this_mu, this_M, this_omega2 = suppressed_model(model, this_wavelengths, this_rest_wavelengths, z_qso)

    out = {}
    out["log_likelihoods_no_dla"] = O.log_mvnpdf_low_rank(this_flux, this_mu, this_M,
                                                          this_omega2 + this_noise_variance)   # :296-298
    out["min_z_dlas"] = O.min_z_dla(this_wavelengths, z_qso)
    out["max_z_dlas"] = O.max_z_dla(this_wavelengths, z_qso)
    sample_z_dlas = out["min_z_dlas"] + (out["max_z_dlas"] - out["min_z_dlas"]) * offset_samples   # :309-311
    padded = O.padded_wavelengths(this_unmasked_wavelengths)                      # :322-330
    mask_ind = ~this_pixel_mask[unmasked_ind]                                     # :335

    def loglik(nhi, partners):
        if engine == "c":
            from .ref import c_sample_loglik
            return c_sample_loglik(padded, mask_ind, this_flux, this_mu, this_M, this_omega2, this_noise_variance,
                                   sample_z_dlas, nhi, num_lines, partners=partners, nthreads=nthreads)
        res = np.empty(S)
        for i in range(S):
            absorption = O.voigt(padded, sample_z_dlas[i], nhi[i], num_lines)
            if partners is not None:
                for j in range(partners.shape[0]):
                    kk = partners[j, i]
                    absorption = absorption * O.voigt(padded, sample_z_dlas[kk], nhi[kk], num_lines)
            absorption = absorption[mask_ind]
            res[i] = O.log_mvnpdf_low_rank(this_flux, this_mu * absorption, this_M * absorption[:, None],
                                           this_omega2 * absorption ** 2 + this_noise_variance)
        return res

    this_sll = np.full((S, max_dlas), np.nan)                                     # :146
    this_base = np.zeros((max_dlas - 1, S), dtype=np.int64)                       # :313
    out["log_likelihoods_dla"] = np.full(max_dlas, np.nan)
    out["MAP_z_dlas"] = np.full((max_dlas, max_dlas), np.nan)
    out["MAP_log_nhis"] = np.full((max_dlas, max_dlas), np.nan)
    out["MAP_inds"] = np.full((max_dlas, max_dlas), -1, dtype=np.int64)
    logS = math.log(S)
    with np.errstate(invalid="ignore", divide="ignore"):
        for num_dlas in range(1, max_dlas + 1):                                   # :337
            partners = this_base[:num_dlas - 1] if num_dlas > 1 else None
            this_sll[:, num_dlas - 1] = loglik(nhi_samples, partners) - logS      # :359-361
            if num_dlas == 1:                                                     # :365-380
                sll_lls = loglik(lls_nhi_samples, None) - logS
                out["sample_log_likelihoods_lls"] = sll_lls
            if num_dlas > 1:                                                      # :386-392
                idx = this_base[:num_dlas - 1]
                all_z_dlas = np.vstack([sample_z_dlas[None, :], sample_z_dlas[idx]])
                all_log_nhis = np.vstack([log_nhi_samples[None, :], log_nhi_samples[idx]])
                bad = np.any(np.diff(np.sort(all_z_dlas, axis=0), axis=0) < min_z_separation, axis=0)
                this_sll[bad, num_dlas - 1] = np.nan
            else:
                all_z_dlas, all_log_nhis = sample_z_dlas[None, :], log_nhi_samples[None, :]
            col = this_sll[:, num_dlas - 1]
            if np.all(np.isnan(col)):
                max_ll = np.nan
            else:
                max_ll = np.nanmax(col)                                           # :400-401
            sample_probabilities = np.exp(col - max_ll)                           # :403-405
            out["log_likelihoods_dla"][num_dlas - 1] = (max_ll + math.log(np.nanmean(sample_probabilities))
                                                        - logS * (num_dlas - 1)) if not np.isnan(max_ll) else np.nan   # :407-409
            if num_dlas == 1:                                                     # :416-430
                mlls = np.nanmax(sll_lls)
                out["log_likelihoods_lls"] = mlls + math.log(np.nanmean(np.exp(sll_lls - mlls)))
            if not np.isnan(max_ll):                                              # :439-445
                maxidx = int(np.nanargmax(col))
                comp = [maxidx] + [int(this_base[j, maxidx]) for j in range(num_dlas - 1)]
                out["MAP_inds"][num_dlas - 1, :num_dlas] = comp
                out["MAP_z_dlas"][num_dlas - 1, :num_dlas] = all_z_dlas[:, maxidx]
                out["MAP_log_nhis"][num_dlas - 1, :num_dlas] = all_log_nhis[:, maxidx]
            if num_dlas == max_dlas:                                              # :452-454
                break
            if np.isnan(out["log_likelihoods_dla"][num_dlas - 1]):                # :460-464
                break
            W = np.where(np.isnan(sample_probabilities), 0.0, sample_probabilities)   # :467-469
            drawn = matlab_randsample_weighted(rs, S, S, W)                       # :471-472
            this_base[num_dlas - 1] = drawn if base_sample_inds is None else base_sample_inds[num_dlas - 1]
    out["sample_log_likelihoods_dla"] = this_sll
    out["base_sample_inds"] = this_base
    return out


def process_qsos_multi(model, samples, spectra, prior, Z_lls, Z_dla, max_dlas=4, num_lines=3,
                       base_sample_inds=None, engine="c", nthreads=0):
    """Whole-script restatement (...meanflux.m:100-495).  ``samples`` additionally holds
    ``lls_nhi_samples`` (set_lls_parameters.m:55-62)."""
    Q = len(spectra["z_qsos"])
    S = len(samples["offset_samples"])
    res = {n: np.full(Q, np.nan) for n in ("min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_lls",
                                           "log_likelihoods_no_dla", "log_likelihoods_lls",
                                           "log_posteriors_no_dla", "log_posteriors_lls")}
    for n in ("log_priors_dla", "log_likelihoods_dla", "log_posteriors_dla"):
        res[n] = np.full((Q, max_dlas), np.nan)
    res["sample_log_likelihoods_dla"] = np.full((Q, S, max_dlas), np.nan)
    res["sample_log_likelihoods_lls"] = np.full((Q, S), np.nan)
    res["base_sample_inds"] = np.zeros((Q, S, max_dlas - 1), dtype=np.int64)
    res["MAP_z_dlas"] = np.full((Q, max_dlas, max_dlas), np.nan)
    res["MAP_log_nhis"] = np.full((Q, max_dlas, max_dlas), np.nan)
    res["MAP_inds"] = np.full((Q, max_dlas, max_dlas), -1, dtype=np.int64)
    for q in range(Q):
        z_qso = float(spectra["z_qsos"][q])
        lp_no, lp_lls, lp_dla = multi_priors(prior["z_qsos"], prior["dla_ind"], z_qso, max_dlas, Z_lls, Z_dla)
        res["log_priors_no_dla"][q], res["log_priors_lls"][q], res["log_priors_dla"][q] = lp_no, lp_lls, lp_dla
        o = process_one_quasar_multi(model, samples, spectra["all_wavelengths"][q], spectra["all_flux"][q],
                                     spectra["all_noise_variance"][q], spectra["all_pixel_mask"][q], z_qso,
                                     max_dlas=max_dlas, num_lines=num_lines,
                                     base_sample_inds=None if base_sample_inds is None else base_sample_inds[q].T,
                                     engine=engine, nthreads=nthreads)
        if o is None:
            continue
        for n in ("min_z_dlas", "max_z_dlas", "log_likelihoods_no_dla", "log_likelihoods_lls"):
            res[n][q] = o[n]
        res["log_likelihoods_dla"][q] = o["log_likelihoods_dla"]
        res["sample_log_likelihoods_dla"][q] = o["sample_log_likelihoods_dla"]
        res["sample_log_likelihoods_lls"][q] = o["sample_log_likelihoods_lls"]
        res["base_sample_inds"][q] = o["base_sample_inds"].T                      # :476
        res["MAP_z_dlas"][q], res["MAP_log_nhis"][q], res["MAP_inds"][q] = o["MAP_z_dlas"], o["MAP_log_nhis"], o["MAP_inds"]
        res["log_posteriors_no_dla"][q] = lp_no + o["log_likelihoods_no_dla"]     # :300-301
        res["log_posteriors_lls"][q] = lp_lls + o["log_likelihoods_lls"]          # :428-430
        res["log_posteriors_dla"][q] = lp_dla + o["log_likelihoods_dla"]          # :411-413
    lp = np.column_stack([res["log_posteriors_no_dla"], res["log_posteriors_lls"], res["log_posteriors_dla"]])
    with np.errstate(invalid="ignore"):
        mx = np.nanmax(lp, axis=1, keepdims=True)                                 # :482-483 (MATLAB max ignores NaN)
        mp = np.exp(lp - mx)                                                      # :485-488
        mp = mp * (1.0 / np.sum(mp, axis=1, keepdims=True))                       # :490-491 (sum propagates NaN)
    res["model_posteriors"] = mp
    res["p_no_dlas"] = mp[:, 0]
    res["p_lls"] = mp[:, 1]
    res["p_dlas"] = 1.0 - mp[:, 0] - mp[:, 1]                                     # :493-495
    return res
