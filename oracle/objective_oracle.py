r"""TEST INFRASTRUCTURE (oracle): literal NumPy restatement of the reference's GP training objective,
spectrum_loss.m:14-74 and objective.m:12-73 and of their Lyman-series variants multi_dlas/spectrum_loss_lyseries.m:14-91,
multi_dlas/objective_lyseries.m:13-87, in the reference's operation order (chol upper, L \ (L' \ ...)).
Parity unpinned against MATLAB outputs; pinned by tests/test_objective.py: dense multivariate-normal log-density
and central finite differences of every gradient block."""
import numpy as np
from scipy.linalg import solve_triangular

log_2pi = 1.83787706640934534


def spectrum_loss(y, lya_1pz, noise_variance, M, omega2, c_0, tau_0, beta, lyseries=None):
    """spectrum_loss.m:14-74; with ``lyseries = (num_forest_lines, all_transition_wavelengths, all_oscillator_strengths,
    zqso_1pz)`` multi_dlas/spectrum_loss_lyseries.m:14-91 (the two differ in the optical depth only, :20-41)."""
    n, k = M.shape
    lya_optical_depth = tau_0 * lya_1pz ** beta                       # :21 / lyseries :23
    if lyseries is not None:
        num_forest_lines, tw, osc, zqso_1pz = lyseries
        for i in range(1, num_forest_lines):                          # lyseries :29-40 (MATLAB i = 2..num_forest_lines)
            lyman_1pz = tw[0] * lya_1pz / tw[i]
            with np.errstate(invalid="ignore"):
                indicator = lyman_1pz <= zqso_1pz
            lyman_1pz = lyman_1pz * indicator
            tau = tau_0 * tw[i] * osc[i] / (tw[0] * osc[0])
            lya_optical_depth = lya_optical_depth + tau * lyman_1pz ** beta
    lya_absorption = np.exp(-lya_optical_depth)                       # :22 / lyseries :41
    scaling_factor = 1 - lya_absorption + c_0                         # :25
    absorption_noise = omega2 * scaling_factor ** 2                   # :26
    d = noise_variance + absorption_noise                             # :28
    d_inv = 1.0 / d
    D_inv_y = d_inv * y
    D_inv_M = d_inv[:, None] * M
    B = M.T @ D_inv_M                                                 # :40
    B[np.diag_indices(k)] += 1                                        # :41
    L = np.linalg.cholesky(B).T                                       # :42  upper, L'L = B
    C = solve_triangular(L, solve_triangular(L.T, D_inv_M.T, lower=True), lower=False)   # :44
    K_inv_y = D_inv_y - D_inv_M @ (C @ y)                             # :46
    log_det_K = np.sum(np.log(d)) + 2 * np.sum(np.log(np.diag(L)))    # :48
    nlog_p = 0.5 * (y @ K_inv_y + log_det_K + n * log_2pi)            # :52
    K_inv_M = D_inv_M - D_inv_M @ (C @ M)                             # :55
    dM = -(np.outer(K_inv_y, K_inv_y @ M) - K_inv_M)                  # :56
    diag_K_inv = d_inv - np.sum(C * D_inv_M.T, axis=0)                # :59
    dlog_omega = -(absorption_noise * (K_inv_y ** 2 - diag_K_inv))    # :62
    da = c_0 * omega2 * scaling_factor                                # :65
    dlog_c_0 = -(K_inv_y * da) @ K_inv_y + diag_K_inv @ da
    da = omega2 * scaling_factor * lya_optical_depth * lya_absorption   # :69
    dlog_tau_0 = -(K_inv_y * da) @ K_inv_y + diag_K_inv @ da
    da = da * np.log(lya_1pz) * beta                                  # :73
    dlog_beta = -(K_inv_y * da) @ K_inv_y + diag_K_inv @ da
    return nlog_p, dM, dlog_omega, dlog_c_0, dlog_tau_0, dlog_beta


def objective(x, centered_rest_fluxes, lya_1pzs, rest_noise_variances, priors=True, lyseries=None):
    """objective.m:12-73; with ``lyseries = (num_forest_lines, all_transition_wavelengths, all_oscillator_strengths)``
    multi_dlas/objective_lyseries.m:13-87 (zqso_1pz = lya_1pzs(i, end), :46)."""
    num_quasars, num_pixels = centered_rest_fluxes.shape
    k = (x.size - 3) // num_pixels - 1                                # :17
    M = x[:num_pixels * k].reshape(k, num_pixels).T                   # :19-20 (column-major reshape)
    log_omega = x[num_pixels * k:num_pixels * (k + 1)]
    log_c_0, log_tau_0, log_beta = x[-3], x[-2], x[-1]
    omega2 = np.exp(2 * log_omega)
    c_0, tau_0, beta = np.exp(log_c_0), np.exp(log_tau_0), np.exp(log_beta)
    f = 0.0
    dM = np.zeros_like(M); dlog_omega = np.zeros_like(log_omega)
    dlog_c_0 = dlog_tau_0 = dlog_beta = 0.0
    for i in range(num_quasars):                                      # :41-57
        ind = ~np.isnan(centered_rest_fluxes[i])
        if not ind.any():
            continue
        r = spectrum_loss(centered_rest_fluxes[i, ind], lya_1pzs[i, ind], rest_noise_variances[i, ind], M[ind], omega2[ind],
                          c_0, tau_0, beta, None if lyseries is None else tuple(lyseries) + (lya_1pzs[i, -1],))
        f += r[0]; dM[ind] += r[1]; dlog_omega[ind] += r[2]
        dlog_c_0 += r[3]; dlog_tau_0 += r[4]; dlog_beta += r[5]
    if priors:
        dlog_tau_0 += tau_0 * (tau_0 - 0.0023) / 0.0007 ** 2          # :59-64
        dlog_beta += beta * (beta - 3.65) / 0.21 ** 2                 # :66-71
    g = np.concatenate([dM.T.ravel(), dlog_omega, [dlog_c_0, dlog_tau_0, dlog_beta]])   # :73
    return f, g


def make_training_set(num_quasars, num_pixels=1217, k=20, seed=0, missing=0.15):
    """Synthetic training matrices shaped like learn_qso_model.m:36-75 and a parameter vector x."""
    rng = np.random.default_rng(seed)
    rest = 911.75 + 0.25 * np.arange(num_pixels)
    z = 2.15 + rng.gamma(2.0, 0.3, num_quasars)
    lya_1pzs = (1 + z)[:, None] * rest[None, :] / 1215.6701
    M = 0.15 / np.sqrt(np.arange(1, k + 1))[None, :] * np.cos(np.pi * np.arange(1, k + 1)[None, :] * (rest[:, None] - 911.75) / 304.0) \
        + 0.02 * rng.standard_normal((num_pixels, k))
    log_omega = np.log(0.08 + 0.04 * rng.random(num_pixels))
    nv = (0.1 + 0.4 * rng.random((num_quasars, num_pixels))) ** 2
    y = (M @ rng.standard_normal((k, num_quasars))).T + np.sqrt(nv + np.exp(2 * log_omega) * 0.04) * rng.standard_normal((num_quasars, num_pixels))
    # missing pixels: a blue cut-off per quasar (low z: spectrograph edge) plus random masks (learn_qso_model.m:44-45,66-69)
    cut = rng.integers(0, num_pixels // 3, num_quasars)
    miss = (np.arange(num_pixels)[None, :] < cut[:, None]) | (rng.random((num_quasars, num_pixels)) < missing)
    y[miss] = np.nan; nv[miss] = np.nan; lya = lya_1pzs.copy(); lya[miss] = np.nan
    x = np.concatenate([M.T.ravel(), log_omega, [np.log(0.1), np.log(0.0023), np.log(3.65)]])
    return x, y, lya, nv
