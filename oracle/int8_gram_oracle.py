"""TEST INFRASTRUCTURE (oracle): CPU restatement (NumPy, exact integer arithmetic) of the numeric scheme of the INT8
tcgen05 Gram path (gp_dla_detection_b200/csrc/gpdla_i8_kernels.cuh):

  w = a^2/d, u = a (y - a mu)/d, d = a^2 omega^2 + v                      (process_qsos.m:192-198)
  W'' = w * cw,  cw_i = CAP (omega2_i + v_i)               in [0, CAP]   (w is maximal at a = 1)
  U'' = u * cu,  cu_i = CAP / (b_i (|y_i| + |mu_i|)),  b_i = max_a a/(a^2 omega2 + v)   in [-CAP, CAP]
  P''_ic = m_ip m_iq / cw_i * 2^-eP_c,   M''_ic = m_ic / cu_i * 2^-eM_c   (column exponents: max |.| <= CAP)
  X = rint(x 2^F), F = 8 L - 1, split into L signed 8-bit digits (bias trick), slice pairs with
  i + j >= L - 1 contracted exactly in integers (s32 accumulators per diagonal), recombined in FP64.

Prints the error of the log-likelihoods against the plain FP64 evaluation for L = 4, 5, 6 on a synthetic
quasar and on a fuzzed one (noise variance over 10 decades, outliers), i.e. what tests/test_gpu_parity.py
asks of the CUDA path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_dla_detection_b200 import synthetic as syn
from oracle import process_qsos_oracle as O

CAP = 0.996      # digits of |x| <= CAP never overflow the top signed digit


def digits(X, L):
    """x in [-CAP, CAP] -> L signed 8-bit digits d_0 (least significant) .. d_{L-1}:  x ~= sum d_i 2^(8 i) / 2^F."""
    F = 8 * L - 1
    Xi = np.rint(X * 2.0 ** F).astype(np.int64)
    bias = sum(128 << (8 * i) for i in range(L))
    Xb = Xi + bias
    assert np.all(Xb >= 0) and np.all(Xb < (1 << (8 * L)))
    return [(((Xb >> (8 * i)) & 0xFF) ^ 0x80).astype(np.uint8).view(np.int8).astype(np.int64) for i in range(L)]


def quasar_terms(model, sp, q, samples, sub):
    z_qso = float(sp["z_qsos"][q]); lam = sp["all_wavelengths"][q]; mask = sp["all_pixel_mask"][q]
    rest = lam / (1 + z_qso); win = (rest >= O.min_lambda) & (rest <= O.max_lambda); ind = win & ~mask
    k = model["M"].shape[1]
    mu = np.interp(rest[ind], model["rest_wavelengths"], model["mu"])
    M = np.stack([np.interp(rest[ind], model["rest_wavelengths"], model["M"][:, j]) for j in range(k)], 1)
    om2 = np.exp(2 * np.interp(rest[ind], model["rest_wavelengths"], model["log_omega"])) * \
        (1 - np.exp(-np.exp(model["log_tau_0"]) * (lam[ind] / O.lya_wavelength) ** np.exp(model["log_beta"])) + np.exp(model["log_c_0"])) ** 2
    y, v = sp["all_flux"][q][ind], sp["all_noise_variance"][q][ind]
    padded = O.padded_wavelengths(lam[win]); keep = ~mask[win]
    zmin, zmax = O.min_z_dla(lam[ind], z_qso), O.max_z_dla(lam[ind], z_qso)
    A = np.stack([O.voigt(padded, zmin + (zmax - zmin) * samples["offset_samples"][s], samples["nhi_samples"][s], 3)[keep] for s in sub])
    A = np.vstack([A, np.ones((1, A.shape[1]))])    # the null model rides along as one more sample
    return y, v, mu, om2, M, A


def loglik(G, g, y, v, mu, om2, A, iu, k):
    d = A ** 2 * om2 + v
    out = np.empty(A.shape[0])
    for s in range(A.shape[0]):
        B = np.zeros((k, k)); B[iu] = G[s]; B = B + B.T - np.diag(np.diag(B)) + np.eye(k)
        Lc = np.linalg.cholesky(B); zz = np.linalg.solve(Lc, g[s])
        out[s] = -0.5 * (np.sum((y - A[s] * mu) ** 2 / d[s]) - zz @ zz + np.sum(np.log(d[s])) + 2 * np.sum(np.log(np.diag(Lc))) + y.size * O.LOG_2PI)
    return out


def study(tag, y, v, mu, om2, M, A, Ls=(4, 5, 6)):
    k = M.shape[1]
    iu = np.triu_indices(k)
    d = A ** 2 * om2 + v
    W, U = A ** 2 / d, A * (y - A * mu) / d
    P = M[:, iu[0]] * M[:, iu[1]]
    ref = loglik(W @ P, U @ M, y, v, mu, om2, A, iu, k)
    cw = CAP * (om2 + v)
    b = np.where(v >= om2, 1.0 / (om2 + v), 0.5 / np.sqrt(om2 * v))
    yy = np.abs(y) + np.abs(mu)
    cu = np.where(yy > 0, CAP / (b * np.where(yy > 0, yy, 1.0)), 0.0)
    Wn, Un = W * cw, U * cu
    assert Wn.min() >= 0 and Wn.max() <= CAP * (1 + 1e-12) and np.abs(Un).max() <= CAP * (1 + 1e-12), (Wn.max(), np.abs(Un).max())
    Pn = P / cw[:, None]
    Mn = np.where(cu[:, None] > 0, M / np.where(cu > 0, cu, 1.0)[:, None], 0.0)
    eP = np.ceil(np.log2(np.maximum(np.abs(Pn).max(0), 1e-300) / CAP)); eM = np.ceil(np.log2(np.maximum(np.abs(Mn).max(0), 1e-300) / CAP))
    Pn, Mn = Pn * 2.0 ** -eP, Mn * 2.0 ** -eM
    print("%s: n = %d, %d samples, |ll| ~ %.0f, dynamic range of 1/cw %.1e" % (tag, y.size, A.shape[0], np.median(np.abs(ref)), cw.max() / cw.min()))
    out = {}
    for L in Ls:
        F = 8 * L - 1
        dW, dU, dP, dM = digits(Wn, L), digits(Un, L), digits(Pn, L), digits(Mn, L)
        G = np.zeros_like(W @ P); g = np.zeros_like(U @ M); pairs = 0; accmax = 0
        for s in range(L - 1, 2 * L - 1):                # diagonals, most significant last
            accG = np.zeros(G.shape, np.int64); accg = np.zeros(g.shape, np.int64)
            for i in range(L):
                j = s - i
                if 0 <= j < L:
                    accG += dW[i] @ dP[j]; accg += dU[i] @ dM[j]; pairs += 1
            accmax = max(accmax, np.abs(accG).max(), np.abs(accg).max())
            G += accG * 2.0 ** (8 * s - 2 * F); g += accg * 2.0 ** (8 * s - 2 * F)
        G *= 2.0 ** eP; g *= 2.0 ** eM
        ll = loglik(G, g, y, v, mu, om2, A, iu, k)
        err = np.max(np.abs(ll - ref) / np.abs(ref))
        out[L] = (err, float(np.log2(accmax)))
        print("  L = %d (%2d int8 MMA pairs, |s32 acc| <= 2^%.1f): max rel err of log-likelihood %.2e;  Gram err / scale %.2e"
              % (L, pairs, np.log2(accmax), err, np.max(np.abs(G - W @ P)) / np.max(np.abs(W @ P))))
    return out


def fuzzed_and_plain(step=50, Ls=(4, 5, 6)):
    """Errors {L: (max rel err of log-likelihood, log2 max |accumulator|)} for a synthetic and a fuzzed quasar."""
    model, samples = syn.make_model(), syn.make_samples(10000)
    sub = np.arange(0, 10000, step)
    sp = syn.make_spectra(model, 2, seed=2, dla_fraction=1.0)
    plain = study("synthetic", *quasar_terms(model, sp, 0, samples, sub), Ls=Ls)
    rng = np.random.default_rng(2024)
    Lq = len(sp["all_flux"][1])
    sp["all_noise_variance"][1] = 10.0 ** rng.uniform(-6, 4, Lq)
    sp["all_flux"][1] = sp["all_flux"][1] + rng.standard_normal(Lq) * np.sqrt(sp["all_noise_variance"][1])
    out = rng.random(Lq) < 0.01
    sp["all_flux"][1][out] = rng.uniform(-50, 50, np.count_nonzero(out))
    fuzz = study("fuzzed   ", *quasar_terms(model, sp, 1, samples, sub), Ls=Ls)
    return plain, fuzz


if __name__ == "__main__":
    fuzzed_and_plain()
