"""Feasibility study (CPU, NumPy) for taking the Gram off the FP64 pipe: split [W|U] and P into signed
7-bit integer slices (Ozaki scheme), contract slice pairs exactly in integer arithmetic (what the INT8
tcgen05 path with s32 accumulators would do), and see how many slices the hot path's parity budget needs.
Prints the error of the log-likelihoods vs the FP64 evaluation for L = 4..8 slices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_dla_detection_b200 import synthetic as syn
from oracle import process_qsos_oracle as O

BITS = 6   # magnitude bits per slice (values in [-64, 63])


def slices(X, axis, L):
    """X ~= 2^e * sum_i S_i 2^{-BITS (i+1)} with integer S_i, scaled by the max |X| along `axis`."""
    e = np.ceil(np.log2(np.max(np.abs(X), axis=axis, keepdims=True) + 1e-300))
    R = X / 2.0 ** e                      # |R| <= 1
    out = []
    for i in range(L):
        R = R * 2.0 ** BITS
        S = np.rint(R)
        R = R - S
        out.append(S.astype(np.int64))
    return out, e


model, samples, prior = syn.make_model(), syn.make_samples(10000), syn.make_prior()
sp = syn.make_spectra(model, 1, seed=2, dla_fraction=1.0)
sub = np.arange(0, 10000, 40)
# per-sample weights exactly as the kernel forms them
z_qso = float(sp["z_qsos"][0]); lam = sp["all_wavelengths"][0]; mask = sp["all_pixel_mask"][0]
rest = lam / (1 + z_qso); win = (rest >= O.min_lambda) & (rest <= O.max_lambda); ind = win & ~mask
mu = np.interp(rest[ind], model["rest_wavelengths"], model["mu"])
M = np.stack([np.interp(rest[ind], model["rest_wavelengths"], model["M"][:, j]) for j in range(20)], 1)
om2 = np.exp(2 * np.interp(rest[ind], model["rest_wavelengths"], model["log_omega"])) * \
    (1 - np.exp(-np.exp(model["log_tau_0"]) * (lam[ind] / O.lya_wavelength) ** np.exp(model["log_beta"])) + np.exp(model["log_c_0"])) ** 2
y, v = sp["all_flux"][0][ind], sp["all_noise_variance"][0][ind]
padded = O.padded_wavelengths(lam[win]); keep = ~mask[win]
zmin, zmax = O.min_z_dla(lam[ind], z_qso), O.max_z_dla(lam[ind], z_qso)
A = np.stack([O.voigt(padded, zmin + (zmax - zmin) * samples["offset_samples"][s], samples["nhi_samples"][s], 3)[keep] for s in sub])
d = A ** 2 * om2 + v
W, U = A ** 2 / d, A * (y - A * mu) / d
iu = np.triu_indices(20)
P = M[:, iu[0]] * M[:, iu[1]]
n = y.size


def loglik(G, g):
    out = np.empty(len(sub))
    for s in range(len(sub)):
        B = np.zeros((20, 20)); B[iu] = G[s]; B = B + B.T - np.diag(np.diag(B)) + np.eye(20)
        L = np.linalg.cholesky(B); zz = np.linalg.solve(L, g[s])
        out[s] = -0.5 * (np.sum((y - A[s] * mu) ** 2 / d[s]) - zz @ zz + np.sum(np.log(d[s])) + 2 * np.sum(np.log(np.diag(L))) + n * O.LOG_2PI)
    return out


ref = loglik(W @ P, U @ M)
print("reference log-likelihoods: n = %d pixels, %d samples, |ll| ~ %.0f" % (n, len(sub), np.median(np.abs(ref))))
for L in (4, 5, 6, 7, 8):
    Ws, eW = slices(W, 1, L); Us, eU = slices(U, 1, L); Ps, eP = slices(P, 0, L); Ms, eM = slices(M, 0, L)
    G = np.zeros((len(sub), P.shape[1])); g = np.zeros((len(sub), 20)); pairs = 0
    for i in range(L):
        for j in range(L - i):            # keep slice pairs with i + j < L
            sc = 2.0 ** (-BITS * (i + j + 2))
            G += (Ws[i] @ Ps[j]) * sc; g += (Us[i] @ Ms[j]) * sc; pairs += 1
    G *= 2.0 ** eW * 2.0 ** eP; g *= 2.0 ** eU * 2.0 ** eM
    ll = loglik(G, g)
    print("L = %d slices (%2d int8 GEMM pairs): max |rel err| of log-likelihood %.2e, max |err| of Gram entries / scale %.2e"
          % (L, pairs, np.max(np.abs(ll - ref) / np.abs(ref)), np.max(np.abs(G - W @ P)) / np.max(np.abs(W @ P))))
