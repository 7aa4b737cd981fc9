#!/usr/bin/env python
"""Regenerate the committed golden vectors (run in the build container, where /root/reference exists).

voigt_reference.npz      outputs of the REFERENCE's own voigt.c (compiled into oracle/_ref by
                         oracle/Makefile, libcerf replaced by SciPy's Faddeeva wofz) on fixed inputs
process_qsos_small.npz   a 3-quasar, 48-sample problem: inputs + outputs of the literal numpy oracle
                         (oracle/process_qsos_oracle.py, per-sample Python loop)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from gp_dla_detection_b200 import synthetic as syn  # noqa: E402
from oracle import process_qsos_oracle as O  # noqa: E402
from oracle.ref import have_ref, ref_voigt  # noqa: E402

VOIGT_CASES = [(2.3, 1e21, 3), (2.1, 10 ** 20.3, 31), (2.5, 1e23, 3), (2.2, 0.0, 3), (2.9, 1e20, 1),
               (3.7, 10 ** 21.7, 3), (2.0, 10 ** 22.4, 5), (4.4, 10 ** 20.05, 3)]


def main():
    assert have_ref(), "oracle/_ref/voigt_ref.so missing: run `make -C oracle`"
    lam = 10.0 ** (3.5563 + 1e-4 * np.arange(1256))
    lam_hi = 10.0 ** (3.70 + 1e-4 * np.arange(700))
    out = dict(lambdas=lam, lambdas_hi=lam_hi, cases=np.array(VOIGT_CASES))
    for i, (z, N, nl) in enumerate(VOIGT_CASES):
        out["profile_%d" % i] = ref_voigt(lam if z < 3.5 else lam_hi, z, N, int(nl))
    np.savez_compressed(os.path.join(HERE, "voigt_reference.npz"), **out)

    model = syn.make_model()
    samples = syn.make_samples(10000)
    sub = np.arange(7, 10000, 211)[:48]
    samples = {k: v[sub] for k, v in samples.items()}
    prior = syn.make_prior(5000)
    sp = syn.make_spectra(model, 3, seed=99, dla_fraction=0.7)
    # make quasar 2 ragged: drop its blue end (BOSS coverage limit) and mask a block of pixels
    cut = 300
    for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
        sp[k][2] = sp[k][2][cut:]
    sp["all_pixel_mask"][1][400:430] = True
    res = O.process_qsos(model, samples, sp, prior)
    pack = dict(model_rest_wavelengths=model["rest_wavelengths"], model_mu=model["mu"], model_M=model["M"],
                model_log_omega=model["log_omega"],
                model_scalars=np.array([model["log_c_0"], model["log_tau_0"], model["log_beta"]]),
                offset_samples=samples["offset_samples"], log_nhi_samples=samples["log_nhi_samples"],
                nhi_samples=samples["nhi_samples"], prior_z_qsos=prior["z_qsos"], prior_dla_ind=prior["dla_ind"],
                z_qsos=sp["z_qsos"])
    for q in range(3):
        for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
            pack["%s_%d" % (k, q)] = sp[k][q]
    for k, v in res.items():
        pack["out_" + k] = v
    np.savez_compressed(os.path.join(HERE, "process_qsos_small.npz"), **pack)
    # ---- multi-DLA + sub-DLA + mean-flux: 2 quasars, 40 samples, 3 levels, literal numpy loop
    from oracle import process_qsos_multi_oracle as MO
    ms = syn.make_samples(10000, with_lls=True)
    sub2 = np.arange(3, 10000, 251)[:40]
    ms = {k: (v[sub2] if isinstance(v, np.ndarray) else v) for k, v in ms.items()}
    msp = syn.make_spectra(model, 2, seed=123, dla_fraction=1.0, meanflux=True, max_injected=2)
    mres = MO.process_qsos_multi(model, ms, msp, prior, Z_lls=ms["Z_lls"], Z_dla=ms["Z_dla"], max_dlas=3,
                                 engine="numpy")
    mpack = dict(offset_samples=ms["offset_samples"], log_nhi_samples=ms["log_nhi_samples"],
                 nhi_samples=ms["nhi_samples"], lls_nhi_samples=ms["lls_nhi_samples"],
                 Z=np.array([ms["Z_lls"], ms["Z_dla"]]), z_qsos=msp["z_qsos"])
    for q in range(2):
        for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
            mpack["%s_%d" % (k, q)] = msp[k][q]
    for k, v in mres.items():
        mpack["out_" + k] = v
    np.savez_compressed(os.path.join(HERE, "process_qsos_multi_small.npz"), **mpack)
    print("wrote golden vectors:", {k: res[k] for k in ("log_likelihoods_no_dla", "log_likelihoods_dla", "p_dlas")})


if __name__ == "__main__":
    main()
