"""Small end-to-end run of every kernel path for compute-sanitizer (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_dla_detection_b200 import api, synthetic as syn
prior = syn.make_prior(2000)
for k in (10, 20, 40):
    m = syn.make_model(k); s = syn.make_samples(200)
    sp = syn.make_spectra(m, 3, seed=k)
    sp["all_pixel_mask"][1][:] = True
    for key in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
        sp[key][2] = sp[key][2][300:1000]
    r = api.process_qsos(m, s, sp, prior)
    print(k, r["log_likelihoods_dla"])
m = syn.make_model(20); s = syn.make_samples(200, with_lls=True)
sp = syn.make_spectra(m, 3, seed=9, meanflux=True, dla_fraction=0.7)
r = api.process_qsos_multiple_dlas_meanflux(m, s, sp, prior, max_dlas=4)
print(r["log_likelihoods_dla"])
from gp_dla_detection_b200.params import Parameters
r = api.process_qsos(m, s, sp, prior, params=Parameters(num_lines=31))
print(r["log_likelihoods_dla"], api.voigt(10 ** (3.56 + 1e-4 * np.arange(40)), 2.0, 1e21, 31)[:3])
