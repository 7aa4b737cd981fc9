"""Quick GPU-vs-oracle check + timing (development aid; the real tests live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_dla_detection_b200 import api, synthetic as syn
from oracle import process_qsos_oracle as O

S = int(os.environ.get("S", 10000)); Q = int(os.environ.get("Q", 3))
m = syn.make_model(); s = syn.make_samples(S); p = syn.make_prior(); sp = syn.make_spectra(m, Q, dla_fraction=0.5)
lam = 10 ** (3.5563 + 1e-4 * np.arange(1256))
for z, N, nl in [(2.3, 1e21, 3), (2.1, 10 ** 20.3, 31), (2.5, 1e23, 3), (2.2, 0.0, 3), (2.9, 10**20.0, 1)]:
    a = api.voigt(lam, z, N, nl); b = O.voigt(lam, z, N, nl)
    print("voigt", z, N, nl, "max abs err", np.max(np.abs(a - b)), "max rel", np.max(np.abs(a - b) / np.maximum(b, 1e-300) * (b > 1e-12)))
t = time.time(); r = api.process_qsos(m, s, sp, p); print("gpu process_qsos", time.time() - t)
t = time.time(); ro = O.process_qsos(m, s, sp, p, engine="c"); print("oracle(c)", time.time() - t)
for k in ["log_likelihoods_no_dla", "log_likelihoods_dla", "log_posteriors_dla", "p_dlas", "map_z_dlas", "map_log_nhis", "min_z_dlas", "max_z_dlas", "log_priors_dla", "log_priors_no_dla"]:
    print(k, r[k], ro[k], np.max(np.abs(r[k] - ro[k])))
print("map_inds", r["map_inds"], ro["map_inds"])
d = np.abs(r["sample_log_likelihoods_dla"] - ro["sample_log_likelihoods_dla"]) / np.abs(ro["sample_log_likelihoods_dla"])
print("sample ll max rel err", d.max(), "argmax", np.unravel_index(d.argmax(), d.shape))
# timing of the device path
proc = api.DLAProcessor(m, s, p)
Qb = int(os.environ.get("QB", 296))
spb = api.pad_spectra(syn.make_spectra(m, Qb))
dev = torch.device("cuda:0")
tt = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in spb.items()}
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = proc.process_device(tt["wavelengths"], tt["flux"], tt["noise_variance"], tt["pixel_mask"], tt["lengths"], tt["z_qsos"]); e1.record()
    torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    n = spb["lengths"]
    print("device path: %d quasars in %.2f ms -> %.1f quasars/s, %.3f ms/quasar" % (Qb, ms, Qb / ms * 1e3, ms / Qb))
flops = sum(S * 20 * 23 * np.count_nonzero((w / (1 + z) >= 911.75) & (w / (1 + z) <= 1215.75)) for w, z in zip(spb["wavelengths"], spb["z_qsos"]))
print("algorithmic TFLOP/s (n_u-based)", flops / ms * 1e-9)
