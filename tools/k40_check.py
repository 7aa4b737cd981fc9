"""k = 40 through the INT8 path (producing kernel + contract-only passes + Cholesky) against the FP64 DMMA path
(development aid).  S=1000 Q=3 python tools/k40_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_dla_detection_b200 import api, synthetic as syn
S = int(os.environ.get("S", 1000)); Q = int(os.environ.get("Q", 3)); K = int(os.environ.get("K", 40))
m = syn.make_model(K); s = syn.make_samples(S); p = syn.make_prior()
sp = syn.make_spectra(m, Q, seed=5, dla_fraction=0.67)
r = {d: api.process_qsos(m, s, sp, p, gram_digits=d) for d in (-1, 6)}
a, b = r[6]["sample_log_likelihoods_dla"], r[-1]["sample_log_likelihoods_dla"]
print("nan int8 / f64:", np.isnan(a).sum(), np.isnan(b).sum())
print("max rel sample ll:", np.nanmax(np.abs(a - b) / np.abs(b)))
print("no_dla:", r[6]["log_likelihoods_no_dla"], r[-1]["log_likelihoods_no_dla"])
print("map equal:", np.array_equal(r[6]["map_inds"], r[-1]["map_inds"]), "p_dla diff:", np.max(np.abs(r[6]["p_dlas"] - r[-1]["p_dlas"])))
