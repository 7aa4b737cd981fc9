"""Config 4 (multi-DLA + sub-DLA + mean flux) on one GPU: wall time of the host-buffer entry for Q quasars (development aid).
  Q=296 S=10000 DIGITS=0 BATCH=0 python tools/multi_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_dla_detection_b200 import api, synthetic as syn
Q = int(os.environ.get("Q", 296)); S = int(os.environ.get("S", 10000)); digits = int(os.environ.get("DIGITS", 0))
model = syn.make_model(20); samples = syn.make_samples(S, with_lls=True); prior = syn.make_prior()
sp = syn.make_spectra(model, Q, seed=4, dla_fraction=0.3, meanflux=True, max_injected=2)
proc = api.DLAProcessor(model, samples, prior, gram_digits=digits, batch_quasars=int(os.environ.get("BATCH", 0)))
proc.process_multi({k: v[:8] for k, v in sp.items()}, return_samples=False)
best = 1e9
for _ in range(int(os.environ.get("REPS", 2))):
    t0 = time.perf_counter(); res = proc.process_multi(sp, return_samples=False); best = min(best, time.perf_counter() - t0)
print("%s multi-DLA Q=%d S=%d digits=%d: %.1f ms -> %.1f quasars/s" % (os.environ.get("TAG", ""), Q, S, digits, best * 1e3, Q / best))
