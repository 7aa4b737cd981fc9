"""Device-path timing of one batch of quasars (development aid).
  QB=296 S=10000 DIGITS=6 REST_TABLE=0 python tools/i8_time.py      (DIGITS: 6 / 5 INT8 Gram, -1 FP64 DMMA; REST_TABLE: 0 / -1)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_dla_detection_b200 import api, synthetic as syn
S = int(os.environ.get("S", 10000)); Qb = int(os.environ.get("QB", 296)); K = int(os.environ.get("K", 20))
digits = int(os.environ.get("DIGITS", 0)); rt = int(os.environ.get("REST_TABLE", 0)); reps = int(os.environ.get("REPS", 4))
m = syn.make_model(K); s = syn.make_samples(S); p = syn.make_prior()
proc = api.DLAProcessor(m, s, p, gram_digits=digits, rest_table=rt)
spb = api.pad_spectra(syn.make_spectra(m, Qb))
dev = torch.device("cuda:0")
tt = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in spb.items()}
best = 1e9
for it in range(reps):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = proc.process_device(tt["wavelengths"], tt["flux"], tt["noise_variance"], tt["pixel_mask"], tt["lengths"], tt["z_qsos"]); e1.record()
    torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print("%s K=%d S=%d digits=%d rest_table=%d: %d quasars in %.2f ms -> %.1f quasars/s" % (os.environ.get("TAG", ""), K, S, digits, rt, Qb, best, Qb / best * 1e3))
