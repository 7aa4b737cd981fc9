"""Device-path timing of one 296-quasar batch (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_dla_detection_b200 import api, synthetic as syn
S = int(os.environ.get("S", 10000)); Qb = int(os.environ.get("QB", 296))
m = syn.make_model(); s = syn.make_samples(S); p = syn.make_prior()
proc = api.DLAProcessor(m, s, p)
spb = api.pad_spectra(syn.make_spectra(m, Qb))
dev = torch.device("cuda:0")
tt = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in spb.items()}
best = 1e9
for it in range(4):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = proc.process_device(tt["wavelengths"], tt["flux"], tt["noise_variance"], tt["pixel_mask"], tt["lengths"], tt["z_qsos"]); e1.record()
    torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print("%s GRAM=%s DEBUG=%s DIGITS=%s: %d quasars in %.2f ms -> %.1f quasars/s" % (os.environ.get("TAG", ""), os.environ.get("GPDLA_GRAM", "i8"), os.environ.get("GPDLA_I8_DEBUG", "0"), os.environ.get("GPDLA_I8_DIGITS", "6"), Qb, best, Qb / best * 1e3))
