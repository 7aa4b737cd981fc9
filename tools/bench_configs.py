"""Throughput of the other BASELINE.json configurations on one GPU (device-resident inputs, CUDA events):
config 4 (multi-DLA, up to 4 DLAs + sub-DLA + mean flux) and config 5 (k = 10/20/40, 1e3..1e5 samples).
Prints one JSON object; numbers are recorded under profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_dla_detection_b200 import api, synthetic as syn

dev = torch.device("cuda:0")
out = {}
prior = syn.make_prior()


def time_single(k, S, Q):
    model = syn.make_model(k); samples = syn.make_samples(S)
    pad = api.pad_spectra(syn.make_spectra(model, Q, seed=1000 + k))
    t = {n: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for n, v in pad.items()}
    proc = api.DLAProcessor(model, samples, prior)
    args = [t[n] for n in ("wavelengths", "flux", "noise_variance", "pixel_mask", "lengths", "z_qsos")]
    proc.process_device(*args); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); proc.process_device(*args); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_u = np.array([np.count_nonzero((w[:l] / (1 + z) >= 911.75) & (w[:l] / (1 + z) <= 1215.75))
                    for w, l, z in zip(pad["wavelengths"], pad["lengths"], pad["z_qsos"])])
    flops = float(np.sum(n_u) * S * k * (k + 3))
    proc.close()
    return {"quasars": Q, "ms": ms, "quasars_per_s": Q / ms * 1e3, "gram_tflops": flops / ms * 1e-9}


for k, S, Q in [(10, 10000, 592), (20, 10000, 592), (40, 10000, 148), (20, 1000, 2960), (20, 100000, 74)]:
    out["single_k%d_S%d" % (k, S)] = time_single(k, S, Q)

# config 4
model = syn.make_model(20); samples = syn.make_samples(10000, with_lls=True)
Q = 296
sp = syn.make_spectra(model, Q, seed=4, dla_fraction=0.3, meanflux=True, max_injected=2)
proc = api.DLAProcessor(model, samples, prior)
proc.process_multi({k: v[:8] for k, v in sp.items()}, return_samples=False)
t0 = time.perf_counter(); res = proc.process_multi(sp, return_samples=False); dt = time.perf_counter() - t0
out["multi_dla_4levels_k20_S10000"] = {"quasars": Q, "ms": dt * 1e3, "quasars_per_s": Q / dt,
                                       "note": "host-buffer entry (H2D + D2H inside), 4 DLA levels + sub-DLA + null",
                                       "p_2dla_or_more": float(np.mean(np.argmax(res["model_posteriors"], axis=1) >= 3))}
print(json.dumps(out, indent=1))
