"""One evaluation of the GP training objective (objective.m) on one GPU vs the NumPy restatement on the host:
python tools/bench_objective.py [num_spectra]   -> one JSON object (recorded under profiles/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_dla_detection_b200 import api
from oracle import objective_oracle as OB

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
P, k = 1217, 20
x, y, lya, nv = OB.make_training_set(N, num_pixels=P, k=k, seed=5)
n_obs = int(np.count_nonzero(~np.isnan(y)))
ev = api.TrainingObjective(y, lya, nv, k)
xd = torch.from_numpy(x).cuda()
for _ in range(3):
    ev.evaluate_device(xd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    f, g = ev.evaluate_device(xd)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# host: the literal restatement on a bounded sample
ns = 48
t0 = time.perf_counter(); fr, gr = OB.objective(x, y[:ns], lya[:ns], nv[:ns], priors=False); dt = time.perf_counter() - t0
f48, g48 = api.objective(x, y[:ns], lya[:ns], nv[:ns])
g48[-2:] = gr[-2:]   # priors excluded on the host side of this comparison
flops = n_obs * (3.0 * k * k + 12.0 * k)          # B, K^-1 M, diag K^-1, projections: ~ n (3 k^2 + 12 k) per spectrum
print(json.dumps({
    "workload": "objective.m: %d training spectra x %d rest pixels, k = %d, %.1f %% of pixels observed" % (N, P, k, 100.0 * n_obs / (N * P)),
    "gpu_ms_per_evaluation": ms, "gpu_spectra_per_s": N / ms * 1e3,
    "algorithmic_tflops": flops / ms * 1e-9, "hbm_read_gb_per_s": 3.0 * N * P * 8 / ms * 1e-6,
    "cpu_numpy_spectra_per_s": ns / dt, "cpu_sample": "%d spectra, single process (%.2f s)" % (ns, dt),
    "f_rel_err_vs_oracle_on_sample": abs(f48 - fr) / abs(fr), "g_err_over_scale_on_sample": float(np.max(np.abs(g48 - gr)) / np.max(np.abs(gr))),
    "evaluations_per_lbfgs_run_reference": "minFunc MaxIter 4000 / MaxFunEvals 8000 (set_parameters.m:44-46)"}, indent=1))
