"""Step-0 library denominators: cuBLAS DGEMM peak, hot-shape bmm, batched Cholesky (torch FP64)."""
import json, os, time, torch
dev = "cuda"
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {"cpu_count": os.cpu_count()}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = t(lambda: a @ b); out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    del a, b
S, n, N = 10000, 1217, 232
W = torch.rand(S, n, dtype=torch.float64, device=dev); P = torch.randn(n, N, dtype=torch.float64, device=dev)
ms = t(lambda: W @ P); out["hot_gemm_1q_ms"] = ms; out["hot_gemm_1q_tflops"] = 2 * S * n * N / ms * 1e-9
W = torch.rand(16, S, n, dtype=torch.float64, device=dev); P = torch.randn(16, n, N, dtype=torch.float64, device=dev)
ms = t(lambda: torch.bmm(W, P)); out["hot_bmm_16q_ms"] = ms; out["hot_bmm_16q_tflops"] = 16 * 2 * S * n * N / ms * 1e-9
del W, P
M = torch.randn(S, n, 20, dtype=torch.float64, device=dev)
ms = t(lambda: torch.bmm(M.transpose(1, 2), M)); out["naive_gram_1q_ms"] = ms
B = torch.bmm(M.transpose(1, 2), M) + torch.eye(20, dtype=torch.float64, device=dev)
ms = t(lambda: torch.linalg.cholesky(B)); out["chol_20x20_x1e4_ms"] = ms
x = torch.rand(1 << 26, dtype=torch.float64, device=dev) + 0.5
for name, f in (("exp", torch.exp), ("log", torch.log), ("rcp", torch.reciprocal)):
    ms = t(lambda: f(x)); out[f"torch_{name}_gelem_s"] = x.numel() / ms * 1e-6
print(json.dumps(out))
