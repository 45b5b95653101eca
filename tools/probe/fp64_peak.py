"""Probe: measured cuBLAS DGEMM ceiling (the FP64 roofline denominator) + potrf reference on this B200."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2): torch.matmul(a, b, out=c)
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best / 1e9
    del a, b, c
for n in (8192,):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    torch.linalg.cholesky(spd)
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.linalg.cholesky(spd); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"torch_potrf_{n}_tflops"] = n**3 / 3 / best / 1e9
print(json.dumps(res))
