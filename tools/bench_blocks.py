"""Micro-benchmarks of the building blocks for BASELINE configs 4 and 5 and the K / dK build (run under gpurun).
Prints one JSON line per measurement; CUDA-event timing on the stream the kernels are launched on, 3 warm-ups,
buffers larger than L2 (or an L2 flush between iterations for the small ones)."""
import ctypes, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autogpc_b200.engine import Engine, Program, LEAF_SE, LEAF_PER

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
eng = Engine(0)
_stream = torch.cuda.Stream()   # explicit stream shared by torch and the engine (the engine's own stream is non-blocking)
torch.cuda.set_stream(_stream)
st = _stream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(f, reps=5, do_flush=False):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

def out(**kw):
    print(json.dumps(kw), flush=True)

# ---- write-only ceiling: a plain fill of the same buffer (the copy peak counts read + write bytes) ----
for n in (8192, 16384):
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    ms = timed(lambda: K.fill_(1.0))
    out(block="fill (torch)", n=n, ms=ms, gbs=8.0 * n * n / ms / 1e6, frac_hbm=8.0 * n * n / ms / 1e6 / HBM)
    ms = timed(lambda: K.zero_())
    out(block="zero (memset)", n=n, ms=ms, gbs=8.0 * n * n / ms / 1e6, frac_hbm=8.0 * n * n / ms / 1e6 / HBM)
    del K

# ---- K and dK build: HBM-write roofline (8 n^2 bytes, 8 n^2 (1+P) with dK) ----
progs = {
    "SE0": Program.build([LEAF_SE], [0], [0], [[0]], 2),
    "SE0xSE1": Program.build([LEAF_SE, LEAF_SE], [0, 1], [0, 2], [[0, 1]], 4),
    "SE0xSE1+PER2+SE3 (config 4)": Program.build([LEAF_SE, LEAF_SE, LEAF_PER, LEAF_SE], [0, 1, 2, 3], [0, 2, 4, 7], [[0, 1], [2], [3]], 9),
}
for n in (8192, 16384):
    X = torch.rand(n, 8, dtype=torch.float64, device="cuda") * 6 - 3
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    for name, p in progs.items():
        th = np.abs(np.random.default_rng(0).normal(1.0, 0.2, p.n_theta)) + 0.3
        ms = timed(lambda: eng.build_K(p, th, X.data_ptr(), n, 8, K.data_ptr(), n, 0, st))
        gbs = 8.0 * n * n / ms / 1e6
        out(block="build_K", program=name, n=n, ms=ms, gbs=gbs, frac_hbm=gbs / HBM)
    if n == 8192:
        for name, p in progs.items():
            dK = torch.empty(p.n_theta, n, n, dtype=torch.float64, device="cuda")
            th = np.abs(np.random.default_rng(0).normal(1.0, 0.2, p.n_theta)) + 0.3
            ms = timed(lambda: eng.build_K(p, th, X.data_ptr(), n, 8, K.data_ptr(), n, dK.data_ptr(), st))
            gbs = 8.0 * n * n * (1 + p.n_theta) / ms / 1e6
            out(block="build_K+dK", program=name, n=n, ms=ms, gbs=gbs, frac_hbm=gbs / HBM)
            del dK
    del X, K
torch.cuda.empty_cache()

if len(sys.argv) > 1 and sys.argv[1] == "kbuild":
    eng.close()
    sys.exit(0)

# ---- potrf / potrf+trtri: single large n (config 4) and batched small (config 5) ----
def spd(n, batch):
    G = torch.randn(n, n, dtype=torch.float64, device="cuda")
    S1 = G @ G.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    del G
    return S1

for n, batch in ((8192, 1), (16384, 1), (32768, 1), (1024, 512), (1024, 4096)):
    try:
        S1 = spd(n, 1)
        A = S1.unsqueeze(0).repeat(batch, 1, 1).contiguous() if batch > 1 else S1.clone().unsqueeze(0)
        def restore():
            A.copy_(S1.unsqueeze(0).expand_as(A))
        t_copy = timed(restore, reps=3)
        t = timed(lambda: (restore(), eng.potrf(A.data_ptr(), n, batch, 0, 0, 0, st)), reps=3) - t_copy
        out(block="potrf", n=n, batch=batch, ms=t, tflops=batch * n ** 3 / 3 / t / 1e9)
        if n <= 16384 or batch > 1:
            Mi, Ui = torch.empty_like(A), torch.empty_like(A)
            t2 = timed(lambda: (restore(), eng.potrf(A.data_ptr(), n, batch, Mi.data_ptr(), Ui.data_ptr(), 0, st)), reps=3) - t_copy
            out(block="potrf+trtri", n=n, batch=batch, ms=t2, tflops=batch * 2 * n ** 3 / 3 / t2 / 1e9)
            del Mi, Ui
        if batch == 1 and n <= 16384:
            t3 = timed(lambda: torch.linalg.cholesky(S1), reps=3)
            out(block="cusolver_potrf", n=n, ms=t3, tflops=n ** 3 / 3 / t3 / 1e9)
        del A, S1
        torch.cuda.empty_cache()
    except Exception as e:  # noqa
        out(block="potrf", n=n, batch=batch, error=str(e)[:200])
        torch.cuda.empty_cache()
eng.close()
