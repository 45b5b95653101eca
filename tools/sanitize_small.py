"""Small end-to-end exercise of the int8 GEMM (persistent and one-tile forms), potrf / trtri / lauum on it, score_batch and
the on-device optimiser, sized for compute-sanitizer (memcheck / racecheck / synccheck), run under gpurun:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
(compute-sanitizer is closed on the round-2 GPU pool; without it the script is a quick end-to-end check with asserts.)"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autogpc_b200.engine import Engine, Program, LEAF_SE, LEAF_PER  # noqa: E402

e = Engine(0)
L, ctx = e._L, e._ctx
P = lambda t: ctypes.c_void_p(t.data_ptr())
D = ctypes.c_double
torch.manual_seed(0)
for (M, N, K, batch, alpha, beta) in [(128, 128, 64, 1, 1.0, 0.0), (256, 384, 160, 2, -1.0, 1.0), (384, 256, 96, 3, 0.5, 0.25)]:
    A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda")
    C = torch.randn(batch, M, N, dtype=torch.float64, device="cuda")
    ref = alpha * A @ B.transpose(1, 2) + beta * C
    torch.cuda.synchronize()
    e._check(L.agpc_gemm_nt_i8x(ctx, P(A), P(B), P(C), M, N, K, D(alpha), D(beta), batch, 7, None))
    torch.cuda.synchronize()
    err = (C - ref).abs().max().item() / (ref.abs().max().item() * K ** 0.5)
    print("gemm_nt_i8x", M, N, K, batch, "err", err, flush=True)
    assert err < 1e-13
for (n, batch) in [(384, 2), (640, 1)]:
    G = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
    S = G @ G.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device="cuda")
    A = S.clone()
    Mi, Ui = torch.zeros_like(S), torch.zeros_like(S)
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    e._check(L.agpc_potrf(ctx, P(A), n, batch, P(Mi), P(Ui), P(info), None))
    torch.cuda.synchronize()
    Lref = torch.linalg.cholesky(S)
    print("potrf", n, batch, (torch.tril(A) - Lref).abs().max().item(), (torch.tril(Mi) - torch.linalg.inv(Lref)).abs().max().item(), flush=True)
rng = np.random.default_rng(0)
X = rng.normal(size=(300, 3))
y = np.where(np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.2 * rng.normal(size=300) > 0, 1.0, -1.0)
did = e.dataset(X, y)
progs = [Program.build([LEAF_SE], [0], [0], [[0]], 2), Program.build([LEAF_SE, LEAF_PER], [0, 1], [0, 2], [[0], [1]], 5)]
thetas = [np.array([1.0, 0.8]), np.array([0.9, 1.1, 1.0, 2.0, 0.7])]
rows = [np.arange(260, dtype=np.int32), np.arange(20, 300, dtype=np.int32)]
r = e.score_batch(did, progs, thetas, rows, damping=0.0)
print("score_batch", r["nlml"], r["status"], r["sweeps"], flush=True)
o = e.optimize_batch(did, progs, thetas, [[1, 1], [1, 1, 1, 2, 1]], [[0, 0], [0, 0, 0, 0.5, 0]], [[1, 1], [1, 1, 1, 6.0, 1]],
                     rows, damping=0.0, max_evals=8)
print("optimize_batch", o["nlml"], o["evals"], o["task"], flush=True)
assert np.all(o["nlml"] <= r["nlml"] + 1e-9)
e.dataset_free(did)
e.close()
print("sanitize_small done", flush=True)
