"""GPU check of the TMA+DMMA GEMM and the blocked potrf/trtri against torch FP64 (run under gpurun)."""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autogpc_b200", "libautogpc_b200.so"))
lib.agpc_last_error.restype = ctypes.c_char_p
ctx = ctypes.c_void_p()
assert lib.agpc_create(0, ctypes.byref(ctx)) == 0
P = lambda t: ctypes.c_void_p(t.data_ptr())
def chk(rc):
    if rc != 0: raise RuntimeError(lib.agpc_last_error(ctx).decode())
def sync(): torch.cuda.synchronize()
torch.manual_seed(0)
dev = "cuda"
# one explicit stream for torch AND the engine (the engine's own stream is non-blocking: no implicit ordering with torch)
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
ST = ctypes.c_void_p(_stream.cuda_stream)
# ---- gemm
for (M, N, K, batch, alpha, beta) in [(128, 128, 16, 1, 1.0, 0.0), (256, 384, 160, 2, -1.0, 1.0), (1024, 512, 1024, 3, 0.5, 0.25)]:
    A = torch.randn(batch, M, K, dtype=torch.float64, device=dev); B = torch.randn(batch, N, K, dtype=torch.float64, device=dev)
    C = torch.randn(batch, M, N, dtype=torch.float64, device=dev); C0 = C.clone()
    chk(lib.agpc_gemm_nt(ctx, P(A), P(B), P(C), M, N, K, ctypes.c_double(alpha), ctypes.c_double(beta), batch, ST)); sync()
    ref = alpha * A @ B.transpose(1, 2) + beta * C0
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    print(f"gemm_nt M={M} N={N} K={K} batch={batch}: rel max err {err:.3e}", flush=True)
    assert err < 1e-13
# ---- potrf + trtri
for (n, batch) in [(128, 1), (256, 2), (640, 3), (1024, 2), (1920, 1)]:
    G = torch.randn(batch, n, n, dtype=torch.float64, device=dev)
    S = G @ G.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device=dev)
    A = S.clone(); Mi = torch.zeros_like(S); Ui = torch.zeros_like(S); info = torch.zeros(batch, dtype=torch.int32, device=dev)
    chk(lib.agpc_potrf(ctx, P(A), n, batch, P(Mi), P(Ui), P(info), ST)); sync()
    Lref = torch.linalg.cholesky(S)
    L = torch.tril(A)
    e1 = (L - Lref).abs().max().item()
    Minv_ref = torch.linalg.inv(Lref)
    e2 = (torch.tril(Mi) - Minv_ref).abs().max().item()
    e3 = (torch.triu(Ui) - Minv_ref.transpose(1, 2)).abs().max().item()
    print(f"potrf n={n} batch={batch}: |L-Lref|={e1:.3e} |M-inv(L)|={e2:.3e} |U-inv(L)^T|={e3:.3e} info={info.tolist()}", flush=True)
    assert e1 < 1e-11 and e2 < 1e-10 and e3 < 1e-10
# not-PD detection
n = 256
S = torch.eye(n, dtype=torch.float64, device=dev); S[200, 200] = -1.0
info = torch.zeros(1, dtype=torch.int32, device=dev)
chk(lib.agpc_potrf(ctx, P(S), n, 1, None, None, P(info), ST)); sync()
print("not-pd info:", info.tolist()); assert info.item() == 2
# ---- timing
def timeit(f, reps=3):
    f(); sync(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); sync(); best = min(best, e0.elapsed_time(e1))
    return best
# the engine launches on its own stream: time with wall-clock + sync instead
def timeit_wall(f, reps=3):
    f(); sync(); best = 1e9
    for _ in range(reps):
        sync(); t = time.perf_counter(); f(); sync(); best = min(best, (time.perf_counter() - t) * 1e3)
    return best
for n in (4096, 8192):
    A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev); C = torch.zeros(n, n, dtype=torch.float64, device=dev)
    ms = timeit_wall(lambda: chk(lib.agpc_gemm_nt(ctx, P(A), P(B), P(C), n, n, n, ctypes.c_double(1.0), ctypes.c_double(0.0), 1, ST)))
    print(f"gemm_nt {n}^3: {ms:.2f} ms  {2*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
    ms = timeit(lambda: torch.matmul(A, B.T, out=C))
    print(f"cublas  {n}^3: {ms:.2f} ms  {2*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
# diagonal-block kernel alone (n = 128: potrf == one potf2_inv launch)
for batch in (1, 40, 148):
    S = (torch.eye(128, dtype=torch.float64, device=dev) * 3 + 0.01).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    A = S.clone(); Mi = torch.empty_like(S); Ui = torch.empty_like(S)
    def run_p():
        for _ in range(20):
            chk(lib.agpc_potrf(ctx, P(A), 128, batch, P(Mi), P(Ui), None, ST))
    ms = timeit_wall(run_p) / 20
    print(f"potf2_inv alone batch={batch}: {ms*1e3:.1f} us per launch (incl. launch overhead)", flush=True)
for n, batch in ((8192, 1), (6400, 1), (6400, 4), (1024, 64)):
    G = torch.randn(n, n, dtype=torch.float64, device=dev)
    S1 = G @ G.T / n + torch.eye(n, dtype=torch.float64, device=dev)
    S = S1.unsqueeze(0).repeat(batch, 1, 1).contiguous()
    A = S.clone()
    def run():
        A.copy_(S)
        chk(lib.agpc_potrf(ctx, P(A), n, batch, None, None, None, ST))
    def copy_only():
        A.copy_(S)
    ms = timeit_wall(run) - timeit_wall(copy_only)
    print(f"potrf n={n} batch={batch}: {ms:.2f} ms  {batch*n**3/3/ms/1e9:.2f} TFLOP/s", flush=True)
    Mi = torch.empty_like(S); Ui = torch.empty_like(S)
    def run2():
        A.copy_(S)
        chk(lib.agpc_potrf(ctx, P(A), n, batch, P(Mi), P(Ui), None, ST))
    ms2 = timeit_wall(run2) - timeit_wall(copy_only)
    print(f"potrf+trtri n={n} batch={batch}: {ms2:.2f} ms  {batch*2*n**3/3/ms2/1e9:.2f} TFLOP/s", flush=True)
    if batch == 1:
        ms3 = timeit(lambda: torch.linalg.cholesky(S1))
        print(f"cusolver potrf n={n}: {ms3:.2f} ms {n**3/3/ms3/1e9:.2f} TFLOP/s", flush=True)
    del S, A, Mi, Ui
print("launches:", lib.agpc_launch_count(ctx))
