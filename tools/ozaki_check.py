"""GPU check of the sliced-integer tcgen05 GEMM (csrc/ozaki.cu) and of potrf / trtri running on it, against torch FP64.
Run under gpurun.  Prints one line per case; exits non-zero on a failed bound.  AGPC_OZAKI selects the slice count of the
factorisation path (0 = DMMA only, 6, 7)."""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = ctypes.CDLL(os.path.join(ROOT, "autogpc_b200", "libautogpc_b200.so"))
lib.agpc_last_error.restype = ctypes.c_char_p
ctx = ctypes.c_void_p()
assert lib.agpc_create(0, ctypes.byref(ctx)) == 0
P = lambda t: ctypes.c_void_p(t.data_ptr())
D = ctypes.c_double


def chk(rc):
    if rc != 0:
        raise RuntimeError(lib.agpc_last_error(ctx).decode())


def sync():
    torch.cuda.synchronize()


torch.manual_seed(0)
dev = "cuda"
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
ST = ctypes.c_void_p(_stream.cuda_stream)
fails = 0
quick = "--quick" in sys.argv

# ---- 1. the product itself ---------------------------------------------------------------------------------------
cases = [(128, 128, 32, 1, 1.0, 0.0), (128, 128, 64, 1, 1.0, 0.0), (256, 384, 160, 2, -1.0, 1.0), (512, 256, 1024, 3, 0.5, 0.25),
         (1024, 1024, 512, 1, -1.0, 1.0)]
for S in (7, 6):
    for (M, N, K, batch, alpha, beta) in cases:
        A = torch.randn(batch, M, K, dtype=torch.float64, device=dev)
        B = torch.randn(batch, N, K, dtype=torch.float64, device=dev)
        # rows of very different magnitude and a few exact zeros: the per-row scale has to cope
        A *= torch.exp(3.0 * torch.randn(batch, M, 1, dtype=torch.float64, device=dev))
        B[:, 3, :] = 0.0
        C = torch.randn(batch, M, N, dtype=torch.float64, device=dev)
        C0 = C.clone()
        sync()
        chk(lib.agpc_gemm_nt_i8x(ctx, P(A), P(B), P(C), M, N, K, D(alpha), D(beta), batch, S, ST))
        sync()
        ref = alpha * A @ B.transpose(1, 2) + beta * C0
        # error measured against the natural scale of each entry: |alpha| max_k|a_ik| max_k|b_jk| sqrt(K)
        scale = abs(alpha) * A.abs().amax(2, keepdim=True) * B.abs().amax(2, keepdim=True).transpose(1, 2) * K ** 0.5 + 1e-300
        err = ((C - ref).abs() / scale).max().item()
        bound = 3e-14 if S == 7 else 3e-13   # FP64 rounding level of the K-term sum itself for S = 7
        ok = err <= bound
        fails += not ok
        print(f"gemm_nt_i8x S={S} M={M} N={N} K={K} batch={batch} alpha={alpha} beta={beta}: scaled max err {err:.3e} "
              f"(bound {bound:.0e}) {'ok' if ok else 'FAIL'}", flush=True)

# the DMMA building-block bound of tests/test_gpu_parity.py::test_gemm_nt_building_block on N(0,1) data
for S in (7, 6):
    for (M, N, K, batch, alpha, beta) in [(256, 384, 160, 2, -1.0, 1.0), (512, 256, 1024, 3, 0.5, 0.25)]:
        A = torch.randn(batch, M, K, dtype=torch.float64, device=dev)
        B = torch.randn(batch, N, K, dtype=torch.float64, device=dev)
        C = torch.randn(batch, M, N, dtype=torch.float64, device=dev)
        ref = alpha * A @ B.transpose(1, 2) + beta * C
        sync()
        chk(lib.agpc_gemm_nt_i8x(ctx, P(A), P(B), P(C), M, N, K, D(alpha), D(beta), batch, S, ST))
        sync()
        err = (C - ref).abs().max().item() / (ref.abs().max().item() * K ** 0.5)
        print(f"gemm_nt_i8x S={S} N(0,1) M={M} N={N} K={K}: max err / (max|ref| sqrt K) = {err:.3e} (DMMA test bound 1e-13)", flush=True)
        if S == 7:
            fails += err > 1e-13


def timeit(f, reps=5):
    f()
    sync()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        sync()
        best = min(best, e0.elapsed_time(e1))
    return best


ms_arr = (ctypes.c_double * 16)()
wk_arr = (ctypes.c_double * 16)()
ln_arr = (ctypes.c_int64 * 16)()


def prof(f):
    """(gemm ms, slice ms) of one call by the engine's per-launch events"""
    f()
    sync()
    chk(lib.agpc_profile_enable(ctx, 1))
    f()
    chk(lib.agpc_profile_read(ctx, ms_arr, wk_arr, ln_arr))
    chk(lib.agpc_profile_enable(ctx, 0))
    return ms_arr[1] + ms_arr[8] + ms_arr[9], ms_arr[10], ms_arr[2]


# ---- 2. speed of the product: whole call (slicing included) and the MMA kernel alone -----------------------------
shapes = [(8192, 8192, 8192, 1), (6400, 6400, 512, 8), (4096, 4096, 4096, 4)] if not quick else [(2048, 2048, 2048, 1)]
for (M, N, K, batch) in shapes:
    A = torch.randn(batch, M, K, dtype=torch.float64, device=dev)
    B = torch.randn(batch, N, K, dtype=torch.float64, device=dev)
    C = torch.zeros(batch, M, N, dtype=torch.float64, device=dev)
    fl = 2.0 * M * N * K * batch
    t_d = timeit(lambda: chk(lib.agpc_gemm_nt(ctx, P(A), P(B), P(C), M, N, K, D(1.0), D(0.0), batch, ST)))
    line = f"speed M={M} N={N} K={K} batch={batch}: DMMA {t_d:.3f} ms = {fl / t_d / 1e9:.1f} TFLOP/s"
    for S in (7, 6):
        f = lambda: chk(lib.agpc_gemm_nt_i8x(ctx, P(A), P(B), P(C), M, N, K, D(1.0), D(0.0), batch, S, ST))
        t = timeit(f)
        g_ms, s_ms, _ = prof(f)
        line += f" | i8x S={S}: call {t:.3f} ms = {fl / t / 1e9:.1f} TF-eq, MMA kernel {g_ms:.3f} ms = {fl / g_ms / 1e9:.1f} TF-eq, slicing {s_ms:.3f} ms"
    print(line, flush=True)
    del A, B, C

# ---- 3. potrf + trtri on it --------------------------------------------------------------------------------------
oz = os.environ.get("AGPC_OZAKI", "7")
for (n, batch) in [(256, 2), (640, 3), (1024, 2), (1920, 1), (3200, 2)]:
    G = torch.randn(batch, n, n, dtype=torch.float64, device=dev)
    Sm = G @ G.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device=dev)
    A = Sm.clone()
    Mi = torch.zeros_like(Sm)
    Ui = torch.zeros_like(Sm)
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    sync()
    chk(lib.agpc_potrf(ctx, P(A), n, batch, P(Mi), P(Ui), P(info), ST))
    sync()
    Lref = torch.linalg.cholesky(Sm)
    L = torch.tril(A)
    e1 = (L - Lref).abs().max().item()
    Minv_ref = torch.linalg.inv(Lref)
    e2 = (torch.tril(Mi) - Minv_ref).abs().max().item()
    e3 = (torch.triu(Ui) - Minv_ref.transpose(1, 2)).abs().max().item()
    ok = e1 < 1e-11 and e2 < 1e-10 and e3 < 1e-10
    fails += not ok
    print(f"potrf(AGPC_OZAKI={oz}) n={n} batch={batch}: |L-Lref|={e1:.3e} |M-inv(L)|={e2:.3e} |U-inv(L)^T|={e3:.3e} "
          f"info={info.tolist()} {'ok' if ok else 'FAIL'}", flush=True)

if not quick:
    for (n, batch) in [(8192, 1), (6400, 8), (16384, 1)]:
        G = torch.randn(n, n, dtype=torch.float64, device=dev)
        S1 = G @ G.T / n + torch.eye(n, dtype=torch.float64, device=dev)
        del G
        Sm = S1.unsqueeze(0).repeat(batch, 1, 1).contiguous()
        A = Sm.clone()
        Mi = torch.zeros_like(Sm)
        Ui = torch.zeros_like(Sm)

        def run(inv):
            A.copy_(Sm)
            chk(lib.agpc_potrf(ctx, P(A), n, batch, P(Mi) if inv else None, P(Ui) if inv else None, None, ST))

        t_copy = timeit(lambda: A.copy_(Sm))
        t_p = timeit(lambda: run(False)) - t_copy
        t_pi = timeit(lambda: run(True)) - t_copy
        fl = n ** 3 / 3.0 * batch
        print(f"potrf(AGPC_OZAKI={oz}) n={n} batch={batch}: potrf {t_p:.2f} ms = {fl / t_p / 1e9:.1f} TFLOP/s; potrf+trtri {t_pi:.2f} ms = "
              f"{2 * fl / t_pi / 1e9:.1f} TFLOP/s", flush=True)
        # residuals on device: |L L^T - B| / |B| and |M L - I| on one matrix
        L = torch.tril(A[0])
        r1 = (L @ L.T - S1).abs().max().item() / S1.abs().max().item()
        r2 = (torch.tril(Mi[0]) @ L - torch.eye(n, dtype=torch.float64, device=dev)).abs().max().item()
        print(f"    residuals: |LL^T-B|/|B| = {r1:.2e}, |ML-I| = {r2:.2e}", flush=True)
        fails += (r1 > 1e-12) or (r2 > 1e-10)
        del Sm, A, Mi, Ui, S1, L
        torch.cuda.empty_cache()

print("ozaki_check done, %d failures" % fails, flush=True)
sync()
lib.agpc_destroy(ctx)
sys.stdout.flush()
os._exit(1 if fails else 0)
