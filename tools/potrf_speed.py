"""Timing of potrf / potrf + trtri through the C ABI for a few (n, batch) shapes (run under gpurun)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "autogpc_b200", "libautogpc_b200.so"))
lib.agpc_last_error.restype = ctypes.c_char_p
ctx = ctypes.c_void_p()
assert lib.agpc_create(0, ctypes.byref(ctx)) == 0
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ST = ctypes.c_void_p(st.cuda_stream)


def chk(rc):
    if rc != 0:
        raise RuntimeError(lib.agpc_last_error(ctx).decode())


def timeit(f, reps=5):
    f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


shapes = [(4096, 1), (8192, 1), (16384, 1), (6400, 8), (6400, 40), (1024, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:] if "x" in a]
for (n, batch) in shapes:
    G = torch.randn(n, n, dtype=torch.float64, device="cuda")
    S1 = G @ G.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    del G
    Sm = S1.unsqueeze(0).expand(batch, n, n)
    A = Sm.clone(memory_format=torch.contiguous_format)
    Mi = torch.zeros_like(A)
    Ui = torch.zeros_like(A)

    def run(inv):
        A.copy_(Sm)
        chk(lib.agpc_potrf(ctx, P(A), n, batch, P(Mi) if inv else None, P(Ui) if inv else None, None, ST))

    t_copy = timeit(lambda: A.copy_(Sm))
    t_p = timeit(lambda: run(False)) - t_copy
    t_pi = timeit(lambda: run(True)) - t_copy
    fl = n ** 3 / 3.0 * batch
    Lf = torch.tril(A[0])
    res = (Lf @ Lf.T - S1).abs().max().item() / S1.abs().max().item()
    print(f"AGPC_OZ_PERSIST={os.environ.get('AGPC_OZ_PERSIST', '')} OUTER={os.environ.get('AGPC_POTRF_OUTER', '')} |LL^T-B|/|B|={res:.1e} n={n} batch={batch}: potrf {t_p:.2f} ms = {fl / t_p / 1e9:.1f} TFLOP/s; "
          f"potrf+trtri {t_pi:.2f} ms = {2 * fl / t_pi / 1e9:.1f} TFLOP/s", flush=True)
    if "--classes" in sys.argv:
        ms_arr = (ctypes.c_double * 16)()
        wk_arr = (ctypes.c_double * 16)()
        ln_arr = (ctypes.c_int64 * 16)()
        names = ["kbuild", "gemm_potrf", "potf2", "matvec", "ep", "grad", "formB", "misc", "gemm_trtri", "gemm_lauum", "slice", "oz_potrf", "oz_trtri", "oz_lauum"]
        for inv in (False, True):
            run(inv)
            torch.cuda.synchronize()
            chk(lib.agpc_profile_enable(ctx, 1))
            run(inv)
            chk(lib.agpc_profile_read(ctx, ms_arr, wk_arr, ln_arr))
            chk(lib.agpc_profile_enable(ctx, 0))
            print("    classes (ms, launches)", "potrf+trtri" if inv else "potrf", {names[i]: (round(ms_arr[i], 2), ln_arr[i]) for i in range(14) if ln_arr[i]}, flush=True)
    del Sm, A, Mi, Ui, S1, Lf
    torch.cuda.empty_cache()
torch.cuda.synchronize()
lib.agpc_destroy(ctx)
sys.stdout.flush()
os._exit(0)
