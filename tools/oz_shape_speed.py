"""MMA-kernel rate of the sliced-integer GEMM (agpc_gemm_nt_i8x, per-launch events of the engine) over the tile shapes the
factorisations launch: short K, narrow N.  Run under gpurun.  usage: oz_shape_speed.py MxNxKxbatch ..."""
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
D = ctypes.c_double
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ST = ctypes.c_void_p(st.cuda_stream)
ms_arr = (ctypes.c_double * 16)()
wk_arr = (ctypes.c_double * 16)()
ln_arr = (ctypes.c_int64 * 16)()


def chk(rc):
    if rc != 0:
        raise RuntimeError(lib.agpc_last_error(ctx).decode())


shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:] if "x" in a] or [
    (6400, 256, 256, 40), (6400, 256, 512, 40), (6400, 256, 2048, 40), (6400, 512, 512, 40), (6400, 6400, 128, 8),
    (6400, 6400, 256, 8), (6400, 6400, 512, 8), (6400, 6400, 2048, 8)]
for (M, N, K, batch) in shapes:
    A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda")
    C = torch.zeros(batch, M, N, dtype=torch.float64, device="cuda")
    f = lambda: chk(lib.agpc_gemm_nt_i8x(ctx, P(A), P(B), P(C), M, N, K, D(-1.0), D(1.0), batch, 7, ST))
    f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        chk(lib.agpc_profile_enable(ctx, 1))
        f()
        chk(lib.agpc_profile_read(ctx, ms_arr, wk_arr, ln_arr))
        chk(lib.agpc_profile_enable(ctx, 0))
        best = min(best, ms_arr[1])
    fl = 2.0 * M * N * K * batch
    tiles = (M // 128) * (N // 64) * batch
    cyc = best * 1e-3 * 1.965e9 * 148 / tiles
    print(f"AGPC_OZ_PERSIST={os.environ.get('AGPC_OZ_PERSIST', 'auto')} M={M} N={N} K={K} batch={batch}: MMA kernel {best:.3f} ms = {fl / best / 1e9:.1f} TF-eq; "
          f"{cyc:.0f} SM-cycles per 128x64 tile = {cyc / (K / 32):.0f} per K step (28 MMAs)", flush=True)
    del A, B, C
torch.cuda.synchronize()
lib.agpc_destroy(ctx)
sys.stdout.flush()
os._exit(0)
