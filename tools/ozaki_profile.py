"""Two launches of the int8 tensor-core GEMM for an ncu --set full capture: the K = 512 trailing-update shape of the
headline bench (6400 x 6400, batch 2) and a square 4096^3 product, slicing included.  Run under gpurun."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autogpc_b200.engine import Engine
eng = Engine(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
torch.manual_seed(0)
for (M, N, K, batch) in [(6400, 6400, 512, 2), (4096, 4096, 4096, 1)]:
    A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda")
    C = torch.randn(batch, M, N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        eng.gemm_nt_i8x(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, -1.0, 1.0, batch, 7, st.cuda_stream)
    torch.cuda.synchronize()
    print("ran", M, N, K, batch, flush=True)
eng.close()
