"""Runs one blocked potrf(+trtri) so ncu can list per-kernel durations."""
import ctypes, os, sys, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, "autogpc_b200", "libautogpc_b200.so"))
ctx = ctypes.c_void_p(); assert lib.agpc_create(0, ctypes.byref(ctx)) == 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
G = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
S = G @ G.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device="cuda")
Mi = torch.empty_like(S); Ui = torch.empty_like(S)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    A = S.clone()
    torch.cuda.synchronize()   # the engine's stream is not ordered with torch's
    assert lib.agpc_potrf(ctx, P(A), n, batch, P(Mi), P(Ui), None, None) == 0
    torch.cuda.synchronize()
print("ok")
