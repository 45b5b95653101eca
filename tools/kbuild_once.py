"""Runs the K build (SE0 program, n = 8192) and one 128 x 128 potf2_inv a few times so ncu can capture them."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autogpc_b200.engine import Engine, Program, LEAF_SE
eng = Engine(0)
n = 8192
X = torch.rand(n, 8, dtype=torch.float64, device="cuda") * 6 - 3
K = torch.empty(n, n, dtype=torch.float64, device="cuda")
p = Program.build([LEAF_SE], [0], [0], [[0]], 2)
torch.cuda.synchronize()
for _ in range(3):
    eng.build_K(p, np.array([1.1, 0.9]), X.data_ptr(), n, 8, K.data_ptr(), n)
S = (torch.eye(128, dtype=torch.float64, device="cuda") * 3 + 0.01).unsqueeze(0).repeat(4, 1, 1).contiguous()
Mi, Ui = torch.empty_like(S), torch.empty_like(S)
for _ in range(3):
    A = S.clone()
    torch.cuda.synchronize()
    eng.potrf(A.data_ptr(), 128, 4, Mi.data_ptr(), Ui.data_ptr(), 0)
torch.cuda.synchronize()
print("ok")
