"""CPU prototype (NumPy, no GPU): how many 7-bit integer slices does an Ozaki-style emulation of the FP64 tile GEMM need
before the blocked Cholesky of B = I + S^1/2 K S^1/2 keeps this engine's parity bar?

Scheme (error-free operand slicing, one power-of-two scale per row): a row of A is scaled into (-1, 1) by 2^-e_i and cut
into `s` signed slices of BETA = 7 bits (int8 range); A ~= sum_p 2^(e_i - BETA (p+1)) S_p.  A B^T is then the sum of the
INTEGER products S_p^A (S_q^B)^T with p + q < s (exact in int32 for K <= 133 k), rescaled and accumulated in FP64 --
s (s + 1) / 2 int8 GEMMs per DGEMM.  Here the integer products are formed in float64 on integer-valued arrays (exact),
so the arithmetic is the one tcgen05 kind::i8 would perform.

The trailing updates of a right-looking blocked Cholesky (128-column panels; panel factorisation and solve in native
FP64, as the design keeps them) are run through that GEMM and compared with LAPACK: relative error of L, of log det B and
of a solve -- for s = 4 ... 8.  Usage: python tools/probe/ozaki_prototype.py [n]
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gpc_oracle as orc  # noqa: E402
from tests.util import random_theta, synth  # noqa: E402

BETA = 7


def slices(A, s):
    """Row-scaled error-free slicing: returns (list of integer-valued float arrays, exponent vector e)."""
    amax = np.max(np.abs(A), axis=1)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)
    R = A * np.exp2(-e)[:, None]                 # |R| < 1/2: exact (power-of-two scaling)
    out = []
    for _ in range(s):
        R = R * (1 << BETA)
        S = np.trunc(R)                          # |S| <= 2^(BETA-1) = 64 fits int8; remainder keeps the sign
        out.append(S)
        R = R - S                                # exact
    return out, e


def gemm_nt_emulated(A, B, s):
    """A @ B.T with s slices per operand, integer products for p + q < s."""
    SA, ea = slices(A, s)
    SB, eb = slices(B, s)
    C = np.zeros((A.shape[0], B.shape[0]))
    for d in range(s):                           # d = p + q, smallest contributions first
        dd = s - 1 - d
        acc = np.zeros_like(C)
        for p in range(dd + 1):
            acc += SA[p] @ SB[dd - p].T          # integer-valued, exact in FP64 (and in int32 on the tensor core)
        C += acc * 2.0 ** (-BETA * (dd + 2))
    return C * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def blocked_cholesky(A, nb, gemm):
    A = A.copy()
    n = A.shape[0]
    for k in range(0, n, nb):
        e = min(k + nb, n)
        A[k:e, k:e] = sla.cholesky(A[k:e, k:e], lower=True)
        if e < n:
            A[e:, k:e] = sla.solve_triangular(A[k:e, k:e], A[e:, k:e].T, lower=True).T
            A[e:, e:] -= gemm(A[e:, k:e], A[e:, k:e])
    return np.tril(A)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    X, y = synth(n, 4, 0)
    prog = orc.KernelProgram.from_tree(("+", ("*", ("SE", 0), ("SE", 1)), ("PER", 2), ("SE", 3)))
    theta = random_theta(prog, np.random.default_rng(1))
    K = orc.kern_K(prog, theta, X)
    ep = orc.ep_parallel(K, y, damping=0.0)
    sq = np.sqrt(ep.tau_t)
    Bm = np.eye(n) + sq[:, None] * K * sq[None, :]
    L_ref = sla.cholesky(Bm, lower=True)
    rhs = np.random.default_rng(2).standard_normal(n)
    x_ref = sla.cho_solve((L_ref, True), rhs)
    print("n = %d, cond(B) = %.2e, EP sweeps %d" % (n, np.linalg.cond(Bm), ep.sweeps))
    L_nat = blocked_cholesky(Bm, 128, lambda a, b: a @ b.T)
    print("native FP64 blocked:   |L - L_ref| / |L_ref| = %.2e" % (np.abs(L_nat - L_ref).max() / np.abs(L_ref).max()))
    for s in range(4, 9):
        L = blocked_cholesky(Bm, 128, lambda a, b: gemm_nt_emulated(a, b, s))
        x = sla.cho_solve((L, True), rhs)
        ld, ld_ref = 2 * np.sum(np.log(np.diag(L))), 2 * np.sum(np.log(np.diag(L_ref)))
        print("s = %d (%2d int8 GEMMs): max|L - L_ref| / max|L_ref| = %.2e   |logdet diff| / |logdet| = %.2e   "
              "solve rel err = %.2e" % (s, s * (s + 1) // 2, np.abs(L - L_ref).max() / np.abs(L_ref).max(),
                                        abs(ld - ld_ref) / abs(ld_ref), np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)))
