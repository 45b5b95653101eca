"""CPU prototype (NumPy, no GPU) of the sliced-integer FP64 GEMM behind csrc/ozaki.cu: how many 8-bit signed-digit
slices does the emulation need before potrf / trtri / lauum of B = I + S^1/2 K S^1/2 keep this engine's parity bar
(NLML 1e-9, gradients 1e-7 relative)?

Scheme -- exactly what the slicing kernel does on the device:
  * one power-of-two scale per operand row: e_i with |a_ik| 2^-e_i < 1/4 for every k of the row;
  * X_ik = rint(a_ik 2^(8 s - e_i)) as a 64-bit integer (exact for the entries within 2^-(8s-53) of the row scale),
    cut into s signed bytes d_0 (most significant) ... d_(s-1) by the carry-propagating split
    d = ((X + 128) & 255) - 128, X = (X - d) >> 8   -- round-to-nearest digits, so the truncation error of the product
    is unbiased (the round-1 prototype truncated toward zero: 7 bits per slice and a systematic bias);
  * A B^T = 2^(ea_i + eb_j) sum_{p+q < s} 2^(-8 (p + q + 2)) D_p^A (D_q^B)^T: s (s + 1) / 2 int8 GEMMs whose products of
    equal weight p + q share one int32 accumulator (|.| <= 2^14 K (p+q+1): exact for K < 18 k), then one FP64 Horner
    pass over the s accumulators.
Integer products are formed in float64 on integer-valued arrays (exact), i.e. the arithmetic tcgen05 kind::i8 performs.

The three GEMM-shaped phases of one posterior evaluation are run through that GEMM -- right-looking blocked Cholesky
(panel factorisation and solve stay native FP64, as in csrc/chol.cu), L^-1 by recursive doubling, B^-1 = L^-T L^-1 --
and compared with LAPACK.   Usage: python tools/probe/ozaki_prototype.py [n]
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gpc_oracle as orc  # noqa: E402
from tests.util import random_theta, synth  # noqa: E402


def slices(A, s):
    """Row-scaled signed-digit slicing: (list of s integer-valued float arrays in [-128, 127], exponent vector e)."""
    amax = np.max(np.abs(A), axis=1)
    _, ex = np.frexp(amax)                      # amax = m 2^ex, m in [0.5, 1)
    e = np.where(amax > 0, ex + 2, 0).astype(np.int64)   # |a| 2^-e < 1/4
    X = np.rint(np.ldexp(A, (8 * s - e)[:, None].astype(np.int32))).astype(np.int64)
    out = []
    for _ in range(s):
        d = ((X + 128) & 255) - 128
        out.append(d.astype(np.float64))
        X = (X - d) >> 8
    assert np.all(X == 0), "top digit overflow"
    return out[::-1], e


def gemm_nt_emulated(A, B, s):
    """A @ B.T with s slices per operand, integer products for p + q < s, Horner accumulation from the smallest weight."""
    DA, ea = slices(A, s)
    DB, eb = slices(B, s)
    acc = np.zeros((A.shape[0], B.shape[0]))
    for d in range(s - 1, -1, -1):
        g = np.zeros_like(acc)
        for p in range(d + 1):
            g += DA[p] @ DB[d - p].T            # integer-valued, exact (int32 on the tensor core)
        acc = acc * 2.0 ** -8 + g
    return np.ldexp(acc, (ea[:, None] + eb[None, :] - 16).astype(np.int32))


def blocked_cholesky(A, nb, gemm, outer=4):
    """Two-level right-looking Cholesky as chol.cu: panels of nb inside outer blocks (native FP64), K = outer*nb
    trailing updates through `gemm`."""
    A = A.copy()
    n = A.shape[0]
    for k0 in range(0, n, nb * outer):
        e0 = min(k0 + nb * outer, n)
        for k in range(k0, e0, nb):
            e = min(k + nb, n)
            if k > k0:
                A[k:, k:e] -= A[k:, k0:k] @ A[k:e, k0:k].T
            A[k:e, k:e] = sla.cholesky(A[k:e, k:e], lower=True)
            if e < n:
                A[e:, k:e] = sla.solve_triangular(A[k:e, k:e], A[e:, k:e].T, lower=True).T
        if e0 < n:
            A[e0:, e0:] -= gemm(A[e0:, k0:e0], A[e0:, k0:e0])
    return np.tril(A)


def trtri_doubling(L, nb, gemm):
    """M = L^-1 by recursive doubling (chol.cu trtri_blocked): M21 = -M22 (L21 M11)."""
    n = L.shape[0]
    M = np.zeros_like(L)
    for k in range(0, n, nb):
        e = min(k + nb, n)
        M[k:e, k:e] = sla.solve_triangular(L[k:e, k:e], np.eye(e - k), lower=True)
    h = nb
    while h < n:
        for g0 in range(0, n, 2 * h):
            m = min(g0 + h, n)
            e = min(g0 + 2 * h, n)
            if m >= e:
                continue
            T = gemm(M[g0:m, g0:m].T, L[m:e, g0:m])           # U11 L21^T  (h1 x h2)
            M[m:e, g0:m] = -gemm(M[m:e, m:e], T)               # M22 T^T
        h *= 2
    return M


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
    Binv_ref = sla.cho_solve((L_ref, True), np.eye(n))
    ld_ref = 2 * np.sum(np.log(np.diag(L_ref)))
    print("n = %d, cond(B) = %.2e, EP sweeps %d" % (n, np.linalg.cond(Bm), ep.sweeps))
    nat = lambda a, b: a @ b.T
    for name, gemm in [("native FP64", nat)] + [("s = %d (%2d int8 GEMMs)" % (s, s * (s + 1) // 2),
                                                 (lambda s_: (lambda a, b: gemm_nt_emulated(a, b, s_)))(s))
                                                for s in range(4, 9)]:
        L = blocked_cholesky(Bm, 128, gemm)
        M = trtri_doubling(L, 128, gemm)
        Binv = gemm(M.T, M.T)
        ld = 2 * np.sum(np.log(np.diag(L)))
        print("%-24s max|L-Lref|/max|L| = %.2e  |logdet diff|/|logdet| = %.2e  max|ML-I| = %.2e  max|Binv-ref|/max|Binv| = %.2e"
              % (name, np.abs(L - L_ref).max() / np.abs(L_ref).max(), abs(ld - ld_ref) / abs(ld_ref),
                 np.abs(M @ L - np.eye(n)).max(), np.abs(Binv - Binv_ref).max() / np.abs(Binv_ref).max()))
