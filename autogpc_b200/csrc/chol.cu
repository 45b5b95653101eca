// Blocked FP64 Cholesky / triangular inverse / B^-1 for B = I + S^1/2 K S^1/2 (batched, lock-step).
//
// Stands in for GPy util/linalg.py jitchol (dpotrf), dtrtrs / dtrtri and pdinv's dpotri + symmetrify, as reached
// from EP.expectation_propagation and EP.inference for gpckernel.py:245-246.
//
//   potrf  : right-looking, 128-wide panels.  Per panel k: potf2_inv (one CTA per matrix) factors the diagonal
//            block and inverts it; the panel solve P <- P W_k^T and the trailing update C -= P P^T both run on the
//            TMA + DMMA tile kernel (gemm_nt.cu).  W_k = L_kk^-1 doubles as level 0 of trtri.
//   trtri  : level-synchronous recursive doubling, M21 = -M22 L21 M11, two batched triangular-aware GEMMs per level;
//            both M = L^-1 and U = M^T are kept so every product is in K-contiguous ("NT") form.
//   lauum  : B^-1 = U U^T, lower tiles only, K range clipped at the tile row.
#include "common.cuh"

namespace agpc {

constexpr int POTF2_THREADS = 512;
constexpr int POTF2_LD = TILE + 1;
constexpr int POTF2_SMEM = TILE * POTF2_LD * 8;

// One CTA per (batch item): S (smem) holds the lower triangle of the diagonal block (-> L) and, transposed into
// the strict upper part shifted by one column, the running inverse W (W[i][c] at S[c][i+1]).
__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, double* __restrict__ M, double* __restrict__ U, long long ld,
                 long long batch_stride, int diag_block, const int* __restrict__ active, int* __restrict__ info) {
  const int b = blockIdx.x;
  if (active != nullptr && active[b] == 0) return;
  extern __shared__ double S[];
  const long long base = (long long)b * batch_stride + (long long)diag_block * TILE * ld + (long long)diag_block * TILE;
  double* Ab = A + base;
  const int tid = threadIdx.x;
  // load lower triangle; W = I
  for (int idx = tid; idx < TILE * TILE; idx += POTF2_THREADS) {
    const int i = idx / TILE, j = idx % TILE;
    if (j <= i) S[i * POTF2_LD + j] = Ab[(long long)i * ld + j];
  }
  __syncthreads();
  for (int idx = tid; idx < TILE * TILE; idx += POTF2_THREADS) {
    const int c = idx / TILE, i = idx % TILE;
    if (i >= c) S[c * POTF2_LD + i + 1] = (i == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  bool bad = false;
  for (int j = 0; j < TILE; ++j) {
    __syncthreads();  // updates of step j-1 are complete
    double d = S[j * POTF2_LD + j];
    if (!(d > 0.0) || !isfinite(d)) {
      bad = true;
      d = 1.0;
    }
    const double piv = sqrt(d);
    const double rinv = 1.0 / piv;
    // scale column j of A (rows > j) and row j of W (cols <= j); nobody reads those entries before the barrier
    if (tid < TILE) {
      const int i = tid;
      if (i > j) S[i * POTF2_LD + j] *= rinv;
    } else if (tid < 2 * TILE) {
      const int c = tid - TILE;
      if (c <= j) S[c * POTF2_LD + j + 1] *= rinv;
    }
    __syncthreads();
    if (tid == 0) S[j * POTF2_LD + j] = piv;  // a_jj is not read again during this step
    const int m = TILE - 1 - j;
    if (m > 0) {
      // A part: m x m (k fastest), W part: (j+1) x m (i fastest)
      const int nA = m * m, nW = (j + 1) * m;
      for (int idx = tid; idx < nA + nW; idx += POTF2_THREADS) {
        if (idx < nA) {
          const int i = j + 1 + idx / m, k = j + 1 + idx % m;
          if (k <= i) S[i * POTF2_LD + k] -= S[i * POTF2_LD + j] * S[k * POTF2_LD + j];
        } else {
          const int r = idx - nA;
          const int c = r / m, i = j + 1 + r % m;
          S[c * POTF2_LD + i + 1] -= S[i * POTF2_LD + j] * S[c * POTF2_LD + j + 1];
        }
      }
    }
  }
  __syncthreads();
  // write back: L (lower) to A; W (lower, zeros above) to M; W^T (upper, zeros below) to U
  double* Mb = M + base;
  double* Ub = U + base;
  for (int idx = tid; idx < TILE * TILE; idx += POTF2_THREADS) {
    const int i = idx / TILE, j = idx % TILE;
    if (j <= i) Ab[(long long)i * ld + j] = S[i * POTF2_LD + j];
    const double w_ij = (j <= i) ? S[j * POTF2_LD + i + 1] : 0.0;  // W[i][j]
    const double w_ji = (i <= j) ? S[i * POTF2_LD + j + 1] : 0.0;  // W[j][i] = U[i][j]
    Mb[(long long)i * ld + j] = w_ij;
    Ub[(long long)i * ld + j] = w_ji;
  }
  if (bad && tid == 0 && info != nullptr) info[b] = AGPC_ST_NOT_PD;
}

cudaError_t potf2_inv_launch(const Launch& L, double* A, double* M, double* U, long long ld, long long batch_stride,
                             int diag_block, int batch, const int* active, int* info) {
  static bool configured = false;
  if (!configured) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    configured = true;
  }
  potf2_inv_kernel<<<batch, POTF2_THREADS, POTF2_SMEM, L.stream>>>(A, M, U, ld, batch_stride, diag_block, active, info);
  L.bump();
  return cudaGetLastError();
}

// ---- drivers -----------------------------------------------------------------------------------------------
cudaError_t potrf_blocked(const Launch& L, const MatSet& ms, const int* active, int* info) {
  const int nblk = ms.np / TILE;
  for (int k = 0; k < nblk; ++k) {
    AGPC_CUDA_TRY(potf2_inv_launch(L, ms.A, ms.M, ms.U, ms.np, ms.stride, k, ms.batch, active, info));
    const int below = nblk - k - 1;
    if (below == 0) break;
    GemmArgs t{};
    t.C = ms.A; t.Ct = nullptr; t.ldc = ms.np; t.c_batch = ms.stride;
    t.a_row0 = (k + 1) * TILE; t.a_col0 = k * TILE;      // P_old (in place)
    t.b_row0 = k * TILE; t.b_col0 = k * TILE;            // W_k from M's diagonal block
    t.c_row0 = (k + 1) * TILE; t.c_col0 = k * TILE;
    t.m_tiles = below; t.n_tiles = 1; t.k_tiles = TILE / GEMM_BK;
    t.k_rule = KRULE_FULL; t.lower_only = 0; t.alpha = 1.0; t.beta = 0.0; t.active = active;
    AGPC_CUDA_TRY(gemm_nt_launch(L, ms.mapA, ms.mapM, t, below, 1, ms.batch));
    GemmArgs s{};
    s.C = ms.A; s.Ct = nullptr; s.ldc = ms.np; s.c_batch = ms.stride;
    s.a_row0 = (k + 1) * TILE; s.a_col0 = k * TILE;
    s.b_row0 = (k + 1) * TILE; s.b_col0 = k * TILE;
    s.c_row0 = (k + 1) * TILE; s.c_col0 = (k + 1) * TILE;
    s.m_tiles = below; s.n_tiles = below; s.k_tiles = TILE / GEMM_BK;
    s.k_rule = KRULE_FULL; s.lower_only = 1; s.alpha = -1.0; s.beta = 1.0; s.active = active;
    AGPC_CUDA_TRY(gemm_nt_launch(L, ms.mapA, ms.mapA, s, below * (below + 1) / 2, 1, ms.batch));
  }
  return cudaSuccess;
}

// M = L^-1 (lower), U = M^T (upper); diagonal 128-blocks already written by potf2_inv.
cudaError_t trtri_blocked(const Launch& L, const MatSet& ms, const int* active) {
  const int nblk = ms.np / TILE;
  for (int h = 1; h < nblk; h *= 2) {
    const int G = 2 * h;
    const int groups = (nblk + G - 1) / G;
    // step 1: T[left rows, right cols] = U11 * L21^T
    GemmArgs a{};
    a.C = ms.T; a.Ct = nullptr; a.ldc = ms.np; a.c_batch = ms.stride;
    a.a_row0 = 0; a.a_col0 = 0;                 // U11
    a.b_row0 = h * TILE; a.b_col0 = 0;          // L21
    a.c_row0 = 0; a.c_col0 = h * TILE;
    a.m_tiles = h; a.n_tiles = 0; a.k_tiles = h * (TILE / GEMM_BK);
    a.m_kind = EXT_FIXED; a.n_kind = EXT_H2; a.k_kind = EXT_FIXED;
    a.grp_blocks = G; a.half_blocks = h; a.nblk = nblk;
    a.k_rule = KRULE_BEGIN_AT_ROW; a.lower_only = 0; a.alpha = 1.0; a.beta = 0.0; a.active = active;
    AGPC_CUDA_TRY(gemm_nt_launch(L, ms.mapU, ms.mapA, a, h * h, groups, ms.batch));
    // step 2: M21 = -M22 * T^T ; U12 = M21^T
    GemmArgs c{};
    c.C = ms.M; c.Ct = ms.U; c.ldc = ms.np; c.c_batch = ms.stride;
    c.a_row0 = h * TILE; c.a_col0 = h * TILE;   // M22
    c.b_row0 = 0; c.b_col0 = h * TILE;          // T (h1 x h2)
    c.c_row0 = h * TILE; c.c_col0 = 0;
    c.ct_row0 = 0; c.ct_col0 = h * TILE;
    c.m_tiles = 0; c.n_tiles = h; c.k_tiles = 0;
    c.m_kind = EXT_H2; c.n_kind = EXT_FIXED; c.k_kind = EXT_H2;
    c.grp_blocks = G; c.half_blocks = h; c.nblk = nblk;
    c.k_rule = KRULE_END_AT_ROW; c.lower_only = 0; c.alpha = -1.0; c.beta = 0.0; c.active = active;
    AGPC_CUDA_TRY(gemm_nt_launch(L, ms.mapM, ms.mapT, c, h * h, groups, ms.batch));
  }
  return cudaSuccess;
}

// dst (lower tiles) = U U^T = B^-1
cudaError_t lauum_blocked(const Launch& L, const MatSet& ms, double* dst, const int* active) {
  const int nblk = ms.np / TILE;
  GemmArgs a{};
  a.C = dst; a.Ct = nullptr; a.ldc = ms.np; a.c_batch = ms.stride;
  a.m_tiles = nblk; a.n_tiles = nblk; a.k_tiles = nblk * (TILE / GEMM_BK);
  a.k_rule = KRULE_BEGIN_AT_ROW; a.lower_only = 1; a.alpha = 1.0; a.beta = 0.0; a.active = active;
  return gemm_nt_launch(L, ms.mapU, ms.mapU, a, nblk * (nblk + 1) / 2, 1, ms.batch);
}

}  // namespace agpc
