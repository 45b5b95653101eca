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
constexpr int POTF2_SMEM = (TILE * POTF2_LD + 2 * TILE + TILE + 16) * 8;

// One CTA per batch item factors the 128 x 128 diagonal block AND inverts its factor, entirely in registers.
//
// The square S (thread (ty, tx) of a 32 x 16 grid owns rows ty + 32a, columns tx + 16b: cyclic, so every thread
// stays busy as the active region shrinks) holds the lower triangle of A (-> L) and, transposed into its strict
// upper triangle, the strict lower triangle of the running inverse W (W[i][c] at S[c][i]); diag(W) = 1 / diag(L) is
// kept aside.  With that packing the right-looking Cholesky step and the forward-substitution step of W = L^-1 are
// ONE masked rank-1 update per column j:
//     v = S[:, j]  (A[j:, j] below the diagonal, W'[j, :j] above it),  u = v with u_j = 1
//     S[r][q] -= u_r v_q / a_jj     for q > j and (r <= j  or  r >= q)
//     S[:, j] *= 1 / sqrt(a_jj)     (final: columns <= j are never touched again)
// Column j is broadcast through a double-buffered shared-memory vector, one __syncthreads per step; the owners of
// column j + 1 update and publish it (and the pivot's reciprocal square root) before the bulk of step j, so the
// sqrt / divide latency of the pivot sits off the critical path.
__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, double* __restrict__ M, double* __restrict__ U, long long ld,
                 long long batch_stride, int diag_block, const int* __restrict__ active, int* __restrict__ info) {
  const int bi = blockIdx.x;
  if (active != nullptr && active[bi] == 0) return;
  extern __shared__ double smem[];
  double* stage = smem;                      // [128][129] output transposition buffer
  double* vbuf = smem + TILE * POTF2_LD;     // [2][128]   broadcast column
  double* winv = vbuf + 2 * TILE;            // [128]      1 / L_jj
  double* ibuf = winv + TILE;                // [2][4]     {1/a_jj, 1/sqrt(a_jj), sqrt(a_jj)}
  int* badflag = reinterpret_cast<int*>(ibuf + 8);
  const long long base = (long long)bi * batch_stride + (long long)diag_block * TILE * ld + (long long)diag_block * TILE;
  double* Ab = A + base;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (tid == 0) *badflag = 0;

  double s[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = ty + 32 * a, q = tx + 16 * b;
      s[a][b] = (q <= r) ? Ab[(long long)r * ld + q] : 0.0;
    }

  // publishes column jn (held in s[.][B] by the threads with tx == jn % 16) and its pivot terms
#define POTF2_PUBLISH(B, jn)                                            \
  {                                                                     \
    double* vb_ = vbuf + ((jn)&1) * TILE;                               \
    _Pragma("unroll") for (int a = 0; a < 4; ++a) {                     \
      const int r = ty + 32 * a;                                        \
      vb_[r] = s[a][B];                                                 \
      if (r == (jn)) {                                                  \
        double d = s[a][B];                                             \
        if (!(d > 0.0) || !isfinite(d)) {                               \
          *badflag = 1;                                                 \
          d = 1.0;                                                      \
        }                                                               \
        double rinv = rsqrt(d); /* MUFU.RSQ64H + Newton: short chain */ \
        rinv = fma(fma(-d * rinv, rinv, 1.0), 0.5 * rinv, rinv);        \
        const double piv = d * rinv;                                    \
        double* ib_ = ibuf + ((jn)&1) * 4;                              \
        ib_[0] = rinv * rinv;                                           \
        ib_[1] = rinv;                                                  \
        ib_[2] = piv;                                                   \
        winv[(jn)] = rinv;                                              \
      }                                                                 \
    }                                                                   \
  }

  if (tx == 0) POTF2_PUBLISH(0, 0)
  __syncthreads();

#pragma unroll
  for (int bj = 0; bj < 8; ++bj) {
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int j = 16 * bj + jj;
      const double* v = vbuf + (j & 1) * TILE;
      const double* ib = ibuf + (j & 1) * 4;
      const double inv = ib[0];
      double u[4];
      bool low[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r = ty + 32 * a;
        const double vr = v[r];
        u[a] = (r == j) ? -1.0 : -vr;
        low[a] = r <= j;
      }
      // column j is final
      if (tx == jj) {
        const double rinv = ib[1], piv = ib[2];
#pragma unroll
        for (int a = 0; a < 4; ++a) s[a][bj] = (ty + 32 * a == j) ? piv : s[a][bj] * rinv;
      }
      // columns of this 16-group first, so that column j + 1 can be published early
      {
        const int q = tx + 16 * bj;
        const double t = v[q] * inv;
        if (tx > jj) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (low[a] || ty + 32 * a >= q) s[a][bj] = fma(u[a], t, s[a][bj]);
        }
      }
      if (jj < 15 && tx == jj + 1) POTF2_PUBLISH(bj, j + 1)
#pragma unroll
      for (int b = bj + 1; b < 8; ++b) {
        const int q = tx + 16 * b;
        const double t = v[q] * inv;
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (low[a] || ty + 32 * a >= q) s[a][b] = fma(u[a], t, s[a][b]);
      }
      if (jj == 15 && bj < 7 && tx == 0) POTF2_PUBLISH((bj < 7 ? bj + 1 : 7), j + 1)
      __syncthreads();
    }
  }
#undef POTF2_PUBLISH

  // transpose through shared memory so that L, W and W^T all leave with coalesced row writes
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) stage[(ty + 32 * a) * POTF2_LD + tx + 16 * b] = s[a][b];
  __syncthreads();
  double* Mb = M + base;
  double* Ub = U + base;
  for (int idx = tid; idx < TILE * TILE; idx += POTF2_THREADS) {
    const int i = idx / TILE, jx = idx % TILE;
    const double sij = stage[i * POTF2_LD + jx];
    if (jx <= i) Ab[(long long)i * ld + jx] = sij;
    const double wd = winv[i];
    Mb[(long long)i * ld + jx] = (jx < i) ? stage[jx * POTF2_LD + i] : (jx == i ? wd : 0.0);  // W[i][jx]
    Ub[(long long)i * ld + jx] = (jx > i) ? sij : (jx == i ? wd : 0.0);                       // W[jx][i]
  }
  if (tid == 0 && *badflag != 0 && info != nullptr) info[bi] = AGPC_ST_NOT_PD;
}

cudaError_t potf2_inv_launch(const Launch& L, double* A, double* M, double* U, long long ld, long long batch_stride,
                             int diag_block, int batch, const int* active, int* info) {
  ProfScope prof_scope(L, CLS_POTF2);
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
  const double wn = L.work_n > 0.0 ? L.work_n : (double)ms.np;
  L.add_work(CLS_GEMM, (double)L.work_items * wn * wn * wn / 3.0);  // LAPACK dpotrf count
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
  const double wn = L.work_n > 0.0 ? L.work_n : (double)ms.np;
  L.add_work(CLS_GEMM, (double)L.work_items * wn * wn * wn / 3.0);  // dtrtri
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
  const double wn = L.work_n > 0.0 ? L.work_n : (double)ms.np;
  L.add_work(CLS_GEMM, (double)L.work_items * wn * wn * wn / 3.0);  // dlauum
  GemmArgs a{};
  a.C = dst; a.Ct = nullptr; a.ldc = ms.np; a.c_batch = ms.stride;
  a.m_tiles = nblk; a.n_tiles = nblk; a.k_tiles = nblk * (TILE / GEMM_BK);
  a.k_rule = KRULE_BEGIN_AT_ROW; a.lower_only = 1; a.alpha = 1.0; a.beta = 0.0; a.active = active;
  return gemm_nt_launch(L, ms.mapU, ms.mapU, a, nblk * (nblk + 1) / 2, 1, ms.batch);
}

}  // namespace agpc
