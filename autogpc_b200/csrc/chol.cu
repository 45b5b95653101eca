// Blocked FP64 Cholesky / triangular inverse / B^-1 for B = I + S^1/2 K S^1/2 (batched, lock-step).
//
// Stands in for GPy util/linalg.py jitchol (dpotrf), dtrtrs / dtrtri and pdinv's dpotri + symmetrify, as reached
// from EP.expectation_propagation and EP.inference for gpckernel.py:245-246.
//
//   potrf  : 128-wide panels in outer blocks of 4 (two-level blocking): inside a block left-looking (in-block update
//            -> potf2_inv, one CTA per matrix, factors the diagonal block AND inverts it -> panel solve through the
//            inverse, P <- P W_k^T), then one K = 512 trailing update per block; all products on the TMA + DMMA tile
//            kernel (gemm_nt.cu).  Look-ahead: panels + the next block's columns on a high-priority stream, the rest
//            of the trailing update on the caller's stream.  W_k = L_kk^-1 doubles as level 0 of trtri.
//   trtri  : level-synchronous recursive doubling, M21 = -M22 L21 M11, two batched triangular-aware GEMMs per level;
//            both M = L^-1 and U = M^T are kept so every product is in K-contiguous ("NT") form.
//   lauum  : B^-1 = U U^T, lower tiles only, K range clipped at the tile row.
#include "common.cuh"

namespace agpc {

constexpr int POTF2_THREADS = 256;
constexpr int POTF2_LD = TILE + 1;
constexpr int POTF2_SMEM = (TILE * POTF2_LD + 2 * TILE + TILE + 16) * 8;

// One CTA per batch item factors the 128 x 128 diagonal block AND inverts its factor, entirely in registers.
//
// The square S (thread (ty, tx) of a 16 x 16 grid owns rows ty + 16a, columns tx + 16b, 8 x 8 each: cyclic, so every thread
// stays busy as the active region shrinks) holds the lower triangle of A (-> L) and, transposed into its strict
// upper triangle, the strict lower triangle of the running inverse W (W[i][c] at S[c][i]); diag(W) = 1 / diag(L) is
// kept aside.  With that packing the right-looking Cholesky step and the forward-substitution step of W = L^-1 are
// ONE masked rank-1 update per column j:
//     v = S[:, j]  (A[j:, j] below the diagonal, W'[j, :j] above it),  u = v with u_j = 1
//     S[r][q] -= u_r v_q / a_jj     for q > j and (r <= j  or  r >= q)
//     S[:, j] *= 1 / sqrt(a_jj)     (final: columns <= j are never touched again)
// Column j is broadcast through a double-buffered shared-memory vector, one __syncthreads per step; the owners of
// column j + 1 update and publish it (and the pivot's reciprocal square root) before the bulk of step j, so the
// rsqrt latency of the pivot sits off the critical path.
//
// Masking is branch-free.  Element (r, q) is "lower" (part of A, live from the start) when r >= q and "upper"
// (W[q][r], created at step r) otherwise; an upper element must not move before step r, which is arranged by
// multiplying it with u_low (u with the rows r > j zeroed) instead of u.  Whether an element is lower is a
// compile-time fact except on the 16 x 16 patches on the diagonal (a == b), where it is ty >= tx.
struct Potf2Smem {
  double* vbuf;   // [2][128] broadcast column
  double* winv;   // [128]    1 / L_jj
  double* ibuf;   // [2][4]   1/a_jj (slot 0)
  int* badflag;
};

// publishes column jn (of 16-group BP), held in col[0..8) by the threads that own it, and its pivot terms; the pivot
// a_jn,jn can only sit in row group BP, on the thread with ty == jn % 16
template <int BP>
__device__ __forceinline__ void potf2_publish(const Potf2Smem& sm, const double (&col)[8], int jn, int ty) {
  double* vb = sm.vbuf + (jn & 1) * TILE;
#pragma unroll
  for (int a = 0; a < 8; ++a) vb[ty + 16 * a] = col[a];
  if (ty == (jn & 15)) {
    double d = col[BP];
    if (!(d > 0.0) || !isfinite(d)) {
      *sm.badflag = 1;
      d = 1.0;
    }
    const double rinv = rsqrt(d);  // 1 ulp (MUFU.RSQ64H + Newton)
    sm.ibuf[(jn & 1) * 4] = rinv * rinv;
    sm.winv[jn] = rinv;
  }
}

// um = multiplier of row group A for the 16 x 16 patch (A, B): below the diagonal (A > B) the element is part of A
// and always live; above it (A < B) it is W[q][r], live once step r has passed; on the diagonal patch it depends on
// the thread (ty >= tx).
template <int A, int B>
__device__ __forceinline__ double potf2_um(const double (&uf)[8], const double (&ul)[8], bool diag_lower) {
  return (A > B) ? uf[A] : (A < B) ? ul[A] : (diag_lower ? uf[A] : ul[A]);
}

template <int BJ, int B>
__device__ __forceinline__ void potf2_load_t(double (&t)[8], const double* v, double inv, int tx, int jj) {
  if constexpr (B < 8) {
    const double x = v[tx + 16 * B] * inv;
    t[B] = (B == BJ) ? ((tx > jj) ? x : 0.0) : x;  // finished columns of the current 16-group get a zero update
    potf2_load_t<BJ, B + 1>(t, v, inv, tx, jj);
  }
}

template <int B>
__device__ __forceinline__ void potf2_col_update(double (&s)[8][8], const double (&uf)[8], const double (&ul)[8],
                                                 bool diag_lower, double t) {
  s[0][B] = fma(-potf2_um<0, B>(uf, ul, diag_lower), t, s[0][B]);
  s[1][B] = fma(-potf2_um<1, B>(uf, ul, diag_lower), t, s[1][B]);
  s[2][B] = fma(-potf2_um<2, B>(uf, ul, diag_lower), t, s[2][B]);
  s[3][B] = fma(-potf2_um<3, B>(uf, ul, diag_lower), t, s[3][B]);
  s[4][B] = fma(-potf2_um<4, B>(uf, ul, diag_lower), t, s[4][B]);
  s[5][B] = fma(-potf2_um<5, B>(uf, ul, diag_lower), t, s[5][B]);
  s[6][B] = fma(-potf2_um<6, B>(uf, ul, diag_lower), t, s[6][B]);
  s[7][B] = fma(-potf2_um<7, B>(uf, ul, diag_lower), t, s[7][B]);
}

template <int BJ, int B>
__device__ __forceinline__ void potf2_rest(double (&s)[8][8], const double (&uf)[8], const double (&ul)[8],
                                           bool diag_lower, const double (&t)[8]) {
  if constexpr (B < 8) {
    potf2_col_update<B>(s, uf, ul, diag_lower, t[B]);
    potf2_rest<BJ, B + 1>(s, uf, ul, diag_lower, t);
  }
}

// the 16 columns j = 16 BJ .. 16 BJ + 15.  Column j is NOT scaled here: columns <= j simply stop receiving updates
// (their t is zero / they are skipped), and the 1 / sqrt(a_jj) scaling of every finished column -- diagonal included,
// a_jj / sqrt(a_jj) = L_jj -- is applied once at write-out.
template <int BJ>
__device__ __forceinline__ void potf2_group(double (&s)[8][8], const Potf2Smem& sm, bool diag_lower, int tx, int ty) {
#pragma unroll 1
  for (int jj = 0; jj < 16; ++jj) {
    const int j = 16 * BJ + jj;
    const double* v = sm.vbuf + (j & 1) * TILE;
    const double inv = sm.ibuf[(j & 1) * 4];
    double t[8];
    potf2_load_t<BJ, BJ>(t, v, inv, tx, jj);
    double uf[8], ul[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double vr = v[ty + 16 * a];
      // rows of 16-row groups a < BJ are all <= j, rows of groups a > BJ are all > j
      if (a < BJ) {
        uf[a] = vr;
        ul[a] = vr;
      } else if (a > BJ) {
        uf[a] = vr;
        ul[a] = 0.0;
      } else {
        uf[a] = (ty == jj) ? 1.0 : vr;
        ul[a] = (ty <= jj) ? uf[a] : 0.0;
      }
    }
    // columns of this 16-group first (lanes tx > jj), so that column j + 1 can be published early
    potf2_col_update<BJ>(s, uf, ul, diag_lower, t[BJ]);
    if (jj < 15 && tx == jj + 1) {
      const double col[8] = {s[0][BJ], s[1][BJ], s[2][BJ], s[3][BJ], s[4][BJ], s[5][BJ], s[6][BJ], s[7][BJ]};
      potf2_publish<BJ>(sm, col, j + 1, ty);
    }
    potf2_rest<BJ, BJ + 1>(s, uf, ul, diag_lower, t);
    if constexpr (BJ < 7) {
      if (jj == 15 && tx == 0) {
        const double col[8] = {s[0][BJ + 1], s[1][BJ + 1], s[2][BJ + 1], s[3][BJ + 1],
                               s[4][BJ + 1], s[5][BJ + 1], s[6][BJ + 1], s[7][BJ + 1]};
        potf2_publish<BJ + 1>(sm, col, j + 1, ty);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, double* __restrict__ M, double* __restrict__ U, long long ld,
                 long long batch_stride, int diag_block, const int* __restrict__ active, int* __restrict__ info) {
  const int bi = blockIdx.x;
  if (active != nullptr && active[bi] == 0) return;
  extern __shared__ double smem[];
  double* stage = smem;                      // [128][129] output transposition buffer
  Potf2Smem sm;
  sm.vbuf = smem + TILE * POTF2_LD;
  sm.winv = sm.vbuf + 2 * TILE;
  sm.ibuf = sm.winv + TILE;
  sm.badflag = reinterpret_cast<int*>(sm.ibuf + 8);
  const long long base = (long long)bi * batch_stride + (long long)diag_block * TILE * ld + (long long)diag_block * TILE;
  double* Ab = A + base;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (tid == 0) *sm.badflag = 0;

  double s[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = ty + 16 * a, q = tx + 16 * b;
      s[a][b] = (q <= r) ? Ab[(long long)r * ld + q] : 0.0;
    }
  const bool diag_lower = ty >= tx;
  __syncthreads();
  if (tx == 0) {
    const double col[8] = {s[0][0], s[1][0], s[2][0], s[3][0], s[4][0], s[5][0], s[6][0], s[7][0]};
    potf2_publish<0>(sm, col, 0, ty);
  }
  __syncthreads();

  potf2_group<0>(s, sm, diag_lower, tx, ty);
  potf2_group<1>(s, sm, diag_lower, tx, ty);
  potf2_group<2>(s, sm, diag_lower, tx, ty);
  potf2_group<3>(s, sm, diag_lower, tx, ty);
  potf2_group<4>(s, sm, diag_lower, tx, ty);
  potf2_group<5>(s, sm, diag_lower, tx, ty);
  potf2_group<6>(s, sm, diag_lower, tx, ty);
  potf2_group<7>(s, sm, diag_lower, tx, ty);

  // transpose through shared memory so that L, W and W^T all leave with coalesced row writes
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const double wq = sm.winv[tx + 16 * b];  // deferred column scaling
#pragma unroll
    for (int a = 0; a < 8; ++a) stage[(ty + 16 * a) * POTF2_LD + tx + 16 * b] = s[a][b] * wq;
  }
  __syncthreads();
  double* Mb = M + base;
  double* Ub = U + base;
  for (int idx = tid; idx < TILE * TILE; idx += POTF2_THREADS) {
    const int i = idx / TILE, jx = idx % TILE;
    const double sij = stage[i * POTF2_LD + jx];
    if (jx <= i) Ab[(long long)i * ld + jx] = sij;
    const double wd = sm.winv[i];
    Mb[(long long)i * ld + jx] = (jx < i) ? stage[jx * POTF2_LD + i] : (jx == i ? wd : 0.0);  // W[i][jx]
    Ub[(long long)i * ld + jx] = (jx > i) ? sij : (jx == i ? wd : 0.0);                       // W[jx][i]
  }
  if (tid == 0 && *sm.badflag != 0 && info != nullptr) info[bi] = AGPC_ST_NOT_PD;
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
// dst[r][c0 .. c0 + 128) = src[r][c0 .. c0 + 128) for rows r0 .. r0 + rows (one 128-wide panel, per batch item)
__global__ void copy_panel_kernel(const double* __restrict__ src, double* __restrict__ dst, long long ld,
                                  long long batch_stride, int r0, int c0, int rows, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one double2 each
  const int r = (int)(idx / (TILE / 2)), c = (int)(idx % (TILE / 2)) * 2;
  if (r >= rows) return;
  const long long o = (long long)b * batch_stride + (long long)(r0 + r) * ld + c0 + c;
  *reinterpret_cast<double2*>(dst + o) = *reinterpret_cast<const double2*>(src + o);
}
static cudaError_t copy_panel_launch(const Launch& L, const double* src, double* dst, long long ld, long long batch_stride,
                                     int r0, int c0, int rows, int batch, const int* active) {
  ProfScope prof_scope(L, CLS_MISC);
  const long long n2 = (long long)rows * (TILE / 2);
  dim3 grid((unsigned)((n2 + 255) / 256), batch);
  copy_panel_kernel<<<grid, 256, 0, L.stream>>>(src, dst, ld, batch_stride, r0, c0, rows, active);
  L.bump();
  return cudaGetLastError();
}

// Two-level blocking: 128-wide panels (the potf2_inv / trsm grain) grouped into outer blocks of POTRF_OUTER panels.
// Inside an outer block the algorithm is left-looking (panel k is updated by the block's earlier panels, K <= 384);
// the trailing matrix is updated once per outer block with K = 512, where the tile GEMM runs close to its DGEMM rate
// (a K = 128 rank update pays the TMA prologue and the C round trip once per 8 k-iterations: 24 vs 33+ TFLOP/s).
constexpr int POTRF_OUTER = 4;
// one-tile-wide launches with fewer CTAs than this run as 128 x 32 tiles (4x the CTAs, a quarter of the latency)
constexpr int THIN_CTA_LIMIT = 160;

// panels k0 .. k0 + kb of one outer block: in-block left-looking update, diagonal block, panel solve
// first_staged: the head update of the previous outer block already duplicated panel k0 (rows > k0) into M
static cudaError_t potrf_panels(const Launch& L, const MatSet& ms, const int* active, int* info, int k0, int kb,
                                bool first_staged) {
  const int nblk = ms.np / TILE;
  for (int t = 0; t < kb; ++t) {
    const int k = k0 + t;
    bool staged = (t == 0) && first_staged;  // is the panel below the diagonal block already duplicated in M?
    if (t > 0) {
      // A[k:, k] -= A[k:, k0:k] A[k, k0:k]^T
      GemmArgs u{};
      u.C = ms.A; u.Ct = nullptr; u.ldc = ms.np; u.c_batch = ms.stride;
      u.a_row0 = k * TILE; u.a_col0 = k0 * TILE;
      u.b_row0 = k * TILE; u.b_col0 = k0 * TILE;
      u.c_row0 = k * TILE; u.c_col0 = k * TILE;
      u.m_tiles = nblk - k; u.n_tiles = 1; u.k_tiles = t * (TILE / GEMM_BK);
      u.k_rule = KRULE_FULL; u.lower_only = 0; u.alpha = -1.0; u.beta = 1.0; u.active = active;
      if ((nblk - k) * ms.batch <= THIN_CTA_LIMIT) {
        u.n_tiles = TILE / 32;
        // the thin solve below cannot run in place: this update stages its result (rows > k) in M as it goes
        u.C2 = ms.M; u.c2_cols = u.n_tiles; u.c2_row_min = 1;
        staged = true;
        AGPC_CUDA_TRY(gemm_nt_thin_launch(L, ms.mapA, ms.mapA32, u, (nblk - k) * u.n_tiles, 1, ms.batch));
      } else {
        AGPC_CUDA_TRY(gemm_nt_launch(L, ms.mapA, ms.mapA, u, nblk - k, 1, ms.batch));
      }
    }
    AGPC_CUDA_TRY(potf2_inv_launch(L, ms.A, ms.M, ms.U, ms.np, ms.stride, k, ms.batch, active, info));
    const int below = nblk - k - 1;
    if (below == 0) break;
    // panel solve through the inverse: A[k+1:, k] <- A[k+1:, k] W_k^T  (W_k = L_kk^-1 is lower: K clipped)
    GemmArgs p{};
    p.C = ms.A; p.Ct = nullptr; p.ldc = ms.np; p.c_batch = ms.stride;
    p.a_row0 = (k + 1) * TILE; p.a_col0 = k * TILE;
    p.b_row0 = k * TILE; p.b_col0 = k * TILE;  // W_k from M's diagonal block
    p.c_row0 = (k + 1) * TILE; p.c_col0 = k * TILE;
    p.m_tiles = below; p.n_tiles = 1; p.k_tiles = TILE / GEMM_BK;
    p.k_rule = KRULE_END_AT_COL; p.lower_only = 0; p.alpha = 1.0; p.beta = 0.0; p.active = active;
    if (below * ms.batch <= THIN_CTA_LIMIT) {
      // four CTAs share one 128-row operand tile here, so the solve cannot run in place: stage the panel in the
      // (still unused) strictly-lower part of M and read it from there
      if (!staged)
        AGPC_CUDA_TRY(copy_panel_launch(L, ms.A, ms.M, ms.np, ms.stride, (k + 1) * TILE, k * TILE, below * TILE,
                                        ms.batch, active));
      p.n_tiles = TILE / 32;
      AGPC_CUDA_TRY(gemm_nt_thin_launch(L, ms.mapM, ms.mapM32, p, below * p.n_tiles, 1, ms.batch));
    } else {
      AGPC_CUDA_TRY(gemm_nt_launch(L, ms.mapA, ms.mapM, p, below, 1, ms.batch));
    }
  }
  return cudaSuccess;
}

// trailing update with outer block [k0, k0 + kb): C[r0:, c0 : c0 + ncols] -= A[., k0:k0+kb] A[., k0:k0+kb]^T for the
// tiles on or below the diagonal.  (r0 == c0; ncols < 0: everything to the right, as a lower-triangular tile list)
static cudaError_t potrf_trailing(const Launch& L, const MatSet& ms, const int* active, int k0, int kb, int r0, int ncols,
                                  bool stage_first_col = false) {
  const int nblk = ms.np / TILE;
  const int rows = nblk - r0;
  if (rows <= 0 || ncols == 0) return cudaSuccess;
  GemmArgs s{};
  s.C = ms.A; s.Ct = nullptr; s.ldc = ms.np; s.c_batch = ms.stride;
  s.a_row0 = r0 * TILE; s.a_col0 = k0 * TILE;
  s.b_row0 = r0 * TILE; s.b_col0 = k0 * TILE;
  s.c_row0 = r0 * TILE; s.c_col0 = r0 * TILE;
  s.k_tiles = kb * (TILE / GEMM_BK);
  s.k_rule = KRULE_FULL; s.alpha = -1.0; s.beta = 1.0; s.active = active;
  if (ncols < 0) {
    s.m_tiles = rows; s.n_tiles = rows; s.lower_only = 1;
    return gemm_nt_launch(L, ms.mapA, ms.mapA, s, rows * (rows + 1) / 2, 1, ms.batch);
  }
  const int nc = ncols < rows ? ncols : rows;
  s.m_tiles = rows; s.n_tiles = nc; s.lower_only = 0; s.skip_upper = 1;
  if (stage_first_col) {  // next panel (first block column, rows below its diagonal block) also goes to the M scratch
    s.C2 = ms.M; s.c2_cols = 1; s.c2_row_min = 1;
  }
  return gemm_nt_launch(L, ms.mapA, ms.mapA, s, rows * nc, 1, ms.batch);
}

// Small batches run with look-ahead: the trailing update of outer block i is split into its first POTRF_OUTER block
// columns (the "head": exactly the panels of outer block i + 1) and the rest.  The latency-bound chain panels -> head
// -> panels ... runs on the context's HIGH-PRIORITY stream, so its few CTAs get the next SM that frees up, while the
// bulk rest(i) keeps every other SM busy on the caller's stream:
//   hi:  panels(i) | head(i) [after rest(i-1)] | panels(i+1) | head(i+1) [after rest(i)] | ...
//   lo:            | rest(i) [after panels(i), rest(i-1)]    | rest(i+1) ...
cudaError_t potrf_blocked(const Launch& L, const MatSet& ms, const int* active, int* info) {
  const int nblk = ms.np / TILE;
  const double wn = L.work_n > 0.0 ? L.work_n : (double)ms.np;
  L.add_work(CLS_GEMM, (double)L.work_items * wn * wn * wn / 3.0);  // LAPACK dpotrf count
  const bool lookahead = L.side != nullptr && nblk > 2 * POTRF_OUTER;
  if (!lookahead) {
    for (int k0 = 0; k0 < nblk; k0 += POTRF_OUTER) {
      const int kb = nblk - k0 < POTRF_OUTER ? nblk - k0 : POTRF_OUTER;
      AGPC_CUDA_TRY(potrf_panels(L, ms, active, info, k0, kb, false));
      if (k0 + kb < nblk) AGPC_CUDA_TRY(potrf_trailing(L, ms, active, k0, kb, k0 + kb, -1));
    }
    return cudaSuccess;
  }
  Launch H = L;  // high-priority chain
  H.stream = L.side;
  AGPC_CUDA_TRY(cudaEventRecord(L.ev_rest, L.stream));  // everything queued so far (form_B ...) precedes the chain
  AGPC_CUDA_TRY(cudaStreamWaitEvent(H.stream, L.ev_rest, 0));
  bool rest_pending = false;
  for (int k0 = 0; k0 < nblk; k0 += POTRF_OUTER) {
    const int kb = nblk - k0 < POTRF_OUTER ? nblk - k0 : POTRF_OUTER;
    // the head update of the previous block staged panel k0 when that panel will be solved with thin tiles
    const bool first_staged = k0 > 0 && (nblk - k0 - 1) * ms.batch <= THIN_CTA_LIMIT;
    AGPC_CUDA_TRY(potrf_panels(H, ms, active, info, k0, kb, first_staged));
    const int r0 = k0 + kb;
    if (r0 >= nblk) break;
    AGPC_CUDA_TRY(cudaEventRecord(L.ev_panel, H.stream));
    if (rest_pending) AGPC_CUDA_TRY(cudaStreamWaitEvent(H.stream, L.ev_rest, 0));  // rest(i-1) wrote the head's columns
    AGPC_CUDA_TRY(potrf_trailing(H, ms, active, k0, kb, r0, POTRF_OUTER, (nblk - r0 - 1) * ms.batch <= THIN_CTA_LIMIT));
    rest_pending = false;
    if (r0 + POTRF_OUTER < nblk) {
      AGPC_CUDA_TRY(cudaStreamWaitEvent(L.stream, L.ev_panel, 0));
      AGPC_CUDA_TRY(potrf_trailing(L, ms, active, k0, kb, r0 + POTRF_OUTER, -1));
      AGPC_CUDA_TRY(cudaEventRecord(L.ev_rest, L.stream));
      rest_pending = true;
    }
  }
  AGPC_CUDA_TRY(cudaEventRecord(L.ev_panel, H.stream));  // join: the caller's stream continues after the chain
  AGPC_CUDA_TRY(cudaStreamWaitEvent(L.stream, L.ev_panel, 0));
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
