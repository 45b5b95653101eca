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
#include <cstdlib>

#include "common.cuh"

namespace agpc {

constexpr int POTF2_THREADS = 256;
constexpr int POTF2_LD = TILE + 1;
constexpr int POTF2_SMEM = (TILE * POTF2_LD + 4 * TILE + TILE + 16) * 8;

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
// Columns are taken two at a time (see potf2_publish_pair): the pair is broadcast through a double-buffered shared-
// memory buffer, ONE __syncthreads per pair; the owners of the next pair update and publish it (and the pivot thread
// the 2 x 2 pivot block's reciprocals) before the bulk of the current pair's update, so the pivot latency sits off
// the critical path.  53.8 us per block (58 with one column per barrier, 405 for the first shared-memory version);
// what remains is FP64 issue / dependency stalls of the update itself, not the barrier chain.
//
// Masking is branch-free.  Element (r, q) is "lower" (part of A, live from the start) when r >= q and "upper"
// (W[q][r], created at step r) otherwise; an upper element must not move before step r, which is arranged by
// multiplying it with u_low (u with the rows r > j zeroed) instead of u.  Whether an element is lower is a
// compile-time fact except on the 16 x 16 patches on the diagonal (a == b), where it is ty >= tx.
struct Potf2Smem {
  double* vbuf;   // [2 buffers][2 columns][128] broadcast column pair
  double* winv;   // [128]    1 / L_jj
  double* ibuf;   // [2][4]   {1/a1, c = S[j+1][j]/a1, 1/a2'}
  int* badflag;
};

// Columns are eliminated in PAIRS (j even, j + 1): one barrier, one broadcast and one pivot chain per two columns.
// With v1 = S[:, j], v2 = S[:, j+1] (both as they stand before the pair), a1 = v1[j], c = v1[j+1] / a1 and u1 = v1 with
// u1[j] = 1, column j + 1 after step j is v2' = v2 - u1 c on EVERY row, its pivot a2' = v2'[j+1], and
//     S[r][q] -= um1(r) v1[q] / a1 + um2(r) v2'[q] / a2'      (q > j + 1)
// where um1 / um2 are u1 / u2' = v2' with u2'[j+1] = 1 under the liveness rule of their own column.
//
// Publishes the raw pair (jn, jn + 1) of 16-group BP.  The three pivot-block entries S[jn][jn], S[jn+1][jn],
// S[jn+1][jn+1] all sit in s[BP][BP] of three lanes of ONE warp (ty = jn % 16 and ty + 1 share a warp), so the pivot
// thread collects them with shuffles -- which every thread must therefore execute.
template <int BP>
__device__ __forceinline__ void potf2_publish_pair(const Potf2Smem& sm, const double (&s)[8][8], int jn, int tx, int ty) {
  const int jj = jn & 15;
  const double d = s[BP][BP];
  const double a11 = __shfl_sync(0xffffffffu, d, jj);
  const double a21 = __shfl_sync(0xffffffffu, d, 16 + jj);
  const double a22 = __shfl_sync(0xffffffffu, d, 16 + jj + 1);
  if (tx == jj || tx == jj + 1) {
    double* vb = sm.vbuf + ((jn >> 1) & 1) * (2 * TILE) + (tx - jj) * TILE;
#pragma unroll
    for (int a = 0; a < 8; ++a) vb[ty + 16 * a] = s[a][BP];
    if (tx == jj && ty == jj) {  // pivot thread of the pair
      double a1 = a11;
      if (!(a1 > 0.0) || !isfinite(a1)) {
        *sm.badflag = 1;
        a1 = 1.0;
      }
      const double r1 = rsqrt(a1);
      const double inv1 = r1 * r1;
      const double c = a21 * inv1;
      double a2 = fma(-a21, c, a22);
      if (!(a2 > 0.0) || !isfinite(a2)) {
        *sm.badflag = 1;
        a2 = 1.0;
      }
      const double r2 = rsqrt(a2);
      double* ib = sm.ibuf + ((jn >> 1) & 1) * 4;
      ib[0] = inv1;
      ib[1] = c;
      ib[2] = r2 * r2;
      sm.winv[jn] = r1;
      sm.winv[jn + 1] = r2;
    }
  }
}

// um = multiplier of row group A for the 16 x 16 patch (A, B): below the diagonal (A > B) the element is part of A
// and always live; above it (A < B) it is W[q][r], live once step r has passed; on the diagonal patch it depends on
// the thread (ty >= tx).
template <int A, int B>
__device__ __forceinline__ double potf2_um(const double (&uf)[8], const double (&ul)[8], bool diag_lower) {
  return (A > B) ? uf[A] : (A < B) ? ul[A] : (diag_lower ? uf[A] : ul[A]);
}

template <int A, int B>
__device__ __forceinline__ void potf2_el_update(double (&s)[8][8], const double (&uf1)[8], const double (&ul1)[8],
                                                const double (&uf2)[8], const double (&ul2)[8], bool diag_lower,
                                                double t1, double t2) {
  double x = fma(-potf2_um<A, B>(uf1, ul1, diag_lower), t1, s[A][B]);
  s[A][B] = fma(-potf2_um<A, B>(uf2, ul2, diag_lower), t2, x);
}

template <int B>
__device__ __forceinline__ void potf2_col_update(double (&s)[8][8], const double (&uf1)[8], const double (&ul1)[8],
                                                 const double (&uf2)[8], const double (&ul2)[8], bool diag_lower,
                                                 double t1, double t2) {
  potf2_el_update<0, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<1, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<2, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<3, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<4, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<5, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<6, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  potf2_el_update<7, B>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
}

template <int BJ, int B>
__device__ __forceinline__ void potf2_load_t(double (&t1)[8], double (&t2)[8], const double* v1, const double* v2,
                                             double inv1, double c, double inv2, int tx, int jj) {
  if constexpr (B < 8) {
    const double x1 = v1[tx + 16 * B];
    const double x2 = fma(-x1, c, v2[tx + 16 * B]);  // column j + 1 after step j, at row q (q > j + 1: u1[q] = v1[q])
    const bool on = (B != BJ) || (tx > jj + 1);      // finished / current columns of this 16-group get a zero update
    t1[B] = on ? x1 * inv1 : 0.0;
    t2[B] = on ? x2 * inv2 : 0.0;
    potf2_load_t<BJ, B + 1>(t1, t2, v1, v2, inv1, c, inv2, tx, jj);
  }
}

template <int BJ, int B>
__device__ __forceinline__ void potf2_rest(double (&s)[8][8], const double (&uf1)[8], const double (&ul1)[8],
                                           const double (&uf2)[8], const double (&ul2)[8], bool diag_lower,
                                           const double (&t1)[8], const double (&t2)[8]) {
  if constexpr (B < 8) {
    potf2_col_update<B>(s, uf1, ul1, uf2, ul2, diag_lower, t1[B], t2[B]);
    potf2_rest<BJ, B + 1>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
  }
}

// the 16 columns j = 16 BJ .. 16 BJ + 15, two at a time.  Finished columns are NOT scaled here: they simply stop
// receiving updates, and the 1 / sqrt(pivot) scaling of every column -- diagonal included, a / sqrt(a) = L_jj -- is
// applied once at write-out.
template <int BJ>
__device__ __forceinline__ void potf2_group(double (&s)[8][8], const Potf2Smem& sm, bool diag_lower, int tx, int ty) {
#pragma unroll 1
  for (int jj = 0; jj < 16; jj += 2) {
    const int j = 16 * BJ + jj;
    const double* v1 = sm.vbuf + ((j >> 1) & 1) * (2 * TILE);
    const double* v2 = v1 + TILE;
    const double* ib = sm.ibuf + ((j >> 1) & 1) * 4;
    const double inv1 = ib[0], c = ib[1], inv2 = ib[2];
    double t1[8], t2[8];
    potf2_load_t<BJ, BJ>(t1, t2, v1, v2, inv1, c, inv2, tx, jj);
    double uf1[8], ul1[8], uf2[8], ul2[8], v2p[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double x1 = v1[ty + 16 * a];
      // rows of 16-row groups a < BJ are all <= j, rows of groups a > BJ are all > j + 1
      const double u1 = (a == BJ && ty == jj) ? 1.0 : x1;
      const double x2 = fma(-u1, c, v2[ty + 16 * a]);  // column j + 1 after step j (every row is live for it)
      const double u2 = (a == BJ && ty == jj + 1) ? 1.0 : x2;
      v2p[a] = x2;
      uf1[a] = u1;
      uf2[a] = u2;
      if (a < BJ) {
        ul1[a] = u1;
        ul2[a] = u2;
      } else if (a > BJ) {
        ul1[a] = 0.0;
        ul2[a] = 0.0;
      } else {
        ul1[a] = (ty <= jj) ? u1 : 0.0;
        ul2[a] = (ty <= jj + 1) ? u2 : 0.0;
      }
    }
    // column j + 1 takes its post-step-j values (its owners published the raw ones)
    if (tx == jj + 1) {
#pragma unroll
      for (int a = 0; a < 8; ++a) s[a][BJ] = v2p[a];
    }
    // columns of this 16-group first (lanes tx > jj + 1), so that the next pair can be published early
    potf2_col_update<BJ>(s, uf1, ul1, uf2, ul2, diag_lower, t1[BJ], t2[BJ]);
    if (jj < 14) potf2_publish_pair<BJ>(sm, s, j + 2, tx, ty);
    potf2_rest<BJ, BJ + 1>(s, uf1, ul1, uf2, ul2, diag_lower, t1, t2);
    if constexpr (BJ < 7) {
      if (jj == 14) potf2_publish_pair<BJ + 1>(sm, s, j + 2, tx, ty);
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
  sm.winv = sm.vbuf + 4 * TILE;
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
  potf2_publish_pair<0>(sm, s, 0, tx, ty);
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
  bool& configured = func_attr_done(__LINE__);  // per device: the attribute belongs to the device's context
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

// ---- int8 tensor-core path (ozaki.cu) -----------------------------------------------------------------------------
// Products with a long K extent run as sliced-integer GEMMs when the MatSet carries slice buffers (ms.oz_S = 6 / 7):
// the operands are re-sliced from their FP64 source right before each product (row scale = exact row maximum over the
// K range that product uses), SA serves the A operand and SB the B operand.
static inline bool oz_on(const MatSet& ms) { return ms.oz_S > 0 && ms.SA != nullptr; }
// which: 0 = SA, 1 = SB, 2 = SC
static inline OzOperand oz_operand(const MatSet& ms, int which) {
  return OzOperand{which == 2 ? ms.SC : which ? ms.SB : ms.SA, which == 2 ? ms.scC : which ? ms.scB : ms.scA,
                   (long long)ms.oz_S * ms.np * ms.np, (long long)ms.np, ms.np / OZ_KB};
}
static inline OzSliceArgs oz_slice_args(const MatSet& ms, const double* X, int which, const int* active) {
  OzSliceArgs a{};
  a.X = X; a.ld = ms.np; a.x_batch = ms.stride;
  a.S = which == 2 ? ms.SC : which ? ms.SB : ms.SA; a.scale = which == 2 ? ms.scC : which ? ms.scB : ms.scA;
  a.s_batch = (long long)ms.oz_S * ms.np * ms.np; a.sc_batch = ms.np; a.nkc = ms.np / OZ_KB;
  a.active = active;
  return a;
}
constexpr int OZ_CPB = TILE / OZ_KB;   // K chunks per 128-block
constexpr int OZ_TRTRI_MIN_H = 2;      // trtri levels below this half-size (K = 128) stay on the DMMA kernel

// slices the factored panel columns [k0, k0 + kb) for the rows from block r0 down: both operands of the trailing update
static cudaError_t potrf_slice_panel(const Launch& L, const MatSet& ms, const int* active, int k0, int kb, int r0) {
  const int nblk = ms.np / TILE;
  if (nblk - r0 <= 0) return cudaSuccess;
  OzSliceArgs a = oz_slice_args(ms, ms.A, 0, active);
  a.row0 = r0 * TILE; a.col0 = k0 * TILE; a.row_blocks = nblk - r0; a.col_chunks = kb * OZ_CPB;
  a.fixed_scale = ms.scA;  // row scales from |L_ik| <= sqrt(B_ii) (oz_row_bound_launch): every panel of L shares them
  L.add_work(CLS_SLICE, 8.0 * (double)(nblk - r0) * TILE * kb * TILE * L.work_items);
  return oz_slice_launch(L, a, nblk - r0, 1, ms.batch, ms.oz_S);
}

static cudaError_t potrf_panels(const Launch& L, const MatSet& ms, const int* active, int* info, int k0, int kb,
                                bool first_staged);

// Left-looking update of one outer block column from the persistent digit slices of the panels factored so far:
//   C[k0:, k0 : k0 + kb] -= L[k0:, kc0 : kc0 + kcb] L[k0 : k0 + kb, kc0 : kc0 + kcb]^T     (tiles on or below the diagonal)
// One call with a long K range replaces kc / 4 right-looking K = 512 passes over the same tiles: the per-tile prologue and
// epilogue of the int8 kernel (a third of a K = 512 launch) are paid once per tile instead of once per outer block.
static cudaError_t potrf_update_oz(const Launch& L, const MatSet& ms, const int* active, int k0, int kb, int kc0, int kcb,
                                   bool stage_first_col, bool concurrent = false) {
  const int nblk = ms.np / TILE;
  const int rows = nblk - k0;
  if (rows <= 0 || kcb <= 0) return cudaSuccess;
  OzGemmArgs z{};
  z.A = oz_operand(ms, 0); z.B = z.A;
  z.C = ms.A; z.Ct = nullptr; z.ldc = ms.np; z.c_batch = ms.stride;
  z.a_row0 = k0 * TILE; z.a_col0 = kc0 * TILE; z.b_row0 = k0 * TILE; z.b_col0 = kc0 * TILE;
  z.c_row0 = k0 * TILE; z.c_col0 = k0 * TILE;
  z.k_chunks = kcb * OZ_CPB; z.k_rule = KRULE_FULL; z.alpha = -1.0; z.beta = 1.0; z.active = active;
  z.concurrent = concurrent ? 1 : 0;
  const int nc = kb < rows ? kb : rows;
  z.m_tiles = rows; z.n_tiles = nc; z.lower_only = 0; z.skip_upper = 1;
  AGPC_CUDA_TRY(oz_gemm_launch(L, z, rows * nc, 1, ms.batch, ms.oz_S));
  if (stage_first_col && rows > 1)  // the thin panel solve reads its panel from the M scratch
    AGPC_CUDA_TRY(copy_panel_launch(L, ms.A, ms.M, ms.np, ms.stride, (k0 + 1) * TILE, k0 * TILE, (rows - 1) * TILE, ms.batch, active));
  return cudaSuccess;
}

// potrf on the int8 tensor-core path: LEFT-looking over outer blocks.  Block column j is brought up to date from all
// earlier panels right before it is factored -- part 1 (every panel but the last outer block's: long K) on the caller's
// stream while the previous block's panel chain still runs on the high-priority stream, part 2 (the last outer block,
// K = 512) on the chain itself.
//   hi:  part2(j) [after part1(j)] | panels(j) | slice(j) | part2(j+1) ...
//   lo:  part1(j+1) [after slice(j-1)]                     | part1(j+2) ...
static int potrf_outer_oz(int nblk);
static cudaError_t potrf_blocked_oz(const Launch& L, const MatSet& ms, const int* active, int* info) {
  const int nblk = ms.np / TILE;
  // panels per outer block of the left-looking form: 2 up to n = 12 k (the persistent int8 kernel makes K = 256 passes cheap
  // and fewer in-block updates stay on the FP64 pipe: -3 % at n = 6400 x 40, -4 % x 8), 4 above (one big matrix: fewer, larger
  // launches beside the panel chain: n = 16384 28.1 ms against 30.2).  A function of n ONLY: the blocking fixes the
  // summation order, and results must not depend on how many other matrices share the batch
  // (tests/test_gpu_parity.py::test_chunked_batches_equal_one_chunk).  profiles/r02_summary.md section 7
  const int OB = potrf_outer_oz(nblk);
  AGPC_CUDA_TRY(oz_row_bound_launch(L, ms.A, ms.np, ms.stride, ms.np, ms.batch, ms.scA, ms.np, active));
  const bool lookahead = L.side != nullptr && nblk > 2 * OB;
  if (!lookahead) {
    for (int k0 = 0; k0 < nblk; k0 += OB) {
      const int kb = nblk - k0 < OB ? nblk - k0 : OB;
      const bool thin = (nblk - k0 - 1) * ms.batch <= THIN_CTA_LIMIT;
      if (k0 > 0) AGPC_CUDA_TRY(potrf_update_oz(L, ms, active, k0, kb, 0, k0, thin));
      AGPC_CUDA_TRY(potrf_panels(L, ms, active, info, k0, kb, k0 > 0 && thin));
      if (k0 + kb < nblk) AGPC_CUDA_TRY(potrf_slice_panel(L, ms, active, k0, kb, k0 + kb));
    }
    return cudaSuccess;
  }
  Launch H = L;
  H.stream = L.side;
  AGPC_CUDA_TRY(cudaEventRecord(L.ev_rest, L.stream));  // everything queued so far (form_B, the row bounds) precedes the chain
  AGPC_CUDA_TRY(cudaStreamWaitEvent(H.stream, L.ev_rest, 0));
  for (int k0 = 0; k0 < nblk; k0 += OB) {
    const int kb = nblk - k0 < OB ? nblk - k0 : OB;
    const int r0 = k0 + kb;
    const bool thin = (nblk - k0 - 1) * ms.batch <= THIN_CTA_LIMIT;
    // (a) the chain needs part 1 of THIS block column (queued on the caller's stream one iteration ago)
    if (k0 > OB) AGPC_CUDA_TRY(cudaStreamWaitEvent(H.stream, L.ev_rest, 0));
    // (b) part 1 of the NEXT block column: all panels before this block; their slices were completed with ev_panel
    if (k0 > 0 && r0 < nblk) {
      const int kbn = nblk - r0 < OB ? nblk - r0 : OB;
      AGPC_CUDA_TRY(cudaStreamWaitEvent(L.stream, L.ev_panel, 0));
      // (AGPC_OZ_KCAP = c > 0 cuts the K range into segments of c panels: shorter tiles free SMs for the chain sooner,
      //  at the price of one more epilogue pass per segment)
      static const int kcap = getenv("AGPC_OZ_KCAP") ? atoi(getenv("AGPC_OZ_KCAP")) : 0;
      if (kcap > 0) {
        for (int kc = 0; kc < k0; kc += kcap) AGPC_CUDA_TRY(potrf_update_oz(L, ms, active, r0, kbn, kc, k0 - kc < kcap ? k0 - kc : kcap, false));
      } else {
        AGPC_CUDA_TRY(potrf_update_oz(L, ms, active, r0, kbn, 0, k0, false, true));
      }
      AGPC_CUDA_TRY(cudaEventRecord(L.ev_rest, L.stream));
    }
    // (c) part 2 of this block column: the previous outer block (sliced at the end of the last iteration)
    if (k0 > 0) AGPC_CUDA_TRY(potrf_update_oz(H, ms, active, k0, kb, k0 - OB, OB, thin));
    // (d) factor it, slice it
    AGPC_CUDA_TRY(potrf_panels(H, ms, active, info, k0, kb, k0 > 0 && thin));
    if (r0 >= nblk) break;
    AGPC_CUDA_TRY(potrf_slice_panel(H, ms, active, k0, kb, r0));
    AGPC_CUDA_TRY(cudaEventRecord(L.ev_panel, H.stream));
  }
  AGPC_CUDA_TRY(cudaEventRecord(L.ev_panel, H.stream));  // join: the caller's stream continues after the chain
  AGPC_CUDA_TRY(cudaStreamWaitEvent(L.stream, L.ev_panel, 0));
  return cudaSuccess;
}

static int potrf_outer_oz(int nblk) {
  static const int ob_env = getenv("AGPC_POTRF_OUTER") ? atoi(getenv("AGPC_POTRF_OUTER")) : 0;
  return ob_env > 0 ? ob_env : (nblk < 96 ? 2 : POTRF_OUTER);
}

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
  if (oz_on(ms)) {  // the panel was sliced into SA by potrf_slice_panel
    OzGemmArgs z{};
    z.A = oz_operand(ms, 0); z.B = z.A;
    z.C = ms.A; z.Ct = nullptr; z.ldc = ms.np; z.c_batch = ms.stride;
    z.a_row0 = r0 * TILE; z.a_col0 = k0 * TILE; z.b_row0 = r0 * TILE; z.b_col0 = k0 * TILE;
    z.c_row0 = r0 * TILE; z.c_col0 = r0 * TILE;
    z.k_chunks = kb * OZ_CPB; z.k_rule = KRULE_FULL; z.alpha = -1.0; z.beta = 1.0; z.active = active;
    if (ncols < 0) {
      z.m_tiles = rows; z.n_tiles = rows; z.lower_only = 1;
      return oz_gemm_launch(L, z, rows * (rows + 1) / 2, 1, ms.batch, ms.oz_S);
    }
    const int nc = ncols < rows ? ncols : rows;
    z.m_tiles = rows; z.n_tiles = nc; z.lower_only = 0; z.skip_upper = 1;
    AGPC_CUDA_TRY(oz_gemm_launch(L, z, rows * nc, 1, ms.batch, ms.oz_S));
    if (stage_first_col && rows > 1)  // the thin panel solve reads its panel from the M scratch
      AGPC_CUDA_TRY(copy_panel_launch(L, ms.A, ms.M, ms.np, ms.stride, (r0 + 1) * TILE, r0 * TILE, (rows - 1) * TILE, ms.batch, active));
    return cudaSuccess;
  }
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
  static const bool right_looking = getenv("AGPC_OZ_RIGHT") && atoi(getenv("AGPC_OZ_RIGHT")) != 0;  // round-2 first form
  if (oz_on(ms) && !right_looking) return potrf_blocked_oz(L, ms, active, info);
  const bool lookahead = L.side != nullptr && nblk > 2 * POTRF_OUTER;
  if (!lookahead) {
    for (int k0 = 0; k0 < nblk; k0 += POTRF_OUTER) {
      const int kb = nblk - k0 < POTRF_OUTER ? nblk - k0 : POTRF_OUTER;
      AGPC_CUDA_TRY(potrf_panels(L, ms, active, info, k0, kb, false));
      if (k0 + kb < nblk) {
        if (oz_on(ms)) AGPC_CUDA_TRY(potrf_slice_panel(L, ms, active, k0, kb, k0 + kb));
        AGPC_CUDA_TRY(potrf_trailing(L, ms, active, k0, kb, k0 + kb, -1));
      }
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
    if (oz_on(ms)) {
      // the slices of this block's panel replace those of the previous one, which rest(i-1) may still be reading
      if (rest_pending) AGPC_CUDA_TRY(cudaStreamWaitEvent(H.stream, L.ev_rest, 0));
      AGPC_CUDA_TRY(potrf_slice_panel(H, ms, active, k0, kb, r0));
    }
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
  L.add_work(CLS_GEMM_TRTRI, (double)L.work_items * wn * wn * wn / 3.0);  // dtrtri
  Launch Lt = L;
  Lt.gemm_cls = CLS_GEMM_TRTRI;
  static const bool right_looking = getenv("AGPC_OZ_RIGHT") && atoi(getenv("AGPC_OZ_RIGHT")) != 0;
  static const bool keep_env = !(getenv("AGPC_OZ_KEEP_L") && atoi(getenv("AGPC_OZ_KEEP_L")) == 0);
  const bool keep_L = oz_on(ms) && ms.SC != nullptr && !right_looking && keep_env;
  const int OBp = potrf_outer_oz(nblk);
  static const bool rowmax_env = !(getenv("AGPC_OZ_ROWMAX") && atoi(getenv("AGPC_OZ_ROWMAX")) == 0);
  // Running row maxima of M and U (rmM, rmU): every entry of row i of U (of M) that exists when level h slices U11 (M22) lies
  // inside that level's range -- the earlier levels' groups are sub-groups of the half the row belongs to -- so the maximum
  // over everything produced so far IS the maximum of the range, and the slice launches of U11 / M22 need no pass of their
  // own over the data (AGPC_OZ_ROWMAX=2 turns this part off).  Producers: oz_rowmax_init_launch after the FP64 levels
  // (diagonal blocks + level h = 1), then the step-2 epilogue of every int8 level (rows of M21, columns of it for U12).
  static const bool run_env = !(getenv("AGPC_OZ_ROWMAX") && atoi(getenv("AGPC_OZ_ROWMAX")) == 2);
  const bool running = oz_on(ms) && ms.rmM != nullptr && ms.rmU != nullptr && rowmax_env && run_env && oz_gemm_rowmax_supported() &&
                       nblk > OZ_TRTRI_MIN_H;
  bool running_ready = false;
  for (int h = 1; h < nblk; h *= 2) {
    const int G = 2 * h;
    const int groups = (nblk + G - 1) / G;
    if (oz_on(ms) && h >= OZ_TRTRI_MIN_H) {
      const double blk = (double)TILE * TILE * 8.0 * Lt.work_items;
      // step 1: T[left rows, right cols] = U11 * L21^T  (U11 upper triangular: K from the tile row on)
      // L21 (rows of the right half, columns of the left half) lies below the outer blocks of its columns whenever h is a
      // multiple of the potrf outer block: the left-looking potrf has already sliced it into SA (static row scales), and
      // with a third image for trtri's own A operands those slices survive -- a third of this level's slicing traffic
      const bool l_kept = keep_L && h % OBp == 0;
      const int wa = keep_L ? 2 : 0;   // image of the A operands (U11, M22)
      if (running && !running_ready) {  // first int8 level: the state the FP64 levels left behind
        AGPC_CUDA_TRY(oz_rowmax_init_launch(Lt, ms.M, ms.U, ms.np, ms.stride, ms.np, h * TILE, ms.rmM, ms.rmU, ms.batch, active));
        running_ready = true;
      }
      OzSliceArgs su = oz_slice_args(ms, ms.U, wa, active);
      if (running) su.given_max = ms.rmU;
      su.row0 = 0; su.col0 = 0; su.row_blocks = h; su.col_chunks = h * OZ_CPB; su.tri = 2;
      su.grp_blocks = G; su.half_blocks = h; su.nblk = nblk; su.need_h2 = 1;
      AGPC_CUDA_TRY(oz_slice_launch(Lt, su, h, groups, ms.batch, ms.oz_S));
      if (!l_kept) {
        OzSliceArgs sl = oz_slice_args(ms, ms.A, 1, active);
        sl.row0 = h * TILE; sl.col0 = 0; sl.rows_kind = EXT_H2; sl.col_chunks = h * OZ_CPB;
        sl.grp_blocks = G; sl.half_blocks = h; sl.nblk = nblk; sl.need_h2 = 1;
        AGPC_CUDA_TRY(oz_slice_launch(Lt, sl, h, groups, ms.batch, ms.oz_S));
      }
      Lt.add_work(CLS_SLICE, blk * (l_kept ? 0.5 : 1.5) * h * h * groups);
      OzGemmArgs a{};
      a.A = oz_operand(ms, wa); a.B = oz_operand(ms, l_kept ? 0 : 1);
      a.C = ms.T; a.Ct = nullptr; a.ldc = ms.np; a.c_batch = ms.stride;
      a.a_row0 = 0; a.a_col0 = 0; a.b_row0 = h * TILE; a.b_col0 = 0; a.c_row0 = 0; a.c_col0 = h * TILE;
      a.m_tiles = h; a.n_tiles = 0; a.k_chunks = h * OZ_CPB;
      a.m_kind = EXT_FIXED; a.n_kind = EXT_H2; a.k_kind = EXT_FIXED;
      a.grp_blocks = G; a.half_blocks = h; a.nblk = nblk;
      a.k_rule = KRULE_BEGIN_AT_ROW; a.alpha = 1.0; a.beta = 0.0; a.active = active;
      // the epilogue collects max |T_ij| per row, so the slicing of T below needs no row-maximum pass over T
      const bool t_max = ms.rmax != nullptr && rowmax_env && oz_gemm_rowmax_supported();
      if (t_max) {
        AGPC_CUDA_TRY(cudaMemsetAsync(ms.rmax, 0, sizeof(double) * (size_t)ms.np * ms.batch, Lt.stream));
        a.row_absmax = ms.rmax; a.absmax_batch = ms.np;
      }
      AGPC_CUDA_TRY(oz_gemm_launch(Lt, a, h * h, groups, ms.batch, ms.oz_S));
      // step 2: M21 = -M22 * T^T ; U12 = M21^T  (M22 lower triangular: K up to the tile row)
      OzSliceArgs sm = oz_slice_args(ms, ms.M, wa, active);
      if (running) sm.given_max = ms.rmM;
      sm.row0 = h * TILE; sm.col0 = h * TILE; sm.rows_kind = EXT_H2; sm.cols_kind = EXT_H2; sm.tri = 1;
      sm.grp_blocks = G; sm.half_blocks = h; sm.nblk = nblk; sm.need_h2 = 1;
      AGPC_CUDA_TRY(oz_slice_launch(Lt, sm, h, groups, ms.batch, ms.oz_S));
      OzSliceArgs st = oz_slice_args(ms, ms.T, 1, active);
      if (t_max) st.given_max = ms.rmax;
      st.row0 = 0; st.col0 = h * TILE; st.row_blocks = h; st.cols_kind = EXT_H2;
      st.grp_blocks = G; st.half_blocks = h; st.nblk = nblk; st.need_h2 = 1;
      AGPC_CUDA_TRY(oz_slice_launch(Lt, st, h, groups, ms.batch, ms.oz_S));
      Lt.add_work(CLS_SLICE, blk * 1.5 * h * h * groups);
      OzGemmArgs c{};
      c.A = oz_operand(ms, wa); c.B = oz_operand(ms, 1);
      c.C = ms.M; c.Ct = ms.U; c.ldc = ms.np; c.c_batch = ms.stride;
      c.a_row0 = h * TILE; c.a_col0 = h * TILE; c.b_row0 = 0; c.b_col0 = h * TILE;
      c.c_row0 = h * TILE; c.c_col0 = 0; c.ct_row0 = 0; c.ct_col0 = h * TILE;
      c.m_tiles = 0; c.n_tiles = h; c.k_chunks = 0;
      c.m_kind = EXT_H2; c.n_kind = EXT_FIXED; c.k_kind = EXT_H2;
      c.grp_blocks = G; c.half_blocks = h; c.nblk = nblk;
      c.k_rule = KRULE_END_AT_ROW; c.alpha = -1.0; c.beta = 0.0; c.active = active;
      if (running) {  // the new blocks M21 (rows) and U12 = M21^T (columns of the product) enter the running maxima
        c.row_absmax = ms.rmM; c.col_absmax = ms.rmU; c.absmax_batch = ms.np;
      }
      AGPC_CUDA_TRY(oz_gemm_launch(Lt, c, h * h, groups, ms.batch, ms.oz_S));
      continue;
    }
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
    AGPC_CUDA_TRY(gemm_nt_launch(Lt, ms.mapU, ms.mapA, a, h * h, groups, ms.batch));
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
    AGPC_CUDA_TRY(gemm_nt_launch(Lt, ms.mapM, ms.mapT, c, h * h, groups, ms.batch));
  }
  return cudaSuccess;
}

// dst (lower tiles) = U U^T = B^-1
cudaError_t lauum_blocked(const Launch& L, const MatSet& ms, double* dst, const int* active) {
  const int nblk = ms.np / TILE;
  const double wn = L.work_n > 0.0 ? L.work_n : (double)ms.np;
  L.add_work(CLS_GEMM_LAUUM, (double)L.work_items * wn * wn * wn / 3.0);  // dlauum
  Launch Ll = L;
  Ll.gemm_cls = CLS_GEMM_LAUUM;
  if (oz_on(ms)) {
    OzSliceArgs su = oz_slice_args(ms, ms.U, 0, active);
    su.row_blocks = nblk; su.col_chunks = nblk * OZ_CPB; su.tri = 2;
    Ll.add_work(CLS_SLICE, 4.0 * wn * wn * Ll.work_items);
    AGPC_CUDA_TRY(oz_slice_launch(Ll, su, nblk, 1, ms.batch, ms.oz_S));
    OzGemmArgs z{};
    z.A = oz_operand(ms, 0); z.B = z.A;
    z.C = dst; z.Ct = nullptr; z.ldc = ms.np; z.c_batch = ms.stride;
    z.m_tiles = nblk; z.n_tiles = nblk; z.k_chunks = nblk * OZ_CPB;
    z.k_rule = KRULE_BEGIN_AT_ROW; z.lower_only = 1; z.alpha = 1.0; z.beta = 0.0; z.active = active;
    return oz_gemm_launch(Ll, z, nblk * (nblk + 1) / 2, 1, ms.batch, ms.oz_S);
  }
  GemmArgs a{};
  a.C = dst; a.Ct = nullptr; a.ldc = ms.np; a.c_batch = ms.stride;
  a.m_tiles = nblk; a.n_tiles = nblk; a.k_tiles = nblk * (TILE / GEMM_BK);
  a.k_rule = KRULE_BEGIN_AT_ROW; a.lower_only = 1; a.alpha = 1.0; a.beta = 0.0; a.active = active;
  return gemm_nt_launch(Ll, ms.mapU, ms.mapU, a, nblk * (nblk + 1) / 2, 1, ms.batch);
}

}  // namespace agpc
