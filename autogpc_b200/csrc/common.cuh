// Shared declarations of the autogpc_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/autogpc_b200.h"

namespace agpc {

constexpr int TILE = 128;       // factorisation block = GEMM CTA tile edge; every matrix is padded to a multiple
constexpr int GEMM_BK = 16;     // K-chunk of the GEMM pipeline: 16 doubles = one 128-byte TMA swizzle span
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_CONSUMER_WARPS = 16;
constexpr int GEMM_THREADS = (GEMM_CONSUMER_WARPS + 1) * 32;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * 2 * TILE * GEMM_BK * 8 + 1024 /*align slack*/ + 128 /*barriers*/;

// What a GEMM tile's extents / K-range depend on (see gemm_nt.cu).
enum : int { EXT_FIXED = 0, EXT_H2 = 1 };
enum : int { KRULE_FULL = 0, KRULE_BEGIN_AT_ROW = 1, KRULE_END_AT_ROW = 2, KRULE_END_AT_COL = 3 };

struct GemmArgs {
  double* C;                 // output (row-major), element (r, c) of batch b at C + b*c_batch + r*ldc + c
  double* Ct;                // optional second, transposed output: Ct + b*c_batch + c*ldc + r   (nullptr = none)
  double* C2;                // optional duplicate of C (same layout and offsets) for the tiles with tj < c2_cols and
  int c2_cols;               //   ti >= c2_row_min: stages the next panel for the thin solve without a copy kernel
  int c2_row_min;
  long long ldc, c_batch;
  int a_row0, a_col0, b_row0, b_col0, c_row0, c_col0, ct_row0, ct_col0;  // element offsets (group-relative)
  int m_tiles, n_tiles, k_tiles;  // extents when EXT_FIXED: tiles of 128 x 128 x 16
  int m_kind, n_kind, k_kind;     // EXT_FIXED or EXT_H2 (right-half size of trtri group g, may be partial)
  int grp_blocks, half_blocks, nblk;  // trtri grouping: group g covers blocks [g*grp_blocks, ...)
  int k_rule, lower_only;
  int skip_upper;  // rectangular tile grid whose C origin lies on the diagonal: tiles with ti < tj are not computed
  double alpha, beta;
  const int* active;  // per-batch-item flag (nullptr = all active)
};

// Kernel classes of the optional per-launch event timing (agpc_profile_*): bench.py's roofline and the share table.
enum : int { CLS_KBUILD = 0, CLS_GEMM, CLS_POTF2, CLS_MATVEC, CLS_EP, CLS_GRAD, CLS_FORMB, CLS_MISC, CLS_GEMM_TRTRI, CLS_GEMM_LAUUM,
              CLS_SLICE, CLS_OZ_POTRF, CLS_OZ_TRTRI, CLS_OZ_LAUUM, CLS_COUNT };  // CLS_GEMM = the tile GEMM inside potrf (and predict); _TRTRI / _LAUUM = the same kernel in
                                       // those phases; CLS_SLICE = operand slicing for the int8 tensor-core GEMM (ozaki.cu);
                                       // CLS_OZ_* = that GEMM per phase: event time, work = EXECUTED int8 operations (2 x MACs)

// When enabled, every launcher brackets its kernel with a cudaEvent pair on the launching stream; the pairs are
// summed per class by agpc_profile_read.  `work` is the ALGORITHMIC work the drivers attribute to a class
// (flops for CLS_GEMM / CLS_POTF2, bytes for the others), not what the kernels happen to execute.
struct Profiler {
  struct Rec {
    cudaEvent_t e0, e1;
    int cls;
  };
  bool enabled = false;
  cudaEvent_t base = nullptr;  // recorded at enable(): launch intervals are measured from it
  std::vector<Rec> recs;
  size_t used = 0;
  double work[CLS_COUNT] = {0};
  long long count[CLS_COUNT] = {0};
  int open(int cls, cudaStream_t st) {
    if (used == recs.size()) {
      Rec r;
      if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
      recs.push_back(r);
    }
    recs[used].cls = cls;
    cudaEventRecord(recs[used].e0, st);
    return (int)used++;
  }
  void close(int idx, cudaStream_t st) { cudaEventRecord(recs[idx].e1, st); }
};

// Launchers (host).  Every launcher bumps *launch_counter when non-null.
struct Launch {
  cudaStream_t stream;
  long long* counter;
  Profiler* prof = nullptr;
  cudaStream_t side = nullptr;        // second stream + two events for the look-ahead of single-matrix factorisations
  cudaEvent_t ev_panel = nullptr, ev_rest = nullptr;
  int gemm_cls = CLS_GEMM;  // profiler class the tile GEMM is booked under (potrf / trtri / lauum phase)
  int work_items = 1;     // items still active in the batch      } algorithmic work accounting only
  double work_n = 0.0;    // mean true n of the batch's items     }
  void bump(int n = 1) const {
    if (counter) *counter += n;
  }
  void add_work(int cls, double w) const {
    if (prof && prof->enabled) prof->work[cls] += w;
  }
};
struct ProfScope {
  const Launch& L;
  int idx = -1;
  ProfScope(const Launch& l, int cls) : L(l) {
    if (L.prof && L.prof->enabled) {
      idx = L.prof->open(cls, L.stream);
      L.prof->count[cls] += 1;
    }
  }
  ~ProfScope() {
    if (idx >= 0) L.prof->close(idx, L.stream);
  }
};

// gemm_nt.cu
cudaError_t gemm_nt_launch(const Launch& L, const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmArgs& args,
                           int tiles_x, int groups, int batch);
// Encodes a 3-D FP64 tensor map {cols, rows, batch} with a {16, 128, 1} box and 128-byte swizzle.
int make_tensor_map(CUtensorMap* out, const double* base, long long cols, long long rows, long long batch,
                    long long ld, long long batch_stride, int box_rows = TILE);
cudaError_t gemm_nt_thin_launch(const Launch& L, const CUtensorMap& mapA, const CUtensorMap& mapB32,
                                const GemmArgs& args, int tiles_x, int groups, int batch);

// ---- ozaki.cu: FP64-accurate tile GEMM on tcgen05 kind::i8 (operands as S signed 8-bit digit slices) ---------------
constexpr int OZ_KB = 32;  // K extent (= bytes) of one operand tile: one 32-byte swizzle span, one kind::i8 MMA deep
constexpr int OZ_BN = 64;  // C tile is 128 x 64: S accumulators x 64 columns must fit the 512 TMEM columns
// Sliced image of one FP64 operand matrix: tile (rb, kc, p) -- rows 128 rb .. +128, columns 32 kc .. +32, digit slice p --
// is the 4 KB block at S + ((rb * nkc + kc) * n_slices + p) * 4096, already in the swizzled K-major layout the MMA
// reads; scale[row] = 2^(e_row - 8).  Tiles are addressed by the coordinates of the source matrix.
struct OzOperand {
  const int8_t* S;
  const double* scale;
  long long s_batch, sc_batch;  // batch strides (bytes / elements)
  int nkc;                      // K chunks per row block = columns of the source matrix / 32
};
struct OzGemmArgs {  // same tile / group / K-range semantics as GemmArgs; n_tiles counts 128-wide tile columns
  OzOperand A, B;
  double* C;
  double* Ct;
  long long ldc, c_batch;
  int a_row0, a_col0, b_row0, b_col0, c_row0, c_col0, ct_row0, ct_col0;
  int m_tiles, n_tiles, k_chunks;  // k_chunks: K / 32
  int m_kind, n_kind, k_kind;
  int grp_blocks, half_blocks, nblk;
  int k_rule, lower_only, skip_upper;
  double alpha, beta;
  const int* active;
  double* row_absmax;      // optional: max |C_ij| over each row of the output region (what is STORED, beta included), by
  long long absmax_batch;  //   atomic max into row_absmax[b * absmax_batch + row]; the caller zeroes it first.  NaN wins
                           //   (bit pattern above every finite value), so a non-finite entry poisons its row as in FP64
  double* col_absmax;      // optional, with Ct: max |Ct_row| (= column maxima of the product), same indexing by Ct row
  int concurrent;  // host hint: this launch shares the GPU with a latency-bound chain on another stream (look-ahead)
  int dbg;  // measurement aid (AGPC_OZ_DBG): 1 = main loop without MMAs (data movement only), 2 = without loads (MMAs on
            // whatever the ring holds); results are garbage in both
};
struct OzSliceArgs {
  const double* X;  // source region: rows row0 .. , columns col0 .. (group g adds g * grp_blocks * 128 to both)
  long long ld, x_batch;
  int8_t* S;
  double* scale;
  long long s_batch, sc_batch;
  int nkc;
  int row0, col0;
  int row_blocks, col_chunks;  // extent (128-row blocks, 32-column chunks) when the kinds are EXT_FIXED
  int rows_kind, cols_kind;    // EXT_H2: right-half size of trtri group g
  int grp_blocks, half_blocks, nblk;
  int need_h2;  // grouped launches: skip groups without a right half (their left half may lie outside the matrix)
  int col_split;              // > 1 (only with fixed_scale / given_max, i.e. no row-maximum pass): the chunk range of a row
  int grid_rows;              //   block is cut into col_split parts, one CTA each: blockIdx.x = part * grid_rows + row block
  const double* given_max;    // non-null: max |x| of every row over the sliced range is GIVEN (indexed like scale[], e.g.
                              // OzGemmArgs::row_absmax of the product that wrote X): no row-maximum pass over the data
  const double* fixed_scale;  // non-null: the row scales are GIVEN (scale[] already holds 2^(e - 8) from a bound on the row):
                              // no row-maximum pass, slices made with different calls share their scales
  int tri;  // 0: whole rectangle; 1: chunks up to the end of the row block's own diagonal block (lower triangular
            // region); 2: chunks from the start of it (upper triangular).  The row scale covers exactly that range.
  const int* active;
};
struct Launch;
cudaError_t oz_gemm_launch(const Launch& L, const OzGemmArgs& a, int tiles128, int groups, int batch, int S);
cudaError_t oz_slice_launch(const Launch& L, const OzSliceArgs& a, int row_blocks, int groups, int batch, int S);
// rmM[i] = max_{k in [G0 * (i / G0), i]} |M[i][k]|, rmU[i] = max_{k in [i, G0 * (i / G0) + G0)} |U[i][k]|: the state of the running
// maxima after the levels that ran without the epilogue hooks (diagonal blocks + FP64 levels; G0 = their group width)
cudaError_t oz_rowmax_init_launch(const Launch& L, const double* M, const double* U, long long ld, long long batch_stride, int np,
                                  int G0, double* rmM, double* rmU, int batch, const int* active);
bool oz_gemm_rowmax_supported();  // false when an experimental kernel form without the row_absmax epilogue is selected
void oz_release_stream(cudaStream_t st);  // persistent kernel's per-stream tile counter: give it back before destroying st
// scale[b][i] = 2^(e_i - 8) with 2^e_i > 4 sqrt(A_ii): |L_ik| <= sqrt(A_ii) for every entry of row i of the Cholesky factor
cudaError_t oz_row_bound_launch(const Launch& L, const double* A, long long ld, long long a_batch, int n, int batch,
                                double* scale, long long sc_batch, const int* active);

// One lock-step batch of padded np x np matrices (np % 128 == 0, ld == np, matrix b at base + b*stride) and the
// tensor maps the GEMM kernel reads them through.
//   K: kernel matrix            A: B = I + S^1/2 K S^1/2 -> L (lower)        M: L^-1 (lower)
//   U: L^-T (upper)             T: scratch (trtri) / B^-1 (lauum output)
struct MatSet {
  int np = 0, batch = 0;
  long long stride = 0;
  double *K = nullptr, *A = nullptr, *M = nullptr, *U = nullptr, *T = nullptr;
  // int8 tensor-core path (ozaki.cu): two slice buffers of oz_S * np * np bytes per item + their row scales; oz_S = 0
  // keeps every product on the FP64 DMMA kernel
  int oz_S = 0;
  int8_t *SA = nullptr, *SB = nullptr;
  double *scA = nullptr, *scB = nullptr;
  // optional third image: when present, trtri puts its A operands (U11, M22) here and takes L21 from the slices of L that
  // the left-looking potrf left in SA (statically scaled) instead of slicing L a second time
  int8_t* SC = nullptr;
  double* scC = nullptr;
  double* rmax = nullptr;  // optional [batch][np]: row maxima collected by a GEMM epilogue for the slicing of its output
  double *rmM = nullptr, *rmU = nullptr;  // optional [batch][np]: running max |M_ik| / |U_ik| over the part of row i that trtri has
                                          // produced so far (kept by its producers): exactly the range the next level slices
  CUtensorMap mapA, mapM, mapU, mapT;
  CUtensorMap mapA32, mapM32;  // 32-row boxes: B operands of the thin GEMM variant
};

// chol.cu
cudaError_t potrf_blocked(const Launch& L, const MatSet& ms, const int* active, int* info);
cudaError_t trtri_blocked(const Launch& L, const MatSet& ms, const int* active);
cudaError_t lauum_blocked(const Launch& L, const MatSet& ms, double* dst, const int* active);
cudaError_t potf2_inv_launch(const Launch& L, double* A, double* M, double* U, long long ld, long long batch_stride,
                             int diag_block, int batch, const int* active, int* info);

// kbuild.cu
struct BuildArgs {
  const agpc_program* progs;   // [batch] (device)
  const double* theta;         // [batch][theta_stride]
  int theta_stride;
  const double* Xa;            // row points  [batch][na_pad][D]
  const double* Xb;            // col points  [batch][nb_pad][D]
  const int* na_of;            // [batch] valid rows (nullptr -> na_fixed)
  const int* nb_of;            // [batch] valid cols (nullptr -> nb_fixed)
  int na_fixed, nb_fixed;
  int na_pad, nb_pad, D;
  long long xa_batch, xb_batch;  // batch strides of Xa / Xb in elements
  double* out;                   // [batch][na_pad][ld]
  long long ld, out_batch;
  double* dK;                    // optional [batch][n_theta][na_pad][ld]
  double* kdiag;                 // optional [batch][na_pad] (only meaningful when Xa == Xb)
  const int* active;
  int dk_tiled;                  // materialised dK: every leaf sits in exactly one term -> streaming tiled kernel
};
struct GradArgs {
  const agpc_program* progs;
  const double* theta;
  int theta_stride;
  const double* Xf;    // [batch][np][D]
  const int* n_of;     // valid n per item (nullptr -> n_fixed)
  int n_fixed, np, D;
  const double* G;     // B^-1 lower tiles (mode 0) or a full weight matrix (mode 1): [batch][np][ld]
  long long ld, g_batch;
  const double* alpha; // [batch][np]   (mode 0)
  const double* s;     // [batch][np]   (mode 0)
  const double* ru;    // optional (mode 0): extra symmetric rank-2 weight 1/2 (ru_i rv_k + ru_k rv_i)  -- the implicit
  const double* rv;    //   term of the Laplace gradient ([R&W] 5.24)
  int mode;
  double* partial;     // [batch][tiles][AGPC_MAX_THETA]
  int tiles_per_item;
  const int* active;
};
cudaError_t build_K_launch(const Launch& L, const BuildArgs& a, int batch, bool grad);
cudaError_t kself_launch(const Launch& L, const agpc_program* progs, const double* theta, int theta_stride,
                         const double* Xs, int m_pad, int D, const int* m_of, int batch, double* kss);
cudaError_t gather_rows_launch(const Launch& L, const double* X, const double* y, const int* rows,
                               const long long* row_off, const int* n_of, int D, int np, int batch, double* Xf,
                               double* yf);
cudaError_t form_B_launch(const Launch& L, const double* K, const double* s, double* A, int np, long long stride,
                          int batch, const int* active, const double* jitter = nullptr);
cudaError_t jitter_step_mat_launch(const Launch& L, const double* A0, double* A, int n, long long stride, int batch, int* info,
                                   double* jitter, int* tries, int* retry);
cudaError_t jitter_step_launch(const Launch& L, const double* kdiag, const double* s, const int* n_of, int np, int batch,
                               int* info, double* jitter, int* tries, int* retry, int* used, const int* active);
cudaError_t scale_cols_launch(const Launch& L, double* Ks, const double* s, int rows, int np, int batch);
cudaError_t grad_contract_launch(const Launch& L, const GradArgs& a, int batch, double scale, double* grad,
                                 int grad_stride);
int grad_tiles(int np, int mode);
cudaError_t dmu_dx_launch(const Launch& L, const agpc_program* prog, const double* theta, const double* Xs, int m,
                          const double* Xf, int n, int D, const double* alpha, double* out);

// ep.cu
cudaError_t rowdot_launch(const Launch& L, int range, bool sq, const double* A, const double* x, double* y,
                          double* y2, int nrows, int ncols, long long a_batch, long long x_batch, long long y_batch,
                          int batch, const int* active, bool strict_sq = false);
cudaError_t sqrt_scale_launch(const Launch& L, const double* tau, const double* w, double* s, double* u, int np,
                              int batch, const int* active);
cudaError_t mul_launch(const Launch& L, const double* a, const double* b, double* out, int np, int batch,
                       const int* active);
cudaError_t marginals_launch(const Launch& L, const double* tau, const double* usq, const double* lsq, const double* Lmat,
                             long long l_stride, const double* w, const double* kv, const double* kdiag, double* sig,
                             double* mu, const int* n_of, int np, int batch, int cold, const int* active);
cudaError_t ep_update_launch(const Launch& L, double* tau, double* nu, const double* sig, const double* mu,
                             const double* yf, const double* kdiag, const int* n_of, int np, int batch, double eps,
                             double damping, int min_sweeps, int* active, int* sweeps, int* status, int* converged,
                             const int* info);
cudaError_t ep_final_launch(const Launch& L, const double* tau, const double* nu, const double* sig, const double* mu,
                            const double* yf, const double* s, const double* z, const double* Lmat, long long stride,
                            const int* n_of, int np, int batch, const agpc_program* progs, double* alpha, double* nlml,
                            double* nlml_gauss, double* bic, int* status, const int* info);
cudaError_t laplace_lik_launch(const Launch& L, const double* f, const double* yf, double* W, double* s, double* b,
                               double* g1, double* d3, const int* n_of, int np, int batch, const int* active);
cudaError_t laplace_step_launch(const Launch& L, double* f, double* a, const double* f_new, const double* a_new,
                                const double* yf, double* psi, const int* n_of, int np, int batch, double tol,
                                int* active, int* iters, int* status, const int* info);
cudaError_t laplace_final_launch(const Launch& L, const double* f, const double* a, const double* W, const double* g1,
                                 const double* d3, const double* dB, const double* yf, const double* Lmat,
                                 long long stride, const int* n_of, int np, int batch, const agpc_program* progs,
                                 double* s2, double* nlml, double* nlml_gauss, double* bic, int* status, const int* info);
cudaError_t axpy_launch(const Launch& L, const double* x, const double* y, double alpha, double* out, int np, int batch);
cudaError_t predict_prob_launch(const Launch& L, const double* mu_s, const double* kss, const double* vsq, double* p,
                                double* var, long long total);

}  // namespace agpc

namespace agpc {
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device context: one "done" flag per (call site, device), so
// a second Engine(device = k) in the same process configures its own context instead of failing at launch
// (internal linkage on purpose: every translation unit gets its own table, so the call sites -- identified by their line
// number -- of different files can never alias)
static inline bool& func_attr_done(int site) {
  static bool done[64][2048] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  return done[dev & 63][site & 2047];
}
}  // namespace agpc

#define AGPC_CUDA_TRY(expr)                         \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) return _e;               \
  } while (0)
