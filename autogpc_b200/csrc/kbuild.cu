// Pairwise kernel-matrix kernels: K(X, X'), materialised dK/dtheta, and the fused gradient contraction.
//
// Stands in for GPy kern.K / kern.update_gradients_full of RBF, StdPeriodic, Bias, Add, Prod (+ the Linear extension)
// -- SURVEY 8a G1/G2 -- as reached from GPClassification's constructor and optimize() (gpckernel.py:245-246) and
// from predict (gpckernel.py:832).  Leaf formulas (theta order = gpss2gpy, gpckernel.py:535-565):
//   SE    k = v exp(-r^2 / 2l^2)                     dk/dv = k/v   dk/dl = k r^2 / l^3
//   PER   k = v exp(-sin^2(pi r / T) / 2l^2)         dk/dv = k/v   dk/dT = k sin b cos b (b/T) / l^2   dk/dl = k sin^2 b / l^3
//   LIN   k = v (x - c)(x' - c)                      dk/dv = k/v   dk/dc = -v (x + x' - 2c)
//   CONST k = v                                      dk/dv = 1
// A product term is evaluated as (prod of scale factors) * exp(sum of exponents): one exp per term, not per leaf.
//
// Kernels here:
//   build_K_tiled_kernel     K only (the scoring path): 64 x 64 tiles, 2 x 4 pairs per thread and pass, symmetric
//                            (lower tiles + mirrored store through shared memory) when both point sets coincide
//   build_KdK_tiled_kernel   K + materialised dK/dtheta, streaming: slabs written as each term is evaluated
//   build_K_kernel<true>     general materialised dK (leaves shared between terms): 32 x 128 tiles, per-pair interpreter
//   grad_contract_kernel     fused <W, dK/dtheta> contraction (never materialises dK): what score_batch uses
// Point coordinates of a tile live in shared memory, dimension-major (column reads conflict-free, row reads broadcast).
// HBM-write-bound for one-exp programs, FP64-ALU / issue-bound once a term carries sin/cos.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace agpc {

constexpr int KB_ROWS = 32, KB_COLS = 128, KB_THREADS = 256;

// exp(x) for x <= 0 -- every exponent of a product term is -(sum of squares) / 2l^2.  libdevice's exp() spends more
// issue slots on materialising its polynomial coefficients (two 32-bit moves per double under the 80-register cap of the
// tiled kernels) and on its range / special-case tests than on arithmetic: 92 instructions per pair, ~25 of them FP64,
// made build_K issue-bound at 54-61 % of HBM.  Here the coefficients are constant-bank operands of the DFMAs:
// k = rint(x log2 e) via the 1.5 * 2^52 shifter, r = x - k ln2 (hi / lo split, exact product), Taylor polynomial on
// the reduced argument, 2^k added into the exponent field.  Below -708 (result < 2.3e-308, subnormal) the
// value is flushed to 0; NaN propagates; positive arguments are never produced by a valid program.
__constant__ double EXP_C[14] = {
    1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,
    1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
__constant__ double EXP_K[4] = {1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01,
                                -1.90821492927058770002e-10};

// select without a branch: ptxas otherwise turns the underflow test into a branch around the whole polynomial, which
// serialises the independent exps of a register tile into one dependent DFMA chain each (ILP 1; "wait" stalls dominate)
__device__ __forceinline__ double zero_if_below(double x, double lim, double v) {
  double out;
  asm("{ .reg .pred p; setp.lt.f64 p, %1, %2; selp.f64 %0, 0d0000000000000000, %3, p; }"
      : "=d"(out) : "d"(x), "d"(lim), "d"(v));
  return out;
}

// N independent exps in lock-step (every step is issued for all N values before the next one, so N dependent chains
// are in flight), table-driven: k = rint(32 x log2 e), x = (k >> 5) ln2 + (k & 31) ln2/32 + r with |r| <= ln2/64, so a
// degree-6 polynomial is exact to 4e-18 and exp(x) = 2^(k>>5) * T[k&31] * (1 + q(r)); 11 FP64 instructions per value
// instead of 17 (the FP64 pipe issues one warp instruction every 2 cycles).  tab = EXP_T staged in shared memory.
__constant__ double EXP_T[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
__constant__ double EXP_K32[4] = {0x1.71547652b82fep+5, 6755399441055744.0, -0x1.62e42fee00000p-6, -0x1.a39ef35793c76p-38};

template <int N>
__device__ __forceinline__ void exp_nonpos_n(double (&x)[N], const double* __restrict__ tab) {
  double r[N], p[N], tj[N];
  int e[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double t = fma(x[j], EXP_K32[0], EXP_K32[1]);
    const int ki = __double2loint(t);
    const double kd = t - EXP_K32[1];
    r[j] = fma(kd, EXP_K32[3], fma(kd, EXP_K32[2], x[j]));
    tj[j] = tab[ki & 31];
    e[j] = ki >> 5;
    p[j] = EXP_C[6];
  }
#pragma unroll
  for (int i = 5; i >= 1; --i)
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], EXP_C[i]);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double v = fma(tj[j], p[j] * r[j], tj[j]);
    x[j] = zero_if_below(x[j], -708.0, __hiloint2double(__double2hiint(v) + (e[j] << 20), __double2loint(v)));
  }
}

// single value, table-free (generic kernels): Taylor degree 13 on |r| <= ln2 / 2
__device__ __forceinline__ double exp_nonpos(double x) {
  const double t = fma(x, EXP_K[0], EXP_K[1]);
  const int k = __double2loint(t);
  const double kd = t - EXP_K[1];
  const double r = fma(kd, EXP_K[3], fma(kd, EXP_K[2], x));
  double p = EXP_C[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, EXP_C[i]);
  return zero_if_below(x, -708.0, __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p)));
}

struct LeafConst {
  int type, dim;
  double var, a, b, c, d;
};

struct ProgShared {
  int n_leaves, n_terms, n_theta;
  int term_len[AGPC_MAX_TERMS];
  int term_leaf[AGPC_MAX_TERMS][AGPC_MAX_TERM_LEN];
  int leaf_off[AGPC_MAX_LEAVES];
  LeafConst leaf[AGPC_MAX_LEAVES];
};

__device__ void load_program(ProgShared& P, const agpc_program& g, const double* __restrict__ theta) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    P.n_leaves = g.n_leaves;
    P.n_terms = g.n_terms;
    P.n_theta = g.n_theta;
  }
  if (tid < AGPC_MAX_TERMS) P.term_len[tid] = g.term_len[tid];
  if (tid < AGPC_MAX_TERMS * AGPC_MAX_TERM_LEN) (&P.term_leaf[0][0])[tid] = (&g.term_leaf[0][0])[tid];
  if (tid < AGPC_MAX_LEAVES && tid < g.n_leaves) {
    LeafConst lc;
    lc.type = g.leaf_type[tid];
    lc.dim = g.leaf_dim[tid];
    const int o = g.leaf_off[tid];
    P.leaf_off[tid] = o;
    lc.var = theta[o];
    lc.a = lc.b = lc.c = lc.d = 0.0;
    if (lc.type == AGPC_LEAF_SE) {
      const double l = theta[o + 1];
      lc.a = 1.0 / (l * l);
      lc.b = 1.0 / (l * l * l);
    } else if (lc.type == AGPC_LEAF_PER) {
      const double T = theta[o + 1], l = theta[o + 2];
      lc.a = 1.0 / T;
      lc.b = 1.0 / (l * l);
      lc.d = 1.0 / (l * l * l);
    } else if (lc.type == AGPC_LEAF_LIN) {
      lc.a = theta[o + 1];
    }
    P.leaf[tid] = lc;
  }
}

// ---- prologue of the tiled kernels: everything a tile needs is requested in ONE round trip to memory --------------
// The X panels go global -> shared with cp.async (no register staging, no wait before the next request), the program
// words and the whole theta vector are loaded by other threads at the same time; the leaf constants (1/l^2, ...) are
// then derived from shared memory.  (The first version chained program -> theta -> X loads: three dependent
// round trips per 64 x 64 tile, ~30 % of the warp-time samples of build_K_tiled_kernel.)
__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = valid ? 8 : 0;   // 0 source bytes: the 8 destination bytes are zero-filled, src is not read
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(src), "r"(bytes) : "memory");
}

// xr[d][r] = Xa[row0 + r][d], xc[d][r] = Xb[col0 + r][d] for r < TILE_W; rows beyond *_pad read as 0.  256 threads.
template <int TILE_W>
__device__ __forceinline__ void load_x_tiles_async(double* xr, double* xc, const double* __restrict__ Xa,
                                                   const double* __restrict__ Xb, int row0, int col0, int na_pad,
                                                   int nb_pad, int D) {
  static_assert(TILE_W == 64, "thread mapping assumes 64-wide tiles and 256 threads");
  const int r = threadIdx.x & 63, which = threadIdx.x >> 7, d0 = (threadIdx.x >> 6) & 1;
  const int g = (which ? col0 : row0) + r;
  const bool ok = g < (which ? nb_pad : na_pad);
  const double* src = (which ? Xb : Xa) + (ok ? (long long)g * D : 0);
  double* dst = (which ? xc : xr) + r;
  for (int d = d0; d < D; d += 2) cp_async8(dst + d * TILE_W, src + d, ok);
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// phase A (any time before the first barrier): raw program words and theta -> shared
__device__ __forceinline__ void load_program_raw(ProgShared& P, double* th_s, const agpc_program& g,
                                                 const double* __restrict__ theta, int theta_stride) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    P.n_leaves = g.n_leaves;
    P.n_terms = g.n_terms;
    P.n_theta = g.n_theta;
  }
  if (tid < AGPC_MAX_TERMS) P.term_len[tid] = g.term_len[tid];
  if (tid < AGPC_MAX_TERMS * AGPC_MAX_TERM_LEN) (&P.term_leaf[0][0])[tid] = (&g.term_leaf[0][0])[tid];
  if (tid >= 128 && tid < 128 + AGPC_MAX_LEAVES) {
    const int l = tid - 128;
    P.leaf[l].type = g.leaf_type[l];
    P.leaf[l].dim = g.leaf_dim[l];
    P.leaf_off[l] = g.leaf_off[l];
  }
  if (tid >= 160 && tid < 160 + AGPC_MAX_THETA && tid - 160 < theta_stride) th_s[tid - 160] = theta[tid - 160];
}

// phase B (after the barrier that publishes phase A): leaf constants from shared memory; needs one more barrier
__device__ __forceinline__ void derive_leaf_consts(ProgShared& P, const double* th_s) {
  const int tid = threadIdx.x;
  if (tid < P.n_leaves) {
    LeafConst& lc = P.leaf[tid];
    const int o = P.leaf_off[tid];
    double a = 0.0, b = 0.0, d = 0.0;
    if (lc.type == AGPC_LEAF_SE) {
      const double l = th_s[o + 1];
      a = 1.0 / (l * l);
      b = 1.0 / (l * l * l);
    } else if (lc.type == AGPC_LEAF_PER) {
      const double T = th_s[o + 1], l = th_s[o + 2];
      a = 1.0 / T;
      b = 1.0 / (l * l);
      d = 1.0 / (l * l * l);
    } else if (lc.type == AGPC_LEAF_LIN) {
      a = th_s[o + 1];
    }
    lc.var = th_s[o];
    lc.a = a;
    lc.b = b;
    lc.c = 0.0;
    lc.d = d;
  }
}

// K(x, x') and (GRAD) its partials w.r.t. theta accumulated into dk[0..n_theta) with weight w.
// xi(d) = xa[d * sa], xj(d) = xb[d * sb].
template <bool GRAD>
__device__ __forceinline__ double eval_pair(const ProgShared& P, const double* xa, int sa, const double* xb, int sb,
                                            double* dk, double w) {
  double K = 0.0;
  for (int t = 0; t < P.n_terms; ++t) {
    const int len = P.term_len[t];
    double arg = 0.0, mult = 1.0;
    bool has_exp = false;
    double sc[AGPC_MAX_TERM_LEN][2];
    double fac[AGPC_MAX_TERM_LEN];
    for (int p = 0; p < len; ++p) {
      const LeafConst& lc = P.leaf[P.term_leaf[t][p]];
      double f = lc.var;
      if (lc.type == AGPC_LEAF_SE) {
        const double r = xa[lc.dim * sa] - xb[lc.dim * sb];
        arg -= 0.5 * r * r * lc.a;
        has_exp = true;
      } else if (lc.type == AGPC_LEAF_PER) {
        const double r = xa[lc.dim * sa] - xb[lc.dim * sb];
        double s, c;
        sincos(M_PI * r * lc.a, &s, &c);  // GPy: np.sin(np.pi * r / T)
        arg -= 0.5 * s * s * lc.b;
        has_exp = true;
        if (GRAD) {
          sc[p][0] = s;
          sc[p][1] = c;
        }
      } else if (lc.type == AGPC_LEAF_LIN) {
        f *= (xa[lc.dim * sa] - lc.a) * (xb[lc.dim * sb] - lc.a);
      }
      mult *= f;
      if (GRAD) fac[p] = f;
    }
    const double e = has_exp ? exp_nonpos(arg) : 1.0;
    const double v = mult * e;
    K += v;
    if (GRAD) {
      const double wv = w * v;
      for (int p = 0; p < len; ++p) {
        const int l = P.term_leaf[t][p];
        const LeafConst& lc = P.leaf[l];
        const int o = P.leaf_off[l];
        if (lc.type == AGPC_LEAF_SE) {
          const double r = xa[lc.dim * sa] - xb[lc.dim * sb];
          dk[o] += wv / lc.var;
          dk[o + 1] += wv * r * r * lc.b;
        } else if (lc.type == AGPC_LEAF_PER) {
          const double r = xa[lc.dim * sa] - xb[lc.dim * sb];
          const double s = sc[p][0], c = sc[p][1];
          const double bb = M_PI * r * lc.a;
          dk[o] += wv / lc.var;
          dk[o + 1] += wv * s * c * (bb * lc.a) * lc.b;
          dk[o + 2] += wv * s * s * lc.d;
        } else if (lc.type == AGPC_LEAF_LIN) {
          double excl = e;
          for (int q = 0; q < len; ++q)
            if (q != p) excl *= fac[q];
          const double xi = xa[lc.dim * sa], xj = xb[lc.dim * sb];
          dk[o] += w * excl * (xi - lc.a) * (xj - lc.a);
          dk[o + 1] += w * excl * (-lc.var * (xi + xj - 2.0 * lc.a));
        } else {
          // CONST: dk/dv = (term without the constant) = v / var
          dk[o] += wv / lc.var;
        }
      }
    }
  }
  return K;
}

// ---- gather fold rows ------------------------------------------------------------------------------------
// Xf[b][i][d] = X[rows[off_b + i]][d], yf[b][i] = y[...]; padding rows: x = 0, y = +1.
__global__ void gather_rows_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                   const int* __restrict__ rows, const long long* __restrict__ row_off,
                                   const int* __restrict__ n_of, int D, int np, double* __restrict__ Xf,
                                   double* __restrict__ yf) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int n = n_of[b];
  double* xo = Xf + ((long long)b * np + i) * D;
  if (i < n) {
    const int r = rows[row_off[b] + i];
    for (int d = 0; d < D; ++d) xo[d] = X[(long long)r * D + d];
    yf[(long long)b * np + i] = y[r];
  } else {
    for (int d = 0; d < D; ++d) xo[d] = 0.0;
    yf[(long long)b * np + i] = 1.0;
  }
}

// ---- K build ---------------------------------------------------------------------------------------------
// out[b][i][j] = K(xa_i, xb_j) for i < na, j < nb, zero in the padding; optional kdiag (when xa == xb) and
// materialised dK slabs dK[b][theta][i][j].
template <bool GRAD>
__global__ void __launch_bounds__(KB_THREADS)
build_K_kernel(const BuildArgs a) {
  const int b = blockIdx.z;
  if (a.active != nullptr && a.active[b] == 0) return;
  extern __shared__ double smem[];
  __shared__ ProgShared P;
  const int D = a.D;
  double* xr = smem;                 // [D][KB_ROWS]
  double* xc = smem + D * KB_ROWS;   // [D][KB_COLS]
  const int row0 = blockIdx.y * KB_ROWS, col0 = blockIdx.x * KB_COLS;
  const int na = a.na_of ? a.na_of[b] : a.na_fixed, nb = a.nb_of ? a.nb_of[b] : a.nb_fixed;
  load_program(P, a.progs[b], a.theta + (long long)b * a.theta_stride);
  const double* Xa = a.Xa + (long long)b * a.xa_batch;
  const double* Xb = a.Xb + (long long)b * a.xb_batch;
  for (int idx = threadIdx.x; idx < KB_ROWS * D; idx += KB_THREADS) {
    const int r = idx / D, d = idx % D;
    xr[d * KB_ROWS + r] = (row0 + r < a.na_pad) ? Xa[(long long)(row0 + r) * D + d] : 0.0;
  }
  for (int idx = threadIdx.x; idx < KB_COLS * D; idx += KB_THREADS) {
    const int c = idx / D, d = idx % D;
    xc[d * KB_COLS + c] = (col0 + c < a.nb_pad) ? Xb[(long long)(col0 + c) * D + d] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 column lanes x 4 row groups
  double* out = a.out + (long long)b * a.out_batch;
  const int n_theta = P.n_theta;
  for (int rr = ty; rr < KB_ROWS; rr += 4) {
    const int i = row0 + rr;
    if (i >= a.na_pad) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = tx + 64 * h, j = col0 + cc;
      if (j >= a.ld) continue;
      const bool valid = (i < na) && (j < nb);
      double dk[GRAD ? AGPC_MAX_THETA : 1];
      if (GRAD)
        for (int q = 0; q < n_theta; ++q) dk[q] = 0.0;
      double k = 0.0;
      if (valid) k = eval_pair<GRAD>(P, xr + rr, KB_ROWS, xc + cc, KB_COLS, dk, 1.0);
      out[(long long)i * a.ld + j] = k;
      if (GRAD) {
        double* dKb = a.dK + (long long)b * n_theta * a.out_batch;
        for (int q = 0; q < n_theta; ++q) dKb[(long long)q * a.out_batch + (long long)i * a.ld + j] = valid ? dk[q] : 0.0;
      }
      if (a.kdiag != nullptr && i == j) a.kdiag[(long long)b * a.na_pad + i] = k;
    }
  }
}

// ---- K build, register micro-tiled and symmetric ------------------------------------------------------------------
// One CTA = one 64 x 64 tile, 256 threads, 4 x 4 pairs per thread (rows ty + 16r, columns tx + 16c).  The kernel
// program is walked ONCE per thread with the leaf constants in registers and applied to all 16 pairs, so the
// interpreter cost (the issue-bound part of build_K_kernel: 89 % SM busy at 24 % of HBM, profiles/r01_summary.md) is
// amortised 16x.  When Xa == Xb (training K) only the tiles on or below the diagonal are evaluated; each is also
// written transposed through shared memory, which halves the exp / sincos work per stored element -- that, not
// HBM, is what bounds this kernel for anything beyond a single SE term.
constexpr int KT = 64, KT_THREADS = 256, KT_LD = KT + 1;
// A CTA walks a STRIP of up to KT_STRIP consecutive tiles of one tile row: the kernel program, theta, the leaf
// constants and the row panel are loaded once per strip instead of once per tile, and the column panel of the next
// tile arrives by cp.async while the current tile is evaluated.  (Per-tile prologue + its two barriers were 28 % + 14 %
// of the warp-time samples of the one-tile-per-CTA form, profiles/r01_build_K_v2_ncu_source_regions.txt.)
constexpr int KT_STRIP = 8;
// PER leaves are evaluated through the angle-difference identity: sin(pi (x_i - x_j) / T) = s_i c_j - c_i s_j with
// (s, c) = sincospi(x / T) tabulated ONCE per tile row / tile column (64 sincos calls per leaf and tile side instead of
// 4096 sin calls per leaf and tile): the per-pair cost of a PER leaf drops from ~60 instructions to 3 FP64 ones, the
// same as an SE leaf.  Both forms round the phase pi x / T to ~1e-16 |x / T|, so they agree to that level (the oracle
// parity tests hold unchanged).  Up to KT_PER leaves get a table; further ones (never produced by the search grammar
// within AGPC_MAX_LEAVES in practice) fall back to the direct form.
constexpr int KT_PER = 8;

// xc[d][r] = Xb[col0 + r][d], r < 64, by all 256 threads
__device__ __forceinline__ void load_xc_async(double* xc, const double* __restrict__ Xb, int col0, int nb_pad, int D) {
  const int r = threadIdx.x & 63, d0 = threadIdx.x >> 6;
  const int g = col0 + r;
  const bool ok = g < nb_pad;
  const double* src = Xb + (ok ? (long long)g * D : 0);
  for (int d = d0; d < D; d += 4) cp_async8(xc + d * KT + r, src + d, ok);
}

template <int MINB>
__global__ void __launch_bounds__(KT_THREADS, MINB)
build_K_tiled_kernel(const BuildArgs a, const int symmetric, const int tiles_n, const int strip) {
  const int b = blockIdx.z;
  if (a.active != nullptr && a.active[b] == 0) return;
  extern __shared__ double smem[];
  __shared__ ProgShared P;
  __shared__ double exp_tab[32];
  __shared__ double th_s[AGPC_MAX_THETA];
  __shared__ int per_slot[AGPC_MAX_LEAVES];   // leaf -> trig table (or -1)
  __shared__ int per_leaf[KT_PER];            // table -> leaf
  __shared__ int n_per_s;
  if (threadIdx.x < 32) exp_tab[threadIdx.x] = EXP_T[threadIdx.x];
  const int D = a.D;
  double* xr = smem;                // [D][KT]
  double* xc0 = smem + D * KT;      // [2][D][KT]: column panel, double-buffered across the strip
  double* st = xc0 + 2 * D * KT;    // [KT][KT_LD] transposition buffer (symmetric mode)
  double* trg_r = st + KT * KT_LD;  // [KT_PER][2][KT]: (1/2l^2)^(1/4) * (sin, cos)(pi x_i / T) of the tile rows
  double* trg_c = trg_r + KT_PER * 2 * KT;  // [KT_PER][2][KT]: the same of the tile columns
  // strip index -> (tile row, first tile column): row ti holds ti + 1 (symmetric) or tiles_n tiles
  int ti = 0, tj0;
  {
    int t = blockIdx.x;
    if (symmetric) {
      // tile row ti holds ceil((ti + 1) / strip) strips: q + 1 for each of the `strip` rows of group q = ti / strip.  Walk
      // the groups (<= tiles_m / strip steps, no division), then one division inside the group.  (A row-by-row walk with a
      // division per row was 9 % of this kernel's instructions and 12 % of its stall samples at n = 6400.)
      int q = 0, base = 0;
      while (base + strip * (q + 1) <= t) {
        base += strip * (q + 1);
        ++q;
      }
      const int rem = (t - base) / (q + 1);
      ti = q * strip + rem;
      t -= base + rem * (q + 1);
    } else {
      const int strips = (tiles_n + strip - 1) / strip;
      ti = t / strips;
      t %= strips;
    }
    tj0 = t * strip;
  }
  const int row_tiles = symmetric ? ti + 1 : tiles_n;
  const int n_tiles = min(strip, row_tiles - tj0);
  const int row0 = ti * KT;
  const double* Xb = a.Xb + (long long)b * a.xb_batch;
  load_x_tiles_async<KT>(xr, xc0, a.Xa + (long long)b * a.xa_batch, Xb, row0, tj0 * KT, a.na_pad, a.nb_pad, D);
  load_program_raw(P, th_s, a.progs[b], a.theta + (long long)b * a.theta_stride, a.theta_stride);
  const int na = a.na_of ? a.na_of[b] : a.na_fixed, nb = a.nb_of ? a.nb_of[b] : a.nb_fixed;
  cp_async_wait_all();
  __syncthreads();
  derive_leaf_consts(P, th_s);
  if (threadIdx.x == KT_THREADS - 1) {
    int np = 0;
    for (int l = 0; l < P.n_leaves; ++l) {
      const bool tab = P.leaf[l].type == AGPC_LEAF_PER && np < KT_PER;
      per_slot[l] = tab ? np : -1;
      if (tab) per_leaf[np++] = l;
    }
    n_per_s = np;
  }
  __syncthreads();
  const int n_per = n_per_s;
  // row tables, once per strip; both sides carry (1 / 2 l^2)^(1/4), so a pair costs  t = sr cc - cr sc; arg -= t t
  for (int idx = threadIdx.x; idx < n_per * KT; idx += KT_THREADS) {
    const int sl = idx >> 6, r = idx & (KT - 1);
    const LeafConst& lc = P.leaf[per_leaf[sl]];
    double sn, cs;
    sincospi(xr[lc.dim * KT + r] * lc.a, &sn, &cs);
    const double w = sqrt(sqrt(0.5 * lc.b));
    trg_r[(sl * 2 + 0) * KT + r] = w * sn;
    trg_r[(sl * 2 + 1) * KT + r] = w * cs;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int kt = 0; kt < n_tiles; ++kt) {
  const int tj = tj0 + kt;
  const int col0 = tj * KT;
  const double* xc = xc0 + (kt & 1) * D * KT;
  if (kt + 1 < n_tiles) load_xc_async(xc0 + ((kt + 1) & 1) * D * KT, Xb, (tj + 1) * KT, a.nb_pad, D);
  if (n_per > 0) {  // column tables of this tile (the previous tile's readers passed the barrier that ended it)
    for (int idx = threadIdx.x; idx < n_per * KT; idx += KT_THREADS) {
      const int sl = idx >> 6, r = idx & (KT - 1);
      const LeafConst& lc = P.leaf[per_leaf[sl]];
      double sn, cs;
      sincospi(xc[lc.dim * KT + r] * lc.a, &sn, &cs);
      const double w = sqrt(sqrt(0.5 * lc.b));
      trg_c[(sl * 2 + 0) * KT + r] = w * sn;
      trg_c[(sl * 2 + 1) * KT + r] = w * cs;
    }
    __syncthreads();
  }

  // two passes of 2 x 4 pairs (rows ty + 16r, r = 2h, 2h + 1): keeps the live set under 80 registers -> 3 CTAs per SM,
  // which is what hides the load -> compute -> store latency chain of a tile (the kernel is latency-, not ALU-bound)
  double* out = a.out + (long long)b * a.out_batch;
  const bool mirror = symmetric && ti != tj;
  // fast path: an off-diagonal tile (no kdiag) that lies completely inside the valid n x n region, mirror included
  const bool interior = ti != tj && row0 + KT <= na && col0 + KT <= nb && row0 + KT <= a.na_pad && col0 + KT <= a.ld &&
                        (!mirror || (col0 + KT <= a.na_pad && row0 + KT <= a.ld));
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    double K[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) K[r][c] = 0.0;
    const int n_terms = P.n_terms;
    for (int t = 0; t < n_terms; ++t) {
      const int len = P.term_len[t];
      double arg[2][4];
      double mr[2] = {1.0, 1.0}, mc[4] = {1.0, 1.0, 1.0, 1.0};  // LIN factors are separable: rows x columns
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) arg[r][c] = 0.0;
      double scale = 1.0;
      bool has_exp = false, has_mult = false;
      for (int p = 0; p < len; ++p) {
        const int leaf_id = P.term_leaf[t][p];
        const LeafConst& lc = P.leaf[leaf_id];
        scale *= lc.var;
        if (lc.type == AGPC_LEAF_CONST) continue;
        if (lc.type == AGPC_LEAF_PER && per_slot[leaf_id] >= 0) {
          has_exp = true;
          const double* tr = trg_r + per_slot[leaf_id] * 2 * KT;
          const double* tc = trg_c + per_slot[leaf_id] * 2 * KT;
          double sr[2], cr[2], sc[4], cc[4];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            sr[r] = tr[ty + 16 * (2 * h + r)];
            cr[r] = tr[KT + ty + 16 * (2 * h + r)];
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sc[c] = tc[tx + 16 * c];
            cc[c] = tc[KT + tx + 16 * c];
          }
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // sqrt(1/2l^2) sin(pi (x_i - x_j) / T); two rounded products and a difference (no fma), so that swapping
              // i and j negates it EXACTLY: the diagonal tiles evaluate both (i, j) and (j, i), and K must be symmetric
              const double sn = __dsub_rn(__dmul_rn(sr[r], cc[c]), __dmul_rn(cr[r], sc[c]));
              arg[r][c] = fma(-sn, sn, arg[r][c]);
            }
          continue;
        }
        double xi[2], xj[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) xi[r] = xr[lc.dim * KT + ty + 16 * (2 * h + r)];
#pragma unroll
        for (int c = 0; c < 4; ++c) xj[c] = xc[lc.dim * KT + tx + 16 * c];
        if (lc.type == AGPC_LEAF_SE) {
          has_exp = true;
          const double ha = 0.5 * lc.a;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double d = xi[r] - xj[c];
              arg[r][c] -= ha * d * d;
            }
        } else if (lc.type == AGPC_LEAF_PER) {
          has_exp = true;
          const double w = M_PI * lc.a, hb = 0.5 * lc.b;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double sn = sin(w * (xi[r] - xj[c]));  // GPy: np.sin(np.pi * r / T)
              arg[r][c] -= hb * sn * sn;
            }
        } else {  // LIN
          has_mult = true;
#pragma unroll
          for (int r = 0; r < 2; ++r) mr[r] *= xi[r] - lc.a;
#pragma unroll
          for (int c = 0; c < 4; ++c) mc[c] *= xj[c] - lc.a;
        }
      }
      if (has_exp) exp_nonpos_n<8>(reinterpret_cast<double(&)[8]>(arg), exp_tab);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double v = scale;
          if (has_mult) v *= mr[r] * mc[c];
          if (has_exp) v *= arg[r][c];
          K[r][c] += v;
        }
    }
    if (interior) {
      // off-diagonal tile fully inside the valid region: no masks, one row pointer + immediate column offsets
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int lr = ty + 16 * (2 * h + r);
        double* o = out + (long long)(row0 + lr) * a.ld + col0 + tx;
        double* sp = st + lr * KT_LD + tx;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          o[16 * c] = K[r][c];
          if (mirror) sp[16 * c] = K[r][c];
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int lr = ty + 16 * (2 * h + r);
        const int i = row0 + lr;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = col0 + tx + 16 * c;
          if (!(i < na && j < nb)) K[r][c] = 0.0;
          if (i < a.na_pad && j < a.ld) out[(long long)i * a.ld + j] = K[r][c];
          if (a.kdiag != nullptr && i == j && i < a.na_pad && ti == tj) a.kdiag[(long long)b * a.na_pad + i] = K[r][c];
          if (mirror) st[lr * KT_LD + tx + 16 * c] = K[r][c];
        }
      }
    }
  }
  if (mirror) {
    // (a direct store of the mirror image from registers -- 8 rows x 32 B per warp instruction -- was measured 20 %
    // slower than this transposition through shared memory: 4x the L1 tag work per store instruction)
    __syncthreads();
    // mirror tile: rows col0.., columns row0..; element (jj, ii) = st[ii][jj]
    if (interior) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double* o = out + (long long)(col0 + ty + 16 * r) * a.ld + row0 + tx;
        const double* sp = st + tx * KT_LD + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < 4; ++c) o[16 * c] = sp[16 * c * KT_LD];
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int jrow = col0 + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int icol = row0 + tx + 16 * c;
          if (jrow < a.na_pad && icol < a.ld) out[(long long)jrow * a.ld + icol] = st[(tx + 16 * c) * KT_LD + ty + 16 * r];
        }
      }
    }
  }
  // the next tile's column panel has landed; the transposition buffer and the panel just used are free again
  cp_async_wait_all();
  __syncthreads();
  }
}

// ---- K AND materialised dK/dtheta, streaming ------------------------------------------------------------------------
// Same 64 x 64 tiling and symmetry as build_K_tiled_kernel.  Each product term is evaluated once per pair (one exp);
// the slabs of its leaves' parameters are that value times a cheap per-leaf factor and are written straight out --
// possible because (checked on the host) every leaf belongs to exactly one term, so no slab is accumulated across
// terms.  (1 + P) x 8 bytes per pair for one exp: this variant is HBM-write-bound.
enum : int { SLAB_K = 0, SLAB_VAR, SLAB_SE_L, SLAB_PER_T, SLAB_PER_L, SLAB_LIN_C };

struct SlabCtx {
  const double (*ex)[4];   // exp part of the term, per pair
  const double* mr;        // LIN factors of the term: rows / columns (for SLAB_LIN_C: without the leaf itself)
  const double* mc;
  const double* xi;        // coordinates of the leaf's dimension
  const double* xj;
  double s;                // scalar prefactor (product of variances as the slab needs it)
  double p0, p1, p2;       // leaf constants
};

template <int KIND>
__device__ __forceinline__ double slab_value(const SlabCtx& k, const double (&Kacc)[4][4], int r, int c) {
  if (KIND == SLAB_K) return Kacc[r][c];
  const double base = k.s * k.mr[r] * k.mc[c] * k.ex[r][c];
  if (KIND == SLAB_VAR) return base;
  const double d = k.xi[r] - k.xj[c];
  if (KIND == SLAB_SE_L) return base * d * d * k.p0;                        // k r^2 / l^3
  if (KIND == SLAB_LIN_C) return -base * (k.xi[r] + k.xj[c] - 2.0 * k.p0);  // -(rest) var (x + x' - 2c)
  const double bb = M_PI * d * k.p0;                                        // PER: b = pi r / T
  double sn, cs;
  sincos(bb, &sn, &cs);
  if (KIND == SLAB_PER_T) return base * sn * cs * (bb * k.p0) * k.p1;       // k sin b cos b (b / T) / l^2
  return base * sn * sn * k.p2;                                             // k sin^2 b / l^3
}

// writes one 64 x 64 tile of one slab (and its mirror image through shared memory when the tile is off-diagonal)
template <int KIND>
__device__ __forceinline__ void emit_slab(const BuildArgs& a, double* __restrict__ dst, const SlabCtx& k,
                                          const double (&Kacc)[4][4], double* st, int row0, int col0, int na, int nb,
                                          int tx, int ty, bool mirror) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int lr = ty + 16 * r, i = row0 + lr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = col0 + tx + 16 * c;
      const double x = (i < na && j < nb) ? slab_value<KIND>(k, Kacc, r, c) : 0.0;
      if (i < a.na_pad && j < a.ld) dst[(long long)i * a.ld + j] = x;
      if (mirror) st[lr * KT_LD + tx + 16 * c] = x;
    }
  }
  if (mirror) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int jrow = col0 + ty + 16 * r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int icol = row0 + tx + 16 * c;
        if (jrow < a.na_pad && icol < a.ld) dst[(long long)jrow * a.ld + icol] = st[(tx + 16 * c) * KT_LD + ty + 16 * r];
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(KT_THREADS, 2)
build_KdK_tiled_kernel(const BuildArgs a, const int symmetric, const int tiles_n) {
  const int b = blockIdx.z;
  if (a.active != nullptr && a.active[b] == 0) return;
  extern __shared__ double smem[];
  __shared__ ProgShared P;
  __shared__ double exp_tab[32];
  if (threadIdx.x < 32) exp_tab[threadIdx.x] = EXP_T[threadIdx.x];
  const int D = a.D;
  double* xr = smem;
  double* xc = smem + D * KT;
  double* st = xc + D * KT;
  int ti, tj;
  if (symmetric) {
    const int t = blockIdx.x;
    ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);  // estimate; the loops below make it exact
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = blockIdx.x / tiles_n;
    tj = blockIdx.x % tiles_n;
  }
  const int row0 = ti * KT, col0 = tj * KT;
  const int na = a.na_of ? a.na_of[b] : a.na_fixed, nb = a.nb_of ? a.nb_of[b] : a.nb_fixed;
  load_program(P, a.progs[b], a.theta + (long long)b * a.theta_stride);
  const double* Xa = a.Xa + (long long)b * a.xa_batch;
  const double* Xb = a.Xb + (long long)b * a.xb_batch;
  for (int idx = threadIdx.x; idx < KT * D; idx += KT_THREADS) {
    const int r = idx / D, d = idx % D;
    xr[d * KT + r] = (row0 + r < a.na_pad) ? Xa[(long long)(row0 + r) * D + d] : 0.0;
    xc[d * KT + r] = (col0 + r < a.nb_pad) ? Xb[(long long)(col0 + r) * D + d] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool mirror = symmetric && ti != tj;
  double* outK = a.out + (long long)b * a.out_batch;
  double* outdK = a.dK + (long long)b * P.n_theta * a.out_batch;

  double K[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) K[r][c] = 0.0;

  const int n_terms = P.n_terms;
  for (int t = 0; t < n_terms; ++t) {
    const int len = P.term_len[t];
    // ---- term value, factored: scale (variances) * mr[r] mc[c] (LIN leaves) * ex[r][c] (exp of the summed exponents) ----
    double ex[4][4], mr[4] = {1.0, 1.0, 1.0, 1.0}, mc[4] = {1.0, 1.0, 1.0, 1.0};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) ex[r][c] = 0.0;
    bool has_exp = false;
    double scale = 1.0;
    for (int p = 0; p < len; ++p) {
      const LeafConst& lc = P.leaf[P.term_leaf[t][p]];
      scale *= lc.var;
      if (lc.type == AGPC_LEAF_CONST) continue;
      double xi[4], xj[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) xi[r] = xr[lc.dim * KT + ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) xj[c] = xc[lc.dim * KT + tx + 16 * c];
      if (lc.type == AGPC_LEAF_SE) {
        has_exp = true;
        const double ha = 0.5 * lc.a;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double d = xi[r] - xj[c];
            ex[r][c] -= ha * d * d;
          }
      } else if (lc.type == AGPC_LEAF_PER) {
        has_exp = true;
        const double w = M_PI * lc.a, hb = 0.5 * lc.b;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double sn = sin(w * (xi[r] - xj[c]));
            ex[r][c] -= hb * sn * sn;
          }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) mr[r] *= xi[r] - lc.a;
#pragma unroll
        for (int c = 0; c < 4; ++c) mc[c] *= xj[c] - lc.a;
      }
    }
    if (has_exp) exp_nonpos_n<16>(reinterpret_cast<double(&)[16]>(ex), exp_tab);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (!has_exp) ex[r][c] = 1.0;
        K[r][c] += scale * mr[r] * mc[c] * ex[r][c];
      }

    // ---- the slabs of this term's leaves ----
    for (int p = 0; p < len; ++p) {
      const int l = P.term_leaf[t][p];
      const LeafConst& lc = P.leaf[l];
      double* slab = outdK + (long long)P.leaf_off[l] * a.out_batch;
      double sx = 1.0;  // product of the OTHER leaves' variances: d/dvar without dividing by var
      for (int q = 0; q < len; ++q)
        if (q != p) sx *= P.leaf[P.term_leaf[t][q]].var;
      double xi[4] = {0.0, 0.0, 0.0, 0.0}, xj[4] = {0.0, 0.0, 0.0, 0.0};
      if (lc.type != AGPC_LEAF_CONST) {
#pragma unroll
        for (int r = 0; r < 4; ++r) xi[r] = xr[lc.dim * KT + ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) xj[c] = xc[lc.dim * KT + tx + 16 * c];
      }
      SlabCtx k{ex, mr, mc, xi, xj, sx, 0.0, 0.0, 0.0};
      emit_slab<SLAB_VAR>(a, slab, k, K, st, row0, col0, na, nb, tx, ty, mirror);
      k.s = scale;
      if (lc.type == AGPC_LEAF_SE) {
        k.p0 = lc.b;
        emit_slab<SLAB_SE_L>(a, slab + a.out_batch, k, K, st, row0, col0, na, nb, tx, ty, mirror);
      } else if (lc.type == AGPC_LEAF_PER) {
        k.p0 = lc.a; k.p1 = lc.b; k.p2 = lc.d;
        emit_slab<SLAB_PER_T>(a, slab + a.out_batch, k, K, st, row0, col0, na, nb, tx, ty, mirror);
        emit_slab<SLAB_PER_L>(a, slab + 2 * a.out_batch, k, K, st, row0, col0, na, nb, tx, ty, mirror);
      } else if (lc.type == AGPC_LEAF_LIN) {
        // the term without this leaf's own factor (x - c)(x' - c): rebuild the LIN products from the other leaves
        double mr2[4] = {1.0, 1.0, 1.0, 1.0}, mc2[4] = {1.0, 1.0, 1.0, 1.0};
        for (int q = 0; q < len; ++q) {
          const LeafConst& lq = P.leaf[P.term_leaf[t][q]];
          if (q == p || lq.type != AGPC_LEAF_LIN) continue;
#pragma unroll
          for (int r = 0; r < 4; ++r) mr2[r] *= xr[lq.dim * KT + ty + 16 * r] - lq.a;
#pragma unroll
          for (int c = 0; c < 4; ++c) mc2[c] *= xc[lq.dim * KT + tx + 16 * c] - lq.a;
        }
        SlabCtx k2{ex, mr2, mc2, xi, xj, scale, lc.a, 0.0, 0.0};
        emit_slab<SLAB_LIN_C>(a, slab + a.out_batch, k2, K, st, row0, col0, na, nb, tx, ty, mirror);
      }
    }
  }
  SlabCtx kk{nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, 0.0, 0.0, 0.0};
  emit_slab<SLAB_K>(a, outK, kk, K, st, row0, col0, na, nb, tx, ty, mirror);
}

cudaError_t build_K_launch(const Launch& L, const BuildArgs& a, int batch, bool grad) {
  ProfScope prof_scope(L, CLS_KBUILD);
  if (grad && a.dk_tiled) {
    bool& configured = func_attr_done(__LINE__);
    const size_t smem = sizeof(double) * (2 * (size_t)a.D * KT + KT * KT_LD);
    if (!configured || smem > 48 * 1024) {
      AGPC_CUDA_TRY(cudaFuncSetAttribute(build_KdK_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(smem > 100 * 1024 ? smem : 100 * 1024)));
      configured = true;
    }
    const bool symmetric = a.Xa == a.Xb && a.xa_batch == a.xb_batch && a.na_of == a.nb_of && a.na_fixed == a.nb_fixed &&
                           a.na_pad == a.nb_pad && a.ld == a.na_pad;
    const int tiles_m = (a.na_pad + KT - 1) / KT;
    const int tiles_n = ((symmetric ? a.na_pad : (int)a.ld) + KT - 1) / KT;
    dim3 grid(symmetric ? (unsigned)(tiles_m * (tiles_m + 1) / 2) : (unsigned)(tiles_m * tiles_n), 1, batch);
    build_KdK_tiled_kernel<<<grid, KT_THREADS, smem, L.stream>>>(a, symmetric ? 1 : 0, tiles_n);
    L.bump();
    return cudaGetLastError();
  }
  if (grad) {
    dim3 grid((unsigned)((a.ld + KB_COLS - 1) / KB_COLS), (unsigned)((a.na_pad + KB_ROWS - 1) / KB_ROWS), batch);
    const size_t smem = sizeof(double) * a.D * (KB_ROWS + KB_COLS);
    build_K_kernel<true><<<grid, KB_THREADS, smem, L.stream>>>(a);
    L.bump();
    return cudaGetLastError();
  }
  bool& configured = func_attr_done(__LINE__);
  const size_t smem = sizeof(double) * (3 * (size_t)a.D * KT + KT * KT_LD + 2 * KT_PER * 2 * KT);
  static const int occ = getenv("AGPC_KB_OCC") ? atoi(getenv("AGPC_KB_OCC")) : 2;   // resident CTAs per SM the kernel is built for
  if (!configured || smem > 48 * 1024) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(build_K_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(smem > 100 * 1024 ? smem : 100 * 1024)));
    AGPC_CUDA_TRY(cudaFuncSetAttribute(build_K_tiled_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(smem > 100 * 1024 ? smem : 100 * 1024)));
    configured = true;
  }
  // symmetric when both point sets are the same array and the output is square enough to hold the mirror tiles
  const bool symmetric = a.Xa == a.Xb && a.xa_batch == a.xb_batch && a.na_of == a.nb_of && a.na_fixed == a.nb_fixed &&
                         a.na_pad == a.nb_pad && a.ld >= a.na_pad;
  const int tiles_m = (a.na_pad + KT - 1) / KT;
  const int cols = symmetric ? a.na_pad : (int)a.ld;
  const int tiles_n = (cols + KT - 1) / KT;
  // strip length: as long as possible (KT_STRIP) while the launch still has ~4 CTAs per resident slot
  const long long tiles_total = (symmetric ? (long long)tiles_m * (tiles_m + 1) / 2 : (long long)tiles_m * tiles_n) * batch;
  int strip = (int)(tiles_total / (148LL * occ * 4));
  strip = strip < 1 ? 1 : (strip > KT_STRIP ? KT_STRIP : strip);
  unsigned strips = 0;
  if (symmetric)
    for (int ti = 0; ti < tiles_m; ++ti) strips += (unsigned)((ti + 1 + strip - 1) / strip);
  else
    strips = (unsigned)(tiles_m * ((tiles_n + strip - 1) / strip));
  dim3 grid(strips, 1, batch);
  if (occ == 2)
    build_K_tiled_kernel<2><<<grid, KT_THREADS, smem, L.stream>>>(a, symmetric ? 1 : 0, tiles_n, strip);
  else
    build_K_tiled_kernel<3><<<grid, KT_THREADS, smem, L.stream>>>(a, symmetric ? 1 : 0, tiles_n, strip);
  L.bump();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (symmetric && a.ld > a.na_pad) {
    // padding columns [na_pad, ld) are not covered by the triangular tile list: clear them
    return cudaMemset2DAsync(a.out + a.na_pad, sizeof(double) * a.ld, 0, sizeof(double) * (a.ld - a.na_pad),
                             (size_t)a.na_pad * batch, L.stream);
  }
  return cudaSuccess;
}

cudaError_t gather_rows_launch(const Launch& L, const double* X, const double* y, const int* rows,
                               const long long* row_off, const int* n_of, int D, int np, int batch, double* Xf,
                               double* yf) {
  ProfScope prof_scope(L, CLS_MISC);
  dim3 grid((np + 255) / 256, batch);
  gather_rows_kernel<<<grid, 256, 0, L.stream>>>(X, y, rows, row_off, n_of, D, np, Xf, yf);
  L.bump();
  return cudaGetLastError();
}

// ---- B = I + S^1/2 K S^1/2 (lower tiles only) ---------------------------------------------------------------
__global__ void form_B_kernel(const double* __restrict__ K, const double* __restrict__ s, double* __restrict__ A,
                              int np, long long stride, const int* __restrict__ active, const double* __restrict__ jitter) {
  const int b = blockIdx.z;
  if (active != nullptr && active[b] == 0) return;
  const int ti = blockIdx.y, tj = blockIdx.x;  // 32 x 128 tiles
  const int row0 = ti * 32, col0 = tj * 128;
  if (col0 > row0 + 31) return;  // strictly above the diagonal
  const double* Kb = K + (long long)b * stride;
  double* Ab = A + (long long)b * stride;
  const double* sb = s + (long long)b * np;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int rr = ty; rr < 32; rr += 4) {
    const int i = row0 + rr;
    const double si = sb[i];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = col0 + tx + 64 * h;
      const long long o = (long long)i * np + j;
      double v = si * Kb[o] * sb[j];
      if (i == j) v += 1.0 + (jitter != nullptr ? jitter[b] : 0.0);  // jitchol's retry: B + jitter I
      Ab[o] = v;
    }
  }
}

cudaError_t form_B_launch(const Launch& L, const double* K, const double* s, double* A, int np, long long stride,
                          int batch, const int* active, const double* jitter) {
  ProfScope prof_scope(L, CLS_FORMB);
  dim3 grid(np / 128, np / 32, batch);
  form_B_kernel<<<grid, 256, 0, L.stream>>>(K, s, A, np, stride, active, jitter);
  L.bump();
  return cudaGetLastError();
}

// ---- jitchol's jitter ladder (GPy util/linalg.py, reached from EP.inference for gpckernel.py:245-246) --------------
// One CTA per item whose factorisation failed (info == NOT_PD).  First failure of this factorisation (tries == 0):
// jitter = mean(diag B) 1e-6 with B_ii = 1 + s_i^2 K_ii over the item's true rows; later failures: jitter *= 10; after
// 5 jittered attempts, or with a non-positive / non-finite diagonal, the item stays NOT_PD.  retry[b] = 1 marks the
// items to re-form and re-factorise (their info is cleared), used[b] records that the result carries jitter.
__global__ void __launch_bounds__(256)
jitter_step_kernel(const double* __restrict__ kdiag, const double* __restrict__ s, const int* __restrict__ n_of, int np,
                   int* __restrict__ info, double* __restrict__ jitter, int* __restrict__ tries, int* __restrict__ retry,
                   int* __restrict__ used, const int* __restrict__ active) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) retry[b] = 0;
  if ((active != nullptr && active[b] == 0) || info[b] != AGPC_ST_NOT_PD) return;
  __shared__ double red[256];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  const int n = n_of[b];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double si = s[(long long)b * np + i];
    const double d = 1.0 + si * si * kdiag[(long long)b * np + i];
    if (!(d > 0.0) || !isfinite(d)) bad = 1;
    acc += d;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int t = tries[b];
    if (bad || t >= 5) return;  // gives up: info stays NOT_PD
    jitter[b] = t == 0 ? red[0] / n * 1e-6 : jitter[b] * 10.0;
    tries[b] = t + 1;
    retry[b] = 1;
    used[b] = 1;
    info[b] = 0;
  }
}
// the same ladder step for a plain matrix (agpc_jitchol): diagonal read from the untouched copy A0; retried items get
// A = A0 + jitter I written back (lower triangle + diagonal; the factorisation reads nothing else)
__global__ void __launch_bounds__(256)
jitter_step_mat_kernel(const double* __restrict__ A0, int n, long long stride, int* __restrict__ info,
                       double* __restrict__ jitter, int* __restrict__ tries, int* __restrict__ retry) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) retry[b] = 0;
  if (info[b] != AGPC_ST_NOT_PD) return;
  __shared__ double red[256];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double d = A0[(long long)b * stride + (long long)i * n + i];
    if (!(d > 0.0) || !isfinite(d)) bad = 1;
    acc += d;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int t = tries[b];
    if (bad || t >= 5) return;
    jitter[b] = t == 0 ? red[0] / n * 1e-6 : jitter[b] * 10.0;
    tries[b] = t + 1;
    retry[b] = 1;
    info[b] = 0;
  }
}
__global__ void restore_jitter_kernel(const double* __restrict__ A0, double* __restrict__ A, int n, long long stride,
                                      const double* __restrict__ jitter, const int* __restrict__ retry) {
  const int b = blockIdx.y;
  if (retry[b] == 0) return;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  A[(long long)b * stride + idx] = A0[(long long)b * stride + idx] + (i == j ? jitter[b] : 0.0);
}
cudaError_t jitter_step_mat_launch(const Launch& L, const double* A0, double* A, int n, long long stride, int batch, int* info,
                                   double* jitter, int* tries, int* retry) {
  ProfScope prof_scope(L, CLS_MISC);
  jitter_step_mat_kernel<<<batch, 256, 0, L.stream>>>(A0, n, stride, info, jitter, tries, retry);
  dim3 grid((unsigned)(((long long)n * n + 255) / 256), batch);
  restore_jitter_kernel<<<grid, 256, 0, L.stream>>>(A0, A, n, stride, jitter, retry);
  L.bump(2);
  return cudaGetLastError();
}

cudaError_t jitter_step_launch(const Launch& L, const double* kdiag, const double* s, const int* n_of, int np, int batch,
                               int* info, double* jitter, int* tries, int* retry, int* used, const int* active) {
  ProfScope prof_scope(L, CLS_MISC);
  jitter_step_kernel<<<batch, 256, 0, L.stream>>>(kdiag, s, n_of, np, info, jitter, tries, retry, used, active);
  L.bump();
  return cudaGetLastError();
}

// ---- fused gradient contraction ---------------------------------------------------------------------------
// partial[b][tile][q] = sum over the tile's pairs of  w_ik * dK_ik/dtheta_q, where
//   mode 0 (EP):   w_ik = c_ik * 1/2 (alpha_i alpha_k - s_i Binv_ik s_k), lower triangle, c = 2 off-diagonal, 1 on it
//   mode 1 (raw):  w_ik = G_ik (full matrix)
constexpr int GT = 64;  // gradient tile edge

__global__ void __launch_bounds__(256)
grad_contract_kernel(const GradArgs a) {
  const int b = blockIdx.y;
  if (a.active != nullptr && a.active[b] == 0) return;
  __shared__ ProgShared P;
  extern __shared__ double smem[];
  const int D = a.D;
  double* xr = smem;             // [D][GT]
  double* xc = smem + D * GT;    // [D][GT]
  double* red = xc + D * GT;     // [8 warps][AGPC_MAX_THETA]
  const int nt = a.np / GT;
  int ti, tj;
  if (a.mode == 0) {
    const int t = blockIdx.x;
    ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);  // estimate; the loops below make it exact
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = blockIdx.x / nt;
    tj = blockIdx.x % nt;
  }
  const int n = a.n_of ? a.n_of[b] : a.n_fixed;
  load_program(P, a.progs[b], a.theta + (long long)b * a.theta_stride);
  const double* Xf = a.Xf + (long long)b * a.np * D;
  const int row0 = ti * GT, col0 = tj * GT;
  for (int idx = threadIdx.x; idx < GT * D; idx += 256) {
    const int r = idx / D, d = idx % D;
    xr[d * GT + r] = Xf[(long long)(row0 + r) * D + d];
    xc[d * GT + r] = Xf[(long long)(col0 + r) * D + d];
  }
  __syncthreads();
  const int n_theta = P.n_theta;
  double acc[AGPC_MAX_THETA];
  for (int q = 0; q < n_theta; ++q) acc[q] = 0.0;
  const double* Gb = a.G + (long long)b * a.g_batch;
  const double* al = a.alpha ? a.alpha + (long long)b * a.np : nullptr;
  const double* sv = a.s ? a.s + (long long)b * a.np : nullptr;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = col0 + tx;
  for (int rr = ty; rr < GT; rr += 4) {
    const int i = row0 + rr;
    if (i >= n || j >= n) continue;
    double w;
    if (a.mode == 0) {
      if (j > i) continue;
      const double g = Gb[(long long)i * a.ld + j];
      w = 0.5 * (al[i] * al[j] - sv[i] * g * sv[j]);
      if (a.ru != nullptr) {
        const double* ru = a.ru + (long long)b * a.np;
        const double* rv = a.rv + (long long)b * a.np;
        w += 0.5 * (ru[i] * rv[j] + ru[j] * rv[i]);
      }
      if (j < i) w *= 2.0;
    } else {
      w = Gb[(long long)i * a.ld + j];
    }
    eval_pair<true>(P, xr + rr, GT, xc + tx, GT, acc, w);
  }
  // deterministic block reduction: shuffle within warps, fixed-order sum across the 8 warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < n_theta; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * AGPC_MAX_THETA + q] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_theta) {
    double v = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8 * AGPC_MAX_THETA + threadIdx.x];
    a.partial[((long long)b * a.tiles_per_item + blockIdx.x) * AGPC_MAX_THETA + threadIdx.x] = v;
  }
}

// ---- the same contraction, register micro-tiled (round 2) --------------------------------------------------------------
// grad_contract_kernel above walks the kernel program once per PAIR and keeps its dk[] in a theta-indexed array, i.e. in
// local memory: 337 instructions per pair, 0.09 of the HBM rate of the B^-1 tiles it reads.  Here a thread owns 2 x 4
// pairs per pass (two passes per 64 x 64 tile, as build_K_tiled_kernel), the program is walked once per pass, and what
// has to be summed over pairs is, per product term t, S0_t = sum w v_t (every variance: dk/dvar_p = v_t / var_p) and per
// leaf one or two more sums (SE: sum w v r^2; PER: sum w v sin cos r and sum w v sin^2; LIN: sum w v/(factor) d factor/dc).
// Each is reduced over the warp as soon as its 8 pairs are done and added by lane 0 into the warp's own row of a
// shared [8 warps][theta] table: no atomics, fixed order, so the result is deterministic.  PER leaves use the same
// per-tile sincospi tables as the K build.
__global__ void __launch_bounds__(KT_THREADS, 2)
grad_contract_tiled_kernel(const GradArgs a) {
  const int b = blockIdx.y;
  if (a.active != nullptr && a.active[b] == 0) return;
  extern __shared__ double smem[];
  __shared__ ProgShared P;
  __shared__ double exp_tab[32];
  __shared__ double th_s[AGPC_MAX_THETA];
  __shared__ int per_slot[AGPC_MAX_LEAVES];
  __shared__ int per_leaf[KT_PER];
  __shared__ int n_per_s;
  if (threadIdx.x < 32) exp_tab[threadIdx.x] = EXP_T[threadIdx.x];
  const int D = a.D;
  double* xr = smem;                          // [D][KT]
  double* xc = smem + D * KT;                 // [D][KT]
  double* trg_r = xc + D * KT;                // [KT_PER][2][KT]  (sin, cos)(pi x_i / T)
  double* trg_c = trg_r + KT_PER * 2 * KT;    // [KT_PER][2][KT]
  double* red = trg_c + KT_PER * 2 * KT;      // [8 warps][AGPC_MAX_THETA]
  const int nt = a.np / KT;
  int ti, tj;
  if (a.mode == 0) {
    const int t = blockIdx.x;
    ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = blockIdx.x / nt;
    tj = blockIdx.x % nt;
  }
  const int n = a.n_of ? a.n_of[b] : a.n_fixed;
  const int row0 = ti * KT, col0 = tj * KT;
  const double* Xf = a.Xf + (long long)b * a.np * D;
  load_x_tiles_async<KT>(xr, xc, Xf, Xf, row0, col0, a.np, a.np, D);
  load_program_raw(P, th_s, a.progs[b], a.theta + (long long)b * a.theta_stride, a.theta_stride);
  for (int i = threadIdx.x; i < 8 * AGPC_MAX_THETA; i += KT_THREADS) red[i] = 0.0;
  cp_async_wait_all();
  __syncthreads();
  derive_leaf_consts(P, th_s);
  if (threadIdx.x == KT_THREADS - 1) {
    int np = 0;
    for (int l = 0; l < P.n_leaves; ++l) {
      const bool tab = P.leaf[l].type == AGPC_LEAF_PER && np < KT_PER;
      per_slot[l] = tab ? np : -1;
      if (tab) per_leaf[np++] = l;
    }
    n_per_s = np;
  }
  __syncthreads();
  const int n_per = n_per_s;
  for (int idx = threadIdx.x; idx < 2 * n_per * KT; idx += KT_THREADS) {
    const int side = idx / (n_per * KT), rem = idx % (n_per * KT);
    const int sl = rem >> 6, r = rem & (KT - 1);
    const LeafConst& lc = P.leaf[per_leaf[sl]];
    double sn, cs;
    sincospi((side ? xc : xr)[lc.dim * KT + r] * lc.a, &sn, &cs);
    double* dst = side ? trg_c : trg_r;
    dst[(sl * 2 + 0) * KT + r] = sn;
    dst[(sl * 2 + 1) * KT + r] = cs;
  }
  __syncthreads();

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* myred = red + warp * AGPC_MAX_THETA;
  const double* Gb = a.G + (long long)b * a.g_batch;
  const double* al = a.alpha ? a.alpha + (long long)b * a.np : nullptr;
  const double* sv = a.s ? a.s + (long long)b * a.np : nullptr;
  const double* ru = a.ru ? a.ru + (long long)b * a.np : nullptr;
  const double* rv = a.rv ? a.rv + (long long)b * a.np : nullptr;
  auto warp_add = [&](int slot, double v) {   // sum over the warp, lane 0 accumulates into the warp's own row
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) myred[slot] += v;
  };

#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    // ---- weights of the 2 x 4 pairs of this pass ----
    double w[2][4];
    int gi[2], gj[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) gi[r] = row0 + ty + 16 * (2 * h + r);
#pragma unroll
    for (int c = 0; c < 4; ++c) gj[c] = col0 + tx + 16 * c;
    // every index lies inside the padded matrix, so all loads are issued unconditionally and together (one memory round
    // trip per pass; a branch per pair made them 8 dependent round trips: 50 % of the first version's stall samples) and
    // the pairs outside the valid triangle are masked afterwards
    double g[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) g[r][c] = Gb[(long long)gi[r] * a.ld + gj[c]];
    if (a.mode == 0) {
      double ali[2], svi[2], alj[4], svj[4], rui[2] = {0.0, 0.0}, rvi[2] = {0.0, 0.0}, ruj[4] = {0.0, 0.0, 0.0, 0.0}, rvj[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        ali[r] = al[gi[r]];
        svi[r] = sv[gi[r]];
        if (ru != nullptr) { rui[r] = ru[gi[r]]; rvi[r] = rv[gi[r]]; }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        alj[c] = al[gj[c]];
        svj[c] = sv[gj[c]];
        if (ru != nullptr) { ruj[c] = ru[gj[c]]; rvj[c] = rv[gj[c]]; }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double wv = 0.5 * (ali[r] * alj[c] - svi[r] * g[r][c] * svj[c]);
          if (ru != nullptr) wv += 0.5 * (rui[r] * rvj[c] + ruj[c] * rvi[r]);
          if (gj[c] < gi[r]) wv *= 2.0;
          const bool ok = gi[r] < n && gj[c] < n && gj[c] <= gi[r];
          w[r][c] = ok ? wv : 0.0;
        }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) w[r][c] = (gi[r] < n && gj[c] < n) ? g[r][c] : 0.0;
    }
    const int n_terms = P.n_terms;
    for (int t = 0; t < n_terms; ++t) {
      const int len = P.term_len[t];
      double arg[2][4];
      double mr[2] = {1.0, 1.0}, mc[4] = {1.0, 1.0, 1.0, 1.0};
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) arg[r][c] = 0.0;
      double scale = 1.0;
      bool has_exp = false;
      // ---- pass A over the leaves: the term's value ----
      for (int p = 0; p < len; ++p) {
        const int leaf_id = P.term_leaf[t][p];
        const LeafConst& lc = P.leaf[leaf_id];
        scale *= lc.var;
        if (lc.type == AGPC_LEAF_CONST) continue;
        if (lc.type == AGPC_LEAF_PER && per_slot[leaf_id] >= 0) {
          has_exp = true;
          const double* tr = trg_r + per_slot[leaf_id] * 2 * KT;
          const double* tc = trg_c + per_slot[leaf_id] * 2 * KT;
          const double hb = 0.5 * lc.b;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const double sr = tr[ty + 16 * (2 * h + r)], cr = tr[KT + ty + 16 * (2 * h + r)];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double sn = sr * tc[KT + tx + 16 * c] - cr * tc[tx + 16 * c];
              arg[r][c] -= hb * sn * sn;
            }
          }
          continue;
        }
        double xi[2], xj[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) xi[r] = xr[lc.dim * KT + ty + 16 * (2 * h + r)];
#pragma unroll
        for (int c = 0; c < 4; ++c) xj[c] = xc[lc.dim * KT + tx + 16 * c];
        if (lc.type == AGPC_LEAF_SE) {
          has_exp = true;
          const double ha = 0.5 * lc.a;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double d = xi[r] - xj[c];
              arg[r][c] -= ha * d * d;
            }
        } else if (lc.type == AGPC_LEAF_PER) {
          has_exp = true;
          const double wq = M_PI * lc.a, hb = 0.5 * lc.b;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double sn = sin(wq * (xi[r] - xj[c]));
              arg[r][c] -= hb * sn * sn;
            }
        } else {  // LIN
#pragma unroll
          for (int r = 0; r < 2; ++r) mr[r] *= xi[r] - lc.a;
#pragma unroll
          for (int c = 0; c < 4; ++c) mc[c] *= xj[c] - lc.a;
        }
      }
      if (has_exp) exp_nonpos_n<8>(reinterpret_cast<double(&)[8]>(arg), exp_tab);
      // arg <- w * (scale * exp part): what every derivative of this term starts from (the LIN factors are separable)
      double s0 = 0.0;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double e = has_exp ? arg[r][c] : 1.0;
          arg[r][c] = w[r][c] * scale * e;
          s0 = fma(arg[r][c], mr[r] * mc[c], s0);
        }
      // S0_t = sum w v_t, booked to every variance of the term
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      // ---- pass B over the leaves: the per-leaf sums ----
      for (int p = 0; p < len; ++p) {
        const int leaf_id = P.term_leaf[t][p];
        const LeafConst& lc = P.leaf[leaf_id];
        const int o = P.leaf_off[leaf_id];
        if (lane == 0) myred[o] += s0 / lc.var;
        if (lc.type == AGPC_LEAF_CONST) continue;
        if (lc.type == AGPC_LEAF_PER && per_slot[leaf_id] >= 0) {
          const double* tr = trg_r + per_slot[leaf_id] * 2 * KT;
          const double* tc = trg_c + per_slot[leaf_id] * 2 * KT;
          const double* xrd = xr + lc.dim * KT;
          const double* xcd = xc + lc.dim * KT;
          double sT = 0.0, sL = 0.0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int lr = ty + 16 * (2 * h + r);
            const double sr = tr[lr], cr = tr[KT + lr], xi = xrd[lr];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int lc4 = tx + 16 * c;
              const double sc = tc[lc4], cc = tc[KT + lc4];
              const double sn = sr * cc - cr * sc, cs = cr * cc + sr * sc;
              const double wv = arg[r][c] * (mr[r] * mc[c]);
              sT = fma(wv * sn * cs, xi - xcd[lc4], sT);
              sL = fma(wv * sn, sn, sL);
            }
          }
          // dk/dT = k sin b cos b (b / T) / l^2 with b = pi r / T;  dk/dl = k sin^2 b / l^3
          warp_add(o + 1, sT * (M_PI * lc.a * lc.a) * lc.b);
          warp_add(o + 2, sL * lc.d);
          continue;
        }
        double xi[2], xj[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) xi[r] = xr[lc.dim * KT + ty + 16 * (2 * h + r)];
#pragma unroll
        for (int c = 0; c < 4; ++c) xj[c] = xc[lc.dim * KT + tx + 16 * c];
        if (lc.type == AGPC_LEAF_SE) {
          double sL = 0.0;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double d = xi[r] - xj[c];
              sL = fma(arg[r][c] * (mr[r] * mc[c]), d * d, sL);
            }
          warp_add(o + 1, sL * lc.b);   // dk/dl = k r^2 / l^3
        } else if (lc.type == AGPC_LEAF_PER) {
          const double wq = M_PI * lc.a;
          double sT = 0.0, sL = 0.0;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              double sn, cs;
              const double d = xi[r] - xj[c];
              sincos(wq * d, &sn, &cs);
              const double wv = arg[r][c] * (mr[r] * mc[c]);
              sT = fma(wv * sn * cs, d, sT);
              sL = fma(wv * sn, sn, sL);
            }
          warp_add(o + 1, sT * (M_PI * lc.a * lc.a) * lc.b);
          warp_add(o + 2, sL * lc.d);
        } else {  // LIN: d/dc of var (x_i - c)(x_j - c) = -var (x_i + x_j - 2c), times the term without this leaf's own factor
          double exr[2] = {1.0, 1.0}, exc[4] = {1.0, 1.0, 1.0, 1.0};
          for (int q = 0; q < len; ++q) {
            if (q == p) continue;
            const LeafConst& lq = P.leaf[P.term_leaf[t][q]];
            if (lq.type != AGPC_LEAF_LIN) continue;
#pragma unroll
            for (int r = 0; r < 2; ++r) exr[r] *= xr[lq.dim * KT + ty + 16 * (2 * h + r)] - lq.a;
#pragma unroll
            for (int c = 0; c < 4; ++c) exc[c] *= xc[lq.dim * KT + tx + 16 * c] - lq.a;
          }
          double sC = 0.0;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              sC = fma(arg[r][c] * (exr[r] * exc[c]), -(xi[r] + xj[c] - 2.0 * lc.a), sC);
          warp_add(o + 1, sC);   // (arg already carries every variance of the term, this leaf's included)
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < P.n_theta) {
    double v = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8 * AGPC_MAX_THETA + threadIdx.x];
    a.partial[((long long)b * a.tiles_per_item + blockIdx.x) * AGPC_MAX_THETA + threadIdx.x] = v;
  }
}

// grad[b][q] = scale * sum_tiles partial[b][tile][q]   (one warp per (b, q), fixed summation order)
__global__ void grad_finalize_kernel(const double* __restrict__ partial, int tiles, const agpc_program* progs,
                                     double scale, double* __restrict__ grad, int grad_stride,
                                     const int* __restrict__ active) {
  const int b = blockIdx.x;
  if (active != nullptr && active[b] == 0) return;
  const int q = blockIdx.y, lane = threadIdx.x & 31;
  if (q >= progs[b].n_theta) return;
  double v = 0.0;
  for (int t = lane; t < tiles; t += 32) v += partial[((long long)b * tiles + t) * AGPC_MAX_THETA + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) grad[(long long)b * grad_stride + q] = scale * v;
}

cudaError_t grad_contract_launch(const Launch& L, const GradArgs& a, int batch, double scale, double* grad,
                                 int grad_stride) {
  ProfScope prof_scope(L, CLS_GRAD);
  const int nt = a.np / GT;
  const int tiles = a.mode == 0 ? nt * (nt + 1) / 2 : nt * nt;
  dim3 grid(tiles, batch);
  static const bool generic = getenv("AGPC_GRAD_GENERIC") && atoi(getenv("AGPC_GRAD_GENERIC")) != 0;   // round-1 kernel, for A/B
  if (generic) {
    const size_t smem = sizeof(double) * (2 * a.D * GT + 8 * AGPC_MAX_THETA);
    grad_contract_kernel<<<grid, 256, smem, L.stream>>>(a);
  } else {
    const size_t smem = sizeof(double) * (2 * (size_t)a.D * KT + 2 * KT_PER * 2 * KT + 8 * AGPC_MAX_THETA);
    bool& configured = func_attr_done(__LINE__);
    if (!configured || smem > 48 * 1024) {
      AGPC_CUDA_TRY(cudaFuncSetAttribute(grad_contract_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(smem > 100 * 1024 ? smem : 100 * 1024)));
      configured = true;
    }
    grad_contract_tiled_kernel<<<grid, KT_THREADS, smem, L.stream>>>(a);
  }
  L.bump();
  AGPC_CUDA_TRY(cudaGetLastError());
  grad_finalize_kernel<<<dim3(batch, AGPC_MAX_THETA), 32, 0, L.stream>>>(a.partial, tiles, a.progs, scale, grad, grad_stride,
                                                                     a.active);
  L.bump();
  return cudaGetLastError();
}

// ---- k(x*, x*) per test point, and column scaling Ks[:, k] *= s_k --------------------------------------------
__global__ void kself_kernel(const agpc_program* __restrict__ progs, const double* __restrict__ theta,
                             int theta_stride, const double* __restrict__ Xs, int m_pad, int D,
                             const int* __restrict__ m_of, double* __restrict__ kss) {
  const int b = blockIdx.y;
  __shared__ ProgShared P;
  load_program(P, progs[b], theta + (long long)b * theta_stride);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_pad) return;
  const double* x = Xs + ((long long)b * m_pad + i) * D;
  double dummy[1];
  kss[(long long)b * m_pad + i] = (i < m_of[b]) ? eval_pair<false>(P, x, 1, x, 1, dummy, 0.0) : 0.0;
}
cudaError_t kself_launch(const Launch& L, const agpc_program* progs, const double* theta, int theta_stride,
                         const double* Xs, int m_pad, int D, const int* m_of, int batch, double* kss) {
  ProfScope prof_scope(L, CLS_MISC);
  dim3 grid((m_pad + 255) / 256, batch);
  kself_kernel<<<grid, 256, 0, L.stream>>>(progs, theta, theta_stride, Xs, m_pad, D, m_of, kss);
  L.bump();
  return cudaGetLastError();
}
__global__ void scale_cols_kernel(double* __restrict__ Ks, const double* __restrict__ s, int rows, int np) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= np || i >= rows) return;
  Ks[((long long)b * rows + i) * np + j] *= s[(long long)b * np + j];
}
cudaError_t scale_cols_launch(const Launch& L, double* Ks, const double* s, int rows, int np, int batch) {
  ProfScope prof_scope(L, CLS_MISC);
  dim3 grid((np + 255) / 256, rows, batch);
  scale_cols_kernel<<<grid, 256, 0, L.stream>>>(Ks, s, rows, np);
  L.bump();
  return cudaGetLastError();
}

// ---- d mu*/d x* = sum_i alpha_i dk(x*, x_i)/dx*  (m.predictive_gradients, gpckernel.py:431) ---------------------
// dk/dx for a product term v = prod f_l * exp(arg):  SE: -v r / l^2;  PER: -v sin b cos b (pi/T) / l^2;
// LIN: (v without the leaf) * var (x' - c);  CONST: 0.  One warp per test point.
constexpr int MAX_D = 32;
__device__ __forceinline__ void eval_pair_dx(const ProgShared& P, const double* xa, const double* xb, double w,
                                             double* gx) {
  for (int t = 0; t < P.n_terms; ++t) {
    const int len = P.term_len[t];
    double arg = 0.0, mult = 1.0;
    double fac[AGPC_MAX_TERM_LEN], dlog[AGPC_MAX_TERM_LEN];
    for (int p = 0; p < len; ++p) {
      const LeafConst& lc = P.leaf[P.term_leaf[t][p]];
      double f = lc.var, dl = 0.0;
      if (lc.type == AGPC_LEAF_SE) {
        const double r = xa[lc.dim] - xb[lc.dim];
        arg -= 0.5 * r * r * lc.a;
        dl = -r * lc.a;
      } else if (lc.type == AGPC_LEAF_PER) {
        const double r = xa[lc.dim] - xb[lc.dim];
        double s, c;
        sincos(M_PI * r * lc.a, &s, &c);
        arg -= 0.5 * s * s * lc.b;
        dl = -s * c * (M_PI * lc.a) * lc.b;
      } else if (lc.type == AGPC_LEAF_LIN) {
        f *= (xa[lc.dim] - lc.a) * (xb[lc.dim] - lc.a);
      }
      mult *= f;
      fac[p] = f;
      dlog[p] = dl;
    }
    const double e = exp(arg);
    const double v = mult * e;
    for (int p = 0; p < len; ++p) {
      const LeafConst& lc = P.leaf[P.term_leaf[t][p]];
      if (lc.type == AGPC_LEAF_SE || lc.type == AGPC_LEAF_PER) {
        gx[lc.dim] += w * v * dlog[p];
      } else if (lc.type == AGPC_LEAF_LIN) {
        double excl = e;
        for (int q = 0; q < len; ++q)
          if (q != p) excl *= fac[q];
        gx[lc.dim] += w * excl * lc.var * (xb[lc.dim] - lc.a);
      }
    }
  }
}
__global__ void __launch_bounds__(256)
dmu_dx_kernel(const agpc_program* __restrict__ prog, const double* __restrict__ theta, const double* __restrict__ Xs,
              int m, const double* __restrict__ Xf, int n, int D, const double* __restrict__ alpha,
              double* __restrict__ out) {
  __shared__ ProgShared P;
  load_program(P, *prog, theta);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= m) return;
  double gx[MAX_D];
  for (int d = 0; d < D; ++d) gx[d] = 0.0;
  const double* xa = Xs + (long long)j * D;
  for (int i = lane; i < n; i += 32) eval_pair_dx(P, xa, Xf + (long long)i * D, alpha[i], gx);
  for (int d = 0; d < D; ++d) {
    double v = gx[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[(long long)j * D + d] = v;
  }
}
cudaError_t dmu_dx_launch(const Launch& L, const agpc_program* prog, const double* theta, const double* Xs, int m,
                          const double* Xf, int n, int D, const double* alpha, double* out) {
  ProfScope prof_scope(L, CLS_MISC);
  if (D > MAX_D) return cudaErrorInvalidValue;
  dmu_dx_kernel<<<(m + 7) / 8, 256, 0, L.stream>>>(prog, theta, Xs, m, Xf, n, D, alpha, out);
  L.bump();
  return cudaGetLastError();
}

int grad_tiles(int np, int mode) {
  const int nt = np / GT;
  return mode == 0 ? nt * (nt + 1) / 2 : nt * nt;
}

}  // namespace agpc
