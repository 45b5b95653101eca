// On-device lock-step optimiser: the loop GPy runs as m.optimize() (gpckernel.py:246) -- paramz
// opt_lbfgsb = SciPy L-BFGS-B (m = 10, factr = 1e7, pgtol = 1e-5, maxfun evaluations) over the transformed parameters
// (paramz Logexp / Logistic) -- for a whole batch of models at once, with the iterates, the limited-memory pairs, the
// line-search state and the EP sites of every model resident in HBM.  One round = one scoring pass over the models that
// still run (score_items_dev, sites warm-started) + one small kernel that advances every model's state machine to its
// next trial point; the host only reads back the "still running" flags to pack the next round.
//
// The problem is unconstrained in optimiser space (the constraints live in the transforms), so L-BFGS-B's generalized
// Cauchy point and subspace minimisation reduce to the quasi-Newton step d = -H g of the limited-memory BFGS matrix with
// H0 = (s'y / y'y) I, which is what this file computes (two-loop recursion).  Everything else follows L-BFGS-B 3.0's
// driver (Zhu, Byrd, Lu, Nocedal, ACM TOMS 23, 1997; Morales & Nocedal 2011): first step 1 / |g|, later steps 1, the
// Moré-Thuente line search (dcsrch / dcstep of MINPACK-2: ftol = 1e-3, gtol = 0.9, xtol = 0.1, at most maxls = 20
// evaluations, then restart from steepest descent with the memory dropped, or stop if the memory is empty), pair
// skipped when s'y <= eps * (-g'd step), stop on max |g| <= pgtol or (f_old - f) <= factr * eps * max(|f_old|, |f|, 1),
// SciPy's maxfun / maxiter tests at every new iterate, paramz's treatment of a failed evaluation (objective 1e300, the
// last good gradient, at most 10 in a row).  Same algorithm, not the same rounding: the host driver (optim.py, SciPy's
// own setulb) stays the default where bit-for-bit trajectories matter; tests/test_gpu_parity.py checks that both end in
// the same optimum and select the same structures.
#include <cfloat>
#include <cstdio>
#include <vector>

#include "engine.h"

using namespace agpc;

namespace {

constexpr int PMAX = AGPC_MAX_THETA;   // 48 parameters at most: lane l of the item's warp owns l and l + 32
constexpr int MMAX = 10;               // limited-memory pairs
constexpr double FAIL_VALUE = 1e300;
constexpr int ALLOWED_FAILURES = 10;
constexpr double LOGEXP_LIM = 36.0;
constexpr double EPSMCH = 2.220446049250313e-16;

enum : int { PH_FIRST = 0, PH_LINESEARCH = 1 };
enum : int { TASK_RUNNING = 0, TASK_CONV_PG = 1, TASK_CONV_F = 2, TASK_STOP_MAXFUN = 3, TASK_STOP_MAXITER = 4, TASK_ABNORMAL = 5,
             TASK_TOO_MANY_FAILURES = 6 };

struct ItemState {
  double x[PMAX];      // current trial point
  double t[PMAX];      // iterate the line search started from
  double r[PMAX];      // gradient at t
  double d[PMAX];      // search direction
  double z[PMAX];      // t + d (the unit step is taken as this sum, like L-BFGS-B's z)
  double g[PMAX];      // gradient at x (optimiser space)
  double glast[PMAX];  // transformed gradient of the last successful evaluation
  double S[MMAX][PMAX], Y[MMAX][PMAX];
  double sy[MMAX];
  double theta, f, fold, gd, gdold, stp, dnorm, sbgnrm;
  // dcsrch
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
  int brackt, stage;
  int col, head, iter, nfev, ifun, iback, nskip, fails, phase, task, have_glast, n;
};

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// dot product over the item's n parameters (lane l: l and l + 32), identical on every lane
__device__ __forceinline__ double wdot(const double* a, const double* b, int n, int lane) {
  double v = 0.0;
  if (lane < n) v = a[lane] * b[lane];
  if (lane + 32 < n) v = fma(a[lane + 32], b[lane + 32], v);
  return wsum(v);
}

// paramz transformations (paramz/transformations.py): kind 0 free, 1 Logexp (positive), 2 Logistic (bounded)
__device__ __forceinline__ double from_x(int kind, double x, double lo, double hi) {
  if (kind == 1) return x > LOGEXP_LIM ? x : log1p(exp(fmax(x, -700.0)));
  if (kind == 2) return x > -700.0 ? lo + (hi - lo) / (1.0 + exp(-x)) : lo;
  return x;
}
__device__ __forceinline__ double to_x(int kind, double v, double lo, double hi) {
  if (kind == 1) return v > LOGEXP_LIM ? v : log(expm1(v));
  if (kind == 2) {
    const double w = hi - lo;
    v = fmin(fmax(v, lo + 1e-10 * w), hi - 1e-10 * w);
    return log((v - lo) / (hi - v));
  }
  return v;
}
__device__ __forceinline__ double gradfactor(int kind, double v, double lo, double hi) {
  if (kind == 1) return v > LOGEXP_LIM ? 1.0 : -expm1(-v);
  if (kind == 2) return (v - lo) * (hi - v) / (hi - lo);
  return 1.0;
}

// ---- Moré-Thuente safeguarded step (MINPACK-2 dcstep): new trial step from the interval (stx, sty) and the point stp ----
__device__ void mt_step(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp,
                        double dp, int& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {  // higher value: the minimum is bracketed; cubic vs quadratic, whichever is closer to stx
    const double th = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(th), fmax(fabs(dx), fabs(dp)));
    double gam = s * sqrt((th / s) * (th / s) - (dx / s) * (dp / s));
    if (stp < stx) gam = -gam;
    const double p = (gam - dx) + th, q = ((gam - dx) + gam) + dp, r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = fabs(stpc - stx) < fabs(stpq - stx) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {  // lower value, derivatives of opposite sign: bracketed; cubic vs secant, the farther from stp
    const double th = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(th), fmax(fabs(dx), fabs(dp)));
    double gam = s * sqrt((th / s) * (th / s) - (dx / s) * (dp / s));
    if (stp > stx) gam = -gam;
    const double p = (gam - dp) + th, q = ((gam - dp) + gam) + dx, r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {  // lower value, same sign, derivative shrinks: cubic (if it tends the right way) vs secant
    const double th = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(th), fmax(fabs(dx), fabs(dp)));
    double gam = s * sqrt(fmax(0.0, (th / s) * (th / s) - (dx / s) * (dp / s)));
    if (stp > stx) gam = -gam;
    const double p = (gam - dp) + th, q = (gam + (dx - dp)) + gam, r = p / q;
    double stpc;
    if (r < 0.0 && gam != 0.0) stpc = stp + r * (stx - stp);
    else stpc = stp > stx ? stpmax : stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = fabs(stpc - stp) < fabs(stpq - stp) ? stpc : stpq;
      stpf = stp > stx ? fmin(stp + 0.66 * (sty - stp), stpf) : fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
      stpf = fmax(stpmin, fmin(stpmax, stpf));
    }
  } else {  // lower value, same sign, derivative does not shrink: cubic through sty if bracketed, else the interval end
    if (brackt) {
      const double th = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = fmax(fabs(th), fmax(fabs(dy), fabs(dp)));
      double gam = s * sqrt((th / s) * (th / s) - (dy / s) * (dp / s));
      if (stp > sty) gam = -gam;
      const double p = (gam - dp) + th, q = ((gam - dp) + gam) + dy, r = p / q;
      stpf = stp + r * (sty - stp);
    } else {
      stpf = stp > stx ? stpmax : stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// one call of the line search with the value f and slope g at the current step; returns true when the step is accepted
// (sufficient decrease + curvature, or one of dcsrch's warnings), false when another evaluation at it.stp is needed
__device__ bool mt_search(ItemState& it, double f, double g, double stpmin, double stpmax, bool start) {
  constexpr double ftol = 1e-3, gtol = 0.9, xtol = 0.1, xtrapl = 1.1, xtrapu = 4.0;
  if (start) {
    it.brackt = 0; it.stage = 1; it.finit = f; it.ginit = g; it.gtest = ftol * g;
    it.width = stpmax - stpmin; it.width1 = it.width / 0.5;
    it.stx = 0.0; it.fx = f; it.gx = g; it.sty = 0.0; it.fy = f; it.gy = g;
    it.stmin = 0.0; it.stmax = it.stp + xtrapu * it.stp;
    return false;
  }
  const double ftest = it.finit + it.stp * it.gtest;
  if (it.stage == 1 && f <= ftest && g >= 0.0) it.stage = 2;
  bool accept = false;
  if (it.brackt && (it.stp <= it.stmin || it.stp >= it.stmax)) accept = true;    // rounding errors prevent progress
  if (it.brackt && it.stmax - it.stmin <= xtol * it.stmax) accept = true;        // xtol test satisfied
  if (it.stp == stpmax && f <= ftest && g <= it.gtest) accept = true;            // stp = stpmax
  if (it.stp == stpmin && (f > ftest || g >= it.gtest)) accept = true;           // stp = stpmin
  if (f <= ftest && fabs(g) <= gtol * (-it.ginit)) accept = true;                // strong Wolfe conditions hold
  if (accept) return true;
  if (it.stage == 1 && f <= it.fx && f > ftest) {  // first stage: work on f(stp) - f(0) - ftol stp f'(0)
    const double fm = f - it.stp * it.gtest;
    double fxm = it.fx - it.stx * it.gtest, fym = it.fy - it.sty * it.gtest;
    const double gm = g - it.gtest;
    double gxm = it.gx - it.gtest, gym = it.gy - it.gtest;
    mt_step(it.stx, fxm, gxm, it.sty, fym, gym, it.stp, fm, gm, it.brackt, it.stmin, it.stmax);
    it.fx = fxm + it.stx * it.gtest; it.fy = fym + it.sty * it.gtest;
    it.gx = gxm + it.gtest; it.gy = gym + it.gtest;
  } else {
    mt_step(it.stx, it.fx, it.gx, it.sty, it.fy, it.gy, it.stp, f, g, it.brackt, it.stmin, it.stmax);
  }
  if (it.brackt) {  // bisect when the interval does not shrink fast enough
    if (fabs(it.sty - it.stx) >= 0.66 * it.width1) it.stp = it.stx + 0.5 * (it.sty - it.stx);
    it.width1 = it.width;
    it.width = fabs(it.sty - it.stx);
  }
  if (it.brackt) {
    it.stmin = fmin(it.stx, it.sty);
    it.stmax = fmax(it.stx, it.sty);
  } else {
    it.stmin = it.stp + xtrapl * (it.stp - it.stx);
    it.stmax = it.stp + xtrapu * (it.stp - it.stx);
  }
  it.stp = fmin(fmax(it.stp, stpmin), stpmax);
  if ((it.brackt && (it.stp <= it.stmin || it.stp >= it.stmax)) || (it.brackt && it.stmax - it.stmin <= xtol * it.stmax))
    it.stp = it.stx;
  return false;
}

struct LbfgsOpts {
  int m, max_evals, maxiter, maxls;
  double factr, pgtol;
};

// ---- kernels -----------------------------------------------------------------------------------------------------
// x = finv(theta0)
__global__ void optim_init_kernel(ItemState* st, const double* __restrict__ theta0, int theta_stride, const int* __restrict__ kinds,
                                  const double* __restrict__ lo, const double* __restrict__ hi, const int* __restrict__ n_theta,
                                  int n_items) {
  const int b = blockIdx.x, lane = threadIdx.x;
  if (b >= n_items) return;
  ItemState& it = st[b];
  const int n = n_theta[b];
  for (int i = lane; i < PMAX; i += 32) {
    const long long o = (long long)b * theta_stride + i;
    it.x[i] = i < n ? to_x(kinds[o], theta0[o], lo[o], hi[o]) : 0.0;
    it.t[i] = it.x[i];
    it.glast[i] = 0.0;
  }
  if (lane == 0) {
    it.n = n; it.col = 0; it.head = 0; it.iter = 0; it.nfev = 0; it.ifun = 0; it.iback = 0; it.nskip = 0; it.fails = 0;
    it.phase = PH_FIRST; it.task = TASK_RUNNING; it.have_glast = 0; it.theta = 1.0; it.f = 0.0; it.fold = 0.0;
  }
}

// theta_c[k] = f(x) of live item idx[k] (or of its accepted iterate t when `accepted`); optional gather of the sites
__global__ void optim_gather_kernel(const ItemState* __restrict__ st, const int* __restrict__ idx, int n_live, int accepted,
                                    const int* __restrict__ kinds, const double* __restrict__ lo, const double* __restrict__ hi,
                                    int theta_stride, double* __restrict__ theta_c) {
  const int k = blockIdx.x, lane = threadIdx.x;
  if (k >= n_live) return;
  const int b = idx[k];
  const ItemState& it = st[b];
  for (int i = lane; i < theta_stride; i += 32) {
    const long long o = (long long)b * theta_stride + i;
    double v = 0.0;
    if (i < it.n) v = from_x(kinds[o], accepted ? it.t[i] : it.x[i], lo[o], hi[o]);
    theta_c[(long long)k * theta_stride + i] = v;
  }
}
// dst[k][i] = src[idx[k]][i]  (gather = 1)   or   dst[idx[k]][i] = src[k][i] where ok[k]  (gather = 0), rows of `len` doubles
__global__ void optim_move_rows_kernel(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ idx,
                                       int len, int gather, const int* __restrict__ status_c) {
  const int k = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const long long b = idx[k];
  if (gather) {
    dst[(long long)k * len + i] = src[b * len + i];
  } else {
    const int s = status_c[k];
    if (s == AGPC_ST_NOT_PD || s == AGPC_ST_NONFINITE) return;  // a failed evaluation leaves the model's sites alone
    dst[b * len + i] = src[(long long)k * len + i];
  }
}

// d = -H g: two-loop recursion over the stored pairs (newest first), H0 = I / theta
__device__ void lbfgs_direction(ItemState& it, int lane) {
  const int n = it.n;
  double q0 = lane < n ? it.g[lane] : 0.0, q1 = lane + 32 < n ? it.g[lane + 32] : 0.0;
  double alpha[MMAX];
  for (int j = 0; j < it.col; ++j) {
    const int p = (it.head + it.col - 1 - j) % MMAX;  // newest pair first
    double v = 0.0;
    if (lane < n) v = it.S[p][lane] * q0;
    if (lane + 32 < n) v = fma(it.S[p][lane + 32], q1, v);
    const double a = wsum(v) / it.sy[p];
    alpha[j] = a;
    if (lane < n) q0 = fma(-a, it.Y[p][lane], q0);
    if (lane + 32 < n) q1 = fma(-a, it.Y[p][lane + 32], q1);
  }
  q0 /= it.theta;
  q1 /= it.theta;
  for (int j = it.col - 1; j >= 0; --j) {
    const int p = (it.head + it.col - 1 - j) % MMAX;
    double v = 0.0;
    if (lane < n) v = it.Y[p][lane] * q0;
    if (lane + 32 < n) v = fma(it.Y[p][lane + 32], q1, v);
    const double bta = wsum(v) / it.sy[p];
    if (lane < n) q0 = fma(alpha[j] - bta, it.S[p][lane], q0);
    if (lane + 32 < n) q1 = fma(alpha[j] - bta, it.S[p][lane + 32], q1);
  }
  if (lane < n) it.d[lane] = -q0;
  if (lane + 32 < n) it.d[lane + 32] = -q1;
  __syncwarp();
}

// Consumes the evaluation of live item k and runs its state machine up to the next evaluation request.  One warp per item.
__global__ void optim_advance_kernel(ItemState* st, const int* __restrict__ idx, int n_live, const double* __restrict__ nlml_c,
                                     const double* __restrict__ grad_c, const int* __restrict__ status_c, int theta_stride,
                                     const int* __restrict__ kinds, const double* __restrict__ lo, const double* __restrict__ hi,
                                     const double* __restrict__ theta_c, LbfgsOpts o, int* __restrict__ live_out) {
  const int k = blockIdx.x, lane = threadIdx.x;
  if (k >= n_live) return;
  const int b = idx[k];
  ItemState& it = st[b];
  const int n = it.n;
  const double stpmx = 1e10;

  // ---- the evaluation: transformed gradient, or paramz's stand-in for a failed one ----
  double f = nlml_c[k];
  const int status = status_c[k];
  const bool failed = status == AGPC_ST_NOT_PD || status == AGPC_ST_NONFINITE || !isfinite(f);
  int fails = it.fails;
  const int nfev = it.nfev + 1;
  __syncwarp();
  if (lane == 0) it.nfev = nfev;
  if (failed) {
    ++fails;
    f = FAIL_VALUE;
    for (int i = lane; i < n; i += 32) it.g[i] = it.have_glast ? it.glast[i] : 0.0;
  } else {
    fails = 0;
    for (int i = lane; i < n; i += 32) {
      const long long oo = (long long)b * theta_stride + i;
      const double th = theta_c[(long long)k * theta_stride + i];
      double gi = grad_c[(long long)k * theta_stride + i] * gradfactor(kinds[oo], th, lo[oo], hi[oo]);
      gi = fmin(fmax(gi, -1e10), 1e10);   // paramz clips the transformed gradient
      it.g[i] = gi;
      it.glast[i] = gi;
    }
    if (lane == 0) it.have_glast = 1;
  }
  __syncwarp();
  if (lane == 0) it.fails = fails;
  if (fails > ALLOWED_FAILURES) {
    if (lane == 0) { it.task = TASK_TOO_MANY_FAILURES; live_out[k] = 0; }
    return;
  }
  int task = TASK_RUNNING;
  bool new_iteration = false;

  if (it.phase == PH_FIRST) {
    double m = 0.0;
    for (int i = lane; i < n; i += 32) m = fmax(m, fabs(it.g[i]));
    const double sbgnrm = wmax(m);
    if (lane == 0) { it.f = f; it.sbgnrm = sbgnrm; }
    if (sbgnrm <= o.pgtol) task = TASK_CONV_PG;
    else new_iteration = true;
  } else {
    // ---- inside a line search: value f, slope g'd at t + stp d ----
    const double gd = wdot(it.g, it.d, n, lane);
    // the scalar recurrences of the line search run on lane 0; the outcome is broadcast
    int acc_i = 0;
    if (lane == 0) acc_i = mt_search(it, f, gd, 0.0, stpmx, false) ? 1 : 0;
    acc_i = __shfl_sync(0xffffffffu, acc_i, 0);
    const bool accepted = acc_i != 0;
    __syncwarp();
    if (!accepted) {
      int ifun = it.ifun + 1;
      const int iback = ifun - 1;
      if (iback >= o.maxls) {
        // line search failed: back to the iterate it started from; drop the memory and retry along -g, or give up
        for (int i = lane; i < n; i += 32) { it.x[i] = it.t[i]; it.g[i] = it.r[i]; }
        __syncwarp();
        f = it.fold;
        if (lane == 0) it.f = f;
        if (it.col == 0) {
          task = TASK_ABNORMAL;
        } else {
          if (lane == 0) { it.col = 0; it.head = 0; it.theta = 1.0; }
          __syncwarp();
          new_iteration = true;
        }
      } else {
        if (lane == 0) { it.ifun = ifun; it.iback = iback; }
        const double stp = it.stp;
        for (int i = lane; i < n; i += 32) it.x[i] = stp == 1.0 ? it.z[i] : fma(stp, it.d[i], it.t[i]);
        if (lane == 0) live_out[k] = 1;
        return;
      }
    } else {
      // ---- new iterate ----
      const int iter = it.iter + 1;
      if (lane == 0) { it.iter = iter; it.f = f; }
      double m = 0.0;
      for (int i = lane; i < n; i += 32) m = fmax(m, fabs(it.g[i]));
      const double sbgnrm = wmax(m);
      if (lane == 0) it.sbgnrm = sbgnrm;
      const double fold = it.fold;
      if (iter >= o.maxiter) task = TASK_STOP_MAXITER;
      else if (nfev > o.max_evals) task = TASK_STOP_MAXFUN;
      else if (sbgnrm <= o.pgtol) task = TASK_CONV_PG;
      else if (fold - f <= o.factr * EPSMCH * fmax(fmax(fabs(fold), fabs(f)), 1.0)) task = TASK_CONV_F;
      // the iterate is accepted either way: t <- x
      // ---- limited-memory update with s = stp d, y = g - r ----
      const double stp = it.stp;
      double rr_l = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double y = it.g[i] - it.r[i];
        it.r[i] = y;
        rr_l = fma(y, y, rr_l);
      }
      const double rr = wsum(rr_l);
      double dr, ddum;
      if (stp == 1.0) { dr = gd - it.gdold; ddum = -it.gdold; }
      else { dr = (gd - it.gdold) * stp; ddum = -it.gdold * stp; }
      __syncwarp();
      if (task == TASK_RUNNING) {
        if (dr <= EPSMCH * ddum) {
          if (lane == 0) it.nskip += 1;
        } else {
          const int col = it.col, head = it.head;
          __syncwarp();
          const int slot = col < o.m ? (head + col) % MMAX : head;
          for (int i = lane; i < n; i += 32) {
            it.S[slot][i] = stp == 1.0 ? it.d[i] : stp * it.d[i];
            it.Y[slot][i] = it.r[i];
          }
          if (lane == 0) {
            it.sy[slot] = dr;
            it.theta = rr / dr;
            if (col < o.m) it.col = col + 1;
            else it.head = (head + 1) % MMAX;
          }
        }
        __syncwarp();
        new_iteration = true;
      }
      for (int i = lane; i < n; i += 32) it.t[i] = it.x[i];
      __syncwarp();
    }
  }

  if (task != TASK_RUNNING) {
    if (lane == 0) { it.task = task; live_out[k] = 0; }
    // the accepted iterate is what the caller gets back
    if (it.phase == PH_FIRST)
      for (int i = lane; i < n; i += 32) it.t[i] = it.x[i];
    return;
  }
  if (new_iteration) {
    for (;;) {
      // ---- search direction and the start of its line search ----
      if (it.col == 0) {
        for (int i = lane; i < n; i += 32) it.d[i] = -it.g[i];
        __syncwarp();
      } else {
        lbfgs_direction(it, lane);
      }
      const double dtd = wdot(it.d, it.d, n, lane);
      const double dnorm = sqrt(dtd);
      const double gd = wdot(it.g, it.d, n, lane);
      if (gd >= 0.0) {  // not a descent direction: same exit as a failed line search
        if (it.col == 0) {
          if (lane == 0) { it.task = TASK_ABNORMAL; live_out[k] = 0; }
          for (int i = lane; i < n; i += 32) it.t[i] = it.x[i];
          return;
        }
        if (lane == 0) { it.col = 0; it.head = 0; it.theta = 1.0; }
        __syncwarp();
        continue;
      }
      const double stp = (it.iter == 0) ? fmin(1.0 / dnorm, stpmx) : 1.0;
      for (int i = lane; i < n; i += 32) {
        it.t[i] = it.x[i];
        it.r[i] = it.g[i];
        it.z[i] = it.x[i] + it.d[i];
      }
      __syncwarp();
      if (lane == 0) {
        it.fold = it.f; it.dnorm = dnorm; it.gd = gd; it.gdold = gd; it.stp = stp; it.ifun = 1; it.iback = 0;
        it.phase = PH_LINESEARCH;
        mt_search(it, it.f, gd, 0.0, stpmx, true);
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) it.x[i] = stp == 1.0 ? it.z[i] : fma(stp, it.d[i], it.t[i]);
      if (lane == 0) live_out[k] = 1;
      return;
    }
  }
}

__global__ void optim_report_kernel(const ItemState* __restrict__ st, int n_items, int* __restrict__ evals, int* __restrict__ iters,
                                    int* __restrict__ task) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_items) return;
  evals[b] = st[b].nfev;
  iters[b] = st[b].iter;
  task[b] = st[b].task;
}

int fail(agpc_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  else snprintf(buf, sizeof buf, "%s", what);
  ctx->err = buf;
  return code;
}
#define CK(expr)                                                      \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) return fail(ctx, AGPC_E_CUDA, #expr, _e);  \
  } while (0)

struct Carve {  // bump allocator over one grow-only device buffer (base == nullptr: size pass)
  char* base;
  size_t off = 0;
  template <class T>
  T* get(size_t n) {
    off = (off + 255) / 256 * 256;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += sizeof(T) * (n ? n : 1);
    return p;
  }
};

}  // namespace

extern "C" int agpc_optimize_batch(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                                   double* theta, int32_t theta_stride, const int32_t* kinds, const double* lo,
                                   const double* hi, const int32_t* rows, const int64_t* row_off,
                                   const agpc_ep_options* ep, const agpc_lbfgs_options* lb, double* tau, double* nu,
                                   int32_t site_stride, double* nlml, double* grad, double* bic, double* nlml_gauss,
                                   int32_t* sweeps, int32_t* status, int32_t* evals, int32_t* iters, int32_t* task) {
  if (!ctx) return AGPC_E_ARG;
  if (n_items <= 0 || !programs || !theta || !kinds || !lo || !hi || !rows || !row_off || !ep || !lb || !nlml || !status)
    return fail(ctx, AGPC_E_ARG, "optimize_batch: null argument");
  if (theta_stride <= 0 || theta_stride > PMAX) return fail(ctx, AGPC_E_ARG, "optimize_batch: theta_stride must be 1..48");
  if (lb->m < 1 || lb->m > MMAX || lb->maxls < 1 || lb->max_evals < 1)
    return fail(ctx, AGPC_E_ARG, "optimize_batch: bad L-BFGS options (m in 1..10)");
  if (ep->inference != AGPC_INFER_EP) return fail(ctx, AGPC_E_UNSUPPORTED, "optimize_batch: EP inference only");
  if (dataset_id < 0 || dataset_id >= (int)ctx->datasets.size() || !ctx->datasets[dataset_id].X)
    return fail(ctx, AGPC_E_ARG, "optimize_batch: bad dataset id");
  const int N = ctx->datasets[dataset_id].N;
  if (row_off[0] != 0) return fail(ctx, AGPC_E_ARG, "optimize_batch: row_off[0] must be 0");
  int n_max = 0;
  for (int b = 0; b < n_items; ++b) {
    if (row_off[b + 1] <= row_off[b]) return fail(ctx, AGPC_E_ARG, "optimize_batch: empty or unordered row list");
    n_max = std::max<int>(n_max, (int)(row_off[b + 1] - row_off[b]));
    if (programs[b].n_theta <= 0 || programs[b].n_theta > theta_stride) return fail(ctx, AGPC_E_ARG, "optimize_batch: bad n_theta");
  }
  for (int64_t i = 0; i < row_off[n_items]; ++i)
    if (rows[i] < 0 || rows[i] >= N) return fail(ctx, AGPC_E_ARG, "optimize_batch: row index outside the dataset");
  if ((tau || nu) && site_stride < n_max) return fail(ctx, AGPC_E_ARG, "optimize_batch: site_stride < n");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int ss = n_max;  // internal site pitch
  const size_t nth = (size_t)n_items * theta_stride, n_rows = (size_t)row_off[n_items];

  ItemState* d_st;
  double *d_theta0, *d_lo, *d_hi, *d_tau, *d_nu, *c_theta, *c_tau, *c_nu, *c_nlml, *c_grad, *c_bic, *c_gauss;
  int *d_kinds, *d_nth, *d_rows, *d_idx, *d_live, *c_sweeps, *c_status, *d_rep;
  auto carve = [&](char* base) {
    Carve D{base};
    d_st = D.get<ItemState>(n_items);
    d_theta0 = D.get<double>(nth); d_kinds = D.get<int>(nth); d_lo = D.get<double>(nth); d_hi = D.get<double>(nth);
    d_nth = D.get<int>(n_items); d_rows = D.get<int>(n_rows);
    d_tau = D.get<double>((size_t)n_items * ss); d_nu = D.get<double>((size_t)n_items * ss);
    d_idx = D.get<int>(n_items); d_live = D.get<int>(n_items);
    c_theta = D.get<double>(nth); c_tau = D.get<double>((size_t)n_items * ss); c_nu = D.get<double>((size_t)n_items * ss);
    c_nlml = D.get<double>(n_items); c_grad = D.get<double>(nth); c_bic = D.get<double>(n_items); c_gauss = D.get<double>(n_items);
    c_sweeps = D.get<int>(n_items); c_status = D.get<int>(n_items); d_rep = D.get<int>(3 * (size_t)n_items);
    return D.off + 256;
  };
  const size_t total = carve(nullptr);
  if (ctx->opt_stage_bytes < total) {  // grow-only: no cudaMalloc / cudaFree on the steady-state path
    if (ctx->opt_stage) cudaFree(ctx->opt_stage);
    ctx->opt_stage = nullptr;
    ctx->opt_stage_bytes = 0;
    const size_t want = total + total / 4;
    if (cudaMalloc(&ctx->opt_stage, want) != cudaSuccess) return fail(ctx, AGPC_E_NOMEM, "optimize_batch: cudaMalloc");
    ctx->opt_stage_bytes = want;
  }
  carve(static_cast<char*>(ctx->opt_stage));

  std::vector<int> h_nth(n_items);
  for (int b = 0; b < n_items; ++b) h_nth[b] = programs[b].n_theta;
  CK(cudaMemcpyAsync(d_theta0, theta, sizeof(double) * nth, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_kinds, kinds, sizeof(int) * nth, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_lo, lo, sizeof(double) * nth, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_hi, hi, sizeof(double) * nth, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_nth, h_nth.data(), sizeof(int) * n_items, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_rows, rows, sizeof(int) * n_rows, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(d_tau, 0, sizeof(double) * (size_t)n_items * ss, s));
  CK(cudaMemsetAsync(d_nu, 0, sizeof(double) * (size_t)n_items * ss, s));
  optim_init_kernel<<<n_items, 32, 0, s>>>(d_st, d_theta0, theta_stride, d_kinds, d_lo, d_hi, d_nth, n_items);
  ctx->launches += 1;
  CK(cudaGetLastError());

  LbfgsOpts o{lb->m, lb->max_evals, lb->maxiter > 0 ? lb->maxiter : 15000, lb->maxls, lb->factr, lb->pgtol};
  std::vector<int> live(n_items), flags(n_items);
  for (int b = 0; b < n_items; ++b) live[b] = b;
  std::vector<agpc_program> progs_c(n_items);
  std::vector<int64_t> start_c(n_items);
  std::vector<int32_t> len_c(n_items);
  agpc_ep_options epo = *ep;
  const dim3 tb(256);

  // one scoring pass over `sel` (theta from the trial point x, or from the accepted iterate t), results in the c_* arrays
  auto score = [&](const std::vector<int>& sel, bool accepted, bool warm) -> int {
    const int nl = (int)sel.size();
    for (int k = 0; k < nl; ++k) {
      progs_c[k] = programs[sel[k]];
      start_c[k] = row_off[sel[k]];
      len_c[k] = (int32_t)(row_off[sel[k] + 1] - row_off[sel[k]]);
    }
    CK(cudaMemcpyAsync(d_idx, sel.data(), sizeof(int) * nl, cudaMemcpyHostToDevice, s));
    optim_gather_kernel<<<nl, 32, 0, s>>>(d_st, d_idx, nl, accepted ? 1 : 0, d_kinds, d_lo, d_hi, theta_stride, c_theta);
    if (warm) {
      const dim3 g((ss + 255) / 256, nl);
      optim_move_rows_kernel<<<g, tb, 0, s>>>(c_tau, d_tau, d_idx, ss, 1, nullptr);
      optim_move_rows_kernel<<<g, tb, 0, s>>>(c_nu, d_nu, d_idx, ss, 1, nullptr);
      ctx->launches += 2;
    }
    ctx->launches += 1;
    CK(cudaGetLastError());
    epo.warm_start = warm ? 1 : 0;
    int rc = score_items_dev(ctx, dataset_id, nl, progs_c.data(), c_theta, theta_stride, d_rows, start_c.data(), len_c.data(),
                             &epo, c_tau, c_nu, ss, c_nlml, c_grad, c_bic, c_gauss, c_sweeps, c_status, s);
    if (rc) return rc;
    const dim3 g((ss + 255) / 256, nl);
    optim_move_rows_kernel<<<g, tb, 0, s>>>(d_tau, c_tau, d_idx, ss, 0, c_status);
    optim_move_rows_kernel<<<g, tb, 0, s>>>(d_nu, c_nu, d_idx, ss, 0, c_status);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return AGPC_OK;
  };

  // ---- rounds ----
  for (int round = 0; !live.empty(); ++round) {
    int rc = score(live, false, round > 0);
    if (rc) return rc;
    const int nl = (int)live.size();
    optim_advance_kernel<<<nl, 32, 0, s>>>(d_st, d_idx, nl, c_nlml, c_grad, c_status, theta_stride, d_kinds, d_lo, d_hi, c_theta,
                                           o, d_live);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(flags.data(), d_live, sizeof(int) * nl, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int> next;
    next.reserve(nl);
    for (int k = 0; k < nl; ++k)
      if (flags[k]) next.push_back(live[k]);
    live.swap(next);
    if (round > 4 * o.max_evals + 64) return fail(ctx, AGPC_E_CUDA, "optimize_batch: round limit exceeded (internal error)");
  }

  // ---- leave every model at its accepted iterate: one more pass over all of them, sites warm-started ----
  std::vector<int> all(n_items);
  for (int b = 0; b < n_items; ++b) all[b] = b;
  int rc = score(all, true, true);
  if (rc) return rc;
  CK(cudaMemcpyAsync(theta, c_theta, sizeof(double) * nth, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(nlml, c_nlml, sizeof(double) * n_items, cudaMemcpyDeviceToHost, s));
  if (grad) CK(cudaMemcpyAsync(grad, c_grad, sizeof(double) * nth, cudaMemcpyDeviceToHost, s));
  if (bic) CK(cudaMemcpyAsync(bic, c_bic, sizeof(double) * n_items, cudaMemcpyDeviceToHost, s));
  if (nlml_gauss) CK(cudaMemcpyAsync(nlml_gauss, c_gauss, sizeof(double) * n_items, cudaMemcpyDeviceToHost, s));
  if (sweeps) CK(cudaMemcpyAsync(sweeps, c_sweeps, sizeof(int) * n_items, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(status, c_status, sizeof(int) * n_items, cudaMemcpyDeviceToHost, s));
  if (tau && nu) {
    CK(cudaMemcpy2DAsync(tau, sizeof(double) * site_stride, c_tau, sizeof(double) * ss, sizeof(double) * ss, n_items, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpy2DAsync(nu, sizeof(double) * site_stride, c_nu, sizeof(double) * ss, sizeof(double) * ss, n_items, cudaMemcpyDeviceToHost, s));
  }
  optim_report_kernel<<<(n_items + 255) / 256, 256, 0, s>>>(d_st, n_items, d_rep, d_rep + n_items, d_rep + 2 * (size_t)n_items);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (evals) CK(cudaMemcpyAsync(evals, d_rep, sizeof(int) * n_items, cudaMemcpyDeviceToHost, s));
  if (iters) CK(cudaMemcpyAsync(iters, d_rep + n_items, sizeof(int) * n_items, cudaMemcpyDeviceToHost, s));
  if (task) CK(cudaMemcpyAsync(task, d_rep + 2 * (size_t)n_items, sizeof(int) * n_items, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return AGPC_OK;
}
