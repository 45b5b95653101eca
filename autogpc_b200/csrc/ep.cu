// EP site updates, posterior marginals and the log-marginal reductions (warp-reduction kernels).
//
// Stands in for GPy EP.expectation_propagation / Bernoulli.moments_match_ep / EP.inference's scalar algebra
//  for gpckernel.py:245-246, restated for the *parallel* EP schedule (oracle.ep_parallel):
//   marginals:  diag(Sigma)_i = (1 - (B^-1)_ii) / tau~_i,   (B^-1)_ii = sum_k U[i,k]^2  (U = L^-T)
//               mu = w - K (s o z),  w = K nu~,  z = B^-1 (s o w) = U (M (s o w))
//   sites:      cavity tau_ = 1/Sigma_ii - tau~_i, nu_ = mu_i/Sigma_ii - nu~_i; probit moments via erfcx;
//               tau~ += damping * (1/sigma^2_hat - 1/Sigma_ii), nu~ += damping * (mu_hat/sigma^2_hat - mu_i/Sigma_ii)
//   log Z_EP:   R&W (3.65) with (3.73)-(3.75); alpha = nu~ - s o z.
#include <math_constants.h>

#include "common.cuh"

namespace agpc {

// ---- row-oriented matrix-vector kernels: one warp per row, coalesced along the row ---------------------------
// y_i = sum_{k in range(i)} A[i,k] x_k ;  optionally y2_i = sum_{k in range(i)} A[i,k]^2
// range: 0 = all k < np, 1 = k <= i (lower), 2 = k >= i (upper); STRICT: the sum of squares leaves the diagonal out
template <int RANGE, bool SQ, bool STRICT = false>
__global__ void __launch_bounds__(256)
rowdot_kernel(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y,
              double* __restrict__ y2, int nrows, int ncols, long long a_batch, long long x_batch, long long y_batch,
              const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= nrows) return;
  const double* row = A + (long long)b * a_batch + (long long)i * ncols;
  const double* xb = x + (long long)b * x_batch;
  int k0 = 0, k1 = ncols;
  if (RANGE == 1) k1 = i + 1;
  if (RANGE == 2) k0 = i;
  // align the start down to a multiple of 2 so double2 loads are 16-byte aligned (ncols is even)
  const int ka = k0 & ~1;
  double s = 0.0, q = 0.0;
  auto edge = [&](int k) {   // one masked step: the first 64 columns (may start below k0) and the last ones (may end at k1)
    const double2 a2 = *reinterpret_cast<const double2*>(row + k);
    const double2 x2 = *reinterpret_cast<const double2*>(xb + k);
    if (k >= k0) {
      s += a2.x * x2.x;
      if (SQ && !(STRICT && k == i)) q += a2.x * a2.x;
    }
    if (k + 1 < k1) {
      s += a2.y * x2.y;
      if (SQ && !(STRICT && k + 1 == i)) q += a2.y * a2.y;
    }
  };
  int k = ka + 2 * lane;
  if (k < k1) edge(k);
  k += 64;
  // interior: four loads of the row in flight per lane, no masks.  k >= k0 after the first step; the bound keeps every
  // element inside the range AND, for a lower row (k1 = i + 1), below the diagonal, which the strict sum of squares must
  // leave out (an upper row has its diagonal in the first step).  One load per step left the kernel at 71 % of the DRAM
  // rate (profiles/r02_rowdot_kernel_ncu_raw.csv: 96 % of the stall samples on the first use of the load).
  const int k_in = k1 - 1;   // last index the interior may touch is k_in - 1
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
  for (; k + 193 < k_in; k += 256) {
    const double2 a0 = *reinterpret_cast<const double2*>(row + k), a1 = *reinterpret_cast<const double2*>(row + k + 64);
    const double2 a2 = *reinterpret_cast<const double2*>(row + k + 128), a3 = *reinterpret_cast<const double2*>(row + k + 192);
    const double2 x0 = *reinterpret_cast<const double2*>(xb + k), x1 = *reinterpret_cast<const double2*>(xb + k + 64);
    const double2 x2 = *reinterpret_cast<const double2*>(xb + k + 128), x3 = *reinterpret_cast<const double2*>(xb + k + 192);
    s = fma(a0.x, x0.x, s); s = fma(a0.y, x0.y, s);
    s1 = fma(a1.x, x1.x, s1); s1 = fma(a1.y, x1.y, s1);
    s2 = fma(a2.x, x2.x, s2); s2 = fma(a2.y, x2.y, s2);
    s3 = fma(a3.x, x3.x, s3); s3 = fma(a3.y, x3.y, s3);
    if (SQ) {
      q = fma(a0.x, a0.x, q); q = fma(a0.y, a0.y, q);
      q1 = fma(a1.x, a1.x, q1); q1 = fma(a1.y, a1.y, q1);
      q2 = fma(a2.x, a2.x, q2); q2 = fma(a2.y, a2.y, q2);
      q3 = fma(a3.x, a3.x, q3); q3 = fma(a3.y, a3.y, q3);
    }
  }
  for (; k < k1; k += 64) edge(k);
  s += (s1 + s2) + s3;
  if (SQ) q += (q1 + q2) + q3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    if (SQ) q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    y[(long long)b * y_batch + i] = s;
    if (SQ) y2[(long long)b * y_batch + i] = q;
  }
}

// range: 0 full row, 1 lower (k <= i), 2 upper (k >= i); sq: also y2_i = sum A[i,k]^2 over the same range
cudaError_t rowdot_launch(const Launch& L, int range, bool sq, const double* A, const double* x, double* y,
                          double* y2, int nrows, int ncols, long long a_batch, long long x_batch, long long y_batch,
                          int batch, const int* active, bool strict_sq) {
  ProfScope prof_scope(L, CLS_MATVEC);
  dim3 grid((nrows + 7) / 8, batch);
#define RD(R, S) rowdot_kernel<R, S><<<grid, 256, 0, L.stream>>>(A, x, y, y2, nrows, ncols, a_batch, x_batch, y_batch, active)
#define RDS(R) rowdot_kernel<R, true, true><<<grid, 256, 0, L.stream>>>(A, x, y, y2, nrows, ncols, a_batch, x_batch, y_batch, active)
  if (strict_sq && range == 1) RDS(1);
  else if (strict_sq && range == 2) RDS(2);
  else if (range == 0 && !sq) RD(0, false);
  else if (range == 0) RD(0, true);
  else if (range == 1) RD(1, false);
  else if (sq) RD(2, true);
  else RD(2, false);
#undef RD
#undef RDS
  L.bump();
  return cudaGetLastError();
}

// ---- probit helpers (same branches as oracle.log_cdf / pdf_over_cdf) ---------------------------------------------
__device__ __forceinline__ double pdf_over_cdf(double z) {
  return 0.79788456080286535588 /* sqrt(2/pi) */ / erfcx(-z * 0.70710678118654752440);
}
__device__ __forceinline__ double log_cdf(double z) {
  const double t = -z * 0.70710678118654752440;
  if (z > -5.0) return log(0.5 * erfc(t));
  return log(0.5 * erfcx(t)) - t * t;
}

// ---- small elementwise steps ---------------------------------------------------------------------------------
// s = sqrt(tau); u = s * w   (u optional)
__global__ void sqrt_scale_kernel(const double* __restrict__ tau, const double* __restrict__ w, double* __restrict__ s,
                                  double* __restrict__ u, int np, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const long long o = (long long)b * np + i;
  const double si = sqrt(tau[o]);
  if (s) s[o] = si;
  if (u) u[o] = si * w[o];
}
// out = a * b (elementwise)
__global__ void mul_kernel(const double* __restrict__ a, const double* __restrict__ bvec, double* __restrict__ out,
                           int np, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const long long o = (long long)b * np + i;
  out[o] = a[o] * bvec[o];
}
// Posterior marginals.  mu = w - kv.  Variance (kdiag when tau == 0 everywhere: cold start):
//   Sigma_ii = (1 - (B^-1)_ii) / tau_i  with  (B^-1)_ii = sum_k M_ki^2 = 1 / L_ii^2 + usq_i,  usq_i = sum_{k>i} U_ik^2.
// Evaluated literally that identity cancels against 1 and loses everything once tau_i Sigma_ii < 1e-16 -- sites at
// the tau floor of a kernel with K_ii ~ 1e6 (two of the 4096 models of configs[4] ended non-finite that way).  With
// L_ii^2 - 1 = tau_i K_ii - lsq_i,  lsq_i = sum_{k<i} L_ik^2, everything is O(tau_i) BEFORE the subtraction:
//   Sigma_ii = ((tau_i K_ii - lsq_i) / L_ii^2 - usq_i) / tau_i
// whose only cancellation is K_ii against the explained variance, as in GPy's  diag(K) - sum(V^2, 0).
__global__ void marginals_kernel(const double* __restrict__ tau, const double* __restrict__ usq,
                                 const double* __restrict__ lsq, const double* __restrict__ Lmat, long long l_stride,
                                 const double* __restrict__ w, const double* __restrict__ kv,
                                 const double* __restrict__ kdiag, double* __restrict__ sig, double* __restrict__ mu,
                                 const int* __restrict__ n_of, int np, int cold, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const long long o = (long long)b * np + i;
  if (i >= n_of[b]) {
    sig[o] = 1.0;
    mu[o] = 0.0;
    return;
  }
  if (cold) {
    sig[o] = kdiag[o];
    mu[o] = 0.0;
  } else {
    const double lii = Lmat[(long long)b * l_stride + (long long)i * np + i];
    const double t = tau[o];
    sig[o] = (fma(t, kdiag[o], -lsq[o]) / (lii * lii) - usq[o]) / t;
    mu[o] = w[o] - kv[o];
  }
}

cudaError_t sqrt_scale_launch(const Launch& L, const double* tau, const double* w, double* s, double* u, int np,
                              int batch, const int* active) {
  ProfScope prof_scope(L, CLS_EP);
  dim3 grid((np + 255) / 256, batch);
  sqrt_scale_kernel<<<grid, 256, 0, L.stream>>>(tau, w, s, u, np, active);
  L.bump();
  return cudaGetLastError();
}
cudaError_t mul_launch(const Launch& L, const double* a, const double* b, double* out, int np, int batch,
                       const int* active) {
  ProfScope prof_scope(L, CLS_EP);
  dim3 grid((np + 255) / 256, batch);
  mul_kernel<<<grid, 256, 0, L.stream>>>(a, b, out, np, active);
  L.bump();
  return cudaGetLastError();
}
cudaError_t marginals_launch(const Launch& L, const double* tau, const double* usq, const double* lsq, const double* Lmat,
                             long long l_stride, const double* w, const double* kv, const double* kdiag, double* sig,
                             double* mu, const int* n_of, int np, int batch, int cold, const int* active) {
  ProfScope prof_scope(L, CLS_EP);
  dim3 grid((np + 255) / 256, batch);
  marginals_kernel<<<grid, 256, 0, L.stream>>>(tau, usq, lsq, Lmat, l_stride, w, kv, kdiag, sig, mu, n_of, np, cold, active);
  L.bump();
  return cudaGetLastError();
}

// ---- block reduction helper (fixed order -> deterministic) -----------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* red /* [32][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if (lane == 0) red[warp * NV + q] = v[q];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double t = (lane < nw) ? red[lane * NV + q] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[q] = t;
    }
  }
  __syncthreads();
}

// ---- EP site update: one CTA per item ------------------------------------------------------------------------
// Reads marginals (sig, mu), updates (tau, nu) in place, decides convergence of the item.
// active[b] <- 0 when converged (or failed); sweeps[b]++.
//
// Posterior reuse: the fixed-point residual is measured at the CURRENT sites, i.e. at the sites whose factorisation
// (L, L^-1, marginals) is still sitting in the item's buffers.  When that residual is already EP_REUSE_FACTOR times
// below the tolerance, the sites are left untouched and converged[b] = 2 tells the caller that the posterior in the
// buffers IS the final one -- no update, no re-factorisation (2 n^3 / 3 flops of the ~2.8 n^3 of a cold evaluation,
// of the 1.7 n^3 of a warm-started one).  Otherwise the update is applied as in GPy (converged[b] = 1, the caller
// re-factorises at the new sites).  oracle/gpc_oracle.py::ep_parallel restates exactly this rule.
constexpr double EP_REUSE_FACTOR = 0.1;

struct SiteStep {
  double new_tau, new_v, r_tau, f_v;
};
__device__ __forceinline__ SiteStep site_step(double sg, double m, double tt, double vt, double y, double kd,
                                              double damping) {
  const double tau_cav = 1.0 / sg - tt;
  const double v_cav = m / sg - vt;
  const double den2 = tau_cav * tau_cav + tau_cav;
  const double den = sqrt(den2);
  const double z = y * v_cav / den;
  const double N = pdf_over_cdf(z);
  const double mu_hat = v_cav / tau_cav + y * N / den;
  const double s2_hat = 1.0 / tau_cav - N * (z + N) / den2;
  const double f_tau = 1.0 / s2_hat - 1.0 / sg;  // full step = fixed-point residual
  const double f_v = mu_hat / s2_hat - m / sg;
  const double floor_tau = 1e-10 / kd;  // keeps (1 - Binv_ii)/tau well defined; see DESIGN.md
  SiteStep o;
  o.new_tau = tt + damping * f_tau;
  if (!(o.new_tau >= floor_tau)) o.new_tau = isnan(o.new_tau) ? o.new_tau : floor_tau;
  double full_tau = tt + f_tau;
  if (!(full_tau >= floor_tau)) full_tau = isnan(full_tau) ? full_tau : floor_tau;
  o.r_tau = full_tau - tt;  // convergence is judged on the undamped residual
  o.new_v = vt + damping * f_v;
  o.f_v = f_v;
  return o;
}

__global__ void __launch_bounds__(1024)
ep_update_kernel(double* __restrict__ tau, double* __restrict__ nu, const double* __restrict__ sig,
                 const double* __restrict__ mu, const double* __restrict__ yf, const double* __restrict__ kdiag,
                 const int* __restrict__ n_of, int np, double eps, double damping, int min_sweeps, int* __restrict__ active,
                 int* __restrict__ sweeps, int* __restrict__ status, int* __restrict__ converged,
                 const int* __restrict__ info) {
  const int b = blockIdx.x;
  if (active[b] == 0) return;
  if (info[b] != 0) {  // the factorisation behind these marginals failed
    if (threadIdx.x == 0) {
      status[b] = info[b];
      active[b] = 0;
    }
    return;
  }
  __shared__ double red[32 * 3];
  __shared__ int keep_sites;
  const int n = n_of[b];
  // damping <= 0: automatic schedule, undamped for the first 3 sweeps and 0.8 afterwards (oracle.ep_damping): undamped
  // parallel EP limit-cycles on strongly correlated priors where GPy's sequential EP converges
  if (damping <= 0.0) damping = (sweeps[b] + 1 <= 3) ? 1.0 : 0.8;
  double acc[3] = {0.0, 0.0, 0.0};  // sum r_tau^2, sum r_nu^2 (undamped residuals), non-finite count
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long o = (long long)b * np + i;
    const SiteStep st = site_step(sig[o], mu[o], tau[o], nu[o], yf[o], kdiag[o], damping);
    acc[0] += st.r_tau * st.r_tau;
    acc[1] += st.f_v * st.f_v;
    if (!isfinite(st.new_tau) || !isfinite(st.new_v)) acc[2] += 1.0;
  }
  block_reduce<3>(acc, red);
  if (threadIdx.x == 0) {
    const int sw = sweeps[b] + 1;
    sweeps[b] = sw;
    int keep = 0;
    if (acc[2] > 0.0) {
      status[b] = AGPC_ST_NONFINITE;
      active[b] = 0;
    } else if (sw >= min_sweeps && acc[0] / n <= eps && acc[1] / n <= eps) {
      keep = (acc[0] / n <= EP_REUSE_FACTOR * eps && acc[1] / n <= EP_REUSE_FACTOR * eps) ? 1 : 0;
      converged[b] = keep ? 2 : 1;
      active[b] = 0;
    }
    keep_sites = keep;
  }
  __syncthreads();
  if (keep_sites) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long o = (long long)b * np + i;
    const SiteStep st = site_step(sig[o], mu[o], tau[o], nu[o], yf[o], kdiag[o], damping);
    tau[o] = st.new_tau;
    nu[o] = st.new_v;
  }
}

cudaError_t ep_update_launch(const Launch& L, double* tau, double* nu, const double* sig, const double* mu,
                             const double* yf, const double* kdiag, const int* n_of, int np, int batch, double eps,
                             double damping, int min_sweeps, int* active, int* sweeps, int* status, int* converged,
                             const int* info) {
  ProfScope prof_scope(L, CLS_EP);
  ep_update_kernel<<<batch, 1024, 0, L.stream>>>(tau, nu, sig, mu, yf, kdiag, n_of, np, eps, damping, min_sweeps, active, sweeps,
                                                  status, converged, info);
  L.bump();
  return cudaGetLastError();
}

// ---- final reductions: log Z_EP, alpha, Gaussian-only value, BIC ----------------------------------------------
__global__ void __launch_bounds__(1024)
ep_final_kernel(const double* __restrict__ tau, const double* __restrict__ nu, const double* __restrict__ sig,
                const double* __restrict__ mu, const double* __restrict__ yf, const double* __restrict__ s,
                const double* __restrict__ z /* B^-1 (s o w) */, const double* __restrict__ Lmat, long long stride,
                const int* __restrict__ n_of, int np, const agpc_program* __restrict__ progs,
                double* __restrict__ alpha, double* __restrict__ nlml, double* __restrict__ nlml_gauss,
                double* __restrict__ bic, int* __restrict__ status, const int* __restrict__ info) {
  const int b = blockIdx.x;
  __shared__ double red[32 * 4];
  const int n = n_of[b];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // site terms + 1/2 nu~.mu ; sum log L_ii ; gauss extras ; nonfinite
  const double* Lb = Lmat + (long long)b * stride;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    const long long o = (long long)b * np + i;
    if (i >= n) {
      alpha[o] = 0.0;
      continue;
    }
    const double sg = sig[o], m = mu[o], tt = tau[o], vt = nu[o], y = yf[o];
    const double tau_cav = 1.0 / sg - tt;
    const double v_cav = m / sg - vt;
    const double zz = y * v_cav / sqrt(tau_cav * tau_cav + tau_cav);
    const double tsum = tau_cav + tt;
    const double site = log_cdf(zz) + 0.5 * log1p(tt / tau_cav) - 0.5 * vt * vt / tsum +
                        0.5 * v_cav * ((tt / tau_cav) * v_cav - 2.0 * vt) / tsum;
    const double a = vt - s[o] * z[o];
    alpha[o] = a;
    const double lii = Lb[(long long)i * np + i];
    acc[0] += site + 0.5 * vt * m;
    acc[1] += log(lii);
    if (tt > 0.0) acc[2] += -log(tt) + (vt / tt) * a;
    if (!isfinite(site) || !isfinite(a) || !(lii > 0.0)) acc[3] += 1.0;
  }
  block_reduce<4>(acc, red);
  if (threadIdx.x == 0) {
    const double log_z = -acc[1] + acc[0];
    const double v = -log_z;
    nlml[b] = v;
    nlml_gauss[b] = 0.5 * (n * 1.8378770664093454836 /* ln 2 pi */ + 2.0 * acc[1] + acc[2]);
    bic[b] = 2.0 * v + progs[b].p_eff * log((double)n);
    if (info[b] != 0) status[b] = info[b];
    else if ((acc[3] > 0.0 || !isfinite(v)) && status[b] != AGPC_ST_NOT_PD) status[b] = AGPC_ST_NONFINITE;
    if (status[b] == AGPC_ST_NOT_PD || status[b] == AGPC_ST_NONFINITE) {
      nlml[b] = CUDART_INF;
      nlml_gauss[b] = CUDART_INF;
      bic[b] = CUDART_INF;
    }
  }
}

cudaError_t ep_final_launch(const Launch& L, const double* tau, const double* nu, const double* sig, const double* mu,
                            const double* yf, const double* s, const double* z, const double* Lmat, long long stride,
                            const int* n_of, int np, int batch, const agpc_program* progs, double* alpha, double* nlml,
                            double* nlml_gauss, double* bic, int* status, const int* info) {
  ProfScope prof_scope(L, CLS_EP);
  ep_final_kernel<<<batch, 1024, 0, L.stream>>>(tau, nu, sig, mu, yf, s, z, Lmat, stride, n_of, np, progs, alpha, nlml,
                                                 nlml_gauss, bic, status, info);
  L.bump();
  return cudaGetLastError();
}

// ---- Laplace approximation ([R&W] alg. 3.1 / 5.1; oracle.laplace_mode, score_laplace) ------------------------------
// log p(y|f) = log Phi(y f):  g1 = y h,  W = h (z + h),  d3 = y h ((z + h)(z + 2h) - 1)  with z = y f, h = phi/Phi(z);
// also s = sqrt(W) and b = W f + g1 (the Newton right-hand side; with tau = W, nu = b the EP posterior formulas give
// the Laplace posterior, which is how predict_* and the outputs reuse the site form).
__global__ void laplace_lik_kernel(const double* __restrict__ f, const double* __restrict__ yf, double* __restrict__ W,
                                   double* __restrict__ s, double* __restrict__ bv, double* __restrict__ g1,
                                   double* __restrict__ d3, const int* __restrict__ n_of, int np,
                                   const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const long long o = (long long)b * np + i;
  if (i >= n_of[b]) {
    W[o] = s[o] = bv[o] = g1[o] = d3[o] = 0.0;
    return;
  }
  const double y = yf[o], fi = f[o];
  const double z = y * fi, h = pdf_over_cdf(z);
  const double w = h * (z + h);
  W[o] = w;
  s[o] = sqrt(w);
  g1[o] = y * h;
  d3[o] = y * h * ((z + h) * (z + 2.0 * h) - 1.0);
  bv[o] = w * fi + y * h;
}
cudaError_t laplace_lik_launch(const Launch& L, const double* f, const double* yf, double* W, double* s, double* b,
                               double* g1, double* d3, const int* n_of, int np, int batch, const int* active) {
  ProfScope prof_scope(L, CLS_EP);
  dim3 grid((np + 255) / 256, batch);
  laplace_lik_kernel<<<grid, 256, 0, L.stream>>>(f, yf, W, s, b, g1, d3, n_of, np, active);
  L.bump();
  return cudaGetLastError();
}

// Newton step with halving on psi(f) = -1/2 a^T f + sum log Phi(y f): one CTA per item.
__global__ void __launch_bounds__(1024)
laplace_step_kernel(double* __restrict__ f, double* __restrict__ a, const double* __restrict__ f_new,
                    const double* __restrict__ a_new, const double* __restrict__ yf, double* __restrict__ psi,
                    const int* __restrict__ n_of, int np, double tol, int* __restrict__ active,
                    int* __restrict__ iters, int* __restrict__ status, const int* __restrict__ info) {
  const int b = blockIdx.x;
  if (active[b] == 0) return;
  if (info[b] != 0) {
    if (threadIdx.x == 0) {
      status[b] = info[b];
      active[b] = 0;
    }
    return;
  }
  __shared__ double red[32 * 2];
  __shared__ double s_step;
  __shared__ int s_ok;
  const int n = n_of[b];
  const long long o0 = (long long)b * np;
  const double psi_old = (iters[b] == 0) ? n * -0.69314718055994530942 : psi[b];  // cold start: f = a = 0
  double step = 1.0, psi_t = 0.0;
  bool ok = false;
  for (int trial = 0; trial <= 10; ++trial) {
    double acc[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double at = a[o0 + i] + step * (a_new[o0 + i] - a[o0 + i]);
      const double ft = f[o0 + i] + step * (f_new[o0 + i] - f[o0 + i]);
      acc[0] += at * ft;
      acc[1] += log_cdf(yf[o0 + i] * ft);
    }
    block_reduce<2>(acc, red);
    if (threadIdx.x == 0) {
      const double p = -0.5 * acc[0] + acc[1];
      s_step = p;
      s_ok = (p >= psi_old - 1e-12 * fabs(psi_old)) ? 1 : 0;
    }
    __syncthreads();
    psi_t = s_step;
    ok = s_ok != 0;
    __syncthreads();
    if (ok || trial == 10) break;
    step *= 0.5;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    a[o0 + i] += step * (a_new[o0 + i] - a[o0 + i]);
    f[o0 + i] += step * (f_new[o0 + i] - f[o0 + i]);
  }
  if (threadIdx.x == 0) {
    iters[b] += 1;
    psi[b] = psi_t;
    if (!isfinite(psi_t)) {
      status[b] = AGPC_ST_NONFINITE;
      active[b] = 0;
    } else if (ok && fabs(psi_t - psi_old) <= tol * fmax(1.0, fabs(psi_t))) {
      active[b] = 0;
    }
  }
}
cudaError_t laplace_step_launch(const Launch& L, double* f, double* a, const double* f_new, const double* a_new,
                                const double* yf, double* psi, const int* n_of, int np, int batch, double tol,
                                int* active, int* iters, int* status, const int* info) {
  ProfScope prof_scope(L, CLS_EP);
  laplace_step_kernel<<<batch, 1024, 0, L.stream>>>(f, a, f_new, a_new, yf, psi, n_of, np, tol, active, iters, status, info);
  L.bump();
  return cudaGetLastError();
}

// log q = -1/2 a^T f + sum log Phi(y f) - sum log L_ii;  s2_i = 1/2 Sigma_ii d3_i with Sigma_ii = (1 - B^-1_ii) / W_i
__global__ void __launch_bounds__(1024)
laplace_final_kernel(const double* __restrict__ f, const double* __restrict__ a, const double* __restrict__ W,
                     const double* __restrict__ g1, const double* __restrict__ d3, const double* __restrict__ dB,
                     const double* __restrict__ yf, const double* __restrict__ Lmat, long long stride,
                     const int* __restrict__ n_of, int np, const agpc_program* __restrict__ progs,
                     double* __restrict__ s2, double* __restrict__ nlml, double* __restrict__ nlml_gauss,
                     double* __restrict__ bic, int* __restrict__ status, const int* __restrict__ info) {
  const int b = blockIdx.x;
  __shared__ double red[32 * 3];
  const int n = n_of[b];
  const long long o0 = (long long)b * np;
  const double* Lb = Lmat + (long long)b * stride;
  double acc[3] = {0.0, 0.0, 0.0};  // psi terms, sum log L_ii, non-finite count
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    if (i >= n) {
      s2[o0 + i] = 0.0;
      continue;
    }
    const double lii = Lb[(long long)i * np + i];
    const double sig = (1.0 - dB[o0 + i]) / W[o0 + i];
    const double v = 0.5 * sig * d3[o0 + i];
    s2[o0 + i] = v;
    acc[0] += -0.5 * a[o0 + i] * f[o0 + i] + log_cdf(yf[o0 + i] * f[o0 + i]);
    acc[1] += log(lii);
    if (!isfinite(v) || !(lii > 0.0)) acc[2] += 1.0;
  }
  block_reduce<3>(acc, red);
  if (threadIdx.x == 0) {
    const double v = -(acc[0] - acc[1]);
    nlml[b] = v;
    nlml_gauss[b] = v;
    bic[b] = 2.0 * v + progs[b].p_eff * log((double)n);
    if (info[b] != 0) status[b] = info[b];
    else if ((acc[2] > 0.0 || !isfinite(v)) && status[b] != AGPC_ST_NOT_PD) status[b] = AGPC_ST_NONFINITE;
    if (status[b] == AGPC_ST_NOT_PD || status[b] == AGPC_ST_NONFINITE) {
      nlml[b] = CUDART_INF;
      nlml_gauss[b] = CUDART_INF;
      bic[b] = CUDART_INF;
    }
  }
}
cudaError_t laplace_final_launch(const Launch& L, const double* f, const double* a, const double* W, const double* g1,
                                 const double* d3, const double* dB, const double* yf, const double* Lmat,
                                 long long stride, const int* n_of, int np, int batch, const agpc_program* progs,
                                 double* s2, double* nlml, double* nlml_gauss, double* bic, int* status, const int* info) {
  ProfScope prof_scope(L, CLS_EP);
  laplace_final_kernel<<<batch, 1024, 0, L.stream>>>(f, a, W, g1, d3, dB, yf, Lmat, stride, n_of, np, progs, s2, nlml,
                                                      nlml_gauss, bic, status, info);
  L.bump();
  return cudaGetLastError();
}

// out = x + alpha * y
__global__ void axpy_kernel(const double* __restrict__ x, const double* __restrict__ y, double alpha,
                            double* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = x[i] + alpha * y[i];
}
cudaError_t axpy_launch(const Launch& L, const double* x, const double* y, double alpha, double* out, int np, int batch) {
  ProfScope prof_scope(L, CLS_EP);
  const long long total = (long long)np * batch;
  axpy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(x, y, alpha, out, total);
  L.bump();
  return cudaGetLastError();
}

// ---- predictive probability p* = Phi(mu* / sqrt(1 + v*)) -------------------------------------------------------
__global__ void predict_prob_kernel(const double* __restrict__ mu_s, const double* __restrict__ kss,
                                    const double* __restrict__ vsq, double* __restrict__ p, double* __restrict__ var,
                                    long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double v = kss[i] - vsq[i];
  if (var) var[i] = v;
  p[i] = 0.5 * erfc(-(mu_s[i] / sqrt(1.0 + v)) * 0.70710678118654752440);
}

cudaError_t predict_prob_launch(const Launch& L, const double* mu_s, const double* kss, const double* vsq, double* p,
                                double* var, long long total) {
  ProfScope prof_scope(L, CLS_MISC);
  predict_prob_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(mu_s, kss, vsq, p, var, total);
  L.bump();
  return cudaGetLastError();
}

}  // namespace agpc
