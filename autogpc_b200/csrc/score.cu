// score_batch / predict_batch: the lock-step batched scoring pipeline behind GPy.models.GPClassification(X, Y, kernel)
// + parameters_changed() (gpckernel.py:126,245-246,322), m.log_likelihood() (gpckernel.py:371) and m.predict
// (gpckernel.py:832).  Host orchestration only -- every numerical step is a kernel in kbuild.cu / ep.cu / chol.cu /
// gemm_nt.cu.  Items of a chunk advance through EP sweeps together; a per-item `active` flag on the device lets
// converged items drop out of every launch, and the host polls that flag once per sweep.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <chrono>
#include <cstdlib>

#include "engine.h"

using namespace agpc;

namespace {

int fail(agpc_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess)
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  else
    snprintf(buf, sizeof buf, "%s", what);
  ctx->err = buf;
  return code;
}

#define CK(expr)                                                      \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) return fail(ctx, AGPC_E_CUDA, #expr, _e);  \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 1024);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += sizeof(T) * count;
    return p;
  }
};

// All device buffers of one lock-step chunk.
struct Chunk {
  int nb = 0, np = 0, D = 0, theta_stride = 0;
  MatSet ms;
  double *tau, *nu, *s, *sig, *mu, *w, *u, *t, *z, *kv, *dB, *alpha, *kdiag, *yf, *tmp, *lsq, *lsq_dummy;
  double* Xf;
  int *n_of, *active, *sweeps, *status, *converged, *info, *ones;
  int *jit_tries, *jit_retry, *jit_used;  // jitchol ladder state per item
  double* jitter;
  long long* row_off;
  int* rows;
  agpc_program* progs;
  double* theta;
  double *nlml, *nlml_gauss, *bic, *grad, *partial;

  size_t carve(char* base, int nb_, int np_, int D_, int theta_stride_, bool with_rows, int oz_S) {
    nb = nb_; np = np_; D = D_; theta_stride = theta_stride_;
    Carver c{base};
    const size_t mat = (size_t)nb * np * np, vec = (size_t)nb * np;
    ms.np = np; ms.batch = nb; ms.stride = (long long)np * np;
    ms.K = c.take<double>(mat); ms.A = c.take<double>(mat); ms.M = c.take<double>(mat);
    ms.U = c.take<double>(mat); ms.T = c.take<double>(mat);
    ms.oz_S = oz_S;
    if (oz_S > 0) {  // digit-slice images of the two operands of the int8 tensor-core GEMM + their row scales
      ms.SA = c.take<int8_t>(mat * oz_S); ms.SB = c.take<int8_t>(mat * oz_S); ms.SC = c.take<int8_t>(mat * oz_S);
      ms.scA = c.take<double>(vec); ms.scB = c.take<double>(vec); ms.scC = c.take<double>(vec); ms.rmax = c.take<double>(vec);
      ms.rmM = c.take<double>(vec); ms.rmU = c.take<double>(vec);
    }
    tau = c.take<double>(vec); nu = c.take<double>(vec); s = c.take<double>(vec); sig = c.take<double>(vec);
    mu = c.take<double>(vec); w = c.take<double>(vec); u = c.take<double>(vec); t = c.take<double>(vec);
    z = c.take<double>(vec); kv = c.take<double>(vec); dB = c.take<double>(vec); alpha = c.take<double>(vec);
    kdiag = c.take<double>(vec); yf = c.take<double>(vec); tmp = c.take<double>(vec);
    lsq = c.take<double>(vec); lsq_dummy = c.take<double>(vec);
    Xf = c.take<double>(vec * D);
    n_of = c.take<int>(nb); active = c.take<int>(nb); sweeps = c.take<int>(nb); status = c.take<int>(nb);
    converged = c.take<int>(nb); info = c.take<int>(nb); ones = c.take<int>(nb);
    jit_tries = c.take<int>(nb); jit_retry = c.take<int>(nb); jit_used = c.take<int>(nb); jitter = c.take<double>(nb);
    row_off = c.take<long long>(nb + 1);
    rows = with_rows ? c.take<int>(vec) : nullptr;
    progs = c.take<agpc_program>(nb);
    theta = c.take<double>((size_t)nb * theta_stride);
    nlml = c.take<double>(nb); nlml_gauss = c.take<double>(nb); bic = c.take<double>(nb);
    grad = c.take<double>((size_t)nb * theta_stride);
    partial = c.take<double>((size_t)nb * grad_tiles(np, 0) * AGPC_MAX_THETA);
    return align_up(c.off, 1024);
  }
  int make_maps() {
    return make_tensor_map(&ms.mapA, ms.A, np, np, nb, np, ms.stride) | make_tensor_map(&ms.mapM, ms.M, np, np, nb, np, ms.stride) |
           make_tensor_map(&ms.mapU, ms.U, np, np, nb, np, ms.stride) | make_tensor_map(&ms.mapT, ms.T, np, np, nb, np, ms.stride) |
           make_tensor_map(&ms.mapA32, ms.A, np, np, nb, np, ms.stride, 32) | make_tensor_map(&ms.mapM32, ms.M, np, np, nb, np, ms.stride, 32);
  }
};

__global__ void fill_int_kernel(int* p, int v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// dst[b*dst_stride + i] = src[b*src_stride + i] for i < n_of[b]; dst zero-filled up to dst_len when pad != 0
__global__ void copy_strided_kernel(double* __restrict__ dst, long long dst_stride, const double* __restrict__ src,
                                    long long src_stride, const int* __restrict__ n_of, int dst_len, int pad) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dst_len) return;
  if (i < n_of[b])
    dst[(long long)b * dst_stride + i] = src[(long long)b * src_stride + i];
  else if (pad)
    dst[(long long)b * dst_stride + i] = 0.0;
}
// warm-started sites must satisfy tau >= 1e-10 / K_ii (same floor as the EP update)
__global__ void clamp_tau_kernel(double* __restrict__ tau, const double* __restrict__ kdiag, const int* __restrict__ n_of,
                                 int np) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_of[b]) return;
  const long long o = (long long)b * np + i;
  const double fl = 1e-10 / kdiag[o];
  if (!(tau[o] >= fl)) tau[o] = fl;
}
// items that are still active after the last sweep did not converge
__global__ void mark_unconverged_kernel(const int* __restrict__ active, int* __restrict__ status, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && active[b] != 0 && status[b] == AGPC_ST_OK) status[b] = AGPC_ST_EP_NOT_CONVERGED;
}
// a result whose factorisation needed jitter is flagged (GPy would only log a warning; SURVEY 8b item 4 status 1)
__global__ void jitter_status_kernel(const int* __restrict__ used, int* __restrict__ status, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && used[b] != 0 && status[b] == AGPC_ST_OK) status[b] = AGPC_ST_JITTER;
}
// need[b] = 1 unless the EP update left the item's sites untouched (converged[b] == 2, ep.cu): then the
// factorisation, inverse and marginals in its buffers already belong to the final sites
__global__ void need_final_kernel(const int* __restrict__ converged, int* __restrict__ need, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) need[b] = converged[b] == 2 ? 0 : 1;
}

int count_flags(agpc_ctx* ctx, const Launch& L, const int* flags_dev, int nb, int value, int* n_out);

// jitchol (GPy util/linalg.py): a factorisation that fails is retried on B + jitter I with jitter = mean(diag B) 1e-6,
// x 10 per further failure, at most 5 times; every factorisation starts its own ladder.  One poll of info[] per
// factorisation (a 4 nb byte copy) finds the failures; the common case costs nothing else.
// fault injection for the tests (a B = I + S^1/2 K S^1/2 built from PSD kernels cannot be made to fail reliably):
// AGPC_FAULT_NOT_PD=k makes the plain attempt and the first k - 1 jittered attempts of item 0 of every factorisation
// report NOT_PD
int fault_budget() {
  static const int fault = getenv("AGPC_FAULT_NOT_PD") ? atoi(getenv("AGPC_FAULT_NOT_PD")) : 0;
  return fault;
}
cudaError_t inject_fault(const Launch& L, Chunk& c, int* fault_left) {
  static const int not_pd = AGPC_ST_NOT_PD;
  if ((*fault_left)-- > 0) return cudaMemcpyAsync(c.info, &not_pd, sizeof(int), cudaMemcpyHostToDevice, L.stream);
  return cudaSuccess;
}

int jitter_ladder(agpc_ctx* ctx, const Launch& L, Chunk& c, const int* active, int* fault_left) {
  const int np = c.np, nb = c.nb;
  int n_bad = 0;
  int rc = count_flags(ctx, L, c.info, nb, AGPC_ST_NOT_PD, &n_bad);
  if (rc) return rc;
  if (n_bad == 0) return AGPC_OK;
  if (cudaMemsetAsync(c.jit_tries, 0, sizeof(int) * nb, L.stream) != cudaSuccess) return AGPC_E_CUDA;
  for (int attempt = 0; attempt < 5 && n_bad > 0; ++attempt) {
    if (jitter_step_launch(L, c.kdiag, c.s, c.n_of, np, nb, c.info, c.jitter, c.jit_tries, c.jit_retry, c.jit_used, active) != cudaSuccess)
      return AGPC_E_CUDA;
    int n_retry = 0;
    rc = count_flags(ctx, L, c.jit_retry, nb, 1, &n_retry);
    if (rc) return rc;
    if (n_retry == 0) break;
    Launch Lr = L;
    Lr.work_items = n_retry;
    if (form_B_launch(Lr, c.ms.K, c.s, c.ms.A, np, c.ms.stride, nb, c.jit_retry, c.jitter) != cudaSuccess) return AGPC_E_CUDA;
    if (potrf_blocked(Lr, c.ms, c.jit_retry, c.info) != cudaSuccess) return AGPC_E_CUDA;
    if (inject_fault(L, c, fault_left) != cudaSuccess) return AGPC_E_CUDA;
    rc = count_flags(ctx, L, c.info, nb, AGPC_ST_NOT_PD, &n_bad);
    if (rc) return rc;
  }
  return AGPC_OK;
}

// posterior at the current sites for the items flagged in `active` (nullptr = all):
// s, A = chol(B), M, U, w = K nu, z = B^-1 (s o w), dB = diag(B^-1), kv = K (s o z), then sig / mu.
cudaError_t posterior(agpc_ctx* ctx, const Launch& L, Chunk& c, const int* active) {
  const int np = c.np, nb = c.nb;
  const long long st = c.ms.stride;
  {
    const double wn = L.work_n > 0.0 ? L.work_n : (double)np, n2 = 8.0 * wn * wn * L.work_items;
    L.add_work(CLS_FORMB, n2);          // read lower K, write lower B
    L.add_work(CLS_MATVEC, 3.5 * n2);   // two full K passes, three triangular passes over M, U and L
  }
  AGPC_CUDA_TRY(sqrt_scale_launch(L, c.tau, nullptr, c.s, nullptr, np, nb, active));
  AGPC_CUDA_TRY(form_B_launch(L, c.ms.K, c.s, c.ms.A, np, st, nb, active));
  AGPC_CUDA_TRY(potrf_blocked(L, c.ms, active, c.info));
  int fault_left = fault_budget();
  AGPC_CUDA_TRY(inject_fault(L, c, &fault_left));
  if (ctx->jitter && jitter_ladder(ctx, L, c, active, &fault_left) != AGPC_OK) return cudaErrorUnknown;
  AGPC_CUDA_TRY(trtri_blocked(L, c.ms, active));
  AGPC_CUDA_TRY(rowdot_launch(L, 0, false, c.ms.K, c.nu, c.w, nullptr, np, np, st, np, np, nb, active));
  AGPC_CUDA_TRY(mul_launch(L, c.s, c.w, c.u, np, nb, active));
  AGPC_CUDA_TRY(rowdot_launch(L, 1, false, c.ms.M, c.u, c.t, nullptr, np, np, st, np, np, nb, active));
  // z = U t and usq_i = sum_{k>i} U_ik^2; lsq_i = sum_{k<i} L_ik^2 (one more triangular pass, over L): see marginals_kernel
  AGPC_CUDA_TRY(rowdot_launch(L, 2, true, c.ms.U, c.t, c.z, c.dB, np, np, st, np, np, nb, active, true));
  AGPC_CUDA_TRY(rowdot_launch(L, 1, true, c.ms.A, c.t, c.lsq_dummy, c.lsq, np, np, st, np, np, nb, active, true));
  AGPC_CUDA_TRY(mul_launch(L, c.s, c.z, c.tmp, np, nb, active));
  AGPC_CUDA_TRY(rowdot_launch(L, 0, false, c.ms.K, c.tmp, c.kv, nullptr, np, np, st, np, np, nb, active));
  AGPC_CUDA_TRY(marginals_launch(L, c.tau, c.dB, c.lsq, c.ms.A, st, c.w, c.kv, c.kdiag, c.sig, c.mu, c.n_of, np, nb, 0, active));
  return cudaSuccess;
}

size_t per_item_bytes(int np, int D, int theta_stride, int oz_S) {
  Chunk tmp;
  const size_t one = tmp.carve(nullptr, 1, np, D, theta_stride, true, oz_S);
  const size_t two = tmp.carve(nullptr, 2, np, D, theta_stride, true, oz_S);
  return two - one + 4096;
}

int poll_active(agpc_ctx* ctx, const Launch& L, const int* active_dev, int nb, int* n_active) {
  if (ctx->pinned_bytes < sizeof(int) * (size_t)nb) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    CK(cudaMallocHost(&ctx->pinned, sizeof(int) * (size_t)nb * 2));
    ctx->pinned_bytes = sizeof(int) * (size_t)nb * 2;
  }
  CK(cudaMemcpyAsync(ctx->pinned, active_dev, sizeof(int) * nb, cudaMemcpyDeviceToHost, L.stream));
  CK(cudaStreamSynchronize(L.stream));
  int cnt = 0;
  const int* a = static_cast<const int*>(ctx->pinned);
  for (int i = 0; i < nb; ++i) cnt += a[i] != 0;
  *n_active = cnt;
  return AGPC_OK;
}

// number of entries of a device flag array equal to `value` (one small D2H copy + stream synchronisation)
int count_flags(agpc_ctx* ctx, const Launch& L, const int* flags_dev, int nb, int value, int* n_out) {
  if (ctx->pinned_bytes < sizeof(int) * (size_t)nb) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    CK(cudaMallocHost(&ctx->pinned, sizeof(int) * (size_t)nb * 2));
    ctx->pinned_bytes = sizeof(int) * (size_t)nb * 2;
  }
  CK(cudaMemcpyAsync(ctx->pinned, flags_dev, sizeof(int) * nb, cudaMemcpyDeviceToHost, L.stream));
  CK(cudaStreamSynchronize(L.stream));
  int cnt = 0;
  const int* a = static_cast<const int*>(ctx->pinned);
  for (int i = 0; i < nb; ++i) cnt += a[i] == value;
  *n_out = cnt;
  return AGPC_OK;
}

// EP for one lock-step chunk (K is already built): parallel-schedule sweeps until every item has converged, then the
// posterior at the final sites, log Z_EP, alpha and (optionally) the gradient.  tau_src / nu_src: warm-start sites.
int ep_chunk(agpc_ctx* ctx, Launch& L, Chunk& c, const agpc_ep_options* opt, int theta_stride, bool want_grad, int D,
             bool warm, bool frozen, const double* tau_src, const double* nu_src, int site_stride) {
  const int np = c.np, nb = c.nb;
  const int tb = 256;
  dim3 gvec((np + tb - 1) / tb, nb);
  CK(cudaMemsetAsync(c.jit_used, 0, sizeof(int) * nb, L.stream));
  if (warm) {
    copy_strided_kernel<<<gvec, tb, 0, L.stream>>>(c.tau, np, tau_src, site_stride, c.n_of, np, 1);
    copy_strided_kernel<<<gvec, tb, 0, L.stream>>>(c.nu, np, nu_src, site_stride, c.n_of, np, 1);
    clamp_tau_kernel<<<gvec, tb, 0, L.stream>>>(c.tau, c.kdiag, c.n_of, np);
    L.bump(3);
  } else {
    CK(cudaMemsetAsync(c.tau, 0, sizeof(double) * (size_t)nb * np, L.stream));
    CK(cudaMemsetAsync(c.nu, 0, sizeof(double) * (size_t)nb * np, L.stream));
  }

  // ---- EP sweeps (parallel schedule) ----
  for (int sweep = 1; sweep <= opt->max_sweeps; ++sweep) {
    if (sweep == 1 && !warm) {
      CK(marginals_launch(L, c.tau, c.dB, c.lsq, c.ms.A, c.ms.stride, c.w, c.kv, c.kdiag, c.sig, c.mu, c.n_of, np, nb, 1, c.active));
    } else {
      CK(posterior(ctx, L, c, c.active));
    }
    // GPy measures convergence as the change between two sweeps (hence >= 2 from a cold start); the residual used here
    // is an absolute fixed-point error, so warm-started sites that are already converged may stop after one sweep
    CK(ep_update_launch(L, c.tau, c.nu, c.sig, c.mu, c.yf, c.kdiag, c.n_of, np, nb, opt->epsilon, opt->damping,
                        warm ? 1 : 2, c.active, c.sweeps, c.status, c.converged, c.info));
    int n_active = 0;
    int rc = poll_active(ctx, L, c.active, nb, &n_active);
    if (rc) return rc;
    if (n_active == 0) break;
    L.work_items = n_active;
  }
  L.work_items = nb;
  if (!frozen) {
    mark_unconverged_kernel<<<(nb + 255) / 256, 256, 0, L.stream>>>(c.active, c.status, nb);
    L.bump();
  }

  // ---- final posterior at the final sites (only for the items whose sites moved after their last factorisation:
  //      frozen or unconverged items and those that converged with a residual above the reuse threshold), log Z, gradient
  CK(cudaMemsetAsync(c.info, 0, sizeof(int) * nb, L.stream));
  need_final_kernel<<<(nb + 255) / 256, 256, 0, L.stream>>>(c.converged, c.active, nb);
  L.bump();
  int n_need = 0;
  {
    int rc = poll_active(ctx, L, c.active, nb, &n_need);
    if (rc) return rc;
  }
  if (n_need > 0) {
    L.work_items = n_need;
    CK(posterior(ctx, L, c, c.active));
    L.work_items = nb;
  }
  CK(ep_final_launch(L, c.tau, c.nu, c.sig, c.mu, c.yf, c.s, c.z, c.ms.A, c.ms.stride, c.n_of, np, nb, c.progs,
                     c.alpha, c.nlml, c.nlml_gauss, c.bic, c.status, c.info));
  jitter_status_kernel<<<(nb + 255) / 256, 256, 0, L.stream>>>(c.jit_used, c.status, nb);
  L.bump();
  if (want_grad) {
    CK(lauum_blocked(L, c.ms, c.ms.T, nullptr));
    GradArgs ga{};
    ga.progs = c.progs; ga.theta = c.theta; ga.theta_stride = theta_stride; ga.Xf = c.Xf; ga.n_of = c.n_of;
    ga.np = np; ga.D = D; ga.G = c.ms.T; ga.ld = np; ga.g_batch = c.ms.stride; ga.alpha = c.alpha; ga.s = c.s;
    ga.mode = 0; ga.partial = c.partial; ga.tiles_per_item = grad_tiles(np, 0); ga.active = nullptr;
    // d nlml / d theta = - sum dL_dK * dK/dtheta
    L.add_work(CLS_GRAD, 4.0 * L.work_n * L.work_n * nb);  // lower triangle of B^-1 read once
    CK(cudaMemsetAsync(c.grad, 0, sizeof(double) * (size_t)nb * theta_stride, L.stream));
    CK(grad_contract_launch(L, ga, nb, -1.0, c.grad, theta_stride));
  }
  return AGPC_OK;
}

// Laplace approximation for one lock-step chunk ([R&W] alg. 3.1 with step halving, then alg. 5.1); K is already built.
// Vector roles: mu = f, alpha = a = K^-1 f, tau = W, s = sqrt(W), nu = b = W f + grad, w = grad log p, sig = d3.
int laplace_chunk(agpc_ctx* ctx, Launch& L, Chunk& c, const agpc_ep_options* opt, int theta_stride, bool want_grad, int D) {
  const int np = c.np, nb = c.nb;
  const long long st = c.ms.stride;
  const size_t vec = sizeof(double) * (size_t)nb * np;
  CK(cudaMemsetAsync(c.mu, 0, vec, L.stream));
  CK(cudaMemsetAsync(c.alpha, 0, vec, L.stream));
  CK(cudaMemsetAsync(c.nlml_gauss, 0, sizeof(double) * nb, L.stream));  // psi per item lives here during the iteration
  const int max_iter = opt->max_sweeps > 0 ? opt->max_sweeps : 50;
  const double tol = opt->epsilon > 0.0 ? opt->epsilon : 1e-10;
  const double n2 = 8.0 * L.work_n * L.work_n;
  for (int it = 1; it <= max_iter; ++it) {
    L.add_work(CLS_FORMB, n2 * L.work_items);
    L.add_work(CLS_MATVEC, 3.0 * n2 * L.work_items);
    CK(laplace_lik_launch(L, c.mu, c.yf, c.tau, c.s, c.nu, c.w, c.sig, c.n_of, np, nb, c.active));
    CK(form_B_launch(L, c.ms.K, c.s, c.ms.A, np, st, nb, c.active));
    CK(potrf_blocked(L, c.ms, c.active, c.info));
    CK(trtri_blocked(L, c.ms, c.active));
    CK(rowdot_launch(L, 0, false, c.ms.K, c.nu, c.kv, nullptr, np, np, st, np, np, nb, c.active));   // K b
    CK(mul_launch(L, c.s, c.kv, c.u, np, nb, c.active));
    CK(rowdot_launch(L, 1, false, c.ms.M, c.u, c.t, nullptr, np, np, st, np, np, nb, c.active));
    CK(rowdot_launch(L, 2, false, c.ms.U, c.t, c.z, nullptr, np, np, st, np, np, nb, c.active));     // B^-1 (s o K b)
    CK(mul_launch(L, c.s, c.z, c.tmp, np, nb, c.active));
    CK(axpy_launch(L, c.nu, c.tmp, -1.0, c.u, np, nb));                                              // a_new
    CK(rowdot_launch(L, 0, false, c.ms.K, c.u, c.kv, nullptr, np, np, st, np, np, nb, c.active));    // f_new = K a_new
    CK(laplace_step_launch(L, c.mu, c.alpha, c.kv, c.u, c.yf, c.nlml_gauss, c.n_of, np, nb, tol, c.active, c.sweeps,
                           c.status, c.info));
    int n_active = 0;
    int rc = poll_active(ctx, L, c.active, nb, &n_active);
    if (rc) return rc;
    if (n_active == 0) break;
    L.work_items = n_active;
  }
  L.work_items = nb;
  mark_unconverged_kernel<<<(nb + 255) / 256, 256, 0, L.stream>>>(c.active, c.status, nb);
  L.bump();
  // quantities at the final mode
  CK(cudaMemsetAsync(c.info, 0, sizeof(int) * nb, L.stream));
  L.add_work(CLS_FORMB, n2 * nb);
  L.add_work(CLS_MATVEC, 2.5 * n2 * nb);
  CK(laplace_lik_launch(L, c.mu, c.yf, c.tau, c.s, c.nu, c.w, c.sig, c.n_of, np, nb, nullptr));
  CK(form_B_launch(L, c.ms.K, c.s, c.ms.A, np, st, nb, nullptr));
  CK(potrf_blocked(L, c.ms, nullptr, c.info));
  CK(trtri_blocked(L, c.ms, nullptr));
  CK(rowdot_launch(L, 2, true, c.ms.U, c.nu, c.tmp, c.dB, np, np, st, np, np, nb, nullptr));         // diag(B^-1)
  CK(laplace_final_launch(L, c.mu, c.alpha, c.tau, c.w, c.sig, c.dB, c.yf, c.ms.A, st, c.n_of, np, nb, c.progs, c.z,
                          c.nlml, c.nlml_gauss, c.bic, c.status, c.info));
  if (want_grad) {
    // u = s2 - R K s2,  R = S^1/2 B^-1 S^1/2
    CK(rowdot_launch(L, 0, false, c.ms.K, c.z, c.kv, nullptr, np, np, st, np, np, nb, nullptr));
    CK(mul_launch(L, c.s, c.kv, c.u, np, nb, nullptr));
    CK(rowdot_launch(L, 1, false, c.ms.M, c.u, c.t, nullptr, np, np, st, np, np, nb, nullptr));
    CK(rowdot_launch(L, 2, false, c.ms.U, c.t, c.tmp, nullptr, np, np, st, np, np, nb, nullptr));
    CK(mul_launch(L, c.s, c.tmp, c.u, np, nb, nullptr));
    CK(axpy_launch(L, c.z, c.u, -1.0, c.dB, np, nb));
    CK(lauum_blocked(L, c.ms, c.ms.T, nullptr));
    GradArgs ga{};
    ga.progs = c.progs; ga.theta = c.theta; ga.theta_stride = theta_stride; ga.Xf = c.Xf; ga.n_of = c.n_of;
    ga.np = np; ga.D = D; ga.G = c.ms.T; ga.ld = np; ga.g_batch = st; ga.alpha = c.alpha; ga.s = c.s;
    ga.ru = c.dB; ga.rv = c.w;
    ga.mode = 0; ga.partial = c.partial; ga.tiles_per_item = grad_tiles(np, 0); ga.active = nullptr;
    L.add_work(CLS_GRAD, 4.0 * L.work_n * L.work_n * nb);
    CK(cudaMemsetAsync(c.grad, 0, sizeof(double) * (size_t)nb * theta_stride, L.stream));
    CK(grad_contract_launch(L, ga, nb, -1.0, c.grad, theta_stride));
  }
  return AGPC_OK;
}

}  // namespace

// The scoring path proper.  Item b reads the row list rows_dev[row_start[b] .. row_start[b] + row_len[b]) -- the lists need
// not be adjacent or ordered, which is what lets the on-device optimiser (optim.cu) pass the still-running subset of a
// resident batch without re-packing its rows.
int score_items_dev(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                    const double* theta_dev, int32_t theta_stride, const int32_t* rows_dev, const int64_t* row_start,
                    const int32_t* row_len, const agpc_ep_options* opt, double* tau_dev, double* nu_dev,
                    int32_t site_stride, double* nlml_dev, double* grad_dev, double* bic_dev, double* nlml_gauss_dev,
                    int32_t* sweeps_dev, int32_t* status_dev, void* stream) {
  if (!ctx) return AGPC_E_ARG;
  if (dataset_id < 0 || dataset_id >= (int)ctx->datasets.size() || !ctx->datasets[dataset_id].X)
    return fail(ctx, AGPC_E_ARG, "score_batch: bad dataset id");
  if (n_items <= 0 || !programs || !theta_dev || !rows_dev || !row_start || !row_len || !opt || !nlml_dev || !status_dev)
    return fail(ctx, AGPC_E_ARG, "score_batch: null argument");
  if (theta_stride <= 0 || theta_stride > 4096) return fail(ctx, AGPC_E_ARG, "score_batch: bad theta_stride");
  const Dataset& ds = ctx->datasets[dataset_id];
  int n_max = 0;
  for (int b = 0; b < n_items; ++b) {
    const long long n = row_len[b];
    if (n <= 0 || n > ds.N * 4LL) return fail(ctx, AGPC_E_ARG, "score_batch: bad row_off");
    n_max = std::max<int>(n_max, (int)n);
    const agpc_program& p = programs[b];
    if (p.n_leaves <= 0 || p.n_leaves > AGPC_MAX_LEAVES || p.n_terms <= 0 || p.n_terms > AGPC_MAX_TERMS ||
        p.n_theta <= 0 || p.n_theta > AGPC_MAX_THETA || p.n_theta > theta_stride)
      return fail(ctx, AGPC_E_ARG, "score_batch: malformed kernel program");
    for (int l = 0; l < p.n_leaves; ++l)
      if (p.leaf_type[l] < 0 || p.leaf_type[l] > AGPC_LEAF_CONST || p.leaf_dim[l] < 0 || p.leaf_dim[l] >= ds.D ||
          p.leaf_off[l] < 0 || p.leaf_off[l] >= p.n_theta)
        return fail(ctx, AGPC_E_ARG, "score_batch: malformed kernel program leaf");
    for (int t = 0; t < p.n_terms; ++t) {
      if (p.term_len[t] <= 0 || p.term_len[t] > AGPC_MAX_TERM_LEN) return fail(ctx, AGPC_E_ARG, "score_batch: bad term");
      for (int q = 0; q < p.term_len[t]; ++q)
        if (p.term_leaf[t][q] < 0 || p.term_leaf[t][q] >= p.n_leaves) return fail(ctx, AGPC_E_ARG, "score_batch: bad term leaf");
    }
  }
  if ((tau_dev || nu_dev) && site_stride < n_max) return fail(ctx, AGPC_E_ARG, "score_batch: site_stride < n");
  if ((opt->warm_start || opt->max_sweeps == 0) && (!tau_dev || !nu_dev))
    return fail(ctx, AGPC_E_ARG, "score_batch: warm start / frozen sites need tau and nu");
  CK(cudaSetDevice(ctx->device));
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};

  // chunking: consecutive items, as many as the arena budget allows at the padded size of the largest item
  const int np_all = (int)align_up((size_t)n_max, TILE);
  const size_t per_item = per_item_bytes(np_all, ds.D, theta_stride, ctx->oz_slices(np_all));
  const size_t budget = ctx->workspace_budget();
  int chunk_cap = (int)std::min<size_t>((size_t)n_items, std::max<size_t>(1, budget / per_item));
  chunk_cap = std::min(chunk_cap, 65535);  // the batch index is gridDim.y / gridDim.z of most launches
  if (per_item > budget) return fail(ctx, AGPC_E_NOMEM, "score_batch: one item does not fit the workspace budget");
  Chunk probe;
  const size_t need = probe.carve(nullptr, chunk_cap, np_all, ds.D, theta_stride, false, ctx->oz_slices(np_all));
  int rc = ctx->ensure_arena(need);
  if (rc) return rc;

  for (int first = 0; first < n_items; first += chunk_cap) {
    const int nb = std::min(chunk_cap, n_items - first);
    int np = 0;
    std::vector<int> n_of(nb);
    std::vector<long long> off(nb + 1);
    for (int b = 0; b < nb; ++b) {
      n_of[b] = row_len[first + b];
      off[b] = row_start[first + b];
      np = std::max(np, n_of[b]);
    }
    off[nb] = off[nb - 1] + n_of[nb - 1];
    np = (int)align_up((size_t)np, TILE);
    Chunk c;
    c.carve((char*)ctx->arena, nb, np, ds.D, theta_stride, false, ctx->oz_slices(np_all));
    if (c.make_maps()) return fail(ctx, AGPC_E_CUDA, "cuTensorMapEncodeTiled failed");
    const int tb = 256;
    dim3 gvec((np + tb - 1) / tb, nb);
    {
      double m3 = 0.0;
      for (int b = 0; b < nb; ++b) m3 += (double)n_of[b] * n_of[b] * n_of[b];
      L.work_n = cbrt(m3 / nb);
      L.work_items = nb;
    }

    CK(cudaMemcpyAsync(c.n_of, n_of.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(c.row_off, off.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(c.progs, programs + first, sizeof(agpc_program) * nb, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpy2DAsync(c.theta, sizeof(double) * theta_stride, theta_dev + (size_t)first * theta_stride,
                         sizeof(double) * theta_stride, sizeof(double) * theta_stride, nb, cudaMemcpyDeviceToDevice, L.stream));
    fill_int_kernel<<<(nb + 255) / 256, 256, 0, L.stream>>>(c.active, 1, nb);
    fill_int_kernel<<<(nb + 255) / 256, 256, 0, L.stream>>>(c.ones, 1, nb);
    L.bump(2);
    CK(cudaMemsetAsync(c.sweeps, 0, sizeof(int) * nb, L.stream));
    CK(cudaMemsetAsync(c.status, 0, sizeof(int) * nb, L.stream));
    CK(cudaMemsetAsync(c.converged, 0, sizeof(int) * nb, L.stream));
    CK(cudaMemsetAsync(c.info, 0, sizeof(int) * nb, L.stream));

    CK(gather_rows_launch(L, ds.X, ds.y, rows_dev, c.row_off, c.n_of, ds.D, np, nb, c.Xf, c.yf));
    BuildArgs ba{};
    ba.progs = c.progs; ba.theta = c.theta; ba.theta_stride = theta_stride;
    ba.Xa = c.Xf; ba.Xb = c.Xf; ba.na_of = c.n_of; ba.nb_of = c.n_of;
    ba.na_pad = np; ba.nb_pad = np; ba.D = ds.D; ba.xa_batch = (long long)np * ds.D; ba.xb_batch = ba.xa_batch;
    ba.out = c.ms.K; ba.ld = np; ba.out_batch = c.ms.stride; ba.dK = nullptr; ba.kdiag = c.kdiag; ba.active = nullptr;
    L.add_work(CLS_KBUILD, 8.0 * L.work_n * L.work_n * nb);
    CK(build_K_launch(L, ba, nb, false));

    const bool laplace = opt->inference == AGPC_INFER_LAPLACE;
    const bool frozen = !laplace && opt->max_sweeps <= 0;
    const bool warm = !laplace && (frozen || opt->warm_start != 0);
    rc = laplace ? laplace_chunk(ctx, L, c, opt, theta_stride, grad_dev != nullptr, ds.D)
                 : ep_chunk(ctx, L, c, opt, theta_stride, grad_dev != nullptr, ds.D, warm, frozen,
                            warm ? tau_dev + (size_t)first * site_stride : nullptr,
                            warm ? nu_dev + (size_t)first * site_stride : nullptr, site_stride);
    if (rc) return rc;
    if (grad_dev)
      CK(cudaMemcpyAsync(grad_dev + (size_t)first * theta_stride, c.grad, sizeof(double) * (size_t)nb * theta_stride,
                         cudaMemcpyDeviceToDevice, L.stream));
    // ---- outputs ----
    CK(cudaMemcpyAsync(nlml_dev + first, c.nlml, sizeof(double) * nb, cudaMemcpyDeviceToDevice, L.stream));
    if (bic_dev) CK(cudaMemcpyAsync(bic_dev + first, c.bic, sizeof(double) * nb, cudaMemcpyDeviceToDevice, L.stream));
    if (nlml_gauss_dev) CK(cudaMemcpyAsync(nlml_gauss_dev + first, c.nlml_gauss, sizeof(double) * nb, cudaMemcpyDeviceToDevice, L.stream));
    if (sweeps_dev) CK(cudaMemcpyAsync(sweeps_dev + first, c.sweeps, sizeof(int) * nb, cudaMemcpyDeviceToDevice, L.stream));
    CK(cudaMemcpyAsync(status_dev + first, c.status, sizeof(int) * nb, cudaMemcpyDeviceToDevice, L.stream));
    if (tau_dev && nu_dev && !frozen) {
      copy_strided_kernel<<<gvec, tb, 0, L.stream>>>(tau_dev + (size_t)first * site_stride, site_stride, c.tau, np, c.n_of, site_stride < np ? site_stride : np, 0);
      copy_strided_kernel<<<gvec, tb, 0, L.stream>>>(nu_dev + (size_t)first * site_stride, site_stride, c.nu, np, c.n_of, site_stride < np ? site_stride : np, 0);
      L.bump(2);
    }
    CK(cudaGetLastError());
    if (first + chunk_cap < n_items) CK(cudaStreamSynchronize(L.stream));  // arena is reused by the next chunk
  }
  return AGPC_OK;
}

extern "C" {

int agpc_score_batch_dev(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                         const double* theta_dev, int32_t theta_stride, const int32_t* rows_dev,
                         const int64_t* row_off, const agpc_ep_options* opt, double* tau_dev, double* nu_dev,
                         int32_t site_stride, double* nlml_dev, double* grad_dev, double* bic_dev,
                         double* nlml_gauss_dev, int32_t* sweeps_dev, int32_t* status_dev, void* stream) {
  if (!ctx) return AGPC_E_ARG;
  if (n_items <= 0 || !row_off) return fail(ctx, AGPC_E_ARG, "score_batch: null argument");
  std::vector<int32_t> len(n_items);
  for (int b = 0; b < n_items; ++b) {
    const long long n = row_off[b + 1] - row_off[b];
    if (n <= 0 || n > 0x7fffffffLL) return fail(ctx, AGPC_E_ARG, "score_batch: bad row_off");
    len[b] = (int32_t)n;
  }
  return score_items_dev(ctx, dataset_id, n_items, programs, theta_dev, theta_stride, rows_dev, row_off, len.data(), opt,
                         tau_dev, nu_dev, site_stride, nlml_dev, grad_dev, bic_dev, nlml_gauss_dev, sweeps_dev, status_dev,
                         stream);
}

int agpc_score_batch(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                     const double* theta, int32_t theta_stride, const int32_t* rows, const int64_t* row_off,
                     const agpc_ep_options* opt, double* tau, double* nu, int32_t site_stride, double* nlml,
                     double* grad, double* bic, double* nlml_gauss, int32_t* sweeps, int32_t* status) {
  if (!ctx) return AGPC_E_ARG;
  if (n_items <= 0 || !theta || !rows || !row_off || !nlml || !status || theta_stride <= 0)
    return fail(ctx, AGPC_E_ARG, "score_batch: null argument");
  if (dataset_id < 0 || dataset_id >= (int)ctx->datasets.size() || !ctx->datasets[dataset_id].X)
    return fail(ctx, AGPC_E_ARG, "score_batch: bad dataset id");
  {  // the row lists index the device-resident dataset: reject anything outside it before a kernel dereferences it
    const int N = ctx->datasets[dataset_id].N;
    if (row_off[0] != 0) return fail(ctx, AGPC_E_ARG, "score_batch: row_off[0] must be 0");
    for (int b = 0; b < n_items; ++b)
      if (row_off[b + 1] <= row_off[b]) return fail(ctx, AGPC_E_ARG, "score_batch: empty or unordered row list");
    for (int64_t i = 0; i < row_off[n_items]; ++i)
      if (rows[i] < 0 || rows[i] >= N) return fail(ctx, AGPC_E_ARG, "score_batch: row index outside the dataset");
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n_rows = (size_t)row_off[n_items];
  const size_t nth = (size_t)n_items * theta_stride;
  const bool sites = tau && nu;
  const size_t nsite = sites ? (size_t)n_items * site_stride : 0;
  // device staging (separate from the arena, which score_batch_dev may re-allocate); persistent in the context
  double *d_theta, *d_tau, *d_nu, *d_nlml, *d_grad, *d_bic, *d_gauss;
  int *d_rows, *d_sweeps, *d_status;
  auto carve = [&](char* base) {
    Carver c{base};
    d_theta = c.take<double>(nth);
    d_rows = c.take<int>(n_rows);
    d_tau = c.take<double>(nsite + 1);
    d_nu = c.take<double>(nsite + 1);
    d_nlml = c.take<double>(n_items);
    d_grad = c.take<double>(nth);
    d_bic = c.take<double>(n_items);
    d_gauss = c.take<double>(n_items);
    d_sweeps = c.take<int>(n_items);
    d_status = c.take<int>(n_items);
    return align_up(c.off, 1024);
  };
  const size_t total = carve(nullptr);
  char* stage = nullptr;
  static const bool trace = getenv("AGPC_TIMING") != nullptr;   // host-phase wall times of this call on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = now();
  if (ctx->stage_bytes < total) {   // grow-only: no cudaMalloc / cudaFree on the steady-state path
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    const size_t want = total + total / 4;
    if (cudaMalloc(&ctx->stage, want) != cudaSuccess) return fail(ctx, AGPC_E_NOMEM, "score_batch staging cudaMalloc");
    ctx->stage_bytes = want;
  }
  stage = static_cast<char*>(ctx->stage);
  carve(stage);
  const auto t1 = now();
  auto t2 = t1, t3 = t1;
  int rc = AGPC_OK;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMemcpyAsync(d_theta, theta, sizeof(double) * nth, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d_rows, rows, sizeof(int) * n_rows, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if (sites && (opt->warm_start || opt->max_sweeps <= 0)) {
      if ((e = cudaMemcpyAsync(d_tau, tau, sizeof(double) * nsite, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(d_nu, nu, sizeof(double) * nsite, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    }
    t2 = now();
    rc = agpc_score_batch_dev(ctx, dataset_id, n_items, programs, d_theta, theta_stride, d_rows, row_off, opt,
                              sites ? d_tau : nullptr, sites ? d_nu : nullptr, site_stride, d_nlml, grad ? d_grad : nullptr,
                              d_bic, d_gauss, d_sweeps, d_status, st);
    t3 = now();
    if (rc) break;
    if ((e = cudaMemcpyAsync(nlml, d_nlml, sizeof(double) * n_items, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if (grad && (e = cudaMemcpyAsync(grad, d_grad, sizeof(double) * nth, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if (bic && (e = cudaMemcpyAsync(bic, d_bic, sizeof(double) * n_items, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if (nlml_gauss && (e = cudaMemcpyAsync(nlml_gauss, d_gauss, sizeof(double) * n_items, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if (sweeps && (e = cudaMemcpyAsync(sweeps, d_sweeps, sizeof(int) * n_items, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(status, d_status, sizeof(int) * n_items, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if (sites && opt->max_sweeps > 0) {
      if ((e = cudaMemcpyAsync(tau, d_tau, sizeof(double) * nsite, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(nu, d_nu, sizeof(double) * nsite, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    }
    e = cudaStreamSynchronize(st);
  } while (0);
  cudaStreamSynchronize(st);
  const auto t4 = now();
  if (trace)
    fprintf(stderr, "[agpc] score_batch host phases (ms): staging %.3f  h2d %.3f  device %.3f  d2h+sync %.3f\n",
            ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, AGPC_E_CUDA, "score_batch host copies", e);
  return AGPC_OK;
}


static int predict_impl(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                        const double* theta, int32_t theta_stride, const int32_t* rows, const int64_t* row_off,
                        const double* tau, const double* nu, int32_t site_stride, const int32_t* test_rows,
                        const int64_t* test_off, const double* Xstar_host, double* p_out, double* mu_out,
                        double* var_out, int32_t* status, double* dmu_dx) {
  if (!ctx) return AGPC_E_ARG;
  if (dataset_id < 0 || dataset_id >= (int)ctx->datasets.size() || !ctx->datasets[dataset_id].X)
    return fail(ctx, AGPC_E_ARG, "predict_batch: bad dataset id");
  if (n_items <= 0 || !programs || !theta || !rows || !row_off || !tau || !nu || (!test_rows && !Xstar_host) || !test_off || !p_out)
    return fail(ctx, AGPC_E_ARG, "predict: null argument");
  if (Xstar_host && n_items != 1) return fail(ctx, AGPC_E_ARG, "predict_points: one item only");
  const Dataset& ds = ctx->datasets[dataset_id];
  for (int64_t i = 0; i < row_off[n_items]; ++i)
    if (rows[i] < 0 || rows[i] >= ds.N) return fail(ctx, AGPC_E_ARG, "predict: training row index outside the dataset");
  if (test_rows)
    for (int64_t i = 0; i < test_off[n_items]; ++i)
      if (test_rows[i] < 0 || test_rows[i] >= ds.N) return fail(ctx, AGPC_E_ARG, "predict: test row index outside the dataset");
  CK(cudaSetDevice(ctx->device));
  Launch L{ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  int n_max = 0, m_max = 0;
  for (int b = 0; b < n_items; ++b) {
    n_max = std::max<int>(n_max, (int)(row_off[b + 1] - row_off[b]));
    m_max = std::max<int>(m_max, (int)(test_off[b + 1] - test_off[b]));
    if (programs[b].n_theta > theta_stride || programs[b].n_theta <= 0) return fail(ctx, AGPC_E_ARG, "predict_batch: bad program");
  }
  if (n_max <= 0 || m_max <= 0 || site_stride < n_max) return fail(ctx, AGPC_E_ARG, "predict_batch: bad sizes");
  const int np_all = (int)align_up((size_t)n_max, TILE), mp_all = (int)align_up((size_t)m_max, TILE);
  const size_t extra_item = sizeof(double) * ((size_t)2 * mp_all * np_all + (size_t)mp_all * (ds.D + 6)) + sizeof(int) * (size_t)(mp_all + np_all) +
                            sizeof(double) * (size_t)site_stride + 16384;
  const size_t per_item = per_item_bytes(np_all, ds.D, theta_stride, 0) + extra_item;
  const size_t budget = ctx->workspace_budget();
  if (per_item > budget) return fail(ctx, AGPC_E_NOMEM, "predict_batch: one item does not fit the workspace budget");
  const int chunk_cap = std::min((int)std::min<size_t>((size_t)n_items, std::max<size_t>(1, budget / per_item)), 65535);
  int rc = ctx->ensure_arena(per_item * chunk_cap + (1 << 20));
  if (rc) return rc;

  for (int first = 0; first < n_items; first += chunk_cap) {
    const int nb = std::min(chunk_cap, n_items - first);
    std::vector<int> n_of(nb), m_of(nb);
    std::vector<long long> off(nb + 1), toff(nb + 1);
    int np = 0, mp = 0;
    for (int b = 0; b < nb; ++b) {
      n_of[b] = (int)(row_off[first + b + 1] - row_off[first + b]);
      m_of[b] = (int)(test_off[first + b + 1] - test_off[first + b]);
      off[b] = row_off[first + b] - row_off[first];
      toff[b] = test_off[first + b] - test_off[first];
      np = std::max(np, n_of[b]);
      mp = std::max(mp, m_of[b]);
    }
    off[nb] = row_off[first + nb] - row_off[first];
    toff[nb] = test_off[first + nb] - test_off[first];
    np = (int)align_up((size_t)np, TILE);
    mp = (int)align_up((size_t)mp, TILE);
    Chunk c;
    const size_t used = c.carve((char*)ctx->arena, nb, np, ds.D, theta_stride, true, 0);
    Carver x{(char*)ctx->arena};
    x.off = used;
    double* Ks = x.take<double>((size_t)nb * mp * np);
    double* V = x.take<double>((size_t)nb * mp * np);
    double* Xs = x.take<double>((size_t)nb * mp * ds.D);
    double* ys = x.take<double>((size_t)nb * mp);
    double* mu_s = x.take<double>((size_t)nb * mp);
    double* kss = x.take<double>((size_t)nb * mp);
    double* vsq = x.take<double>((size_t)nb * mp);
    double* pbuf = x.take<double>((size_t)nb * mp);
    double* varbuf = x.take<double>((size_t)nb * mp);
    int* trows = x.take<int>((size_t)nb * mp);
    int* m_dev = x.take<int>(nb);
    long long* toff_dev = x.take<long long>(nb + 1);
    double* site_tmp = x.take<double>((size_t)nb * site_stride);
    if (x.off > ctx->arena_bytes) return fail(ctx, AGPC_E_NOMEM, "predict_batch: arena carve overflow");
    if (c.make_maps()) return fail(ctx, AGPC_E_CUDA, "cuTensorMapEncodeTiled failed");
    CUtensorMap mapKs;
    if (make_tensor_map(&mapKs, Ks, np, mp, nb, np, (long long)mp * np)) return fail(ctx, AGPC_E_CUDA, "cuTensorMapEncodeTiled failed");
    const int tb = 256;
    dim3 gvec((np + tb - 1) / tb, nb);
    const long long nrows_chunk = off[nb], ntest_chunk = toff[nb];
    CK(cudaMemcpyAsync(c.n_of, n_of.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(m_dev, m_of.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(c.row_off, off.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(toff_dev, toff.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(c.progs, programs + first, sizeof(agpc_program) * nb, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(c.theta, theta + (size_t)first * theta_stride, sizeof(double) * (size_t)nb * theta_stride, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemcpyAsync(c.rows, rows + row_off[first], sizeof(int) * nrows_chunk, cudaMemcpyHostToDevice, L.stream));
    if (!Xstar_host) CK(cudaMemcpyAsync(trows, test_rows + test_off[first], sizeof(int) * ntest_chunk, cudaMemcpyHostToDevice, L.stream));
    CK(cudaMemsetAsync(c.status, 0, sizeof(int) * nb, L.stream));
    CK(cudaMemsetAsync(c.info, 0, sizeof(int) * nb, L.stream));
    CK(gather_rows_launch(L, ds.X, ds.y, c.rows, c.row_off, c.n_of, ds.D, np, nb, c.Xf, c.yf));
    if (!Xstar_host) {
      CK(gather_rows_launch(L, ds.X, ds.y, trows, toff_dev, m_dev, ds.D, mp, nb, Xs, ys));
    } else {
      CK(cudaMemsetAsync(Xs, 0, sizeof(double) * (size_t)mp * ds.D, L.stream));
      CK(cudaMemcpyAsync(Xs, Xstar_host, sizeof(double) * (size_t)m_of[0] * ds.D, cudaMemcpyHostToDevice, L.stream));
    }
    BuildArgs ba{};
    ba.progs = c.progs; ba.theta = c.theta; ba.theta_stride = theta_stride;
    ba.Xa = c.Xf; ba.Xb = c.Xf; ba.na_of = c.n_of; ba.nb_of = c.n_of;
    ba.na_pad = np; ba.nb_pad = np; ba.D = ds.D; ba.xa_batch = (long long)np * ds.D; ba.xb_batch = ba.xa_batch;
    ba.out = c.ms.K; ba.ld = np; ba.out_batch = c.ms.stride; ba.kdiag = c.kdiag;
    CK(build_K_launch(L, ba, nb, false));
    // sites in
    CK(cudaMemcpyAsync(site_tmp, tau + (size_t)first * site_stride, sizeof(double) * (size_t)nb * site_stride, cudaMemcpyHostToDevice, L.stream));
    copy_strided_kernel<<<gvec, tb, 0, L.stream>>>(c.tau, np, site_tmp, site_stride, c.n_of, np, 1);
    CK(cudaStreamSynchronize(L.stream));
    CK(cudaMemcpyAsync(site_tmp, nu + (size_t)first * site_stride, sizeof(double) * (size_t)nb * site_stride, cudaMemcpyHostToDevice, L.stream));
    copy_strided_kernel<<<gvec, tb, 0, L.stream>>>(c.nu, np, site_tmp, site_stride, c.n_of, np, 1);
    clamp_tau_kernel<<<gvec, tb, 0, L.stream>>>(c.tau, c.kdiag, c.n_of, np);
    L.bump(3);
    CK(posterior(ctx, L, c, nullptr));
    CK(ep_final_launch(L, c.tau, c.nu, c.sig, c.mu, c.yf, c.s, c.z, c.ms.A, c.ms.stride, c.n_of, np, nb, c.progs,
                       c.alpha, c.nlml, c.nlml_gauss, c.bic, c.status, c.info));
    // cross kernel, latent mean and variance
    BuildArgs bs = ba;
    bs.Xa = Xs; bs.na_of = m_dev; bs.na_pad = mp; bs.xa_batch = (long long)mp * ds.D;
    bs.out = Ks; bs.ld = np; bs.out_batch = (long long)mp * np; bs.kdiag = nullptr;
    CK(build_K_launch(L, bs, nb, false));
    CK(rowdot_launch(L, 0, false, Ks, c.alpha, mu_s, nullptr, mp, np, (long long)mp * np, np, mp, nb, nullptr));
    CK(kself_launch(L, c.progs, c.theta, theta_stride, Xs, mp, ds.D, m_dev, nb, kss));
    CK(scale_cols_launch(L, Ks, c.s, mp, np, nb));
    GemmArgs g{};
    g.C = V; g.ldc = np; g.c_batch = (long long)mp * np;
    g.m_tiles = mp / TILE; g.n_tiles = np / TILE; g.k_tiles = np / GEMM_BK;
    g.k_rule = KRULE_END_AT_COL; g.alpha = 1.0; g.beta = 0.0;
    CK(gemm_nt_launch(L, mapKs, c.ms.mapM, g, g.m_tiles * g.n_tiles, 1, nb));
    CK(rowdot_launch(L, 0, true, V, c.alpha, pbuf /*unused dot*/, vsq, mp, np, (long long)mp * np, np, mp, nb, nullptr));
    CK(predict_prob_launch(L, mu_s, kss, vsq, pbuf, varbuf, (long long)nb * mp));
    if (dmu_dx) {
      // V is free again: use it as the m x D output buffer
      CK(dmu_dx_launch(L, c.progs, c.theta, Xs, m_of[0], c.Xf, n_of[0], ds.D, c.alpha, V));
      CK(cudaMemcpyAsync(dmu_dx, V, sizeof(double) * (size_t)m_of[0] * ds.D, cudaMemcpyDeviceToHost, L.stream));
    }
    CK(cudaStreamSynchronize(L.stream));
    for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)test_off[first + b];
      CK(cudaMemcpyAsync(p_out + o, pbuf + (size_t)b * mp, sizeof(double) * m_of[b], cudaMemcpyDeviceToHost, L.stream));
      if (mu_out) CK(cudaMemcpyAsync(mu_out + o, mu_s + (size_t)b * mp, sizeof(double) * m_of[b], cudaMemcpyDeviceToHost, L.stream));
      if (var_out) CK(cudaMemcpyAsync(var_out + o, varbuf + (size_t)b * mp, sizeof(double) * m_of[b], cudaMemcpyDeviceToHost, L.stream));
    }
    if (status) CK(cudaMemcpyAsync(status + first, c.status, sizeof(int) * nb, cudaMemcpyDeviceToHost, L.stream));
    CK(cudaStreamSynchronize(L.stream));
  }
  return AGPC_OK;
}

int agpc_predict_batch(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                       const double* theta, int32_t theta_stride, const int32_t* rows, const int64_t* row_off,
                       const double* tau, const double* nu, int32_t site_stride, const int32_t* test_rows,
                       const int64_t* test_off, double* p_out, double* mu_out, double* var_out, int32_t* status) {
  if (!test_rows) return ctx ? fail(ctx, AGPC_E_ARG, "predict_batch: null test_rows") : AGPC_E_ARG;
  return predict_impl(ctx, dataset_id, n_items, programs, theta, theta_stride, rows, row_off, tau, nu, site_stride,
                      test_rows, test_off, nullptr, p_out, mu_out, var_out, status, nullptr);
}

int agpc_predict_points(agpc_ctx* ctx, int32_t dataset_id, const agpc_program* program, const double* theta,
                        const int32_t* rows, int32_t n_rows, const double* tau, const double* nu,
                        const double* Xstar_host, int32_t m, double* p_out, double* mu_out, double* var_out,
                        double* dmu_dx) {
  if (!ctx) return AGPC_E_ARG;
  if (!program || !Xstar_host || m <= 0 || n_rows <= 0) return fail(ctx, AGPC_E_ARG, "predict_points: bad argument");
  const int64_t row_off[2] = {0, n_rows};
  const int64_t test_off[2] = {0, m};
  int32_t status = 0;
  return predict_impl(ctx, dataset_id, 1, program, theta, program->n_theta, rows, row_off, tau, nu, n_rows, nullptr,
                      test_off, Xstar_host, p_out, mu_out, var_out, &status, dmu_dx);
}

}  // extern "C"
