/*
 * autogpc_b200 -- C ABI of the B200-native candidate-scoring engine for AutoGPC.
 *
 * The reference (charles92/autogpc) has no FFI: its scoring path is the slice of the GPy *Python object API* that
 * gpckernel.py touches.  Each entry point below names the reference call site(s) it stands in for; the Python host
 * (autogpc_b200/gpy_compat.py) binds them with ctypes and re-creates that object API on top, see INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and a negative AGPC_E_* code on API misuse / CUDA failure
 * (text via agpc_last_error).  Per-item numerical failures are reported in status[] (AGPC_ST_*) and never abort a
 * batch -- this is what lets GPCKernel.train's per-fold skip (gpckernel.py:196-199) and paramz's <=10 tolerated
 * linear-algebra failures be reproduced by the host.  One context per host thread; contexts are not shared.
 * All matrices are FP64 row-major.  "dev" pointers are CUDA device pointers on the context's device, "host"
 * pointers are ordinary host memory.  `stream` is a cudaStream_t passed as void*; NULL means the CONTEXT'S OWN stream,
 * which is non-blocking: it is not ordered with the legacy default stream, so a caller that prepared device buffers on
 * another stream must either pass that stream or synchronise before the call (and after it, before reading results).
 *
 * Memory: the context owns a workspace arena (grown on demand up to agpc_set_workspace_limit), a grow-only staging
 * buffer for the host-buffer entry points and a small pool of destroyed datasets' buffers, so the steady-state path
 * performs no cudaMalloc / cudaFree.  Environment: AGPC_TIMING=1 prints the host-phase wall times of every
 * agpc_score_batch call on stderr (staging, H2D, device call, D2H + sync).
 */
#ifndef AUTOGPC_B200_H
#define AUTOGPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGPC_VERSION 101

/* error codes */
#define AGPC_OK 0
#define AGPC_E_ARG (-1)
#define AGPC_E_CUDA (-2)
#define AGPC_E_NOMEM (-3)
#define AGPC_E_UNSUPPORTED (-4)

/* per-item status */
#define AGPC_ST_OK 0
#define AGPC_ST_JITTER 1
#define AGPC_ST_NOT_PD 2
#define AGPC_ST_EP_NOT_CONVERGED 3
#define AGPC_ST_NONFINITE 4

/* leaf opcodes of a kernel program; theta layout per leaf follows gpss2gpy (gpckernel.py:535-565):
 *   SE    = GPy.kern.RBF          [variance, lengthscale]          (gpckernel.py:539)
 *   PER   = GPy.kern.StdPeriodic  [variance, period, lengthscale]  (gpckernel.py:551)
 *   LIN   = ff.LinearKernel       [variance, location]             (flexible_function.py:1284; no GPy analogue in
 *                                                                   the reference: gpckernel.py:573 raises)
 *   CONST = GPy.kern.Bias         [variance]                       (gpckernel.py:564)                              */
#define AGPC_LEAF_SE 0
#define AGPC_LEAF_PER 1
#define AGPC_LEAF_LIN 2
#define AGPC_LEAF_CONST 3

#define AGPC_MAX_LEAVES 16
#define AGPC_MAX_TERMS 16
#define AGPC_MAX_TERM_LEN 8
#define AGPC_MAX_THETA 48

/* Kernel program: K = sum_t prod_{l in term t} k_l  -- GPy.kern.Add / GPy.kern.Prod (gpckernel.py:568,571) with
 * products distributed over sums.  Fixed-width POD so an item table is a plain array. */
typedef struct agpc_program {
  int32_t n_leaves, n_terms, n_theta, p_eff; /* p_eff: flexible_function.py:69-74,1554,1647 (BIC parameter count) */
  int32_t leaf_type[AGPC_MAX_LEAVES];
  int32_t leaf_dim[AGPC_MAX_LEAVES];
  int32_t leaf_off[AGPC_MAX_LEAVES]; /* offset of the leaf's first parameter in theta */
  int32_t term_len[AGPC_MAX_TERMS];
  int32_t term_leaf[AGPC_MAX_TERMS][AGPC_MAX_TERM_LEN];
} agpc_program;

/* EP options: GPy EP(epsilon=1e-6, eta=1, delta=1) (expectation_propagation.py); damping = delta. */
typedef struct agpc_ep_options {
  double epsilon;     /* stop when mean(d tau^2) <= eps and mean(d nu^2) <= eps (tested from sweep 2 on)          */
  double damping;     /* site update step: 1.0 = GPy's delta; <= 0 = automatic schedule (1.0 for 3 sweeps, then   */
                      /* 0.8), which is what makes the parallel schedule converge wherever sequential EP does     */
  int32_t max_sweeps; /* 0 = keep the supplied sites frozen (2016-GPy optimisation semantics)                     */
  int32_t warm_start; /* 1 = start from tau/nu as supplied, 0 = cold start (tau = nu = 0)                         */
  int32_t inference;  /* AGPC_INFER_EP (GPy's default for GPClassification) or AGPC_INFER_LAPLACE (R&W alg. 3.1 + */
                      /* 5.1; epsilon = relative tolerance on the Newton objective, max_sweeps = Newton iterations,*/
                      /* always cold-started; tau / nu return W and W f + grad log p, the site form of the         */
                      /* Laplace posterior, so agpc_predict_* work unchanged)                                       */
  int32_t reserved;
} agpc_ep_options;
#define AGPC_INFER_EP 0
#define AGPC_INFER_LAPLACE 1

typedef struct agpc_ctx agpc_ctx;

/* ---- context -------------------------------------------------------------------------------------------- */
int agpc_version(void);
int agpc_create(int device, agpc_ctx** out);
int agpc_destroy(agpc_ctx* ctx);
const char* agpc_last_error(agpc_ctx* ctx);
/* cap on the workspace arena (bytes); 0 = 80% of free device memory at first use */
int agpc_set_workspace_limit(agpc_ctx* ctx, uint64_t bytes);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
int64_t agpc_launch_count(agpc_ctx* ctx);

/* Optional per-launch event timing (bench.py's roofline; SURVEY 5 "tracing / profiling": the reference only has the
 * @timing wall-clock decorator, gpcexperiment.py:13-20).  enable(1) resets and starts bracketing every kernel launch
 * with a cudaEvent pair on its stream; read() synchronises and returns, per kernel class (AGPC_CLS_*), the summed
 * device time in ms, the ALGORITHMIC work attributed to the class (flops for GEMM, bytes otherwise) and the launch
 * count.  Arrays have AGPC_CLS_COUNT entries. */
#define AGPC_CLS_KBUILD 0
#define AGPC_CLS_GEMM 1
#define AGPC_CLS_POTF2 2
#define AGPC_CLS_MATVEC 3
#define AGPC_CLS_EP 4
#define AGPC_CLS_GRAD 5
#define AGPC_CLS_FORMB 6
#define AGPC_CLS_MISC 7
#define AGPC_CLS_GEMM_TRTRI 8 /* the tile GEMM inside trtri (AGPC_CLS_GEMM: inside potrf / predict) */
#define AGPC_CLS_GEMM_LAUUM 9 /* ... inside lauum */
#define AGPC_CLS_SLICE 10     /* operand slicing for the int8 tensor-core GEMM (bytes of FP64 read) */
#define AGPC_CLS_OZ_POTRF 11  /* the int8 tensor-core GEMM (csrc/ozaki.cu) inside potrf / trtri / lauum: its event time is  */
#define AGPC_CLS_OZ_TRTRI 12  /* ALSO part of AGPC_CLS_GEMM / _TRTRI / _LAUUM (whose work stays the algorithmic FP64 flop   */
#define AGPC_CLS_OZ_LAUUM 13  /* count of the phase); work here = int8 operations the kernel EXECUTED (2 x MACs)           */
#define AGPC_CLS_COUNT 14
int agpc_profile_enable(agpc_ctx* ctx, int32_t on);
int agpc_profile_read(agpc_ctx* ctx, double* ms, double* work, int64_t* launches);

/* ---- dataset: GPCData.X / .Y handed to GPy.models.GPClassification(X, Y, kernel) (gpckernel.py:126,245,322) - */
/* X: N x D row-major FP64; y: N labels, 1 -> +1, anything else (0 / -1) -> -1 (bernoulli.py).  Copied to device. */
int agpc_dataset_create(agpc_ctx* ctx, const double* X_host, const double* y_host, int32_t N, int32_t D,
                        int32_t* dataset_id);
int agpc_dataset_destroy(agpc_ctx* ctx, int32_t dataset_id);

/* ---- the hot entry: one GPClassification construction / parameters_changed() per item ------------------------
 * item b = (programs[b], theta[b*theta_stride ...], rows[row_off[b] .. row_off[b+1]))  -- the training rows are a
 * fold of the dataset (GPCData.kFoldSplits, gpcdata.py:149-180; all rows for add/toSummands, gpckernel.py:126,322).
 * For every item: build K, run parallel EP (cold or warm), factorise B = I + S^1/2 K S^1/2, and return
 *   nlml[b]        = -log Z_EP                          (m.log_likelihood(), gpckernel.py:371)
 *   grad[b*theta_stride + j] = d nlml / d theta_j       (constrained theta; paramz transforms stay on the host)
 *   bic[b]         = 2 nlml + p_eff ln n                (flexible_function.py:685-687)
 *   nlml_gauss[b]  = 2016-era GPy objective (Gaussian marginal of the pseudo-targets, sites frozen)
 *   sweeps[b], status[b]
 *   tau/nu [b*site_stride + i]: EP site parameters in/out (row order of the item's row list).
 * All array arguments are HOST pointers; the call copies in, runs on the context's stream, copies out and
 * synchronises.  Items are processed in lock-step chunks sized to the workspace arena. */
int agpc_score_batch(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                     const double* theta, int32_t theta_stride, const int32_t* rows, const int64_t* row_off,
                     const agpc_ep_options* opt, double* tau, double* nu, int32_t site_stride, double* nlml,
                     double* grad, double* bic, double* nlml_gauss, int32_t* sweeps, int32_t* status);

/* same, with theta / outputs resident on the device (no host round trip of the n-sized site vectors):
 * programs, rows (int32, device), row_off (host) as above; theta, tau, nu, nlml, grad, bic, nlml_gauss, sweeps,
 * status are DEVICE pointers.  Asynchronous on `stream` except for the per-sweep convergence poll. */
int agpc_score_batch_dev(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                         const double* theta_dev, int32_t theta_stride, const int32_t* rows_dev,
                         const int64_t* row_off, const agpc_ep_options* opt, double* tau_dev, double* nu_dev,
                         int32_t site_stride, double* nlml_dev, double* grad_dev, double* bic_dev,
                         double* nlml_gauss_dev, int32_t* sweeps_dev, int32_t* status_dev, void* stream);

/* ---- m.optimize() (gpckernel.py:246) for a whole batch, on the device ------------------------
 * paramz opt_lbfgsb = SciPy L-BFGS-B over the transformed parameters; here the limited-memory pairs, the Moré-Thuente
 * line search state, the iterates and the EP sites of every model stay in HBM, one round = one scoring pass over the
 * models still running + one kernel that advances every model to its next trial point.  theta (host, in/out): start
 * values in, optimum out.  kinds / lo / hi (host, [n_items][theta_stride]): the parameter's paramz constraint -- 0 free,
 * 1 positive (Logexp), 2 bounded (Logistic on [lo, hi]).  Outputs as agpc_score_batch at the optimum, plus per item the
 * number of objective evaluations, L-BFGS iterations and the reason it stopped (AGPC_OPT_*). */
typedef struct agpc_lbfgs_options {
  int32_t m;         /* limited-memory pairs, 1..10 (paramz: 10)                                   */
  int32_t max_evals; /* SciPy maxfun: stop at the first new iterate after this many evaluations    */
  int32_t maxiter;   /* SciPy maxiter (<= 0: 15000)                                                */
  int32_t maxls;     /* evaluations per line search before it is abandoned (SciPy: 20)             */
  double factr;      /* stop when (f_old - f) <= factr * eps * max(|f_old|, |f|, 1)  (paramz: 1e7) */
  double pgtol;      /* stop when max |g| <= pgtol                                   (paramz: 1e-5) */
} agpc_lbfgs_options;
#define AGPC_OPT_CONV_PGTOL 1
#define AGPC_OPT_CONV_FACTR 2
#define AGPC_OPT_STOP_MAXFUN 3
#define AGPC_OPT_STOP_MAXITER 4
#define AGPC_OPT_ABNORMAL_LINESEARCH 5
#define AGPC_OPT_TOO_MANY_FAILURES 6
int agpc_optimize_batch(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                        double* theta, int32_t theta_stride, const int32_t* kinds, const double* lo, const double* hi,
                        const int32_t* rows, const int64_t* row_off, const agpc_ep_options* ep,
                        const agpc_lbfgs_options* lb, double* tau, double* nu, int32_t site_stride, double* nlml,
                        double* grad, double* bic, double* nlml_gauss, int32_t* sweeps, int32_t* status,
                        int32_t* evals, int32_t* iters, int32_t* task);

/* ---- m.predict(XT) (gpckernel.py:832): p* = Phi(mu* / sqrt(1 + v*)) per test row ------------------------------
 * test rows are dataset rows `test_rows[test_off[b] .. test_off[b+1])`; p_out is laid out the same way.
 * Optionally also returns the latent mean / variance (m._raw_predict, gpcplot.py:110-114). */
int agpc_predict_batch(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                       const double* theta, int32_t theta_stride, const int32_t* rows, const int64_t* row_off,
                       const double* tau, const double* nu, int32_t site_stride, const int32_t* test_rows,
                       const int64_t* test_off, double* p_out, double* mu_out, double* var_out, int32_t* status);
/* predict at arbitrary points Xstar (m x D, host) for ONE item trained on `rows`; also d mu* / d x* (m x D) when
 * dmu_dx != NULL (m.predictive_gradients, gpckernel.py:431) */
int agpc_predict_points(agpc_ctx* ctx, int32_t dataset_id, const agpc_program* program, const double* theta,
                        const int32_t* rows, int32_t n_rows, const double* tau, const double* nu,
                        const double* Xstar_host, int32_t m, double* p_out, double* mu_out, double* var_out,
                        double* dmu_dx);

/* ---- building blocks, exported for the metric's micro-benchmarks and for unit parity tests (device pointers) -- */
/* kern.K(X) (+ materialised dK/dtheta_j when dK_dev != NULL: n_theta slabs of n x ld): G1/G2 of SURVEY 8a */
int agpc_build_K(agpc_ctx* ctx, const agpc_program* program, const double* theta_host, const double* X_dev,
                 int32_t n, int32_t D, double* K_dev, double* dK_dev, int64_t ld, void* stream);
/* fused gradient contraction grad_j = sum_ik G_ik dK_ik/dtheta_j without materialising dK (kern.update_gradients_full) */
int agpc_kernel_grad(agpc_ctx* ctx, const agpc_program* program, const double* theta_host, const double* X_dev,
                     int32_t n, int32_t D, const double* G_dev, int64_t ld, double* grad_host, void* stream);
/* util.linalg.jitchol (dpotrf, lower): in-place blocked Cholesky of `batch` n x n matrices (n % 128 == 0,
 * ld == n, matrix b at A_dev + b*n*n).  Also leaves Linv (lower, = dtrtri) in Minv_dev and its transpose in
 * Uinv_dev when those are non-NULL.  info_dev[b] = 0 or AGPC_ST_NOT_PD. */
int agpc_potrf(agpc_ctx* ctx, double* A_dev, int32_t n, int32_t batch, double* Minv_dev, double* Uinv_dev,
               int32_t* info_dev, void* stream);
/* util.linalg.jitchol(A, maxtries=5) as GPy defines it: agpc_potrf, and for every matrix whose plain factorisation
 * fails a retry on A + jitter I with jitter = mean(diag A) * 1e-6, times 10 per further failure, at most 5 jittered
 * attempts; a non-positive or non-finite diagonal fails at once.  jitter_dev[b] = the jitter that succeeded (0 = none
 * needed), info_dev[b] = 0 or AGPC_ST_NOT_PD.  The same ladder runs inside agpc_score_batch on B = I + S^1/2 K S^1/2
 * (status AGPC_ST_JITTER); AGPC_JITTER=0 in the environment switches that one off. */
int agpc_jitchol(agpc_ctx* ctx, double* A_dev, int32_t n, int32_t batch, double* jitter_dev, int32_t* info_dev, void* stream);
/* C = alpha * A * B^T + beta * C on the TMA + DMMA tile kernel (M, N % 128 == 0, K % 16 == 0, contiguous) */
int agpc_gemm_nt(agpc_ctx* ctx, const double* A_dev, const double* B_dev, double* C_dev, int32_t M, int32_t N,
                 int32_t K, double alpha, double beta, int32_t batch, void* stream);

/* The same product on the Blackwell tensor core proper (tcgen05.mma kind::i8, accumulators in TMEM): both operands are
 * cut into n_slices (6 or 7) signed 8-bit digit planes with one power-of-two scale per row, the integer products are
 * accumulated exactly in int32 and recombined in FP64 (csrc/ozaki.cu).  n_slices = 7 is accurate to the FP64 rounding
 * level of a DGEMM, 6 to ~1e-12 of |a_i|max |b_j|max K^1/2.  M, N % 128 == 0, K % 32 == 0, contiguous.  This is the
 * kernel behind the K = 512 trailing updates of agpc_potrf, the upper levels of its triangular inverse and lauum. */
int agpc_gemm_nt_i8x(agpc_ctx* ctx, const double* A_dev, const double* B_dev, double* C_dev, int32_t M, int32_t N,
                     int32_t K, double alpha, double beta, int32_t batch, int32_t n_slices, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUTOGPC_B200_H */
