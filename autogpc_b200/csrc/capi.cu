// C ABI of the engine (include/autogpc_b200.h): context, workspace arena, datasets, building-block entry points.
// The scoring pipeline itself is in score.cu.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <cstring>
#include <new>

#include "engine.h"

using namespace agpc;

static int fail(agpc_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (ctx) {
    char buf[512];
    if (e != cudaSuccess)
      snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else
      snprintf(buf, sizeof buf, "%s", what);
    ctx->err = buf;
  }
  return code;
}

#define CK(ctx, expr)                                                  \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) return fail(ctx, AGPC_E_CUDA, #expr, _e);   \
  } while (0)

int agpc_ctx::ensure_arena(size_t bytes) {
  if (bytes <= arena_bytes) return AGPC_OK;
  if (arena) {
    cudaStreamSynchronize(stream);
    cudaFree(arena);
    arena = nullptr;
    arena_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&arena, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(this, AGPC_E_NOMEM, "workspace arena cudaMalloc", e);
  }
  arena_bytes = bytes;
  return AGPC_OK;
}

size_t agpc_ctx::workspace_budget() {
  if (limit) return limit;
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return (size_t)8 << 30;
  return (size_t)((double)(fr + arena_bytes) * 0.80);
}

extern "C" {

int agpc_version(void) { return AGPC_VERSION; }

int agpc_create(int device, agpc_ctx** out) {
  if (!out) return AGPC_E_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return AGPC_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return AGPC_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AGPC_E_CUDA;
  if (prop.major != 10) return AGPC_E_UNSUPPORTED;  // sm_100a only: no fallback path exists
  agpc_ctx* c = new (std::nothrow) agpc_ctx();
  if (!c) return AGPC_E_NOMEM;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("AGPC_JITTER")) c->jitter = atoi(e) != 0;
  if (const char* e = getenv("AGPC_OZAKI")) {
    const int v = atoi(e);
    c->oz_S = (v == 6 || v == 7) ? v : 0;
  }
  if (const char* e = getenv("AGPC_OZ_MIN_N")) c->oz_min_np = atoi(e);
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // the look-ahead (panel) stream outranks the bulk updates
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_panel, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_rest, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return AGPC_E_CUDA;
  }
  *out = c;
  return AGPC_OK;
}

int agpc_destroy(agpc_ctx* ctx) {
  if (!ctx) return AGPC_E_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& d : ctx->datasets) {
    if (d.X) cudaFree(d.X);
    if (d.y) cudaFree(d.y);
  }
  for (auto& d : ctx->dataset_pool) {
    cudaFree(d.X);
    cudaFree(d.y);
  }
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->opt_stage) cudaFree(ctx->opt_stage);
  if (ctx->prof.base) cudaEventDestroy(ctx->prof.base);
  for (auto& r : ctx->prof.recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  cudaStreamSynchronize(ctx->side);
  cudaEventDestroy(ctx->ev_panel);
  cudaEventDestroy(ctx->ev_rest);
  agpc::oz_release_stream(ctx->side);
  agpc::oz_release_stream(ctx->stream);
  cudaStreamDestroy(ctx->side);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AGPC_OK;
}

const char* agpc_last_error(agpc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int agpc_set_workspace_limit(agpc_ctx* ctx, uint64_t bytes) {
  if (!ctx) return AGPC_E_ARG;
  ctx->limit = (size_t)bytes;
  return AGPC_OK;
}

int64_t agpc_launch_count(agpc_ctx* ctx) { return ctx ? ctx->launches : -1; }

int agpc_profile_enable(agpc_ctx* ctx, int32_t on) {
  if (!ctx) return AGPC_E_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  Profiler& p = ctx->prof;
  p.enabled = on != 0;
  p.used = 0;
  if (p.base == nullptr && cudaEventCreate(&p.base) != cudaSuccess) return fail(ctx, AGPC_E_CUDA, "profile: event create");
  if (on) {
    cudaEventRecord(p.base, ctx->stream);
    cudaEventSynchronize(p.base);
  }
  for (int c = 0; c < CLS_COUNT; ++c) {
    p.work[c] = 0.0;
    p.count[c] = 0;
  }
  return AGPC_OK;
}

int agpc_profile_read(agpc_ctx* ctx, double* ms, double* work, int64_t* launches) {
  if (!ctx || !ms || !work || !launches) return AGPC_E_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaDeviceSynchronize());
  Profiler& p = ctx->prof;
  for (int c = 0; c < CLS_COUNT; ++c) {
    ms[c] = 0.0;
    work[c] = p.work[c];
    launches[c] = p.count[c];
  }
  // per class: length of the UNION of its launch intervals (launches of one class may overlap when the look-ahead
  // stream and the bulk stream run the same kernel concurrently; a plain sum would count that time twice)
  std::vector<std::pair<float, float>> iv[CLS_COUNT];
  for (size_t i = 0; i < p.used; ++i) {
    float t0 = 0.f, t1 = 0.f;
    CK(ctx, cudaEventElapsedTime(&t0, p.base, p.recs[i].e0));
    CK(ctx, cudaEventElapsedTime(&t1, p.base, p.recs[i].e1));
    iv[p.recs[i].cls].push_back({t0, t1});
  }
  for (int c = 0; c < CLS_COUNT; ++c) {
    std::sort(iv[c].begin(), iv[c].end());
    double total = 0.0;
    float cur0 = 0.f, cur1 = -1.f;
    for (auto& x : iv[c]) {
      if (cur1 < cur0 || x.first > cur1) {
        if (cur1 >= cur0) total += cur1 - cur0;
        cur0 = x.first;
        cur1 = x.second;
      } else if (x.second > cur1) {
        cur1 = x.second;
      }
    }
    if (cur1 >= cur0) total += cur1 - cur0;
    ms[c] = total;
  }
  return AGPC_OK;
}

int agpc_dataset_create(agpc_ctx* ctx, const double* X_host, const double* y_host, int32_t N, int32_t D,
                        int32_t* dataset_id) {
  if (!ctx || !X_host || !y_host || N <= 0 || D <= 0 || !dataset_id) return fail(ctx, AGPC_E_ARG, "dataset_create args");
  CK(ctx, cudaSetDevice(ctx->device));
  Dataset d;
  d.N = N;
  d.D = D;
  for (size_t i = 0; i < ctx->dataset_pool.size(); ++i)
    if (ctx->dataset_pool[i].N == N && ctx->dataset_pool[i].D == D) {
      d = ctx->dataset_pool[i];
      ctx->dataset_pool.erase(ctx->dataset_pool.begin() + i);
      break;
    }
  if (!d.X) {
    CK(ctx, cudaMalloc(&d.X, sizeof(double) * (size_t)N * D));
    if (cudaMalloc(&d.y, sizeof(double) * (size_t)N) != cudaSuccess) {
      cudaFree(d.X);
      return fail(ctx, AGPC_E_NOMEM, "dataset_create cudaMalloc");
    }
  }
  CK(ctx, cudaMemcpyAsync(d.X, X_host, sizeof(double) * (size_t)N * D, cudaMemcpyHostToDevice, ctx->stream));
  // labels: 1 -> +1, everything else -> -1 (GPy bernoulli.py moments_match_ep)
  std::vector<double> ypm(N);
  for (int i = 0; i < N; ++i) ypm[i] = (y_host[i] == 1.0) ? 1.0 : -1.0;
  CK(ctx, cudaMemcpyAsync(d.y, ypm.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  // reuse a free slot if any
  for (size_t i = 0; i < ctx->datasets.size(); ++i)
    if (ctx->datasets[i].X == nullptr) {
      ctx->datasets[i] = d;
      *dataset_id = (int32_t)i;
      return AGPC_OK;
    }
  ctx->datasets.push_back(d);
  *dataset_id = (int32_t)ctx->datasets.size() - 1;
  return AGPC_OK;
}

int agpc_dataset_destroy(agpc_ctx* ctx, int32_t id) {
  if (!ctx || id < 0 || id >= (int)ctx->datasets.size() || !ctx->datasets[id].X) return fail(ctx, AGPC_E_ARG, "bad dataset id");
  cudaSetDevice(ctx->device);
  // the buffers go to the pool instead of cudaFree: keep cudaFree's implicit device-wide synchronisation, so that work
  // a caller queued on its own stream has finished reading the dataset before a later create re-uses the memory
  cudaDeviceSynchronize();
  if (ctx->dataset_pool.size() >= 4) {
    cudaFree(ctx->dataset_pool.front().X);
    cudaFree(ctx->dataset_pool.front().y);
    ctx->dataset_pool.erase(ctx->dataset_pool.begin());
  }
  ctx->dataset_pool.push_back(ctx->datasets[id]);
  ctx->datasets[id] = Dataset();
  return AGPC_OK;
}

int agpc_gemm_nt(agpc_ctx* ctx, const double* A, const double* B, double* C, int32_t M, int32_t N, int32_t K,
                 double alpha, double beta, int32_t batch, void* stream) {
  if (!ctx || !A || !B || !C || M <= 0 || N <= 0 || K <= 0 || batch <= 0 || M % TILE || N % TILE || K % GEMM_BK)
    return fail(ctx, AGPC_E_ARG, "gemm_nt: M,N must be multiples of 128 and K of 16");
  CK(ctx, cudaSetDevice(ctx->device));
  CUtensorMap ma, mb;
  if (make_tensor_map(&ma, A, K, M, batch, K, (long long)M * K) || make_tensor_map(&mb, B, K, N, batch, K, (long long)N * K))
    return fail(ctx, AGPC_E_CUDA, "cuTensorMapEncodeTiled failed");
  GemmArgs g{};
  g.C = C; g.ldc = N; g.c_batch = (long long)M * N;
  g.m_tiles = M / TILE; g.n_tiles = N / TILE; g.k_tiles = K / GEMM_BK;
  g.alpha = alpha; g.beta = beta;
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  CK(ctx, gemm_nt_launch(L, ma, mb, g, g.m_tiles * g.n_tiles, 1, batch));
  return AGPC_OK;
}

int agpc_gemm_nt_i8x(agpc_ctx* ctx, const double* A, const double* B, double* C, int32_t M, int32_t N, int32_t K,
                     double alpha, double beta, int32_t batch, int32_t n_slices, void* stream) {
  if (!ctx || !A || !B || !C || M <= 0 || N <= 0 || K <= 0 || batch <= 0 || M % TILE || N % TILE || K % OZ_KB)
    return fail(ctx, AGPC_E_ARG, "gemm_nt_i8x: M,N must be multiples of 128 and K of 32");
  if (n_slices != 6 && n_slices != 7) return fail(ctx, AGPC_E_ARG, "gemm_nt_i8x: n_slices must be 6 or 7");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t sa_bytes = (size_t)n_slices * M * K, sb_bytes = (size_t)n_slices * N * K;
  const size_t need = (sa_bytes + sb_bytes) * batch + sizeof(double) * (size_t)(M + N) * batch + 8192;
  int rc = ctx->ensure_arena(need);
  if (rc) return rc;
  char* p = (char*)ctx->arena;
  int8_t* SA = (int8_t*)p; p += (sa_bytes * batch + 1023) / 1024 * 1024;
  int8_t* SB = (int8_t*)p; p += (sb_bytes * batch + 1023) / 1024 * 1024;
  double* scA = (double*)p; p += sizeof(double) * (size_t)M * batch;
  double* scB = (double*)p;
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  OzSliceArgs sa{};
  sa.X = A; sa.ld = K; sa.x_batch = (long long)M * K; sa.S = SA; sa.scale = scA; sa.s_batch = (long long)sa_bytes; sa.sc_batch = M;
  sa.nkc = K / OZ_KB; sa.row_blocks = M / TILE; sa.col_chunks = K / OZ_KB;
  CK(ctx, oz_slice_launch(L, sa, M / TILE, 1, batch, n_slices));
  OzSliceArgs sb = sa;
  sb.X = B; sb.x_batch = (long long)N * K; sb.S = SB; sb.scale = scB; sb.s_batch = (long long)sb_bytes; sb.sc_batch = N;
  sb.row_blocks = N / TILE;
  CK(ctx, oz_slice_launch(L, sb, N / TILE, 1, batch, n_slices));
  OzGemmArgs g{};
  g.A = OzOperand{SA, scA, (long long)sa_bytes, M, K / OZ_KB};
  g.B = OzOperand{SB, scB, (long long)sb_bytes, N, K / OZ_KB};
  g.C = C; g.ldc = N; g.c_batch = (long long)M * N;
  g.m_tiles = M / TILE; g.n_tiles = N / TILE; g.k_chunks = K / OZ_KB;
  g.alpha = alpha; g.beta = beta;
  L.add_work(CLS_GEMM, 2.0 * M * N * (double)K * batch);
  L.add_work(CLS_SLICE, 8.0 * ((double)M + N) * K * batch);
  CK(ctx, oz_gemm_launch(L, g, g.m_tiles * g.n_tiles, 1, batch, n_slices));
  return AGPC_OK;
}

int agpc_potrf(agpc_ctx* ctx, double* A, int32_t n, int32_t batch, double* Minv, double* Uinv, int32_t* info,
               void* stream) {
  if (!ctx || !A || n <= 0 || batch <= 0 || n % TILE) return fail(ctx, AGPC_E_ARG, "potrf: n must be a multiple of 128");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t mat = sizeof(double) * (size_t)n * n * batch;
  const bool want_inv = (Minv != nullptr) || (Uinv != nullptr);
  size_t need = 0;
  if (!Minv) need += mat;
  if (!Uinv) need += mat;
  if (want_inv) need += mat;  // T
  const int oz_S = ctx->oz_S;
  const size_t oz_bytes = (size_t)oz_S * n * n * batch, sc_bytes = sizeof(double) * (size_t)n * batch;
  if (oz_S) need += 3 * (oz_bytes + 1024) + 6 * sc_bytes;
  int rc = ctx->ensure_arena(need + 4096);
  if (rc) return rc;
  char* p = (char*)ctx->arena;
  MatSet ms;
  ms.np = n; ms.batch = batch; ms.stride = (long long)n * n;
  ms.A = A;
  ms.M = Minv ? Minv : (double*)p; if (!Minv) p += mat;
  ms.U = Uinv ? Uinv : (double*)p; if (!Uinv) p += mat;
  ms.T = want_inv ? (double*)p : nullptr;
  if (want_inv) p += mat;
  if (oz_S) {
    ms.oz_S = oz_S;
    ms.scA = (double*)p; p += sc_bytes;
    ms.scB = (double*)p; p += sc_bytes;
    ms.scC = (double*)p; p += sc_bytes;
    ms.rmax = (double*)p; p += sc_bytes;
    ms.rmM = (double*)p; p += sc_bytes;
    ms.rmU = (double*)p; p += sc_bytes;
    p = (char*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
    ms.SA = (int8_t*)p; p += (oz_bytes + 1023) / 1024 * 1024;
    ms.SB = (int8_t*)p; p += (oz_bytes + 1023) / 1024 * 1024;
    ms.SC = (int8_t*)p;
  }
  if (make_tensor_map(&ms.mapA, ms.A, n, n, batch, n, ms.stride) || make_tensor_map(&ms.mapM, ms.M, n, n, batch, n, ms.stride) ||
      make_tensor_map(&ms.mapU, ms.U, n, n, batch, n, ms.stride) ||
      make_tensor_map(&ms.mapA32, ms.A, n, n, batch, n, ms.stride, 32) ||
      make_tensor_map(&ms.mapM32, ms.M, n, n, batch, n, ms.stride, 32) ||
      (want_inv && make_tensor_map(&ms.mapT, ms.T, n, n, batch, n, ms.stride)))
    return fail(ctx, AGPC_E_CUDA, "cuTensorMapEncodeTiled failed");
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  if (info) CK(ctx, cudaMemsetAsync(info, 0, sizeof(int32_t) * batch, L.stream));
  CK(ctx, potrf_blocked(L, ms, nullptr, info));
  if (want_inv) CK(ctx, trtri_blocked(L, ms, nullptr));
  return AGPC_OK;
}


int agpc_jitchol(agpc_ctx* ctx, double* A, int32_t n, int32_t batch, double* jitter_out, int32_t* info_out, void* stream) {
  if (!ctx || !A || n <= 0 || batch <= 0 || n % TILE) return fail(ctx, AGPC_E_ARG, "jitchol: n must be a multiple of 128");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t mat = sizeof(double) * (size_t)n * n * batch;
  const int oz_S = ctx->oz_S;
  const size_t oz_bytes = (size_t)oz_S * n * n * batch, sc_bytes = sizeof(double) * (size_t)n * batch;
  size_t need = 3 * mat + (oz_S ? 2 * (oz_bytes + 1024) + 2 * sc_bytes : 0) + (sizeof(double) + 3 * sizeof(int)) * (size_t)batch + 16384;
  int rc = ctx->ensure_arena(need);
  if (rc) return rc;
  char* p = (char*)ctx->arena;
  MatSet ms;
  ms.np = n; ms.batch = batch; ms.stride = (long long)n * n;
  ms.A = A;
  ms.M = (double*)p; p += mat;
  ms.U = (double*)p; p += mat;
  double* A0 = (double*)p; p += mat;
  double* jit = (double*)p; p += sizeof(double) * (size_t)batch;
  int* info = (int*)p; p += sizeof(int) * (size_t)batch;
  int* tries = (int*)p; p += sizeof(int) * (size_t)batch;
  int* retry = (int*)p; p += sizeof(int) * (size_t)batch;
  if (oz_S) {
    p = (char*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
    ms.oz_S = oz_S;
    ms.scA = (double*)p; p += sc_bytes;
    ms.scB = (double*)p; p += sc_bytes;
    p = (char*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
    ms.SA = (int8_t*)p; p += (oz_bytes + 1023) / 1024 * 1024;
    ms.SB = (int8_t*)p;
  }
  if (make_tensor_map(&ms.mapA, ms.A, n, n, batch, n, ms.stride) || make_tensor_map(&ms.mapM, ms.M, n, n, batch, n, ms.stride) ||
      make_tensor_map(&ms.mapU, ms.U, n, n, batch, n, ms.stride) ||
      make_tensor_map(&ms.mapA32, ms.A, n, n, batch, n, ms.stride, 32) ||
      make_tensor_map(&ms.mapM32, ms.M, n, n, batch, n, ms.stride, 32))
    return fail(ctx, AGPC_E_CUDA, "cuTensorMapEncodeTiled failed");
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  CK(ctx, cudaMemcpyAsync(A0, A, mat, cudaMemcpyDeviceToDevice, L.stream));
  CK(ctx, cudaMemsetAsync(info, 0, sizeof(int) * batch, L.stream));
  CK(ctx, cudaMemsetAsync(tries, 0, sizeof(int) * batch, L.stream));
  CK(ctx, cudaMemsetAsync(jit, 0, sizeof(double) * batch, L.stream));
  CK(ctx, potrf_blocked(L, ms, nullptr, info));
  std::vector<int> h(batch);
  for (int attempt = 0; attempt <= 5; ++attempt) {
    CK(ctx, cudaMemcpyAsync(h.data(), info, sizeof(int) * batch, cudaMemcpyDeviceToHost, L.stream));
    CK(ctx, cudaStreamSynchronize(L.stream));
    int n_bad = 0;
    for (int b = 0; b < batch; ++b) n_bad += h[b] == AGPC_ST_NOT_PD;
    if (n_bad == 0 || attempt == 5) break;
    CK(ctx, jitter_step_mat_launch(L, A0, A, n, ms.stride, batch, info, jit, tries, retry));
    CK(ctx, potrf_blocked(L, ms, retry, info));
  }
  if (jitter_out) CK(ctx, cudaMemcpyAsync(jitter_out, jit, sizeof(double) * batch, cudaMemcpyDeviceToDevice, L.stream));
  if (info_out) CK(ctx, cudaMemcpyAsync(info_out, info, sizeof(int) * batch, cudaMemcpyDeviceToDevice, L.stream));
  return AGPC_OK;
}

// small persistent device scratch for a program + theta (building-block entry points)
static int param_scratch(agpc_ctx* ctx, const agpc_program* prog, const double* theta_host, cudaStream_t st,
                         agpc_program** d_prog, double** d_theta) {
  if (!ctx->scratch) {
    if (cudaMalloc(&ctx->scratch, 1 << 16) != cudaSuccess) return fail(ctx, AGPC_E_NOMEM, "scratch cudaMalloc");
  }
  *d_prog = (agpc_program*)ctx->scratch;
  *d_theta = (double*)((char*)ctx->scratch + 4096);
  if (prog->n_theta <= 0 || prog->n_theta > AGPC_MAX_THETA || prog->n_leaves <= 0 || prog->n_leaves > AGPC_MAX_LEAVES ||
      prog->n_terms <= 0 || prog->n_terms > AGPC_MAX_TERMS)
    return fail(ctx, AGPC_E_ARG, "malformed kernel program");
  CK(ctx, cudaMemcpyAsync(*d_prog, prog, sizeof(agpc_program), cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(*d_theta, theta_host, sizeof(double) * prog->n_theta, cudaMemcpyHostToDevice, st));
  return AGPC_OK;
}

int agpc_build_K(agpc_ctx* ctx, const agpc_program* program, const double* theta_host, const double* X_dev, int32_t n,
                 int32_t D, double* K_dev, double* dK_dev, int64_t ld, void* stream) {
  if (!ctx || !program || !theta_host || !X_dev || !K_dev || n <= 0 || D <= 0 || ld < n || (ld & 1))
    return fail(ctx, AGPC_E_ARG, "build_K: bad argument (ld must be even and >= n)");
  CK(ctx, cudaSetDevice(ctx->device));
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  agpc_program* d_prog;
  double* d_theta;
  int rc = param_scratch(ctx, program, theta_host, L.stream, &d_prog, &d_theta);
  if (rc) return rc;
  BuildArgs a{};
  a.progs = d_prog; a.theta = d_theta; a.theta_stride = AGPC_MAX_THETA;
  a.Xa = X_dev; a.Xb = X_dev; a.na_fixed = n; a.nb_fixed = n; a.na_pad = n; a.nb_pad = n; a.D = D;
  a.out = K_dev; a.ld = ld; a.out_batch = (long long)n * ld; a.dK = dK_dev;
  {  // the streaming dK kernel writes each parameter slab once: needs every leaf in exactly one term
    int uses[AGPC_MAX_LEAVES] = {0};
    bool unique = true;
    for (int t = 0; t < program->n_terms && unique; ++t) {
      if (program->term_len[t] <= 0 || program->term_len[t] > AGPC_MAX_TERM_LEN) return fail(ctx, AGPC_E_ARG, "build_K: bad term");
      for (int q = 0; q < program->term_len[t]; ++q) {
        const int l = program->term_leaf[t][q];
        if (l < 0 || l >= program->n_leaves) return fail(ctx, AGPC_E_ARG, "build_K: bad term leaf");
        if (++uses[l] > 1) unique = false;
      }
    }
    a.dk_tiled = unique ? 1 : 0;
  }
  CK(ctx, build_K_launch(L, a, 1, dK_dev != nullptr));
  return AGPC_OK;
}

int agpc_kernel_grad(agpc_ctx* ctx, const agpc_program* program, const double* theta_host, const double* X_dev,
                     int32_t n, int32_t D, const double* G_dev, int64_t ld, double* grad_host, void* stream) {
  if (!ctx || !program || !theta_host || !X_dev || !G_dev || !grad_host || n <= 0 || D <= 0 || ld < n)
    return fail(ctx, AGPC_E_ARG, "kernel_grad: bad argument");
  CK(ctx, cudaSetDevice(ctx->device));
  Launch L{stream ? (cudaStream_t)stream : ctx->stream, &ctx->launches, &ctx->prof, ctx->side, ctx->ev_panel, ctx->ev_rest};
  agpc_program* d_prog;
  double* d_theta;
  int rc = param_scratch(ctx, program, theta_host, L.stream, &d_prog, &d_theta);
  if (rc) return rc;
  const int np = (n + 63) / 64 * 64;
  const int tiles = grad_tiles(np, 1);
  const size_t need = sizeof(double) * ((size_t)np * D + (size_t)tiles * AGPC_MAX_THETA + AGPC_MAX_THETA) + 8192;
  rc = ctx->ensure_arena(need);
  if (rc) return rc;
  double* Xp = (double*)ctx->arena;
  double* partial = Xp + (((size_t)np * D + 127) / 128) * 128;
  double* g_dev = partial + (size_t)tiles * AGPC_MAX_THETA;
  CK(ctx, cudaMemsetAsync(Xp, 0, sizeof(double) * (size_t)np * D, L.stream));
  CK(ctx, cudaMemcpyAsync(Xp, X_dev, sizeof(double) * (size_t)n * D, cudaMemcpyDeviceToDevice, L.stream));
  GradArgs ga{};
  ga.progs = d_prog; ga.theta = d_theta; ga.theta_stride = AGPC_MAX_THETA; ga.Xf = Xp; ga.n_fixed = n; ga.np = np; ga.D = D;
  ga.G = G_dev; ga.ld = ld; ga.g_batch = 0; ga.mode = 1; ga.partial = partial; ga.tiles_per_item = tiles;
  CK(ctx, grad_contract_launch(L, ga, 1, 1.0, g_dev, AGPC_MAX_THETA));
  CK(ctx, cudaMemcpyAsync(grad_host, g_dev, sizeof(double) * program->n_theta, cudaMemcpyDeviceToHost, L.stream));
  CK(ctx, cudaStreamSynchronize(L.stream));
  return AGPC_OK;
}

}  // extern "C"
