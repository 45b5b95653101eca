// Host-side engine state behind the opaque agpc_ctx handle.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

struct Dataset {
  double* X = nullptr;  // N x D row-major (device)
  double* y = nullptr;  // N, +-1 (device)
  int N = 0, D = 0;
};

struct agpc_ctx {
  int device = 0;
  int sm_count = 148;
  int jitter = 1;  // jitchol's jitter ladder on a failed factorisation (AGPC_JITTER=0 disables: failures stay NOT_PD)
  int oz_S = 7;  // digit slices of the int8 tensor-core GEMM path (ozaki.cu); 0 = FP64 DMMA kernel only.  AGPC_OZAKI=0|6|7
  int oz_min_np = 1536;  // scoring path: padded matrix size from which the int8 path is used (AGPC_OZ_MIN_N).  Below it the
                         // slicing passes cost more than the tensor core saves: 4096 models at n = 1024 run 8 % faster on the
                         // FP64 DMMA kernel (0.75 s against 0.81 s), from n = 1536 on the int8 path wins by 20 % and more
                         // (profiles/r02_summary.md section 3).  The building-block entry points are not affected.
  int oz_slices(int np) const { return np >= oz_min_np ? oz_S : 0; }
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // look-ahead stream of potrf_blocked
  cudaEvent_t ev_panel = nullptr, ev_rest = nullptr;
  std::string err;
  long long launches = 0;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  size_t limit = 0;
  void* pinned = nullptr;  // small pinned staging buffer for per-sweep flags
  size_t pinned_bytes = 0;
  void* scratch = nullptr;  // 64 KB device scratch: program + theta of the building-block entry points
  void* stage = nullptr;    // grow-only device staging of the host-buffer entry point (cudaMalloc / cudaFree per
  size_t stage_bytes = 0;   // call cost 0.5-350 ms on a box with > 100 GB allocated: profiles/r01_summary.md)
  void* opt_stage = nullptr;  // the same for agpc_optimize_batch (optimiser state, compact staging of the running subset)
  size_t opt_stage_bytes = 0;
  std::vector<Dataset> datasets;
  std::vector<Dataset> dataset_pool;  // buffers of destroyed datasets kept for re-use (at most 4): no cudaMalloc /
                                      // cudaFree when a caller re-uploads a same-shaped dataset per step
  agpc::Profiler prof;

  int ensure_arena(size_t bytes);
  size_t workspace_budget();
};

// score.cu: the scoring path with explicit per-item row lists (agpc_score_batch_dev = adjacent lists); used by optim.cu
int score_items_dev(agpc_ctx* ctx, int32_t dataset_id, int32_t n_items, const agpc_program* programs,
                    const double* theta_dev, int32_t theta_stride, const int32_t* rows_dev, const int64_t* row_start,
                    const int32_t* row_len, const agpc_ep_options* opt, double* tau_dev, double* nu_dev,
                    int32_t site_stride, double* nlml_dev, double* grad_dev, double* bic_dev, double* nlml_gauss_dev,
                    int32_t* sweeps_dev, int32_t* status_dev, void* stream);
