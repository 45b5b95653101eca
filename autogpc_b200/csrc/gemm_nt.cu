// FP64 tile GEMM  C = alpha * A * B^T + beta * C  for the blocked Cholesky / trtri / lauum of B = I + S^1/2 K S^1/2.
//
// Replaces the dsyrk / dgemm / dtrsm / dtrtri / dpotri calls GPy reaches through util/linalg.py (jitchol, dtrtrs,
// pdinv, tdot) from EP.inference -- SURVEY 2.1 -- for gpckernel.py:245-246.
//
// sm_100a design: tcgen05 has no FP64 kind, so the FP64 tensor path on Blackwell is warp-level DMMA
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; measured issue ceiling 37.05 TFLOP/s on B200, tools/probe).  One CTA owns
// a 128 x 128 tile of C.  A dedicated producer warp streams 128 x 16 slices of both operands (both K-contiguous,
// "NT") into a 4-stage shared-memory ring with TMA (cp.async.bulk.tensor, 128-byte swizzle, mbarrier
// complete_tx); 16 consumer warps each own a 32 x 32 sub-tile (16 DMMA accumulators) and read their fragments
// bank-conflict-free straight out of the swizzled ring.  The k-index permutation inside a 16-wide slice
// (lane t reads k = 8*(t>>1) + 2*j + (t&1)) is what makes the 64-bit fragment loads conflict-free under the TMA
// swizzle; A and B use the same permutation so the contraction is unchanged.
//
// Triangular structure is exploited at tile granularity (skip upper tiles for syrk/lauum; clip the K range for the
// triangular operand of trtri/lauum); zeros inside diagonal 128-blocks are real zeros written by potf2_inv.
#include "common.cuh"

namespace agpc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// NJ = 4: 128 x 128 C tile (the throughput shape).  NJ = 1: 128 x 32 C tile for the thin, latency-bound launches of
// a single-matrix factorisation (panel solve, in-block panel update), which would otherwise leave most SMs idle;
// the B operand then comes through a tensor map with a 32-row box.  n_tiles / tj count BN-wide column tiles.
template <int NJ>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const GemmArgs args) {
  constexpr int BN = 32 * NJ;
  const int b = blockIdx.z;
  if (args.active != nullptr && args.active[b] == 0) return;
  const int g = blockIdx.y;
  // group geometry (trtri levels); for plain problems grp_blocks == 0 and everything is EXT_FIXED
  int h2 = 0;
  if (args.grp_blocks > 0) {
    h2 = args.nblk - g * args.grp_blocks - args.half_blocks;
    h2 = h2 < 0 ? 0 : (h2 > args.half_blocks ? args.half_blocks : h2);
  }
  const int shift = g * args.grp_blocks * TILE;
  const int mt = args.m_kind == EXT_H2 ? h2 : args.m_tiles;
  const int nt = args.n_kind == EXT_H2 ? h2 : args.n_tiles;
  const int kt = args.k_kind == EXT_H2 ? h2 * (TILE / GEMM_BK) : args.k_tiles;
  if (mt <= 0 || nt <= 0) return;
  int ti, tj;
  if (args.lower_only) {
    // linear index -> (ti, tj), tj <= ti
    const int t = blockIdx.x;
    if (t >= mt * (mt + 1) / 2) return;
    ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = blockIdx.x / nt;
    tj = blockIdx.x % nt;
    if (ti >= mt) return;
    if (args.skip_upper && ti * TILE < tj * BN) return;
  }
  int k_begin = 0, k_end = kt;
  if (args.k_rule == KRULE_BEGIN_AT_ROW) k_begin = ti * (TILE / GEMM_BK);
  if (args.k_rule == KRULE_END_AT_ROW) k_end = min(kt, (ti + 1) * (TILE / GEMM_BK));
  if (args.k_rule == KRULE_END_AT_COL) k_end = min(kt, (tj + 1) * (BN / GEMM_BK));
  const int n_iter = k_end - k_begin;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t STAGE_BYTES = TILE * GEMM_BK * 8;  // 16 KB per operand per stage (B uses BN / 128 of it)
  constexpr uint32_t B_BYTES = BN * GEMM_BK * 8;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + GEMM_STAGES * STAGE_BYTES;
  const uint32_t bars = sB + GEMM_STAGES * STAGE_BYTES;  // full[0..S), empty[0..S): 8 bytes each

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (GEMM_STAGES + s), GEMM_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == GEMM_CONSUMER_WARPS) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      const int a_row = args.a_row0 + shift + ti * TILE, a_col = args.a_col0 + shift;
      const int b_row = args.b_row0 + shift + tj * BN, b_col = args.b_col0 + shift;
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % GEMM_STAGES;
        const uint32_t ph = (uint32_t)(it / GEMM_STAGES) & 1u;
        mbar_wait(bars + 8 * (GEMM_STAGES + s), ph ^ 1u);
        const uint32_t full = bars + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES + B_BYTES);
        const int kc = (k_begin + it) * GEMM_BK;
        tma_load_3d(sA + s * STAGE_BYTES, &mapA, full, a_col + kc, a_row, b);
        tma_load_3d(sB + s * STAGE_BYTES, &mapB, full, b_col + kc, b_row, b);
      }
    }
    return;
  }

  // ---------------- DMMA consumers ----------------
  const int wm = warp >> 2, wn = warp & 3;  // 4 x 4 warps, 32 x 32 each
  const int gid = lane >> 2, tig = lane & 3;
  // byte offset of this lane's element inside a stage for 8-row block 0, k-step 0 (see header comment)
  const uint32_t lane_off = (uint32_t)gid * 128u + (uint32_t)(tig & 1) * 8u;
  const uint32_t chunk0 = ((uint32_t)((tig >> 1) << 2) ^ (uint32_t)gid) << 4;
  const uint32_t a_off = (uint32_t)(wm * 32) * 128u + lane_off;
  const uint32_t b_off = (uint32_t)(wn * 8 * NJ) * 128u + lane_off;

  // C is folded in up front (acc = (beta / alpha) C, result = alpha acc): its load latency then overlaps the TMA
  // prologue instead of sitting between the last DMMA and the store, which matters for the K = 128 rank updates.
  const long long row0 = (long long)args.c_row0 + shift + ti * TILE + wm * 32 + gid;
  const long long col0 = (long long)args.c_col0 + shift + tj * BN + wn * (8 * NJ) + tig * 2;
  double* Cb = args.C + (long long)b * args.c_batch;
  const double alpha = args.alpha;
  double acc[4][NJ][2];
  if (args.beta != 0.0) {
    const double pre = args.beta / alpha;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const double2 old = *reinterpret_cast<const double2*>(Cb + (row0 + i * 8) * args.ldc + col0 + j * 8);
        acc[i][j][0] = pre * old.x;
        acc[i][j][1] = pre * old.y;
      }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  for (int it = 0; it < n_iter; ++it) {
    const int s = it % GEMM_STAGES;
    const uint32_t ph = (uint32_t)(it / GEMM_STAGES) & 1u;
    mbar_wait(bars + 8 * s, ph);
    const uint32_t pa = sA + s * STAGE_BYTES + a_off;
    const uint32_t pb = sB + s * STAGE_BYTES + b_off;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const uint32_t ck = chunk0 ^ (uint32_t)(jp << 4);
      double af[4], bf[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = lds_f64(pa + i * 1024 + ck);
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = lds_f64(pb + j * 1024 + ck);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (GEMM_STAGES + s));
  }

  // ---------------- epilogue: registers -> global ----------------
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      double2 v;
      v.x = alpha * acc[i][j][0];
      v.y = alpha * acc[i][j][1];
      *reinterpret_cast<double2*>(Cb + (row0 + i * 8) * args.ldc + col0 + j * 8) = v;
      acc[i][j][0] = v.x;
      acc[i][j][1] = v.y;
    }
  }
  if (args.C2 != nullptr && tj < args.c2_cols && ti >= args.c2_row_min) {
    double* C2b = args.C2 + (long long)b * args.c_batch;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        *reinterpret_cast<double2*>(C2b + (row0 + i * 8) * args.ldc + col0 + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
  }
  if (args.Ct != nullptr) {
    const long long trow0 = (long long)args.ct_row0 + shift + tj * BN + wn * (8 * NJ) + tig * 2;
    const long long tcol0 = (long long)args.ct_col0 + shift + ti * TILE + wm * 32 + gid;
    double* Tb = args.Ct + (long long)b * args.c_batch;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        Tb[(trow0 + j * 8) * args.ldc + tcol0 + i * 8] = acc[i][j][0];
        Tb[(trow0 + j * 8 + 1) * args.ldc + tcol0 + i * 8] = acc[i][j][1];
      }
  }
}

cudaError_t gemm_nt_launch(const Launch& L, const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmArgs& args,
                           int tiles_x, int groups, int batch) {
  ProfScope prof_scope(L, L.gemm_cls);
  bool& configured = func_attr_done(__LINE__);  // per device: the attribute belongs to the device's context
  if (!configured) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    AGPC_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    configured = true;
  }
  if (tiles_x <= 0 || groups <= 0 || batch <= 0) return cudaSuccess;
  dim3 grid(tiles_x, groups, batch);
  gemm_nt_kernel<4><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, L.stream>>>(mapA, mapB, args);
  L.bump();
  return cudaGetLastError();
}

// thin variant: args.n_tiles counts 32-wide column tiles; mapB must have a 32-row box (make_tensor_map box_rows = 32)
cudaError_t gemm_nt_thin_launch(const Launch& L, const CUtensorMap& mapA, const CUtensorMap& mapB32,
                                const GemmArgs& args, int tiles_x, int groups, int batch) {
  ProfScope prof_scope(L, L.gemm_cls);
  bool& configured = func_attr_done(__LINE__);  // per device: the attribute belongs to the device's context
  if (!configured) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    configured = true;
  }
  if (tiles_x <= 0 || groups <= 0 || batch <= 0) return cudaSuccess;
  dim3 grid(tiles_x, groups, batch);
  gemm_nt_kernel<1><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, L.stream>>>(mapA, mapB32, args);
  L.bump();
  return cudaGetLastError();
}

// ---- tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tensor_map(CUtensorMap* out, const double* base, long long cols, long long rows, long long batch,
                    long long ld, long long batch_stride, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return AGPC_E_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 8ull, (cuuint64_t)batch_stride * 8ull};
  if (batch == 1) strides[1] = (cuuint64_t)ld * 8ull * (cuuint64_t)rows;
  cuuint32_t box[3] = {GEMM_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AGPC_OK : AGPC_E_CUDA;
}

}  // namespace agpc
