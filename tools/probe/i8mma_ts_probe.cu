// Probe 2 for the sliced-integer GEMM: A operand of tcgen05.mma kind::i8 read from TENSOR MEMORY (filled by
// tcgen05.cp 128x256b from the swizzled shared-memory tile) instead of shared memory.  The 128 x 64 x 32 MMA is
// shared-memory-bound at 48 cycles (4 KB of A + 2 KB of B per instruction at 128 B/clk); with A in TMEM only B is read.
// Questions: (1) does cp.128x256b with the MMA's own K-major 32-byte-swizzle descriptor land row r in lane r, 32 bytes in 8
// columns, as the TS-mode MMA expects; (2) is cp -> mma -> cp(overwrite) ordered when issued back to back by one thread
// (single-buffered A) or does A need double buffering; (3) the MMA rate.
//   NT  = number of A slices (of S) kept in TMEM, the rest stay SS-mode;  DB = 1: two TMEM buffers alternate per k-step.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  } while (!done);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
               ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_cp(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
constexpr int KB = 32, N = 64;
__host__ __device__ inline uint32_t swz_off(int row, int kbyte) {
  const uint32_t a = (uint32_t)row * KB + kbyte;
  return a ^ (((a >> 7) & 1u) << 4);
}
__host__ __device__ inline uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((8 * KB) >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
__host__ __device__ inline uint32_t make_idesc() {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// whole problem resident in shared memory: KS k-steps x (S A-tiles of 4 KB + S B-tiles of 2 KB); every MMA / cp is issued
// back to back by one thread with no wait in between (REPS > 1 repeats the whole sequence for the rate measurement)
template <int S, int NT, int DB, int KS, int MODE = 0>
__global__ void __launch_bounds__(128, 1)
ts_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ out, int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  constexpr int A_T = 128 * KB, B_T = N * KB, STEP = S * (A_T + B_T);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int Ktot = KS * KB;
  for (int idx = tid; idx < KS * S * 128 * KB; idx += 128) {
    const int ks = idx / (S * 128 * KB), p = (idx / (128 * KB)) % S, r = (idx / KB) % 128, c = idx % KB;
    sm[ks * STEP + p * A_T + swz_off(r, c)] = (uint8_t)A[((size_t)p * 128 + r) * Ktot + ks * KB + c];
  }
  for (int idx = tid; idx < KS * S * N * KB; idx += 128) {
    const int ks = idx / (S * N * KB), q = (idx / (N * KB)) % S, r = (idx / KB) % N, c = idx % KB;
    sm[ks * STEP + S * A_T + q * B_T + swz_off(r, c)] = (uint8_t)B[((size_t)q * N + r) * Ktot + ks * KB + c];
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc();
  const uint32_t ta_base = tmem + S * N;  // A buffers right after the S accumulators
  if (tid == 0) {
    if (MODE == 1)
      for (int p = 0; p < NT; ++p) tmem_cp(ta_base + p * 8, make_desc(base + p * A_T));
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t sa = base + ks * STEP, sb = sa + S * A_T;
        const uint32_t ta = ta_base + (DB ? (uint32_t)(ks & 1) * NT * 8 : 0u);
        if (MODE != 1)
          for (int p = 0; p < NT; ++p) tmem_cp(ta + p * 8, make_desc(sa + p * A_T));
        if (MODE != 2)
        for (int p = 0; p < S; ++p)
          for (int q = 0; p + q < S; ++q) {
            const uint64_t db = make_desc(sb + q * B_T);
            const uint32_t acc = (rep == 0 && ks == 0 && p == 0) ? 0u : 1u;
            if (p < NT) umma_ts(tmem + (uint32_t)(p + q) * N, ta + p * 8, db, idesc, acc);
            else umma_ss(tmem + (uint32_t)(p + q) * N, make_desc(sa + p * A_T), db, idesc, acc);
          }
      }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    if (cycles) cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (out != nullptr && blockIdx.x == 0)
    for (int d = 0; d < S; ++d)
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(d * N + c0), v);
        for (int j = 0; j < 16; ++j) out[((size_t)d * 128 + tid) * N + c0 + j] = (int32_t)v[j];
      }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int S, int NT, int DB, int KS, int MODE = 0>
int run(int reps_rate, int data_mode = 0) {
  const int Ktot = KS * KB;
  std::vector<int8_t> hA((size_t)S * 128 * Ktot), hB((size_t)S * N * Ktot);
  srand(99 + S + NT);
  // data_mode: 0 = random full-range digits, 1 = small digits 0..3, 2 = all zero (the MMA timing depends on the data)
  for (auto& x : hA) x = data_mode == 2 ? 0 : data_mode == 1 ? (int8_t)(rand() % 4) : (int8_t)(rand() % 256 - 128);
  for (auto& x : hB) x = data_mode == 2 ? 0 : data_mode == 1 ? (int8_t)(rand() % 4) : (int8_t)(rand() % 256 - 128);
  std::vector<int32_t> ref((size_t)S * 128 * N, 0), got((size_t)S * 128 * N, -1);
  for (int p = 0; p < S; ++p)
    for (int q = 0; p + q < S; ++q)
      for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) {
          int32_t s = 0;
          for (int k = 0; k < Ktot; ++k) s += (int32_t)hA[((size_t)p * 128 + i) * Ktot + k] * (int32_t)hB[((size_t)q * N + j) * Ktot + k];
          ref[((size_t)(p + q) * 128 + i) * N + j] += s;
        }
  int8_t *dA, *dB;
  int32_t* dO;
  long long* dC;
  CK(cudaMalloc(&dA, hA.size()));
  CK(cudaMalloc(&dB, hB.size()));
  CK(cudaMalloc(&dO, got.size() * 4));
  CK(cudaMalloc(&dC, 148 * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, got.size() * 4));
  const int smem = KS * S * (128 + N) * KB + 1024;
  CK(cudaFuncSetAttribute(ts_kernel<S, NT, DB, KS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ts_kernel<S, NT, DB, KS, MODE><<<1, 128, smem>>>(dA, dB, dO, 1, nullptr);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < got.size(); ++i) bad += got[i] != ref[i];
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ts_kernel<S, NT, DB, KS, MODE><<<148, 128, smem>>>(dA, dB, nullptr, reps_rate, dC);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc[148];
  CK(cudaMemcpy(cyc, dC, sizeof cyc, cudaMemcpyDeviceToHost));
  const double ksteps = (double)reps_rate * KS;
  if (MODE) printf("[mode %d: %s] ", MODE, MODE == 1 ? "A copied to TMEM once, MMAs only in the loop (results not comparable)" : "cp only, no MMAs");
  printf("data=%s S=%d A-slices in TMEM=%d double-buffered=%d k-steps=%d : %s (%zu / %zu mismatches) | %.0f cycles per k-step "
         "(%d MMAs + %d cp), %.1f TF-equivalent chip (FP64 flops of the 128x64x32 tile-step / time)\n",
         data_mode == 2 ? "zero" : data_mode == 1 ? "small" : "random", S, NT, DB, KS, bad ? "FAIL" : "ok", bad, got.size(), cyc[0] / ksteps, S * (S + 1) / 2, NT,
         2.0 * 128 * N * 32 * ksteps * 148 / (ms * 1e-3) / 1e12);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
  return bad ? 1 : 0;
}

int main() {
  int bad = 0;
  bad += run<7, 0, 0, 4>(200);   // all SS: baseline (1344 cycles per k-step expected)
  bad += run<7, 0, 0, 4>(200, 1);
  bad += run<7, 0, 0, 4>(200, 2);
  bad += run<7, 0, 0, 4>(2000, 0);   // 10 x longer: set-up amortised, wall-clock rate ~ cycle rate if the clock holds
  bad += run<7, 0, 0, 4>(2000, 2);
  run<7, 7, 0, 4, 1>(200);       // pure TS-mode MMA rate
  run<7, 7, 0, 4, 2>(200);       // pure tcgen05.cp rate
  run<6, 6, 0, 4, 1>(200);
  bad += run<7, 7, 0, 4>(200);   // all A slices in TMEM, single buffer: relies on mma -> cp ordering
  bad += run<7, 4, 1, 4>(200);   // A_0..A_3 in TMEM, double-buffered (64 spare columns)
  bad += run<7, 4, 0, 4>(200);
  bad += run<7, 2, 1, 4>(200);
  bad += run<6, 6, 1, 4>(200);   // S = 6: 384 + 2 x 48 = 480 columns
  bad += run<6, 6, 0, 4>(200);
  bad += run<6, 0, 0, 4>(200);
  printf("ts probe done, %d failing\n", bad);
  return 0;
}
