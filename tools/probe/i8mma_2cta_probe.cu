// Probe 3 for the sliced-integer GEMM: tcgen05.mma.cta_group::2 kind::i8, M = 256 (two CTAs of a cluster, 128 rows each),
// N = 64.  The single-CTA 128 x 64 x 32 MMA is bound by operand delivery into the tensor core: 32 cycles for the 4 KB
// of A + 16 for the 2 KB of B = 48 (profiles/r02_i8mma_probe.txt).  In the pair each CTA still delivers its own A rows
// but only HALF of B (32 rows, 1 KB): 32 + 8 = 40 cycles expected for the same 128 x 64 of output per CTA.
// (1) correctness of the operand placement (B rows [32 r, 32 r + 32) in CTA r, same shared-memory offsets in both CTAs,
// descriptors of the leader) and of the accumulator layout (CTA r holds rows [128 r, 128 r + 128)); (2) the rate.
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
__device__ __forceinline__ void umma2_i8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
constexpr int KB = 32, N = 64, NH = N / 2;
__host__ __device__ inline uint32_t swz_off(int row, int kbyte) {
  const uint32_t a = (uint32_t)row * KB + kbyte;
  return a ^ (((a >> 7) & 1u) << 4);
}
__host__ __device__ inline uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((8 * KB) >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
__host__ __device__ inline uint32_t make_idesc(int M, int Nn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A: [S][256][Ktot] int8, B: [S][64][Ktot]; out: [S][256][64] int32 (only cluster 0 writes).  KS k-steps resident.
template <int S, int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ out, int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  constexpr int A_T = 128 * KB, B_T = NH * KB, STEP = S * (A_T + B_T);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  const int Ktot = KS * KB;
  for (int idx = tid; idx < KS * S * 128 * KB; idx += 128) {
    const int ks = idx / (S * 128 * KB), p = (idx / (128 * KB)) % S, r = (idx / KB) % 128, c = idx % KB;
    sm[ks * STEP + p * A_T + swz_off(r, c)] = (uint8_t)A[((size_t)p * 256 + rank * 128 + r) * Ktot + ks * KB + c];
  }
  for (int idx = tid; idx < KS * S * NH * KB; idx += 128) {
    const int ks = idx / (S * NH * KB), q = (idx / (NH * KB)) % S, r = (idx / KB) % NH, c = idx % KB;
    sm[ks * STEP + S * A_T + q * B_T + swz_off(r, c)] = (uint8_t)B[((size_t)q * N + rank * NH + r) * Ktot + ks * KB + c];
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc(256, N);
  if (rank == 0 && tid == 0) {
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t sa = base + ks * STEP, sb = sa + S * A_T;
        for (int p = 0; p < S; ++p)
          for (int q = 0; p + q < S; ++q)
            umma2_i8(tmem + (uint32_t)(p + q) * N, make_desc(sa + p * A_T), make_desc(sb + q * B_T), idesc,
                     (rep == 0 && ks == 0 && p == 0) ? 0u : 1u);
      }
    umma2_commit(smem_u32(&bar));
  }
  if (tid == 0) {
    mbar_wait(smem_u32(&bar), 0);
    if (rank == 0 && cycles) cycles[blockIdx.x / 2] = 0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (out != nullptr && blockIdx.x < 2)
    for (int d = 0; d < S; ++d)
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(d * N + c0), v);
        for (int j = 0; j < 16; ++j) out[((size_t)d * 256 + rank * 128 + tid) * N + c0 + j] = (int32_t)v[j];
      }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int S, int KS>
int run(int reps_rate) {
  const int Ktot = KS * KB;
  std::vector<int8_t> hA((size_t)S * 256 * Ktot), hB((size_t)S * N * Ktot);
  srand(7 + S);
  for (auto& x : hA) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : hB) x = (int8_t)(rand() % 256 - 128);
  std::vector<int32_t> ref((size_t)S * 256 * N, 0), got((size_t)S * 256 * N, -1);
  for (int p = 0; p < S; ++p)
    for (int q = 0; p + q < S; ++q)
      for (int i = 0; i < 256; ++i)
        for (int j = 0; j < N; ++j) {
          int32_t s = 0;
          for (int k = 0; k < Ktot; ++k) s += (int32_t)hA[((size_t)p * 256 + i) * Ktot + k] * (int32_t)hB[((size_t)q * N + j) * Ktot + k];
          ref[((size_t)(p + q) * 256 + i) * N + j] += s;
        }
  int8_t *dA, *dB;
  int32_t* dO;
  CK(cudaMalloc(&dA, hA.size()));
  CK(cudaMalloc(&dB, hB.size()));
  CK(cudaMalloc(&dO, got.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, got.size() * 4));
  const int smem = KS * S * (128 + NH) * KB + 1024;
  CK(cudaFuncSetAttribute(pair_kernel<S, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pair_kernel<S, KS><<<2, 128, smem>>>(dA, dB, dO, 1, nullptr);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0, bad_lo = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    bad += got[i] != ref[i];
    if (((i / N) % 256) < 128) bad_lo += got[i] != ref[i];
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  pair_kernel<S, KS><<<148, 128, smem>>>(dA, dB, nullptr, 2, nullptr);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  pair_kernel<S, KS><<<148, 128, smem>>>(dA, dB, nullptr, reps_rate, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ksteps = (double)reps_rate * KS;
  const double cyc_per_kstep = ms * 1e-3 * 1.965e9 / ksteps;   // at the 1965 MHz the bench runs at
  printf("cta_group::2 M=256 N=64 S=%d k-steps=%d : %s (%zu / %zu mismatches, %zu in CTA 0's rows) | ~%.0f cycles per k-step "
         "(%d MMAs, %.1f per MMA), %.1f TF-equivalent chip\n",
         S, KS, bad ? "FAIL" : "ok", bad, got.size(), bad_lo, cyc_per_kstep, S * (S + 1) / 2, cyc_per_kstep / (S * (S + 1) / 2),
         2.0 * 256 * N * 32 * ksteps * 74 / (ms * 1e-3) / 1e12);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return bad ? 1 : 0;
}

int main() {
  int bad = 0;
  bad += run<7, 4>(400);
  bad += run<6, 4>(400);
  printf("2cta probe done, %d failing\n", bad);
  return 0;
}
