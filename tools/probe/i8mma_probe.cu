// Probe: tcgen05.mma.kind::i8 on B200 (sm_100a) -- (1) which shared-memory descriptor encodings give the right
// int32 result for K-major operands under the 32/64/128-byte swizzles, with the K advance done by bumping the
// descriptor start address inside the swizzle atom; (2) the MMA rate the tensor core sustains out of resident
// shared memory for N = 64 / 128 / 256 (shared-memory operand bandwidth is the limiter for narrow N).
// Feeds the design of the sliced-integer FP64 GEMM (csrc/ozaki.cu).   nvcc -gencode arch=compute_100a,code=sm_100a
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
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// mode: 0 = plain (A discarded), 1 = collector::a::fill, 2 = ::use, 3 = ::lastuse -- A operand kept in the tensor core's
// collector buffer across consecutive MMAs that share it (SASS: A_KEEP / A_REUSE.A_KEEP / A_REUSE)
#define UMMA_I8_ASM(MOD)                                                                              \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"                                           \
               "tcgen05.mma.cta_group::1.kind::i8" MOD " [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" \
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory")
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum, int mode = 0) {
  if (mode == 1) UMMA_I8_ASM(".collector::a::fill");
  else if (mode == 2) UMMA_I8_ASM(".collector::a::use");
  else if (mode == 3) UMMA_I8_ASM(".collector::a::lastuse");
  else UMMA_I8_ASM("");
}
__device__ __forceinline__ int coll_mode(int q, int nq) { return nq == 1 ? 0 : q == 0 ? 1 : q == nq - 1 ? 3 : 2; }
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

// K-major operand tile: `rows` rows x KB bytes (KB = swizzle span: 32 / 64 / 128), row pitch KB, 16-byte chunk index
// XOR-ed with (row % 8) restricted to the chunks of one row (Swizzle<log2(KB/16), 4, 3> on the byte address).
__host__ __device__ inline uint32_t swz_off(int row, int kbyte, int KB) {
  const uint32_t a = (uint32_t)row * KB + kbyte;
  const uint32_t mask = (KB / 16 - 1);                 // 1, 3, 7
  return a ^ (((a >> 7) & mask) << 4);
}
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, int KB) {
  const uint64_t layout = KB == 128 ? 2 : KB == 64 ? 4 : 6;
  const uint64_t sbo = (uint64_t)(8 * KB) >> 4;         // 8-row group pitch
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- (1) correctness: D_d = sum_{p+q=d} A_p B_q^T over Ktot, S slices, one CTA ------------------------------------
template <int KB, int S, int N, int COLL>
__global__ void __launch_bounds__(128, 1)
check_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int Ktot, int32_t* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  constexpr int A_BYTES = 128 * KB, B_BYTES = N * KB;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc(128, N);
  uint32_t phase = 0;
  for (int k0 = 0; k0 < Ktot; k0 += KB) {
    // stage the chunk: slice p of A at p*A_BYTES, slice q of B at S*A_BYTES + q*B_BYTES
    for (int idx = tid; idx < S * 128 * KB; idx += 128) {
      const int p = idx / (128 * KB), r = (idx / KB) % 128, c = idx % KB;
      sm[p * A_BYTES + swz_off(r, c, KB)] = (uint8_t)A[((size_t)p * 128 + r) * Ktot + k0 + c];
    }
    for (int idx = tid; idx < S * N * KB; idx += 128) {
      const int q = idx / (N * KB), r = (idx / KB) % N, c = idx % KB;
      sm[S * A_BYTES + q * B_BYTES + swz_off(r, c, KB)] = (uint8_t)B[((size_t)q * N + r) * Ktot + k0 + c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < KB; kk += 32)
        for (int p = 0; p < S; ++p)
          for (int q = 0; p + q < S; ++q) {
            const uint64_t da = make_desc(base + p * A_BYTES + kk, KB);
            const uint64_t db = make_desc(base + S * A_BYTES + q * B_BYTES + kk, KB);
            // accumulator d = p + q is initialised by its first pair in issue order: (p = 0, q = d)
            umma_i8(tmem + (uint32_t)(p + q) * N, da, db, idesc, (k0 == 0 && kk == 0 && p == 0) ? 0u : 1u, COLL ? coll_mode(q, S - p) : 0);
          }
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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

template <int KB, int S, int N, int COLL = 0>
int run_check(int Ktot) {
  std::vector<int8_t> hA((size_t)S * 128 * Ktot), hB((size_t)S * N * Ktot);
  srand(1234 + KB + N);
  for (auto& x : hA) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : hB) x = (int8_t)(rand() % 256 - 128);
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
  CK(cudaMalloc(&dA, hA.size()));
  CK(cudaMalloc(&dB, hB.size()));
  CK(cudaMalloc(&dO, got.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, got.size() * 4));
  const int smem = S * (128 + N) * KB + 1024;
  CK(cudaFuncSetAttribute(check_kernel<KB, S, N, COLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  check_kernel<KB, S, N, COLL><<<1, 128, smem>>>(dA, dB, Ktot, dO);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < got.size(); ++i) bad += got[i] != ref[i];
  printf("check KB=%3d S=%d N=%3d coll=%d Ktot=%4d : %s (%zu / %zu mismatches; got[0]=%d ref[0]=%d)\n", KB, S, N, COLL, Ktot,
         bad ? "FAIL" : "ok", bad, got.size(), got[0], ref[0]);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return bad ? 1 : 0;
}

// ---- (2) rate: one CTA per SM issues `reps` x (S(S+1)/2 pairs x KB/32 k-steps) MMAs on resident operands -----------
template <int KB, int S, int N, int COLL>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, int nacc, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr int A_BYTES = 128 * KB, B_BYTES = N * KB;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < S * (128 + N) * KB / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x01010101u * (i & 3);
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
  const uint32_t idesc = make_idesc(128, N);
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int kk = 0; kk < KB; kk += 32)
        for (int p = 0; p < S; ++p)
          for (int q = 0; p + q < S; ++q) {
            const uint64_t da = make_desc(base + p * A_BYTES + kk, KB);
            const uint64_t db = make_desc(base + S * A_BYTES + q * B_BYTES + kk, KB);
            umma_i8(tmem + (uint32_t)((p + q) % nacc) * N, da, db, idesc, 1u, COLL ? coll_mode(q, S - p) : 0);
          }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int KB, int S, int N, int COLL = 0>
int run_rate(int reps, int nacc) {
  const int smem = S * (128 + N) * KB + 1024;
  if (smem > 227 * 1024) return 0;
  CK(cudaFuncSetAttribute(rate_kernel<KB, S, N, COLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* dcyc;
  CK(cudaMalloc(&dcyc, 148 * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<KB, S, N, COLL><<<148, 128, smem>>>(2, nacc, dcyc);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  rate_kernel<KB, S, N, COLL><<<148, 128, smem>>>(reps, nacc, dcyc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc[148];
  CK(cudaMemcpy(cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost));
  const double mmas = (double)reps * (KB / 32) * (S * (S + 1) / 2);
  const double macs = mmas * 128.0 * N * 32.0;
  printf("rate  KB=%3d S=%d N=%3d coll=%d nacc=%d : %.1f cycles/MMA (SM 0), %.2f POP/s chip (%.3f ms), %.0f MAC/clk/SM\n", KB, S, N, COLL, nacc,
         cyc[0] / mmas, 2.0 * macs * 148 / (ms * 1e-3) / 1e15, ms, macs / cyc[0]);
  cudaFree(dcyc);
  return 0;
}

int main() {
  int bad = 0;
  bad += run_check<128, 7, 64>(256);
  bad += run_check<64, 7, 64>(256);
  bad += run_check<32, 7, 64>(256);
  bad += run_check<128, 3, 128>(256);
  bad += run_check<64, 2, 256>(128);
  bad += run_check<32, 8, 64>(64);
  bad += run_check<64, 7, 64, 1>(256);
  bad += run_check<32, 7, 64, 1>(256);
  bad += run_check<128, 7, 64, 1>(256);
  bad += run_check<64, 6, 80, 1>(128);
  run_rate<64, 7, 64, 1>(400, 7);
  run_rate<32, 7, 64, 1>(800, 7);
  run_rate<128, 7, 64, 1>(200, 7);
  run_rate<64, 6, 80, 1>(500, 6);
  run_rate<64, 8, 64, 1>(400, 8);
  run_rate<64, 3, 128, 1>(1500, 3);
  run_rate<128, 7, 64>(200, 7);
  run_rate<64, 7, 64>(400, 7);
  run_rate<32, 7, 64>(800, 7);
  run_rate<32, 7, 64>(800, 1);
  run_rate<64, 3, 128>(1500, 3);
  run_rate<128, 3, 128>(800, 3);
  run_rate<32, 4, 128>(2000, 4);
  run_rate<64, 2, 256>(3000, 2);
  run_rate<128, 2, 256>(1500, 2);
  run_rate<64, 6, 80>(500, 6);
  printf("probe done, %d failing checks\n", bad);
  return 0;
}
