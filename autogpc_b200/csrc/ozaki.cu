// FP64-accurate tile GEMM  C = alpha * A * B^T + beta * C  on the Blackwell tensor core proper: operands cut into S
// signed 8-bit digit slices, products formed by tcgen05.mma kind::i8 with exact int32 accumulation in TMEM, recombined
// in FP64 in the epilogue (Ozaki-style error-free slicing).
//
// Same role as gemm_nt.cu (the dsyrk / dgemm / dtrtri / dlauum calls GPy reaches through util/linalg.py jitchol, dtrtrs,
// pdinv from EP.inference, for gpckernel.py:245-246) for the products whose K extent amortises the slicing: the
// K = 512 trailing updates of potrf, the upper levels of trtri, lauum.  The DMMA pipe that gemm_nt.cu runs on tops out
// at 37 TFLOP/s on B200 and that kernel already fills it to 90 %; kind::i8 runs at 4.5 POP/s, so S (S + 1) / 2 = 28
// integer MMAs per FP64 product still come out 3x ahead.
//
// Numerics.  Per operand row i one power-of-two scale 2^e_i with |a_ik| 2^-e_i < 1/4 over the K range the product
// uses; X_ik = rint(a_ik 2^(8S - e_i)) is an integer below 2^(8S-2) (exact for S = 7 wherever a_ik is within 2^-1 of the
// row scale: 53 bits fit), written as sum_p d_p 256^(S-1-p) with round-to-nearest signed digits d_p in [-128, 127]
// (bytes of (X + BIAS) ^ BIAS, BIAS = 0x80 in every byte).  Then
//     (A B^T)_ij = 2^(ea_i + eb_j - 16) sum_{p+q <= S-1} 256^-(p+q) (D_p^A D_q^B^T)_ij   + O(S 2^(-8 S + 4)) |a_i|_max |b_j|_max K
// and the integer products of equal weight p + q share ONE int32 accumulator (|.| <= (p+q+1) 2^14 K: exact for
// K < 18 k).  tools/probe/ozaki_prototype.py: with S = 7 potrf / trtri / lauum of B = I + S^1/2 K S^1/2 agree with
// LAPACK to 7e-15 / 5e-14 / 1e-14 (native blocked FP64: 4e-15 / 2e-15 / 7e-15); S = 6 to 3e-12 / 2e-11 / 1e-11.
//
// sm_100a design (profiles/r02_i8mma_probe.txt).  One CTA owns a 128 x 64 tile of C and all S accumulators for it:
// S x 64 TMEM columns (448 of 512 for S = 7) -- that, not shared memory, is what fixes the tile width; the MMA then
// runs shared-memory-bound at 48 cycles per 128 x 64 x 32 instruction (4 KB of A + 2 KB of B at 128 B/clk), 67 % of
// the kind::i8 peak.  Operand slices live in HBM already in the shared-memory image the MMA wants (K-major, 32-byte
// swizzle, one 4 KB tile per slice per 128 rows x 32 k), so a stage is filled by S + 1 plain bulk copies
// (cp.async.bulk, mbarrier complete_tx) with no tensor map; 5-stage ring, one producer thread, one MMA-issuing
// thread, 8 epilogue warps that read TMEM (tcgen05.ld), run the FP64 Horner recombination, and go through a padded
// shared-memory transposition so that every global access of C is a full-line row access.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace agpc {
namespace oz {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded: a protocol error must surface as a launch failure (trap), never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  } while (!done);
}
// the same for waits that are long by design (epilogue warps waiting out a whole main loop): sleep between polls, so that
// sixteen polling warps do not compete with the MMA operand reads for shared-memory cycles
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, unsigned ns) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) {
      if (++spins > (1u << 24)) __trap();
      __nanosleep(ns);
    }
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// the same copy delivered to the same CTA-relative address (and mbarrier) of every CTA in cta_mask of the cluster
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mode 0: A operand discarded after the MMA; 1 / 2 / 3: collector::a::fill / use / lastuse -- consecutive MMAs that share
// their A slice keep it in the tensor core's operand collector (SASS A_KEEP / A_REUSE)
#define AGPC_UMMA_I8(MOD)                                                                              \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"                                             \
               "tcgen05.mma.cta_group::1.kind::i8" MOD " [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" \
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory")
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum, int mode) {
  if (mode == 1) AGPC_UMMA_I8(".collector::a::fill");
  else if (mode == 2) AGPC_UMMA_I8(".collector::a::use");
  else if (mode == 3) AGPC_UMMA_I8(".collector::a::lastuse");
  else AGPC_UMMA_I8("");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// max |v| into *dst (non-negative doubles order like their bit patterns; `bad` = a non-finite value was seen: NaN is stored)
__device__ __forceinline__ void absmax_commit(double* dst, double m, bool bad) {
  const unsigned long long bits = bad ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(m);
  atomicMax(reinterpret_cast<unsigned long long*>(dst), bits);
}
// exact int32 -> double without the conversion pipe: 2^52 + 2^31 + v has v in its low mantissa bits
__device__ __forceinline__ double i32_to_f64(uint32_t v) {
  return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;
}

}  // namespace oz

constexpr int OZ_STAGES = 5;
constexpr int OZ_THREADS = 320;  // warp 0: bulk-copy producer, warp 1: TMEM owner + MMA issuer, warps 2-9: epilogue
constexpr int OZ_EPI_LD = OZ_BN + 1;  // padded row pitch (doubles) of the epilogue transposition buffer
template <int S>
constexpr int oz_stage_bytes() {
  return S * (TILE + OZ_BN) * OZ_KB;
}
template <int S>
constexpr int oz_smem_bytes() {
  return OZ_STAGES * oz_stage_bytes<S>() + 1024 /*align slack*/ + 256 /*barriers*/;
}
static_assert(TILE * OZ_EPI_LD * 8 <= OZ_STAGES * oz_stage_bytes<6>(), "epilogue buffer must fit the ring");

// K-major 32-byte-swizzled tile image: 16-byte chunk index XOR-ed with bit 7 of the byte offset (Swizzle<1,4,3>)
__device__ __forceinline__ uint32_t oz_swz(uint32_t row, uint32_t kbyte) {
  const uint32_t a = row * OZ_KB + kbyte;
  return a ^ (((a >> 7) & 1u) << 4);
}

// PAIR = true: launched as clusters of two CTAs -- the two 64-column halves of one 128 x 128 tile.  They need the same A
// stage, so each loads half of it and multicasts it into both shared memories (a third less L2 -> SM traffic per MMA,
// which is what the single-CTA form is bound by at 6.3 TB/s); a stage is released when BOTH CTAs' MMAs have read it.
template <int S, bool PAIR>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const OzGemmArgs args) {
  using namespace oz;
  constexpr int CPB = TILE / OZ_KB;  // K chunks per 128-block
  const int b = blockIdx.z;
  if (args.active != nullptr && args.active[b] == 0) return;
  const int g = blockIdx.y;
  int h2 = 0;
  if (args.grp_blocks > 0) {
    h2 = args.nblk - g * args.grp_blocks - args.half_blocks;
    h2 = h2 < 0 ? 0 : (h2 > args.half_blocks ? args.half_blocks : h2);
  }
  const int shift = g * args.grp_blocks * TILE;
  const int mt = args.m_kind == EXT_H2 ? h2 : args.m_tiles;
  const int nt = args.n_kind == EXT_H2 ? h2 : args.n_tiles;
  const int kt = args.k_kind == EXT_H2 ? h2 * CPB : args.k_chunks;
  if (mt <= 0 || nt <= 0) return;
  const int half = blockIdx.x & 1;  // which 64-column half of the 128-wide tile column
  const int t = blockIdx.x >> 1;
  int ti, tj;
  if (args.lower_only) {
    if (t >= mt * (mt + 1) / 2) return;
    ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = t / nt;
    tj = t % nt;
    if (ti >= mt) return;
    if (args.skip_upper && tj > ti) return;
  }
  int k_begin = 0, k_end = kt;
  if (args.k_rule == KRULE_BEGIN_AT_ROW) k_begin = ti * CPB;
  if (args.k_rule == KRULE_END_AT_ROW) k_end = min(kt, (ti + 1) * CPB);
  if (args.k_rule == KRULE_END_AT_COL) k_end = min(kt, (2 * tj + half + 1) * (OZ_BN / OZ_KB));
  const int n_iter = k_end - k_begin;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr uint32_t STAGE = oz_stage_bytes<S>();
  constexpr uint32_t A_SLICE = TILE * OZ_KB, B_SLICE = OZ_BN * OZ_KB;
  const uint32_t bars = smem_base + OZ_STAGES * STAGE;  // full[0..ST), empty[0..ST), accf
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + OZ_STAGES * STAGE + 8 * (2 * OZ_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * OZ_STAGES + 1; ++s)
      mbar_init(bars + 8 * s, (PAIR && s >= OZ_STAGES && s < 2 * OZ_STAGES) ? 2 : 1);  // empty[]: one commit per CTA of the pair
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything of ours can land on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int a_row = args.a_row0 + shift + ti * TILE;
  const int b_row = args.b_row0 + shift + tj * TILE + half * OZ_BN;

  if (warp == 0) {
    // ---------------- producer: S + 1 bulk copies per stage ----------------
    if (lane == 0) {
      const long long tileA = (long long)(a_row >> 7) * args.A.nkc + ((args.a_col0 + shift) / OZ_KB) + k_begin;
      const long long tileB = (long long)(b_row >> 7) * args.B.nkc + ((args.b_col0 + shift) / OZ_KB) + k_begin;
      const int8_t* srcA = args.A.S + (long long)b * args.A.s_batch + tileA * (S * (long long)A_SLICE);
      const int8_t* srcB = args.B.S + (long long)b * args.B.s_batch + tileB * (S * (long long)A_SLICE) + ((b_row >> 6) & 1) * B_SLICE;
      for (int it = 0; it < n_iter && args.dbg != 2; ++it) {
        const int s = it % OZ_STAGES;
        const uint32_t ph = (uint32_t)(it / OZ_STAGES) & 1u;
        mbar_wait(bars + 8 * (OZ_STAGES + s), ph ^ 1u);
        const uint32_t full = bars + 8 * s;
        mbar_expect_tx(full, STAGE);
        const uint32_t dst = smem_base + s * STAGE;
        if (PAIR) {  // my half of the A stage, into both CTAs
          constexpr uint32_t HALF_A = (S * A_SLICE / 2 + 15u) & ~15u;
          const uint32_t off = half ? HALF_A : 0u, len = half ? S * A_SLICE - HALF_A : HALF_A;
          bulk_g2s_mc(dst + off, srcA + (long long)it * (S * (long long)A_SLICE) + off, len, full, (uint16_t)3);
        } else {
          bulk_g2s(dst, srcA + (long long)it * (S * (long long)A_SLICE), S * A_SLICE, full);
        }
#pragma unroll
        for (int q = 0; q < S; ++q)
          bulk_g2s(dst + S * A_SLICE + q * B_SLICE, srcB + (long long)it * (S * (long long)A_SLICE) + q * (long long)A_SLICE, B_SLICE, full);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 64, M = 128
      constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
      // shared-memory descriptor: 32-byte swizzle (6), 8-row group pitch 256 B, LBO field 1, version 1
      constexpr uint64_t DESC_HI = (1ull << 16) | ((uint64_t)((8 * OZ_KB) >> 4) << 32) | (1ull << 46) | (6ull << 61);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % OZ_STAGES;
        const uint32_t ph = (uint32_t)(it / OZ_STAGES) & 1u;
        if (args.dbg != 2) mbar_wait(bars + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * STAGE, sb = sa + S * A_SLICE;
        if (args.dbg != 1)
#pragma unroll
        for (int p = 0; p < S; ++p) {
          const uint64_t da = DESC_HI | (uint64_t)(((sa + p * A_SLICE) >> 4) & 0x3FFF);
#pragma unroll
          for (int q = 0; q < S - p; ++q) {
            const uint64_t db = DESC_HI | (uint64_t)(((sb + q * B_SLICE) >> 4) & 0x3FFF);
            const int nq = S - p;
            const int mode = nq == 1 ? 0 : q == 0 ? 1 : q == nq - 1 ? 3 : 2;
            // accumulator p + q: overwritten by its first product of the tile (it == 0, p == 0), accumulated after
            umma_i8(tmem + (uint32_t)(p + q) * OZ_BN, da, db, IDESC, (it == 0 && p == 0) ? 0u : 1u, mode);
          }
        }
        if (PAIR) umma_commit_mc(bars + 8 * (OZ_STAGES + s), (uint16_t)3);  // frees the stage in both CTAs of the pair ...
        else umma_commit(bars + 8 * (OZ_STAGES + s));                       // ... when these MMAs have read it
      }
      if (n_iter > 0) umma_commit(bars + 8 * (2 * OZ_STAGES));  // accumulators complete
    }
  } else {
    // ---------------- epilogue: TMEM -> FP64 Horner -> padded smem -> global ----------------
    const int ew = warp - 2;           // 0..7
    const int q4 = warp & 3;           // TMEM lane quarter this warp may read
    const int ch = ew >> 2;            // column half (32 columns) of the 64-wide tile
    const int row = q4 * 32 + lane;
    double* stage = reinterpret_cast<double*>(smem_gen);
    if (n_iter > 0) {
      mbar_wait(bars + 8 * (2 * OZ_STAGES), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const double sa = args.A.scale[(long long)b * args.A.sc_batch + a_row + row] * args.alpha;
    const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ch * 32);
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 16) {
      double sum[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) sum[j] = 0.0;
      if (n_iter > 0) {
#pragma unroll
        for (int d = S - 1; d >= 0; --d) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)(d * OZ_BN + c0), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[j] = fma(sum[j], 0.00390625, i32_to_f64(v[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) stage[row * OZ_EPI_LD + ch * 32 + c0 + j] = sum[j] * sa;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const long long c_row = (long long)args.c_row0 + shift + ti * TILE;
    const long long c_col = (long long)args.c_col0 + shift + tj * TILE + half * OZ_BN;
    const double2 sb2 = *reinterpret_cast<const double2*>(args.B.scale + (long long)b * args.B.sc_batch + b_row + 2 * lane);
    double* Cb = args.C + (long long)b * args.c_batch;
    const double beta = args.beta;
    for (int rr = ew; rr < TILE; rr += 8) {
      double2 v;
      v.x = stage[rr * OZ_EPI_LD + 2 * lane] * sb2.x;
      v.y = stage[rr * OZ_EPI_LD + 2 * lane + 1] * sb2.y;
      double2* dst = reinterpret_cast<double2*>(Cb + (c_row + rr) * args.ldc + c_col + 2 * lane);
      if (beta != 0.0) {
        const double2 old = *dst;
        v.x = fma(beta, old.x, v.x);
        v.y = fma(beta, old.y, v.y);
      }
      *dst = v;
      if (args.row_absmax != nullptr) {
        double m = fmax(fabs(v.x), fabs(v.y));
        int bad = !(fabs(v.x) + fabs(v.y) <= 1.7976931348623157e308);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
          bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (lane == 0) absmax_commit(args.row_absmax + (long long)b * args.absmax_batch + c_row + rr, m, bad != 0);
      }
    }
    if (args.Ct != nullptr) {
      const long long t_row = (long long)args.ct_row0 + shift + tj * TILE + half * OZ_BN;
      const long long t_col = (long long)args.ct_col0 + shift + ti * TILE;
      double* Tb = args.Ct + (long long)b * args.c_batch;
      for (int cc = ew; cc < OZ_BN; cc += 8) {
        const double sbc = args.B.scale[(long long)b * args.B.sc_batch + b_row + cc];
        double m = 0.0, t = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double v = stage[(lane + 32 * i) * OZ_EPI_LD + cc] * sbc;
          Tb[(t_row + cc) * args.ldc + t_col + lane + 32 * i] = v;
          m = fmax(m, fabs(v));
          t += fabs(v);
        }
        if (args.col_absmax != nullptr) {
          int bad = !(t <= 1.7976931348623157e308);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            bad |= __shfl_xor_sync(0xffffffffu, bad, o);
          }
          if (lane == 0) absmax_commit(args.col_absmax + (long long)b * args.absmax_batch + t_row + cc, m, bad != 0);
        }
      }
    }
  }
  __syncwarp();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer may still arrive on our barriers / multicast until it is done as well
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

// ---- persistent form: tiles streamed through a resident CTA ---------------------------------------------------------
// One CTA per SM stays resident and pulls 128 x 64 tiles from a per-stream atomic counter.  What the one-tile-per-CTA form
// pays per tile -- CTA launch, barrier init, TMEM allocation, the first stage's load latency and the whole epilogue with
// the tensor core idle, about a third of a K = 512 tile -- is overlapped here:
//   * the producer keeps loading across the tile boundary (the stage ring belongs to the main loop alone: the epilogue
//     goes from TMEM through registers straight to global memory, 32 B per thread per store, full sectors);
//   * TMEM holds 8 accumulator slots of 64 columns for the S <= 7 accumulators of a tile, and the slot of accumulator d
//     rotates by one per tile: tile j keeps weight d in slot (d + j) mod 8, i.e. where tile j-1 kept weight d + 1.  The
//     epilogue drains weights S-1, S-2, ... 0 (its Horner order anyway) and signals each slot as it is read; the first
//     K step of the next tile issues its products weight-major in the same order, so its MMAs chase the drain one slot
//     behind and the tensor pipe never waits for an epilogue.
constexpr int OZP_RING = 4;      // tile records in flight between producer, MMA issuer and epilogue
constexpr int OZS_EPI_WARPS = 16;
constexpr int OZS_THREADS = 64 + 32 * OZS_EPI_WARPS;  // warp 0: producer, warp 1: TMEM owner + MMA issuer, warps 2-17: epilogue
constexpr int OZP_MAX_BATCH = 4096;  // compact list of active items in shared memory (2 B each)
struct OzTileRec {
  int b, a_row, b_row, a_kc, b_kc, n_iter;  // a_kc / b_kc: first K chunk in each operand's own columns; n_iter < 0: no more tiles
  int c_row, c_col, t_row, t_col;
  int pad[6];
};
static_assert(sizeof(OzTileRec) == 64, "one record per 64-byte slot");
constexpr int OZP_EXTRA = 256 /*barriers*/ + 256 /*tmem slot, counts*/ + OZP_RING * 64 + 2 * OZP_MAX_BATCH;
template <int S>
constexpr int ozp_smem_bytes() {
  return OZ_STAGES * oz_stage_bytes<S>() + 1024 /*align slack*/ + OZP_EXTRA;
}
static_assert(ozp_smem_bytes<7>() <= 227 * 1024, "persistent kernel: shared memory");

namespace oz {
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_global_v4(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
}  // namespace oz

// (x, group, item) -> tile record; false when the index names no work (ozaki_gemm_kernel's early returns)
__device__ __forceinline__ bool oz_decode_tile(const OzGemmArgs& args, int x, int g, int b, OzTileRec& T) {
  constexpr int CPB = TILE / OZ_KB;
  int h2 = 0;
  if (args.grp_blocks > 0) {
    h2 = args.nblk - g * args.grp_blocks - args.half_blocks;
    h2 = h2 < 0 ? 0 : (h2 > args.half_blocks ? args.half_blocks : h2);
  }
  const int shift = g * args.grp_blocks * TILE;
  const int mt = args.m_kind == EXT_H2 ? h2 : args.m_tiles;
  const int nt = args.n_kind == EXT_H2 ? h2 : args.n_tiles;
  const int kt = args.k_kind == EXT_H2 ? h2 * CPB : args.k_chunks;
  if (mt <= 0 || nt <= 0) return false;
  const int half = x & 1, t = x >> 1;
  int ti, tj;
  if (args.lower_only) {
    if (t >= mt * (mt + 1) / 2) return false;
    ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = t / nt;
    tj = t % nt;
    if (ti >= mt) return false;
    if (args.skip_upper && tj > ti) return false;
  }
  int k_begin = 0, k_end = kt;
  if (args.k_rule == KRULE_BEGIN_AT_ROW) k_begin = ti * CPB;
  if (args.k_rule == KRULE_END_AT_ROW) k_end = min(kt, (ti + 1) * CPB);
  if (args.k_rule == KRULE_END_AT_COL) k_end = min(kt, (2 * tj + half + 1) * (OZ_BN / OZ_KB));
  T.b = b;
  T.a_row = args.a_row0 + shift + ti * TILE;
  T.b_row = args.b_row0 + shift + tj * TILE + half * OZ_BN;
  T.a_kc = (args.a_col0 + shift) / OZ_KB + k_begin;
  T.b_kc = (args.b_col0 + shift) / OZ_KB + k_begin;
  T.n_iter = max(k_end - k_begin, 0);
  T.c_row = args.c_row0 + shift + ti * TILE;
  T.c_col = args.c_col0 + shift + tj * TILE + half * OZ_BN;
  T.t_row = args.ct_row0 + shift + tj * TILE + half * OZ_BN;
  T.t_col = args.ct_col0 + shift + ti * TILE;
  return true;
}

template <int S>
__global__ void __launch_bounds__(OZS_THREADS, 1)
ozaki_gemm_stream_kernel(const OzGemmArgs args, const int gx, const int gy, const int batch, int* __restrict__ counter) {
  using namespace oz;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr uint32_t STAGE = oz_stage_bytes<S>();
  constexpr uint32_t A_SLICE = TILE * OZ_KB, B_SLICE = OZ_BN * OZ_KB;
  constexpr int ST = OZ_STAGES, R = OZP_RING;
  // barriers: full[ST], empty[ST], accf, drained[S] (weight d of the current tile has been read out of TMEM),
  //           tfull[R] / tempty[R] (tile record ring)
  const uint32_t bars = smem_base + ST * STAGE;
  const uint32_t bar_accf = bars + 8 * (2 * ST), bar_drained = bar_accf + 8, bar_tfull = bar_drained + 8 * S,
                 bar_tempty = bar_tfull + 8 * R;
  static_assert(8 * (2 * ST + 1 + S + 2 * R) <= 256, "barrier block");
  uint8_t* extra = smem_gen + ST * STAGE + 256;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(extra);
  int* n_act_s = reinterpret_cast<int*>(extra + 8);
  volatile int* ring = reinterpret_cast<volatile int*>(extra + 256);  // R records of 16 ints
  uint16_t* act = reinterpret_cast<uint16_t*>(extra + 256 + R * 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * ST + 1; ++s) mbar_init(bars + 8 * s, 1);
    for (int d = 0; d < S; ++d) mbar_init(bar_drained + 8 * d, OZS_EPI_WARPS);  // one arrival per epilogue warp
    for (int r = 0; r < R; ++r) {
      mbar_init(bar_tfull + 8 * r, 1);
      mbar_init(bar_tempty + 8 * r, OZS_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 2) {  // compact list of the batch items that still take part
    int cnt = 0;
    for (int base = 0; base < batch; base += 32) {
      const int i = base + lane;
      const bool on = i < batch && (args.active == nullptr || args.active[i] != 0);
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) act[cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
      cnt += __popc(m);
    }
    if (lane == 0) *n_act_s = cnt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------- producer: tile fetch, then S + 1 bulk copies per stage ----------------
    if (lane == 0) {
      const int per_item = gx * gy;
      const long long total = (long long)per_item * *n_act_s;
      const bool longest_last = args.k_rule == KRULE_END_AT_ROW || args.k_rule == KRULE_END_AT_COL;
      int gs = 0, pub = 0;
      long long w = atomicAdd(counter, 1);
      for (;;) {
        OzTileRec T = {};
        bool have = false;
        while (w < total) {
          const int bi = (int)(w / per_item);
          int rem = (int)(w % per_item);
          if (longest_last) rem = per_item - 1 - rem;  // hand out an item's long tiles first: short ones fill the tail
          if (oz_decode_tile(args, rem % gx, rem / gx, (int)act[bi], T)) {
            have = true;
            break;
          }
          w = atomicAdd(counter, 1);
        }
        const int r = pub % R;
        if (pub >= R) mbar_wait(bar_tempty + 8 * r, (uint32_t)((pub / R) - 1) & 1u);
        if (!have) T.n_iter = -1;
        volatile int* rec = ring + 16 * r;
        rec[0] = T.b; rec[1] = T.a_row; rec[2] = T.b_row; rec[3] = T.n_iter;
        rec[4] = T.c_row; rec[5] = T.c_col; rec[6] = T.t_row; rec[7] = T.t_col;
        mbar_arrive(bar_tfull + 8 * r);
        ++pub;
        if (!have) break;
        const long long w_next = atomicAdd(counter, 1);  // its latency hides behind this tile's copies
        const int8_t* srcA = args.A.S + (long long)T.b * args.A.s_batch +
                             ((long long)(T.a_row >> 7) * args.A.nkc + T.a_kc) * (S * (long long)A_SLICE);
        const int8_t* srcB = args.B.S + (long long)T.b * args.B.s_batch +
                             ((long long)(T.b_row >> 7) * args.B.nkc + T.b_kc) * (S * (long long)A_SLICE) + ((T.b_row >> 6) & 1) * B_SLICE;
        for (int it = 0; it < T.n_iter; ++it, ++gs) {
          const int s = gs % ST;
          const uint32_t ph = (uint32_t)(gs / ST) & 1u;
          mbar_wait(bars + 8 * (ST + s), ph ^ 1u);
          const uint32_t full = bars + 8 * s;
          mbar_expect_tx(full, STAGE);
          const uint32_t dst = smem_base + s * STAGE;
          bulk_g2s(dst, srcA + (long long)it * (S * (long long)A_SLICE), S * A_SLICE, full);
#pragma unroll
          for (int q = 0; q < S; ++q)
            bulk_g2s(dst + S * A_SLICE + q * B_SLICE, srcB + (long long)it * (S * (long long)A_SLICE) + q * (long long)A_SLICE, B_SLICE, full);
        }
        w = w_next;
      }
      if (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1) {  // last CTA out re-arms the counter for the stream's next launch
        counter[0] = 0;
        counter[1] = 0;
        __threadfence();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
      constexpr uint64_t DESC_HI = (1ull << 16) | ((uint64_t)((8 * OZ_KB) >> 4) << 32) | (1ull << 46) | (6ull << 61);
      int gs = 0;
      for (int j = 0;; ++j) {
        const int r = j % R;
        mbar_wait(bar_tfull + 8 * r, (uint32_t)(j / R) & 1u);
        const int n_iter = ring[16 * r + 3];
        if (n_iter < 0) break;
        const uint32_t rot = (uint32_t)j & 7u;
        const uint32_t prev = (uint32_t)(j - 1) & 1u;
        const bool no_chase = args.dbg == 3;  // measurement aid: drain the previous tile completely before the first MMA
        if (j >= 1 && (n_iter == 0 || no_chase)) {  // (n_iter == 0: nothing to issue, still keep in step with the drain)
          for (int d = 1; d < S; ++d) mbar_wait(bar_drained + 8 * d, prev);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int it = 0; it < n_iter; ++it, ++gs) {
          const int s = gs % ST;
          const uint32_t ph = (uint32_t)(gs / ST) & 1u;
          mbar_wait(bars + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + s * STAGE, sb = sa + S * A_SLICE;
          if (it == 0) {
            // weight-major, lightest weight first: accumulator d takes over the slot weight d + 1 of the previous tile had
#pragma unroll
            for (int d = S - 1; d >= 0; --d) {
              if (j >= 1 && d <= S - 2 && !no_chase) {
                mbar_wait(bar_drained + 8 * (d + 1), prev);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              }
              const uint32_t acc = tmem + (((uint32_t)d + rot) & 7u) * OZ_BN;
#pragma unroll
              for (int p = 0; p <= d; ++p) {
                const uint64_t da = DESC_HI | (uint64_t)(((sa + p * A_SLICE) >> 4) & 0x3FFF);
                const uint64_t db = DESC_HI | (uint64_t)(((sb + (d - p) * B_SLICE) >> 4) & 0x3FFF);
                umma_i8(acc, da, db, IDESC, p == 0 ? 0u : 1u, 0);
              }
            }
          } else {
#pragma unroll
            for (int p = 0; p < S; ++p) {
              const uint64_t da = DESC_HI | (uint64_t)(((sa + p * A_SLICE) >> 4) & 0x3FFF);
#pragma unroll
              for (int q = 0; q < S - p; ++q) {
                const uint64_t db = DESC_HI | (uint64_t)(((sb + q * B_SLICE) >> 4) & 0x3FFF);
                const int nq = S - p;
                const int mode = nq == 1 ? 0 : q == 0 ? 1 : q == nq - 1 ? 3 : 2;  // A slice p stays in the operand collector
                umma_i8(tmem + (((uint32_t)(p + q) + rot) & 7u) * OZ_BN, da, db, IDESC, 1u, mode);
              }
            }
          }
          umma_commit(bars + 8 * (ST + s));
        }
        // the slot the NEXT tile writes first (its weight S-1) was weight 0 of the previous tile: make sure that drain is
        // over before this tile is handed to the epilogue (which also keeps every drained[] barrier within one phase)
        if (j >= 1) mbar_wait(bar_drained, prev);
        umma_commit(bar_accf);
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> FP64 Horner in registers -> global ----------------
    // 16 warps: lane quarter q4 (the TMEM lanes a warp may read) x 16-column group.  Everything that does not depend on
    // the accumulators -- the tile's old C values when beta != 0, the row and column scales -- is requested BEFORE the
    // wait for the MMAs, so its latency runs under the main loop.
    const int ew = warp - 2;  // 0..15
    const int q4 = warp & 3;
    const int cgp = ew >> 2;  // columns 16 cgp .. 16 cgp + 15 of the 64-wide tile
    const int row = q4 * 32 + lane;
    const double beta = args.beta;
    const unsigned idle_ns = args.dbg >= 100 ? (unsigned)(args.dbg - 100) : 0u;  // measurement aid: AGPC_OZ_DBG = 100 + ns
    for (int j = 0;; ++j) {
      const int r = j % R;
      if (idle_ns) mbar_wait_idle(bar_tfull + 8 * r, (uint32_t)(j / R) & 1u, idle_ns);
      else mbar_wait(bar_tfull + 8 * r, (uint32_t)(j / R) & 1u);
      volatile int* rec = ring + 16 * r;
      const int b = rec[0], a_row = rec[1], b_row = rec[2], n_iter = rec[3];
      const int c_row = rec[4], c_col = rec[5], t_row = rec[6], t_col = rec[7];
      if (n_iter < 0) break;
      const uint32_t rot = (uint32_t)j & 7u;
      double* crow = args.C + (long long)b * args.c_batch + (long long)(c_row + row) * args.ldc + c_col + cgp * 16;
      double old[16];
      if (beta != 0.0) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) ld_global_v4(crow + i, old[i], old[i + 1], old[i + 2], old[i + 3]);
      }
      const double sa = args.A.scale[(long long)b * args.A.sc_batch + a_row + row] * args.alpha;
      const double* sbp = args.B.scale + (long long)b * args.B.sc_batch + b_row + cgp * 16;
      if (idle_ns) mbar_wait_idle(bar_accf, (uint32_t)j & 1u, idle_ns);
      else mbar_wait(bar_accf, (uint32_t)j & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      double sum[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) sum[i] = 0.0;
      const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(cgp * 16);
#pragma unroll
      for (int d = S - 1; d >= 0; --d) {
        uint32_t v[16];
        if (n_iter > 0) {
          tmem_ld16(tbase + (((uint32_t)d + rot) & 7u) * OZ_BN, v);
          tmem_ld_wait();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_drained + 8 * d);
        if (n_iter > 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) sum[i] = fma(sum[i], 0.00390625, i32_to_f64(v[i]));
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) sum[i] = (sum[i] * sa) * __ldg(sbp + i);
      if (beta != 0.0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) old[i] = fma(beta, old[i], sum[i]);
#pragma unroll
        for (int i = 0; i < 16; i += 4) st_global_v4(crow + i, old[i], old[i + 1], old[i + 2], old[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 4) st_global_v4(crow + i, sum[i], sum[i + 1], sum[i + 2], sum[i + 3]);
      }
      if (args.row_absmax != nullptr) {
        double m = 0.0, t = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const double v = fabs(beta != 0.0 ? old[i] : sum[i]);
          m = fmax(m, v);
          t += v;
        }
        absmax_commit(args.row_absmax + (long long)b * args.absmax_batch + c_row + row, m, !(t <= 1.7976931348623157e308));
      }
      if (args.Ct != nullptr) {  // transposed copy (alpha A B^T alone, as in the one-tile kernel): lanes = consecutive columns of Ct
        double* Tb = args.Ct + (long long)b * args.c_batch + (long long)(t_row + cgp * 16) * args.ldc + t_col + row;
#pragma unroll
        for (int i = 0; i < 16; ++i) Tb[(long long)i * args.ldc] = sum[i];
        if (args.col_absmax != nullptr) {  // rows of Ct = columns of this tile: maximum over the warp's 32 rows, one atomic each
          double* cm = args.col_absmax + (long long)b * args.absmax_batch + t_row + cgp * 16;
          // 16 column maxima over 32 lanes by a halving butterfly: at distance 16 / 8 / 4 / 2 every lane hands the half of its
          // columns that its partner keeps to the partner (8 + 4 + 2 + 1 exchanges), then one exchange at distance 1; lane l
          // ends with column 8 b4 + 4 b3 + 2 b2 + b1 of its bits (31 shuffles instead of 16 x 5).  Non-finite values travel
          // as +inf (fmax would drop a NaN) and are stored as NaN.
          double v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const double t = fabs(sum[i]);
            v[i] = t <= 1.7976931348623157e308 ? t : __longlong_as_double(0x7ff0000000000000LL);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool up = (lane & 16) != 0;
            const double give = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
            v[i] = fmax(keep, __shfl_xor_sync(0xffffffffu, give, 16));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool up = (lane & 8) != 0;
            const double give = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
            v[i] = fmax(keep, __shfl_xor_sync(0xffffffffu, give, 8));
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const bool up = (lane & 4) != 0;
            const double give = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
            v[i] = fmax(keep, __shfl_xor_sync(0xffffffffu, give, 4));
          }
          {
            const bool up = (lane & 2) != 0;
            const double give = up ? v[0] : v[1], keep = up ? v[1] : v[0];
            v[0] = fmax(keep, __shfl_xor_sync(0xffffffffu, give, 2));
          }
          v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], 1));
          if ((lane & 1) == 0) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            absmax_commit(cm + col, v[0], !(v[0] <= 1.7976931348623157e308));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * r);
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

// ---- 128 x 128 tile, accumulators split over a CTA pair ------------------------------------------------------------------
// The 128 x 64 kernel above is bound by operand delivery: a 128 x 64 x 32 instruction takes 48-55 cycles, a 128 x 128 x
// 32 one 64 (profiles/r02_summary.md), but S accumulators of 128 columns do not fit the 512 TMEM columns of one SM.
// Here the two CTAs of a cluster own the SAME 128 x 128 tile and split the weight groups d = p + q between them (S = 7:
// CTA 0 takes d in {0, 1, 2, 6} = 13 products, CTA 1 d in {3, 4, 5} = 15): each issues N = 128 instructions into 4 / 3
// accumulators of 128 columns.  Both need (nearly) all operand slices, so CTA 0 loads the A stage and CTA 1 the B stage
// and each copy is multicast into both shared memories.  Epilogue: every CTA forms its partial FP64 sum for all 128
// columns, keeps the 64 columns it finalises (CTA r: columns 64 r ..) and writes the other 64 into the peer's shared
// memory (DSMEM); after a cluster barrier each adds the two partials, scales and updates its half of C.
constexpr int OZ2_STAGES = 4;
template <int S>
constexpr int oz2_stage_bytes() {
  return S * 2 * TILE * OZ_KB;
}
template <int S>
constexpr int oz2_smem_bytes() {
  return OZ2_STAGES * oz2_stage_bytes<S>() + 1024 + 256;
}
constexpr int OZ2_EPI_LD = OZ_BN + 1;
static_assert(2 * TILE * OZ2_EPI_LD * 8 <= OZ2_STAGES * oz2_stage_bytes<6>(), "the two epilogue buffers must fit the ring");

// weight groups of CTA `rank`, as a bit mask over d = 0 .. S-1 (product counts 13 | 15 for S = 7, 12 | 9 for S = 6)
template <int S>
__host__ __device__ constexpr uint32_t oz2_groups(int rank) {
  return S == 7 ? (rank == 0 ? 0x47u : 0x38u) : (rank == 0 ? 0x27u : 0x18u);
}

// MC = true: each CTA loads one operand stage and multicasts it (half the L2 traffic, but every stage then waits for
// both CTAs' producers and both CTAs' MMAs); MC = false: each CTA loads both operand stages for itself.
template <int S, bool MC>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm128_kernel(const OzGemmArgs args) {
  using namespace oz;
  constexpr int CPB = TILE / OZ_KB;
  const int b = blockIdx.z;
  if (args.active != nullptr && args.active[b] == 0) return;
  const int g = blockIdx.y;
  int h2 = 0;
  if (args.grp_blocks > 0) {
    h2 = args.nblk - g * args.grp_blocks - args.half_blocks;
    h2 = h2 < 0 ? 0 : (h2 > args.half_blocks ? args.half_blocks : h2);
  }
  const int shift = g * args.grp_blocks * TILE;
  const int mt = args.m_kind == EXT_H2 ? h2 : args.m_tiles;
  const int nt = args.n_kind == EXT_H2 ? h2 : args.n_tiles;
  const int kt = args.k_kind == EXT_H2 ? h2 * CPB : args.k_chunks;
  if (mt <= 0 || nt <= 0) return;
  const int rank = blockIdx.x & 1;  // position in the CTA pair (cluster of 2 along x)
  const int t = blockIdx.x >> 1;
  int ti, tj;
  if (args.lower_only) {
    if (t >= mt * (mt + 1) / 2) return;
    ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = t / nt;
    tj = t % nt;
    if (ti >= mt) return;
    if (args.skip_upper && tj > ti) return;
  }
  int k_begin = 0, k_end = kt;
  if (args.k_rule == KRULE_BEGIN_AT_ROW) k_begin = ti * CPB;
  if (args.k_rule == KRULE_END_AT_ROW) k_end = min(kt, (ti + 1) * CPB);
  if (args.k_rule == KRULE_END_AT_COL) k_end = min(kt, (tj + 1) * CPB);
  const int n_iter = k_end - k_begin;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr uint32_t SLICE = TILE * OZ_KB;          // 4 KB: one digit slice of 128 rows x 32 k
  constexpr uint32_t STAGE = oz2_stage_bytes<S>();  // A slices 0..S-1, then B slices 0..S-1
  const uint32_t bars = smem_base + OZ2_STAGES * STAGE;  // full[0..ST), empty[0..ST), accf
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + OZ2_STAGES * STAGE + 8 * (2 * OZ2_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * OZ2_STAGES + 1; ++s) mbar_init(bars + 8 * s, (MC && s >= OZ2_STAGES && s < 2 * OZ2_STAGES) ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int a_row = args.a_row0 + shift + ti * TILE;
  const int b_row = args.b_row0 + shift + tj * TILE;
  constexpr uint32_t GROUPS0 = oz2_groups<S>(0), GROUPS1 = oz2_groups<S>(1);
  const uint32_t my_groups = rank ? GROUPS1 : GROUPS0;

  if (warp == 0) {
    // ---------------- producer: CTA 0 brings the A stage, CTA 1 the B stage, each into BOTH shared memories ------------
    if (lane == 0) {
      const long long tileA = (long long)(a_row >> 7) * args.A.nkc + ((args.a_col0 + shift) / OZ_KB) + k_begin;
      const long long tileB = (long long)(b_row >> 7) * args.B.nkc + ((args.b_col0 + shift) / OZ_KB) + k_begin;
      const int8_t* srcA = args.A.S + (long long)b * args.A.s_batch + tileA * (S * (long long)SLICE);
      const int8_t* srcB = args.B.S + (long long)b * args.B.s_batch + tileB * (S * (long long)SLICE);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % OZ2_STAGES;
        const uint32_t ph = (uint32_t)(it / OZ2_STAGES) & 1u;
        mbar_wait(bars + 8 * (OZ2_STAGES + s), ph ^ 1u);
        const uint32_t full = bars + 8 * s;
        mbar_expect_tx(full, STAGE);
        const long long off = (long long)it * (S * (long long)SLICE);
        if (MC) {
          bulk_g2s_mc(smem_base + s * STAGE + rank * (S * SLICE), (rank ? srcB : srcA) + off, S * SLICE, full, (uint16_t)3);
        } else {
          bulk_g2s(smem_base + s * STAGE, srcA + off, S * SLICE, full);
          bulk_g2s(smem_base + s * STAGE + S * SLICE, srcB + off, S * SLICE, full);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: this CTA's weight groups, N = 128 ----------------
    if (lane == 0) {
      constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
      constexpr uint64_t DESC_HI = (1ull << 16) | ((uint64_t)((8 * OZ_KB) >> 4) << 32) | (1ull << 46) | (6ull << 61);
      // accumulator column of weight group d in this CTA: its rank among the CTA's groups, times 128
      uint32_t col_of[S];
      {
        uint32_t c = 0;
#pragma unroll
        for (int d = 0; d < S; ++d) {
          col_of[d] = c * TILE;
          if ((my_groups >> d) & 1u) ++c;
        }
      }
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % OZ2_STAGES;
        const uint32_t ph = (uint32_t)(it / OZ2_STAGES) & 1u;
        mbar_wait(bars + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * STAGE, sb = sa + S * SLICE;
        uint32_t touched = 0;
#pragma unroll
        for (int p = 0; p < S; ++p) {
          const uint64_t da = DESC_HI | (uint64_t)(((sa + p * SLICE) >> 4) & 0x3FFF);
#pragma unroll
          for (int q = 0; q < S - p; ++q) {
            const int d = p + q;
            if (!((my_groups >> d) & 1u)) continue;
            const uint64_t db = DESC_HI | (uint64_t)(((sb + q * SLICE) >> 4) & 0x3FFF);
            // an accumulator is overwritten by its first product of the tile and accumulated into afterwards
            const uint32_t acc = (it == 0 && !((touched >> d) & 1u)) ? 0u : 1u;
            touched |= 1u << d;
            umma_i8(tmem + col_of[d], da, db, IDESC, acc, 0);
          }
        }
        if (MC) umma_commit_mc(bars + 8 * (OZ2_STAGES + s), (uint16_t)3);  // the stage is free once BOTH CTAs have read it
        else umma_commit(bars + 8 * (OZ2_STAGES + s));
      }
      if (n_iter > 0) umma_commit(bars + 8 * (2 * OZ2_STAGES));
    }
  }
  // ---------------- epilogue ----------------
  double* mine = reinterpret_cast<double*>(smem_gen);          // [128][65]: my partial sums of the 64 columns I finalise
  double* inbox = mine + TILE * OZ2_EPI_LD;                    // [128][65]: the peer's partial sums of the same columns
  if (warp >= 2 && n_iter > 0) {
    mbar_wait(bars + 8 * (2 * OZ2_STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncwarp();
  __syncthreads();
  cluster_sync_all();  // both CTAs are past their main loops: the rings (and the peer's) are free to become epilogue buffers
  if (warp >= 2) {
    const int ew = warp - 2, q4 = warp & 3, ch = ew >> 2;  // TMEM lane quarter, 64-column half
    const int row = q4 * 32 + lane;
    // the half this warp handles stays here when it is mine (ch == rank), else it goes to the peer's inbox
    uint32_t dst_local = smem_u32(ch == rank ? mine : inbox) + (uint32_t)(row * OZ2_EPI_LD) * 8u;
    uint32_t dst = dst_local;
    if (ch != rank) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(dst_local), "r"((uint32_t)(rank ^ 1)));
    const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ch * 64);
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
      double sum[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) sum[j] = 0.0;
      if (n_iter > 0) {
        // Horner over this CTA's groups from the smallest weight up; the step between consecutive owned groups d' > d
        // is 256^-(d' - d), and the lowest owned group carries 256^-d_min at the end
        int prev = -1;
        uint32_t c = 0;
#pragma unroll
        for (int d = 0; d < S; ++d)
          if ((my_groups >> d) & 1u) ++c;
#pragma unroll
        for (int d = S - 1; d >= 0; --d) {
          if (!((my_groups >> d) & 1u)) continue;
          --c;
          uint32_t v[16];
          tmem_ld16(taddr + c * TILE + (uint32_t)c0, v);
          tmem_ld_wait();
          const double w = prev < 0 ? 0.0 : scalbn(1.0, -8 * (prev - d));
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[j] = fma(sum[j], w, i32_to_f64(v[j]));
          prev = d;
        }
        const double w0 = scalbn(1.0, -8 * prev);  // prev = lowest owned group
#pragma unroll
        for (int j = 0; j < 16; ++j) sum[j] *= w0;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst + (uint32_t)(c0 + j) * 8u), "d"(sum[j]) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // partial sums exchanged
  if (warp >= 2) {
    const int ew = warp - 2;
    const long long c_row = (long long)args.c_row0 + shift + ti * TILE;
    const long long c_col = (long long)args.c_col0 + shift + tj * TILE + rank * OZ_BN;
    const int bcol = b_row + rank * OZ_BN;  // B-operand rows = C columns of my half
    const double2 sb2 = *reinterpret_cast<const double2*>(args.B.scale + (long long)b * args.B.sc_batch + bcol + 2 * lane);
    const double* sA = args.A.scale + (long long)b * args.A.sc_batch + a_row;
    double* Cb = args.C + (long long)b * args.c_batch;
    const double alpha = args.alpha, beta = args.beta;
    for (int rr = ew; rr < TILE; rr += 8) {
      const double sa = sA[rr] * alpha;
      double2 v;
      v.x = (mine[rr * OZ2_EPI_LD + 2 * lane] + inbox[rr * OZ2_EPI_LD + 2 * lane]) * sa * sb2.x;
      v.y = (mine[rr * OZ2_EPI_LD + 2 * lane + 1] + inbox[rr * OZ2_EPI_LD + 2 * lane + 1]) * sa * sb2.y;
      double2* dstp = reinterpret_cast<double2*>(Cb + (c_row + rr) * args.ldc + c_col + 2 * lane);
      if (beta != 0.0) {
        const double2 old = *dstp;
        v.x = fma(beta, old.x, v.x);
        v.y = fma(beta, old.y, v.y);
      }
      *dstp = v;
    }
    if (args.Ct != nullptr) {
      const long long t_row = (long long)args.ct_row0 + shift + tj * TILE + rank * OZ_BN;
      const long long t_col = (long long)args.ct_col0 + shift + ti * TILE;
      double* Tb = args.Ct + (long long)b * args.c_batch;
      for (int cc = ew; cc < OZ_BN; cc += 8) {
        const double sbc = args.B.scale[(long long)b * args.B.sc_batch + bcol + cc] * alpha;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = lane + 32 * i;
          Tb[(t_row + cc) * args.ldc + t_col + r] = (mine[r * OZ2_EPI_LD + cc] + inbox[r * OZ2_EPI_LD + cc]) * sA[r] * sbc;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

// ---- slicing: FP64 region -> S int8 digit tiles (shared-memory image) + one scale per row --------------------------
// One CTA per 128-row block of the region.  Pass 1: row maxima over the K range (8 warps x 16 rows, coalesced row reads)
// -> e_i, 2^(8S - e_i) kept in shared memory, 2^(e_i - 8) written out for the GEMM epilogue.  Pass 2: per K chunk every
// thread converts 4 x 4 elements: X = rint(a 2^(8S-e)) as int64, bytes of (X + BIAS) ^ BIAS are the signed digits,
// gathered by PRMT into one 32-bit word per slice.
template <int S>
__global__ void __launch_bounds__(256) ozaki_slice_kernel(const OzSliceArgs a) {
  constexpr int CPB = TILE / OZ_KB;
  const int split = a.col_split > 1 ? a.col_split : 1;
  const int rbi = split > 1 ? (int)blockIdx.x % a.grid_rows : (int)blockIdx.x, part = split > 1 ? (int)blockIdx.x / a.grid_rows : 0;
  const int g = blockIdx.y, b = blockIdx.z;
  if (a.active != nullptr && a.active[b] == 0) return;
  int h2 = 0;
  if (a.grp_blocks > 0) {
    h2 = a.nblk - g * a.grp_blocks - a.half_blocks;
    h2 = h2 < 0 ? 0 : (h2 > a.half_blocks ? a.half_blocks : h2);
  }
  if (a.need_h2 && h2 <= 0) return;
  const int nrb = a.rows_kind == EXT_H2 ? h2 : a.row_blocks;
  const int ncc = a.cols_kind == EXT_H2 ? h2 * CPB : a.col_chunks;
  if (rbi >= nrb || ncc <= 0) return;
  const int shift = g * a.grp_blocks * TILE;
  int c_lo = 0, c_hi = ncc;
  if (a.tri == 1) c_hi = min(ncc, (rbi + 1) * CPB);
  if (a.tri == 2) c_lo = min(ncc, rbi * CPB);
  if (split > 1) {  // this CTA's share of the chunk range (the scales are given: no pass needs the whole row)
    const int per = (c_hi - c_lo + split - 1) / split;
    c_lo = min(c_hi, c_lo + part * per);
    c_hi = min(c_hi, c_lo + per);
    if (c_lo >= c_hi && part > 0) return;
  }
  const int row_base = a.row0 + shift + rbi * TILE, col_base = a.col0 + shift;
  const double* Xb = a.X + (long long)b * a.x_batch;
  __shared__ double s_mult[TILE];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (a.fixed_scale != nullptr) {
    // given scales 2^(e - 8): multiplier 2^(8S - e) = 2^(8S - 8) / scale, exact
    if (tid < TILE) s_mult[tid] = scalbn(1.0, 8 * S - 8) / a.fixed_scale[(long long)b * a.sc_batch + row_base + tid];
  } else if (a.given_max != nullptr) {
    // row maxima collected by the epilogue of the product that wrote X: no pass over the data for them
    if (tid < TILE) {
      const double amax = a.given_max[(long long)b * a.sc_batch + row_base + tid];
      const bool bad = !(amax <= 1.7976931348623157e308);
      const int e = amax > 0.0 && !bad ? ilogb(amax) + 3 : 0;
      s_mult[tid] = bad ? 0.0 : scalbn(1.0, 8 * S - e);
      if (part == 0) a.scale[(long long)b * a.sc_batch + row_base + tid] = bad ? __longlong_as_double(0x7ff8000000000000LL) : scalbn(1.0, e - 8);
    }
  } else
  for (int rr = 0; rr < 16; ++rr) {
    const int r = warp * 16 + rr;
    const double* p = Xb + (long long)(row_base + r) * a.ld + col_base;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;  // four independent chains: the loads of a row are all in flight
    int bad = 0;
    const int c_end = c_hi * OZ_KB;
    int c = c_lo * OZ_KB + 2 * lane;
    for (; c + 192 < c_end; c += 256) {
      const double2 v0 = *reinterpret_cast<const double2*>(p + c), v1 = *reinterpret_cast<const double2*>(p + c + 64);
      const double2 v2 = *reinterpret_cast<const double2*>(p + c + 128), v3 = *reinterpret_cast<const double2*>(p + c + 192);
      m0 = fmax(m0, fmax(fabs(v0.x), fabs(v0.y)));
      m1 = fmax(m1, fmax(fabs(v1.x), fabs(v1.y)));
      m2 = fmax(m2, fmax(fabs(v2.x), fabs(v2.y)));
      m3 = fmax(m3, fmax(fabs(v3.x), fabs(v3.y)));
      bad |= !(fabs(v0.x) + fabs(v0.y) + fabs(v1.x) + fabs(v1.y) + fabs(v2.x) + fabs(v2.y) + fabs(v3.x) + fabs(v3.y) <= 1.7976931348623157e308);
    }
    for (; c < c_end; c += 64) {
      const double2 v = *reinterpret_cast<const double2*>(p + c);
      m0 = fmax(m0, fmax(fabs(v.x), fabs(v.y)));
      bad |= !(fabs(v.x) + fabs(v.y) <= 1.7976931348623157e308);
    }
    double amax = fmax(fmax(m0, m1), fmax(m2, m3));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane == 0) {
      const int e = amax > 0.0 ? ilogb(amax) + 3 : 0;  // |a| 2^-e < 1/4
      double mult = scalbn(1.0, 8 * S - e), sc = scalbn(1.0, e - 8);
      if (bad) {  // a non-finite entry poisons the row, as it would in FP64
        mult = 0.0;
        sc = __longlong_as_double(0x7ff8000000000000LL);
      }
      s_mult[r] = mult;
      a.scale[(long long)b * a.sc_batch + row_base + r] = sc;
    }
  }
  __syncthreads();
  // pass 2.  A warp takes 4 rows x 32 columns per step: lane l -> row l / 8, columns 4 (l % 8) .. + 4 (32 contiguous
  // bytes in, 256 contiguous bytes per row); it writes ONE 32-bit word per slice, so a warp store is 128 contiguous
  // bytes of the tile image (the 32-byte swizzle only permutes 16-byte chunks inside such a block).
  constexpr unsigned long long BIAS = S == 7 ? 0x0080808080808080ull : S == 6 ? 0x0000808080808080ull : 0x8080808080808080ull;
  const int lr = lane >> 3, lc = (lane & 7) * 4;
  int8_t* tiles = a.S + (long long)b * a.s_batch +
                  ((long long)(row_base >> 7) * a.nkc + (col_base / OZ_KB)) * (S * (long long)(TILE * OZ_KB));
  for (int kc = c_lo; kc < c_hi; ++kc) {
    double2 v[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = warp * 16 + i * 4 + lr;
      const double* src = Xb + (long long)(row_base + r) * a.ld + col_base + kc * OZ_KB + lc;
      v[i][0] = *reinterpret_cast<const double2*>(src);
      v[i][1] = *reinterpret_cast<const double2*>(src + 2);
    }
    int8_t* dst = tiles + (long long)kc * (S * (long long)(TILE * OZ_KB));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = warp * 16 + i * 4 + lr;
      const double mult = s_mult[r];
      const double x[4] = {v[i][0].x, v[i][0].y, v[i][1].x, v[i][1].y};
      uint32_t lo[4], hi[4];
      if (a.fixed_scale != nullptr) {
        // no row-maximum pass looked at the data: a non-finite entry poisons the row's scale here, and an entry
        // above the bound saturates instead of wrapping around
        constexpr double LIM = (double)(1ull << (8 * S - 2)) * 1.9;
        bool bad = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) bad |= !(fabs(x[e]) <= 1.7976931348623157e308);
        if (bad) a.scale[(long long)b * a.sc_batch + row_base + r] = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double t = fmin(fmax(x[e] * mult, -LIM), LIM);
          const unsigned long long z = ((unsigned long long)__double2ll_rn(bad ? 0.0 : t) + BIAS) ^ BIAS;
          lo[e] = (uint32_t)z;
          hi[e] = (uint32_t)(z >> 32);
        }
      } else
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned long long z = ((unsigned long long)__double2ll_rn(x[e] * mult) + BIAS) ^ BIAS;
        lo[e] = (uint32_t)z;
        hi[e] = (uint32_t)(z >> 32);
      }
      const uint32_t off = oz_swz(r, lc);
#pragma unroll
      for (int p = 0; p < S; ++p) {
        const int k = S - 1 - p;  // byte k of z is digit p (p = 0 most significant)
        const uint32_t* srcw = k < 4 ? lo : hi;
        const uint32_t sel = (uint32_t)(k & 3) | ((uint32_t)(4 + (k & 3)) << 4);
        const uint32_t t0 = __byte_perm(srcw[0], srcw[1], sel), t1 = __byte_perm(srcw[2], srcw[3], sel);
        *reinterpret_cast<uint32_t*>(dst + p * (TILE * OZ_KB) + off) = __byte_perm(t0, t1, 0x5410);
      }
    }
  }
}

// scale[i] = 2^(e_i - 8), 2^e_i > 4 sqrt(A_ii) (A before factorisation): a bound on every entry of row i of its Cholesky factor
__global__ void oz_row_bound_kernel(const double* __restrict__ A, long long ld, long long a_batch, int n,
                                    double* __restrict__ scale, long long sc_batch, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = A[(long long)b * a_batch + (long long)i * ld + i];
  double sc = scalbn(1.0, -8);
  if (d > 0.0 && d <= 1.7976931348623157e308) sc = scalbn(1.0, ilogb(sqrt(d)) + 3 - 8);
  scale[(long long)b * sc_batch + i] = sc;  // (a non-positive or non-finite diagonal fails in potf2_inv anyway)
}
cudaError_t oz_row_bound_launch(const Launch& L, const double* A, long long ld, long long a_batch, int n, int batch,
                                double* scale, long long sc_batch, const int* active) {
  ProfScope prof_scope(L, CLS_SLICE);
  dim3 grid((n + 255) / 256, batch);
  oz_row_bound_kernel<<<grid, 256, 0, L.stream>>>(A, ld, a_batch, n, scale, sc_batch, active);
  L.bump();
  return cudaGetLastError();
}

// ---- launchers ---------------------------------------------------------------------------------------------------
template <int S, bool PAIR>
static cudaError_t oz_gemm_launch_t(const Launch& L, const OzGemmArgs& a, int tiles128, int groups, int batch) {
  bool& configured = func_attr_done(__LINE__ + 2 * S + (PAIR ? 1 : 0));
  if (!configured) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(ozaki_gemm_kernel<S, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz_smem_bytes<S>()));
    configured = true;
  }
  if (!PAIR) {
    dim3 grid(2 * tiles128, groups, batch);
    ozaki_gemm_kernel<S, false><<<grid, OZ_THREADS, oz_smem_bytes<S>(), L.stream>>>(a);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * tiles128, groups, batch);
  cfg.blockDim = dim3(OZ_THREADS);
  cfg.dynamicSmemBytes = oz_smem_bytes<S>();
  cfg.stream = L.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<S, PAIR>, a);
}

// int8 MACs one launch executes per batch item: the kernel's own tile / K-range rules, evaluated on the host
static double oz_executed_macs(const OzGemmArgs& a, int tiles128, int groups, int S) {
  constexpr int CPB = TILE / OZ_KB;
  double chunks = 0.0;  // sum over 128 x 64 tiles of their K chunks
  for (int g = 0; g < groups; ++g) {
    int h2 = 0;
    if (a.grp_blocks > 0) {
      h2 = a.nblk - g * a.grp_blocks - a.half_blocks;
      h2 = h2 < 0 ? 0 : (h2 > a.half_blocks ? a.half_blocks : h2);
    }
    const int mt = a.m_kind == EXT_H2 ? h2 : a.m_tiles, nt = a.n_kind == EXT_H2 ? h2 : a.n_tiles;
    const int kt = a.k_kind == EXT_H2 ? h2 * CPB : a.k_chunks;
    if (mt <= 0 || nt <= 0) continue;
    for (int ti = 0; ti < mt; ++ti) {
      const int tj_end = a.lower_only ? ti + 1 : (a.skip_upper ? (ti + 1 < nt ? ti + 1 : nt) : nt);
      if (!a.lower_only && ti * nt >= tiles128) break;
      for (int tj = 0; tj < tj_end; ++tj)
        for (int half = 0; half < 2; ++half) {
          int kb = 0, ke = kt;
          if (a.k_rule == KRULE_BEGIN_AT_ROW) kb = ti * CPB;
          if (a.k_rule == KRULE_END_AT_ROW) ke = ke < (ti + 1) * CPB ? ke : (ti + 1) * CPB;
          if (a.k_rule == KRULE_END_AT_COL) ke = ke < (2 * tj + half + 1) * (OZ_BN / OZ_KB) ? ke : (2 * tj + half + 1) * (OZ_BN / OZ_KB);
          if (ke > kb) chunks += ke - kb;
        }
    }
  }
  return chunks * (S * (S + 1) / 2) * (double)TILE * OZ_BN * OZ_KB;
}

template <int S, bool MC>
static cudaError_t oz_gemm128_launch_t(const Launch& L, const OzGemmArgs& a, int tiles128, int groups, int batch) {
  bool& configured = func_attr_done(__LINE__ + 2 * S + (MC ? 1 : 0));
  if (!configured) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(ozaki_gemm128_kernel<S, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz2_smem_bytes<S>()));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * tiles128, groups, batch);
  cfg.blockDim = dim3(OZ_THREADS);
  cfg.dynamicSmemBytes = oz2_smem_bytes<S>();
  cfg.stream = L.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ozaki_gemm128_kernel<S, MC>, a);
}

// Tile counters of the persistent kernel: {next tile, CTAs finished} per stream.  Launches on one stream run one after the
// other and the last CTA of each re-arms its pair, so one pair per (device, stream) is enough.
constexpr int OZP_STREAM_SLOTS = 1024;
__device__ int oz_tile_counters[OZP_STREAM_SLOTS][2];

namespace {
struct OzpSlot {
  int dev;
  cudaStream_t st;
  bool used;
};
OzpSlot ozp_slots[OZP_STREAM_SLOTS];
int* ozp_base[64] = {};
std::mutex ozp_mu;
}  // namespace

static int* ozp_counter_for(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(ozp_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (ozp_base[dev & 63] == nullptr) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, oz_tile_counters) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, sizeof(int) * 2 * OZP_STREAM_SLOTS) != cudaSuccess) return nullptr;
    ozp_base[dev & 63] = static_cast<int*>(p);
  }
  int free_slot = -1;
  for (int i = 0; i < OZP_STREAM_SLOTS; ++i) {
    if (ozp_slots[i].used && ozp_slots[i].dev == dev && ozp_slots[i].st == st) return ozp_base[dev & 63] + 2 * i;
    if (!ozp_slots[i].used && free_slot < 0) free_slot = i;
  }
  if (free_slot < 0) return nullptr;  // table full: the caller falls back to the one-tile-per-CTA kernel
  ozp_slots[free_slot] = OzpSlot{dev, st, true};
  return ozp_base[dev & 63] + 2 * free_slot;
}

// A stream that is about to be destroyed gives its counter pair back (agpc_destroy; the stream must be idle: its last
// launch re-armed the pair to zero).
void oz_release_stream(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(ozp_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < OZP_STREAM_SLOTS; ++i)
    if (ozp_slots[i].used && ozp_slots[i].dev == dev && ozp_slots[i].st == st) ozp_slots[i].used = false;
}

template <int S>
static cudaError_t oz_gemm_stream_launch_t(const Launch& L, const OzGemmArgs& a, int tiles128, int groups, int batch, int* counter) {
  bool& configured = func_attr_done(__LINE__ + S);
  if (!configured) {
    AGPC_CUDA_TRY(cudaFuncSetAttribute(ozaki_gemm_stream_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, ozp_smem_bytes<S>()));
    configured = true;
  }
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (sms[dev & 63] == 0) AGPC_CUDA_TRY(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = 2LL * tiles128 * groups * batch;
  const int grid = (int)(tiles < sms[dev & 63] ? tiles : sms[dev & 63]);
  ozaki_gemm_stream_kernel<S><<<grid, OZS_THREADS, ozp_smem_bytes<S>(), L.stream>>>(a, 2 * tiles128, groups, batch, counter);
  return cudaGetLastError();
}

cudaError_t oz_gemm_launch(const Launch& L, const OzGemmArgs& a, int tiles128, int groups, int batch, int S) {
  ProfScope prof_scope(L, L.gemm_cls);
  const int oz_cls = L.gemm_cls == CLS_GEMM_TRTRI ? CLS_OZ_TRTRI : L.gemm_cls == CLS_GEMM_LAUUM ? CLS_OZ_LAUUM : CLS_OZ_POTRF;
  ProfScope oz_scope(L, oz_cls);
  if (tiles128 <= 0 || groups <= 0 || batch <= 0) return cudaSuccess;
  if (L.prof && L.prof->enabled)
    L.add_work(oz_cls, 2.0 * oz_executed_macs(a, tiles128, groups, S) * (a.active ? (double)(L.work_items < batch ? L.work_items : batch) : (double)batch));
  L.bump();
  // AGPC_OZ_PAIR=1: 2-CTA multicast pairs.  Measured no faster (profiles/r02_summary.md: the kernel is bound by operand
  // delivery into the tensor core, not by L2 -> SM traffic), so single CTAs are the default.  Pairs need both halves of a
  // tile to run the same K range (KRULE_END_AT_COL clips per half).
  // AGPC_OZ_TILE=128: the 128 x 128 tile with its accumulators split over a CTA pair (ozaki_gemm128_kernel)
  static const int dbg_env = getenv("AGPC_OZ_DBG") ? atoi(getenv("AGPC_OZ_DBG")) : 0;
  if ((dbg_env == 3 || dbg_env >= 100) && batch <= OZP_MAX_BATCH) {
    OzGemmArgs d = a;
    d.dbg = dbg_env;
    int* counter = ozp_counter_for(L.stream);
    if (counter != nullptr)
      return S == 6 ? oz_gemm_stream_launch_t<6>(L, d, tiles128, groups, batch, counter) : oz_gemm_stream_launch_t<7>(L, d, tiles128, groups, batch, counter);
  }
  if (dbg_env) {
    OzGemmArgs d = a;
    d.dbg = dbg_env;
    return S == 6 ? oz_gemm_launch_t<6, false>(L, d, tiles128, groups, batch) : oz_gemm_launch_t<7, false>(L, d, tiles128, groups, batch);
  }
  // AGPC_OZ_PERSIST: 1 = persistent tile-streaming kernel (ozaki_gemm_stream_kernel) for every launch, 0 = one CTA per
  // tile, 2 = persistent except launches that run beside a look-ahead chain; unset: 1 for batches of >= 16 matrices (every
  // kernel of the chain fills the GPU by itself, resident CTAs cost the chain nothing), 2 below (the chain's small
  // kernels need the SMs that one-tile CTAs hand back as they finish).  profiles/r02_summary.md section 7.
  static const int persist_set = getenv("AGPC_OZ_PERSIST") ? atoi(getenv("AGPC_OZ_PERSIST")) : -1;
  const int persist_env = persist_set >= 0 ? persist_set : (batch >= 16 ? 1 : 2);
  if (persist_env && batch <= OZP_MAX_BATCH && !(persist_env == 2 && a.concurrent)) {
    int* counter = ozp_counter_for(L.stream);
    if (counter != nullptr)
      return S == 6 ? oz_gemm_stream_launch_t<6>(L, a, tiles128, groups, batch, counter) : oz_gemm_stream_launch_t<7>(L, a, tiles128, groups, batch, counter);
  }
  static const int tile_env = getenv("AGPC_OZ_TILE") ? atoi(getenv("AGPC_OZ_TILE")) : 64;
  if (tile_env == 128) return S == 6 ? oz_gemm128_launch_t<6, true>(L, a, tiles128, groups, batch) : oz_gemm128_launch_t<7, true>(L, a, tiles128, groups, batch);
  if (tile_env == 129) return S == 6 ? oz_gemm128_launch_t<6, false>(L, a, tiles128, groups, batch) : oz_gemm128_launch_t<7, false>(L, a, tiles128, groups, batch);
  static const bool pair_env = getenv("AGPC_OZ_PAIR") && atoi(getenv("AGPC_OZ_PAIR")) != 0;
  const bool pair = pair_env && a.k_rule != KRULE_END_AT_COL;
  if (S == 6) return pair ? oz_gemm_launch_t<6, true>(L, a, tiles128, groups, batch) : oz_gemm_launch_t<6, false>(L, a, tiles128, groups, batch);
  return pair ? oz_gemm_launch_t<7, true>(L, a, tiles128, groups, batch) : oz_gemm_launch_t<7, false>(L, a, tiles128, groups, batch);
}

// one warp per row: the extent of row i inside its G0-wide diagonal group (lower part for M, upper part for U)
__global__ void __launch_bounds__(256)
oz_rowmax_init_kernel(const double* __restrict__ M, const double* __restrict__ U, long long ld, long long batch_stride, int np,
                      int G0, double* __restrict__ rmM, double* __restrict__ rmU, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= np) return;
  const int g0 = (i / G0) * G0, g1 = min(g0 + G0, np);
  const double* mrow = M + (long long)b * batch_stride + (long long)i * ld;
  const double* urow = U + (long long)b * batch_stride + (long long)i * ld;
  double mm = 0.0, mu = 0.0, t = 0.0;
  for (int k = g0 + lane; k <= i; k += 32) {
    const double v = fabs(mrow[k]);
    mm = fmax(mm, v);
    t += v;
  }
  int bad_m = !(t <= 1.7976931348623157e308);
  t = 0.0;
  for (int k = i + lane; k < g1; k += 32) {
    const double v = fabs(urow[k]);
    mu = fmax(mu, v);
    t += v;
  }
  int bad_u = !(t <= 1.7976931348623157e308);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mm = fmax(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    mu = fmax(mu, __shfl_xor_sync(0xffffffffu, mu, o));
    bad_m |= __shfl_xor_sync(0xffffffffu, bad_m, o);
    bad_u |= __shfl_xor_sync(0xffffffffu, bad_u, o);
  }
  if (lane == 0) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    rmM[(long long)b * np + i] = bad_m ? nan : mm;
    rmU[(long long)b * np + i] = bad_u ? nan : mu;
  }
}
cudaError_t oz_rowmax_init_launch(const Launch& L, const double* M, const double* U, long long ld, long long batch_stride, int np,
                                  int G0, double* rmM, double* rmU, int batch, const int* active) {
  ProfScope prof_scope(L, CLS_SLICE);
  dim3 grid((np + 7) / 8, batch);
  oz_rowmax_init_kernel<<<grid, 256, 0, L.stream>>>(M, U, ld, batch_stride, np, G0, rmM, rmU, active);
  L.bump();
  return cudaGetLastError();
}

bool oz_gemm_rowmax_supported() {
  static const int tile_env = getenv("AGPC_OZ_TILE") ? atoi(getenv("AGPC_OZ_TILE")) : 64;
  static const int dbg_env = getenv("AGPC_OZ_DBG") ? atoi(getenv("AGPC_OZ_DBG")) : 0;
  return tile_env != 128 && tile_env != 129 && dbg_env == 0;
}

cudaError_t oz_slice_launch(const Launch& L, const OzSliceArgs& a, int row_blocks, int groups, int batch, int S) {
  ProfScope prof_scope(L, CLS_SLICE);
  if (row_blocks <= 0 || groups <= 0 || batch <= 0) return cudaSuccess;
  // Launches without a row-maximum pass are a chain of load -> convert -> store rounds per CTA, one per 32-column chunk:
  // with few CTAs they are latency-bound (a 256-column potrf panel: 30 us whatever the number of active items), so the
  // chunk range is cut into parts of >= 2 chunks until the grid has a few CTAs per SM slot.
  OzSliceArgs b = a;
  b.col_split = 0;
  static const bool split_env = !(getenv("AGPC_OZ_SLICE_SPLIT") && atoi(getenv("AGPC_OZ_SLICE_SPLIT")) == 0);
  if (split_env && (a.fixed_scale != nullptr || a.given_max != nullptr) && a.cols_kind == EXT_FIXED && a.col_chunks >= 4) {
    int split = 1;
    while (split * 2 * 2 <= a.col_chunks && (long long)row_blocks * groups * batch * split < 148LL * 3 * 4) split *= 2;
    if (split > 1) {
      b.col_split = split;
      b.grid_rows = row_blocks;
    }
  }
  dim3 grid(row_blocks * (b.col_split > 1 ? b.col_split : 1), groups, batch);
  if (S == 6)
    ozaki_slice_kernel<6><<<grid, 256, 0, L.stream>>>(b);
  else
    ozaki_slice_kernel<7><<<grid, 256, 0, L.stream>>>(b);
  L.bump();
  return cudaGetLastError();
}

}  // namespace agpc
