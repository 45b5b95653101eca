// Write-pattern probe: how fast can an n x n FP64 matrix be written when the writers are TH x TW tiles (the K build's
// pattern: lower tiles + mirrored upper tiles) compared with a linear fill?  nvcc -arch=sm_100a -O3 tile_write.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int TH, int TW, bool SYM>
__global__ void __launch_bounds__(256) tile_write(double* out, int n, int tiles_n) {
  int ti, tj;
  if (SYM) {
    const int t = blockIdx.x;
    ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    ti = blockIdx.x / tiles_n;
    tj = blockIdx.x % tiles_n;
  }
  const int row0 = ti * TH, col0 = tj * TW;
  constexpr int TPR = TW / 2;            // threads per row (double2 stores)
  constexpr int RPP = 256 / TPR;         // rows per pass
  const int tx = threadIdx.x % TPR, ty = threadIdx.x / TPR;
  const double v = (double)blockIdx.x;
  for (int r = ty; r < TH; r += RPP) {
    double2* o = reinterpret_cast<double2*>(out + (long long)(row0 + r) * n + col0) + tx;
    *o = make_double2(v, v + 1.0);
  }
  if (SYM && ti != tj) {
    // mirrored tile: rows col0.., cols row0.. (TW x TH)
    constexpr int TPR2 = TH / 2, RPP2 = 256 / TPR2;
    const int tx2 = threadIdx.x % TPR2, ty2 = threadIdx.x / TPR2;
    for (int r = ty2; r < TW; r += RPP2) {
      double2* o = reinterpret_cast<double2*>(out + (long long)(col0 + r) * n + row0) + tx2;
      *o = make_double2(v, v + 1.0);
    }
  }
}
template <int TH, int TW, bool SYM>
void run(double* d, int n, const char* name) {
  const int tr = n / TH, tc = n / TW;
  int grid = SYM ? tr * (tr + 1) / 2 : tr * tc;   // SYM assumes TH == TW
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) tile_write<TH, TW, SYM><<<grid, 256>>>(d, n, tc);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) tile_write<TH, TW, SYM><<<grid, 256>>>(d, n, tc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  printf("{\"probe\": \"tile_write\", \"pattern\": \"%s\", \"n\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", name, n, ms, 8.0 * n * n / ms / 1e6);
}
int main() {
  for (int n : {8192, 16384}) {
    double* d; cudaMalloc(&d, 8ull * n * n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaMemsetAsync(d, 0, 8ull * n * n);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemsetAsync(d, 0, 8ull * n * n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("{\"probe\": \"tile_write\", \"pattern\": \"memset\", \"n\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", n, ms, 8.0 * n * n / ms / 1e6);
    run<64, 64, false>(d, n, "64x64 full");
    run<64, 64, true>(d, n, "64x64 lower+mirror");
    run<128, 128, true>(d, n, "128x128 lower+mirror");
    run<32, 128, false>(d, n, "32x128 full");
    run<64, 128, false>(d, n, "64x128 full");
    run<16, 256, false>(d, n, "16x256 full");
    run<8, 512, false>(d, n, "8x512 full");
    cudaFree(d);
  }
  return 0;
}
