// Probe: FP64 issue-rate ceilings on B200 (sm_100a): DFMA vs DMMA m8n8k4 vs m16n8k16.
// Register-only loops; reports TFLOP/s. Used to choose the Cholesky trailing-update instruction.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters) {
  double a[16]; double x = threadIdx.x * 1e-9, y = 1.0000001;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], y, x);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma884(double* out, int iters) {
  double c[16][2]; double a = threadIdx.x * 1e-9, b = 1.0000001;
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void k_dmma16816(double* out, int iters) {
  double c[8][4]; double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0000001 + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__global__ void k_dmma1688(double* out, int iters) {
  double c[8][4]; double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-9 + i;
  b[0] = 1.0000001; b[1] = 2.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32, blocks = sms * 2;
    double nthr = (double)blocks * threads, nwarp = nthr / 32;
    float ms;
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("warps/CTA=%2d x2 CTA/SM  DFMA        %8.2f TFLOP/s\n", warps, nthr * iters * 16 * 2 / ms / 1e9);
    ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
    printf("warps/CTA=%2d x2 CTA/SM  DMMA m8n8k4 %8.2f TFLOP/s\n", warps, nwarp * iters * 16 * (8 * 8 * 4 * 2.0) / ms / 1e9);
    ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
    printf("warps/CTA=%2d x2 CTA/SM  DMMA m16n8k8 %7.2f TFLOP/s\n", warps, nwarp * iters * 8 * (16 * 8 * 8 * 2.0) / ms / 1e9);
    ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters); });
    printf("warps/CTA=%2d x2 CTA/SM  DMMA m16n8k16 %6.2f TFLOP/s\n", warps, nwarp * iters * 8 * (16 * 8 * 16 * 2.0) / ms / 1e9);
  }
  CK(cudaGetLastError());
  return 0;
}
