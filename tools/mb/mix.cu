// Micro-benchmark: can the SMSP issue integer instructions in the shadow of FP64 instructions (2 pipe cycles each)?
// mode r: per loop iteration, 8 independent DFMA + r*8 independent integer ops (LOP3/IMAD mix).
#include <cstdio>
#include <cuda_runtime.h>
template <int R>
__global__ void __launch_bounds__(256) mix(double* out, unsigned* outi, int iters, double a, double b, unsigned k) {
  double x[8]; unsigned y[16];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int i = 0; i < 16; i++) y[i] = threadIdx.x * 7 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
#pragma unroll
      for (int i = 0; i < 8 * R; i++) y[i % 16] = (y[i % 16] ^ k) * 0x9E3779B9u + (y[(i + 1) % 16] >> 3);
    }
  }
  double s = 0; unsigned t = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  for (int i = 0; i < 16; i++) t ^= y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int R> __global__ void __launch_bounds__(256) intonly(unsigned* outi, int iters, unsigned k) {
  unsigned y[16];
  for (int i = 0; i < 16; i++) y[i] = threadIdx.x * 7 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8 * R; i++) y[i % 16] = (y[i % 16] ^ k) * 0x9E3779B9u + (y[(i + 1) % 16] >> 3);
  }
  unsigned t = 0;
  for (int i = 0; i < 16; i++) t ^= y[i];
  outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  const int blocks = 148 * 8, iters = 2000;
  double* d; unsigned* di; cudaMalloc(&d, blocks * 256 * 8); cudaMalloc(&di, blocks * 256 * 4);
  float t0 = timeit([&] { mix<0><<<blocks, 256>>>(d, di, iters, 0.999, 1e-3, 5); });
  float t1 = timeit([&] { mix<1><<<blocks, 256>>>(d, di, iters, 0.999, 1e-3, 5); });
  float t2 = timeit([&] { mix<2><<<blocks, 256>>>(d, di, iters, 0.999, 1e-3, 5); });
  float i1 = timeit([&] { intonly<1><<<blocks, 256>>>(di, iters, 5); });
  float i2 = timeit([&] { intonly<2><<<blocks, 256>>>(di, iters, 5); });
  printf("dfma only %.3f ms | dfma+int(1:1 stmts) %.3f | dfma+int(1:2) %.3f | int only x1 %.3f | int only x2 %.3f\n", t0, t1, t2, i1, i2);
  double flops = (double)blocks * 256 * iters * 4 * 8 * 2;
  printf("dfma TFLOP/s %.2f\n", flops / t0 / 1e9);
  return 0;
}
