// Does IMAD.WIDE contend with DFMA?  Per group: 8 DFMA (r,r,const) + 8 integer statements of one kind.
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void __launch_bounds__(512) mix(double* out, int iters, double a, double b, unsigned k) {
  double x[8]; unsigned y[8];
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 2654435761u + i * 40503u; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (K != 9) x[i] = fma(x[i], a, b);
        if (K == 1 || K == 9) { unsigned long long p = (unsigned long long)y[i] * 0xD2511F53u; y[i] = (unsigned)p ^ (unsigned)(p >> 32) ^ k; }  // IMAD.WIDE + LOP3
        if (K == 2) { unsigned lo = y[i] * 0xD2511F53u, hi = __umulhi(y[i], 0xD2511F53u); asm volatile("" : "+r"(lo), "+r"(hi)); y[i] = lo ^ hi ^ k; }  // IMAD + IMAD.HI + LOP3
        if (K == 3) { y[i] = y[i] * 0x9E3779B9u + k; y[i] = y[i] * 0x85EBCA6Bu + k; }   // 2 IMAD
        if (K == 4) { y[i] = (y[i] ^ y[(i + 1) & 7] ^ k); y[i] = (y[i] ^ y[(i + 3) & 7] ^ 0x1234567u); }  // 2 LOP3
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += x[i] + y[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int K> void run(double* d, const char* name) {
  const int blocks = 148 * 4, iters = 500;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<K><<<blocks, 512>>>(d, 10, 0.999, 1e-3, 5); cudaDeviceSynchronize();
  cudaEventRecord(e0); mix<K><<<blocks, 512>>>(d, iters, 0.999, 1e-3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s %7.3f ms  %6.2f cycles per (DFMA + int statement) per SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / (16.0 * iters * 4 * 8));
}
int main() {
  double* d; cudaMalloc(&d, 148 * 4 * 512 * 8);
  run<0>(d, "DFMA only");
  run<9>(d, "IMAD.WIDE + LOP3 only");
  run<1>(d, "DFMA + IMAD.WIDE + LOP3");
  run<2>(d, "DFMA + IMAD.LO + IMAD.HI + LOP3");
  run<3>(d, "DFMA + 2 IMAD");
  run<4>(d, "DFMA + 2 LOP3");
  return 0;
}
