// Per-instruction cost table on sm_100a: cycles per warp-instruction per SMSP at full occupancy,
// for the instruction kinds the path kernels are made of.  16 warps/SMSP, 8 independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define REP 8
enum { DFMA_RRU, DFMA_RRR, DMUL_RR, IMADW, LOP3X, FSELX, I2F64S64, I2F64U64, I2F64S32, RSQ64H, SHFX, IADD3X, DSETPSEL, LDS128, FFMAX, MUFULG2, IMADLO, NKIND };
const char* NAMES[NKIND] = {"DFMA r,r,const", "DFMA r,r,r", "DMUL r,r", "IMAD.WIDE.U32 + LOP3", "LOP3 (xor3)", "FSEL", "I2F.F64.S64", "I2F.F64.U64", "I2F.F64.S32",
  "MUFU.RSQ64H", "SHF.L.W (funnel)", "IADD3", "DSETP+2xSEL", "LDS.128 random idx", "FFMA", "MUFU.LG2", "IMAD (lo)"};
template <int K>
__global__ void __launch_bounds__(512) kern(double* out, int iters, double a, double b, unsigned k, const double* tab) {
  __shared__ double s_tab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = tab[i];
  __syncthreads();
  double x[REP]; unsigned y[REP]; unsigned long long w[REP]; float f[REP];
  for (int i = 0; i < REP; i++) { x[i] = 1.0 + threadIdx.x * 1e-3 + i; y[i] = threadIdx.x * 2654435761u + i * 40503u; w[i] = y[i] * 0x9E3779B97F4A7C15ull; f[i] = 1.f + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < REP; i++) {
        if (K == DFMA_RRU) x[i] = fma(x[i], a, b);
        if (K == DFMA_RRR) x[i] = fma(x[i], x[(i + 1) % REP], x[(i + 3) % REP]);
        if (K == DMUL_RR) x[i] = x[i] * x[(i + 1) % REP];
        if (K == IMADW) { unsigned long long p = (unsigned long long)y[i] * 0xD2511F53u; y[i] = (unsigned)p ^ (unsigned)(p >> 32) ^ k; }
        if (K == LOP3X) y[i] = y[i] ^ y[(i + 1) % REP] ^ k;
        if (K == FSELX) y[i] = (y[(i + 1) % REP] & 1) ? y[i] : y[(i + 2) % REP];
        if (K == I2F64S64) { x[i] = (double)(long long)w[i]; w[i] += __double2hiint(x[i]); }
        if (K == I2F64U64) { x[i] = (double)w[i]; w[i] += __double2hiint(x[i]); }
        if (K == I2F64S32) { x[i] = (double)(int)y[i]; y[i] += __double2hiint(x[i]); }
        if (K == RSQ64H) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r; }
        if (K == SHFX) y[i] = __funnelshift_l(y[i], y[(i + 1) % REP], 21);
        if (K == IADD3X) y[i] = y[i] + y[(i + 1) % REP] + k;
        if (K == DSETPSEL) x[i] = (x[i] == a) ? b : x[(i + 1) % REP];
        if (K == LDS128) { double2 t = *(const double2*)&s_tab[2 * (y[i] & 127)]; y[i] = y[i] * 1664525u + __double2loint(t.x) + __double2loint(t.y); }
        if (K == FFMAX) f[i] = fmaf(f[i], f[(i + 1) % REP], 1.0001f);
        if (K == MUFULG2) f[i] = __log2f(f[i]);
        if (K == IMADLO) y[i] = y[i] * 0x9E3779B9u + y[(i + 1) % REP];
      }
    }
  }
  double s = 0;
  for (int i = 0; i < REP; i++) s += x[i] + y[i] + (double)w[i] + f[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int K> void run(double* d, const double* tab, int clock_khz) {
  const int blocks = 148 * 4, iters = 500;  // 4 blocks x 16 warps per SM = 16 warps per SMSP
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<K><<<blocks, 512>>>(d, 10, 0.999, 1e-3, 5, tab); cudaDeviceSynchronize();
  cudaEventRecord(e0); kern<K><<<blocks, 512>>>(d, iters, 0.999, 1e-3, 5, tab); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr_per_smsp = 16.0 * iters * 4 * REP;  // statements; each statement may be >1 SASS instruction
  double cyc = ms * 1e-3 * clock_khz * 1e3 / warp_instr_per_smsp;
  printf("%-22s %8.3f ms  %6.2f cycles per statement per SMSP\n", NAMES[K], ms, cyc);
}
int main() {
  double* d; cudaMalloc(&d, 148 * 4 * 512 * 8);
  double h[256]; for (int i = 0; i < 256; i++) h[i] = i * 1.5; double* tab; cudaMalloc(&tab, sizeof h); cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice);
  int khz = 1965000;
  run<DFMA_RRU>(d, tab, khz); run<DFMA_RRR>(d, tab, khz); run<DMUL_RR>(d, tab, khz); run<IMADW>(d, tab, khz); run<IMADLO>(d, tab, khz); run<LOP3X>(d, tab, khz); run<FSELX>(d, tab, khz);
  run<SHFX>(d, tab, khz); run<IADD3X>(d, tab, khz); run<I2F64S64>(d, tab, khz); run<I2F64U64>(d, tab, khz); run<I2F64S32>(d, tab, khz); run<RSQ64H>(d, tab, khz);
  run<DSETPSEL>(d, tab, khz); run<LDS128>(d, tab, khz); run<FFMAX>(d, tab, khz); run<MUFULG2>(d, tab, khz);
  return 0;
}
