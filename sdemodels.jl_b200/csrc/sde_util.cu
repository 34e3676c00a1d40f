// sde_util.cu — model-independent device kernels: the exact accumulator behind the payoff reduction,
// the row sub-sampler of _multilevel_simulate (src/multilevel.jl:95), and the FP64/FP32 peak
// micro-benchmarks that provide the roofline denominators MEASURED_PEAKS.json lacks.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sde_args.h"

// ---------------------------------------------------------------------------------------------------
// Exact accumulation of the per-CTA partial sums.  Each double is split into three 32-bit limbs of a
// 2176-bit fixed-point number (LSB = 2^-1074) and added with integer atomics into int64 bins, so the
// result does not depend on the order of the additions: 1, 2, 4 or 8 GPUs (ncclAllReduce of the bins with
// ncclSum on int64) give bit-identical sums.  bins[q][SDE_NBINS]; partials[chunk*stride + q].
// ---------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256) sde_accumulate_kernel(const double *partials, int64_t nchunks,
                                                                         int stride, int nq,
                                                                         unsigned long long *bins,
                                                                         unsigned long long *overflow) {
  __shared__ unsigned long long sb[2 * SDE_MAXF * SDE_NBINS];
  for (int i = threadIdx.x; i < nq * SDE_NBINS; i += blockDim.x) sb[i] = 0ull;
  __syncthreads();
  const int64_t total = nchunks * nq;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / nq;
    const int q = (int)(idx - c * nq);
    const unsigned long long b = (unsigned long long)__double_as_longlong(partials[c * stride + q]);
    const int e = (int)((b >> 52) & 0x7ff);
    unsigned long long mant = b & 0xFFFFFFFFFFFFFull;
    if (e == 0x7ff) {  // inf / nan partial: cannot be represented — flagged, the host reports an error
      atomicAdd(overflow, 1ull);
      continue;
    }
    if (mant == 0 && e == 0) continue;
    int sh = e - 1;
    if (e == 0) sh = 0; else mant |= (1ull << 52);
    const int bin = sh >> 5, off = sh & 31;
    const unsigned long long lo64 = mant << off;
    const unsigned long long hi = off ? (mant >> (64 - off)) : 0ull;
    unsigned long long l0 = lo64 & 0xffffffffull, l1 = lo64 >> 32, l2 = hi;
    if (b >> 63) { l0 = 0ull - l0; l1 = 0ull - l1; l2 = 0ull - l2; }  // two's complement adds
    unsigned long long *dst = sb + q * SDE_NBINS + bin;
    if (l0) atomicAdd(dst, l0);
    if (l1) atomicAdd(dst + 1, l1);
    if (l2) atomicAdd(dst + 2, l2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nq * SDE_NBINS; i += blockDim.x)
    if (sb[i]) atomicAdd(bins + i, sb[i]);
}

// xc = xf[1:substeps:end, :]  (multilevel.jl:95), reference layout [path][row][dim]
template <typename T>
__global__ void sde_subsample_kernel(const T *src, T *dst, int64_t npaths, int64_t nrows_dst, int64_t nrows_src,
                                     int64_t substeps, int dim) {
  const int64_t total = npaths * nrows_dst * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % dim);
    const int64_t r = (i / dim) % nrows_dst, k = i / (dim * nrows_dst);
    dst[i] = src[(k * nrows_src + r * substeps) * dim + d];
  }
}
extern "C" void sde_launch_subsample(int precision, const void *src, void *dst, int64_t npaths, int64_t nrows_dst,
                                     int64_t nrows_src, int64_t substeps, int dim, cudaStream_t st) {
  const int64_t total = npaths * nrows_dst * dim;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  if (precision == 0)
    sde_subsample_kernel<double><<<grid, 256, 0, st>>>((const double *)src, (double *)dst, npaths, nrows_dst,
                                                       nrows_src, substeps, dim);
  else
    sde_subsample_kernel<float><<<grid, 256, 0, st>>>((const float *)src, (float *)dst, npaths, nrows_dst,
                                                      nrows_src, substeps, dim);
}

extern "C" void sde_launch_accumulate(const double *partials, int64_t nchunks, int stride, int nq,
                                      unsigned long long *bins, unsigned long long *overflow, cudaStream_t st) {
  int64_t want = (nchunks * nq + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > 148 ? 148 : want));
  sde_accumulate_kernel<<<grid, 256, 0, st>>>(partials, nchunks, stride, nq, bins, overflow);
}

// ---------------------------------------------------------------------------------------------------
// Peak micro-benchmarks: 8 independent FMA chains per thread, 8 warps per SMSP.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) sde_peak_kernel(T *out, int iters, T a, T b) {
  T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// returns flops executed by one launch
extern "C" double sde_launch_peak(int precision, void *scratch, int blocks, int iters, cudaStream_t st) {
  if (precision == 0)
    sde_peak_kernel<double><<<blocks, 1024, 0, st>>>((double *)scratch, iters, 0.999999, 1e-6);
  else
    sde_peak_kernel<float><<<blocks, 1024, 0, st>>>((float *)scratch, iters, 0.999999f, 1e-6f);
  return (double)blocks * 1024.0 * (double)iters * 8.0 * 8.0 * 2.0;
}
