// Prototype of the Wiener-driven trajectory kernel on per-lane bulk async copies (cp.async.bulk, 1-D TMA):
// every lane owns one path ([path][row][D] order, rows contiguous), moves its own 16-byte-aligned row windows
// global -> shared -> global through a ring of NSTAGE buffers with one mbarrier per lane and stage, integrates in
// place in shared memory, and stores the window back with a bulk store.  Head rows (up to the first 16-byte
// boundary of the path) and the sub-16-byte leftover at the end go through plain loads / stores.
// Checks bit-equality with a naive thread-per-path kernel and times a sweep of (window bytes, stages, warps per CTA).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ---- the "model": strict (no contraction) Euler–Maruyama steps ---------------------------------------------
template <typename T, int D> struct Step;
template <> struct Step<double, 1> {  // BlackScholes
  static __device__ __forceinline__ void go(double dt, const double *dw, double *x) {
    const double mu = __dmul_rn(0.01, x[0]), sg = __dmul_rn(0.3, x[0]);
    x[0] = __dadd_rn(__dadd_rn(x[0], __dmul_rn(mu, dt)), __dmul_rn(sg, dw[0]));
  }
};
template <> struct Step<float, 1> {
  static __device__ __forceinline__ void go(float dt, const float *dw, float *x) {
    const float mu = __fmul_rn(0.01f, x[0]), sg = __fmul_rn(0.3f, x[0]);
    x[0] = __fadd_rn(__fadd_rn(x[0], __fmul_rn(mu, dt)), __fmul_rn(sg, dw[0]));
  }
};
template <> struct Step<double, 2> {  // Heston, literal form
  static __device__ __forceinline__ void go(double dt, const double *dw, double *x) {
    const double S = x[0], V = x[1], sv = sqrt(fabs(V));
    const double mu1 = __dmul_rn(0.01, S), mu2 = __dmul_rn(10.0, __dsub_rn(0.3, V));
    const double s11 = __dmul_rn(sv, S), s21 = __dmul_rn(__dmul_rn(0.1, sv), 0.4), s22 = __dmul_rn(__dmul_rn(0.1, sv), 0.9165151389911681);
    x[0] = __dadd_rn(__dadd_rn(S, __dmul_rn(mu1, dt)), __dmul_rn(s11, dw[0]));
    x[1] = __dadd_rn(__dadd_rn(V, __dmul_rn(mu2, dt)), __dadd_rn(__dmul_rn(s21, dw[0]), __dmul_rn(s22, dw[1])));
  }
};
template <> struct Step<float, 2> {
  static __device__ __forceinline__ void go(float dt, const float *dw, float *x) {
    const float S = x[0], V = x[1], sv = sqrtf(fabsf(V));
    const float mu1 = __fmul_rn(0.01f, S), mu2 = __fmul_rn(10.0f, __fsub_rn(0.3f, V));
    const float s11 = __fmul_rn(sv, S), s21 = __fmul_rn(__fmul_rn(0.1f, sv), 0.4f), s22 = __fmul_rn(__fmul_rn(0.1f, sv), 0.91651514f);
    x[0] = __fadd_rn(__fadd_rn(S, __fmul_rn(mu1, dt)), __fmul_rn(s11, dw[0]));
    x[1] = __fadd_rn(__fadd_rn(V, __fmul_rn(mu2, dt)), __fadd_rn(__fmul_rn(s21, dw[0]), __fmul_rn(s22, dw[1])));
  }
};
template <typename T> __device__ __forceinline__ T x0v(int d) { return d == 0 ? (T)100 : (T)0.4; }

// ---- naive reference: one thread per path, direct global accesses, out of place -------------------------------
template <typename T, int D>
__global__ void naive(const T *w, T *xo, int64_t nrows, int64_t npaths, T dt) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npaths) return;
  const T *wp = w + p * nrows * D;
  T *xp = xo + p * nrows * D;
  T x[D], wprev[D], dw[D];
  for (int d = 0; d < D; d++) { x[d] = x0v<T>(d); wprev[d] = wp[d]; xp[d] = x[d]; }
  for (int64_t i = 1; i < nrows; i++) {
    for (int d = 0; d < D; d++) { const T wn = wp[i * D + d]; dw[d] = wn - wprev[d]; wprev[d] = wn; }
    Step<T, D>::go(dt, dw, x);
    for (int d = 0; d < D; d++) xp[i * D + d] = x[d];
  }
}

// ---- PTX helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1 << 22)) { printf("mbarrier timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
__host__ __device__ constexpr int clcm(int a, int b) { return a / cgcd(a, b) * b; }

// ---- bulk kernel ------------------------------------------------------------------------------------------
// SEG: target bytes per lane and window; NSTAGE: ring depth (loads in flight = NSTAGE - 2); in place in shared memory.
template <typename T, int D, int SEG, int NSTAGE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) simw_bulk(const T *w, T *xo, int64_t nrows, int64_t npaths, T dt) {
  constexpr int ROWB = D * (int)sizeof(T);
  constexpr int UNIT = clcm(ROWB, 128);                 // window bytes are a multiple of 128: conflict-free 16-byte accesses
  constexpr int WB = UNIT * (SEG / UNIT > 0 ? SEG / UNIT : 1);
  constexpr int TS = WB / ROWB;                         // rows per window
  constexpr int GB = clcm(ROWB, 16);                    // bytes per register group
  constexpr int G = GB / ROWB, GQ = GB / 16;            // rows / 16-byte chunks per group
  constexpr int STRIDE = WB + 16;                       // lane stride in shared memory
  constexpr int E = 16 / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char *wbuf = smem + (size_t)wid * (NSTAGE * 32 * STRIDE);           // [stage][lane][STRIDE]
  uint64_t *bars = (uint64_t *)(smem + (size_t)WARPS * (NSTAGE * 32 * STRIDE)) + wid * (NSTAGE * 32);  // [stage][lane]
  const int64_t p = ((int64_t)blockIdx.x * WARPS + wid) * 32 + lane;
  const bool live = p < npaths;
#pragma unroll
  for (int s = 0; s < NSTAGE; s++) mbar_init(smem_u32(bars + s * 32 + lane), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  __syncwarp();
  if (!live) return;
  const T *wp = w + p * nrows * D;
  T *xp = xo + p * nrows * D;
  // head rows up to the first 16-byte boundary (same for w and x: the caller guarantees (w - x) % 16 == 0)
  const int a = (int)(((uintptr_t)wp / sizeof(T)) % E);
  int h = 0;
  while (((a + h * D) % E) != 0) h++;
  if (h > nrows) h = (int)nrows;
  T x[D], wprev[D], dw[D];
#pragma unroll
  for (int d = 0; d < D; d++) { x[d] = x0v<T>(d); wprev[d] = (T)0; }
  auto plain_row = [&](int64_t row) {
    if (row == 0) {
#pragma unroll
      for (int d = 0; d < D; d++) wprev[d] = wp[d];
    } else {
#pragma unroll
      for (int d = 0; d < D; d++) { const T wn = wp[row * D + d]; dw[d] = wn - wprev[d]; wprev[d] = wn; }
      Step<T, D>::go(dt, dw, x);
    }
#pragma unroll
    for (int d = 0; d < D; d++) xp[row * D + d] = x[d];
  };
  for (int i = 0; i < h; i++) plain_row(i);
  // windows: full ones, then one partial window of whole 16-byte chunks, then leftover rows
  const int64_t rem_rows = nrows - h;
  const int64_t nfull = rem_rows / TS;
  const int tail_rows = (int)(rem_rows - nfull * TS);
  const int tail_bytes = (tail_rows * ROWB) & ~(GB - 1);   // whole register groups only
  const int tail_grp_rows = tail_bytes / ROWB;
  const int64_t nwin = nfull + (tail_bytes > 0 ? 1 : 0);
  const int64_t nwin_max = (nrows / TS) + 1;               // uniform trip count (lanes differ by at most one window)
  auto win_bytes = [&](int64_t k) -> uint32_t { return k < nfull ? (uint32_t)WB : (k < nwin ? (uint32_t)tail_bytes : 0u); };
  auto issue_load = [&](int64_t k) {
    const uint32_t nb = win_bytes(k);
    if (nb) {
      const int s = (int)(k % NSTAGE);
      const uint32_t bar = smem_u32(bars + s * 32 + lane);
      mbar_expect_tx(bar, nb);
      bulk_g2s(smem_u32(wbuf + (size_t)(s * 32 + lane) * STRIDE), wp + (h + k * TS) * D, nb, bar);
    }
  };
#pragma unroll
  for (int k = 0; k < NSTAGE - 2; k++) issue_load(k);
  for (int64_t k = 0; k < nwin_max; k++) {
    // the stage that window k + NSTAGE - 2 goes into was stored from at iteration k - 2: its read must be done
    bulk_wait_read<1>();
    issue_load(k + NSTAGE - 2);
    const uint32_t nb = win_bytes(k);
    if (nb) {
      const int s = (int)(k % NSTAGE);
      mbar_wait(smem_u32(bars + s * 32 + lane), (uint32_t)((k / NSTAGE) & 1));
      uint4 *buf = (uint4 *)(wbuf + (size_t)(s * 32 + lane) * STRIDE);
      const int ngrp = (int)(nb / GB);
      const int64_t row0 = h + k * TS;
      for (int g = 0; g < ngrp; g++) {
        union { uint4 q[GQ]; T v[G * D]; } u;
#pragma unroll
        for (int c = 0; c < GQ; c++) u.q[c] = buf[g * GQ + c];
#pragma unroll
        for (int r = 0; r < G; r++) {
          if (row0 + g * G + r == 0) {
#pragma unroll
            for (int d = 0; d < D; d++) wprev[d] = u.v[r * D + d];
          } else {
#pragma unroll
            for (int d = 0; d < D; d++) { const T wn = u.v[r * D + d]; dw[d] = wn - wprev[d]; wprev[d] = wn; }
            Step<T, D>::go(dt, dw, x);
          }
#pragma unroll
          for (int d = 0; d < D; d++) u.v[r * D + d] = x[d];
        }
#pragma unroll
        for (int c = 0; c < GQ; c++) buf[g * GQ + c] = u.q[c];
      }
      fence_async_smem();
      bulk_s2g(xp + row0 * D, smem_u32(buf), nb);
    }
    bulk_commit();
  }
  for (int64_t row = h + nfull * TS + tail_grp_rows; row < nrows; row++) plain_row(row);
  bulk_wait_read<0>();
}

// deterministic input
template <typename T> __global__ void fill(T *w, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t hsh = (uint32_t)i * 2654435761u ^ (uint32_t)(i >> 32) * 40503u;
    hsh ^= hsh >> 15; hsh *= 0x2c1b3c6du; hsh ^= hsh >> 12;
    w[i] = (T)((double)(hsh & 0xFFFFFF) / 16777216.0 - 0.5) * (T)0.08 + (T)(i & 1023) * (T)0.001;
  }
}

template <typename T, int D, int SEG, int NSTAGE, int WARPS>
void run(const T *w0, T *work, const T *xref, int64_t nrows, int64_t npaths, const char *name, int off) {
  constexpr int ROWB = D * (int)sizeof(T);
  constexpr int UNIT = clcm(ROWB, 128);
  constexpr int WB = UNIT * (SEG / UNIT > 0 ? SEG / UNIT : 1);
  const size_t smem = (size_t)WARPS * NSTAGE * 32 * (WB + 16) + (size_t)WARPS * NSTAGE * 32 * 8;
  auto k = simw_bulk<T, D, SEG, NSTAGE, WARPS>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WARPS * 32, smem));
  const size_t n = (size_t)nrows * npaths * D;
  const unsigned grid = (unsigned)((npaths + WARPS * 32 - 1) / (WARPS * 32));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemcpy(work + off, w0 + off, n * sizeof(T), cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    k<<<grid, WARPS * 32, smem>>>(work + off, work + off, nrows, npaths, (T)(1.0 / 512));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  // bitwise check against the naive kernel
  const size_t ncheck = std::min<size_t>(n, (size_t)8 << 20);
  std::vector<T> a(ncheck), b(ncheck);
  CK(cudaMemcpy(a.data(), work + off, ncheck * sizeof(T), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), xref, ncheck * sizeof(T), cudaMemcpyDeviceToHost));
  const bool same = memcmp(a.data(), b.data(), ncheck * sizeof(T)) == 0;
  // and the very end of the buffer
  const size_t tail = std::min<size_t>(n, 1 << 20);
  CK(cudaMemcpy(a.data(), work + off + (n - tail), tail * sizeof(T), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), xref + (n - tail), tail * sizeof(T), cudaMemcpyDeviceToHost));
  const bool same2 = memcmp(a.data(), b.data(), tail * sizeof(T)) == 0;
  printf("%-10s D=%d SEG=%4d WB=%4d NSTAGE=%d WARPS=%d smem/CTA=%6zu CTAs/SM=%d warps/SM=%2d  %8.3f ms  %7.1f GB/s r+w  bitwise %s\n", name, D, SEG,
         WB, NSTAGE, WARPS, smem, occ, occ * WARPS, best, 2.0 * n * sizeof(T) / best / 1e6, (same && same2) ? "OK" : "MISMATCH");
  fflush(stdout);
}

template <typename T, int D> void suite(const char *name, int64_t nrows, int64_t npaths, int off) {
  const size_t n = (size_t)nrows * npaths * D + 16;
  T *w0, *work, *xref;
  CK(cudaMalloc(&w0, n * sizeof(T))); CK(cudaMalloc(&work, n * sizeof(T))); CK(cudaMalloc(&xref, n * sizeof(T)));
  fill<T><<<148 * 8, 256>>>(w0, (int64_t)n);
  naive<T, D><<<(unsigned)((npaths + 127) / 128), 128>>>(w0 + off, xref, nrows, npaths, (T)(1.0 / 512));
  CK(cudaDeviceSynchronize());
  // ceiling: device-to-device copy of the same bytes
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemcpy(work, w0, n * sizeof(T), cudaMemcpyDeviceToDevice));
  CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(work, w0, n * sizeof(T), cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("# %s D=%d nrows=%lld npaths=%lld off=%d : cudaMemcpy D2D %.3f ms = %.1f GB/s r+w\n", name, D, (long long)nrows, (long long)npaths, off, ms, 2.0 * n * sizeof(T) / ms / 1e6);
  run<T, D, 128, 4, 4>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 256, 3, 4>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 256, 4, 4>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 256, 4, 2>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 256, 5, 2>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 256, 6, 2>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 512, 3, 2>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 512, 4, 2>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 512, 4, 1>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 512, 6, 1>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 1024, 4, 1>(w0, work, xref, nrows, npaths, name, off);
  run<T, D, 1024, 3, 2>(w0, work, xref, nrows, npaths, name, off);
  CK(cudaFree(w0)); CK(cudaFree(work)); CK(cudaFree(xref));
}

int main(int argc, char **argv) {
  const int64_t npaths = argc > 1 ? atoll(argv[1]) : 2000000;
  const int64_t nrows = argc > 2 ? atoll(argv[2]) : 513;
  suite<double, 1>("f64", nrows, npaths, 0);
  suite<double, 1>("f64+8B", nrows, npaths, 1);   // base pointer 8 mod 16
  suite<double, 2>("f64", nrows, npaths / 2, 0);
  suite<float, 1>("f32", nrows, npaths, 0);
  suite<float, 1>("f32+4B", nrows, npaths, 1);
  suite<float, 2>("f32", nrows, npaths / 2, 0);
  suite<double, 1>("f64short", 37, npaths * 8, 0);
  return 0;
}
