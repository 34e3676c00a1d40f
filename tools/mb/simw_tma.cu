// Prototype of the Wiener-driven trajectory kernel on TMA tensor tiles (cp.async.bulk.tensor.2d, 128-byte swizzle):
// a warp owns 32 paths ([path][row][D] order); one elected lane moves (32 paths) x (128 bytes of rows) tiles
// global -> shared -> global through a ring of NSTAGE stages (one mbarrier per stage), every lane integrates its own
// path in place in shared memory (conflict-free 16-byte accesses thanks to the swizzle).  The path pitch
// nrows*D*sizeof(T) is usually not a multiple of 16 bytes, which a tensor map needs, so g = 1, 2 or 4 consecutive
// paths form one row of the map and a tile is g boxes.  Rows past the last full tile and paths past the last full
// group go through plain loads / stores.  Checks bit-equality with a naive thread-per-path kernel.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
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

__device__ int g_timeout = 0;
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
    if (!done && ++spins > (1 << 20)) { if (!g_timeout) { g_timeout = 1; printf("mbarrier timeout: block %d thread %d bar %x parity %u\n", blockIdx.x, threadIdx.x, bar, parity); } return; }
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


__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tm), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

// ---- TMA kernel -------------------------------------------------------------------------------------------
template <typename T, int D, int NSTAGE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    simw_tma(const __grid_constant__ CUtensorMap tm, const T *w, T *xo, int nrows, int64_t npaths, int g, T dt) {
  constexpr int ROWB = D * (int)sizeof(T);
  constexpr int TS = 128 / ROWB;                        // rows per tile (128 bytes per path)
  constexpr int GB = ROWB > 16 ? ROWB : 16;             // bytes per register group (ROWB divides 128)
  constexpr int G = GB / ROWB, GQ = GB / 16, NG = 128 / GB;
  constexpr int TILE = 32 * 128;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // the swizzle pattern is a function of the address
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char *tiles = smem + (size_t)wid * (NSTAGE * TILE);
  uint64_t *bars = (uint64_t *)(smem + (size_t)WARPS * (NSTAGE * TILE)) + wid * NSTAGE;
  const int64_t wbase = ((int64_t)blockIdx.x * WARPS + wid) * 32;   // first path of this warp (a multiple of g)
  const int rpb = 32 / g;                                            // map rows per box
  const int64_t p = wbase + (int64_t)(lane % rpb) * g + lane / rpb;  // lane <-> tile row `lane`
  const int64_t ncov = (npaths / g) * g;                             // paths inside the tensor map
  const bool live = p < npaths, tiled = p < ncov;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; s++) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncwarp();
  if (wbase >= npaths) return;
  // box j (paths = j mod g) starts at the first row whose address is 16-byte aligned: TMA needs aligned box starts
  auto head = [&](int j) { int h = 0; while (((size_t)j * nrows * ROWB + (size_t)h * ROWB) % 16) h++; return h; };
  int hmax = 0;
  for (int j = 0; j < g; j++) hmax = max(hmax, head(j));
  const int ntile = nrows > hmax ? (nrows - hmax) / TS : 0;
  const int h = head(lane / rpb);
  const int c1 = (int)(wbase / g);
  auto issue_load = [&](int k) {   // lane 0
    if (k < ntile) {
      const int s = k % NSTAGE;
      const uint32_t bar = smem_u32(bars + s);
      mbar_expect_tx(bar, TILE);
      for (int j = 0; j < g; j++)
        tma_load_2d(smem_u32(tiles + s * TILE + j * rpb * 128), &tm, j * nrows * D + (head(j) + k * TS) * D, c1, bar);
    }
  };
  if (lane == 0)
    for (int k = 0; k < NSTAGE - 2; k++) issue_load(k);
  const T *wp = w + p * (int64_t)nrows * D;
  T *xp = xo + p * (int64_t)nrows * D;
  T x[D], wprev[D], dw[D];
#pragma unroll
  for (int d = 0; d < D; d++) { x[d] = x0v<T>(d); wprev[d] = (T)0; }
  auto plain_row = [&](int row) {
    if (row == 0) {
#pragma unroll
      for (int d = 0; d < D; d++) wprev[d] = wp[d];
    } else {
#pragma unroll
      for (int d = 0; d < D; d++) { const T wn = wp[(int64_t)row * D + d]; dw[d] = wn - wprev[d]; wprev[d] = wn; }
      Step<T, D>::go(dt, dw, x);
    }
#pragma unroll
    for (int d = 0; d < D; d++) xp[(int64_t)row * D + d] = x[d];
  };
  const int sw = lane & 7;
  if (tiled)
    for (int rowi = 0; rowi < h && rowi < nrows; rowi++) plain_row(rowi);   // head rows before the first aligned row
  for (int k = 0; k < ntile; k++) {
    if (lane == 0) {
      bulk_wait_read<1>();   // the stage window k + NSTAGE - 2 goes into was stored from at iteration k - 2
      issue_load(k + NSTAGE - 2);
    }
    const int s = k % NSTAGE;
    mbar_wait(smem_u32(bars + s), (uint32_t)((k / NSTAGE) & 1));
    unsigned char *row = tiles + s * TILE + lane * 128;
    if (tiled) {
#pragma unroll
      for (int gi = 0; gi < NG; gi++) {
        union { uint4 q[GQ]; T v[G * D]; } u;
#pragma unroll
        for (int c = 0; c < GQ; c++) u.q[c] = *(const uint4 *)(row + (((gi * GQ + c) ^ sw) << 4));
#pragma unroll
        for (int r = 0; r < G; r++) {
          if (k == 0 && gi == 0 && r == 0 && h == 0) {
#pragma unroll
            for (int d = 0; d < D; d++) wprev[d] = u.v[d];
          } else {
#pragma unroll
            for (int d = 0; d < D; d++) { const T wn = u.v[r * D + d]; dw[d] = wn - wprev[d]; wprev[d] = wn; }
            Step<T, D>::go(dt, dw, x);
          }
#pragma unroll
          for (int d = 0; d < D; d++) u.v[r * D + d] = x[d];
        }
#pragma unroll
        for (int c = 0; c < GQ; c++) *(uint4 *)(row + (((gi * GQ + c) ^ sw) << 4)) = u.q[c];
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      for (int j = 0; j < g; j++)
        tma_store_2d(&tm, j * nrows * D + (head(j) + k * TS) * D, c1, smem_u32(tiles + s * TILE + j * rpb * 128));
      bulk_commit();
    }
  }
  if (live) {
    if (tiled) {
      for (int rowi = max(h, 0) + ntile * TS; rowi < nrows; rowi++) plain_row(rowi);   // rows past the last full tile
    } else {
      for (int rowi = 0; rowi < nrows; rowi++) plain_row(rowi);            // paths past the last full group
    }
  }
  if (lane == 0) bulk_wait_read<0>();
}
// deterministic input
template <typename T> __global__ void fill(T *w, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t hsh = (uint32_t)i * 2654435761u ^ (uint32_t)(i >> 32) * 40503u;
    hsh ^= hsh >> 15; hsh *= 0x2c1b3c6du; hsh ^= hsh >> 12;
    w[i] = (T)((double)(hsh & 0xFFFFFF) / 16777216.0 - 0.5) * (T)0.08 + (T)(i & 1023) * (T)0.001;
  }
}


typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeFn)fn;
}

template <typename T, int D, int NSTAGE, int WARPS>
void run(const T *w0, T *work, const T *xref, int nrows, int64_t npaths, const char *name, int l2promo) {
  const size_t pitch = (size_t)nrows * D * sizeof(T);
  int g = 1;
  while ((g * pitch) % 16) g *= 2;
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)g * nrows * D, (cuuint64_t)(npaths / g)};
  const cuuint64_t strides[1] = {(cuuint64_t)g * pitch};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(T)), (cuuint32_t)(32 / g)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&tm, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)work, dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)l2promo,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  const size_t smem = (size_t)WARPS * NSTAGE * 4096 + (size_t)WARPS * NSTAGE * 8 + 1024;
  auto k = simw_tma<T, D, NSTAGE, WARPS>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WARPS * 32, smem));
  const size_t n = (size_t)nrows * npaths * D;
  const unsigned grid = (unsigned)((npaths + WARPS * 32 - 1) / (WARPS * 32));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemcpy(work, w0, n * sizeof(T), cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    k<<<grid, WARPS * 32, smem>>>(tm, work, work, nrows, npaths, g, (T)(1.0 / 512));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const size_t ncheck = std::min<size_t>(n, (size_t)8 << 20);
  std::vector<T> a(ncheck), b(ncheck);
  CK(cudaMemcpy(a.data(), work, ncheck * sizeof(T), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), xref, ncheck * sizeof(T), cudaMemcpyDeviceToHost));
  const bool same = memcmp(a.data(), b.data(), ncheck * sizeof(T)) == 0;
  const size_t tail = std::min<size_t>(n, 1 << 20);
  CK(cudaMemcpy(a.data(), work + (n - tail), tail * sizeof(T), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), xref + (n - tail), tail * sizeof(T), cudaMemcpyDeviceToHost));
  const bool same2 = memcmp(a.data(), b.data(), tail * sizeof(T)) == 0;
  printf("%-9s D=%d g=%d NSTAGE=%d WARPS=%d l2promo=%d smem/CTA=%6zu CTAs/SM=%2d warps/SM=%2d  %8.3f ms  %7.1f GB/s r+w  bitwise %s\n", name, D, g, NSTAGE,
         WARPS, l2promo, smem, occ, occ * WARPS, best, 2.0 * n * sizeof(T) / best / 1e6, (same && same2) ? "OK" : "MISMATCH");
  fflush(stdout);
}

template <typename T, int D> void suite(const char *name, int nrows, int64_t npaths) {
  const size_t n = (size_t)nrows * npaths * D + 16;
  T *w0, *work, *xref;
  CK(cudaMalloc(&w0, n * sizeof(T))); CK(cudaMalloc(&work, n * sizeof(T))); CK(cudaMalloc(&xref, n * sizeof(T)));
  fill<T><<<148 * 8, 256>>>(w0, (int64_t)n);
  naive<T, D><<<(unsigned)((npaths + 127) / 128), 128>>>(w0, xref, nrows, npaths, (T)(1.0 / 512));
  CK(cudaDeviceSynchronize());
  printf("# %s D=%d nrows=%d npaths=%lld\n", name, D, nrows, (long long)npaths);
  run<T, D, 3, 4>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 4, 4>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 4, 4>(w0, work, xref, nrows, npaths, name, 2);
  run<T, D, 6, 4>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 8, 4>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 4, 2>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 4, 8>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 6, 8>(w0, work, xref, nrows, npaths, name, 0);
  run<T, D, 12, 4>(w0, work, xref, nrows, npaths, name, 0);
  CK(cudaFree(w0)); CK(cudaFree(work)); CK(cudaFree(xref));
}

template <typename T, int D> void dbg(const char *name, int nrows, int64_t npaths) {
  const size_t n = (size_t)nrows * npaths * D + 16;
  T *w0, *work, *xref;
  CK(cudaMalloc(&w0, n * sizeof(T))); CK(cudaMalloc(&work, n * sizeof(T))); CK(cudaMalloc(&xref, n * sizeof(T)));
  fill<T><<<148 * 8, 256>>>(w0, (int64_t)n);
  naive<T, D><<<(unsigned)((npaths + 127) / 128), 128>>>(w0, xref, nrows, npaths, (T)(1.0 / 512));
  CK(cudaDeviceSynchronize());
  printf("# dbg %s D=%d nrows=%d npaths=%lld\n", name, D, nrows, (long long)npaths);
  run<T, D, 3, 4>(w0, work, xref, nrows, npaths, name, 0);
  CK(cudaFree(w0)); CK(cudaFree(work)); CK(cudaFree(xref));
}
int main(int argc, char **argv) {
  if (argc > 3) {
    dbg<double, 1>("f64even", 512, 128);
    dbg<double, 1>("f64", 513, 128);
    dbg<float, 1>("f32", 513, 128);
    dbg<double, 2>("f64", 513, 128);
    dbg<float, 2>("f32", 513, 100);
    dbg<float, 1>("f32", 7, 1000);
    dbg<double, 1>("f64", 2, 1000);
  }
  const int64_t npaths = argc > 1 ? atoll(argv[1]) : 2000000;
  const int nrows = argc > 2 ? atoi(argv[2]) : 513;
  suite<double, 1>("f64", nrows, npaths);
  suite<double, 1>("f64odd", nrows, npaths - 37);   // npaths not a multiple of 32 nor of g
  suite<double, 2>("f64", nrows, npaths / 2);
  suite<float, 1>("f32", nrows, npaths);
  suite<float, 1>("f32odd", nrows, npaths - 37);
  suite<float, 2>("f32", nrows, npaths / 2);
  suite<double, 1>("f64even", nrows - 1, npaths);
  suite<double, 1>("f64short", 37, npaths * 8);
  return 0;
}
