// sde_device.cuh — device layer of sdeb200: the Monte Carlo path-simulation hot path of SDEModels.jl
// (src/simulation.jl:39-48,59-67,104-134; src/schemes/euler_maruyama.jl:1-6; src/schemes/milstein.jl:4-10;
// src/multilevel.jl:65-89) written B200-first: one thread per path, state in registers, a counter-based
// Philox4x32-10 normal generator fused into the step, warp-shuffle + CTA reduction of payoff sums,
// coalesced / shared-memory-staged stores when full trajectories are kept.
//
// The same header is compiled ahead of time by nvcc (built-in BlackScholes / Heston functors) and at run
// time by NVRTC (functors emitted by the @sde_model code generator, sde_dsl.cpp), and — math helpers only —
// by g++ for CPU unit tests of the generator (tests/test_device_math.py).  It therefore includes nothing
// but its own siblings.
#pragma once
#include "sde_args.h"
#include "sde_coeffs.inc"

#if !defined(__CUDACC_RTC__)
#include <math.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define SDE_HD __host__ __device__ __forceinline__
#define SDE_D __device__ __forceinline__
#else
#define SDE_HD inline
#endif

#ifndef SDE_V2_UNROLL
#define SDE_V2_UNROLL 1 /* batches per iteration of the stream-v2 terminal kernel's time loop (2: 285.3 against 294.5 G path-steps/s, profiles/r02zz_ab_unroll.txt) */
#endif
#ifndef SDE_PHILOX_ROUNDS
#define SDE_PHILOX_ROUNDS 10 /* Philox4x32-10, the Random123 / cuRAND default; other values are tuning experiments only */
#endif
#ifndef SDE_SOA_PP
#define SDE_SOA_PP 0 /* [time][D][path] trajectory kernel: per-path Philox form (1: measured 2.5 % slower, 16 more registers) or the plain block function (0) */
#endif
#ifndef SDE_STRICT
#define SDE_STRICT 0 /* 1: reference evaluation order, translation unit compiled with -fmad=false */
#endif

namespace sdeb200 {

// FP64 polynomial coefficients: one __constant__ array on the device (ptxas keeps them in uniform registers
// as DFMA operands instead of rebuilding 64-bit immediates with two moves per use), literals on the host.
#if defined(__CUDACC__)
static __constant__ double sde_kd[SDE_K_N] = SDE_K_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define SDE_K(name) sdeb200::sde_kd[name##_I]
#else
#define SDE_K(name) (name##_V)
#endif

// ---------------------------------------------------------------------------------------------------
// bit helpers (device intrinsics with host twins so the transform can be unit-tested on the CPU)
// ---------------------------------------------------------------------------------------------------
SDE_HD double hilo2double(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi, (int)lo);
#else
  uint64_t b = ((uint64_t)hi << 32) | lo;
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
SDE_HD uint32_t double2hi(double d) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2hiint(d);
#else
  uint64_t b;
  memcpy(&b, &d, 8);
  return (uint32_t)(b >> 32);
#endif
}
SDE_HD uint32_t double2lo(double d) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2loint(d);
#else
  uint64_t b;
  memcpy(&b, &d, 8);
  return (uint32_t)b;
#endif
}
SDE_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}
SDE_HD float bits2float(uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
// ~2^-22-accurate reciprocal square root seed (MUFU.RSQ64H: reads/writes the high word only)
SDE_HD double rsqrt_seed(double a) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return y;
#else
  double y = 1.0 / sqrt(hilo2double(double2hi(a), 0));
  return hilo2double(double2hi(y), 0);
#endif
}

// sqrt by one Goldschmidt iteration + one correction on the MUFU.RSQ64H seed: 6 FP64-pipe instructions, no
// slow-path branch.  <= 1 ulp for positive normal v.
SDE_HD double sqrt_from_seed(double v, double y) {
  double g = v * y;
  double h = y * 0.5; /* exact; measured 0.8 % faster than building y/2 with an integer subtract on the high word (which
                         needs a second zeroed low-word register per seed) */
  const double e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double d = fma(-g, g, v);
  return fma(d, h, g);
}
SDE_HD double sqrt_core(double v) { return sqrt_from_seed(v, rsqrt_seed(v)); }  // v > 0
// Model-side sqrt of the FAST math mode: sqrt(0) = 0 (the +inf seed is clamped to 2^1023 with one integer min, so
// g = 0 * 2^1023 = 0 exactly); v < 0 gives a non-finite result (NaN or -inf) that propagates to the state and is
// counted as a non-finite path — the reference throws DomainError there.  Subnormal v is flushed.
#ifndef SDE_V2_IMAD
#define SDE_V2_IMAD 1
#endif
#ifndef SDE_V2_ROT1
#define SDE_V2_ROT1 1
#endif
#ifndef SDE_SIMRT2_DIAG
#define SDE_SIMRT2_DIAG 0
#endif
#ifndef SDE_SQRT_FAST6
#define SDE_SQRT_FAST6 1 /* six-operation form (sqrt_from_seed6, <= 1 ulp) instead of the seven-operation one */
#endif
SDE_HD double sqrt_from_seed6(double v, double y);
SDE_HD double sqrt_fast(double v) {
  const double y = rsqrt_seed(v);
  const int32_t hy = (int32_t)double2hi(y);
  const double yc = hilo2double((uint32_t)(hy < 0x7FE00000 ? hy : 0x7FE00000), 0);
#if SDE_SQRT_FAST6
  return sqrt_from_seed6(v, yc);
#else
  return sqrt_from_seed(v, yc);
#endif
}
SDE_HD float sqrt_fast(float v) {  // FAST math, FP32: one MUFU.SQRT (max relative error 2^-23) instead of the IEEE sequence
#if defined(__CUDA_ARCH__)
  float g;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(v));
  return g;
#else
  return sqrtf(v);
#endif
}
// sqrt as emitted by the @sde_model code generator: IEEE in STRICT math, the branch-free form in FAST math
#if SDE_STRICT
#define SDE_SQRT(x) sqrt(x)
#else
#define SDE_SQRT(x) sdeb200::sqrt_fast(x)
#endif

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (path_lo, path_hi, block, stream), key = seed.
// The key schedule is loop-invariant per launch; ptxas keeps the ten bumped keys in uniform registers.
// ---------------------------------------------------------------------------------------------------
SDE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t (&r)[4]) {
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < SDE_PHILOX_ROUNDS; i++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// The same block function with the key schedule expanded once per thread: the 2 x ROUNDS round keys sit in
// ordinary registers for the whole time loop (left to itself ptxas re-derives them every iteration with ~14
// uniform-datapath adds, which cost issue slots like any other instruction).
struct PhiloxKeys {
  uint32_t k[2 * SDE_PHILOX_ROUNDS];
};
SDE_HD void philox_expand(uint32_t k0, uint32_t k1, PhiloxKeys &ks) {
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < SDE_PHILOX_ROUNDS; i++) {
    ks.k[2 * i] = k0 + (uint32_t)i * 0x9E3779B9u;
    ks.k[2 * i + 1] = k1 + (uint32_t)i * 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(ks.k[2 * i]), "+r"(ks.k[2 * i + 1]));  // opaque: keep them, do not rematerialise
#endif
  }
}
SDE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &ks, uint32_t (&r)[4]) {
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < SDE_PHILOX_ROUNDS; i++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks.k[2 * i];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks.k[2 * i + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// The same block function for a FIXED path (counter words c0 = path_lo, c1 = path_hi, c3 = stream) and a varying
// block index c2: everything of rounds 0 and 1 that depends on the path only is computed once per path —
// M0*path_lo (round 0) and M1*(hi(M0*path_lo) ^ stream ^ k1) (round 1) — which leaves 17 instead of 18 per-thread
// 32x32->64 products per block (the product with the block index is warp-uniform).  Bit-identical outputs.
struct PhiloxPath {
  uint32_t a, b, c, e;   // path_hi ^ k0;  hi(Q) ^ k2;  lo(Q);  lo(M0*path_lo) ^ k3   with Q = M1 * n2(round 0)
};
SDE_HD void philox_path_init(uint64_t path, uint32_t stream, const PhiloxKeys &ks, PhiloxPath &pp, bool used = true) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * (uint32_t)path;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ stream ^ ks.k[1];
  const uint64_t q = (uint64_t)0xCD9E8D57u * n2;
  pp.a = (uint32_t)(path >> 32) ^ ks.k[0];
  pp.b = (uint32_t)(q >> 32) ^ ks.k[2];
  pp.c = (uint32_t)q;
  pp.e = (uint32_t)p0 ^ ks.k[3];
#if defined(__CUDA_ARCH__)
  if (used) asm volatile("" : "+r"(pp.a), "+r"(pp.b), "+r"(pp.c), "+r"(pp.e));  // keep them, do not rematerialise (FP32: dead)
#endif
}
SDE_HD void philox4x32_10(const PhiloxPath &pp, uint32_t blk, const PhiloxKeys &ks, uint32_t (&r)[4]) {
  static_assert(SDE_PHILOX_ROUNDS >= 2, "rounds 0 and 1 are peeled");
  const uint64_t u = (uint64_t)0xCD9E8D57u * blk;                 // round 0, block half (warp-uniform)
  const uint32_t n0 = (uint32_t)(u >> 32) ^ pp.a;
  const uint64_t t = (uint64_t)0xD2511F53u * n0;                  // round 1, the half that depends on the block
  uint32_t c0 = pp.b ^ (uint32_t)u, c1 = pp.c, c2 = (uint32_t)(t >> 32) ^ pp.e, c3 = (uint32_t)t;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 2; i < SDE_PHILOX_ROUNDS; i++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t m0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks.k[2 * i];
    const uint32_t m2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks.k[2 * i + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = m0;
    c2 = m2;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// ---------------------------------------------------------------------------------------------------
// "sdeb200 normal stream v1", FP64: 128 Philox bits -> two N(0, dt) variates (Box–Muller).
//   X1 = r1:r0  -> u1 = RN(X1) 2^-64 = m 2^-k, m in [1,2)     (one I2F.F64.U64; X1 == 0 reads as u1 = 2^-1087)
//   X2 = r3:r2  -> angle = X2 / 2^64 turns = n/4 + x, |x| <= 1/8 (nearest quarter turn by sign extension)
//   w = dt (-2 ln u1 + 2^-50) = k (2 ln2 dt) + r Q_dt(r) + T_j,   r = m/c_j - 1, |r| <= 2^-8   (128-entry table)
//       the 2^-50 keeps the computed w > 0 when u1 rounds to 1 (no clamp); dt is folded into the constants
//   radius sqrt(w): Goldschmidt on the MUFU.RSQ64H seed; (cos, sin)(2 pi X2 / 2^64): 1024-entry table of cell
//       centres + a two-term residual rotation in t^2, t the integer offset itself (scaled coefficients, exact).
// 25 FP64-pipe instructions per pair; the rest is a handful of integer bit-slicing instructions.
//   lk[6]  = dt * {Q1, Q2, Q3, Q4, -2, 2 ln 2}   (make_log_consts; kernel-parameter space on the device)
//   tab    = [0,256): {1/c_j, dt (2 ln(1/c_j) + 2^-50)} pairs; [256, 2304): {cos, sin} of the 1024 cell centres
//            (shared memory on the device, load_tables)
// ---------------------------------------------------------------------------------------------------
SDE_HD void make_log_consts(double dt, double *lk) {
  lk[0] = SDE_LOG_C1_V * dt;
  lk[1] = SDE_LOG_C2_V * dt;
  lk[2] = SDE_LOG_C3_V * dt;
  lk[3] = SDE_LOG_C4_V * dt;
  lk[4] = -2.0 * dt;
  lk[5] = SDE_TWO_LN2_V * dt;
}

// kr (optional): register-resident copies of the three leading polynomial coefficients {lk[3], ROT_S2, ROT_C2}; a DFMA takes
// one uniform-register operand only, so a leading coefficient read from constant memory costs two moves per use.
// Polar form: radius^2 `w` (already scaled by the step's variance) and the unit vector (ca, sb); the pair is sqrt(w) (ca, sb).
SDE_HD void normal_polar_f64(const uint32_t (&r)[4], const double *tab, const double *lk, double &w, double &ca, double &sb,
                             const double *kr = nullptr) {
  const uint64_t X1 = ((uint64_t)r[1] << 32) | r[0];
  const double d1 = (double)X1;
  const uint32_t h1 = double2hi(d1);
#if defined(__CUDA_ARCH__)
  uint32_t one_hi, mh;   // (h1 & 0xFFFFF) | 0x3FF00000 as ONE three-input LOP3: the second constant sits in a register
  asm("mov.u32 %0, 0x3FF00000;" : "=r"(one_hi));
  asm("lop3.b32 %0, %1, 0x000FFFFF, %2, 0xEA;" : "=r"(mh) : "r"(h1), "r"(one_hi));
  const double m = hilo2double(mh, double2lo(d1));
#else
  const double m = hilo2double((h1 & 0x000FFFFFu) | 0x3FF00000u, double2lo(d1));
#endif
  const int j = (int)((h1 >> 13) & 0x7Fu);
  const double kd = (double)(int)(1087u - (h1 >> 20));
  const double inv_c = tab[2 * j], tl = tab[2 * j + 1];
  const double rr = fma(m, inv_c, -1.0);
  double q = kr ? kr[0] : lk[3];
  q = fma(q, rr, lk[2]);
  q = fma(q, rr, lk[1]);
  q = fma(q, rr, lk[0]);
  q = fma(q, rr, lk[4]);
  w = fma(rr, q, tl);
  w = fma(kd, lk[5], w);
  // angle = X2 / 2^64 turns = (i + 1/2 + d) / 1024: the top 10 bits pick {cos, sin} of the cell centre from a
  // 1024-entry table, the low 54 bits (re-centred) are the offset t = 2^54 d, |t| <= 2^53 (exact in a double);
  // the residual rotation |phi| <= pi/1024 needs two-term polynomials; no quadrant logic.
  const double *rot = tab + 256 + 2 * (r[3] >> 22);
  const double C = rot[0], S = rot[1];
  const int64_t f = (int64_t)(((uint64_t)((r[3] & 0x003FFFFFu) - 0x00200000u) << 32) | r[2]);
  const double t = (double)f;
  const double u = t * t;
  double sp = fma(kr ? kr[1] : SDE_K(SDE_ROT_S2), u, SDE_K(SDE_ROT_S1));
  sp = fma(sp, u, SDE_K(SDE_ROT_S0));
  const double sn = t * sp;                       // sin(phi)
  double cm = fma(kr ? kr[2] : SDE_K(SDE_ROT_C2), u, SDE_K(SDE_ROT_C1));
  cm = cm * u;                                    // cos(phi) - 1
  ca = fma(C, cm, fma(-S, sn, C));   // cos(theta_i + phi) = C + C (cos phi - 1) - S sin phi
  sb = fma(S, cm, fma(C, sn, S));    // sin(theta_i + phi) = S + S (cos phi - 1) + C sin phi
}
SDE_HD void normal_pair_f64(const uint32_t (&r)[4], const double *tab, const double *lk, double &z0, double &z1,
                            const double *kr = nullptr) {
  double w, ca, sb;
  normal_polar_f64(r, tab, lk, w, ca, sb, kr);
  const double g = sqrt_core(w);  // radius, already scaled by sqrt(dt)
  z0 = g * ca;
  z1 = g * sb;
}

// ---------------------------------------------------------------------------------------------------
// "sdeb200 normal stream v2", FP64 (opt-in: sdeb200_opts.rng = 2): 96 Philox bits per pair instead of 128, i.e. three
// Philox4x32-10 blocks per four pairs (-25 % of the generator's integer work, which is what bounds the FP64 kernels:
// an IMAD.WIDE costs 5.5 issue cycles on sm_100a, profiles/r01_mb_cost.txt).  The path's Philox output is read as a
// stream of 32-bit words W = 0, 1, 2, ... (word W = lane W % 4 of block W / 4); pair P takes words (3P, 3P+1, 3P+2):
//   X1 = w[3P+1]:w[3P]  -> u1 exactly as in v1 (64-bit uniform, full relative precision near 0, tail to 9.4 sigma)
//   c  = w[3P+2]        -> angle = c / 2^32 turns = (i + 1/2 + d) / 1024: the top 10 bits pick the cell of the same
//                          1024-entry table, the low 22 bits (re-centred) are the offset; t = 2^32 d is produced by ONE
//                          integer multiply-add, c * 2^10 + 2^31 (mod 2^32), and one 32-bit I2F.
// The radius uses the 6-operation Goldschmidt form (sqrt_from_seed6).  FP32 kernels: v2 == v1 (32 bits per variate already).
// ---------------------------------------------------------------------------------------------------
// sqrt_from_seed without the refinement of h: the last correction d*h only needs h to 2^-22 (d is already 2^-45 g).
SDE_HD double sqrt_from_seed6(double v, double y) {
  double g = v * y;
  const double h = y * 0.5;
  const double e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  const double d = fma(-g, g, v);
  return fma(d, h, g);
}
// ---------------------------------------------------------------------------------------------------
// Natural logarithm and reciprocal for the log-density kernel (K7), built from the pieces of the normal generator:
//   ln x = e ln2 - (T_j + r Q(r))/2 + 2^-51,  x = 2^e m, m in [1,2), r = m/c_j - 1, |r| <= 2^-8, T_j = 2 ln(1/c_j) + 2^-50
// (the 128-entry table and the degree-4 Q of stream v1: |r Q(r) + 2 log1p(r)| <= 7.8e-17).  ABSOLUTE error <= 4e-16 for
// |ln x| <= 1 and <= 1 ulp + 2e-16 beyond; no relative guarantee at x -> 1, which a log-determinant inside
// -(D log 2π + log det Σ + q)/2 does not need.  9 FP64-pipe instructions instead of the ~30 + branches of the library
// routine; everything that is not a positive normal number (0, subnormal, negative, inf, NaN) goes to the library routine.
// 1/x: MUFU.RCP64H seed + two Newton steps (<= 1 ulp for normal x whose reciprocal is normal; library division otherwise).
// ---------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)   // out of line: the library routines would otherwise be inlined once per unrolled tile row
static __device__ __noinline__ double log_slow(double x) { return log(x); }
static __device__ __noinline__ double rcp_slow(double x) { return 1.0 / x; }
#else
static inline double log_slow(double x) { return log(x); }
static inline double rcp_slow(double x) { return 1.0 / x; }
#endif
SDE_HD bool log_fast_ok(double x) { return double2hi(x) - 0x00100000u < 0x7FE00000u; }      // positive normal number
SDE_HD double log_fast_core(double x, const double *ltab /* SDE_LOG_TAB_INIT: {1/c_j, 2 ln(1/c_j) + 2^-50}, j < 128 */) {
  const uint32_t h = double2hi(x);      // branch-free; garbage (but no fault: the index is masked) unless log_fast_ok(x)
  const double m = hilo2double((h & 0x000FFFFFu) | 0x3FF00000u, double2lo(x));
  const int j = (int)((h >> 13) & 0x7Fu);
  const double kd = (double)((int)(h >> 20) - 1023);
  const double inv_c = ltab[2 * j], tl = ltab[2 * j + 1];
  const double rr = fma(m, inv_c, -1.0);
  double q = SDE_K(SDE_LOG_C4);
  q = fma(q, rr, SDE_K(SDE_LOG_C3));
  q = fma(q, rr, SDE_K(SDE_LOG_C2));
  q = fma(q, rr, SDE_K(SDE_LOG_C1));
  q = fma(q, rr, -2.0);
  const double w = fma(rr, q, tl);                                 // -2 ln m + 2^-50
  return fma(kd, 0.5 * SDE_TWO_LN2_V, fma(w, -0.5, 0x1.0p-51));
}
SDE_HD double log_fast(double x, const double *ltab) { return log_fast_ok(x) ? log_fast_core(x, ltab) : log_slow(x); }
SDE_HD bool rcp_fast_ok(double x) { return (double2hi(x) & 0x7FFFFFFFu) - 0x00200000u < 0x7FC00000u; }   // |x| in [2^-1021, 2^1021)
SDE_HD double rcp_fast_core(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  double y = hilo2double(double2hi(1.0 / hilo2double(double2hi(x), 0)), 0);
#endif
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);                                                // 2^-40
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
SDE_HD double rcp_fast(double x) { return rcp_fast_ok(x) ? rcp_fast_core(x) : rcp_slow(x); }

//   lk (v2, fine) = noise_var * {Q2_0, Q2_1, Q2_2, Q2_3, -, 2 ln 2};  tab = [0, 2 LOGN): log pairs, then the rotation pairs
SDE_HD void make_log_consts_v2(double dt, double *lk) {
#if SDE_V2_FINE
  lk[0] = SDE_LOG2_C0_V * dt;
  lk[1] = SDE_LOG2_C1_V * dt;
  lk[2] = SDE_LOG2_C2_V * dt;
  lk[3] = SDE_LOG2_C3_V * dt;
  lk[4] = 0.0;
  lk[5] = SDE_TWO_LN2_V * dt;
#else
  lk[0] = SDE_LOG_C1_V * dt;
  lk[1] = SDE_LOG_C2_V * dt;
  lk[2] = SDE_LOG_C3_V * dt;
  lk[3] = SDE_LOG_C4_V * dt;
  lk[4] = -2.0 * dt;
  lk[5] = SDE_TWO_LN2_V * dt;
#endif
}
// Polar form of the pair: radius^2 `w` (already scaled by the step's variance) and the unit vector (ca, sb); the pair is
// (sqrt(w) ca, sqrt(w) sb).  Models whose noise terms share a square-root factor (Heston: sqrt(V) dW) take this form and
// fold the two square roots into one (Model::step_polar, sde_models_builtin.cuh).
SDE_HD void normal_polar_f64_v2(uint32_t a, uint32_t b, uint32_t c, const double *tab, const double *lk, double &w, double &ca,
                                double &sb, const double *kr = nullptr) {
  const uint64_t X1 = ((uint64_t)b << 32) | a;
  const double d1 = (double)X1;
  const uint32_t h1 = double2hi(d1);
#if defined(__CUDA_ARCH__)
  uint32_t one_hi, mh;
  asm("mov.u32 %0, 0x3FF00000;" : "=r"(one_hi));
  asm("lop3.b32 %0, %1, 0x000FFFFF, %2, 0xEA;" : "=r"(mh) : "r"(h1), "r"(one_hi));
  const double m = hilo2double(mh, double2lo(d1));
#else
  const double m = hilo2double((h1 & 0x000FFFFFu) | 0x3FF00000u, double2lo(d1));
#endif
  const double kd = (double)(int)(1087u - (h1 >> 20));
#if SDE_V2_FINE
  const int j = (int)((h1 >> 11) & 0x1FFu);
  const double inv_c = tab[2 * j], tl = tab[2 * j + 1];
  const double rr = fma(m, inv_c, -1.0);
  double q = kr ? kr[0] : lk[3];
  q = fma(q, rr, lk[2]);
  q = fma(q, rr, lk[1]);
  q = fma(q, rr, lk[0]);
#else
  const int j = (int)((h1 >> 13) & 0x7Fu);
  const double inv_c = tab[2 * j], tl = tab[2 * j + 1];
  const double rr = fma(m, inv_c, -1.0);
  double q = kr ? kr[0] : lk[3];
  q = fma(q, rr, lk[2]);
  q = fma(q, rr, lk[1]);
  q = fma(q, rr, lk[0]);
  q = fma(q, rr, lk[4]);
#endif
  w = fma(rr, q, tl);
  w = fma(kd, lk[5], w);
#if SDE_V2_FINE
  const double *rot = tab + 2 * SDE_V2_LOGN + 2 * (c >> 21);
  const double C = rot[0], S = rot[1];
#if defined(__CUDA_ARCH__) && SDE_V2_IMAD
  int32_t ti;   // ((c & 0x1FFFFF) - 2^20) * 2^11 as ONE integer multiply-add (the compiler splits c * 2048u + 0x80000000u into shift + xor)
  asm("mad.lo.s32 %0, %1, 2048, 0x80000000;" : "=r"(ti) : "r"(c));
  const double t = (double)ti;
#else
  const double t = (double)(int32_t)(c * 2048u + 0x80000000u);   // ((c & 0x1FFFFF) - 2^20) * 2^11, exact
#endif
  const double u = t * t;
  const double sp = fma(kr ? kr[1] : SDE_K(SDE_ROT3_S1), u, SDE_K(SDE_ROT3_S0));
  const double sn = t * sp;
  double cm = fma(kr ? kr[2] : SDE_K(SDE_ROT3_C2), u, SDE_K(SDE_ROT3_C1));
#if SDE_V2_ROT1
  // rotation with the cosine of the residual formed explicitly (1 + cm: one rounding at 2^-53 absolute): two of the
  // three-register FMAs become two-register multiplies, which dispatch faster (profiles/r01_mb_cost.txt)
  const double c1 = fma(cm, u, 1.0);
  ca = fma(C, c1, -(S * sn));
  sb = fma(S, c1, C * sn);
  return;
#endif
#else
  const double *rot = tab + 2 * SDE_V2_LOGN + 2 * (c >> 22);
  const double C = rot[0], S = rot[1];
  const double t = (double)(int32_t)(c * 1024u + 0x80000000u);   // ((c & 0x3FFFFF) - 2^21) * 2^10, exact
  const double u = t * t;
  double sp = fma(kr ? kr[1] : SDE_K(SDE_ROT2_S2), u, SDE_K(SDE_ROT2_S1));
  sp = fma(sp, u, SDE_K(SDE_ROT2_S0));
  const double sn = t * sp;
  double cm = fma(kr ? kr[2] : SDE_K(SDE_ROT2_C2), u, SDE_K(SDE_ROT2_C1));
#endif
  cm = cm * u;
  ca = fma(C, cm, fma(-S, sn, C));
  sb = fma(S, cm, fma(C, sn, S));
}
SDE_HD void normal_pair_f64_v2(uint32_t a, uint32_t b, uint32_t c, const double *tab, const double *lk, double &z0, double &z1,
                               const double *kr = nullptr) {
  double w, ca, sb;
  normal_polar_f64_v2(a, b, c, tab, lk, w, ca, sb, kr);
  const double g = sqrt_from_seed6(w, rsqrt_seed(w));
  z0 = g * ca;
  z1 = g * sb;
}

// FP32 stream: 128 Philox bits -> four N(0, scale2) variates.  u1 = (r + 1/2) 2^-32, angle = r' / 2^32 turns.
// log2, sqrt, sin and cos on the MUFU unit (device, SDE_F32_MUFU); the host twin and the SDE_F32_MUFU = 0 build take
// (cos, sin) from a 256-entry table of cell centres (top 8 bits of r') plus a one-term residual rotation in the 24-bit
// offset.  rotf: {cos, sin} float pairs.
#ifndef SDE_F32_MUFU
#define SDE_F32_MUFU 1 /* device: (cos, sin) of the FP32 stream on the MUFU unit (sin.approx / cos.approx, |x| <= pi: abs error
                          ~2^-21) instead of the 256-entry table + residual rotation: 10 instructions fewer per pair; the
                          full-trajectory FP32 kernels are issue-bound (profiles/README.md) */
#endif
// radius of an FP32 pair from its square (one MUFU.SQRT on the device; -0 passes through)
SDE_HD float radius_f32(float w) {
#if defined(__CUDA_ARCH__)
  float g;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(w));
  return g;
#else
  return sqrtf(w);
#endif
}
// polar form: radius^2 `w` (scaled by scale2) and the unit vector (cs, sn); the pair is radius_f32(w) (cs, sn)
SDE_HD void normal_polar_f32(uint32_t ra, uint32_t rb, float scale2, const float *rotf, float &w, float &cs, float &sn) {
#if defined(__CUDA_ARCH__)
  const float u1 = fmaf(__uint2float_rn(ra), 0x1.0p-32f, 0x1.0p-33f);
#if SDE_F32_MUFU
  // u1 <= 1 (the largest word rounds to exactly 1): log2 <= 0, so w >= 0 without a clamp; -0 passes through sqrt.approx
  float l2;   // u1 >= 2^-33 is a normal number: the flush-to-zero form saves the denormal pre-scaling (3 instructions)
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  w = (-0x1.62e430p+0f /* 2 ln 2 */ * scale2) * l2;
  const float ang = __int2float_rn((int32_t)rb) * 0x1.921fb6p-30f;   // 2 pi rb / 2^32 with rb read as signed: [-pi, pi)
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
  (void)rotf;
  return;
#else
  w = -0x1.62e430p+0f /* 2 ln 2 */ * __log2f(u1);
#endif
#else
  const float u1 = fmaf((float)ra, 0x1.0p-32f, 0x1.0p-33f);
  w = -0x1.62e430p+0f * log2f(u1);
#endif
#if !(defined(__CUDA_ARCH__) && SDE_F32_MUFU)
  w = fmaxf(w * scale2, 0.0f);
  const float *rot = rotf + 2 * (rb >> 24);
  const float C = rot[0], S = rot[1];
  const float t = (float)((int32_t)(rb & 0x00FFFFFFu) - 0x00800000);
  const float u = t * t;
  const float sr = t * fmaf(SDE_ROTF_S1, u, SDE_ROTF_S0);
  const float cm = u * fmaf(SDE_ROTF_C2, u, SDE_ROTF_C1);
  cs = fmaf(C, cm, fmaf(-S, sr, C));
  sn = fmaf(S, cm, fmaf(C, sr, S));
#endif
}
SDE_HD void normal_pair_f32(uint32_t ra, uint32_t rb, float scale2, const float *rotf, float &z0, float &z1) {
  float w, cs, sn;
  normal_polar_f32(ra, rb, scale2, rotf, w, cs, sn);
  const float g = radius_f32(w);
  z0 = g * cs;
  z1 = g * sn;
}

// Normals per Philox block for each precision.
template <typename T> struct NPC;
template <> struct NPC<double> { static constexpr int value = 2; };
template <> struct NPC<float> { static constexpr int value = 4; };

SDE_HD void normals_from_block(const uint32_t (&r)[4], const double *tab, const double *lk, double scale2, double *z,
                               const double *kr = nullptr) {
  (void)scale2;
  normal_pair_f64(r, tab, lk, z[0], z[1], kr);
}
SDE_HD void normals_from_block(const uint32_t (&r)[4], const double *tab, const double *lk, float scale2, float *z,
                               const double *kr = nullptr) {
  (void)lk;
  (void)kr;
  const float *rotf = (const float *)tab;
  normal_pair_f32(r[0], r[1], scale2, rotf, z[0], z[1]);
  normal_pair_f32(r[2], r[3], scale2, rotf, z[2], z[3]);
}

#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int sde_gcd(int a, int b) { return b == 0 ? a : sde_gcd(b, a % b); }
template <int N> struct IntC { static constexpr int value = N; };   // a compile-time integer passed to a generic lambda

#if !defined(__CUDACC__)
// host twin of the log table, for the CPU unit tests of the transform (tests/host_math_harness.cpp)
static const double h_log_tab[256] = SDE_LOG_TAB_INIT;
static const double h_rot_tab[2048] = SDE_ROT_TAB_INIT;
static const float h_rotf_tab[512] = SDE_ROTF_TAB_INIT;
#if SDE_V2_FINE
static const double h_log2_tab[2 * SDE_V2_LOGN] = SDE_LOG2_TAB_INIT;
static const double h_rot3_tab[2 * SDE_V2_ROTN] = SDE_ROT3_TAB_INIT;
#endif
static inline void h_scaled_tables_v2(double dt, double *tab) {  // tab[SDE_TAB2_DOUBLES]
#if SDE_V2_FINE
  for (int i = 0; i < SDE_V2_LOGN; i++) { tab[2 * i] = h_log2_tab[2 * i]; tab[2 * i + 1] = h_log2_tab[2 * i + 1] * dt; }
  for (int i = 0; i < 2 * SDE_V2_ROTN; i++) tab[2 * SDE_V2_LOGN + i] = h_rot3_tab[i];
#else
  for (int i = 0; i < 128; i++) { tab[2 * i] = h_log_tab[2 * i]; tab[2 * i + 1] = h_log_tab[2 * i + 1] * dt; }
  for (int i = 0; i < 2048; i++) tab[256 + i] = h_rot_tab[i];
#endif
}
static inline void h_scaled_tables(double dt, double *tab) {  // tab[SDE_TAB_DOUBLES]
  for (int i = 0; i < 128; i++) { tab[2 * i] = h_log_tab[2 * i]; tab[2 * i + 1] = h_log_tab[2 * i + 1] * dt; }
  for (int i = 0; i < 2048; i++) tab[256 + i] = h_rot_tab[i];
}
#endif

// A batch of U steps' worth of increments for one path: dW[u*M + j] ~ N(0, dt), generated from the path's
// sub-stream blocks [blk0, blk0 + NB).  Flat normal index q = step*M + j follows the reference's draw order
// (path outer, step inner, M draws left to right: src/simulation.jl:4-11,42-46).
// V = normal-stream version (sdeb200_opts.rng): 1 = two FP64 normals per Philox block; 2 = eight per three blocks
// (FP64 only; the FP32 stream is the same in both versions).
template <typename T, int M, int V = 1> struct Batch {
  static constexpr bool V2 = (V == 2) && sizeof(T) == 8;
  static constexpr int NPB = NPC<T>::value;                              // normals per Philox block (v1)
  static constexpr int U = (M == 0) ? 1 : (V2 ? 8 / sde_gcd(M, 8) : NPB / sde_gcd(M, NPB));   // steps per batch
  static constexpr int NB = (M == 0) ? 0 : (V2 ? (U * M * 3) / 8 : (U * M) / NPB);             // Philox blocks per batch
  static constexpr int NZ = (M == 0) ? 1 : U * M;
};

// Does the model functor take the polar form of a pair (Model::POLAR, Model::step_polar)?  Only hand-written FAST functors
// declare it; generated functors and the STRICT build go through step<SCHEME>(…, dw, …).
template <class Model> struct PolarOf {
  template <class U> static constexpr int f(decltype(U::POLAR) *) { return U::POLAR; }
  template <class U> static constexpr int f(...) { return 0; }
  static constexpr int value = f<Model>(nullptr);
};
// The noise of one step as the kernels hand it to the model: M increments, or — for POLAR models (M == 2: one pair per
// step; FAST math, either precision, either normal stream) — the triple (w, ca, sb) with dW = sqrt(w) (ca, sb).  EVERY kernel form takes the same
// branch for a given (model, precision, stream), so terminal values, trajectories and the fine path of a coupled level
// stay bit-identical to one another.
template <class Model, typename T, int V> struct Noise {
  typedef Batch<T, Model::M, V> B;
  static constexpr bool POLAR = PolarOf<Model>::value != 0 && Model::M == 2 && !SDE_STRICT;
  static constexpr bool POLAR2 = POLAR && PolarOf<Model>::value >= 2;   // Model::step_polar2: a coarse step from two triples
  static constexpr int STRIDE = POLAR ? 3 : (Model::M > 0 ? Model::M : 0);
  static constexpr int NV = POLAR ? 3 * B::U : B::NZ;   // values per batch
};
template <class Model, int SCHEME, bool POLAR, int V = 1> struct NoiseStep {
  template <typename T> SDE_HD static void run(const T *c, T t, T dt, const T *nz, T *x) {
    Model::template step<SCHEME>(c, t, dt, nz, x);
  }
  template <typename T> SDE_HD static void run2(const T *, T, T, const T *, const T *, T *) {}   // never taken (POLAR2 false)
  template <typename T> SDE_HD static void increments(const T *nz, T *dw) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < Model::M; j++) dw[j] = nz[j];
  }
};
template <class Model, int SCHEME, int V> struct NoiseStep<Model, SCHEME, true, V> {
  template <typename T> SDE_HD static void run(const T *c, T t, T dt, const T *nz, T *x) {
    Model::step_polar(c, t, dt, nz[0], nz[1], nz[2], x);
  }
  template <typename T> SDE_HD static void run2(const T *c, T t, T dt, const T *na, const T *nb, T *x) {
    Model::step_polar2(c, t, dt, na, nb, x);
  }
  // the pair itself, exactly as normal_pair_f64 / normal_pair_f64_v2 / normal_pair_f32 form it
  SDE_HD static void increments(const double *nz, double *dw) {
    const double g = V == 2 ? sqrt_from_seed6(nz[0], rsqrt_seed(nz[0])) : sqrt_core(nz[0]);
    dw[0] = g * nz[1];
    dw[1] = g * nz[2];
  }
  SDE_HD static void increments(const float *nz, float *dw) {
    const float g = radius_f32(nz[0]);
    dw[0] = g * nz[1];
    dw[1] = g * nz[2];
  }
};

#if defined(__CUDACC__)
// ===================================================================================================
// Device-only part: kernels
// ===================================================================================================
static __device__ const double g_log_tab[256] = SDE_LOG_TAB_INIT;
static __device__ const double g_rot_tab[2048] = SDE_ROT_TAB_INIT;
static __device__ const float g_rotf_tab[512] = SDE_ROTF_TAB_INIT;

#if SDE_V2_FINE
static __device__ const double g_log2_tab[2 * SDE_V2_LOGN] = SDE_LOG2_TAB_INIT;
static __device__ const double g_rot3_tab[2 * SDE_V2_ROTN] = SDE_ROT3_TAB_INIT;
#endif
// tables of the transforms -> shared memory: FP64 18 KB (log + rotation; stream v2: 40 KB), FP32 2 KB (rotation, float pairs)
template <typename T, int V = 1> SDE_D void load_tables(double *s_tab, double dt) {
  if (sizeof(T) == 8 && V == 2 && SDE_V2_FINE) {
#if SDE_V2_FINE
    for (int i = threadIdx.x; i < 2 * SDE_V2_LOGN; i += blockDim.x) s_tab[i] = (i & 1) ? g_log2_tab[i] * dt : g_log2_tab[i];
    for (int i = threadIdx.x; i < 2 * SDE_V2_ROTN; i += blockDim.x) s_tab[2 * SDE_V2_LOGN + i] = g_rot3_tab[i];
#endif
  } else if (sizeof(T) == 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = (i & 1) ? g_log_tab[i] * dt : g_log_tab[i];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_tab[256 + i] = g_rot_tab[i];
  } else {
    float *f = (float *)s_tab;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) f[i] = g_rotf_tab[i];
  }
  __syncthreads();
}
template <typename T, int V = 1> struct TabSize {
  static constexpr int value = sizeof(T) == 8 ? (V == 2 ? SDE_TAB2_DOUBLES : SDE_TAB_DOUBLES) : 256;
};

template <typename T> struct Consts {
  T c[SDE_MAXC];
};

// normals of one batch from its Philox words (V2: w[4*NB] holds the NB blocks back to back)
template <typename T, int M, int V>
SDE_HD void normals_from_words_v2(const uint32_t *w, const double *s_tab, const double *lk, T *dw, const double *kr) {
  typedef Batch<T, M, V> B;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int P = 0; P < B::NZ / 2; P++) {
    double z0, z1;
    normal_pair_f64_v2(w[3 * P], w[3 * P + 1], w[3 * P + 2], s_tab, lk, z0, z1, kr);
    dw[2 * P] = (T)z0;
    dw[2 * P + 1] = (T)z1;
  }
}

// stream v1, FP64: the pair of one Philox block in polar form (M == 2: one block per step)
SDE_HD void polar_from_block(const uint32_t (&r)[4], const double *tab, const double *lk, double scale2, double *nz, const double *kr) {
  (void)scale2;
  normal_polar_f64(r, tab, lk, nz[0], nz[1], nz[2], kr);
}
// FP32: the two pairs of a block, one triple per step
SDE_HD void polar_from_block(const uint32_t (&r)[4], const double *tab, const double *lk, float scale2, float *nz, const double *kr) {
  (void)lk;
  (void)kr;
  const float *rotf = (const float *)tab;
  normal_polar_f32(r[0], r[1], scale2, rotf, nz[0], nz[1], nz[2]);
  normal_polar_f32(r[2], r[3], scale2, rotf, nz[3], nz[4], nz[5]);
}
// the same pairs in polar form, one triple (w, ca, sb) per step (M == 2; see Noise<> below)
template <typename T, int M, int V>
SDE_HD void polar_from_words_v2(const uint32_t *w, const double *s_tab, const double *lk, T *nz, const double *kr) {
  typedef Batch<T, M, V> B;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int P = 0; P < B::NZ / 2; P++) {
    double rw, ca, sb;
    normal_polar_f64_v2(w[3 * P], w[3 * P + 1], w[3 * P + 2], s_tab, lk, rw, ca, sb, kr);
    nz[3 * P] = (T)rw;
    nz[3 * P + 1] = (T)ca;
    nz[3 * P + 2] = (T)sb;
  }
}

template <typename T, int M, int V = 1, bool POL = false>
SDE_D void gen_batch(uint64_t path, uint32_t blk0, uint32_t stream, const PhiloxKeys &ks, const double *s_tab,
                     const double *lk, T scale2, T *dw, const double *kr = nullptr) {
  typedef Batch<T, M, V> B;
  if (M == 0) {
    dw[0] = (T)0;
    return;
  }
  if (B::V2) {
    uint32_t w[4 * (B::NB > 0 ? B::NB : 1)];
#pragma unroll
    for (int b = 0; b < B::NB; b++) {
      uint32_t r[4];
      philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), blk0 + (uint32_t)b, stream, ks, r);
#pragma unroll
      for (int i = 0; i < 4; i++) w[4 * b + i] = r[i];
    }
    if (POL) polar_from_words_v2<T, M, V>(w, s_tab, lk, dw, kr);
    else normals_from_words_v2<T, M, V>(w, s_tab, lk, dw, kr);
  } else {
#pragma unroll
    for (int b = 0; b < B::NB; b++) {
      uint32_t r[4];
      philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), blk0 + (uint32_t)b, stream, ks, r);
      if (POL) polar_from_block(r, s_tab, lk, scale2, dw + b * (3 * B::NPB / 2), kr);
      else normals_from_block(r, s_tab, lk, scale2, dw + b * B::NPB, kr);
    }
  }
}

// Register-resident copies of the three leading coefficients of the FP64 transform (normal_pair_f64's `kr`).  The fake
// per-thread input keeps the compiler from proving the values warp-uniform and moving them back to uniform registers,
// from where every use as a DFMA multiplicand costs two moves (measured: +1.3 % on the headline kernel).
struct PinnedCoef {
  double v[3];
  SDE_D PinnedCoef(const double *lk, bool fp64, int ver = 1) {
    v[0] = lk[3];
#if SDE_V2_FINE
    v[1] = ver == 2 ? SDE_K(SDE_ROT3_S1) : SDE_K(SDE_ROT_S2);
    v[2] = ver == 2 ? SDE_K(SDE_ROT3_C2) : SDE_K(SDE_ROT_C2);
#else
    v[1] = ver == 2 ? SDE_K(SDE_ROT2_S2) : SDE_K(SDE_ROT_S2);
    v[2] = ver == 2 ? SDE_K(SDE_ROT2_C2) : SDE_K(SDE_ROT_C2);
#endif
    if (fp64) asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]) : "r"(threadIdx.x));   // FP32 kernels: dead code
  }
};

// gen_batch for a path whose round-0/1 invariants are precomputed (philox_path_init)
template <typename T, int M, int V = 1, bool POL = false>
SDE_D void gen_batch(const PhiloxPath &pp, uint32_t blk0, const PhiloxKeys &ks, const double *s_tab, const double *lk, T scale2,
                     T *dw, const double *kr = nullptr) {
  typedef Batch<T, M, V> B;
  if (M == 0) {
    dw[0] = (T)0;
    return;
  }
  if (B::V2) {
    uint32_t w[4 * (B::NB > 0 ? B::NB : 1)];
#pragma unroll
    for (int b = 0; b < B::NB; b++) {
      uint32_t r[4];
      philox4x32_10(pp, blk0 + (uint32_t)b, ks, r);
#pragma unroll
      for (int i = 0; i < 4; i++) w[4 * b + i] = r[i];
    }
    if (POL) polar_from_words_v2<T, M, V>(w, s_tab, lk, dw, kr);
    else normals_from_words_v2<T, M, V>(w, s_tab, lk, dw, kr);
  } else {
#pragma unroll
    for (int b = 0; b < B::NB; b++) {
      uint32_t r[4];
      philox4x32_10(pp, blk0 + (uint32_t)b, ks, r);
      if (POL) polar_from_block(r, s_tab, lk, scale2, dw + b * (3 * B::NPB / 2), kr);
      else normals_from_block(r, s_tab, lk, scale2, dw + b * B::NPB, kr);
    }
  }
}

// FP64 kernels take the per-path Philox form and the pinned coefficients, FP32 kernels the plain block function
// (measured 0.45 % faster there: fewer live registers); `pp` is dead code in the FP32 instantiations.
template <typename T, int M, int V = 1, bool POL = false>
SDE_D void gen_batch(const PhiloxPath &pp, uint64_t path, uint32_t blk0, uint32_t stream, const PhiloxKeys &ks, const double *s_tab,
                     const double *lk, T scale2, T *dw, const double *kr) {
  if (sizeof(T) == 8) gen_batch<T, M, V, POL>(pp, blk0, ks, s_tab, lk, scale2, dw, kr);
  else gen_batch<T, M, V, POL>(path, blk0, stream, ks, s_tab, lk, scale2, dw);
}

// Stream v2, step by step: the normals of step u of a batch are produced just before the step needs them — pair P is
// generated when the first step that reads it comes up, Philox block b when the first pair that reads one of its words
// does — so that a thread carrying several paths holds at most three left-over words and one spare normal per path
// between steps instead of a whole batch (8 normals, 12 words).  u is a constant after unrolling; every loop bound and
// condition below folds at compile time.  Same words -> same normals as gen_batch<T, M, 2>.
template <typename T, int M> struct LazyV2 {
  typedef Batch<T, M, 2> B;
  uint32_t w[4 * (B::NB > 0 ? B::NB : 1)];
  T z[B::NZ];
  static SDE_HD constexpr int pairs_before(int u) { return (u * M + 1) / 2; }                 // pairs steps [0, u) read
  static SDE_HD constexpr int blocks_before(int P) { return P == 0 ? 0 : (3 * P - 1) / 4 + 1; }  // blocks pairs [0, P) read
  SDE_D void produce(int u, const PhiloxPath &pp, uint32_t blk0, const PhiloxKeys &ks, const double *s_tab, const double *lk,
                     const double *kr) {
#pragma unroll
    for (int P = 0; P < B::NZ / 2; P++) {
      if (P >= pairs_before(u) && P < pairs_before(u + 1)) {
#pragma unroll
        for (int b = 0; b < B::NB; b++) {
          if (b >= blocks_before(P) && b < blocks_before(P + 1)) {
            uint32_t r[4];
            philox4x32_10(pp, blk0 + (uint32_t)b, ks, r);
#pragma unroll
            for (int i = 0; i < 4; i++) w[4 * b + i] = r[i];
          }
        }
        double z0, z1;
        normal_pair_f64_v2(w[3 * P], w[3 * P + 1], w[3 * P + 2], s_tab, lk, z0, z1, kr);
        z[2 * P] = (T)z0;
        z[2 * P + 1] = (T)z1;
      }
    }
  }
  // M == 2, POLAR models: step u reads pair P = u; the triple (w, ca, sb) goes straight to the model, nothing is kept in z
  SDE_D void produce_polar(int u, const PhiloxPath &pp, uint32_t blk0, const PhiloxKeys &ks, const double *s_tab, const double *lk,
                           const double *kr, T *nz) {
#pragma unroll
    for (int b = 0; b < B::NB; b++) {
      if (b >= blocks_before(u) && b < blocks_before(u + 1)) {
        uint32_t r[4];
        philox4x32_10(pp, blk0 + (uint32_t)b, ks, r);
#pragma unroll
        for (int i = 0; i < 4; i++) w[4 * b + i] = r[i];
      }
    }
    double rw, ca, sb;
    normal_polar_f64_v2(w[3 * u], w[3 * u + 1], w[3 * u + 2], s_tab, lk, rw, ca, sb, kr);
    nz[0] = (T)rw;
    nz[1] = (T)ca;
    nz[2] = (T)sb;
  }
};

template <typename T> SDE_D bool all_finite(const T *x, int D) {
  bool ok = true;
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++)
    if (d < D) ok = ok && isfinite(x[d]);
  return ok;
}

// Trajectory kernels: count the paths whose FINAL state is non-finite (the reference raises DomainError from sqrt inside
// simulate / simulate! as well); outputs are written regardless.
SDE_D void count_nonfinite_n(const sde_args &a, unsigned bad) {
  if (!a.nonfinite) return;
  const unsigned mask = __activemask();
  bad = __reduce_add_sync(mask, bad);
  if ((int)(threadIdx.x & 31) == __ffs((int)mask) - 1 && bad) atomicAdd(a.nonfinite, (unsigned long long)bad);
}
template <typename T> SDE_D void count_nonfinite(const sde_args &a, bool live, const T *x, int D) {
  count_nonfinite_n(a, (live && !all_finite(x, D)) ? 1u : 0u);
}

// payoff features of one terminal state
SDE_D int payoff_features(const sde_args &a, int D, const double *x, double *f) {
  switch (a.payoff_kind) {
    case SDE_PAYOFF_CALL: f[0] = a.payoff_disc * fmax(x[a.payoff_comp] - a.payoff_k, 0.0); return 1;
    case SDE_PAYOFF_PUT: f[0] = a.payoff_disc * fmax(a.payoff_k - x[a.payoff_comp], 0.0); return 1;
    case SDE_PAYOFF_COMPONENT: f[0] = x[a.payoff_comp]; return 1;
    default:
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) f[d] = d < D ? x[d] : 0.0;
      return D;
  }
}

// Warp-shuffle + CTA reduction of per-thread sums; thread 0 writes the CTA's partial (fixed order =>
// deterministic for a given path chunk, independent of how chunks are spread over GPUs).
template <int NQ> SDE_D void block_reduce_store(double (&v)[NQ], int nq, double *s_red, double *dst) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    if (q < nq) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
      if (lane == 0) s_red[wid * NQ + q] = v[q];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < nq; q++) {
      double s = s_red[q];
      for (int w = 1; w < nw; w++) s += s_red[w * NQ + q];
      dst[q] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K1  terminal values: sample(model, scheme, t0, x0, nsteps, npaths)   [simulation.jl:39-48,59-67]
//     + optional fused payoff reduction (north_star (c)).  PPT paths per thread for ILP.
//     chunk = blockDim.x * PPT consecutive paths per CTA; path = chunk_base + p*blockDim.x + tid.
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T, int SCHEME, int PPT, int V = 1>
SDE_D void sample_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M;
  typedef Batch<T, M, V> B;
  typedef Noise<Model, T, V> NS;
  // transform tables: static shared memory, except the 40 KB of stream v2's fine tables, which come from the launch's
  // dynamic shared memory (static + the one-path-per-thread form's staging arrays would pass the 48 KB static limit)
  constexpr bool DYN_TAB = B::V2 && SDE_V2_FINE;
  __shared__ double s_tab_static[DYN_TAB ? 1 : TabSize<T, V>::value];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double *s_tab = DYN_TAB ? (double *)s_dyn : s_tab_static;
  __shared__ double s_red[(SDE_BLOCK * SDE_PPT / PPT / 32) * 2 * SDE_MAXF];
  __shared__ double s_stage[PPT == SDE_PPT ? 1 : SDE_MAXF * SDE_BLOCK * SDE_PPT];   // features of a chunk (PPT < SDE_PPT forms)
  __shared__ unsigned char s_ok[PPT == SDE_PPT ? 1 : SDE_BLOCK * SDE_PPT];
  load_tables<T, V>(s_tab, a.noise_var);
  PhiloxKeys ks;
  philox_expand(a.seed_lo, a.seed_hi, ks);

  Consts<T> cst;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dt = (T)a.dt, t0 = (T)a.t0;
  Model::prepare(prm, dt, cst.c);

  // a CTA walks over chunks (grid-stride): the tables are staged once per CTA, partials stay per chunk
  const int64_t nchunks = (a.npaths + (int64_t)blockDim.x * PPT - 1) / ((int64_t)blockDim.x * PPT);
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
  const int64_t base = chunk * (blockDim.x * PPT);
  T x[PPT][SDE_MAXD];
  bool live[PPT];
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    live[p] = base + p * (int)blockDim.x + threadIdx.x < a.npaths;
#pragma unroll
    for (int d = 0; d < SDE_MAXD; d++) x[p][d] = d < D ? (T)a.x0[d] : (T)0;
  }
  const uint64_t gpath0 = (uint64_t)(a.path0 + base) + threadIdx.x;
  const int64_t nbatch = a.nsteps / B::U;
  const int rem = (int)(a.nsteps - nbatch * B::U);
  PhiloxPath pp[PPT];
#pragma unroll
  for (int p = 0; p < PPT; p++) philox_path_init(gpath0 + (uint64_t)(p * (int)blockDim.x), a.stream, ks, pp[p], sizeof(T) == 8);
  PinnedCoef pc(a.lk, sizeof(T) == 8, V);
  const double *kr = pc.v;

  if (B::V2 && M > 0) {
    // stream v2: step-major order inside a batch, normals produced lazily (LazyV2) — the same arithmetic per path
    LazyV2<T, (M > 0 ? M : 1)> lz[PPT];
    auto one_step = [&](int64_t it, int u) {   // u is a constant after unrolling
      const T t = t0 + (T)(it * B::U + u) * dt;
#pragma unroll
      for (int p = 0; p < PPT; p++) {
        if (NS::POLAR) {
          T nz[3];
          lz[p].produce_polar(u, pp[p], (uint32_t)(it * B::NB), ks, s_tab, a.lk, kr, nz);
          NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, nz, x[p]);
        } else {
          lz[p].produce(u, pp[p], (uint32_t)(it * B::NB), ks, s_tab, a.lk, kr);
          NoiseStep<Model, SCHEME, false>::run(cst.c, t, dt, lz[p].z + u * (M > 0 ? M : 0), x[p]);
        }
      }
    };
    constexpr int UNR = SDE_V2_UNROLL;   // (a macro is not expanded inside the pragma)
#pragma unroll UNR
    for (int64_t it = 0; it < nbatch; it++) {   // whole batches: no per-step conditions
#pragma unroll
      for (int u = 0; u < B::U; u++) one_step(it, u);
    }
    if (rem > 0) {
#pragma unroll
      for (int u = 0; u < B::U; u++)
        if (u < rem) one_step(nbatch, u);
    }
  } else {
  for (int64_t it = 0; it < nbatch; it++) {
#pragma unroll
    for (int p = 0; p < PPT; p++) {
      T dw[NS::NV];
      gen_batch<T, M, V, NS::POLAR>(pp[p], gpath0 + (uint64_t)(p * (int)blockDim.x), (uint32_t)(it * B::NB), a.stream, ks, s_tab,
                                    a.lk, (T)a.noise_var, dw, kr);
#pragma unroll
      for (int u = 0; u < B::U; u++) {
        const T t = t0 + (T)(it * B::U + u) * dt;
        NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, dw + u * NS::STRIDE, x[p]);
      }
    }
  }
  if (rem > 0) {
#pragma unroll
    for (int p = 0; p < PPT; p++) {
      T dw[NS::NV];
      gen_batch<T, M, V, NS::POLAR>(pp[p], gpath0 + (uint64_t)(p * (int)blockDim.x), (uint32_t)(nbatch * B::NB), a.stream, ks, s_tab,
                                    a.lk, (T)a.noise_var, dw, kr);
#pragma unroll
      for (int u = 0; u < B::U; u++) {
        if (u < rem) {
          const T t = t0 + (T)(nbatch * B::U + u) * dt;
          NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, dw + u * NS::STRIDE, x[p]);
        }
      }
    }
  }
  }  // stream v1 order

  // terminal states, reference layout [path][D]  (Vector{SVector{D}}: simulation.jl:59-60)
  if (a.out) {
    T *out = (T *)a.out;
#pragma unroll
    for (int p = 0; p < PPT; p++) {
      if (live[p]) {
        const int64_t k = base + p * (int)blockDim.x + threadIdx.x;
#pragma unroll
        for (int d = 0; d < D; d++) out[k * D + d] = x[p][d];
      }
    }
  }
  if (a.partials) {
    // per-chunk partial sums.  The summation tree is a function of the chunk alone (512 consecutive paths: path
    // i = p*128 + t is added p = 0..3 in order for each t < 128, then lanes, then the four warps), NOT of how many
    // paths a thread carries: the PPT = 1 form (512 threads, small launches) stages its features in shared memory
    // and lets threads t < 128 run the same loop, so both forms write bit-identical partials.
    double acc[2 * SDE_MAXF];
#pragma unroll
    for (int q = 0; q < 2 * SDE_MAXF; q++) acc[q] = 0.0;
    unsigned bad = 0;
    int nf = 1;
    if (PPT == SDE_PPT) {
#pragma unroll
      for (int p = 0; p < PPT; p++) {
        double xd[SDE_MAXD], f[SDE_MAXF];
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xd[d] = (double)x[p][d];
        nf = payoff_features(a, D, xd, f);
        if (live[p]) {
          if (all_finite(xd, D)) {
#pragma unroll
            for (int q = 0; q < SDE_MAXF; q++)
              if (q < nf) { acc[2 * q] += f[q]; acc[2 * q + 1] = fma(f[q], f[q], acc[2 * q + 1]); }
          } else {
            bad++;
          }
        }
      }
    } else {
      constexpr int NT = SDE_BLOCK * SDE_PPT / PPT;   // threads of this form
#pragma unroll
      for (int p = 0; p < PPT; p++) {
        double xd[SDE_MAXD], f[SDE_MAXF];
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xd[d] = (double)x[p][d];
        nf = payoff_features(a, D, xd, f);
        const bool ok = live[p] && all_finite(xd, D);
        if (live[p] && !ok) bad++;
        const int i = p * NT + threadIdx.x;           // path index inside the chunk
        s_ok[i] = ok ? 1 : 0;
#pragma unroll
        for (int q = 0; q < SDE_MAXF; q++)
          if (q < nf) s_stage[q * (SDE_BLOCK * SDE_PPT) + i] = f[q];
      }
      __syncthreads();
      if (threadIdx.x < SDE_BLOCK) {
#pragma unroll
        for (int p = 0; p < SDE_PPT; p++) {
          const int i = p * SDE_BLOCK + threadIdx.x;
          if (s_ok[i]) {
#pragma unroll
            for (int q = 0; q < SDE_MAXF; q++)
              if (q < nf) {
                const double fq = s_stage[q * (SDE_BLOCK * SDE_PPT) + i];
                acc[2 * q] += fq;
                acc[2 * q + 1] = fma(fq, fq, acc[2 * q + 1]);
              }
          }
        }
      }
    }
    block_reduce_store<2 * SDE_MAXF>(acc, 2 * nf, s_red, a.partials + chunk * (2 * SDE_MAXF));
    __syncthreads();  // s_red / s_stage are reused by the next chunk
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(a.nonfinite, (unsigned long long)bad);
  } else if (a.nonfinite) {
    unsigned bad = 0;
#pragma unroll
    for (int p = 0; p < PPT; p++) bad += (live[p] && !all_finite(x[p], D)) ? 1u : 0u;
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(a.nonfinite, (unsigned long long)bad);
  }
  }  // chunk loop
}

// ---------------------------------------------------------------------------------------------------
// K4  coupled fine/coarse level pair: _multilevel_sample   [multilevel.jl:65-89]
//     a.nsteps = number of FINE steps (ncoarse*substeps), a.dt = fine Δt, a.dt_coarse = coarse Δt.
//     ΔW accumulates the fine increments sequentially from zero exactly as multilevel.jl:79-83 does.
//     out = fine terminals, out2 = coarse terminals; partial features (Pf - Pc, Pf, Pc).
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T, int SCHEME, int V = 1>
SDE_D void mlmc_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M;
  typedef Batch<T, M, V> B;
  typedef Noise<Model, T, V> NS;
  __shared__ double s_tab[TabSize<T, V>::value];
  __shared__ double s_red[(SDE_BLOCK / 32) * 2 * SDE_MAXF];
  load_tables<T, V>(s_tab, a.noise_var);
  PhiloxKeys ks;
  philox_expand(a.seed_lo, a.seed_hi, ks);

  Consts<T> cf, cc;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dtf = (T)a.dt, dtc = (T)a.dt_coarse, t0 = (T)a.t0;
  Model::prepare(prm, dtf, cf.c);
  Model::prepare(prm, dtc, cc.c);
  PinnedCoef pc(a.lk, sizeof(T) == 8, V);

  const int64_t nchunks = (a.npaths + blockDim.x - 1) / blockDim.x;
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
  const int64_t k = chunk * blockDim.x + threadIdx.x;
  const bool live = k < a.npaths;
  const uint64_t gpath = (uint64_t)(a.path0 + k);
  PhiloxPath pp;
  philox_path_init(gpath, a.stream, ks, pp, sizeof(T) == 8);
  T xf[SDE_MAXD], xc[SDE_MAXD], DW[M > 0 ? M : 1];
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++) xf[d] = xc[d] = d < D ? (T)a.x0[d] : (T)0;
#pragma unroll
  for (int j = 0; j < (M > 0 ? M : 1); j++) DW[j] = (T)0;

  const int64_t nbatch = (a.nsteps + B::U - 1) / B::U;
  const int substeps = (int)a.substeps;
  int sub = 0;          // fine steps taken inside the current coarse step
  int64_t ncoarse = 0;  // coarse steps completed
  int64_t it = 0;
  if (substeps == 2) {
    // the usual refinement factor, unrolled over groups of G = lcm(U, 2) fine steps: the position inside the coarse
    // step is a compile-time constant, no counters, no branch.  Same operations in the same order as the general loop
    // below (which finishes what is left when nsteps is not a multiple of G).
    constexpr int G = (B::U % 2 == 0) ? B::U : 2 * B::U, NBG = G / B::U;
    const int64_t ngroups = a.nsteps / G;
    for (int64_t gi = 0; gi < ngroups; gi++) {
      T dw[NBG][NS::NV];
#pragma unroll
      for (int b = 0; b < NBG; b++)
        gen_batch<T, M, V, NS::POLAR>(pp, gpath, (uint32_t)((gi * NBG + b) * B::NB), a.stream, ks, s_tab, a.lk, (T)a.noise_var, dw[b],
                                      pc.v);
#pragma unroll
      for (int j = 0; j < G; j++) {
        const T tc = t0 + (T)(gi * (G / 2) + j / 2) * dtc;
        const T tf = tc + (T)(j & 1) * dtf;
        const T *nzu = dw[j / B::U] + (j % B::U) * NS::STRIDE;
        if (NS::POLAR2) {
          // the coarse step straight from the two polar triples of its fine steps (Model::step_polar2): no increments,
          // no sum ΔW, one square root per fine increment instead of radius + radius + model root
          NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cf.c, tf, dtf, nzu, xf);
          if (j & 1) NoiseStep<Model, SCHEME, NS::POLAR2, V>::run2(cc.c, tc, dtc, dw[(j - 1) / B::U] + ((j - 1) % B::U) * NS::STRIDE, nzu, xc);
        } else {
          T dwu[M > 0 ? M : 1];
          NoiseStep<Model, SCHEME, NS::POLAR, V>::increments(nzu, dwu);
#pragma unroll
          for (int q = 0; q < M; q++) DW[q] = ((j & 1) ? DW[q] : (T)0) + dwu[q];
          NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cf.c, tf, dtf, nzu, xf);
          if (j & 1) Model::template step<SCHEME>(cc.c, tc, dtc, DW, xc);
        }
      }
    }
    it = ngroups * NBG;
    ncoarse = ngroups * (G / 2);
#pragma unroll
    for (int q = 0; q < (M > 0 ? M : 1); q++) DW[q] = (T)0;
  }
  for (; it < nbatch; it++) {
    T dw[NS::NV];
    gen_batch<T, M, V, NS::POLAR>(pp, gpath, (uint32_t)(it * B::NB), a.stream, ks, s_tab, a.lk, (T)a.noise_var, dw, pc.v);
#pragma unroll
    for (int u = 0; u < B::U; u++) {
      const int64_t fstep = it * B::U + u;
      if (fstep < a.nsteps) {
        const T tc = t0 + (T)ncoarse * dtc;
        const T tf = tc + (T)sub * dtf;
        const T *nzu = dw + u * NS::STRIDE;
        T dwu[M > 0 ? M : 1];
        NoiseStep<Model, SCHEME, NS::POLAR, V>::increments(nzu, dwu);
#pragma unroll
        for (int j = 0; j < M; j++) DW[j] = DW[j] + dwu[j];
        NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cf.c, tf, dtf, nzu, xf);
        if (++sub == substeps) {
          Model::template step<SCHEME>(cc.c, tc, dtc, DW, xc);
#pragma unroll
          for (int j = 0; j < (M > 0 ? M : 1); j++) DW[j] = (T)0;
          sub = 0;
          ncoarse++;
        }
      }
    }
  }
  if (live) {
    if (a.out) {
#pragma unroll
      for (int d = 0; d < D; d++) ((T *)a.out)[k * D + d] = xf[d];
    }
    if (a.out2) {
#pragma unroll
      for (int d = 0; d < D; d++) ((T *)a.out2)[k * D + d] = xc[d];
    }
  }
  if (a.partials) {
    double acc[2 * SDE_MAXF];
#pragma unroll
    for (int q = 0; q < 2 * SDE_MAXF; q++) acc[q] = 0.0;
    double xd[SDE_MAXD], ff[SDE_MAXF], fc[SDE_MAXF];
#pragma unroll
    for (int d = 0; d < SDE_MAXD; d++) xd[d] = (double)xf[d];
    payoff_features(a, D, xd, ff);
    bool ok = all_finite(xd, D);
#pragma unroll
    for (int d = 0; d < SDE_MAXD; d++) xd[d] = (double)xc[d];
    payoff_features(a, D, xd, fc);
    ok = ok && all_finite(xd, D);
    unsigned bad = 0;
    if (live) {
      if (ok) {
        const double y = ff[0] - fc[0];
        acc[0] = y; acc[1] = y * y;
        acc[2] = ff[0]; acc[3] = ff[0] * ff[0];
        acc[4] = fc[0]; acc[5] = fc[0] * fc[0];
      } else {
        bad = 1;
      }
    }
    block_reduce_store<2 * SDE_MAXF>(acc, 6, s_red, a.partials + chunk * (2 * SDE_MAXF));
    __syncthreads();  // s_red is reused by the next chunk
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(a.nonfinite, (unsigned long long)bad);
  } else if (a.nonfinite) {
    unsigned bad = (live && !(all_finite(xf, D) && all_finite(xc, D))) ? 1u : 0u;
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(a.nonfinite, (unsigned long long)bad);
  }
  }  // chunk loop
}

// ---------------------------------------------------------------------------------------------------
// K2a full trajectories, device layout [time][D][path]:  simulate / simulate!   [simulation.jl:71-76,104-117]
//     VEC = 4 consecutive paths per thread -> 16-byte streaming stores (one in FP32, two in FP64) per (time, d).
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC> struct VecStore;
template <> struct VecStore<double, 2> {
  static SDE_D void st(double *p, const double *v) { __stcs((double2 *)p, make_double2(v[0], v[1])); }
};
template <> struct VecStore<float, 4> {
  static SDE_D void st(float *p, const float *v) { __stcs((float4 *)p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct VecStore<double, 4> {  // two 16-byte stores
  static SDE_D void st(double *p, const double *v) {
    __stcs((double2 *)p, make_double2(v[0], v[1]));
    __stcs((double2 *)p + 1, make_double2(v[2], v[3]));
  }
};
template <> struct VecStore<float, 8> {
  static SDE_D void st(float *p, const float *v) {
    __stcs((float4 *)p, make_float4(v[0], v[1], v[2], v[3]));
    __stcs((float4 *)p + 1, make_float4(v[4], v[5], v[6], v[7]));
  }
};

template <class Model, typename T, int SCHEME, int V = 1>
SDE_D void simulate_soa_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M, VEC = SDE_SOA_VEC;
  typedef Batch<T, M, V> B;
  typedef Noise<Model, T, V> NS;
  __shared__ double s_tab[TabSize<T, V>::value];
  load_tables<T, V>(s_tab, a.noise_var);
  PhiloxKeys ks;
  philox_expand(a.seed_lo, a.seed_hi, ks);
  Consts<T> cst;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dt = (T)a.dt, t0 = (T)a.t0;
  Model::prepare(prm, dt, cst.c);

  const int64_t k0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;  // first local path of this thread
  if (k0 >= a.npaths) return;
  const int nlive = (int)((a.npaths - k0) < VEC ? (a.npaths - k0) : VEC);
  T *out = (T *)a.out;
  const int64_t ld = a.ld_paths, col = a.path_off + k0;
  constexpr int ALN = 16 / (int)sizeof(T);  // 16-byte store alignment in elements
  const bool vec_ok = nlive == VEC && ((ld % ALN) == 0) && ((col % ALN) == 0) &&
                      ((((uint64_t)out) & 15u) == 0);
  T x[VEC][SDE_MAXD];
#pragma unroll
  for (int v = 0; v < VEC; v++)
#pragma unroll
    for (int d = 0; d < SDE_MAXD; d++) x[v][d] = d < D ? (T)a.x0[d] : (T)0;

  // rows are written in order, so the row pointer is bumped (one 64-bit add per row) instead of being rebuilt from the row
  // index, and the vectorised / scalar choice is made once, outside the time loop (two instantiations of the loop body):
  // measured 11 of 94 instructions per warp and row in the FP32 form were store addressing and the branch around it
  T *prow = out + col;
  const int64_t rstride = (int64_t)D * ld;
  auto store_next = [&](auto vec) {
    constexpr bool VECST = decltype(vec)::value != 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
      T *p = prow + d * ld;
      T vals[VEC];
#pragma unroll
      for (int v = 0; v < VEC; v++) vals[v] = x[v][d];
      if (VECST) {
        VecStore<T, VEC>::st(p, vals);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; v++)
          if (v < nlive) __stcs(p + v, vals[v]);
      }
    }
    prow += rstride;
  };
  if (vec_ok) store_next(IntC<1>());
  else store_next(IntC<0>());
  const uint64_t gpath = (uint64_t)(a.path0 + k0);
  PhiloxPath pp[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) philox_path_init(gpath + (uint64_t)v, a.stream, ks, pp[v], sizeof(T) == 8 && SDE_SOA_PP);
  PinnedCoef pc(a.lk, sizeof(T) == 8, V);
  const int64_t nbatch = (a.nsteps + B::U - 1) / B::U, nfull = a.nsteps / B::U;
  const T nv = (T)a.noise_var;
  auto batch = [&](int64_t it, bool full, auto vec) {
    T dw[VEC][NS::NV];
#pragma unroll
    for (int v = 0; v < VEC; v++)
#if SDE_SOA_PP
      gen_batch<T, M, V, NS::POLAR>(pp[v], gpath + (uint64_t)v, (uint32_t)(it * B::NB), a.stream, ks, s_tab, a.lk, nv, dw[v], pc.v);
#else
      gen_batch<T, M, V, NS::POLAR>(gpath + (uint64_t)v, (uint32_t)(it * B::NB), a.stream, ks, s_tab, a.lk, nv, dw[v],
                      sizeof(T) == 8 ? pc.v : nullptr);
#endif
#pragma unroll
    for (int u = 0; u < B::U; u++) {
      const int64_t i = it * B::U + u;
      if (full || i < a.nsteps) {
        const T t = t0 + (T)i * dt;
#pragma unroll
        for (int v = 0; v < VEC; v++) NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, dw[v] + u * NS::STRIDE, x[v]);
        store_next(vec);
      }
    }
  };
  if (vec_ok) {
    for (int64_t it = 0; it < nfull; it++) batch(it, true, IntC<1>());      // whole batches: no per-row bound checks
    for (int64_t it = nfull; it < nbatch; it++) batch(it, false, IntC<1>());
  } else {
    for (int64_t it = 0; it < nfull; it++) batch(it, true, IntC<0>());
    for (int64_t it = nfull; it < nbatch; it++) batch(it, false, IntC<0>());
  }
  unsigned bad = 0;
#pragma unroll
  for (int v = 0; v < VEC; v++) bad += (v < nlive && !all_finite(x[v], D)) ? 1u : 0u;
  count_nonfinite_n(a, bad);
}

// Cooperative copy between a warp's shared-memory tile [32 paths][LD] and global memory where path r's row
// segment starts at g + r*gstride: element e = k*32 + lane of the flattened (32 x ROW) tile, fully unrolled, so a
// warp instruction moves whole 128-byte row segments and all of a lane's loads are in flight together.
template <typename T, int ROW, int LD, int NROWS = 32>
SDE_D void tile_load(T *tile, const T *g, int64_t gstride, int nlive, int ncols, int lane) {
  constexpr int NK = ROW * (NROWS / 32);
  T v[NK];
  if ((32 % ROW) == 0 && nlive == NROWS && ncols == ROW) {  // full tile: no predicates, pointer bumps only
    constexpr int RPI = (32 % ROW) == 0 ? 32 / ROW : 1;     // rows covered by one warp instruction
    const int rr = lane / ROW, cc = lane - rr * ROW;
    const T *gp = g + rr * gstride + cc;
#pragma unroll
    for (int k = 0; k < NK; k++) { v[k] = __ldcs(gp); gp += RPI * gstride; }
    T *tp = tile + rr * LD + cc;
#pragma unroll
    for (int k = 0; k < NK; k++) tp[k * RPI * LD] = v[k];
    return;
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const int e = k * 32 + lane, r = e / ROW, c = e - r * ROW;
    v[k] = (r < nlive && c < ncols) ? __ldcs(g + r * gstride + c) : (T)0;
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const int e = k * 32 + lane, r = e / ROW, c = e - r * ROW;
    tile[r * LD + c] = v[k];
  }
}
template <typename T, int ROW, int LD, int NROWS = 32>
SDE_D void tile_store(const T *tile, T *g, int64_t gstride, int nlive, int ncols, int lane) {
  constexpr int NK = ROW * (NROWS / 32);
  if ((32 % ROW) == 0 && nlive == NROWS && ncols == ROW) {
    constexpr int RPI = (32 % ROW) == 0 ? 32 / ROW : 1;
    const int rr = lane / ROW, cc = lane - rr * ROW;
    const T *tp = tile + rr * LD + cc;
    T *gp = g + rr * gstride + cc;
#pragma unroll
    for (int k = 0; k < NK; k++) { __stcs(gp, tp[k * RPI * LD]); gp += RPI * gstride; }
    return;
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const int e = k * 32 + lane, r = e / ROW, c = e - r * ROW;
    if (r < nlive && c < ncols) __stcs(g + r * gstride + c, tile[r * LD + c]);
  }
}

// ---------------------------------------------------------------------------------------------------
// K2b full trajectories written directly in the reference's memory order [path][time][D]:
//     each warp stages a (32 paths) x (TS time rows) tile in shared memory and writes whole path rows,
//     so global stores stay coalesced (32 consecutive elements per instruction) without a transpose pass.
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T, int SCHEME, int V = 1>
SDE_D void simulate_ref_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M;
  constexpr int PW = SDE_SIMREF_PW;           // paths per lane: a warp owns 32 * PW consecutive paths
  constexpr int TS = (sizeof(T) == 8 ? 16 : 32) / D;  // time rows per tile: 128-byte row segments per path
  constexpr int ROW = TS * D;                 // elements per path per tile
  constexpr int LDS = ROW | 1;                // odd stride: conflict-free column writes
  typedef Batch<T, M, V> B;
  typedef Noise<Model, T, V> NS;
  __shared__ double s_tab[TabSize<T, V>::value];
  extern __shared__ __align__(16) unsigned char s_dyn[];   // [warps][32 * PW rows][LDS]: sde_simref_smem_bytes()
  load_tables<T, V>(s_tab, a.noise_var);
  PhiloxKeys ks;
  philox_expand(a.seed_lo, a.seed_hi, ks);
  Consts<T> cst;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dt = (T)a.dt, t0 = (T)a.t0;
  Model::prepare(prm, dt, cst.c);

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wbase = ((int64_t)blockIdx.x * blockDim.x + (wid << 5)) * PW;  // first path of this warp
  if (wbase >= a.npaths) return;
  const int nrow_live = (int)((a.npaths - wbase) < 32 * PW ? (a.npaths - wbase) : 32 * PW);
  T *tile = (T *)s_dyn + wid * (32 * PW * LDS);
  T *out = (T *)a.out;
  const int64_t nrows = a.nsteps + 1;
  T x[PW][SDE_MAXD];
#pragma unroll
  for (int p = 0; p < PW; p++)
#pragma unroll
    for (int d = 0; d < SDE_MAXD; d++) x[p][d] = d < D ? (T)a.x0[d] : (T)0;

  int fill = 0;          // time rows currently staged
  int64_t row0 = 0;      // first staged time row
  auto stage = [&]() {
#pragma unroll
    for (int p = 0; p < PW; p++)
#pragma unroll
      for (int d = 0; d < D; d++) tile[(lane + 32 * p) * LDS + fill * D + d] = x[p][d];
    fill++;
  };
  auto flush = [&]() {
    __syncwarp();
    tile_store<T, ROW, LDS, 32 * PW>(tile, out + (wbase * nrows + row0) * D, nrows * D, nrow_live, fill * D, lane);
    __syncwarp();
    row0 += fill;
    fill = 0;
  };
  stage();
  const uint64_t gpath = (uint64_t)(a.path0 + wbase + lane);
  PhiloxPath pp[PW];
#pragma unroll
  for (int p = 0; p < PW; p++) philox_path_init(gpath + (uint64_t)(32 * p), a.stream, ks, pp[p], sizeof(T) == 8);
  PinnedCoef pc(a.lk, sizeof(T) == 8, V);
  const int64_t nbatch = (a.nsteps + B::U - 1) / B::U;
  for (int64_t it = 0; it < nbatch; it++) {
    T dw[PW][NS::NV];
#pragma unroll
    for (int p = 0; p < PW; p++)
      gen_batch<T, M, V, NS::POLAR>(pp[p], gpath + (uint64_t)(32 * p), (uint32_t)(it * B::NB), a.stream, ks, s_tab, a.lk, (T)a.noise_var, dw[p],
                      pc.v);
#pragma unroll
    for (int u = 0; u < B::U; u++) {
      const int64_t i = it * B::U + u;
      if (i < a.nsteps) {
        const T t = t0 + (T)i * dt;
#pragma unroll
        for (int p = 0; p < PW; p++) NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, dw[p] + u * NS::STRIDE, x[p]);
        if (fill == TS) flush();
        stage();
      }
    }
  }
  flush();
  unsigned bad = 0;
#pragma unroll
  for (int p = 0; p < PW; p++) bad += (lane + 32 * p < nrow_live && !all_finite(x[p], D)) ? 1u : 0u;
  count_nonfinite_n(a, bad);
}

// ---------------------------------------------------------------------------------------------------
// K3  Wiener-driven trajectories, in place allowed: simulate!(x, model, scheme, t0, x0, w)  [simulation.jl:119-134]
//     w (a.out2) and x (a.out) both in reference order [path][row][.]; increments = w[i] - w[i-1].
//     The only HBM-bound kernel of the path (one load + one store per element): each warp owns 32 paths and
//     moves (32 paths) x (TS rows) tiles through shared memory, so that global loads and stores are whole
//     row segments of one path (coalesced) while every lane walks its own path inside the tile
//     (odd row stride: conflict-free).  A tile is loaded completely before any of it is stored => x === w is safe.
//     Parity entry point: with SDE_STRICT=1 the arithmetic is the reference's, operation for operation.
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T, int SCHEME>
SDE_D void simulate_wiener_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M, MW = (M > 0 ? M : 1);
  constexpr int DM = D > MW ? D : MW;
  constexpr int TS = (SDE_SIMW_ROWBYTES / (int)sizeof(T)) / DM;  // rows per tile
  constexpr int WROW = TS * MW, XROW = TS * D;
  constexpr int WLD = WROW | 1, XLD = XROW | 1;
  extern __shared__ __align__(16) unsigned char s_dyn[];   // [warps][32][WLD] then [warps][32][XLD]: sde_simw_smem_bytes()
  T (*s_w)[32 * WLD] = (T (*)[32 * WLD])s_dyn;
  T (*s_x)[32 * XLD] = (T (*)[32 * XLD])((T *)s_dyn + (SDE_BLOCK / 32) * 32 * WLD);
  Consts<T> cst;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dt = (T)a.dt, t0 = (T)a.t0;
  Model::prepare(prm, dt, cst.c);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wbase = (int64_t)blockIdx.x * blockDim.x + (wid << 5);
  if (wbase >= a.npaths) return;
  const int nlive = (int)((a.npaths - wbase) < 32 ? (a.npaths - wbase) : 32);
  const int64_t nrows = a.nsteps + 1;
  const T *w = (const T *)a.out2;
  T *xo = (T *)a.out;
  T *tw = s_w[wid], *tx = s_x[wid];
  T x[SDE_MAXD], wprev[MW], dw[MW];
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++) x[d] = d < D ? (T)a.x0[d] : (T)0;
#pragma unroll
  for (int j = 0; j < MW; j++) wprev[j] = (T)0;

  for (int64_t row0 = 0; row0 < nrows; row0 += TS) {
    const int nr = (int)((nrows - row0) < TS ? (nrows - row0) : TS);
    // cooperative, coalesced load of the w tile
    tile_load<T, WROW, WLD>(tw, w + (wbase * nrows + row0) * MW, nrows * MW, nlive, nr * MW, lane);
    __syncwarp();
    // every lane integrates its own path over the tile
    for (int i = 0; i < nr; i++) {
      const int64_t row = row0 + i;
      if (row == 0) {
#pragma unroll
        for (int j = 0; j < MW; j++) wprev[j] = tw[lane * WLD + j];
      } else {
        const T t = t0 + (T)(row - 1) * dt;
#pragma unroll
        for (int j = 0; j < MW; j++) {
          const T wn = tw[lane * WLD + i * MW + j];
          dw[j] = wn - wprev[j];
          wprev[j] = wn;
        }
        Model::template step<SCHEME>(cst.c, t, dt, dw, x);
      }
#pragma unroll
      for (int d = 0; d < D; d++) tx[lane * XLD + i * D + d] = x[d];
    }
    __syncwarp();
    tile_store<T, XROW, XLD>(tx, xo + (wbase * nrows + row0) * D, nrows * D, nlive, nr * D, lane);
    __syncwarp();
  }
  count_nonfinite(a, lane < nlive, x, D);
}

// ---------------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (inline PTX, sm_90+): tensor tiles global <-> shared, completion on an mbarrier
// (loads) or a bulk async-group (stores).  SASS: UTMALDG / UTMASTG / SYNCS.
// ---------------------------------------------------------------------------------------------------
SDE_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SDE_D void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
SDE_D void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
SDE_D void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SDE_D void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
}
SDE_D void tma_load_2d(uint32_t dst, const sde_tmap *tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tm), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
SDE_D void tma_store_2d(const sde_tmap *tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
SDE_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> SDE_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
SDE_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Cooperative store of a warp's ragged tail tile [32 paths][ROW] (runs once per warp).  Row r is written only from
// column cs_{r % g} on: the columns before belong to time rows that were already written by TMA stores.
template <typename T, int ROW, int LD>
SDE_D void tail_tile_store(const T *tile, T *g, int64_t gstride, int nlive, int ncols, uint32_t hpack, int mult, int gm1, int lane) {
#pragma unroll 4
  for (int k = 0; k < ROW; k++) {
    const int e = k * 32 + lane, r = e / ROW, c = e - r * ROW, j = r & gm1;
    const int cs = (int)((hpack >> (4 * j)) & 15u) * mult;   // head rows h_j, four bits each
    if (r < nlive && c < ncols && c >= cs) __stcs(g + r * gstride + c, tile[r * LD + c]);
  }
}

// ---------------------------------------------------------------------------------------------------
// K3t Wiener-driven trajectories on TMA tensor tiles (the HBM-bound kernel of the path; same semantics as K3):
//     a warp owns 32 paths; its lane 0 moves tiles of w global -> shared through a ring of SDE_SIMWT_STAGES stages
//     (one mbarrier per stage, SDE_SIMWT_STAGES - 2 loads in flight), every lane integrates its own path IN PLACE in
//     the swizzled tile (16-byte shared-memory accesses, conflict-free), lane 0 stores the tile to x.  A tile is
//     complete in shared memory before any of it is stored => x === w is safe.  Rows before the first aligned row
//     of a path (h_j <= 3) use plain accesses; the rows after the last full tile go through one cooperative
//     shared-memory tile (loads via registers, tail_tile_store).  Needs D == M (eltype(x) == eltype(w)) and D*sizeof(T) | 128;
//     the host routes everything else, unaligned buffers, short paths and the npaths % g leftover to K3.
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T, int SCHEME>
SDE_D void simulate_wiener_tma_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M, MW = (M > 0 ? M : 1);
  constexpr int ROWB = D * (int)sizeof(T);
  constexpr bool OK = (D == MW) && (128 % ROWB == 0);
  if (!OK) return;
  constexpr int NST = SDE_SIMWT_STAGES;
  constexpr int TS = 128 / ROWB;                          // rows per tile
  constexpr int GB = ROWB > 16 ? ROWB : 16;               // bytes per register group
  constexpr int G = GB / ROWB > 0 ? GB / ROWB : 1, GQ = GB / 16, NG = 128 / GB;
  constexpr int TILE = 32 * 128;
  constexpr int TALIGN = SDE_SIMWT_ALIGN;                 // alignment of the tiles' row segments (16: what TMA needs)
  constexpr int HMAX = TALIGN > ROWB ? TALIGN / ROWB - 1 : 0;   // bound on the head rows h_j
  constexpr int TROW = (TS + HMAX) * D, TLD = TROW | 1;   // the cooperative tail tile [32][TLD] (fits in two stages)
  static_assert(!OK || 32 * TLD * (int)sizeof(T) <= 2 * TILE, "tail tile does not fit");
  extern __shared__ __align__(16) unsigned char s_dyn[];
  unsigned char *smem = s_dyn + ((1024u - (smem_u32(s_dyn) & 1023u)) & 1023u);  // the swizzle is a function of the address
  const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for ptxas (see K2u)
  unsigned char *tiles = smem + (size_t)wid * (NST * TILE);
  uint64_t *bars = (uint64_t *)(smem + (size_t)(SDE_BLOCK / 32) * (NST * TILE)) + wid * NST;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; s++) mbar_init(smem_u32(bars + s), 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int64_t wbase = ((int64_t)blockIdx.x * (SDE_BLOCK / 32) + wid) * 32;   // a.npaths is a multiple of g
  if (wbase >= a.npaths) return;
  const int nrows = (int)(a.nsteps + 1);
  const int64_t pitch = (int64_t)nrows * ROWB;
  const int g = sde_tma_group(pitch, TALIGN), rpb = 32 / g;
  // head rows of the g <= 8 sub-paths, four bits each in one word (not an indexed array: no local memory)
  uint32_t hpack = 0;
  int hmax = 0;
  for (int j = 1; j < g; j++) {
    const int hj = sde_tma_head(j, pitch, ROWB, TALIGN);
    hpack |= (uint32_t)hj << (4 * j);
    hmax = hj > hmax ? hj : hmax;
  }
  auto hh = [&](int j) { return (int)((hpack >> (4 * j)) & 15u); };
  const int ntile = nrows > hmax ? (nrows - hmax) / TS : 0;
  const int r = (lane % rpb) * g + lane / rpb;        // lane <-> tile row `lane` <-> path wbase + r
  const int h = hh(lane / rpb);
  const int64_t p = wbase + r;
  const bool live = p < a.npaths;
  const int c1 = (int)(wbase / g);
  const sde_tmap *tmw = &a.tm_w, *tmx = &a.tm_x;
  auto issue_load = [&](int k) {  // lane 0
    if (k < ntile) {
      const int s = k % NST;
      const uint32_t bar = smem_u32(bars + s);
      mbar_expect_tx(bar, TILE);
      for (int j = 0; j < g; j++)
        tma_load_2d(smem_u32(tiles + s * TILE + j * rpb * 128), tmw, j * nrows * D + (hh(j) + k * TS) * D, c1, bar);
    }
  };
  if (lane == 0)
    for (int k = 0; k < NST - 2; k++) issue_load(k);

  Consts<T> cst;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dt = (T)a.dt, t0 = (T)a.t0;
  Model::prepare(prm, dt, cst.c);
  const T *wp = (const T *)a.out2 + p * (int64_t)nrows * D;
  T *xp = (T *)a.out + p * (int64_t)nrows * D;
  T x[SDE_MAXD], wprev[D], dw[D];
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++) x[d] = d < D ? (T)a.x0[d] : (T)0;
#pragma unroll
  for (int d = 0; d < D; d++) wprev[d] = (T)0;
  // one row of the path: v[] holds w[row] on entry and x[row] on exit
  auto row_step = [&](int row, T *v) {
    if (row == 0) {
#pragma unroll
      for (int d = 0; d < D; d++) wprev[d] = v[d];
    } else {
      const T t = t0 + (T)(row - 1) * dt;
#pragma unroll
      for (int d = 0; d < D; d++) { dw[d] = v[d] - wprev[d]; wprev[d] = v[d]; }
      Model::template step<SCHEME>(cst.c, t, dt, dw, x);
    }
#pragma unroll
    for (int d = 0; d < D; d++) v[d] = x[d];
  };
  if (live)
    for (int row = 0; row < h && row < nrows; row++) {   // head rows before the first aligned row (h <= 3)
      T v[D];
#pragma unroll
      for (int d = 0; d < D; d++) v[d] = wp[(int64_t)row * D + d];
      row_step(row, v);
#pragma unroll
      for (int d = 0; d < D; d++) xp[(int64_t)row * D + d] = v[d];
    }
  const int sw = lane & 7;
  for (int k = 0; k < ntile; k++) {
    if (lane == 0) {
      bulk_wait_read<1>();  // the stage tile k + NST - 2 goes into was stored from at iteration k - 2
      issue_load(k + NST - 2);
    }
    const int s = k % NST;
    mbar_wait(smem_u32(bars + s), (uint32_t)((k / NST) & 1));
    unsigned char *trow = tiles + s * TILE + lane * 128;
    if (live) {
      const int row0 = h + k * TS;
#pragma unroll
      for (int gi = 0; gi < NG; gi++) {
        union { uint4 q[GQ]; T v[G * D]; } u;
#pragma unroll
        for (int c = 0; c < GQ; c++) u.q[c] = *(const uint4 *)(trow + (((gi * GQ + c) ^ sw) << 4));
#pragma unroll
        for (int i = 0; i < G; i++) row_step(row0 + gi * G + i, u.v + i * D);
#pragma unroll
        for (int c = 0; c < GQ; c++) *(uint4 *)(trow + (((gi * GQ + c) ^ sw) << 4)) = u.q[c];
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      for (int j = 0; j < g; j++)
        tma_store_2d(tmx, j * nrows * D + (hh(j) + k * TS) * D, c1, smem_u32(tiles + s * TILE + j * rpb * 128));
      bulk_commit();
    }
  }
  // rows past the last full tile: one cooperative tile over rows [T0, nrows) of the warp's paths (natural order).
  // Its loads go to registers first so that all of them are in flight together (measured: 4.45 -> 5.5 TB/s for 4-byte
  // rows, where the tail is 33 of 513 rows; issuing them before the drain of the last tensor stores gained nothing).
  if (lane == 0) bulk_wait_read<0>();   // the ring is free again (and must outlive the stores' reads)
  __syncwarp();
  const int T0 = ntile * TS, ncols = (nrows - T0) * D;
  const int64_t rem = a.npaths - wbase;
  const int nlive = (int)(rem < 32 ? rem : 32);
  T tv[TROW];
  if (ncols > 0) {
    const T *gw = (const T *)a.out2 + (wbase * nrows + T0) * D;
#pragma unroll
    for (int k = 0; k < TROW; k++) {
      const int e = k * 32 + lane, rr = e / TROW, c = e - rr * TROW;
      tv[k] = (rr < nlive && c < ncols) ? __ldcs(gw + (int64_t)rr * nrows * D + c) : (T)0;
    }
  }
  __syncwarp();   // also a scheduling fence: without it the compiler pairs each load with its shared-memory store and the
                  // in-order issue serialises the round trips (measured 4.4 instead of 5.5 TB/s for 4-byte rows)
  if (ncols > 0) {
    T *tt = (T *)tiles;
#pragma unroll
    for (int k = 0; k < TROW; k++) {
      const int e = k * 32 + lane, rr = e / TROW, c = e - rr * TROW;
      tt[rr * TLD + c] = tv[k];
    }
    __syncwarp();
    if (live)
      for (int row = T0 + h; row < nrows; row++) row_step(row, tt + r * TLD + (row - T0) * D);
    __syncwarp();
    tail_tile_store<T, TROW, TLD>(tt, (T *)a.out + (wbase * nrows + T0) * D, (int64_t)nrows * D, nlive, ncols, hpack, D, g - 1, lane);
  }
  __syncwarp();
  count_nonfinite(a, live, x, D);
}

// ---------------------------------------------------------------------------------------------------
// K2u full trajectories in the reference's order [path][time][D] with TMA tensor stores, ONE SUB-PATH PER WARP
//     (round 2; same results as K2b / K2t, bit for bit).  In K2t a warp's 32 lanes cover the g sub-paths of 32/g map
//     rows, so lanes sit at different phases h_j of the 16-byte grid and every row costs its own narrow shared-memory
//     store plus ring arithmetic.  Here a warp owns sub-path j of 32 consecutive map rows (paths (32 br + lane) g + j): all
//     lanes share the phase h_j, and the rows of a Philox batch are written as WHOLE 16-byte chunks — one STS.128 per two
//     FP64 rows / four FP32 rows — at chunk positions that follow from warp-uniform counters.  Rows of a batch that do not
//     fill a chunk (phase PH = (1 - h_j) mod rows-per-chunk) ride in registers into the next batch (`carry`).
//     Tiles: [32 lanes][128 B], 128-byte swizzle, two stages; lane 0 stores a finished tile with one tensor store (box =
//     128 B x 32 map rows) while the warp fills the other stage.  Rows before the first aligned row and after the last
//     full tile are plain stores.  Needs rows of 4, 8 or 16 bytes and batches of whole chunks (U * ROWB a multiple of 16
//     dividing 128); the host routes everything else to K2b.
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T, int SCHEME, int V = 1>
SDE_D void simulate_ref_tma2_body(const sde_args &a) {
  constexpr int D = Model::D, M = Model::M;
  typedef Batch<T, M, V> B;
  typedef Noise<Model, T, V> NS;
  constexpr int U = B::U;
  constexpr int ROWB = D * (int)sizeof(T);
  constexpr int BB = U * ROWB;                                   // bytes per batch and lane
  constexpr int TROWB = sizeof(T) == 8 ? SDE_SIMRT2_ROWB_F64 : SDE_SIMRT2_ROWB_F32;   // bytes per path and tile (128 or 64)
  static_assert(TROWB == 128 || TROWB == 64, "tile rows of 128 bytes (128-byte swizzle) or 64 bytes (64-byte swizzle)");
  constexpr int NCH = TROWB / 16;                                // 16-byte chunks per tile row
  constexpr int TILEB = 32 * TROWB;                              // bytes per tile
  constexpr bool OK = (16 % ROWB == 0) && (BB % 16 == 0) && (TROWB % BB == 0) && M > 0;
  constexpr int RPC = OK ? 16 / ROWB : 1;                        // rows per 16-byte chunk
  constexpr int NCB = OK ? BB / 16 : 1;                          // chunks per batch
  constexpr int TS = OK ? TROWB / ROWB : 1;                      // rows per tile
  constexpr int EPC = 16 / (int)sizeof(T);                       // elements per chunk
  __shared__ double s_tab[TabSize<T, V>::value];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  load_tables<T, V>(s_tab, a.noise_var);
  if (!OK) return;
  unsigned char *smem = s_dyn + ((1024u - (smem_u32(s_dyn) & 1023u)) & 1023u);
  // the warp index through a shuffle: ptxas then knows it is warp-uniform, and the tensor-store operands derived from it (tile
  // address, coordinates) sit in uniform registers instead of being made uniform by a vote loop before every UTMASTG
  const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  // MANUAL: a finished tile leaves through 8 LDS.128 + 8 STG.128 per warp (16-byte chunks, four 128-byte row segments per
  // warp instruction) instead of a tensor store — synchronous, so one stage is enough
  constexpr bool MANUAL = sizeof(T) == 8 ? (SDE_SIMRT2_MANUAL_F64 != 0) : (SDE_SIMRT2_MANUAL_F32 != 0);
  constexpr int NSTG = MANUAL ? 1 : (sizeof(T) == 8 ? SDE_SIMRT2_STAGES_F64 : SDE_SIMRT2_STAGES_F32);   // ring depth: tiles in flight per warp
  static_assert(!MANUAL || TROWB == 128, "the manual flush is written for 128-byte tile rows");
  unsigned char *tiles = smem + (size_t)wid * (NSTG * TILEB);
  PhiloxKeys ks;
  philox_expand(a.seed_lo, a.seed_hi, ks);
  Consts<T> cst;
  T prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = i < Model::NP ? (T)a.params[i] : (T)0;
  const T dt = (T)a.dt, t0 = (T)a.t0;
  Model::prepare(prm, dt, cst.c);

  const int nrows = (int)(a.nsteps + 1);
  const int64_t pitch = (int64_t)nrows * ROWB;
  constexpr int TALIGN = sizeof(T) == 8 ? SDE_SIMRT2_ALIGN_F64 : SDE_SIMRT2_ALIGN_F32;
  const int g = sde_tma_group(pitch, TALIGN);
  const int64_t maprows = a.npaths / g;                           // a.npaths is a multiple of g
  const int64_t unit = (int64_t)blockIdx.x * (SDE_BLOCK / 32) + wid;
  const int j = (int)(unit % g);
  const int64_t br = unit / g;
  if (br * 32 >= maprows) return;
  const int64_t mr = br * 32 + lane;
  const bool live = mr < maprows;
  const int64_t p = (live ? mr : maprows - 1) * g + j;            // dead lanes shadow a live path (their tile rows are clipped)
  const int h = sde_tma_head(j, pitch, ROWB, TALIGN);             // first aligned row of this sub-path: warp-uniform
  const int ntile = nrows > h ? (nrows - h) / TS : 0;
  const int qend = ntile * TS;                                    // ring rows are q = row - h in [0, qend)
  const sde_tmap *tmx = &a.tm_x;
  T *xp = (T *)a.out + p * (int64_t)nrows * D;
  unsigned char *trow = tiles + lane * TROWB;
  const uint32_t tiles_s = smem_u32(tiles);
  const int c0base = j * nrows * D + h * D, c1 = (int)(br * 32);
  // chunk c of this lane's tile row under the shared-memory swizzle: 128-byte mode XORs address bits 4-6 with bits 7-9
  // (= lane & 7 for 128-byte rows), 64-byte mode bits 4-5 with bits 7-8 (= (lane >> 1) & 3 for 64-byte rows)
  const uint32_t sw = TROWB == 128 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);

  auto acquire = [&](int k) {   // about to write tile k: the tensor store of tile k - NSTG must have read its stage
    if (!MANUAL && k >= NSTG) {
      if (lane == 0) bulk_wait_read<(NSTG > 1 ? NSTG - 1 : 0)>();
      __syncwarp();
    }
  };
  auto flush = [&](int k) {     // tile k is complete in shared memory
    if (MANUAL) {
      __syncwarp();
      const unsigned char *src = tiles + (k % NSTG) * TILEB;
      unsigned char *gbase = (unsigned char *)a.out + ((int64_t)j * nrows + h + (int64_t)k * TS) * ROWB;   // sub-path j, first row of tile k
      const int cc = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int r = 4 * i + (lane >> 3);                                 // tile row = map row br*32 + r
        const uint4 v = *(const uint4 *)(src + r * 128 + ((cc ^ (r & 7)) << 4));
        const int64_t mrr = br * 32 + r;
        if (mrr < maprows) __stcs((uint4 *)(gbase + mrr * (int64_t)g * pitch + cc * 16), v);
      }
      __syncwarp();
      return;
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
#if SDE_SIMRT2_DIAG == 1      /* timing experiment only (wrong results): no tensor store at all = the kernel's compute ceiling */
#elif SDE_SIMRT2_DIAG == 2    /* timing experiment only: every warp rewrites its first two tiles = tensor stores that stay in L2 */
      tma_store_2d(tmx, c0base + (k & 1) * TS * D, c1, tiles_s + (uint32_t)(k % NSTG) * (uint32_t)TILEB);
#else
      tma_store_2d(tmx, c0base + k * TS * D, c1, tiles_s + (uint32_t)(k % NSTG) * (uint32_t)TILEB);
#endif
      bulk_commit();
    }
  };
  T x[SDE_MAXD];
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++) x[d] = d < D ? (T)a.x0[d] : (T)0;
  // one row, any position: ring (narrow store) or plain global store
  auto emit_row = [&](int r, const T *v) {
    const int q = r - h;
    if (q >= 0 && q < qend) {
      const int k = q / TS, qq = q - k * TS;
      if (qq == 0) acquire(k);
      const uint32_t off = (uint32_t)qq * ROWB;
      unsigned char *dst = trow + (k % NSTG) * TILEB + ((((off >> 4) ^ sw) << 4) | (off & 15u));
#pragma unroll
      for (int d = 0; d < D; d++) ((T *)dst)[d] = v[d];
      if (qq == TS - 1) flush(k);
    } else if (live) {
#pragma unroll
      for (int d = 0; d < D; d++) __stcs(xp + (int64_t)r * D + d, v[d]);
    }
  };
  emit_row(0, x);

  const uint64_t gpath = (uint64_t)(a.path0 + p);
  PhiloxPath pp;
  philox_path_init(gpath, a.stream, ks, pp, sizeof(T) == 8);
  PinnedCoef pc(a.lk, sizeof(T) == 8, V);
  const int64_t nbatch = (a.nsteps + U - 1) / U, nfull = a.nsteps / U;
  // the last RPC - 1 rows produced (newest last): what a batch leaves over for the next one's first chunk
  T carry[RPC > 1 ? RPC - 1 : 1][D];
#pragma unroll
  for (int c = 0; c < (RPC > 1 ? RPC - 1 : 1); c++)
#pragma unroll
    for (int d = 0; d < D; d++) carry[c][d] = x[d];
  auto slow_batch = [&](int64_t it) {
    T dw[NS::NV];
    gen_batch<T, M, V, NS::POLAR>(pp, gpath, (uint32_t)(it * B::NB), a.stream, ks, s_tab, a.lk, (T)a.noise_var, dw, pc.v);
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t i = it * U + u;
      if (i < a.nsteps) {
        const T t = t0 + (T)i * dt;
        NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, dw + u * NS::STRIDE, x);
        emit_row((int)i + 1, x);
#pragma unroll
        for (int c = 0; c + 1 < RPC - 1; c++)
#pragma unroll
          for (int d = 0; d < D; d++) carry[c][d] = carry[c + 1][d];
        if (RPC > 1) {
#pragma unroll
          for (int d = 0; d < D; d++) carry[RPC - 2][d] = x[d];
        }
      }
    }
  };
  // PH rows of a batch (the newest ones) wait in `carry` for the next batch; chunk number of a batch's first chunk:
  // cs(it) = it * NCB - mshift (q = cs * RPC is the ring row its first element lands on)
  const int PH = (((1 - h) % RPC) + RPC) % RPC;
  const int mshift = (PH - (1 - h)) / RPC;                         // 0, or 1 when h >= 2 (4-byte rows)
  int64_t it = 0;
  const int64_t it0 = (mshift + NCB - 1) / NCB;                     // first batch whose chunks are all ring rows (q >= 0)
  for (; it < it0 && it < nbatch; it++) slow_batch(it);
  // fast batches: it0 <= it < it1, all chunks inside full tiles and all U steps real
  int64_t it1 = ((int64_t)ntile * NCH + mshift) / NCB;
  if (it1 > nfull) it1 = nfull;
  // one batch in the steady state, phase PHC: NCB aligned 16-byte stores, new carry.  Warp-uniform running state:
  // cpos = chunk position inside the tile (0..7), kt = tile index, blk = Philox block of the batch, tb = first step
  const T nv = (T)a.noise_var;
  int cpos = 0, kt = 0, ks_ = 0;                                   // ks_ = kt % NSTG
  uint32_t blk = 0;
  int tb = 0;
  T dwn[NS::NV];                                                    // SDE_SIMRT2_PREFETCH: the normals of the next batch
  auto fast_batch = [&](auto phc, auto msc) {
    constexpr int PHC = decltype(phc)::value;
    constexpr int MS = decltype(msc)::value;   // mshift % NCB: chunks of a tile-crossing batch that still belong to the old tile
    T dw[NS::NV];
#if SDE_SIMRT2_PREFETCH
    // software pipeline: the normals of the NEXT batch are produced while this batch's steps run — the generator does not
    // depend on the state, so the two chains interleave in one basic block (one path per lane has no other parallelism)
#pragma unroll
    for (int i = 0; i < NS::NV; i++) dw[i] = dwn[i];
    gen_batch<T, M, V, NS::POLAR>(pp, gpath, blk + (uint32_t)B::NB, a.stream, ks, s_tab, a.lk, nv, dwn, pc.v);
#else
    gen_batch<T, M, V, NS::POLAR>(pp, gpath, blk, a.stream, ks, s_tab, a.lk, nv, dw, pc.v);
#endif
    T seq[PHC + U][D];                                             // carry rows, then the U new rows
#pragma unroll
    for (int c = 0; c < PHC; c++)
#pragma unroll
      for (int d = 0; d < D; d++) seq[c][d] = carry[(RPC - 1) - PHC + c][d];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const T t = t0 + (T)(tb + u) * dt;
      NoiseStep<Model, SCHEME, NS::POLAR, V>::run(cst.c, t, dt, dw + u * NS::STRIDE, x);
#pragma unroll
      for (int d = 0; d < D; d++) seq[PHC + u][d] = x[d];
    }
    auto put = [&](int c, unsigned char *stage, int pos) {   // chunk c of the batch -> chunk position pos of the lane's tile row
      union { uint4 q4; T v[EPC]; } w;
#pragma unroll
      for (int e = 0; e < EPC; e++) w.v[e] = seq[(c * EPC + e) / D][(c * EPC + e) % D];
      *(uint4 *)(stage + ((((uint32_t)pos) ^ sw) << 4)) = w.q4;
    };
    auto next_tile = [&]() {   // the stage of the next tile is acquired right away (once per tile, not a check per batch)
      flush(kt);
      kt++;
      ks_ = ks_ + 1 == NSTG ? 0 : ks_ + 1;
      acquire(kt);
    };
    // 16-byte rows with several rows per batch (RPC == 1, NCB > 1: row r is chunk r, one past the batch grid) may cross a
    // tile boundary INSIDE a batch at any chunk: boundary checks per chunk there.  Everywhere else the first chunk of a
    // batch sits at cpos = -mshift (mod NCB): on the batch grid when MS == 0 (boundary check once per batch); off it when the
    // head rows are h_j >= 2 (only the 32-byte alignment produces them), and then every NCH / NCB-th batch crosses into the
    // next tile after exactly MS chunks — one warp-uniform branch per batch.
    constexpr bool PERCHUNK = (RPC == 1) && (NCB > 1);
    if (PERCHUNK) {
#pragma unroll
      for (int c = 0; c < NCB; c++) {
        put(c, trow + ks_ * TILEB, cpos);
        if (++cpos == NCH) {
          next_tile();
          cpos = 0;
        }
      }
    } else if (MS == 0 || cpos + NCB <= NCH) {
      unsigned char *stage = trow + ks_ * TILEB;
#pragma unroll
      for (int c = 0; c < NCB; c++) put(c, stage, cpos + c);
      cpos += NCB;
      if (MS == 0 && cpos == NCH) {
        next_tile();
        cpos = 0;
      }
    } else {   // cpos == NCH - MS
      unsigned char *stage = trow + ks_ * TILEB;
#pragma unroll
      for (int c = 0; c < MS; c++) put(c, stage, NCH - MS + c);
      next_tile();
      stage = trow + ks_ * TILEB;
#pragma unroll
      for (int c = MS; c < NCB; c++) put(c, stage, c - MS);
      cpos = NCB - MS;
    }
    if (RPC > 1) {
#pragma unroll
      for (int c = 0; c < RPC - 1; c++)
#pragma unroll
        for (int d = 0; d < D; d++) carry[c][d] = seq[U + PHC - (RPC - 1) + c >= 0 ? U + PHC - (RPC - 1) + c : 0][d];
    }
    blk += (uint32_t)B::NB;
    tb += U;
  };
  constexpr int HMAXR = TALIGN > ROWB ? TALIGN / ROWB - 1 : 0;                       // bound on the head rows h_j
  constexpr int MSMAX = (RPC > 1 && NCB > 1) ? ((HMAXR - 1 + RPC - 1) / RPC) : 0;    // bound on mshift where it matters
  auto run_fast = [&](auto phc) {
    constexpr int PHC = decltype(phc)::value;
    const int64_t cs0 = it * NCB - mshift;                          // >= 0 by the choice of it0
    cpos = (int)(cs0 % NCH);
    kt = (int)(cs0 / NCH);
    ks_ = kt % NSTG;
    blk = (uint32_t)(it * B::NB);
    tb = (int)(it * U);
    if (cpos == 0) acquire(kt);
#if SDE_SIMRT2_PREFETCH
    gen_batch<T, M, V, NS::POLAR>(pp, gpath, blk, a.stream, ks, s_tab, a.lk, nv, dwn, pc.v);
#endif
    const int ms = MSMAX > 0 ? mshift % NCB : 0;
    int nb = (int)(it1 - it);                                       // 32-bit loop counter (nsteps < 2^28)
    it = it1;
    if (MSMAX >= 1 && ms == 1) {
      for (; nb > 0; nb--) fast_batch(phc, IntC<(MSMAX >= 1 && NCB > 1 ? 1 : 0)>());
    } else if (MSMAX >= 2 && ms == 2) {
      for (; nb > 0; nb--) fast_batch(phc, IntC<(MSMAX >= 2 && NCB > 2 ? 2 : 0)>());
    } else {
      for (; nb > 0; nb--) fast_batch(phc, IntC<0>());
    }
    // the newest PHC rows are still in registers only: rows it*U - PHC + 1 .. it*U (compile-time indices: no local memory)
#pragma unroll
    for (int c = 0; c < PHC; c++) emit_row((int)(it * U) - PHC + 1 + c, carry[(RPC - 1) - PHC + c]);
  };
  if (it < it1) {
    if (PH == 0) run_fast(IntC<0>());
    else if (PH == 1) run_fast(IntC<(RPC > 1 ? 1 : 0)>());
    else if (PH == 2) run_fast(IntC<(RPC > 2 ? 2 : 0)>());
    else run_fast(IntC<(RPC > 3 ? 3 : 0)>());
  }
  for (; it < nbatch; it++) slow_batch(it);                          // partial last tile, tail rows
  if (!MANUAL && lane == 0) bulk_wait_read<0>();   // shared memory must outlive the stores' reads
  __syncwarp();
  count_nonfinite(a, live, x, D);
}

// ---------------------------------------------------------------------------------------------------
// K7  logtransition(model, ::EulerMaruyama, times, data)  [schemes.jl:64-76, euler_maruyama.jl:8-13]:
//     Gaussian one-step log-densities along stored trajectories, x (a.out2) [path][row][D] -> a.out [path][row-1].
//     Streaming and HBM-bound (D reads, 1 write per transition); same tile staging as K3.  The arithmetic is done
//     in double for both storage precisions (it is free next to the memory traffic).
//       μ = x0 + drift Δt;  Σ = Δt σ σ';  z = μ - x1;  -0.5 (D log 2π + log det Σ + z' Σ⁻¹ z)
// ---------------------------------------------------------------------------------------------------
#ifndef SDE_LOGT_FASTMATH
#define SDE_LOGT_FASTMATH 1 /* log_fast / rcp_fast instead of the library log and IEEE divisions for D <= 2 (ncu of the library form:
                               FP64 pipe 42 %, issue slots 63 %, 2.7 TB/s — bound by those two, not by its loads) */
#endif
// Every operation of the D <= 2 forms is written with an explicit rounding (no contraction): the branch-free twin below,
// the batched callers and the library fallback then agree bit for bit, and the order is the oracle's.
template <int D> SDE_D double gauss_logpdf_terms(const double (&S)[SDE_MAXD][SDE_MAXD], const double *z, const double *ltab) {
  double logdet, q;
  constexpr double LOG2PI = 0x1.d67f1c864beb5p+0;
  if (D == 1) {
#if SDE_LOGT_FASTMATH
    logdet = log_fast(S[0][0], ltab);
    q = __dmul_rn(z[0], __dmul_rn(z[0], rcp_fast(S[0][0])));
#else
    logdet = log(S[0][0]);
    q = __dmul_rn(z[0], z[0] / S[0][0]);
#endif
    return __dmul_rn(-0.5, __dadd_rn(__dadd_rn(LOG2PI, logdet), q));
  } else if (D == 2) {  // closed forms (StaticArrays: det = a11 a22 - a12 a21, Cramer solve)
    const double det = __dsub_rn(__dmul_rn(S[0][0], S[1][1]), __dmul_rn(S[0][1], S[1][0]));
    const double n0 = __dsub_rn(__dmul_rn(S[1][1], z[0]), __dmul_rn(S[0][1], z[1]));
    const double n1 = __dsub_rn(__dmul_rn(S[0][0], z[1]), __dmul_rn(S[1][0], z[0]));
#if SDE_LOGT_FASTMATH
    const double rd = rcp_fast(det);
    const double y0 = __dmul_rn(n0, rd), y1 = __dmul_rn(n1, rd);
    logdet = log_fast(det, ltab);
#else
    const double y0 = n0 / det, y1 = n1 / det;
    logdet = log(det);
#endif
    q = __dadd_rn(__dmul_rn(z[0], y0), __dmul_rn(z[1], y1));
    return __dmul_rn(-0.5, __dadd_rn(__dadd_rn(2.0 * LOG2PI, logdet), q));
  } else {  // Cholesky
    double L[SDE_MAXD][SDE_MAXD], y[SDE_MAXD];
    double ld = 0.0;
    bool ok = true;
#pragma unroll
    for (int a = 0; a < D; a++) {
#pragma unroll
      for (int b = 0; b <= a; b++) {
        double s2 = S[a][b];
#pragma unroll
        for (int c = 0; c < b; c++) s2 -= L[a][c] * L[b][c];
        if (a == b) {
          ok = ok && (s2 > 0.0);
          L[a][a] = sqrt(s2);
          ld += 2.0 * log(L[a][a]);
        } else {
          L[a][b] = s2 / L[b][b];
        }
      }
    }
    q = 0.0;
#pragma unroll
    for (int a = 0; a < D; a++) {
      double s2 = z[a];
#pragma unroll
      for (int c = 0; c < a; c++) s2 -= L[a][c] * y[c];
      y[a] = s2 / L[a][a];
      q += y[a] * y[a];
    }
    logdet = ok ? ld : nan("");
  }
  return -0.5 * (((double)D * 0x1.d67f1c864beb5p+0 /* log 2π */ + logdet) + q);
}
// The same values without a branch (D <= 2, fast math): `ok` is cleared when an argument leaves the range of log_fast_core /
// rcp_fast_core and the caller then takes gauss_logpdf_terms for the whole batch.  Several independent transitions evaluated
// side by side in one basic block give the FP64 pipe the instruction-level parallelism a single dependent chain does not.
template <int D> SDE_D double gauss_logpdf_terms_nb(const double (&S)[SDE_MAXD][SDE_MAXD], const double *z, const double *ltab, bool &ok) {
  constexpr double LOG2PI = 0x1.d67f1c864beb5p+0;
#if SDE_LOGT_FASTMATH
  if (D == 1) {
    ok = ok && log_fast_ok(S[0][0]) && rcp_fast_ok(S[0][0]);
    const double logdet = log_fast_core(S[0][0], ltab);
    const double q = __dmul_rn(z[0], __dmul_rn(z[0], rcp_fast_core(S[0][0])));
    return __dmul_rn(-0.5, __dadd_rn(__dadd_rn(LOG2PI, logdet), q));
  } else if (D == 2) {
    const double det = __dsub_rn(__dmul_rn(S[0][0], S[1][1]), __dmul_rn(S[0][1], S[1][0]));
    const double n0 = __dsub_rn(__dmul_rn(S[1][1], z[0]), __dmul_rn(S[0][1], z[1]));
    const double n1 = __dsub_rn(__dmul_rn(S[0][0], z[1]), __dmul_rn(S[1][0], z[0]));
    ok = ok && log_fast_ok(det) && rcp_fast_ok(det);
    const double rd = rcp_fast_core(det);
    const double y0 = __dmul_rn(n0, rd), y1 = __dmul_rn(n1, rd);
    const double logdet = log_fast_core(det, ltab);
    const double q = __dadd_rn(__dmul_rn(z[0], y0), __dmul_rn(z[1], y1));
    return __dmul_rn(-0.5, __dadd_rn(__dadd_rn(2.0 * LOG2PI, logdet), q));
  }
#endif
  return gauss_logpdf_terms<D>(S, z, ltab);
}
// NB consecutive transitions of one path: xs[0 .. NB] are rows row1 - 1 .. row1 + NB - 1, res[i] the log-density of xs[i + 1]
// given xs[i].  μ, Σ and z are formed once; the result of a transition does not depend on how rows are batched.
template <class Model, int NB>
SDE_D void logt_batch(const double *prm, double t0, double dt, int row1, const double (&xs)[NB + 1][SDE_MAXD], const double *ltab,
                      double (&res)[NB]) {
  constexpr int D = Model::D, M = Model::M;
  double S[NB][SDE_MAXD][SDE_MAXD], z[NB][SDE_MAXD];
#pragma unroll
  for (int i = 0; i < NB; i++) {
    double mu[SDE_MAXD], sig[SDE_MAXD * SDE_MAXM];
#pragma unroll
    for (int q = 0; q < SDE_MAXD * SDE_MAXM; q++) sig[q] = 0.0;
    Model::coefficients(prm, t0 + (double)(row1 + i - 1) * dt, xs[i], mu, sig);
#pragma unroll
    for (int p = 0; p < D; p++) z[i][p] = (xs[i][p] + mu[p] * dt) - xs[i + 1][p];
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < M; j++) {
          const double term = (dt * sig[p * SDE_MAXM + j]) * sig[q * SDE_MAXM + j];
          acc = j == 0 ? term : acc + term;
        }
        S[i][p][q] = acc;
      }
  }
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NB; i++) res[i] = gauss_logpdf_terms_nb<D>(S[i], z[i], ltab, ok);
  if (!ok) {
#pragma unroll
    for (int i = 0; i < NB; i++) res[i] = gauss_logpdf_terms<D>(S[i], z[i], ltab);
  }
}

template <class Model, typename T>
SDE_D void logtransition_body(const sde_args &a) {
  constexpr int D = Model::D;
  constexpr int TS = (SDE_SIMW_ROWBYTES / (int)sizeof(T)) / D;  // rows per tile
  constexpr int XROW = TS * D, XLD = XROW | 1, OLD = TS | 1;
  extern __shared__ __align__(16) unsigned char s_dyn[];   // [warps][32][XLD] then [warps][32][OLD]
  T (*s_x)[32 * XLD] = (T (*)[32 * XLD])s_dyn;
  T (*s_o)[32 * OLD] = (T (*)[32 * OLD])((T *)s_dyn + (SDE_BLOCK / 32) * 32 * XLD);
  __shared__ double s_ltab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ltab[i] = g_log_tab[i];
  __syncthreads();
  double prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = a.params[i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wbase = (int64_t)blockIdx.x * blockDim.x + (wid << 5);
  if (wbase >= a.npaths) return;
  const int nlive = (int)((a.npaths - wbase) < 32 ? (a.npaths - wbase) : 32);
  const int64_t nrows = a.nsteps + 1;
  const T *xin = (const T *)a.out2;
  T *out = (T *)a.out;
  T *tx = s_x[wid], *to = s_o[wid];
  double xp[SDE_MAXD];
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++) xp[d] = 0.0;
  constexpr int NB = 4;   // transitions evaluated side by side (instruction-level parallelism for the FP64 pipe)
  for (int64_t row0 = 0; row0 < nrows; row0 += TS) {
    const int nr = (int)((nrows - row0) < TS ? (nrows - row0) : TS);
    tile_load<T, XROW, XLD>(tx, xin + (wbase * nrows + row0) * D, nrows * D, nlive, nr * D, lane);
    __syncwarp();
    int i = 0;
    if (row0 == 0) {   // row 0 starts the path: no transition ends there
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) xp[d] = d < D ? (double)tx[lane * XLD + d] : 0.0;
      i = 1;
    }
    if (lane >= nlive) i = nr;   // rows of paths past the end are zero-filled: nothing to evaluate
    for (; i + NB <= nr; i += NB) {
      double xs[NB + 1][SDE_MAXD], res[NB];
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) xs[0][d] = xp[d];
#pragma unroll
      for (int b = 0; b < NB; b++)
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xs[b + 1][d] = d < D ? (double)tx[lane * XLD + (i + b) * D + d] : 0.0;
      logt_batch<Model, NB>(prm, a.t0, a.dt, (int)row0 + i, xs, s_ltab, res);
#pragma unroll
      for (int b = 0; b < NB; b++) to[lane * OLD + i + b] = (T)res[b];   // transition ending at tile row i + b
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) xp[d] = xs[NB][d];              // carried into the next tile as well
    }
    for (; i < nr; i++) {
      double xs[2][SDE_MAXD], res[1];
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) { xs[0][d] = xp[d]; xs[1][d] = d < D ? (double)tx[lane * XLD + i * D + d] : 0.0; }
      logt_batch<Model, 1>(prm, a.t0, a.dt, (int)row0 + i, xs, s_ltab, res);
      to[lane * OLD + i] = (T)res[0];
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) xp[d] = xs[1][d];
    }
    __syncwarp();
    // tile row i holds the transition (row0 + i - 1) -> (row0 + i): output column row0 + i - 1; row 0 has none
    const int i0 = row0 == 0 ? 1 : 0;
    if (nr > i0)
      tile_store<T, TS, OLD>(to + i0, out + wbase * (nrows - 1) + (row0 - 1 + i0), nrows - 1, nlive, nr - i0, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// K7t logtransition on TMA tensor LOADS (same results as K7, bit for bit): a warp owns 32 paths; lane 0 moves tiles of x
//     global -> shared through a ring of SDE_LOGTT_STAGES stages (geometry of K3t: g paths per map row, head rows h_j,
//     128-byte swizzle; NST - 1 loads in flight per warp), every lane walks its own path in the swizzled tile and leaves
//     the log-density of the transition ENDING at tile row i in element slot i of its tile row (slot i lies at or before
//     row i's own bytes, which are in registers by then), and the warp writes the TS results of two / one / ... paths
//     per instruction as whole row segments of out [path][row-1].  The outputs' pitch (nrows-1) differs from the inputs',
//     so their 16-byte phases differ too: plain coalesced stores, no tensor map.  Head rows (< h_j <= 3) are plain
//     accesses, the rows after the last full tile go through one cooperative shared-memory tile like K3t's.
//     Needs D*sizeof(T) | 128; the host routes everything else and the npaths % g leftover to K7.
// ---------------------------------------------------------------------------------------------------
template <class Model, typename T>
SDE_D void logtransition_tma_body(const sde_args &a) {
  constexpr int D = Model::D;
  constexpr int ROWB = D * (int)sizeof(T);
  constexpr bool OK = (128 % ROWB == 0);
  if (!OK) return;
  constexpr int NST = SDE_LOGTT_STAGES;
  constexpr int TS = OK ? 128 / ROWB : 1;                 // rows per tile
  constexpr int NB = TS < 4 ? TS : 4;                     // transitions evaluated side by side
  constexpr int NQ = OK ? NB * ROWB / 16 : 1;             // 16-byte chunks per batch (ROWB >= 4, NB = 4 or ROWB >= 64)
  constexpr int TILE = 32 * 128;
  constexpr int TALIGN = SDE_SIMWT_ALIGN;
  constexpr int HMAX = TALIGN > ROWB ? TALIGN / ROWB - 1 : 0;
  constexpr int TOUT = TS + HMAX, TROW = TOUT * D, TLD = TROW | 1;   // the cooperative tail tile [32][TLD]
  static_assert(!OK || NST < 2 || 32 * TLD * (int)sizeof(T) <= 2 * TILE, "tail tile does not fit");
  __shared__ double s_ltab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ltab[i] = g_log_tab[i];
  __syncthreads();
  extern __shared__ __align__(16) unsigned char s_dyn[];
  unsigned char *smem = s_dyn + ((1024u - (smem_u32(s_dyn) & 1023u)) & 1023u);
  const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for ptxas (see K2u)
  unsigned char *tiles = smem + (size_t)wid * (NST * TILE);
  uint64_t *bars = (uint64_t *)(smem + (size_t)(SDE_BLOCK / 32) * (NST * TILE)) + wid * NST;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; s++) mbar_init(smem_u32(bars + s), 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int64_t wbase = ((int64_t)blockIdx.x * (SDE_BLOCK / 32) + wid) * 32;   // a.npaths is a multiple of g
  if (wbase >= a.npaths) return;
  const int nrows = (int)(a.nsteps + 1);
  const int64_t pitch = (int64_t)nrows * ROWB;
  const int g = sde_tma_group(pitch, TALIGN), lg = g == 1 ? 0 : (g == 2 ? 1 : (g == 4 ? 2 : 3)), rpb = 32 >> lg;
  uint32_t hpack = 0;   // head rows of the g <= 8 sub-paths, four bits each
  int hmax = 0;
  for (int j = 1; j < g; j++) {
    const int hj = sde_tma_head(j, pitch, ROWB, TALIGN);
    hpack |= (uint32_t)hj << (4 * j);
    hmax = hj > hmax ? hj : hmax;
  }
  auto hh = [&](int j) { return (int)((hpack >> (4 * j)) & 15u); };
  const int ntile = nrows > hmax ? (nrows - hmax) / TS : 0;
  const int r = (lane & (rpb - 1)) * g + (lane >> (5 - lg));   // lane <-> tile row `lane` <-> path wbase + r
  const int h = hh(lane >> (5 - lg));
  const int64_t p = wbase + r;
  const bool live = p < a.npaths;
  const int c1 = (int)(wbase >> lg);
  const sde_tmap *tmw = &a.tm_w;
  auto issue_load = [&](int k) {  // lane 0
    if (k < ntile) {
      const int s = k % NST;
      const uint32_t bar = smem_u32(bars + s);
      mbar_expect_tx(bar, TILE);
      for (int j = 0; j < g; j++)
        tma_load_2d(smem_u32(tiles + s * TILE + j * rpb * 128), tmw, j * nrows * D + (hh(j) + k * TS) * D, c1, bar);
    }
  };
  if (lane == 0)
    for (int k = 0; k < NST - 1; k++) issue_load(k);

  double prm[SDE_MAXP];
#pragma unroll
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = a.params[i];
  const T *xin = (const T *)a.out2 + p * (int64_t)nrows * D;
  T *out = (T *)a.out;
  T *op = out + p * (int64_t)(nrows - 1);
  double xp[SDE_MAXD];
#pragma unroll
  for (int d = 0; d < SDE_MAXD; d++) xp[d] = 0.0;
  // log-density of x[row] = xc given x[row - 1] = xp (row >= 1), the arithmetic of K7
  auto transition = [&](int row, const double *xc) {
    double xs[2][SDE_MAXD], res[1];
#pragma unroll
    for (int d = 0; d < SDE_MAXD; d++) { xs[0][d] = xp[d]; xs[1][d] = xc[d]; }
    logt_batch<Model, 1>(prm, a.t0, a.dt, row, xs, s_ltab, res);
    return res[0];
  };
  if (live)
    for (int row = 0; row < h && row < nrows; row++) {   // head rows before the first aligned row (h <= 3)
      double xc[SDE_MAXD];
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) xc[d] = d < D ? (double)xin[(int64_t)row * D + d] : 0.0;
      if (row > 0) op[row - 1] = (T)transition(row, xc);
#pragma unroll
      for (int d = 0; d < SDE_MAXD; d++) xp[d] = xc[d];
    }
  const int sw = lane & 7;
  const int nlive = (int)(a.npaths - wbase < 32 ? a.npaths - wbase : 32);
  auto store_tile = [&](auto gc, int k, const unsigned char *tile) {
    constexpr int GC = decltype(gc)::value;          // sub-paths per map row
    constexpr int RPI = OK ? 32 / TS : 1, RPB = 32 / GC;   // tile rows per warp instruction, tile rows per sub-path
    static_assert(!OK || (RPB % RPI == 0 && RPB >= RPI), "a warp instruction stays inside one sub-path");
    const int cc = lane & (TS - 1), tr0 = lane / TS;
    const uint32_t off = (uint32_t)(cc * (int)sizeof(T));
#pragma unroll
    for (int j = 0; j < GC; j++) {
      const int hj = hh(j);
      T *ptr = out + (wbase + (int64_t)tr0 * GC + j) * (int64_t)(nrows - 1) + (hj + k * TS + cc - 1);
      const bool skip = k == 0 && hj + cc == 0;      // column -1: row 0 starts the path
#pragma unroll
      for (int q = 0; q < RPB / RPI; q++) {
        const int tr = j * RPB + q * RPI + tr0;
        const T v = *(const T *)(tile + tr * 128 + ((((off >> 4) ^ (uint32_t)(tr & 7)) << 4) | (off & 15u)));
        if ((q * RPI + tr0) * GC + j < nlive && !skip) __stcs(ptr, v);
        ptr += (int64_t)RPI * GC * (nrows - 1);
      }
    }
  };
  for (int k = 0; k < ntile; k++) {
    if (lane == 0) issue_load(k + NST - 1);   // into the stage of tile k - 1: consumed and fenced at the end of that iteration
    const int s = k % NST;
    mbar_wait(smem_u32(bars + s), (uint32_t)((k / NST) & 1));
    unsigned char *tile = tiles + s * TILE, *trow = tile + lane * 128;
    if (live) {
      const int row0 = h + k * TS;
#pragma unroll
      for (int bi = 0; bi < TS / NB; bi++) {   // NB rows = NQ 16-byte chunks of the lane's tile row, evaluated side by side
        union { uint4 q[NQ]; T v[NB * D]; } u;
#pragma unroll
        for (int c = 0; c < NQ; c++) u.q[c] = *(const uint4 *)(trow + (((bi * NQ + c) ^ sw) << 4));
        double xs[NB + 1][SDE_MAXD], res[NB];
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
          for (int d = 0; d < SDE_MAXD; d++) xs[b + 1][d] = d < D ? (double)u.v[b * D + d] : 0.0;
        const bool first = row0 + bi * NB == 0;   // row 0 starts the path: its slot is never stored; evaluate x0 -> x0 there
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xs[0][d] = first ? xs[1][d] : xp[d];
        logt_batch<Model, NB>(prm, a.t0, a.dt, row0 + bi * NB, xs, s_ltab, res);
#pragma unroll
        for (int b = 0; b < NB; b++) {   // slot qi lies at or before the bytes of row qi, which are in registers
          const uint32_t off = (uint32_t)((bi * NB + b) * (int)sizeof(T));
          *(T *)(trow + ((((off >> 4) ^ (uint32_t)sw) << 4) | (off & 15u))) = (T)res[b];
        }
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xp[d] = xs[NB][d];
      }
    }
    __syncwarp();
    // the tile's results leave as whole row segments of out: one warp instruction covers 32 / TS tile rows; tile row tr holds
    // sub-path tr / rpb of map row tr % rpb, and its slot cc is the transition ending at row h_j + k TS + cc, i.e. output
    // column ... - 1.  g is a compile-time constant inside store_tile: what is left per instruction is one shared-memory
    // address, one pointer bump and the bound check of a partial warp.
    if (g == 1) store_tile(IntC<1>(), k, tile);
    else if (g == 2) store_tile(IntC<2>(), k, tile);
    else store_tile(IntC<4>(), k, tile);
    fence_async_smem();   // these generic-proxy reads come before the tensor load that refills the stage
    __syncwarp();
  }
  // rows past the last full tile: one cooperative tile over rows [T0, nrows) of the warp's paths (natural order)
  const int T0 = ntile * TS, ncols = (nrows - T0) * D;
  T tv[TROW];
  if (ncols > 0) {
    const T *gw = (const T *)a.out2 + (wbase * nrows + T0) * D;
#pragma unroll
    for (int k = 0; k < TROW; k++) {
      const int e = k * 32 + lane, rr = e / TROW, c = e - rr * TROW;
      tv[k] = (rr < nlive && c < ncols) ? __ldcs(gw + (int64_t)rr * nrows * D + c) : (T)0;
    }
  }
  __syncwarp();   // also a scheduling fence (see K3t)
  if (ncols > 0) {
    T *tt = (T *)tiles;
#pragma unroll
    for (int k = 0; k < TROW; k++) {
      const int e = k * 32 + lane, rr = e / TROW, c = e - rr * TROW;
      tt[rr * TLD + c] = tv[k];
    }
    __syncwarp();
    if (live)
      for (int row = T0 + h; row < nrows; row++) {
        double xc[SDE_MAXD];
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xc[d] = d < D ? (double)tt[r * TLD + (row - T0) * D + d] : 0.0;
        if (row > 0) tt[r * TLD + (row - T0)] = (T)transition(row, xc);   // slot i <= i*D: behind the rows still to be read
#pragma unroll
        for (int d = 0; d < SDE_MAXD; d++) xp[d] = xc[d];
      }
    __syncwarp();
    // tail tile column i = transition ending at row T0 + i -> output column T0 + i - 1; the columns before h_j were written above
    tail_tile_store<T, TOUT, TLD>(tt, out + wbase * (int64_t)(nrows - 1) + (T0 - 1), (int64_t)(nrows - 1), nlive, nrows - T0, hpack, 1,
                                  g - 1, lane);
  }
}

// K6  drift / diffusion evaluation at one point (functor known-answer tests, test/runtests.jl:61-82).
//     out[0..D) = drift, out[D .. D + D*max(M,1)) = diffusion, row-major.
template <class Model>
SDE_D void coef_eval_body(const sde_args &a) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double prm[SDE_MAXP], x[SDE_MAXD], mu[SDE_MAXD], sig[SDE_MAXD * SDE_MAXM];
  for (int i = 0; i < SDE_MAXP; i++) prm[i] = a.params[i];
  for (int d = 0; d < SDE_MAXD; d++) { x[d] = a.x0[d]; mu[d] = 0.0; }
  for (int i = 0; i < SDE_MAXD * SDE_MAXM; i++) sig[i] = 0.0;
  Model::coefficients(prm, a.t0, x, mu, sig);
  double *out = (double *)a.out;
  constexpr int D = Model::D, MW = Model::M > 0 ? Model::M : 1;
  for (int d = 0; d < D; d++) out[d] = mu[d];
  for (int i = 0; i < D; i++)
    for (int j = 0; j < MW; j++) out[D + i * MW + j] = sig[i * SDE_MAXM + j];
}

#define SDE_EM 0
#define SDE_MILSTEIN 1
// Resident CTAs per SM the FP64 terminal-value kernel is compiled for.  1 = "no occupancy target": ptxas then spends
// registers on scheduling (154 for Heston) and the kernel runs 12 warps per SM — measured FASTER (2.20e11 vs 2.10e11
// path-steps/s at 64 registers / 30 warps): the loop is dispatch-bound, not latency-bound.  The FP32 kernel keeps the
// default heuristics (0 = unspecified; with 1 it took 124 registers and lost 2 %).
#ifndef SDE_SAMPLE_MINB
#define SDE_SAMPLE_MINB 1
#endif
#ifndef SDE_SAMPLE_MINB_V2
#define SDE_SAMPLE_MINB_V2 2 /* stream v2 terminal kernel: two CTAs per SM.  Measured with the polar Heston step, G path-steps/s
                                (profiles/r02x_ab_occupancy.txt, BEFORE the time loop lost its per-step conditions): 4 paths per
                                thread 270.6 (204 registers, 8 warps per SM; 168 registers / 12 warps: the same), 2 paths per
                                thread 275.9 (118 registers, 16 warps) / 269.7 (96, 20 warps) / 256.1 (80, 24 warps), 1 path per
                                thread 250.1 (75 registers) / 231.2 (64); with the unconditional loop (r02y_ab_loop.txt) 4 paths
                                per thread lead again: 295.2 against 292.4 */
#endif
#define SDE_PPT_OF(T, V) ((sizeof(T) == 8 && (V) == 2) ? SDE_PPT_V2 : SDE_PPT)
#ifndef SDE_SAMPLE1_MINB
#define SDE_SAMPLE1_MINB 0
#endif
#ifndef SDE_SAMPLE_MINB_F32
#define SDE_SAMPLE_MINB_F32 0
#endif
#ifndef SDE_SIMRT_MINB
#define SDE_SIMRT_MINB 4
#endif
#ifndef SDE_SIMREF_MINB
#define SDE_SIMREF_MINB 5
#endif
#ifndef SDE_F64_MINB
#define SDE_F64_MINB 1 /* the same switch for the FP64 coupled-level and [time][D][path] kernels (+0.6 % / +1.1 %) */
#endif

// Kernel entry points for one model type, scheme, arithmetic type T, normal-stream version V and symbol suffix SUF
// (f64 | f32 | f64v2).  TAG makes the symbols unique.  They come in FAMILIES so that NVRTC can compile a generated model
// lazily, one family per (precision, stream, math) on first use (sde_nvrtc.cpp): TERMINAL = sample / sample1 / mlmc,
// TRAJ = the RNG-driven trajectory kernels, WIENER = the Wiener-driven ones (no noise: no stream version).
// The reference defines Milstein for M == 1 models only (milstein.jl:4): instantiate SCH = 1, 2 only there.
#define SDE_KERNELS_TERMINAL(MODEL, TAG, SCH, SNAME, T, V, SUF)                                                  \
  extern "C" __global__ void __launch_bounds__(SDE_SAMPLE_THREADS(sizeof(T) == 8 && V == 2), sizeof(T) == 8 ? (V == 2 ? SDE_SAMPLE_MINB_V2 : SDE_SAMPLE_MINB) : SDE_SAMPLE_MINB_F32) \
      sdek_##TAG##_sample_##SNAME##_##SUF(                                                                       \
      const __grid_constant__ sde_args a) { sdeb200::sample_body<MODEL, T, SCH, SDE_PPT_OF(T, V), V>(a); }      \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK * SDE_PPT, sizeof(T) == 8 ? SDE_SAMPLE1_MINB : 0)        \
      sdek_##TAG##_sample1_##SNAME##_##SUF(                                                                      \
      const __grid_constant__ sde_args a) { sdeb200::sample_body<MODEL, T, SCH, 1, V>(a); }                     \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, sizeof(T) == 8 ? SDE_F64_MINB : 0)                      \
      sdek_##TAG##_mlmc_##SNAME##_##SUF(                                                                         \
      const __grid_constant__ sde_args a) { sdeb200::mlmc_body<MODEL, T, SCH, V>(a); }
#define SDE_KERNELS_TRAJ(MODEL, TAG, SCH, SNAME, T, V, SUF)                                                      \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, sizeof(T) == 8 ? SDE_F64_MINB : 0)                      \
      sdek_##TAG##_simsoa_##SNAME##_##SUF(                                                                       \
      const __grid_constant__ sde_args a) { sdeb200::simulate_soa_body<MODEL, T, SCH, V>(a); }                  \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMREF_MINB) sdek_##TAG##_simref_##SNAME##_##SUF(    \
      const __grid_constant__ sde_args a) { sdeb200::simulate_ref_body<MODEL, T, SCH, V>(a); }                  \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMRT_MINB) sdek_##TAG##_simrt_##SNAME##_##SUF(      \
      const __grid_constant__ sde_args a) { sdeb200::simulate_ref_tma2_body<MODEL, T, SCH, V>(a); }
#define SDE_KERNELS_WIENER(MODEL, TAG, SCH, SNAME, T, SUF)                                                       \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMW_MINBLOCKS) sdek_##TAG##_simw_##SNAME##_##SUF(   \
      const __grid_constant__ sde_args a) { sdeb200::simulate_wiener_body<MODEL, T, SCH>(a); }                  \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK) sdek_##TAG##_simwt_##SNAME##_##SUF(                    \
      const __grid_constant__ sde_args a) { sdeb200::simulate_wiener_tma_body<MODEL, T, SCH>(a); }

#define SDE_KERNELS(MODEL, TAG, SCH, SNAME, T, PNAME)        \
  SDE_KERNELS_TERMINAL(MODEL, TAG, SCH, SNAME, T, 1, PNAME)  \
  SDE_KERNELS_TRAJ(MODEL, TAG, SCH, SNAME, T, 1, PNAME)      \
  SDE_KERNELS_WIENER(MODEL, TAG, SCH, SNAME, T, PNAME)
// The RNG-driven FP64 kernels again for normal stream v2 (sdeb200_opts.rng = 2); FP32 and the Wiener-driven kernels
// do not depend on the stream version.
#define SDE_KERNELS_V2(MODEL, TAG, SCH, SNAME)                     \
  SDE_KERNELS_TERMINAL(MODEL, TAG, SCH, SNAME, double, 2, f64v2)   \
  SDE_KERNELS_TRAJ(MODEL, TAG, SCH, SNAME, double, 2, f64v2)

// Exact one-step samplers exist for a few models only and have no Wiener-driven or coupled form
// (src/simulation.jl:50-57,78-89): terminal and full-trajectory kernels, scheme id 3.
#define SDE_KERNELS_EXACT(MODEL, TAG, T, PNAME)                                                                  \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK) sdek_##TAG##_sample_exact_##PNAME(                    \
      const __grid_constant__ sde_args a) { sdeb200::sample_body<MODEL, T, 3, SDE_PPT>(a); }                   \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK * SDE_PPT) sdek_##TAG##_sample1_exact_##PNAME(         \
      const __grid_constant__ sde_args a) { sdeb200::sample_body<MODEL, T, 3, 1>(a); }                          \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK) sdek_##TAG##_simsoa_exact_##PNAME(                    \
      const __grid_constant__ sde_args a) { sdeb200::simulate_soa_body<MODEL, T, 3>(a); }                      \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMREF_MINB) sdek_##TAG##_simref_exact_##PNAME(                 \
      const __grid_constant__ sde_args a) { sdeb200::simulate_ref_body<MODEL, T, 3>(a); }                      \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMRT_MINB) sdek_##TAG##_simrt_exact_##PNAME(                  \
      const __grid_constant__ sde_args a) { sdeb200::simulate_ref_tma2_body<MODEL, T, 3>(a); }

#define SDE_KERNELS_EXACT_V2(MODEL, TAG)                                                                          \
  extern "C" __global__ void __launch_bounds__(SDE_SAMPLE_THREADS(1)) sdek_##TAG##_sample_exact_f64v2(          \
      const __grid_constant__ sde_args a) { sdeb200::sample_body<MODEL, double, 3, SDE_PPT_V2, 2>(a); }         \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK * SDE_PPT) sdek_##TAG##_sample1_exact_f64v2(           \
      const __grid_constant__ sde_args a) { sdeb200::sample_body<MODEL, double, 3, 1, 2>(a); }                  \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK) sdek_##TAG##_simsoa_exact_f64v2(                      \
      const __grid_constant__ sde_args a) { sdeb200::simulate_soa_body<MODEL, double, 3, 2>(a); }               \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMREF_MINB) sdek_##TAG##_simref_exact_f64v2(        \
      const __grid_constant__ sde_args a) { sdeb200::simulate_ref_body<MODEL, double, 3, 2>(a); }               \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_SIMRT_MINB) sdek_##TAG##_simrt_exact_f64v2(          \
      const __grid_constant__ sde_args a) { sdeb200::simulate_ref_tma2_body<MODEL, double, 3, 2>(a); }

#define SDE_KERNELS_SCHEME(MODEL, TAG, SCH, SNAME)  \
  SDE_KERNELS(MODEL, TAG, SCH, SNAME, double, f64)  \
  SDE_KERNELS(MODEL, TAG, SCH, SNAME, float, f32)   \
  SDE_KERNELS_V2(MODEL, TAG, SCH, SNAME)

#define SDE_KERNELS_LOGT(MODEL, TAG, T, PNAME)                                                                   \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 4) sdek_##TAG##_logt_##PNAME(                          \
      const __grid_constant__ sde_args a) { sdeb200::logtransition_body<MODEL, T>(a); }                              \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 3) sdek_##TAG##_logtt_##PNAME(                         \
      const __grid_constant__ sde_args a) { sdeb200::logtransition_tma_body<MODEL, T>(a); }

#define SDE_KERNELS_COEF(MODEL, TAG)                                                                             \
  extern "C" __global__ void sdek_##TAG##_coef(const __grid_constant__ sde_args a) {                            \
    sdeb200::coef_eval_body<MODEL>(a);                                                                           \
  }

#endif  // __CUDACC__

}  // namespace sdeb200
