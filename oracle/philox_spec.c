/*
 * philox_spec.c — independent CPU statement of the "sdeb200 normal stream v1" specification
 * (DESIGN.md §RNG).  TEST INFRASTRUCTURE ONLY (see sde_oracle.c header).
 *
 * This is NOT reference behaviour: the reference draws from Julia's process-global
 * Base.randn (src/simulation.jl:4-18) whose stream no test pins.  The device
 * kernels replace it by a counter-based generator; this file restates that generator with
 * libm in extended precision so that RNG-driven kernels can be checked path by path
 * (not only within 3 standard errors).
 *
 * Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
 * (SC'11); constants and known-answer vectors from the Random123 distribution (kat_vectors).
 */
#include <math.h>
#include <stdint.h>

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static const long double TWO_PI_L = 6.283185307179586476925286766559005768L;

/* FP64 stream: normal number q of path `path` comes from Philox block P = q/2,
 * counter (path_lo, path_hi, P, stream), key (seed_lo, seed_hi); q even -> cos, q odd -> sin. */
static void pair_f64(uint64_t seed, uint32_t stream, uint64_t path, uint32_t P, double z[2]) {
  uint32_t ctr[4] = {(uint32_t)path, (uint32_t)(path >> 32), P, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  oracle_philox4x32_10(ctr, key, r);
  uint64_t X1 = ((uint64_t)r[1] << 32) | r[0], X2 = ((uint64_t)r[3] << 32) | r[2];
  /* u1 = RN(X1) 2^-64 (round-to-nearest conversion, as cvt.rn.f64.u64); X1 == 0 reads as u1 = 2^-1087 */
  long double w;
  if (X1 == 0) {
    w = 2.0L * 1087.0L * 0.693147180559945309417232121458176568L;
  } else {
    double d1 = (double)X1;
    w = -2.0L * logl((long double)d1 * 0x1.0p-64L);
  }
  w += 0x1.0p-50L; /* spec: keeps w > 0 when u1 rounds to 1; the device folds it into its log table */
  uint64_t n = (X2 + (1ull << 61)) >> 62;
  int64_t f = (int64_t)(X2 - (n << 62));
  double t = (double)f; /* round-to-nearest, as cvt.rn.f64.s64 */
  long double phi = TWO_PI_L * ((long double)t * 0x1.0p-64L);
  long double c = cosl(phi), s = sinl(phi), cc, ss;
  switch (n & 3) {
    case 0: cc = c; ss = s; break;
    case 1: cc = -s; ss = c; break;
    case 2: cc = -c; ss = -s; break;
    default: cc = s; ss = -c; break;
  }
  long double rad = sqrtl(w);
  z[0] = (double)(rad * cc);
  z[1] = (double)(rad * ss);
}

void oracle_normals_f64(uint64_t seed, uint32_t stream, uint64_t path, uint64_t q0, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t q = q0 + (uint64_t)i;
    double z[2];
    pair_f64(seed, stream, path, (uint32_t)(q >> 1), z);
    out[i] = z[q & 1];
  }
}

/* FP32 stream: normal q from Philox block P = q/4; words (r0,r1) -> normals 0,1; (r2,r3) -> 2,3.
 * u1 = (r + 0.5) 2^-32, theta = 2 pi (int32)r' 2^-32.  Returned in double (the exact value
 * the float kernel approximates). */
static void quad_f32(uint64_t seed, uint32_t stream, uint64_t path, uint32_t P, double z[4]) {
  uint32_t ctr[4] = {(uint32_t)path, (uint32_t)(path >> 32), P, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  oracle_philox4x32_10(ctr, key, r);
  for (int h = 0; h < 2; h++) {
    long double u1 = ((long double)r[2 * h] + 0.5L) * 0x1.0p-32L;
    long double th = TWO_PI_L * ((long double)(int32_t)r[2 * h + 1] * 0x1.0p-32L);
    long double rad = sqrtl(-2.0L * logl(u1));
    z[2 * h] = (double)(rad * cosl(th));
    z[2 * h + 1] = (double)(rad * sinl(th));
  }
}

void oracle_normals_f32(uint64_t seed, uint32_t stream, uint64_t path, uint64_t q0, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t q = q0 + (uint64_t)i;
    double z[4];
    quad_f32(seed, stream, path, (uint32_t)(q >> 2), z);
    out[i] = z[q & 3];
  }
}

/* "sdeb200 normal stream v2" (FP64; opt-in, sdeb200_opts.rng = 2): the path's Philox output is a stream of 32-bit words,
 * word W = lane W % 4 of block W / 4 (counter (path_lo, path_hi, W / 4, stream)); pair P reads words 3P, 3P+1, 3P+2:
 * X1 = w[3P+1]:w[3P] gives u1 exactly as in v1, c = w[3P+2] is the angle in units of 2^-32 turns.  Normal q of the
 * path is element q % 2 of pair q / 2 (even -> cos, odd -> sin). */
static uint32_t stream_word(uint64_t seed, uint32_t stream, uint64_t path, uint64_t W) {
  uint32_t ctr[4] = {(uint32_t)path, (uint32_t)(path >> 32), (uint32_t)(W >> 2), stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  oracle_philox4x32_10(ctr, key, r);
  return r[W & 3];
}

static void pair_f64_v2(uint64_t seed, uint32_t stream, uint64_t path, uint64_t P, double z[2]) {
  const uint32_t a = stream_word(seed, stream, path, 3 * P), b = stream_word(seed, stream, path, 3 * P + 1),
                 c = stream_word(seed, stream, path, 3 * P + 2);
  const uint64_t X1 = ((uint64_t)b << 32) | a;
  long double w;
  if (X1 == 0) {
    w = 2.0L * 1087.0L * 0.693147180559945309417232121458176568L;
  } else {
    double d1 = (double)X1; /* round-to-nearest conversion, as cvt.rn.f64.u64 */
    w = -2.0L * logl((long double)d1 * 0x1.0p-64L);
  }
  w += 0x1.0p-50L;
  const long double phi = TWO_PI_L * ((long double)c * 0x1.0p-32L);
  const long double rad = sqrtl(w);
  z[0] = (double)(rad * cosl(phi));
  z[1] = (double)(rad * sinl(phi));
}

void oracle_normals_f64_v2(uint64_t seed, uint32_t stream, uint64_t path, uint64_t q0, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t q = q0 + (uint64_t)i;
    double z[2];
    pair_f64_v2(seed, stream, path, q >> 1, z);
    out[i] = z[q & 1];
  }
}
