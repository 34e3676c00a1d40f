// Compiles the math helpers of sde_device.cuh with g++ (no CUDA) so the Philox block function and the
// fused normal transforms can be checked on the CPU against oracle/philox_spec.c.
#include "../sdemodels.jl_b200/csrc/sde_models_builtin.cuh"
using namespace sdeb200;

extern "C" void h_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  uint32_t r[4];
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], r);
  for (int i = 0; i < 4; i++) out[i] = r[i];
}
extern "C" void h_normals_f64(uint64_t seed, uint32_t stream, uint64_t path, uint64_t q0, int64_t n, double scale2,
                              double *out) {
  for (int64_t i = 0; i < n; i++) {
    const uint64_t q = q0 + (uint64_t)i;
    uint32_t r[4];
    philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)(q >> 1), stream, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    double z[2], tab[SDE_TAB_DOUBLES], lk[6];
    h_scaled_tables(scale2, tab);
    make_log_consts(scale2, lk);
    normal_pair_f64(r, tab, lk, z[0], z[1]);
    out[i] = z[q & 1];
  }
}
extern "C" void h_normals_f32(uint64_t seed, uint32_t stream, uint64_t path, uint64_t q0, int64_t n, float scale2,
                              float *out) {
  for (int64_t i = 0; i < n; i++) {
    const uint64_t q = q0 + (uint64_t)i;
    uint32_t r[4];
    philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)(q >> 2), stream, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    float z[4];
    normals_from_block(r, (const double *)h_rotf_tab, nullptr, scale2, z);
    out[i] = z[q & 3];
  }
}
// raw transform on chosen bits (edge cases: tails, u1 -> 1, quadrant boundaries)
extern "C" void h_pair_from_bits(const uint32_t *r4, double scale2, double *z) {
  uint32_t r[4] = {r4[0], r4[1], r4[2], r4[3]};
  double tab[SDE_TAB_DOUBLES], lk[6];
  h_scaled_tables(scale2, tab);
  make_log_consts(scale2, lk);
  normal_pair_f64(r, tab, lk, z[0], z[1]);
}
// stream v2: pair P of a path from words 3P..3P+2 of its Philox word stream
static uint32_t h_word(uint64_t seed, uint32_t stream, uint64_t path, uint64_t W) {
  uint32_t r[4];
  philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)(W >> 2), stream, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  return r[W & 3];
}
extern "C" void h_normals_f64_v2(uint64_t seed, uint32_t stream, uint64_t path, uint64_t q0, int64_t n, double scale2,
                                 double *out) {
  static double tab[SDE_TAB2_DOUBLES];
  double lk[6];
  h_scaled_tables_v2(scale2, tab);
  make_log_consts_v2(scale2, lk);
  for (int64_t i = 0; i < n; i++) {
    const uint64_t q = q0 + (uint64_t)i, P = q >> 1;
    double z[2];
    normal_pair_f64_v2(h_word(seed, stream, path, 3 * P), h_word(seed, stream, path, 3 * P + 1), h_word(seed, stream, path, 3 * P + 2),
                       tab, lk, z[0], z[1]);
    out[i] = z[q & 1];
  }
}
extern "C" void h_pair_from_words_v2(const uint32_t *w3, double scale2, double *z) {
  static double tab[SDE_TAB2_DOUBLES];
  double lk[6];
  h_scaled_tables_v2(scale2, tab);
  make_log_consts_v2(scale2, lk);
  normal_pair_f64_v2(w3[0], w3[1], w3[2], tab, lk, z[0], z[1]);
}
extern "C" double h_sqrt6(double v) { return sqrt_from_seed6(v, rsqrt_seed(v)); }
extern "C" double h_sqrt7(double v) { return sqrt_core(v); }
// geometry of the TMA trajectory tiles (sde_args.h): paths per tensor-map row and the first aligned row of a sub-path
extern "C" int h_tma_group(int64_t pitch_bytes, int align) { return sde_tma_group(pitch_bytes, align); }
extern "C" int h_tma_head(int j, int64_t pitch_bytes, int rowb, int align) { return sde_tma_head(j, pitch_bytes, rowb, align); }
// the per-path form of the block function (round 0/1 invariants hoisted) must equal the plain one
extern "C" void h_philox_path(uint64_t path, uint32_t stream, uint32_t blk, const uint32_t *key, uint32_t *out) {
  PhiloxKeys ks;
  philox_expand(key[0], key[1], ks);
  PhiloxPath pp;
  philox_path_init(path, stream, ks, pp);
  uint32_t r[4];
  philox4x32_10(pp, blk, ks, r);
  for (int i = 0; i < 4; i++) out[i] = r[i];
}
// the log-density kernel's logarithm and reciprocal (K7, sde_device.cuh: log_fast / rcp_fast)
extern "C" void h_log_fast(const double *x, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) out[i] = log_fast(x[i], h_log_tab);
}
extern "C" void h_rcp_fast(const double *x, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) out[i] = rcp_fast(x[i]);
}
// the polar form of the Heston step (FAST math, FP64: one sqrt of V*w) beside the plain form on the pair sqrt(w) (ca, sb);
// version = normal stream (1: words r[0..3] of one Philox block, 2: words (3P, 3P+1, 3P+2))
extern "C" void h_heston_polar(int version, const double *params, double dt, const uint32_t *words, const double *x_in, double *x_polar,
                               double *x_plain) {
  static_assert(PolarOf<Heston>::value == 2 && PolarOf<BlackScholes>::value == 0, "only Heston declares the polar forms");
  static_assert(Noise<Heston, double, 2>::POLAR2 && Noise<Heston, float, 1>::POLAR2 && !Noise<BlackScholes, double, 2>::POLAR2, "");
  static_assert(Noise<Heston, double, 2>::POLAR && Noise<Heston, double, 1>::POLAR && Noise<Heston, float, 2>::POLAR, "both precisions, both streams");
  static_assert(Noise<Heston, double, 2>::NV == 12 && Noise<Heston, double, 1>::NV == 3 && Noise<Heston, float, 1>::NV == 6 &&
                Noise<BlackScholes, double, 2>::NV == 8 && !Noise<BlackScholes, float, 1>::POLAR, "values per batch");
  static double tab[SDE_TAB2_DOUBLES > SDE_TAB_DOUBLES ? SDE_TAB2_DOUBLES : SDE_TAB_DOUBLES];
  double lk[6], c[SDE_MAXC], w, ca, sb, z[2], dw[2];
  Heston::prepare(params, dt, c);
  if (version == 2) {
    h_scaled_tables_v2(dt, tab);
    make_log_consts_v2(dt, lk);
    normal_polar_f64_v2(words[0], words[1], words[2], tab, lk, w, ca, sb);
    normal_pair_f64_v2(words[0], words[1], words[2], tab, lk, z[0], z[1]);
  } else {
    const uint32_t r[4] = {words[0], words[1], words[2], words[3]};
    h_scaled_tables(dt, tab);
    make_log_consts(dt, lk);
    normal_polar_f64(r, tab, lk, w, ca, sb);
    normal_pair_f64(r, tab, lk, z[0], z[1]);
  }
  const double nz[3] = {w, ca, sb};
  if (version == 2) NoiseStep<Heston, 0, true, 2>::increments(nz, dw);
  else NoiseStep<Heston, 0, true, 1>::increments(nz, dw);
  if (dw[0] != z[0] || dw[1] != z[1]) { x_polar[0] = x_polar[1] = -1.0; return; }   // increments() must be the pair itself
  x_polar[0] = x_plain[0] = x_in[0];
  x_polar[1] = x_plain[1] = x_in[1];
  NoiseStep<Heston, 0, true, 1>::run(c, 0.0, dt, nz, x_polar);
  NoiseStep<Heston, 0, false, 1>::run(c, 0.0, dt, z, x_plain);
}
// the same in FP32 (host twin of the FP32 stream: table rotation, sqrtf): words (ra, rb) of one pair
extern "C" void h_heston_polar_f32(const float *params, float dt, const uint32_t *words, const float *x_in, float *x_polar, float *x_plain) {
  float c[SDE_MAXC], w, cs, sn, z[2], dw[2];
  Heston::prepare(params, dt, c);
  normal_polar_f32(words[0], words[1], dt, h_rotf_tab, w, cs, sn);
  normal_pair_f32(words[0], words[1], dt, h_rotf_tab, z[0], z[1]);
  const float nz[3] = {w, cs, sn};
  NoiseStep<Heston, 0, true, 1>::increments(nz, dw);
  if (dw[0] != z[0] || dw[1] != z[1]) { x_polar[0] = x_polar[1] = -1.0f; return; }
  x_polar[0] = x_plain[0] = x_in[0];
  x_polar[1] = x_plain[1] = x_in[1];
  NoiseStep<Heston, 0, true, 1>::run(c, 0.0f, dt, nz, x_polar);
  NoiseStep<Heston, 0, false, 1>::run(c, 0.0f, dt, z, x_plain);
}
// the coarse step of a coupled level with two substeps from the two polar triples (Heston::step_polar2) beside the reference
// form: increments summed from zero, then one plain step (multilevel.jl:79-86); stream v2 words of the two fine pairs
extern "C" void h_heston_polar2(const double *params, double dt_fine, const uint32_t *wa, const uint32_t *wb, const double *x_in,
                                double *x_p2, double *x_plain) {
  static double tab[SDE_TAB2_DOUBLES];
  double lk[6], c[SDE_MAXC], na[3], nb[3], za[2], zb[2];
  h_scaled_tables_v2(dt_fine, tab);
  make_log_consts_v2(dt_fine, lk);
  Heston::prepare(params, 2.0 * dt_fine, c);
  normal_polar_f64_v2(wa[0], wa[1], wa[2], tab, lk, na[0], na[1], na[2]);
  normal_polar_f64_v2(wb[0], wb[1], wb[2], tab, lk, nb[0], nb[1], nb[2]);
  NoiseStep<Heston, 0, true, 2>::increments(na, za);
  NoiseStep<Heston, 0, true, 2>::increments(nb, zb);
  const double DW[2] = {(0.0 + za[0]) + zb[0], (0.0 + za[1]) + zb[1]};
  x_p2[0] = x_plain[0] = x_in[0];
  x_p2[1] = x_plain[1] = x_in[1];
  NoiseStep<Heston, 0, true, 2>::run2(c, 0.0, 2.0 * dt_fine, na, nb, x_p2);
  NoiseStep<Heston, 0, false, 2>::run(c, 0.0, 2.0 * dt_fine, DW, x_plain);
}
// the variance row of the Heston FAST step without noise (dw = 0), FP32 and Float64: mean reversion only
extern "C" float h_heston_drift_f32(const float *params, float dt, float V) {
  float c[SDE_MAXC], x[2] = {100.0f, V}, dw[2] = {0.0f, 0.0f};
  Heston::prepare(params, dt, c);
  Heston::step<0>(c, 0.0f, dt, dw, x);
  return x[1];
}
extern "C" double h_heston_drift_f64(const double *params, double dt, double V) {
  double c[SDE_MAXC], x[2] = {100.0, V}, dw[2] = {0.0, 0.0};
  Heston::prepare(params, dt, c);
  Heston::step<0>(c, 0.0, dt, dw, x);
  return x[1];
}
