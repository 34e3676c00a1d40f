// sde_aot.cu — ahead-of-time instantiation of every path kernel for the built-in functors.
// Compiled twice: -DSDE_STRICT=0 (FMA contraction, hoisted forms; the throughput path) and
// -DSDE_STRICT=1 -fmad=false (reference operation order; bit-identical to the CPU oracle).
#include "sde_models_builtin.cuh"
#include "sde_ktable.h"

using sdeb200::BlackScholes;
using sdeb200::Heston;
using sdeb200::OrnsteinUhlenbeck;

namespace sdeb200 {
// wiener!(w, Δt) (src/simulation.jl:22-35) is the in-place running sum w[i] = w[i-1] + sqrt(Δt)*z: the
// Euler–Maruyama recursion of dX_d = dW_d.  Only the reference-layout trajectory kernel is instantiated.
template <int DIM> struct WienerN {
  static constexpr int D = DIM, M = DIM, NP = 0;
  template <typename T> SDE_HD static void prepare(const T *, T, T *) {}
  // w[i] = w[i-1] + sqrt(Δt)*z with the product and the sum rounded separately, as the reference does: the add is
  // never contracted into the generator's last multiply, so every kernel form writes the same bits in FAST math too.
  SDE_D static double add_rn(double a, double b) { return __dadd_rn(a, b); }
  SDE_D static float add_rn(float a, float b) { return __fadd_rn(a, b); }
  template <int SCHEME, typename T> SDE_D static void step(const T *, T, T, const T *dw, T *x) {
#pragma unroll
    for (int d = 0; d < DIM; d++) x[d] = add_rn(x[d], dw[d]);
  }
};
}  // namespace sdeb200

#if SDE_STRICT
#define TAGGED(name) name##_s
#define SDE_AOT_TABLE sde_aot_table_strict
#define SDE_AOT_WIENER sde_aot_wiener_strict
#else
#define TAGGED(name) name##_f
#define SDE_AOT_TABLE sde_aot_table_fast
#define SDE_AOT_WIENER sde_aot_wiener_fast
#endif
#define XSCHEME(MODEL, TAG, SCH, SNAME) SDE_KERNELS_SCHEME(MODEL, TAG, SCH, SNAME)
#define XCOEF(MODEL, TAG) SDE_KERNELS_COEF(MODEL, TAG) SDE_KERNELS_LOGT(MODEL, TAG, double, f64) SDE_KERNELS_LOGT(MODEL, TAG, float, f32)
#define XFILL(T, TAG, s, SNAME) XFILL2(T, TAG, s, SNAME)
#define XFILL2(T, TAG, s, SNAME)                                                      \
  T.sample[s][0] = (const void *)sdek_##TAG##_sample_##SNAME##_f64;                    \
  T.sample[s][1] = (const void *)sdek_##TAG##_sample_##SNAME##_f32;                    \
  T.sample1[s][0] = (const void *)sdek_##TAG##_sample1_##SNAME##_f64;                  \
  T.sample1[s][1] = (const void *)sdek_##TAG##_sample1_##SNAME##_f32;                  \
  T.mlmc[s][0] = (const void *)sdek_##TAG##_mlmc_##SNAME##_f64;                        \
  T.mlmc[s][1] = (const void *)sdek_##TAG##_mlmc_##SNAME##_f32;                        \
  T.simsoa[s][0] = (const void *)sdek_##TAG##_simsoa_##SNAME##_f64;                    \
  T.simsoa[s][1] = (const void *)sdek_##TAG##_simsoa_##SNAME##_f32;                    \
  T.simref[s][0] = (const void *)sdek_##TAG##_simref_##SNAME##_f64;                    \
  T.simref[s][1] = (const void *)sdek_##TAG##_simref_##SNAME##_f32;                    \
  T.simrt[s][0] = (const void *)sdek_##TAG##_simrt_##SNAME##_f64;                      \
  T.simrt[s][1] = (const void *)sdek_##TAG##_simrt_##SNAME##_f32;                      \
  T.simw[s][0] = (const void *)sdek_##TAG##_simw_##SNAME##_f64;                        \
  T.simw[s][1] = (const void *)sdek_##TAG##_simw_##SNAME##_f32;                        \
  T.simwt[s][0] = (const void *)sdek_##TAG##_simwt_##SNAME##_f64;                      \
  T.simwt[s][1] = (const void *)sdek_##TAG##_simwt_##SNAME##_f32;                      \
  T.sample[s][2] = (const void *)sdek_##TAG##_sample_##SNAME##_f64v2;                  \
  T.sample1[s][2] = (const void *)sdek_##TAG##_sample1_##SNAME##_f64v2;                \
  T.mlmc[s][2] = (const void *)sdek_##TAG##_mlmc_##SNAME##_f64v2;                      \
  T.simsoa[s][2] = (const void *)sdek_##TAG##_simsoa_##SNAME##_f64v2;                  \
  T.simref[s][2] = (const void *)sdek_##TAG##_simref_##SNAME##_f64v2;                  \
  T.simrt[s][2] = (const void *)sdek_##TAG##_simrt_##SNAME##_f64v2;                    \
  T.simw[s][2] = T.simw[s][0];                                                         \
  T.simwt[s][2] = T.simwt[s][0];
#define XCOEFFILL(T, TAG) XCOEFFILL2(T, TAG)
#define XCOEFFILL2(T, TAG)                              \
  T.coef = (const void *)sdek_##TAG##_coef;             \
  T.logt[0] = (const void *)sdek_##TAG##_logt_f64;      \
  T.logt[1] = (const void *)sdek_##TAG##_logt_f32;      \
  T.logtt[0] = (const void *)sdek_##TAG##_logtt_f64;    \
  T.logtt[1] = (const void *)sdek_##TAG##_logtt_f32;

XSCHEME(BlackScholes, TAGGED(bs), 0, em)
XSCHEME(BlackScholes, TAGGED(bs), 1, mil)
XSCHEME(BlackScholes, TAGGED(bs), 2, rk)
XCOEF(BlackScholes, TAGGED(bs))
#define XEXACT(MODEL, TAG) XEXACT2(MODEL, TAG)
#define XEXACT2(MODEL, TAG) SDE_KERNELS_EXACT(MODEL, TAG, double, f64) SDE_KERNELS_EXACT(MODEL, TAG, float, f32) SDE_KERNELS_EXACT_V2(MODEL, TAG)
#define XEXACTFILL(T, TAG) XEXACTFILL2(T, TAG)
#define XEXACTFILL2(T, TAG)                                              \
  T.sample[3][0] = (const void *)sdek_##TAG##_sample_exact_f64;          \
  T.sample[3][1] = (const void *)sdek_##TAG##_sample_exact_f32;          \
  T.sample1[3][0] = (const void *)sdek_##TAG##_sample1_exact_f64;        \
  T.sample1[3][1] = (const void *)sdek_##TAG##_sample1_exact_f32;        \
  T.simsoa[3][0] = (const void *)sdek_##TAG##_simsoa_exact_f64;          \
  T.simsoa[3][1] = (const void *)sdek_##TAG##_simsoa_exact_f32;          \
  T.simref[3][0] = (const void *)sdek_##TAG##_simref_exact_f64;          \
  T.simref[3][1] = (const void *)sdek_##TAG##_simref_exact_f32;          \
  T.simrt[3][0] = (const void *)sdek_##TAG##_simrt_exact_f64;            \
  T.simrt[3][1] = (const void *)sdek_##TAG##_simrt_exact_f32;            \
  T.sample[3][2] = (const void *)sdek_##TAG##_sample_exact_f64v2;        \
  T.sample1[3][2] = (const void *)sdek_##TAG##_sample1_exact_f64v2;      \
  T.simsoa[3][2] = (const void *)sdek_##TAG##_simsoa_exact_f64v2;        \
  T.simref[3][2] = (const void *)sdek_##TAG##_simref_exact_f64v2;        \
  T.simrt[3][2] = (const void *)sdek_##TAG##_simrt_exact_f64v2;
XEXACT(BlackScholes, TAGGED(bs))
XSCHEME(OrnsteinUhlenbeck, TAGGED(ou), 0, em)
XSCHEME(OrnsteinUhlenbeck, TAGGED(ou), 1, mil)
XSCHEME(OrnsteinUhlenbeck, TAGGED(ou), 2, rk)
XEXACT(OrnsteinUhlenbeck, TAGGED(ou))
XCOEF(OrnsteinUhlenbeck, TAGGED(ou))
XSCHEME(Heston, TAGGED(heston), 0, em)
XCOEF(Heston, TAGGED(heston))

#define XWIENER(DIM, TAG) XWIENER2(DIM, TAG)
#define XWIENER2(DIM, TAG)                                                                                  \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 5) sdek_##TAG##_f64(const __grid_constant__ sde_args a) { \
    sdeb200::simulate_ref_body<sdeb200::WienerN<DIM>, double, 0>(a);                                         \
  }                                                                                                          \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 5) sdek_##TAG##_f32(const __grid_constant__ sde_args a) { \
    sdeb200::simulate_ref_body<sdeb200::WienerN<DIM>, float, 0>(a);                                          \
  }                                                                                                          \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 4) sdek_##TAG##t_f64(const __grid_constant__ sde_args a) { \
    sdeb200::simulate_ref_tma2_body<sdeb200::WienerN<DIM>, double, 0>(a);                                     \
  }                                                                                                          \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 4) sdek_##TAG##t_f32(const __grid_constant__ sde_args a) { \
    sdeb200::simulate_ref_tma2_body<sdeb200::WienerN<DIM>, float, 0>(a);                                      \
  }                                                                                                          \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 5) sdek_##TAG##_f64v2(const __grid_constant__ sde_args a) { \
    sdeb200::simulate_ref_body<sdeb200::WienerN<DIM>, double, 0, 2>(a);                                      \
  }                                                                                                          \
  extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 4) sdek_##TAG##t_f64v2(const __grid_constant__ sde_args a) { \
    sdeb200::simulate_ref_tma2_body<sdeb200::WienerN<DIM>, double, 0, 2>(a);                                  \
  }
XWIENER(1, TAGGED(wiener1))
XWIENER(2, TAGGED(wiener2))
XWIENER(3, TAGGED(wiener3))
XWIENER(4, TAGGED(wiener4))

// model: 0 = BlackScholes, 1 = Heston, 2 = OrnsteinUhlenbeck
extern "C" int SDE_AOT_TABLE(int model, sde_ktable *t) {
  sde_ktable z = {};
  *t = z;
  if (model == 0) {
    XFILL((*t), TAGGED(bs), 0, em)
    XFILL((*t), TAGGED(bs), 1, mil)
    XFILL((*t), TAGGED(bs), 2, rk)
    XEXACTFILL((*t), TAGGED(bs))
    XCOEFFILL((*t), TAGGED(bs))
    return 0;
  }
  if (model == 2) {
    XFILL((*t), TAGGED(ou), 0, em)
    XFILL((*t), TAGGED(ou), 1, mil)
    XFILL((*t), TAGGED(ou), 2, rk)
    XEXACTFILL((*t), TAGGED(ou))
    XCOEFFILL((*t), TAGGED(ou))
    return 0;
  }
  if (model == 1) {
    XFILL((*t), TAGGED(heston), 0, em)
    XCOEFFILL((*t), TAGGED(heston))
    return 0;
  }
  return -1;
}

// wiener kernels: dim 1..4 (+16: the TMA tensor-store form; dim 3 has none: 3*sizeof(T) does not divide 128),
// slot 0 = f64 stream v1, 1 = f32, 2 = f64 stream v2
#define XW(TAG, P) XW2(TAG, P)
#define XW2(TAG, P) (const void *)sdek_##TAG##_##P
#define XWT(TAG, P) XWT2(TAG, P)
#define XWT2(TAG, P) (const void *)sdek_##TAG##t_##P
extern "C" const void *SDE_AOT_WIENER(int dim, int slot) {
  if (slot < 0 || slot > 2) return 0;
  if (dim >= 16) {
    switch ((dim - 16) * 3 + slot) {
      case 3: return XWT(TAGGED(wiener1), f64);
      case 4: return XWT(TAGGED(wiener1), f32);
      case 5: return XWT(TAGGED(wiener1), f64v2);
      case 6: return XWT(TAGGED(wiener2), f64);
      case 7: return XWT(TAGGED(wiener2), f32);
      case 8: return XWT(TAGGED(wiener2), f64v2);
      case 12: return XWT(TAGGED(wiener4), f64);
      case 13: return XWT(TAGGED(wiener4), f32);
      case 14: return XWT(TAGGED(wiener4), f64v2);
    }
    return 0;
  }
  switch (dim * 3 + slot) {
    case 3: return XW(TAGGED(wiener1), f64);
    case 4: return XW(TAGGED(wiener1), f32);
    case 5: return XW(TAGGED(wiener1), f64v2);
    case 6: return XW(TAGGED(wiener2), f64);
    case 7: return XW(TAGGED(wiener2), f32);
    case 8: return XW(TAGGED(wiener2), f64v2);
    case 9: return XW(TAGGED(wiener3), f64);
    case 10: return XW(TAGGED(wiener3), f32);
    case 11: return XW(TAGGED(wiener3), f64v2);
    case 12: return XW(TAGGED(wiener4), f64);
    case 13: return XW(TAGGED(wiener4), f32);
    case 14: return XW(TAGGED(wiener4), f64v2);
  }
  return 0;
}
