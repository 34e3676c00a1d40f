// sde_models_builtin.cuh — hand-written device functors for the two registry models the benchmark
// configurations name: BlackScholes (src/models/black_scholes.jl:1) and Heston (src/models/models.jl:44-48).
// They have exactly the interface the @sde_model code generator emits for arbitrary models (sde_dsl.cpp):
//
//   D, M, NP                         dim(model) and the number of struct fields (parameter order pinned by
//                                    test/runtests.jl:51-53)
//   prepare(p, dt, c)                hoists parameter-only sub-expressions once per thread
//   step<SCHEME>(c, t, dt, dw, x)    one scheme step in place: euler_maruyama.jl:1-6 / milstein.jl:4-10
//   coefficients(p, t, x, mu, sig)   drift / diffusion (Cholesky of the correlation folded in, codegen.jl:112)
//
// SDE_STRICT=1 (translation unit compiled with -fmad=false): the reference's literal operation order
//  — bit-identical to the CPU oracle.  SDE_STRICT=0: algebraically identical forms with
// loop invariants hoisted and FMAs, the throughput path.
#pragma once
#include "sde_device.cuh"

namespace sdeb200 {

SDE_HD double sde_sqrt(double v) { return sqrt(v); }
SDE_HD float sde_sqrt(float v) { return sqrtf(v); }
SDE_HD double sde_fma(double a, double b, double c) { return fma(a, b, c); }
SDE_HD float sde_fma(float a, float b, float c) { return fmaf(a, b, c); }
SDE_HD double sde_exp(double v) { return exp(v); }
SDE_HD float sde_exp(float v) { return expf(v); }
SDE_HD double sde_log(double v) { return log(v); }
SDE_HD float sde_log(float v) { return logf(v); }

// x + κΔt (θ - x), FAST math.  Float64: x (1 - κΔt) + κθΔt, one FMA (the rounding of 1 - κΔt, 2^-53, is harmless).  FP32 keeps
// the difference form: 1 - κΔt rounded to 24 bits would carry κΔt itself to only 2^-24 / (κΔt) relative — 6e-4 for κΔt =
// 1e-4 — and shift the level the state reverts to by that much, the same way at every step.
SDE_HD double mean_reversion(double x, double one_minus_kdt, double kthdt, double kdt, double theta) {
  (void)kdt;
  (void)theta;
  return fma(x, one_minus_kdt, kthdt);
}
SDE_HD float mean_reversion(float x, float one_minus_kdt, float kthdt, float kdt, float theta) {
  (void)one_minus_kdt;
  (void)kthdt;
  return fmaf(kdt, theta - x, x);
}

struct BlackScholes {  // dS = r*S*dt + σ*S*dW, params (r, σ)
  static constexpr int D = 1, M = 1, NP = 2;
  template <typename T> SDE_HD static void prepare(const T *p, T dt, T *c) {
    c[0] = p[0];
    c[1] = p[1];
    c[5] = (p[0] - (T)0.5 * (p[1] * p[1])) * dt;  // (r - 0.5σ²)Δt      Exact: black_scholes.jl:5
    c[6] = sde_sqrt(dt) * p[1];                   // d = sqrt(Δt)σ      Exact: black_scholes.jl:6
#if !SDE_STRICT
    c[2] = p[0] * dt;                 // r Δt
    c[3] = (T)0.5 * p[1] * p[1];      // σ²/2                                   (Milstein)
    const T sq = sde_sqrt(dt);
    c[4] = p[1] * (p[0] * dt + p[1] * sq) / ((T)2 * sq);  // σ (rΔt + σ√Δt) / (2√Δt)   (RungeKutta: (σpred-σ)/(2√Δt) per unit S)
#else
    (void)dt;
#endif
  }
  template <int SCHEME, typename T> SDE_HD static void step(const T *c, T t, T dt, const T *dw, T *x) {
    (void)t;
    const T S = x[0], w = dw[0];
    if (SCHEME == 3) {  // rand(LogNormal(log(x0) + (r-σ²/2)Δt, sqrt(Δt)σ)) = exp(randn()*d + μ); dw is a plain N(0,1) here
#if SDE_STRICT
      x[0] = sde_exp(w * c[6] + (sde_log(S) + c[5]));
#else
      x[0] = sde_exp(sde_fma(w, c[6], sde_log(S) + c[5]));
#endif
      return;
    }
#if SDE_STRICT
    const T mu = c[0] * S, sg = c[1] * S;
    T y = (S + mu * dt) + sg * w;
    if (SCHEME == 1) y = y + ((c[1] * sg) * (w * w - dt)) / (T)2;  // ∂σ = σ exactly (ForwardDiff dual of σ*x)
    if (SCHEME == 2) {                                              // runge_kutta.jl:7-9
      const T sq = sde_sqrt(dt);
      const T xh = (S + mu * dt) + sg * sq;
      const T sp = c[1] * xh;
      y = y + ((sp - sg) * (w * w - dt)) / ((T)2 * sq);
    }
    x[0] = y;
#else
    T a = sde_fma(c[1], w, c[2]);
    if (SCHEME == 1) a = sde_fma(c[3], sde_fma(w, w, -dt), a);
    if (SCHEME == 2) a = sde_fma(c[4], sde_fma(w, w, -dt), a);
    x[0] = sde_fma(S, a, S);
#endif
  }
  SDE_HD static void coefficients(const double *p, double t, const double *x, double *mu, double *sig) {
    (void)t;
    mu[0] = p[0] * x[0];
    sig[0] = p[1] * x[0];
  }
};

struct OrnsteinUhlenbeck {  // dx = θ*(μ-x)*dt + σ*dW, params (θ, μ, σ)   ornstein_uhlenbeck.jl:1
  static constexpr int D = 1, M = 1, NP = 3;
  template <typename T> SDE_HD static void prepare(const T *p, T dt, T *c) {
    c[0] = p[0];
    c[1] = p[1];
    c[2] = p[2];
    // Exact (ornstein_uhlenbeck.jl:3-8): Normal(x0 e^{-θΔt} + μ(1 - e^{-θΔt}), sqrt(1 - e^{-2θΔt}) σ / sqrt(2θ))
    c[3] = sde_exp(-p[0] * dt);
    c[4] = p[1] * ((T)1 - sde_exp(-p[0] * dt));
    c[5] = sde_sqrt((T)1 - sde_exp((T)-2 * p[0] * dt)) * p[2] / sde_sqrt((T)2 * p[0]);
#if !SDE_STRICT
    c[6] = (T)1 - p[0] * dt;   // 1 - θΔt
    c[7] = p[0] * p[1] * dt;   // θμΔt
    c[8] = p[0] * dt;          // θΔt (FP32: see mean_reversion)
#endif
  }
  template <int SCHEME, typename T> SDE_HD static void step(const T *c, T t, T dt, const T *dw, T *x) {
    (void)t;
    const T X = x[0], w = dw[0];
#if SDE_STRICT
    if (SCHEME == 3) {  // rand(Normal(m, d)) = m + d*randn(); dw is a plain N(0,1) here
      x[0] = (X * c[3] + c[4]) + c[5] * w;
      return;
    }
    const T mu = c[0] * (c[1] - X);
    // σ does not depend on x: the Milstein correction ∂σ*σ*(Δw²-Δt)/2 and the RungeKutta one (σpred-σ)(…) are exactly 0
    T y = (X + mu * dt) + c[2] * w;
    if (SCHEME == 1) y = y + (((T)0 * c[2]) * (w * w - dt)) / (T)2;
    if (SCHEME == 2) y = y + ((c[2] - c[2]) * (w * w - dt)) / ((T)2 * sde_sqrt(dt));
    x[0] = y;
#else
    // FAST forms are written with explicit FMAs only (like BlackScholes and Heston): the compiler has no contraction
    // left to choose, so every kernel form produces the same bits
    (void)dt;
    if (SCHEME == 3) x[0] = sde_fma(c[5], w, sde_fma(X, c[3], c[4]));
    else x[0] = sde_fma(c[2], w, mean_reversion(X, c[6], c[7], c[8], c[1]));   // the Milstein / RungeKutta corrections are exactly 0
#endif
  }
  SDE_HD static void coefficients(const double *p, double t, const double *x, double *mu, double *sig) {
    (void)t;
    mu[0] = p[0] * (p[1] - x[0]);
    sig[0] = p[2];
  }
};

struct Heston {  // dS = r*S*dt + sqrt(V)*S*dW1 ; dV = κ*(θ-V)*dt + σ*sqrt(V)*dW2 ; dW1*dW2 = ρ*dt ; params (r, κ, θ, σ, ρ)
  static constexpr int D = 2, M = 2, NP = 5;
  template <typename T> SDE_HD static void prepare(const T *p, T dt, T *c) {
#if SDE_STRICT
    (void)dt;
    for (int i = 0; i < 5; i++) c[i] = p[i];
    c[5] = sde_sqrt((T)1 - p[4] * p[4]);  // sqrt(1 - ρ^2): parameter-only, same value every step
#else
    c[0] = p[0] * dt;                      // r Δt
    c[1] = (T)1 - p[1] * dt;               // 1 - κ Δt
    c[2] = p[1] * p[2] * dt;               // κ θ Δt
    c[3] = p[3] * p[4];                    // σ ρ
    c[4] = p[3] * sde_sqrt((T)1 - p[4] * p[4]);  // σ sqrt(1 - ρ²)
    c[5] = p[1] * dt;                      // κ Δt   (FP32: see mean_reversion)
    c[6] = p[2];                           // θ
#endif
  }
  template <int SCHEME, typename T> SDE_HD static void step(const T *c, T t, T dt, const T *dw, T *x) {
    (void)t;
    static_assert(SCHEME == 0, "the reference defines Milstein / RungeKutta for M == 1 models only (milstein.jl:4)");
    const T S = x[0], V = x[1];
#if SDE_STRICT
    const T sv = sde_sqrt(V);
    const T mu1 = c[0] * S, mu2 = c[1] * (c[2] - V);
    const T s11 = sv * S, s21 = (c[3] * sv) * c[4], s22 = (c[3] * sv) * c[5];
    x[0] = (S + mu1 * dt) + s11 * dw[0];                  // + 0.0*dw[1] is exact for finite increments
    x[1] = (V + mu2 * dt) + (s21 * dw[0] + s22 * dw[1]);
#else
    (void)dt;
    const T sv = sqrt_fast(V);  // <= 1 ulp, branch-free; sqrt(0) = 0, sqrt(V < 0) = NaN like the reference's DomainError
    x[0] = sde_fma(S, sde_fma(sv, dw[0], c[0]), S);
    const T n = sde_fma(c[4], dw[1], c[3] * dw[0]);
    x[1] = sde_fma(sv, n, mean_reversion(V, c[1], c[2], c[5], c[6]));
#endif
  }
#if !SDE_STRICT
  // Polar form of the step (FAST math, FP64, either normal stream: sde_device.cuh, Noise<>): both noise terms carry
  // sqrt(V) sqrt(w), so ONE square root of V*w replaces the radius' and the model's (7 FP64 operations and one MUFU fewer
  // per step; sqrt(V w) is rounded once, sqrt(V) sqrt(w) three times).  V < 0 still gives a non-finite state.
  static constexpr int POLAR = 2;   // 1: step_polar; 2: step_polar2 as well
  // One step over the SUM of two increments given as polar triples a, b (the coarse step of a coupled level with two
  // substeps, multilevel.jl:79-86): sqrt(V)(dWa + dWb) = sqrt(V wa)(ca, sb)_a + sqrt(V wb)(ca, sb)_b — two roots instead
  // of the two radii plus sqrt(V), and the sum ΔW is never formed.  c = the constants of the coarse Δt.
  template <typename T> SDE_HD static void step_polar2(const T *c, T t, T dt, const T *a, const T *b, T *x) {
    (void)t;
    (void)dt;
    const T S = x[0], V = x[1];
    const T Ga = sqrt_fast(V * a[0]), Gb = sqrt_fast(V * b[0]);
    x[0] = sde_fma(S, sde_fma(Gb, b[1], sde_fma(Ga, a[1], c[0])), S);
    const T na = sde_fma(c[4], a[2], c[3] * a[1]), nb = sde_fma(c[4], b[2], c[3] * b[1]);
    x[1] = sde_fma(Gb, nb, sde_fma(Ga, na, mean_reversion(V, c[1], c[2], c[5], c[6])));
  }
  template <typename T> SDE_HD static void step_polar(const T *c, T t, T dt, T w, T ca, T sb, T *x) {
    (void)t;
    (void)dt;
    const T S = x[0], V = x[1];
    const T G = sqrt_fast(V * w);
    x[0] = sde_fma(S, sde_fma(G, ca, c[0]), S);
    const T n = sde_fma(c[4], sb, c[3] * ca);
    x[1] = sde_fma(G, n, mean_reversion(V, c[1], c[2], c[5], c[6]));
  }
#endif
  SDE_HD static void coefficients(const double *p, double t, const double *x, double *mu, double *sig) {
    (void)t;
    const double sv = sqrt(x[1]);
    mu[0] = p[0] * x[0];
    mu[1] = p[1] * (p[2] - x[1]);
    sig[0 * SDE_MAXM + 0] = sv * x[0];
    sig[0 * SDE_MAXM + 1] = 0.0;
    sig[1 * SDE_MAXM + 0] = (p[3] * sv) * p[4];
    sig[1 * SDE_MAXM + 1] = (p[3] * sv) * sqrt(1.0 - p[4] * p[4]);
  }
};

}  // namespace sdeb200
