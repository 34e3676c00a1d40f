/*
 * sde_oracle.c — CPU restatement of the SDEModels.jl Monte Carlo path-simulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call it, and there only as the checker / the CPU baseline.  The product path
 * (sdemodels.jl_b200/) never links or imports this file.
 *
 * PARITY STATUS.  The reference is pure Julia and Julia is not installed in this image, so
 * the reference itself cannot be run.  This restatement is pinned against every numeric
 * known-answer the reference's own tests hold for this path (test/runtests.jl:51-58 dims
 * and parameter order, :65-71 drift values, :77-81 diffusion values incl. the Heston
 * Cholesky row, :89/:91 the noise-free trajectories, :104-121 the multilevel coupling
 * identities) — see tests/test_oracle_kat.py.  The reference holds NO golden stochastic
 * trajectory and no test pins its RNG (Base.randn) output, so trajectory parity beyond
 * those vectors is "parity unpinned" in the task's sense: the restatement follows the
 * source line by line (citations on every function) with Julia's evaluation order
 * (left-to-right, (a+b)+c, no FMA contraction: compile with -ffp-contract=off).
 *
 * Layout conventions are Julia's memory order: Array{SVector{D}}(nrows, npaths) column-major
 * == C array [npaths][nrows][D].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 4
#define MAXM 4

/* Arithmetic type of the restated loops.  The reference is Float64 only; the FP32 variants of the CUDA kernels
 * are a new-framework addition.  Compiling this file a second time with -DORACLE_REAL_FLOAT restates the SAME loops in
 * float (state, increments, parameters, Δt and times all rounded to float first, every operation a float operation,
 * -ffp-contract=off) and exports them as oracle_f32_*: the STRICT FP32 kernels must equal that bit for bit, which separates
 * kernel errors from the conditioning of FP32 arithmetic itself.  API: parameters / t0 / Δt / x0 arrive as double and are
 * rounded once, arrays are `real`. */
#ifdef ORACLE_REAL_FLOAT
typedef float real;
#define FN(name) oracle_f32_##name
#define R_SQRT sqrtf
#define R_EXP expf
#define R_LOG logf
#else
typedef double real;
#define FN(name) oracle_##name
#define R_SQRT sqrt
#define R_EXP exp
#define R_LOG log
#endif
#define MAXP 16

/* ------------------------------------------------------------------------------------------
 * Model functors: what `@sde_model` generates (src/models/codegen.jl:57-75,105-112) for the
 * registry models (src/models/models.jl:39-54, black_scholes.jl:1, cox_ingersoll_ross.jl:1,
 * ornstein_uhlenbeck.jl:1) and for the ad-hoc models of test/runtests.jl:17-41.
 * diffusion is D x M, stored sig[i*MAXM + j]; correlation Cholesky already folded in
 * (codegen.jl:112).  dsig[i*MAXD + k] = d sigma_i / d x_k for M == 1 models (milstein.jl:1-2).
 * ------------------------------------------------------------------------------------------ */
typedef void (*coef_fn)(const real *p, real t, const real *x, real *out);

/* Exact one-step sampler given the standard normal the distribution's rand() would draw
 * (Distributions 0.21: rand(LogNormal(μ,σ)) = exp(randn()*σ + μ), rand(Normal(μ,σ)) = μ + σ*randn()). */
typedef void (*exact_fn)(const real *p, real dt, const real *x0, real z, real *x1);

typedef struct {
  const char *name;
  int D, M, nparams;
  coef_fn drift, diffusion, ddiffusion;
  exact_fn exact;
} model_def;

/* BlackScholes: dS = r*S*dt + σ*S*dW            params (r, σ)   black_scholes.jl:1 */
static void bs_drift(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[0] * x[0]; }
static void bs_diff(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[1] * x[0]; }
static void bs_ddiff(const real *p, real t, const real *x, real *o) { (void)t; (void)x; o[0] = p[1]; }

/* sample(model::BlackScholes, scheme::Exact, t0, x0): black_scholes.jl:3-13 */
static void bs_exact(const real *p, real dt, const real *x0, real z, real *x1) {
  real mu = R_LOG(x0[0]) + (p[0] - (real)0.5 * (p[1] * p[1])) * dt;
  real d = R_SQRT(dt) * p[1];
  x1[0] = R_EXP(z * d + mu);
}

/* Heston: models.jl:44-48, params (r, κ, θ, σ, ρ); generated form README.md:42-59 / SURVEY §(real)3.5 */
static void heston_drift(const real *p, real t, const real *x, real *o) {
  (void)t;
  o[0] = p[0] * x[0];
  o[1] = p[1] * (p[2] - x[1]);
}
static void heston_diff(const real *p, real t, const real *x, real *o) {
  (void)t;
  real sv = R_SQRT(x[1]);
  o[0 * MAXM + 0] = sv * x[0];
  o[0 * MAXM + 1] = (real)0.0;
  o[1 * MAXM + 0] = (p[3] * sv) * p[4];
  o[1 * MAXM + 1] = (p[3] * sv) * R_SQRT((real)1.0 - p[4] * p[4]);
}

/* Deterministic: dX = X*dt (runtests.jl:24), D=1, M=0 */
static void det_drift(const real *p, real t, const real *x, real *o) { (void)p; (void)t; o[0] = x[0]; }
static void zero_diff(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0] = (real)0.0; }

/* Wiener: dX = dW (runtests.jl:23) */
static void wien_drift(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0] = (real)0.0; }
static void wien_diff(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0] = (real)1.0; }
static void wien_ddiff(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0] = (real)0.0; }

/* TimeDependent: dX = t*a*X*dt + b*X*dW (runtests.jl:25), params (a, b) */
static void td_drift(const real *p, real t, const real *x, real *o) { o[0] = (t * p[0]) * x[0]; }
static void td_diff(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[1] * x[0]; }
static void td_ddiff(const real *p, real t, const real *x, real *o) { (void)t; (void)x; o[0] = p[1]; }

/* TestModel (runtests.jl:17-21), params (a, c, e, b, d, f), D=3, M=4 */
static void tm_drift(const real *p, real t, const real *x, real *o) {
  (void)t;
  o[0] = x[1] * p[0];
  o[1] = x[2] * p[1];
  o[2] = x[0] * p[2];
}
static void tm_diff(const real *p, real t, const real *x, real *o) {
  (void)t; (void)x;
  for (int i = 0; i < 3 * MAXM; i++) o[i] = (real)0.0;
  o[0 * MAXM + 0] = p[3]; o[0 * MAXM + 1] = p[4];
  o[1 * MAXM + 1] = (real)1.0;  o[1 * MAXM + 2] = (real)1.0;
  o[2 * MAXM + 2] = p[5]; o[2 * MAXM + 3] = -(real)1.0;
}

/* SpecialCaseModel1-3 (runtests.jl:28-41) */
static void zero_drift3(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0] = o[1] = o[2] = (real)0.0; }
static void sc1_diff(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0] = (real)1.0; o[1] = (real)1.0; }
static void sc2_diff(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; o[0 * MAXM] = (real)1.0; o[1 * MAXM] = (real)1.0; }
static void sc2_ddiff(const real *p, real t, const real *x, real *o) { (void)p; (void)t; (void)x; for (int i = 0; i < MAXD * MAXD; i++) o[i] = (real)0.0; }
static void sc3_diff(const real *p, real t, const real *x, real *o) {
  (void)p; (void)t; (void)x;
  o[0 * MAXM + 0] = (real)1.0; o[0 * MAXM + 1] = (real)0.0;
  o[1 * MAXM + 0] = (real)1.0; o[1 * MAXM + 1] = (real)1.0;
  o[2 * MAXM + 0] = (real)0.0; o[2 * MAXM + 1] = (real)1.0;
}

/* CoxIngersollRoss: dr = κ*(θ-r)*dt + σ*R_SQRT(r)*dW (cox_ingersoll_ross.jl:1), params (κ, θ, σ) */
static void cir_drift(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[0] * (p[1] - x[0]); }
static void cir_diff(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[2] * R_SQRT(x[0]); }
static void cir_ddiff(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[2] * ((real)0.5 / R_SQRT(x[0])); }

/* OrnsteinUhlenbeck: dx = θ*(μ-x)*dt + σ*dW (ornstein_uhlenbeck.jl:1), params (θ, μ, σ) */
static void ou_drift(const real *p, real t, const real *x, real *o) { (void)t; o[0] = p[0] * (p[1] - x[0]); }
static void ou_diff(const real *p, real t, const real *x, real *o) { (void)t; (void)x; o[0] = p[2]; }

/* sample(model::OrnsteinUhlenbeck, scheme::Exact, t0, x0): ornstein_uhlenbeck.jl:3-13 */
static void ou_exact(const real *p, real dt, const real *x0, real z, real *x1) {
  real m = x0[0] * R_EXP(-p[0] * dt) + p[1] * ((real)1.0 - R_EXP(-p[0] * dt));
  real d = R_SQRT((real)1.0 - R_EXP(-(real)2.0 * p[0] * dt)) * p[2] / R_SQRT((real)2.0 * p[0]);
  x1[0] = m + d * z;
}

/* FitzHughNagumo (models.jl:39-42), params (ɛ, s, γ, β, σ1, σ2), D=2, M=2 */
static void fhn_drift(const real *p, real t, const real *x, real *o) {
  (void)t;
  o[0] = p[0] * (((x[0] - x[0] * x[0] * x[0]) - x[1]) + p[1]);
  o[1] = ((p[2] * x[0]) - x[1]) + p[3];
}
static void fhn_diff(const real *p, real t, const real *x, real *o) {
  (void)t; (void)x;
  o[0 * MAXM + 0] = p[4]; o[0 * MAXM + 1] = (real)0.0;
  o[1 * MAXM + 0] = (real)0.0;  o[1 * MAXM + 1] = p[5];
}

/* Lorenz (models.jl:50-54), params (σ, ρ, β, σ1, σ2, σ3), D=3, M=3 */
static void lor_drift(const real *p, real t, const real *x, real *o) {
  (void)t;
  o[0] = p[0] * (x[1] - x[0]);
  o[1] = x[0] * (p[1] - x[2]) - x[1];
  o[2] = x[0] * x[1] - p[2] * x[2];
}
static void lor_diff(const real *p, real t, const real *x, real *o) {
  (void)t; (void)x;
  for (int i = 0; i < 3 * MAXM; i++) o[i] = (real)0.0;
  o[0 * MAXM + 0] = p[3]; o[1 * MAXM + 1] = p[4]; o[2 * MAXM + 2] = p[5];
}

enum {
  M_BLACKSCHOLES = 0, M_HESTON, M_DETERMINISTIC, M_WIENER, M_TIMEDEPENDENT, M_TESTMODEL,
  M_SPECIAL1, M_SPECIAL2, M_SPECIAL3, M_CIR, M_OU, M_FHN, M_LORENZ, M_COUNT
};

static const model_def MODELS[M_COUNT] = {
  {"BlackScholes", 1, 1, 2, bs_drift, bs_diff, bs_ddiff, bs_exact},
  {"Heston", 2, 2, 5, heston_drift, heston_diff, 0, 0},
  {"Deterministic", 1, 0, 0, det_drift, zero_diff, 0, 0},
  {"Wiener", 1, 1, 0, wien_drift, wien_diff, wien_ddiff, 0},
  {"TimeDependent", 1, 1, 2, td_drift, td_diff, td_ddiff, 0},
  {"TestModel", 3, 4, 6, tm_drift, tm_diff, 0, 0},
  {"SpecialCaseModel1", 1, 2, 0, wien_drift, sc1_diff, 0, 0},
  {"SpecialCaseModel2", 2, 1, 0, zero_drift3, sc2_diff, sc2_ddiff, 0},
  {"SpecialCaseModel3", 3, 2, 0, zero_drift3, sc3_diff, 0, 0},
  {"CoxIngersollRoss", 1, 1, 3, cir_drift, cir_diff, cir_ddiff, 0},
  {"OrnsteinUhlenbeck", 1, 1, 3, ou_drift, ou_diff, wien_ddiff, ou_exact},
  {"FitzHughNagumo", 2, 2, 6, fhn_drift, fhn_diff, 0, 0},
  {"Lorenz", 3, 3, 6, lor_drift, lor_diff, 0, 0},
};

#ifndef ORACLE_REAL_FLOAT
int oracle_model_count(void) { return M_COUNT; }
const char *oracle_model_name(int id) { return (id >= 0 && id < M_COUNT) ? MODELS[id].name : 0; }
int oracle_model_dims(int id, int *D, int *M, int *np) {
  if (id < 0 || id >= M_COUNT) return -1;
  *D = MODELS[id].D; *M = MODELS[id].M; *np = MODELS[id].nparams;
  return 0;
}
/* dense (unpadded) outputs: mu[D]; sig[D*M] row-major */
int oracle_drift(int id, const real *p, real t, const real *x, real *mu) {
  if (id < 0 || id >= M_COUNT) return -1;
  real o[MAXD];
  MODELS[id].drift(p, t, x, o);
  for (int i = 0; i < MODELS[id].D; i++) mu[i] = o[i];
  return 0;
}
int oracle_diffusion(int id, const real *p, real t, const real *x, real *sig) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  real o[MAXD * MAXM];
  memset(o, 0, sizeof o);
  m->diffusion(p, t, x, o);
  if (m->M == 0) { sig[0] = (real)0.0; return 0; } /* `diffusion = 0` for M == 0, codegen.jl:112 */
  for (int i = 0; i < m->D; i++)
    for (int j = 0; j < m->M; j++) sig[i * m->M + j] = o[i * MAXM + j];
  return 0;
}

#endif /* !ORACLE_REAL_FLOAT */

/* API boundary: double arguments are rounded to `real` once (identity in the Float64 build), as the kernels do with
 * their argument block ((T)a.params[i], (T)a.dt, (T)a.t0, (T)a.x0[d]: sde_device.cuh). */
#define CONV_ARGS(m)                                                           \
  real p[MAXP], x0[MAXD];                                                      \
  for (int i_ = 0; i_ < MAXP; i_++) p[i_] = i_ < (m)->nparams ? (real)p_in[i_] : (real)0; \
  for (int i_ = 0; i_ < (m)->D; i_++) x0[i_] = (real)x0_in[i_];

/* ------------------------------------------------------------------------------------------
 * step: src/schemes/euler_maruyama.jl:1-6 and src/schemes/milstein.jl:4-10.
 * Evaluation order per SURVEY Appendix A: x0 + μ*Δt + σ*Δw == (x0 + μ*Δt) + (σ*Δw),
 * static mat-vec row i = σ[i,1]*Δw[1] + σ[i,2]*Δw[2] + ... left to right.
 * scheme: 0 = EulerMaruyama, 1 = Milstein, 2 = RungeKutta (src/schemes/runge_kutta.jl:4-11, the derivative-free
 * Milstein); 1 and 2 for M == 1 only, as in the reference.
 * ------------------------------------------------------------------------------------------ */
static int step(const model_def *m, int scheme, const real *p, real dt, real t, const real *x0,
                const real *dw, real *x1) {
  if (scheme == 3) { /* Exact: `dw` carries the plain standard normal (src/simulation.jl:50-57 calls sample(model, Exact, t, x)) */
    if (!m->exact) return -2;
    m->exact(p, dt, x0, dw[0], x1);
    return 0;
  }
  real mu[MAXD], sig[MAXD * MAXM];
  memset(sig, 0, sizeof sig);
  m->drift(p, t, x0, mu);
  m->diffusion(p, t, x0, sig);
  if (scheme == 0) {
    for (int i = 0; i < m->D; i++) {
      real a = x0[i] + mu[i] * dt;
      if (m->M > 0) {
        real s = sig[i * MAXM] * dw[0];
        for (int j = 1; j < m->M; j++) s = s + sig[i * MAXM + j] * dw[j];
        a = a + s;
      } else {
        a = a + (real)0.0 * dw[0]; /* σ == (real)0.0, Δw == (real)0.0 for AbstractSDE{1,0} (simulation.jl:6) */
      }
      x1[i] = a;
    }
    return 0;
  }
  if (scheme == 1) {
    if (m->M != 1 || !m->ddiffusion) return -2; /* no Milstein method for M != 1 (milstein.jl:4) */
    real ds[MAXD * MAXD];
    m->ddiffusion(p, t, x0, ds);
    real c = dw[0] * dw[0] - dt; /* Δw.^2 - Δt */
    for (int i = 0; i < m->D; i++) {
      real js; /* (∂σ*σ)[i] */
      if (m->D == 1) js = ds[0] * sig[0];
      else {
        js = ds[i * MAXD] * sig[0];
        for (int k = 1; k < m->D; k++) js = js + ds[i * MAXD + k] * sig[k * MAXM];
      }
      x1[i] = ((x0[i] + mu[i] * dt) + sig[i * MAXM] * dw[0]) + (js * c) / (real)2.0;
    }
    return 0;
  }
  if (scheme == 2) {
    /* x0hat = x0 + μ*Δt + vec(σ)*R_SQRT(Δt); σpred = diffusion(model, t0, x0hat);
     * x = x0 + μ*Δt + σ*Δw + (σpred - σ)*(Δw.^2 - Δt)/(2*R_SQRT(Δt))          runge_kutta.jl:7-9 */
    if (m->M != 1) return -2;
    real xh[MAXD], sp[MAXD * MAXM];
    const real sq = R_SQRT(dt);
    memset(sp, 0, sizeof sp);
    for (int i = 0; i < m->D; i++) xh[i] = (x0[i] + mu[i] * dt) + sig[i * MAXM] * sq;
    m->diffusion(p, t, xh, sp);
    const real c = dw[0] * dw[0] - dt, den = (real)2.0 * sq;
    for (int i = 0; i < m->D; i++)
      x1[i] = ((x0[i] + mu[i] * dt) + sig[i * MAXM] * dw[0]) + ((sp[i * MAXM] - sig[i * MAXM]) * c) / den;
    return 0;
  }
  return -3;
}

int FN(step)(int id, int scheme, const double *p_in, double dt, double t, const double *x0_in, const real *dw, real *x1) {
  if (id < 0 || id >= M_COUNT) return -1;
  CONV_ARGS(&MODELS[id])
  return step(&MODELS[id], scheme, p, (real)dt, (real)t, x0, dw, x1);
}

/* ------------------------------------------------------------------------------------------
 * sample with supplied standard normals: src/simulation.jl:39-48 (one path) and :59-67.
 * z[npaths][nsteps][max(M,1)] are the values `_randn(T)` would have returned (simulation.jl:4-11),
 * in the reference's draw order (path outer, step inner, M draws left to right).
 * out[npaths][D].
 * ------------------------------------------------------------------------------------------ */
int FN(sample_z)(int id, int scheme, const double *p_in, double t0_in, double dt_in, const double *x0_in, int64_t nsteps,
                 int64_t npaths, const real *z, real *out) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  CONV_ARGS(m)
  const real t0 = (real)t0_in, dt = (real)dt_in;
  const int M = m->M > 0 ? m->M : 1;
  const real sdt = R_SQRT(dt);
  int rc = 0;
  for (int64_t k = 0; k < npaths; k++) {
    real x[MAXD], xn[MAXD], dw[MAXM];
    for (int i = 0; i < m->D; i++) x[i] = x0[i];
    for (int64_t i = 1; i <= nsteps; i++) {
      real ti = t0 + (real)(i - 1) * dt;
      for (int j = 0; j < M; j++) dw[j] = (scheme == 3 ? (real)1.0 : sdt) * (m->M > 0 ? z[(k * nsteps + (i - 1)) * M + j] : (real)0.0);
      rc |= step(m, scheme, p, dt, ti, x, dw, xn);
      for (int d = 0; d < m->D; d++) x[d] = xn[d];
    }
    for (int d = 0; d < m->D; d++) out[k * m->D + d] = x[d];
  }
  return rc;
}

/* wiener!: src/simulation.jl:22-35.  w[npaths][nrows][dim]; z[npaths][nrows-1][dim]. */
void FN(wiener_z)(double dt_in, int64_t nrows, int64_t npaths, int dim, const real *z, real *w) {
  const real dt = (real)dt_in;
  const real s = R_SQRT(dt);
  for (int64_t k = 0; k < npaths; k++) {
    real sum[MAXM > MAXD ? MAXM : MAXD];
    for (int j = 0; j < dim; j++) { sum[j] = (real)0.0; w[(k * nrows) * dim + j] = (real)0.0; }
    for (int64_t i = 1; i < nrows; i++)
      for (int j = 0; j < dim; j++) {
        sum[j] = sum[j] + s * z[(k * (nrows - 1) + (i - 1)) * dim + j];
        w[(k * nrows + i) * dim + j] = sum[j];
      }
  }
}

/* simulate! driven by standard normals: src/simulation.jl:104-117.  x[npaths][nrows][D]. */
int FN(simulate_z)(int id, int scheme, const double *p_in, double t0_in, double dt_in, const double *x0_in, int64_t nrows,
                   int64_t npaths, const real *z, real *x) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  CONV_ARGS(m)
  const real t0 = (real)t0_in, dt = (real)dt_in;
  const int M = m->M > 0 ? m->M : 1, D = m->D;
  const real sdt = R_SQRT(dt);
  int rc = 0;
  for (int64_t k = 0; k < npaths; k++) {
    real *xk = x + k * nrows * D;
    for (int d = 0; d < D; d++) xk[d] = x0[d];
    for (int64_t i = 2; i <= nrows; i++) {
      real t = t0 + (real)(i - 2) * dt, dw[MAXM];
      for (int j = 0; j < M; j++) dw[j] = (scheme == 3 ? (real)1.0 : sdt) * (m->M > 0 ? z[(k * (nrows - 1) + (i - 2)) * M + j] : (real)0.0);
      rc |= step(m, scheme, p, dt, t, xk + (i - 2) * D, dw, xk + (i - 1) * D);
    }
  }
  return rc;
}

/* simulate! driven by a cumulative Wiener path, x === w allowed: src/simulation.jl:119-134.
 * w[npaths][nrows][M], x[npaths][nrows][D] (D == M when aliased). */
int FN(simulate_wiener)(int id, int scheme, const double *p_in, double t0_in, double dt_in, const double *x0_in,
                        int64_t nrows, int64_t npaths, const real *w, real *x) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  CONV_ARGS(m)
  const real t0 = (real)t0_in, dt = (real)dt_in;
  const int M = m->M > 0 ? m->M : 1, D = m->D;
  int rc = 0;
  for (int64_t k = 0; k < npaths; k++) {
    const real *wk = w + k * nrows * M;
    real *xk = x + k * nrows * D;
    real wprev[MAXM], wnext[MAXM], dw[MAXM], xprev[MAXD], xn[MAXD];
    for (int j = 0; j < M; j++) wprev[j] = wk[j];
    for (int d = 0; d < D; d++) { xk[d] = x0[d]; xprev[d] = x0[d]; }
    for (int64_t i = 2; i <= nrows; i++) {
      real t = t0 + (real)(i - 2) * dt;
      for (int j = 0; j < M; j++) { wnext[j] = wk[(i - 1) * M + j]; dw[j] = wnext[j] - wprev[j]; }
      rc |= step(m, scheme, p, dt, t, xprev, dw, xn);
      for (int d = 0; d < D; d++) { xk[(i - 1) * D + d] = xn[d]; xprev[d] = xn[d]; }
      for (int j = 0; j < M; j++) wprev[j] = wnext[j];
    }
  }
  return rc;
}

#ifndef ORACLE_REAL_FLOAT
/* ------------------------------------------------------------------------------------------
 * logtransition(model, ::EulerMaruyama, times, data): src/schemes/schemes.jl:64-76 with
 * _normal_transition_params of euler_maruyama.jl:8-13 — Gaussian one-step log-densities along
 * stored trajectories.  x[npaths][nrows][D] -> out[npaths][nrows-1]; times t0 + i*dt.
 *   μ = x0 + drift*Δt;  Σ = Δt*σ*σ';  z = μ - x1;  -0.5*(D*log(2π) + log(det(Σ)) + dot(z, Σ\z))
 * D == 1 and D == 2 use the closed forms StaticArrays uses (det = a11 a22 - a12 a21, Cramer solve);
 * larger D go through a Cholesky factorisation (equal up to rounding).
 * ------------------------------------------------------------------------------------------ */
int oracle_logtransition(int id, const double *p, double t0, double dt, int64_t nrows, int64_t npaths,
                         const double *x, double *out) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  const int D = m->D, M = m->M;
  const double log2pi = log(2.0 * 3.14159265358979323846);
  for (int64_t k = 0; k < npaths; k++)
    for (int64_t i = 1; i < nrows; i++) {
      const double *xa = x + (k * nrows + i - 1) * D, *xb = xa + D;
      const double t = t0 + (double)(i - 1) * dt;
      double mu[MAXD], sig[MAXD * MAXM], S[MAXD][MAXD], z[MAXD];
      memset(sig, 0, sizeof sig);
      m->drift(p, t, xa, mu);
      m->diffusion(p, t, xa, sig);
      for (int a = 0; a < D; a++) z[a] = (xa[a] + mu[a] * dt) - xb[a];
      for (int a = 0; a < D; a++)
        for (int b = 0; b < D; b++) {
          double acc = 0.0;
          for (int j = 0; j < M; j++) acc = (j == 0) ? (dt * sig[a * MAXM + j]) * sig[b * MAXM + j]
                                                     : acc + (dt * sig[a * MAXM + j]) * sig[b * MAXM + j];
          S[a][b] = acc;
        }
      double logdet, q;
      if (D == 1) {
        logdet = log(S[0][0]);
        q = z[0] * (z[0] / S[0][0]);
      } else if (D == 2) {
        double det = S[0][0] * S[1][1] - S[0][1] * S[1][0];
        double y0 = (S[1][1] * z[0] - S[0][1] * z[1]) / det, y1 = (S[0][0] * z[1] - S[1][0] * z[0]) / det;
        logdet = log(det);
        q = z[0] * y0 + z[1] * y1;
      } else {
        double L[MAXD][MAXD];
        double ld = 0.0;
        int ok = 1;
        memset(L, 0, sizeof L);
        for (int a = 0; a < D && ok; a++)
          for (int b = 0; b <= a; b++) {
            double s2 = S[a][b];
            for (int c = 0; c < b; c++) s2 -= L[a][c] * L[b][c];
            if (a == b) { if (!(s2 > 0.0)) { ok = 0; break; } L[a][a] = sqrt(s2); ld += 2.0 * log(L[a][a]); }
            else L[a][b] = s2 / L[b][b];
          }
        if (!ok) { out[k * (nrows - 1) + i - 1] = NAN; continue; }
        double y[MAXD];
        q = 0.0;
        for (int a = 0; a < D; a++) {
          double s2 = z[a];
          for (int c = 0; c < a; c++) s2 -= L[a][c] * y[c];
          y[a] = s2 / L[a][a];
          q += y[a] * y[a];
        }
        logdet = ld;
      }
      out[k * (nrows - 1) + i - 1] = -0.5 * (((double)D * log2pi + logdet) + q);
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Multilevel: src/multilevel.jl
 * ------------------------------------------------------------------------------------------ */
static int64_t ipow(int64_t b, int64_t e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

/* subdivide(scheme, substeps^l).Δt : schemes.jl:53-54 with multilevel.jl:7-10 — a single division. */
double oracle_level_dt(double base_dt, int64_t substeps, int64_t level) { return base_dt / (double)ipow(substeps, level); }

/* npaths(s::MultilevelScheme, max_budget): multilevel.jl:12-24.  levels = level_lo:level_hi. */
void oracle_npaths(int64_t substeps, int64_t level_lo, int64_t level_hi, double max_budget, int64_t *n_total) {
  const int nl = (int)(level_hi - level_lo + 1);
  int64_t steps[64];
  double budget = max_budget;
  for (int i = 0; i < nl; i++) { steps[i] = ipow(substeps, level_lo + i); n_total[i] = 0; }
  for (int i = nl; i >= 1; i--) {
    int64_t level_budget = INT64_MAX;
    for (int j = 0; j < i; j++) {
      int64_t v = (int64_t)floor(budget / (double)i / (double)steps[j]) * steps[j];
      if (v < level_budget) level_budget = v;
    }
    int64_t dot = 0;
    for (int j = 0; j < i; j++) {
      int64_t n_partial = level_budget / steps[j];
      n_total[j] += n_partial;
      dot += n_partial * steps[j];
    }
    budget -= (double)dot;
  }
}

/* cost(s, npaths): multilevel.jl:26-29 */
void oracle_cost(int64_t substeps, int64_t level_lo, int64_t level_hi, const int64_t *npaths, int64_t *cost) {
  for (int64_t l = level_lo; l <= level_hi; l++) cost[l - level_lo] = ipow(substeps, l) * npaths[l - level_lo];
}

#endif /* !ORACLE_REAL_FLOAT */

/* _multilevel_sample, one coupled level pair: multilevel.jl:65-89.
 * z[npaths][nsteps*substeps][M] fine-step normals; nsteps = number of COARSE steps.
 * xc[npaths][D], xf[npaths][D].  Requires D == M (or scalar), as the reference does (Δw = zero(T2)). */
int FN(multilevel_sample_z)(int id, int scheme, const double *p_in, double t0_in, double dtc_in, double dtf_in,
                            int64_t substeps, const double *x0_in, int64_t nsteps, int64_t npaths,
                            const real *z, real *xc_out, real *xf_out) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  CONV_ARGS(m)
  const real t0 = (real)t0_in, dt_coarse = (real)dtc_in, dt_fine = (real)dtf_in;
  const int M = m->M > 0 ? m->M : 1, D = m->D;
  const real s = R_SQRT(dt_fine);
  int rc = 0;
  for (int64_t k = 0; k < npaths; k++) {
    real xc[MAXD], xf[MAXD], xn[MAXD], DW[MAXM], dw[MAXM];
    for (int d = 0; d < D; d++) { xc[d] = x0[d]; xf[d] = x0[d]; }
    const real *zk = z + k * nsteps * substeps * M;
    for (int64_t n = 1; n <= nsteps; n++) {
      for (int j = 0; j < M; j++) DW[j] = (real)0.0;
      real tc = t0 + (real)(n - 1) * dt_coarse;
      for (int64_t i = 1; i <= substeps; i++) {
        real tf = tc + (real)(i - 1) * dt_fine;
        for (int j = 0; j < M; j++) {
          dw[j] = s * (m->M > 0 ? zk[((n - 1) * substeps + (i - 1)) * M + j] : (real)0.0);
          DW[j] = DW[j] + dw[j];
        }
        rc |= step(m, scheme, p, dt_fine, tf, xf, dw, xn);
        for (int d = 0; d < D; d++) xf[d] = xn[d];
      }
      rc |= step(m, scheme, p, dt_coarse, tc, xc, DW, xn);
      for (int d = 0; d < D; d++) xc[d] = xn[d];
    }
    for (int d = 0; d < D; d++) { xc_out[k * D + d] = xc[d]; xf_out[k * D + d] = xf[d]; }
  }
  return rc;
}

/* _multilevel_simulate: multilevel.jl:91-100.  wiener!(xf) draws eltype(xf)==D normals per row
 * (simulation.jl:12-18,30), xc = xf[1:substeps:end,:] then both integrated in place.
 * z[npaths][nsteps*substeps][D]; xf[npaths][nsteps*substeps+1][D]; xc[npaths][nsteps+1][D]. */
int FN(multilevel_simulate_z)(int id, int scheme, const double *p, double t0, double dt_coarse, double dt_fine,
                              int64_t substeps, const double *x0, int64_t nsteps, int64_t npaths,
                              const real *z, real *xc, real *xf) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  const int D = m->D;
  if (m->M != D) return -4;
  const int64_t nf = nsteps * substeps + 1, nc = nsteps + 1;
  FN(wiener_z)(dt_fine, nf, npaths, D, z, xf);
  for (int64_t k = 0; k < npaths; k++)
    for (int64_t i = 0; i < nc; i++)
      for (int d = 0; d < D; d++) xc[(k * nc + i) * D + d] = xf[(k * nf + i * substeps) * D + d];
  int rc = FN(simulate_wiener)(id, scheme, p, t0, dt_fine, x0, nf, npaths, xf, xf);
  rc |= FN(simulate_wiener)(id, scheme, p, t0, dt_coarse, x0, nc, npaths, xc, xc);
  return rc;
}

#ifndef ORACLE_REAL_FLOAT
/* ------------------------------------------------------------------------------------------
 * CPU baseline: the reference's execution model (scalar per path, simulation.jl:59-67 calling
 * :39-48) with a generator in the same cost class as Julia's randn() (ziggurat over a
 * 64-bit generator).  Reference RNG stream parity is unobtainable; these runs
 * are compared statistically only.  threads <= 0: all cores.  Returns path terminal sums.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t s[4]; } xo_state;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(xo_state *st) { /* xoshiro256++ (Julia >= 1.7 default generator family) */
  uint64_t *s = st->s, r = rotl64(s[0] + s[3], 23) + s[0], t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
  return r;
}
static inline uint64_t splitmix64(uint64_t *x) {
  uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static void xo_seed(xo_state *st, uint64_t seed) { for (int i = 0; i < 4; i++) st->s[i] = splitmix64(&seed); }

/* Ziggurat (256 layers, Marsaglia-Tsang construction with the Doornik tail), tables built at load. */
#define ZN 256
static double zig_x[ZN + 1], zig_r[ZN];
static int zig_ready = 0;
static void zig_init(void) {
  const double R = 3.6541528853610088, V = 0.00492867323399;
  double f = exp(-0.5 * R * R);
  zig_x[0] = V / f; zig_x[1] = R; zig_x[ZN] = 0.0;
  for (int i = 2; i < ZN; i++) {
    double xx = -2.0 * log(V / zig_x[i - 1] + f);
    zig_x[i] = sqrt(xx);
    f = exp(-0.5 * xx);
  }
  for (int i = 0; i < ZN; i++) zig_r[i] = zig_x[i + 1] / zig_x[i];
  zig_ready = 1;
}
static inline double u01(xo_state *st) { return (double)(xo_next(st) >> 11) * 0x1.0p-53; }
static inline double zig_randn(xo_state *st) {
  for (;;) {
    uint64_t b = xo_next(st);
    int i = (int)(b & 0xff);
    double u = 2.0 * ((double)(b >> 11) * 0x1.0p-53) - 1.0;
    if (fabs(u) < zig_r[i]) return u * zig_x[i];
    if (i == 0) {
      const double R = 3.6541528853610088;
      double xx, yy;
      do { xx = log(1.0 - u01(st)) / R; yy = log(1.0 - u01(st)); } while (-2.0 * yy < xx * xx);
      return u < 0 ? xx - R : R - xx;
    }
    double xx = u * zig_x[i];
    double f0 = exp(-0.5 * (zig_x[i] * zig_x[i] - xx * xx)), f1 = exp(-0.5 * (zig_x[i + 1] * zig_x[i + 1] - xx * xx));
    if (f1 + u01(st) * (f0 - f1) < 1.0) return xx;
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* sums[2*D]: per component Σx, Σx²; out (may be NULL): terminal states [npaths][D].  Returns threads used. */
int oracle_sample_rng(int id, int scheme, const double *p, double t0, double dt, const double *x0, int64_t nsteps,
                      int64_t npaths, uint64_t seed, int threads, double *sums, double *out) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  const int D = m->D, M = m->M > 0 ? m->M : 1;
  if (!zig_ready) zig_init();
  const double sdt = sqrt(dt);
  double acc[2 * MAXD] = {0};
  int used = 1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
  used = threads;
#pragma omp parallel num_threads(threads)
#endif
  {
    double loc[2 * MAXD] = {0};
    int tid = 0, nt = 1;
#ifdef _OPENMP
    tid = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
    xo_state st;
    xo_seed(&st, seed * 0x100000001b3ull + (uint64_t)tid);
    int64_t lo = npaths * tid / nt, hi = npaths * (tid + 1) / nt;
    for (int64_t k = lo; k < hi; k++) {
      double x[MAXD], xn[MAXD], dw[MAXM];
      for (int d = 0; d < D; d++) x[d] = x0[d];
      for (int64_t i = 1; i <= nsteps; i++) {
        double ti = t0 + (double)(i - 1) * dt;
        for (int j = 0; j < M; j++) dw[j] = sdt * (m->M > 0 ? zig_randn(&st) : 0.0);
        step(m, scheme, p, dt, ti, x, dw, xn);
        for (int d = 0; d < D; d++) x[d] = xn[d];
      }
      for (int d = 0; d < D; d++) {
        loc[2 * d] += x[d]; loc[2 * d + 1] += x[d] * x[d];
        if (out) out[k * D + d] = x[d];
      }
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    for (int q = 0; q < 2 * D; q++) acc[q] += loc[q];
  }
  for (int q = 0; q < 2 * D; q++) sums[q] = acc[q];
  return used;
}

/* Full-trajectory CPU baseline (simulate!, simulation.jl:104-117) — x[npaths][nrows][D] written. */
int oracle_simulate_rng(int id, int scheme, const double *p, double t0, double dt, const double *x0, int64_t nrows,
                        int64_t npaths, uint64_t seed, int threads, double *x) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  const int D = m->D, M = m->M > 0 ? m->M : 1;
  if (!zig_ready) zig_init();
  const double sdt = sqrt(dt);
  int used = 1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
  used = threads;
#pragma omp parallel num_threads(threads)
#endif
  {
    int tid = 0, nt = 1;
#ifdef _OPENMP
    tid = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
    xo_state st;
    xo_seed(&st, seed * 0x100000001b3ull + (uint64_t)tid);
    int64_t lo = npaths * tid / nt, hi = npaths * (tid + 1) / nt;
    for (int64_t k = lo; k < hi; k++) {
      double *xk = x + k * nrows * D, dw[MAXM];
      for (int d = 0; d < D; d++) xk[d] = x0[d];
      for (int64_t i = 2; i <= nrows; i++) {
        double t = t0 + (double)(i - 2) * dt;
        for (int j = 0; j < M; j++) dw[j] = sdt * (m->M > 0 ? zig_randn(&st) : 0.0);
        step(m, scheme, p, dt, t, xk + (i - 2) * D, dw, xk + (i - 1) * D);
      }
    }
  }
  return used;
}

/* Coupled-level CPU baseline (multilevel.jl:65-89): per path features Pc, Pf for a call payoff on x[0]. */
int oracle_multilevel_rng(int id, int scheme, const double *p, double t0, double dt_coarse, double dt_fine,
                          int64_t substeps, const double *x0, int64_t nsteps, int64_t npaths, uint64_t seed,
                          int threads, double *xc_out, double *xf_out) {
  if (id < 0 || id >= M_COUNT) return -1;
  const model_def *m = &MODELS[id];
  const int D = m->D, M = m->M > 0 ? m->M : 1;
  if (!zig_ready) zig_init();
  const double s = sqrt(dt_fine);
  int used = 1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
  used = threads;
#pragma omp parallel num_threads(threads)
#endif
  {
    int tid = 0, nt = 1;
#ifdef _OPENMP
    tid = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
    xo_state st;
    xo_seed(&st, seed * 0x100000001b3ull + (uint64_t)tid);
    int64_t lo = npaths * tid / nt, hi = npaths * (tid + 1) / nt;
    for (int64_t k = lo; k < hi; k++) {
      double xc[MAXD], xf[MAXD], xn[MAXD], DW[MAXM], dw[MAXM];
      for (int d = 0; d < D; d++) { xc[d] = x0[d]; xf[d] = x0[d]; }
      for (int64_t n = 1; n <= nsteps; n++) {
        for (int j = 0; j < M; j++) DW[j] = 0.0;
        double tc = t0 + (double)(n - 1) * dt_coarse;
        for (int64_t i = 1; i <= substeps; i++) {
          double tf = tc + (double)(i - 1) * dt_fine;
          for (int j = 0; j < M; j++) { dw[j] = s * (m->M > 0 ? zig_randn(&st) : 0.0); DW[j] = DW[j] + dw[j]; }
          step(m, scheme, p, dt_fine, tf, xf, dw, xn);
          for (int d = 0; d < D; d++) xf[d] = xn[d];
        }
        step(m, scheme, p, dt_coarse, tc, xc, DW, xn);
        for (int d = 0; d < D; d++) xc[d] = xn[d];
      }
      for (int d = 0; d < D; d++) { xc_out[k * D + d] = xc[d]; xf_out[k * D + d] = xf[d]; }
    }
  }
  return used;
}

#endif /* !ORACLE_REAL_FLOAT */
