/*
 * sdeb200.h — C ABI of libsdeb200.so: the B200 (sm_100a) device layer behind the Monte Carlo
 * path-simulation hot path of SDEModels.jl.  Citations are file:line in the reference repository.
 *
 * The reference has NO FFI / plugin interface (pure Julia): its extension mechanism is
 * multiple dispatch on `sample` / `simulate` / `simulate!` (src/SDEModels.jl:21-25).  Each entry point
 * below is what a Julia method of those generic functions `ccall`s (INTEGRATION.md shows the binding);
 * the comment on each names the reference method it replaces.
 *
 * Conventions
 *   - every function returns an int status (SDEB200_OK or a negative SDEB200_E_*); nothing throws or exits.
 *     sdeb200_last_error() returns the message of the last failure on the calling thread.
 *   - arrays are in Julia memory order: Array{SVector{D,Float64}}(nrows, npaths) column-major is the
 *     C array [npaths][nrows][D] (src/simulation.jl:71-76,149-154).
 *   - `params` are the model struct's fields in declaration order (test/runtests.jl:51-53), always double.
 *   - data pointers (`out`, `w`, `x`, ...) may be HOST or DEVICE memory; the library detects which
 *     (cudaPointerGetAttributes) and stages host buffers through pinned memory overlapped with compute.
 *     Element type is double for SDEB200_F64 and float for SDEB200_F32.
 *   - the caller owns every array; the library owns model handles (sdeb200_model_free).
 *   - calls are blocking; ONE call at a time per process (the reference itself is single-threaded and uses a
 *     process-global RNG, src/simulation.jl:4-18).  Any host thread may make that call: every entry point selects the
 *     primary device's context and makes its device current for the calling thread.  Internally a call may use one
 *     worker thread per GPU (sdeb200_init_devices); sdeb200_last_error() is per calling thread.
 *   - there is no CPU fallback: without a CUDA device every compute call returns SDEB200_E_CUDA.
 */
#ifndef SDEB200_H
#define SDEB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEB200_VERSION 100

/* status codes */
#define SDEB200_OK 0
#define SDEB200_E_ARG (-1)     /* bad argument */
#define SDEB200_E_PARSE (-2)   /* @sde_model text could not be parsed / classified */
#define SDEB200_E_NVRTC (-3)   /* run-time compilation of the generated functor failed (log in last_error) */
#define SDEB200_E_CUDA (-4)    /* CUDA runtime error, or no device */
#define SDEB200_E_NCCL (-5)    /* NCCL error */
#define SDEB200_E_DOMAIN (-6)  /* some path ended non-finite (Julia: DomainError from sqrt(V<0));
                                  outputs are still written */
#define SDEB200_E_UNSUPPORTED (-7) /* not defined by the reference either (e.g. Milstein with M != 1, milstein.jl:4) */

/* schemes: src/schemes/schemes.jl:34,37,42 */
#define SDEB200_EULER_MARUYAMA 0
#define SDEB200_MILSTEIN 1
#define SDEB200_RUNGE_KUTTA 2 /* src/schemes/runge_kutta.jl:4-11 — derivative-free Milstein, M == 1 like Milstein */
#define SDEB200_EXACT 3       /* closed-form one-step samplers (src/simulation.jl:50-57,78-89): built-in BlackScholes
                                 (black_scholes.jl:3-13) and OrnsteinUhlenbeck (ornstein_uhlenbeck.jl:3-13) only */
/* arithmetic type of state and noise */
#define SDEB200_F64 0
#define SDEB200_F32 1
/* math mode: FAST = the same scheme step with another rounding sequence — FMA contraction, parameter-only factors and dt
 * folded into prepared constants, geometric rows as x + x*A, mean reversion as x*(1 - k*dt) + k*theta*dt (Float64 only), and
 * for models whose noise terms share one state-dependent square root (Heston) ONE sqrt(A*w) per step with w the squared
 * radius of the step's normal pair: <= 1 ulp of the new state from the plain step, <= 1e-11 from the
 * oracle over 1024 steps.  STRICT = the reference's literal operation order without FMA contraction (bit-identical to a
 * CPU evaluation of the Julia expressions, euler_maruyama.jl:4, milstein.jl:8, runge_kutta.jl:7-9). */
#define SDEB200_MATH_FAST 0
#define SDEB200_MATH_STRICT 1
/* Normal stream versions.  Both are counter-based Philox4x32-10 streams keyed by (seed, stream, global
 * path index, step); they differ in how many Philox bits a pair of FP64 normals consumes: v1 128 (64-bit uniform + 64-bit
 * angle, two normals per block), v2 96 (64-bit uniform + 32-bit angle, eight normals per three blocks: -25 % generator
 * work).  FP32 runs use the same stream under both.  A given (seed, ...) produces DIFFERENT numbers under v1 and v2. */
#define SDEB200_RNG_DEFAULT 0
#define SDEB200_RNG_V1 1
#define SDEB200_RNG_V2 2
/* full-trajectory layouts */
#define SDEB200_LAYOUT_REFERENCE 0 /* [path][time][D] — what simulate returns (simulation.jl:71-76) */
#define SDEB200_LAYOUT_SOA 1       /* [time][D][path] — coalesced device layout */
/* payoffs for the fused estimators (new functionality: the reference returns raw states) */
#define SDEB200_PAYOFF_MOMENTS 0   /* every state component */
#define SDEB200_PAYOFF_CALL 1      /* disc * max(x[comp] - strike, 0) */
#define SDEB200_PAYOFF_PUT 2       /* disc * max(strike - x[comp], 0) */
#define SDEB200_PAYOFF_COMPONENT 3 /* x[comp] */

typedef struct sdeb200_model sdeb200_model;

/* Run options shared by all drivers.  Zero-initialise, then set what differs. */
typedef struct {
  int32_t scheme;      /* SDEB200_EULER_MARUYAMA | SDEB200_MILSTEIN | SDEB200_RUNGE_KUTTA | SDEB200_EXACT */
  int32_t precision;   /* SDEB200_F64 | SDEB200_F32 */
  int32_t math;        /* SDEB200_MATH_FAST | SDEB200_MATH_STRICT */
  int32_t rng;         /* SDEB200_RNG_DEFAULT (= v1) | SDEB200_RNG_V1 | SDEB200_RNG_V2: version of the normal stream */
  double t0;           /* start time (argument t0 of sample/simulate) */
  double dt;           /* scheme.Δt (src/schemes/schemes.jl:5-7) */
  uint64_t seed;       /* Philox key — replaces Random.seed! on the process-global RNG */
  uint32_t stream;     /* independent sub-stream selector; multilevel runs need stream < 2^24 (SDEB200_LEVEL_STREAM) */
  uint32_t reserved2;
  int64_t path_offset; /* global index of this call's first path: shards of one logical run pass their offset
                          so results do not depend on how paths are spread over GPUs */
} sdeb200_opts;

/* Philox stream word of level pair i of a multilevel run on the caller's stream: pair 0 (a plain `sample`,
 * multilevel.jl:56) is the stream itself, pair i >= 1 gets i in the top byte — a plain run on stream 1 no longer shares
 * its noise with level 1. */
#define SDEB200_LEVEL_STREAM(stream, i) ((uint32_t)(stream) + ((uint32_t)(i) << 24))

typedef struct {
  int32_t kind, comp;
  double strike, discount;
} sdeb200_payoff;

/* ---- library / device ---- */
int sdeb200_version(void);
const char *sdeb200_last_error(void);
int sdeb200_device_count(int *n);
int sdeb200_init(int device);              /* bind this process to one GPU (one process per GPU) */
/* One process, several GPUs: bind devices 0 .. n-1 (ndev_requested <= 0: all).  Afterwards every driver
 * whose data pointers are HOST memory (or NULL) shards its paths over the devices — contiguous ranges on
 * SDEB200_SHARD_ALIGN boundaries, one worker thread per device, the estimators' exact integer sums added on the host —
 * and returns the same bits as on one GPU.  Calls on DEVICE pointers run on device 0.  Not combinable with
 * sdeb200_comm_init.  *ndev_out = number of devices bound. */
int sdeb200_init_devices(int ndev_requested, int *ndev_out);
int sdeb200_device_info(int *ndev, int *primary);
int sdeb200_shutdown(void);
int sdeb200_set_stream(void *cuda_stream); /* run on a caller-owned cudaStream_t (NULL: internal stream) */
int sdeb200_synchronize(void);
double sdeb200_last_kernel_ms(void);       /* CUDA-event time of the path kernels of the last call */
int64_t sdeb200_launch_count(void);        /* kernels launched by this library so far */
int sdeb200_measure_peak(int precision, double *tflops, double *ms); /* FMA-chain micro-benchmark */
int sdeb200_set_tma(int enabled);          /* tuning switch: 0 routes the trajectory kernels that have a TMA (tensor-tile)
                                              form to their plain shared-memory-tile form; returns the previous value */
int64_t sdeb200_set_sample1_max(int64_t npaths); /* tuning switch: terminal-value launches of at most this many paths use the
                                              one-path-per-thread kernel form (default 262144; 0: never); returns the
                                              previous value.  Results do not depend on it (bit-identical forms). */
int64_t sdeb200_tma_launch_count(void);    /* launches of TMA kernels so far (tests assert that the path was taken) */

/* ---- models: what `@sde_model` produces (src/models/codegen.jl:87-159) ---- */
/* Parse the equations, build drift / diffusion (correlation Cholesky folded in, codegen.jl:105-112) and emit
 * a CUDA device functor; kernels are compiled with NVRTC on first use.  `eqs_utf8` is the macro body, one
 * equation per line or ';'-separated, e.g. "dS = r*S*dt + σ*S*dW".  param_names (may be NULL / nparams 0):
 * explicit parameter order, the macro's trailing arguments (codegen.jl:116-118). */
int sdeb200_model_create(const char *name, const char *eqs_utf8, const char *const *param_names, int nparams,
                         sdeb200_model **out);
/* "BlackScholes" (black_scholes.jl:1), "Heston" (models.jl:44-48) or "OrnsteinUhlenbeck" (ornstein_uhlenbeck.jl:1):
 * ahead-of-time compiled functors. */
int sdeb200_model_builtin(const char *name, sdeb200_model **out);
void sdeb200_model_free(sdeb200_model *m);
/* dim(model) == (D, M) (models.jl:3-5); kind: 0 AbstractSDE, 1 StateIndependentDiffusion, 2 JumpSDE (rejected) */
int sdeb200_model_info(const sdeb200_model *m, int *D, int *M, int *nparams, int *kind);
const char *sdeb200_model_name(const sdeb200_model *m);
const char *sdeb200_model_param_name(const sdeb200_model *m, int i);    /* fieldnames(Model)[i] */
const char *sdeb200_model_variable_name(const sdeb200_model *m, int i); /* variables(model)[i] */
const char *sdeb200_model_drift_expr(const sdeb200_model *m, int i);          /* symbolic, after simplification */
const char *sdeb200_model_diffusion_expr(const sdeb200_model *m, int i, int j);
const char *sdeb200_model_cuda_source(const sdeb200_model *m);          /* the generated functor */
/* drift(model, t, x) / diffusion(model, t, x) evaluated by the compiled functor ON THE DEVICE
 * (codegen.jl:57-75; known answers test/runtests.jl:61-82).  mu[D]; sigma[D*max(M,1)] row-major. */
int sdeb200_model_coefficients(sdeb200_model *m, const double *params, double t, const double *x, double *mu,
                               double *sigma);
/* force compilation of the (precision, math) kernel variant now; returns the NVRTC log on failure */
int sdeb200_model_compile(sdeb200_model *m, int precision, int math);

/* ---- drivers ---- */
/* sample(model, scheme, t0, x0, nsteps, npaths) -> Vector{T}: terminal states (simulation.jl:39-48,59-67).
 * out[npaths][D]. */
int sdeb200_sample(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0, int64_t nsteps,
                   int64_t npaths, void *out);

/* simulate(model, scheme, t0, x0, nsteps, npaths) -> x (simulation.jl:71-76,104-117): full trajectories,
 * start point included.  layout REFERENCE: out[npaths][nsteps+1][D]; SOA: out[nsteps+1][D][npaths]. */
int sdeb200_simulate(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                     int64_t nsteps, int64_t npaths, int layout, void *out);

/* simulate!(x, model, scheme, t0, x0, w) (simulation.jl:119-134): trajectories driven by the cumulative
 * Wiener path w[npaths][nrows][max(M,1)]; x[npaths][nrows][D]; x == w allowed (needs D == M). */
int sdeb200_simulate_wiener(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                            int64_t nrows, int64_t npaths, const void *w, void *x);

/* wiener!(w, Δt) (simulation.jl:22-35): w[npaths][nrows][dim], row 0 = 0, running sums of N(0, Δt). */
int sdeb200_wiener(const sdeb200_opts *o, int dim, int64_t nrows, int64_t npaths, void *w);

/* _multilevel_sample(model, coarse, fine, substeps, t0, x0, nsteps, npaths) (multilevel.jl:65-89):
 * coupled coarse / fine terminal states from the same Brownian increments.  o->dt is the COARSE Δt, the
 * fine one is dt_fine (the caller computes it as subdivide does, schemes.jl:53-54); ncoarse = number of
 * coarse steps.  xc[npaths][D], xf[npaths][D] (either may be NULL). */
int sdeb200_mlmc_sample_level(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                              double dt_fine, int64_t substeps, int64_t ncoarse, int64_t npaths, void *xc,
                              void *xf);

/* _multilevel_simulate (multilevel.jl:91-100): coupled full trajectories through a shared Wiener path.
 * xc[npaths][ncoarse+1][D], xf[npaths][ncoarse*substeps+1][D]; requires D == M like the reference. */
int sdeb200_mlmc_simulate_level(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                double dt_fine, int64_t substeps, int64_t ncoarse, int64_t npaths, void *xc,
                                void *xf);

/* logtransition(model, scheme::EulerMaruyama, times, data) (src/schemes/schemes.jl:64-76, euler_maruyama.jl:8-13):
 * Gaussian one-step log-densities along stored trajectories on the uniform grid t0 + i*dt.
 * x[npaths][nrows][D] -> out[npaths][nrows-1]; out[k][i] is the log-density of x[k][i+1] given x[k][i]. */
int sdeb200_logtransition(sdeb200_model *m, const sdeb200_opts *o, const double *params, int64_t nrows,
                          int64_t npaths, const void *x, void *out);

/* ---- fused estimators (north_star (c); no reference analogue, oracle = mean(payoff.(sample(...)))) ---- */
/* Terminal-value run with the payoff reduced on the fly: warp shuffles + CTA reduction to per-chunk partials,
 * exact integer accumulation, and — when a communicator is up — ONE ncclAllReduce over all ranks.
 * npaths is this rank's share.  result[2*nf + 2] = { sum f_0, sum f_0^2, ..., n, n_nonfinite } over ALL ranks,
 * nf = D for MOMENTS else 1.  Sums are bit-identical for any number of ranks when shards start on multiples
 * of SDEB200_SHARD_ALIGN paths.  terminal (may be NULL): also store the states, as sdeb200_sample does. */
#define SDEB200_SHARD_ALIGN 1024
int sdeb200_estimate(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                     int64_t nsteps, int64_t npaths, const sdeb200_payoff *payoff, void *terminal, double *result);

/* One MLMC level pair reduced on the fly: features Y = P(xf) - P(xc), P(xf), P(xc).
 * result[8] = { sum Y, sum Y^2, sum Pf, sum Pf^2, sum Pc, sum Pc^2, n, n_nonfinite }; one allreduce. */
int sdeb200_mlmc_estimate_level(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                double dt_fine, int64_t substeps, int64_t ncoarse, int64_t npaths,
                                const sdeb200_payoff *payoff, double *result);

/* Whole multilevel estimator: sample(model, MultilevelScheme(scheme, substeps, levels), t0, x0, nsteps, npaths)
 * (multilevel.jl:52-63) with the payoff reduced per level.  o->dt is the BASE scheme's Δt; level l uses
 * Δt / substeps^l (multilevel.jl:7-10); level index 0 of the call is a plain `sample`, level i >= 1 couples
 * levels i-1 and i with nsteps*substeps^(level_first+i-1) coarse steps.  Levels run concurrently on separate
 * streams; each ends in its own allreduce.  npaths[nlevels] = this rank's share per level.
 * result[nlevels][8] as sdeb200_mlmc_estimate_level (level 0: Y = Pf, Pc = 0). */
int sdeb200_mlmc_estimate(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                          int64_t substeps, int64_t level_first, int nlevels, int64_t nsteps,
                          const int64_t *npaths, const sdeb200_payoff *payoff, double *result);

/* sdeb200_mlmc_estimate with one path offset per level (path_offsets[nlevels], added to o->path_offset): the share of one
 * rank of a multi-process run in which every level is sharded over all ranks. */
int sdeb200_mlmc_estimate_sharded(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                  int64_t substeps, int64_t level_first, int nlevels, int64_t nsteps,
                                  const int64_t *npaths, const int64_t *path_offsets, const sdeb200_payoff *payoff,
                                  double *result);

/* ---- library-owned device buffers: results that stay on the GPU for callers without a device array type ---- */
/* shape[ndim] in C order (last extent contiguous); pass sdeb200_buffer_ptr(b) wherever a driver takes a data pointer. */
typedef struct sdeb200_buffer sdeb200_buffer;
int sdeb200_buffer_alloc(int precision, int ndim, const int64_t *shape, sdeb200_buffer **out);
void *sdeb200_buffer_ptr(const sdeb200_buffer *b);
int sdeb200_buffer_info(const sdeb200_buffer *b, int *precision, int *ndim, int64_t *shape, int64_t *strides_bytes);
int sdeb200_buffer_copy_to_host(const sdeb200_buffer *b, int64_t offset_elems, int64_t count_elems, void *host);
int sdeb200_buffer_copy_from_host(sdeb200_buffer *b, int64_t offset_elems, int64_t count_elems, const void *host);
int sdeb200_buffer_free(sdeb200_buffer *b);

/* ---- multi-GPU: one process per GPU, one NCCL communicator ---- */
int sdeb200_comm_unique_id(void *id128);                          /* rank 0: ncclGetUniqueId (128 bytes) */
int sdeb200_comm_init(const void *id128, int nranks, int rank);   /* all ranks: ncclCommInitRank */
int sdeb200_comm_destroy(void);
int sdeb200_comm_info(int *nranks, int *rank);

/* ---- exact accumulator, exposed for tests and for host-side combination of shards ---- */
/* Round the exact sum held in bins[SDEB200_NBINS] (int64 limbs, LSB 2^-1074) to the nearest double. */
#define SDEB200_NBINS 68
double sdeb200_bins_to_double(const int64_t *bins);
/* Add one double into bins exactly (host twin of the device accumulator). */
void sdeb200_bins_add(int64_t *bins, double v);

#ifdef __cplusplus
}
#endif
#endif
