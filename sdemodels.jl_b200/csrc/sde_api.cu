// sde_api.cu — host side of the C ABI declared in include/sdeb200.h.
// Owns: device binding, streams/events, scratch memory, host<->device staging pipelines, the kernel tables
// (AOT for the built-in functors, NVRTC for @sde_model-generated ones), the exact accumulator's host half,
// and the NCCL communicator (one process per GPU).  No compute happens on the CPU: every driver launches
// the kernels of sde_device.cuh and fails with SDEB200_E_CUDA when no device is available.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdeb200.h"
#include "sde_args.h"
#include "sde_dsl.h"
#include "sde_ktable.h"
#include "sde_coeffs.inc"

extern "C" int sde_aot_table_fast(int model, sde_ktable *t);
extern "C" int sde_aot_table_strict(int model, sde_ktable *t);
extern "C" const void *sde_aot_wiener_fast(int dim, int slot);
extern "C" const void *sde_aot_wiener_strict(int dim, int slot);
extern "C" void sde_launch_accumulate(const double *partials, int64_t nchunks, int stride, int nq,
                                      unsigned long long *bins, unsigned long long *overflow, cudaStream_t st);
extern "C" void sde_launch_subsample(int precision, const void *src, void *dst, int64_t npaths, int64_t nrows_dst,
                                     int64_t nrows_src, int64_t substeps, int dim, cudaStream_t st);
extern "C" double sde_launch_peak(int precision, void *scratch, int blocks, int iters, cudaStream_t st);
// sde_nvrtc.cpp
int sde_nvrtc_compile(const std::string &functor_src, int D, int M, int slot, int math, int family, sde_ktable *kt,
                      void **library_out, std::string *log);
enum { FAM_TERMINAL = 0, FAM_TRAJ = 1, FAM_WIENER = 2, FAM_MISC = 3, FAM_COUNT = 4 };  // kernel families (sde_device.cuh)

static_assert(SDEB200_NBINS == SDE_NBINS, "bins");
static_assert(SDEB200_SHARD_ALIGN == SDE_SHARD_ALIGN, "align");

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return fail(SDEB200_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// NCCL, resolved at run time so that the library loads on machines without it (and shares the copy
// PyTorch already mapped when the host process is Python)
// ---------------------------------------------------------------------------------------------------
struct nccl_uid { char b[128]; };
typedef void *nccl_comm;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(nccl_uid *) = nullptr;
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  for (const char *n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return fail(SDEB200_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                       \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                           \
  if (!g_nccl.field) return fail(SDEB200_E_NCCL, "libnccl lacks %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
#define NC(call)                                                                                         \
  do {                                                                                                   \
    int r_ = (call);                                                                                     \
    if (r_ != 0) return fail(SDEB200_E_NCCL, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(SDEB200_E_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

static const int MAX_LEVELS = 24;
static const int RED_WORDS = 2 * SDE_MAXF * SDE_NBINS + 4;  // bins + {n, nonfinite, overflow, pad}

struct Ctx {
  bool up = false;
  int device = -1;
  cudaStream_t own = nullptr, copy = nullptr, ext = nullptr;
  bool use_ext = false;
  cudaStream_t lvl[MAX_LEVELS] = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, kdone[2] = {}, cdone[2] = {}, lvl_done[MAX_LEVELS] = {};
  DevBuf slab[2], win[2], partials[MAX_LEVELS], red, misc;
  void *pin[2] = {nullptr, nullptr};   // pinned staging ring for pageable host outputs (run_slabs)
  size_t pin_cap[2] = {0, 0};
  std::vector<const void *> opted;     // kernels whose dynamic shared-memory limit was raised on THIS device
  double last_ms = 0.0;
  std::atomic<int64_t> launches{0}, tma_launches{0};
  nccl_comm comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream() const { return use_ext ? ext : own; }
};
// One context per device.  A call runs on the context of the calling thread: the primary device (sdeb200_init /
// sdeb200_init_devices), or — inside a sharded call — the device a worker thread was bound to (run_sharded).
static const int MAX_DEV = 16;
static Ctx g_ctx[MAX_DEV];
static int g_primary = 0;              // device of single-device calls
static int g_ndev = 1;                 // devices host-memory calls are spread over (sdeb200_init_devices)
static thread_local Ctx *tl_ctx = &g_ctx[0];
static thread_local bool tl_shard = false;                       // this thread is a worker of a sharded call
static thread_local unsigned long long *tl_raw = nullptr;        // worker: where finish_reductions leaves the raw integer records
static thread_local const int64_t *tl_level_off = nullptr;       // worker: per-level path offsets of a sharded multilevel call
#define g (*tl_ctx)
static std::mutex g_compile_mutex;     // NVRTC compilation / kernel-table updates of a model

static int ctx_init(Ctx &c, int device) {
  if (c.up) return 0;
  CU(cudaSetDevice(device));
  c.device = device;
  CU(cudaStreamCreateWithFlags(&c.own, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c.ev0));
  CU(cudaEventCreate(&c.ev1));
  for (int i = 0; i < 2; i++) {
    CU(cudaEventCreateWithFlags(&c.kdone[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c.cdone[i], cudaEventDisableTiming));
  }
  c.up = true;
  return 0;
}
static void ctx_release(Ctx &c) {
  if (!c.up) return;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  if (c.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c.comm);
  c.comm = nullptr;
  c.nranks = 1;
  c.rank = 0;
  for (int i = 0; i < 2; i++) {
    c.slab[i].release();
    c.win[i].release();
    if (c.pin[i]) cudaFreeHost(c.pin[i]);
    c.pin[i] = nullptr;
    c.pin_cap[i] = 0;
    if (c.kdone[i]) cudaEventDestroy(c.kdone[i]);
    if (c.cdone[i]) cudaEventDestroy(c.cdone[i]);
    c.kdone[i] = c.cdone[i] = nullptr;
  }
  for (int i = 0; i < MAX_LEVELS; i++) {
    c.partials[i].release();
    if (c.lvl[i]) cudaStreamDestroy(c.lvl[i]);
    if (c.lvl_done[i]) cudaEventDestroy(c.lvl_done[i]);
    c.lvl[i] = nullptr;
    c.lvl_done[i] = nullptr;
  }
  c.red.release();
  c.misc.release();
  if (c.ev0) cudaEventDestroy(c.ev0);
  if (c.ev1) cudaEventDestroy(c.ev1);
  if (c.own) cudaStreamDestroy(c.own);
  if (c.copy) cudaStreamDestroy(c.copy);
  c.ev0 = c.ev1 = nullptr;
  c.own = c.copy = nullptr;
  c.use_ext = false;
  c.opted.clear();
  c.up = false;
}

// Every entry point starts here: select the calling thread's context (worker threads keep the one they were bound to)
// and make its device current for this thread (the CUDA current device is per host thread).
static int ensure_up() {
  if (!tl_shard) tl_ctx = &g_ctx[g_primary];
  if (!g.up) {
    int rc = sdeb200_init(-1);
    if (rc) return rc;
  }
  CU(cudaSetDevice(g.device));
  return 0;
}

extern "C" int sdeb200_version(void) { return SDEB200_VERSION; }
extern "C" const char *sdeb200_last_error(void) { return g_err.c_str(); }
extern "C" int sdeb200_device_count(int *n) {
  if (!n) return fail(SDEB200_E_ARG, "n is NULL");
  *n = 0;
  CU(cudaGetDeviceCount(n));
  return SDEB200_OK;
}
extern "C" int sdeb200_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SDEB200_E_CUDA, "no CUDA device (%s); sdeb200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0) {
    if (g_ctx[g_primary].up) return SDEB200_OK;
    CU(cudaGetDevice(&device));
  }
  if (device >= n || device >= MAX_DEV) return fail(SDEB200_E_ARG, "device %d out of range (%d devices)", device, n);
  if (g_ctx[g_primary].up && g_primary != device && g_ndev == 1) ctx_release(g_ctx[g_primary]);   // re-bind: one GPU per process
  g_primary = device;
  g_ndev = 1;
  tl_ctx = &g_ctx[device];
  if (g_ctx[device].up && g_ctx[device].device != device) ctx_release(g_ctx[device]);   // left over from SDEB200_FAKE_DEVICES
  return ctx_init(g_ctx[device], device);
}
// one host process owning several GPUs.  Binds devices 0 .. n-1 (ndev_requested <= 0: all); calls whose
// data pointers are HOST memory (or null) are then sharded over them — contiguous path ranges on SDEB200_SHARD_ALIGN
// boundaries, one worker thread per device, exact integer sums combined on the host — with results identical, bit for
// bit, to a single-GPU run (the global path index is the Philox counter).  Calls on DEVICE pointers run on the primary
// device (0).  Not to be mixed with sdeb200_comm_init (one process per GPU).
extern "C" int sdeb200_init_devices(int ndev_requested, int *ndev_out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SDEB200_E_CUDA, "no CUDA device (%s); sdeb200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  int want = ndev_requested <= 0 ? n : std::min(ndev_requested, n);
  // test hook: SDEB200_FAKE_DEVICES=k binds k logical devices round-robin onto the physical ones, so that the sharding,
  // the worker threads and the host-side combination of the exact sums can be exercised on a single-GPU box
  if (const char *fake = getenv("SDEB200_FAKE_DEVICES")) want = std::max(1, atoi(fake));
  want = std::min(want, MAX_DEV);
  if (g_ctx[g_primary].comm) return fail(SDEB200_E_ARG, "a NCCL communicator is up: one process per GPU and one process for all GPUs do not mix");
  if (g_ndev == 1 && g_primary != 0) ctx_release(g_ctx[g_primary]);
  for (int d = 0; d < want; d++) {
    if (g_ctx[d].up && g_ctx[d].device != d % n) ctx_release(g_ctx[d]);
    int rc = ctx_init(g_ctx[d], d % n);
    if (rc) return rc;
  }
  for (int d = want; d < MAX_DEV; d++) ctx_release(g_ctx[d]);
  g_primary = 0;
  g_ndev = want;
  tl_ctx = &g_ctx[0];
  CU(cudaSetDevice(0));
  if (ndev_out) *ndev_out = want;
  return SDEB200_OK;
}
extern "C" int sdeb200_device_info(int *ndev, int *primary) {
  if (ndev) *ndev = g_ndev;
  if (primary) *primary = g_primary;
  return SDEB200_OK;
}
extern "C" int sdeb200_shutdown(void) {
  for (int d = 0; d < MAX_DEV; d++) ctx_release(g_ctx[d]);
  g_ndev = 1;
  return SDEB200_OK;
}
extern "C" int sdeb200_set_stream(void *s) {
  int rc = ensure_up();
  if (rc) return rc;
  g.ext = (cudaStream_t)s;
  g.use_ext = true;  // NULL selects the legacy default stream, which is what a torch default stream is
  return SDEB200_OK;
}
extern "C" int sdeb200_synchronize(void) {
  int rc = ensure_up();
  if (rc) return rc;
  CU(cudaStreamSynchronize(g.stream()));
  return SDEB200_OK;
}
extern "C" double sdeb200_last_kernel_ms(void) { return g_ctx[g_primary].last_ms; }
extern "C" int64_t sdeb200_launch_count(void) {
  int64_t n = 0;
  for (int d = 0; d < MAX_DEV; d++) n += g_ctx[d].launches.load();
  return n;
}

extern "C" int sdeb200_measure_peak(int precision, double *tflops, double *ms_out) {
  int rc = ensure_up();
  if (rc) return rc;
  if (!tflops) return fail(SDEB200_E_ARG, "tflops is NULL");
  const int blocks = 148 * 2, iters = 4096;
  rc = g.misc.ensure((size_t)blocks * 1024 * 8);
  if (rc) return rc;
  cudaStream_t st = g.stream();
  sde_launch_peak(precision, g.misc.p, blocks, 64, st);  // warm-up
  double best = 0.0, best_ms = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    CU(cudaEventRecord(g.ev0, st));
    double flops = sde_launch_peak(precision, g.misc.p, blocks, iters, st);
    g.launches++;
    CU(cudaEventRecord(g.ev1, st));
    CU(cudaEventSynchronize(g.ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, g.ev0, g.ev1));
    double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) { best = tf; best_ms = ms; }
  }
  CU(cudaGetLastError());
  *tflops = best;
  if (ms_out) *ms_out = best_ms;
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// models
// ---------------------------------------------------------------------------------------------------
struct sdeb200_model {
  std::string name;
  int D = 0, M = 0, kind = 0, builtin = -1;
  bool has_exact = false;
  std::vector<std::string> params, vars, drift, diffusion;  // diffusion row-major D x max(M,1)
  std::string source;                                       // generated functor (empty for built-ins)
  std::string milstein_error;                               // why Milstein is unavailable (diffusion not differentiable symbolically)
  sde_ktable kt[2];                                         // [math]
  bool have[2][3][FAM_COUNT] = {};                          // [math][slot][family]: compiled (NVRTC, lazily) or built in
  std::vector<void *> libs;
};

extern "C" int sdeb200_model_builtin(const char *name, sdeb200_model **out) {
  if (!name || !out) return fail(SDEB200_E_ARG, "null argument");
  sdeb200_model *m = new sdeb200_model();
  m->name = name;
  if (m->name == "BlackScholes") {
    m->builtin = 0; m->D = 1; m->M = 1;
    m->params = {"r", "σ"}; m->vars = {"S"};
    m->drift = {"r * S"}; m->diffusion = {"σ * S"};
  } else if (m->name == "Heston") {
    m->builtin = 1; m->D = 2; m->M = 2;
    m->params = {"r", "κ", "θ", "σ", "ρ"}; m->vars = {"S", "V"};
    m->drift = {"r * S", "κ * (θ - V)"};
    m->diffusion = {"sqrt(V) * S", "0", "σ * sqrt(V) * ρ", "σ * sqrt(V) * sqrt(1 - ρ ^ 2)"};
  } else if (m->name == "OrnsteinUhlenbeck") {
    m->builtin = 2; m->D = 1; m->M = 1; m->kind = 1;
    m->params = {"θ", "μ", "σ"}; m->vars = {"x"};
    m->drift = {"θ * (μ - x)"}; m->diffusion = {"σ"};
  } else {
    delete m;
    return fail(SDEB200_E_ARG, "no built-in model named '%s' (BlackScholes, Heston, OrnsteinUhlenbeck)", name);
  }
  m->has_exact = m->builtin == 0 || m->builtin == 2;
  sde_aot_table_fast(m->builtin, &m->kt[0]);
  sde_aot_table_strict(m->builtin, &m->kt[1]);
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 3; b++)
      for (int f = 0; f < FAM_COUNT; f++) m->have[a][b][f] = true;
  *out = m;
  return SDEB200_OK;
}

extern "C" int sdeb200_model_create(const char *name, const char *eqs, const char *const *param_names, int nparams,
                                    sdeb200_model **out) {
  if (!name || !eqs || !out) return fail(SDEB200_E_ARG, "null argument");
  std::vector<std::string> pn;
  for (int i = 0; i < nparams; i++) pn.push_back(param_names[i]);
  sde_dsl_model dm;
  std::string err;
  if (!sde_dsl_build(name, eqs, pn, &dm, &err)) return fail(SDEB200_E_PARSE, "%s", err.c_str());
  if (dm.kind == 2)
    return fail(SDEB200_E_UNSUPPORTED, "model '%s' is a JumpSDE; jump diffusions are outside the accelerated path", name);
  if (dm.D > SDE_MAXD || dm.M > SDE_MAXM || (int)dm.params.size() > SDE_MAXP)
    return fail(SDEB200_E_UNSUPPORTED, "model '%s' exceeds D<=%d, M<=%d, %d parameters", name, SDE_MAXD, SDE_MAXM, SDE_MAXP);
  sdeb200_model *m = new sdeb200_model();
  m->name = name;
  m->D = dm.D; m->M = dm.M; m->kind = dm.kind;
  m->params = dm.params; m->vars = dm.vars; m->drift = dm.drift_str; m->diffusion = dm.diffusion_str;
  m->source = dm.cuda_source;
  m->milstein_error = dm.milstein_error;
  memset(m->kt, 0, sizeof m->kt);
  *out = m;
  return SDEB200_OK;
}

extern "C" void sdeb200_model_free(sdeb200_model *m) {
  if (!m) return;
  for (void *l : m->libs) cudaLibraryUnload((cudaLibrary_t)l);
  delete m;
}
extern "C" int sdeb200_model_info(const sdeb200_model *m, int *D, int *M, int *np, int *kind) {
  if (!m) return fail(SDEB200_E_ARG, "model is NULL");
  if (D) *D = m->D;
  if (M) *M = m->M;
  if (np) *np = (int)m->params.size();
  if (kind) *kind = m->kind;
  return SDEB200_OK;
}
extern "C" const char *sdeb200_model_name(const sdeb200_model *m) { return m ? m->name.c_str() : nullptr; }
extern "C" const char *sdeb200_model_param_name(const sdeb200_model *m, int i) {
  return (m && i >= 0 && i < (int)m->params.size()) ? m->params[i].c_str() : nullptr;
}
extern "C" const char *sdeb200_model_variable_name(const sdeb200_model *m, int i) {
  return (m && i >= 0 && i < (int)m->vars.size()) ? m->vars[i].c_str() : nullptr;
}
extern "C" const char *sdeb200_model_drift_expr(const sdeb200_model *m, int i) {
  return (m && i >= 0 && i < (int)m->drift.size()) ? m->drift[i].c_str() : nullptr;
}
extern "C" const char *sdeb200_model_diffusion_expr(const sdeb200_model *m, int i, int j) {
  if (!m) return nullptr;
  const int MW = m->M > 0 ? m->M : 1;
  const int k = i * MW + j;
  return (i >= 0 && j >= 0 && j < MW && k < (int)m->diffusion.size()) ? m->diffusion[k].c_str() : nullptr;
}
extern "C" const char *sdeb200_model_cuda_source(const sdeb200_model *m) { return m ? m->source.c_str() : nullptr; }

// Compile one kernel family of a generated model for (math, slot) if that has not happened yet: NVRTC costs seconds, so
// a model pays only for what it is used for ("lazily per (scheme, precision, kernel-kind)").
static int ensure_family(sdeb200_model *m, int math, int slot, int family) {
  if (m->have[math][slot][family]) return SDEB200_OK;
  std::lock_guard<std::mutex> lock(g_compile_mutex);
  if (m->have[math][slot][family]) return SDEB200_OK;
  int rc = ensure_up();
  if (rc) return rc;
  void *lib = nullptr;
  std::string log;
  rc = sde_nvrtc_compile(m->source, m->D, m->M, slot, math, family, &m->kt[math], &lib, &log);
  if (rc) return fail(rc, "NVRTC/load of model '%s' failed:\n%s", m->name.c_str(), log.c_str());
  m->libs.push_back(lib);
  m->have[math][slot][family] = true;
  return SDEB200_OK;
}

extern "C" int sdeb200_model_compile(sdeb200_model *m, int precision, int math) {
  if (!m) return fail(SDEB200_E_ARG, "model is NULL");
  if (precision < 0 || precision > 1 || math < 0 || math > 1) return fail(SDEB200_E_ARG, "bad precision/math");
  for (int f = 0; f < FAM_COUNT; f++) {
    int rc = ensure_family(m, math, precision == SDEB200_F32 ? 1 : 0, f);
    if (rc) return rc;
  }
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
// kernel-table slot of a run: 0 = f64 on normal stream v1, 1 = f32 (one stream definition), 2 = f64 on normal stream v2
static inline int kslot(const sdeb200_opts *o) { return o->precision == SDEB200_F32 ? 1 : (o->rng == SDEB200_RNG_V2 ? 2 : 0); }

static int check_common(const sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0) {
  if (!m || !o || !x0) return fail(SDEB200_E_ARG, "null argument");
  if (!params && !m->params.empty()) return fail(SDEB200_E_ARG, "params is NULL but the model has %zu", m->params.size());
  if (o->scheme < SDEB200_EULER_MARUYAMA || o->scheme > SDEB200_EXACT)
    return fail(SDEB200_E_ARG, "unknown scheme %d", o->scheme);
  if (o->scheme == SDEB200_EXACT && !m->has_exact)
    return fail(SDEB200_E_UNSUPPORTED, "the reference defines an Exact sampler for BlackScholes, OrnsteinUhlenbeck, "
                "CoxIngersollRoss and Merton only; the B200 path implements the first two (built-in models), not '%s'",
                m->name.c_str());
  if (o->scheme != SDEB200_EULER_MARUYAMA && o->scheme != SDEB200_EXACT && m->M != 1)
    return fail(SDEB200_E_UNSUPPORTED,
                "%s is defined for models with one Wiener process only (M == 1), as in the reference "
                "(src/schemes/milstein.jl:4, runge_kutta.jl:4); '%s' has M == %d",
                o->scheme == SDEB200_MILSTEIN ? "Milstein" : "RungeKutta", m->name.c_str(), m->M);
  if (o->scheme == SDEB200_MILSTEIN && !m->milstein_error.empty())
    return fail(SDEB200_E_UNSUPPORTED, "Milstein needs the derivative of the diffusion (src/schemes/milstein.jl:1-2) and the "
                "front end %s in '%s'; use RungeKutta (derivative-free, runge_kutta.jl:4-11) or EulerMaruyama",
                m->milstein_error.c_str(), m->name.c_str());
  if (o->precision < 0 || o->precision > 1 || o->math < 0 || o->math > 1) return fail(SDEB200_E_ARG, "bad precision/math");
  if (o->rng < 0 || o->rng > SDEB200_RNG_V2) return fail(SDEB200_E_ARG, "unknown normal stream version %d", o->rng);
  if (!(o->dt > 0.0) || !isfinite(o->dt)) return fail(SDEB200_E_ARG, "scheme Δt must be positive and finite");
  return 0;
}

// a->dt and the dt-scaled constants of the FP64 normal transform (sde_device.cuh: make_log_consts / make_log_consts_v2)
static void set_dt(sde_args *a, double dt, int scheme = SDEB200_EULER_MARUYAMA, int rng = SDEB200_RNG_V1) {
  a->dt = dt;
  const double v = scheme == SDEB200_EXACT ? 1.0 : dt;  // Exact samplers draw plain N(0,1) (rand(LogNormal / Normal))
  a->noise_var = v;
  if (rng == SDEB200_RNG_V2 && SDE_V2_FINE) {
    a->lk[0] = SDE_LOG2_C0_V * v;
    a->lk[1] = SDE_LOG2_C1_V * v;
    a->lk[2] = SDE_LOG2_C2_V * v;
    a->lk[3] = SDE_LOG2_C3_V * v;
    a->lk[4] = 0.0;
  } else {
    a->lk[0] = SDE_LOG_C1_V * v;
    a->lk[1] = SDE_LOG_C2_V * v;
    a->lk[2] = SDE_LOG_C3_V * v;
    a->lk[3] = SDE_LOG_C4_V * v;
    a->lk[4] = -2.0 * v;
  }
  a->lk[5] = SDE_TWO_LN2_V * v;
}

static void fill_args(sde_args *a, const sdeb200_model *m, const sdeb200_opts *o, const double *params,
                      const double *x0) {
  memset(a, 0, sizeof *a);
  for (size_t i = 0; i < m->params.size(); i++) a->params[i] = params[i];
  for (int d = 0; d < m->D; d++) a->x0[d] = x0[d];
  a->t0 = o->t0;
  a->dt = o->dt;
  a->seed_lo = (uint32_t)o->seed;
  a->seed_hi = (uint32_t)(o->seed >> 32);
  a->stream = o->stream;
  a->path0 = o->path_offset;
  set_dt(a, o->dt, o->scheme, o->rng);
}

// sample / mlmc kernels walk over chunks grid-stride: cap the grid so each CTA stages its tables once for many chunks
static int64_t capped(int64_t nchunks) { return std::min<int64_t>(nchunks, 148 * 64); }

static inline size_t esize(int precision);
static inline int64_t cdiv(int64_t a, int64_t b);
static int launch(const void *k, int64_t grid, const sde_args *a, cudaStream_t st, size_t smem = 0, int block = SDE_BLOCK) {
  if (!k) return fail(SDEB200_E_UNSUPPORTED, "kernel not available for this model / scheme");
  if (grid <= 0) return 0;
  if (grid > 2147483647LL) return fail(SDEB200_E_ARG, "too many paths for one launch");
  if (smem > 0) {  // static + dynamic shared memory above 48 KB needs an opt-in, once per kernel
    if (std::find(g.opted.begin(), g.opted.end(), k) == g.opted.end()) {   // per device
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      g.opted.push_back(k);
    }
  }
  void *args[1] = {(void *)a};
  CU(cudaLaunchKernel(k, dim3((unsigned)grid), dim3((unsigned)block), args, smem, st));
  g.launches++;
  return 0;
}

// reference-layout trajectory kernels (simulate_ref_body): a warp stages 64 paths x TS rows; see sde_device.cuh
static const int SIMREF_PW = SDE_SIMREF_PW;
static size_t simref_smem(int D, int precision) {
  const int ts = (precision == SDEB200_F64 ? 16 : 32) / D, row = ts * D, lds = row | 1;
  return (size_t)(SDE_BLOCK / 32) * 32 * SIMREF_PW * lds * esize(precision);
}
// Wiener-driven kernel (simulate_wiener_body): per warp a w tile [32][WLD] and an x tile [32][XLD]
static size_t simw_smem(int D, int M, int precision) {
  const int mw = M > 0 ? M : 1, dm = D > mw ? D : mw;
  const int ts = (SDE_SIMW_ROWBYTES / (int)esize(precision)) / dm;
  const int wld = (ts * mw) | 1, xld = (ts * D) | 1;
  return (size_t)(SDE_BLOCK / 32) * 32 * (wld + xld) * esize(precision);
}
static int64_t g_sample1_max = getenv("SDEB200_SAMPLE1_MAX") ? atoll(getenv("SDEB200_SAMPLE1_MAX")) : 262144;
static bool g_tma_enabled = getenv("SDEB200_NO_TMA") == nullptr;
extern "C" int sdeb200_set_tma(int enabled) {
  const int prev = g_tma_enabled ? 1 : 0;
  g_tma_enabled = enabled != 0;
  return prev;
}
extern "C" int64_t sdeb200_tma_launch_count(void) {
  int64_t n = 0;
  for (int d = 0; d < MAX_DEV; d++) n += g_ctx[d].tma_launches.load();
  return n;
}
extern "C" int64_t sdeb200_set_sample1_max(int64_t npaths) {
  const int64_t prev = g_sample1_max;
  g_sample1_max = npaths;
  return prev;
}

// Wiener-driven kernel on TMA tensor tiles (simulate_wiener_tma_body): per warp a ring of SDE_SIMWT_STAGES 4 KB tiles
static size_t simwt_smem() { return (size_t)(SDE_BLOCK / 32) * SDE_SIMWT_STAGES * (4096 + 8) + 1024; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int tmap_encoder(tmap_encode_fn *out) {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    if (!p || qr != cudaDriverEntryPointSuccess) return fail(SDEB200_E_CUDA, "the driver has no cuTensorMapEncodeTiled");
    fn = (tmap_encode_fn)p;
  }
  *out = fn;
  return 0;
}
// Tensor map of a trajectory buffer in reference order [path][row][D]: g paths per map row (sde_args.h), boxes of
// (32/g map rows) x (128 bytes), 128-byte shared-memory swizzle.  npaths must be a multiple of g.
static_assert(sizeof(sde_tmap) == sizeof(CUtensorMap) && alignof(sde_tmap) == alignof(CUtensorMap), "sde_tmap mirrors CUtensorMap");
static int make_traj_tmap(sde_tmap *tm, const void *base, int precision, int D, int64_t nrows, int64_t npaths, int grp,
                          int boxrows = 0, int rowbytes = 128) {   // rowbytes: 128 (128-byte swizzle) or 64 (64-byte swizzle)
  tmap_encode_fn enc = nullptr;
  int rc = tmap_encoder(&enc);
  if (rc) return rc;
  const size_t es = precision == SDEB200_F64 ? 8 : 4;
  const cuuint64_t dims[2] = {(cuuint64_t)grp * nrows * D, (cuuint64_t)(npaths / grp)};
  const cuuint64_t strides[1] = {(cuuint64_t)grp * nrows * D * es};
  const cuuint32_t box[2] = {(cuuint32_t)(rowbytes / es), (cuuint32_t)(boxrows > 0 ? boxrows : 32 / grp)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap *)tm, precision == SDEB200_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   rowbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SDEB200_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// reference-order trajectories with TMA tensor stores (simulate_ref_tma_body): per warp a ring of two 4 KB tiles
static int simrt_rowbytes(int precision) { return precision == SDEB200_F64 ? SDE_SIMRT2_ROWB_F64 : SDE_SIMRT2_ROWB_F32; }
static size_t simrt_smem(int precision) {
  const bool manual = precision == SDEB200_F64 ? SDE_SIMRT2_MANUAL_F64 != 0 : SDE_SIMRT2_MANUAL_F32 != 0;
  const int stages = manual ? 1 : (precision == SDEB200_F64 ? SDE_SIMRT2_STAGES_F64 : SDE_SIMRT2_STAGES_F32);
  return (size_t)(SDE_BLOCK / 32) * stages * 32 * simrt_rowbytes(precision) + 1024;
}
static int64_t simref_grid(int64_t npaths);
static size_t simref_smem(int D, int precision);
// steps per Philox batch of a (precision, stream version, M) combination: Batch<T, M, V>::U of sde_device.cuh
static int batch_steps(int M, int slot) {
  if (M <= 0) return 1;
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  if (slot == 2) return 8 / gcd(M, 8);
  const int npb = slot == 1 ? 4 : 2;
  return npb / gcd(M, npb);
}
// One RNG-driven reference-order launch into a device buffer: the TMA-store kernel K2u (one sub-path per warp, whole
// 16-byte chunks per Philox batch) where its geometry applies — rows of 4, 8 or 16 bytes, batches of whole chunks that
// divide a 128-byte tile row, 16-byte aligned buffer, paths of at least two tiles — the shared-memory-tile kernel K2b
// otherwise and for the npaths % g leftover paths.  `a` carries everything but npaths / out (and path0 of the first path).
static int launch_simref(const void *k, const void *k_tma, int D, int M, int slot, sde_args a, int64_t npaths, void *out,
                         cudaStream_t st) {
  const int precision = slot == 1 ? SDEB200_F32 : SDEB200_F64;
  const size_t es = precision == SDEB200_F64 ? 8 : 4;
  const int rowb = (int)(D * es);
  const int64_t nrows = a.nsteps + 1;
  const int bb = batch_steps(M, slot) * rowb;
  int rc;
  int64_t done = 0;
  const int talign = precision == SDEB200_F64 ? SDE_SIMRT2_ALIGN_F64 : SDE_SIMRT2_ALIGN_F32;
  if (k_tma && g_tma_enabled && M > 0 && 16 % rowb == 0 && bb % 16 == 0 && simrt_rowbytes(precision) % bb == 0 &&
      ((uintptr_t)out % talign) == 0 &&
      nrows >= 2 * (128 / rowb) + 4 && nrows * D < (1 << 28) && npaths < ((int64_t)1 << 31)) {
    const int grp = sde_tma_group(nrows * rowb, talign);
    const int64_t ncov = npaths - npaths % grp;
    if (ncov >= 32) {
      sde_args b = a;
      b.npaths = ncov;
      b.out = out;
      if ((rc = make_traj_tmap(&b.tm_x, out, precision, D, nrows, ncov, grp, 32, simrt_rowbytes(precision)))) return rc;
      const int64_t units = (int64_t)grp * cdiv(ncov / grp, 32);     // warps: sub-path j of 32 map rows each
      if ((rc = launch(k_tma, cdiv(units, SDE_BLOCK / 32), &b, st, simrt_smem(precision)))) return rc;
      g.tma_launches++;
      done = ncov;
    }
  }
  if (done < npaths) {
    a.npaths = npaths - done;
    a.path0 += done;
    a.out = (char *)out + (size_t)done * nrows * D * es;
    if ((rc = launch(k, simref_grid(npaths - done), &a, st, simref_smem(D, precision)))) return rc;
  }
  return 0;
}

static int64_t simref_grid(int64_t npaths) { return cdiv(npaths, (int64_t)SDE_BLOCK * SIMREF_PW); }

static bool is_device_ptr(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static inline size_t esize(int precision) { return precision == SDEB200_F64 ? 8 : 4; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

static int get_ktable(sdeb200_model *m, const sdeb200_opts *o, const sde_ktable **kt, int family) {
  int rc = ensure_family(m, o->math, kslot(o), family);
  if (rc) return rc;
  *kt = &m->kt[o->math];
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// one process, several GPUs: shard a call over the bound devices
// ---------------------------------------------------------------------------------------------------
static bool is_host_or_null(const void *p) { return p == nullptr || !is_device_ptr(p); }
static bool sharded_call(int64_t npaths) { return g_ndev > 1 && !tl_shard && npaths >= 2 * SDE_SHARD_ALIGN; }

// Contiguous path range of device d: boundaries on multiples of SDE_SHARD_ALIGN (the chunking of the payoff reduction,
// and with it every rounding, is then independent of the number of devices) — the same split distributed.shard_range
// applies across processes.
static void shard_range(int64_t npaths, int d, int nd, int64_t *lo, int64_t *hi) {
  const int64_t units = cdiv(npaths, SDE_SHARD_ALIGN);
  *lo = std::min<int64_t>(npaths, units * d / nd * SDE_SHARD_ALIGN);
  *hi = std::min<int64_t>(npaths, units * (d + 1) / nd * SDE_SHARD_ALIGN);
}

// fn(d, lo, n) runs on a worker thread bound to device d (its own context, streams and scratch buffers) for the paths
// [lo, lo + n); all devices run concurrently.  Errors: the first hard error wins; SDEB200_E_DOMAIN (outputs written, some
// paths non-finite) is reported after everything else succeeded, like the single-device call does.
template <class Fn> static int run_devices(Fn &&fn) {   // fn(d) on a worker thread bound to device d, all devices at once
  const int nd = g_ndev;
  std::vector<int> rcs(nd, 0);
  std::vector<std::string> errs(nd);
  std::vector<double> ms(nd, 0.0);
  std::vector<std::thread> th;
  for (int d = 0; d < nd; d++) {
    th.emplace_back([&, d]() {
      tl_ctx = &g_ctx[d];
      tl_shard = true;
      cudaSetDevice(g_ctx[d].device);
      g_ctx[d].last_ms = 0.0;
      rcs[d] = fn(d);
      if (rcs[d]) errs[d] = g_err;
      ms[d] = g_ctx[d].last_ms;
    });
  }
  for (auto &t : th) t.join();
  cudaSetDevice(g_ctx[g_primary].device);
  g_ctx[g_primary].last_ms = *std::max_element(ms.begin(), ms.end());
  int domain = -1;
  for (int d = 0; d < nd; d++) {
    if (rcs[d] == SDEB200_E_DOMAIN) { if (domain < 0) domain = d; continue; }
    if (rcs[d]) return fail(rcs[d], "device %d: %s", d, errs[d].c_str());
  }
  if (domain >= 0) return fail(SDEB200_E_DOMAIN, "%s", errs[domain].c_str());
  return 0;
}
template <class Fn> static int run_sharded(int64_t npaths, Fn &&fn) {
  const int nd = g_ndev;
  return run_devices([&](int d) -> int {
    int64_t lo, hi;
    shard_range(npaths, d, nd, &lo, &hi);
    return hi > lo ? fn(d, lo, hi - lo) : 0;
  });
}

// memcpy with a few threads: a single core moves ~10 GB/s, less than one PCIe 5 link delivers
static void par_memcpy(void *dst, const void *src, size_t n) {
  const size_t MINB = (size_t)8 << 20;
  const int nt = (int)std::min<size_t>(4, n / MINB);
  if (nt <= 1) { memcpy(dst, src, n); return; }
  std::vector<std::thread> th;
  const size_t per = (n / nt + 63) & ~(size_t)63;
  for (int i = 0; i < nt; i++) {
    const size_t o = (size_t)i * per, len = i == nt - 1 ? n - o : per;
    if (o >= n) break;
    th.emplace_back([=]() { memcpy((char *)dst + o, (const char *)src + o, len); });
  }
  for (auto &t : th) t.join();
}

static bool is_pinned_host(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}
static int ensure_pin(int b, size_t bytes) {
  if (bytes <= g.pin_cap[b]) return 0;
  if (g.pin[b]) cudaFreeHost(g.pin[b]);
  g.pin[b] = nullptr;
  g.pin_cap[b] = 0;
  cudaError_t e = cudaHostAlloc(&g.pin[b], bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(SDEB200_E_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
  g.pin_cap[b] = bytes;
  return 0;
}

// Generic slab pipeline for path-major outputs: kernel(slab i) on the compute stream overlaps the device->host copy of
// slab i-1 on the copy stream.  bytes_per_path of contiguous host memory per path.  PINNED host memory (CUDA-allocated or
// registered) is the copy's destination itself.  PAGEABLE memory — what a Julia `Array` or a numpy array is — would make
// cudaMemcpyAsync a staged, host-blocking copy and lose the overlap, so the copy goes into the library's own pinned ring
// and the host moves slab i-2 from the ring into the caller's array (par_memcpy) while the GPU computes slab i-1.
template <class LaunchFn>
static int run_slabs(int64_t npaths, size_t bytes_per_path, void *host_out, LaunchFn &&fn) {
  cudaStream_t st = g.stream();
  int64_t slab = (int64_t)((size_t)(256u << 20) / std::max<size_t>(bytes_per_path, 1));
  slab = std::max<int64_t>(SDE_SHARD_ALIGN, slab / SDE_SHARD_ALIGN * SDE_SHARD_ALIGN);
  slab = std::min<int64_t>(slab, cdiv(npaths, SDE_SHARD_ALIGN) * SDE_SHARD_ALIGN);
  const int nb = npaths > slab ? 2 : 1;
  const bool staged = !is_pinned_host(host_out);
  for (int b = 0; b < nb; b++) {
    int rc = g.slab[b].ensure((size_t)slab * bytes_per_path);
    if (rc) return rc;
    if (staged && (rc = ensure_pin(b, (size_t)slab * bytes_per_path))) return rc;
  }
  bool used[2] = {false, false};
  int64_t first[2] = {0, 0}, count[2] = {0, 0};
  auto drain = [&](int b) -> int {   // the copy engine released buffer b; staged: hand its slab over to the caller's array
    CU(cudaEventSynchronize(g.cdone[b]));
    if (staged) par_memcpy((char *)host_out + (size_t)first[b] * bytes_per_path, g.pin[b], (size_t)count[b] * bytes_per_path);
    used[b] = false;
    return 0;
  };
  int b = 0;
  for (int64_t s0 = 0; s0 < npaths; s0 += slab, b ^= 1) {
    const int64_t n = std::min<int64_t>(slab, npaths - s0);
    int rc;
    if (used[b] && (rc = drain(b))) return rc;
    if ((rc = fn(s0, n, g.slab[b].p, st))) return rc;
    CU(cudaEventRecord(g.kdone[b], st));
    CU(cudaStreamWaitEvent(g.copy, g.kdone[b], 0));
    void *dst = staged ? g.pin[b] : (void *)((char *)host_out + (size_t)s0 * bytes_per_path);
    CU(cudaMemcpyAsync(dst, g.slab[b].p, (size_t)n * bytes_per_path, cudaMemcpyDeviceToHost, g.copy));
    CU(cudaEventRecord(g.cdone[b], g.copy));
    used[b] = true;
    first[b] = s0;
    count[b] = n;
  }
  for (int k = 0; k < 2; k++, b ^= 1) {   // oldest first
    int rc;
    if (used[b] && (rc = drain(b))) return rc;
  }
  CU(cudaStreamSynchronize(g.copy));
  return 0;
}

static int timed_begin() {
  CU(cudaEventRecord(g.ev0, g.stream()));
  return 0;
}
static int timed_end() {
  cudaStream_t st = g.stream();
  CU(cudaEventRecord(g.ev1, st));
  CU(cudaEventSynchronize(g.ev1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, g.ev0, g.ev1));
  g.last_ms = ms;
  CU(cudaGetLastError());
  return 0;
}

static int read_nonfinite(unsigned long long *dev, unsigned long long *host) {
  CU(cudaMemcpyAsync(host, dev, sizeof *host, cudaMemcpyDeviceToHost, g.stream()));
  CU(cudaStreamSynchronize(g.stream()));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// drift / diffusion on the device
// ---------------------------------------------------------------------------------------------------
extern "C" int sdeb200_model_coefficients(sdeb200_model *m, const double *params, double t, const double *x,
                                          double *mu, double *sigma) {
  if (!m || !x) return fail(SDEB200_E_ARG, "null argument");
  int rc = ensure_up();
  if (rc) return rc;
  sdeb200_opts o;
  memset(&o, 0, sizeof o);
  o.dt = 1.0;
  o.t0 = t;
  o.math = SDEB200_MATH_STRICT;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, &o, &kt, FAM_MISC))) return rc;
  const int MW = m->M > 0 ? m->M : 1, nout = m->D + m->D * MW;
  if ((rc = g.misc.ensure(256 * 8))) return rc;
  sde_args a;
  fill_args(&a, m, &o, params, x);
  a.out = g.misc.p;
  cudaStream_t st = g.stream();
  void *args[1] = {&a};
  if (!kt->coef) return fail(SDEB200_E_UNSUPPORTED, "no coefficient kernel");
  CU(cudaLaunchKernel(kt->coef, dim3(1), dim3(32), args, 0, st));
  g.launches++;
  double host[SDE_MAXD + SDE_MAXD * SDE_MAXM];
  CU(cudaMemcpyAsync(host, g.misc.p, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (mu) memcpy(mu, host, m->D * sizeof(double));
  if (sigma) memcpy(sigma, host + m->D, m->D * MW * sizeof(double));
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// sample
// ---------------------------------------------------------------------------------------------------
static int64_t sample_grid(int64_t npaths) { return cdiv(npaths, (int64_t)SDE_BLOCK * SDE_PPT); }
// Terminal-value launch: chunks of SDE_BLOCK*SDE_PPT = 512 paths either way; launches too small to fill the GPU with
// SDE_PPT paths per thread (148 SMs x ~32 warps x 32 lanes x 4 paths ~ 600 k) use the one-path-per-thread form,
// which has the same per-path arithmetic and the same summation tree (bit-identical outputs and partials).
// Bit-identical means: STRICT math (no FMA contraction anywhere) or a built-in functor (FAST forms written with
// explicit FMAs only); generated functors in FAST math leave contraction to the compiler, which may choose
// differently in the two instantiations, so they always take the SDE_PPT form and stay shard-invariant.
static int launch_sample(const sdeb200_model *m, const sde_ktable *kt, const sdeb200_opts *o, const sde_args *a, cudaStream_t st) {
  const void *k1 = kt->sample1[o->scheme][kslot(o)];
  const int64_t grid = capped(sample_grid(a->npaths));
  const bool invariant = o->math == SDEB200_MATH_STRICT || m->builtin >= 0;
  // stream v2's fine transform tables live in dynamic shared memory (sample_body)
  const size_t smem = (kslot(o) == 2 && SDE_V2_FINE) ? (size_t)SDE_TAB2_DOUBLES * 8 : 0;
  if (k1 && invariant && a->npaths <= g_sample1_max) return launch(k1, grid, a, st, smem, SDE_BLOCK * SDE_PPT);
  // the FP64 kernel of stream v2 carries SDE_PPT_V2 paths per thread in correspondingly larger CTAs (same chunks)
  return launch(kt->sample[o->scheme][kslot(o)], grid, a, st, smem, SDE_SAMPLE_THREADS(kslot(o) == 2));
}

extern "C" int sdeb200_sample(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                              int64_t nsteps, int64_t npaths, void *out) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if (nsteps < 0 || npaths < 0 || (!out && npaths > 0)) return fail(SDEB200_E_ARG, "bad nsteps/npaths/out");
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_TERMINAL))) return rc;
  if (npaths == 0) return SDEB200_OK;
  const size_t bpp = (size_t)m->D * esize(o->precision);
  if (sharded_call(npaths) && is_host_or_null(out))   // one process, several GPUs: contiguous shards, one worker per device
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      return sdeb200_sample(m, &os, params, x0, nsteps, n, (char *)out + (size_t)lo * bpp);
    });
  if ((rc = g.red.ensure(RED_WORDS * 8))) return rc;
  cudaStream_t st = g.stream();
  unsigned long long *d_bad = (unsigned long long *)g.red.p;
  CU(cudaMemsetAsync(d_bad, 0, 8, st));
  sde_args a;
  fill_args(&a, m, o, params, x0);
  a.nsteps = nsteps;
  a.nonfinite = d_bad;
  if ((rc = timed_begin())) return rc;
  if (is_device_ptr(out)) {
    a.npaths = npaths;
    a.out = out;
    if ((rc = launch_sample(m, kt, o, &a, st))) return rc;
  } else {
    rc = run_slabs(npaths, bpp, out, [&](int64_t s0, int64_t n, void *dev, cudaStream_t s) {
      sde_args b = a;
      b.npaths = n;
      b.path0 = o->path_offset + s0;
      b.out = dev;
      return launch_sample(m, kt, o, &b, s);
    });
    if (rc) return rc;
  }
  if ((rc = timed_end())) return rc;
  unsigned long long bad = 0;
  if ((rc = read_nonfinite(d_bad, &bad))) return rc;
  if (bad)
    return fail(SDEB200_E_DOMAIN, "%llu of %lld paths ended non-finite (the reference raises DomainError, e.g. sqrt of a "
                "negative variance)", bad, (long long)npaths);
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// simulate (RNG-driven full trajectories)
// ---------------------------------------------------------------------------------------------------
extern "C" int sdeb200_simulate(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                int64_t nsteps, int64_t npaths, int layout, void *out) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if (nsteps < 0 || npaths < 0 || (!out && npaths > 0)) return fail(SDEB200_E_ARG, "bad nsteps/npaths/out");
  if (layout != SDEB200_LAYOUT_REFERENCE && layout != SDEB200_LAYOUT_SOA) return fail(SDEB200_E_ARG, "bad layout");
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_TRAJ))) return rc;
  if (npaths == 0) return SDEB200_OK;
  const size_t es = esize(o->precision);
  const int64_t nrows = nsteps + 1;
  if (layout == SDEB200_LAYOUT_REFERENCE && sharded_call(npaths) && is_host_or_null(out))
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      return sdeb200_simulate(m, &os, params, x0, nsteps, n, layout, (char *)out + (size_t)lo * nrows * m->D * es);
    });
  cudaStream_t st = g.stream();
  if ((rc = g.red.ensure(RED_WORDS * 8))) return rc;
  unsigned long long *d_bad = (unsigned long long *)g.red.p;
  CU(cudaMemsetAsync(d_bad, 0, 8, st));
  sde_args a;
  fill_args(&a, m, o, params, x0);
  a.nsteps = nsteps;
  a.nonfinite = d_bad;
  const bool dev = is_device_ptr(out);
  if ((rc = timed_begin())) return rc;
  if (layout == SDEB200_LAYOUT_REFERENCE) {
    const void *k = kt->simref[o->scheme][kslot(o)], *k_tma = kt->simrt[o->scheme][kslot(o)];
    const size_t bpp = (size_t)nrows * m->D * es;
    if (dev) {
      if ((rc = launch_simref(k, k_tma, m->D, m->M, kslot(o), a, npaths, out, st))) return rc;
    } else {
      rc = run_slabs(npaths, bpp, out, [&](int64_t s0, int64_t n, void *d, cudaStream_t s) {
        sde_args b = a;
        b.path0 = o->path_offset + s0;
        return launch_simref(k, k_tma, m->D, m->M, kslot(o), b, n, d, s);
      });
      if (rc) return rc;
    }
  } else {
    const void *k = kt->simsoa[o->scheme][kslot(o)];
    const int vec = SDE_SOA_VEC;
    if (dev) {
      a.npaths = npaths;
      a.out = out;
      a.ld_paths = npaths;
      a.path_off = 0;
      if ((rc = launch(k, cdiv(cdiv(npaths, vec), SDE_BLOCK), &a, st))) return rc;
    } else {
      // host SoA: slabs of columns, strided in host memory -> 2-D copies
      const size_t rowsD = (size_t)nrows * m->D;
      int64_t slab = (int64_t)((size_t)(256u << 20) / (rowsD * es));
      slab = std::max<int64_t>(SDE_SHARD_ALIGN, slab / SDE_SHARD_ALIGN * SDE_SHARD_ALIGN);
      slab = std::min<int64_t>(slab, cdiv(npaths, SDE_SHARD_ALIGN) * SDE_SHARD_ALIGN);
      if ((rc = g.slab[0].ensure((size_t)slab * rowsD * es))) return rc;
      for (int64_t s0 = 0; s0 < npaths; s0 += slab) {
        const int64_t n = std::min<int64_t>(slab, npaths - s0);
        sde_args b = a;
        b.npaths = n;
        b.path0 = o->path_offset + s0;
        b.out = g.slab[0].p;
        b.ld_paths = slab;
        b.path_off = 0;
        if ((rc = launch(k, cdiv(cdiv(n, vec), SDE_BLOCK), &b, st))) return rc;
        CU(cudaMemcpy2DAsync((char *)out + (size_t)s0 * es, (size_t)npaths * es, g.slab[0].p, (size_t)slab * es,
                             (size_t)n * es, rowsD, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
      }
    }
  }
  if ((rc = timed_end())) return rc;
  unsigned long long bad = 0;
  if ((rc = read_nonfinite(d_bad, &bad))) return rc;
  if (bad)
    return fail(SDEB200_E_DOMAIN, "%llu of %lld trajectories ended non-finite (the reference raises DomainError, e.g. sqrt "
                "of a negative variance); the trajectories are written", bad, (long long)npaths);
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// simulate! driven by a Wiener path
// ---------------------------------------------------------------------------------------------------
// One Wiener-driven launch over device buffers: the TMA kernel where its tiles apply (D == M, D*sizeof(T) | 128,
// 16-byte aligned buffers, paths of at least two tiles), the plain tile kernel for everything else and for the
// npaths % g leftover paths.  `a` carries everything but npaths / out / out2.
static int launch_simw(const sde_ktable *kt, const sdeb200_model *m, const sdeb200_opts *o, sde_args a, int64_t nrows,
                       int64_t npaths, const void *w, void *x, cudaStream_t st) {
  const int MW = m->M > 0 ? m->M : 1, D = m->D;
  const size_t es = esize(o->precision);
  const int rowb = (int)(D * es);
  const void *kt_tma = kt->simwt[o->scheme][kslot(o)], *k = kt->simw[o->scheme][kslot(o)];
  int rc;
  int64_t done = 0;
  if (kt_tma && g_tma_enabled && D == MW && 128 % rowb == 0 && ((uintptr_t)w % SDE_SIMWT_ALIGN) == 0 && ((uintptr_t)x % SDE_SIMWT_ALIGN) == 0 &&
      nrows >= 2 * (128 / rowb) + 4 && nrows * D < (1 << 28) && npaths < ((int64_t)1 << 31)) {
    const int grp = sde_tma_group((int64_t)nrows * rowb, SDE_SIMWT_ALIGN);
    const int64_t ncov = npaths - npaths % grp;
    if (ncov >= 32) {
      a.npaths = ncov;
      a.out = x;
      a.out2 = (void *)w;
      if ((rc = make_traj_tmap(&a.tm_w, w, o->precision, D, nrows, ncov, grp))) return rc;
      if ((rc = make_traj_tmap(&a.tm_x, x, o->precision, D, nrows, ncov, grp))) return rc;
      if ((rc = launch(kt_tma, cdiv(ncov, SDE_BLOCK), &a, st, simwt_smem()))) return rc;
      g.tma_launches++;
      done = ncov;
    }
  }
  if (done < npaths) {
    a.npaths = npaths - done;
    a.out = (char *)x + (size_t)done * nrows * D * es;
    a.out2 = (char *)w + (size_t)done * nrows * MW * es;
    if ((rc = launch(k, cdiv(npaths - done, SDE_BLOCK), &a, st, simw_smem(m->D, m->M, o->precision)))) return rc;
  }
  return 0;
}

extern "C" int sdeb200_simulate_wiener(sdeb200_model *m, const sdeb200_opts *o, const double *params,
                                       const double *x0, int64_t nrows, int64_t npaths, const void *w, void *x) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if (nrows < 1 || npaths < 0 || ((!w || !x) && npaths > 0)) return fail(SDEB200_E_ARG, "bad nrows/npaths/w/x");
  const int MW = m->M > 0 ? m->M : 1;
  if (w == x && m->D != MW)
    return fail(SDEB200_E_ARG, "x === w needs eltype(x) == eltype(w), i.e. D == M (D=%d, M=%d)", m->D, m->M);
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_WIENER))) return rc;
  if (npaths == 0) return SDEB200_OK;
  const size_t es = esize(o->precision);
  if (sharded_call(npaths) && is_host_or_null(w) && is_host_or_null(x))
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      return sdeb200_simulate_wiener(m, o, params, x0, nrows, n, (const char *)w + (size_t)lo * nrows * MW * es,
                                     (char *)x + (size_t)lo * nrows * m->D * es);
    });
  cudaStream_t st = g.stream();
  if ((rc = g.red.ensure(RED_WORDS * 8))) return rc;
  unsigned long long *d_bad = (unsigned long long *)g.red.p;
  CU(cudaMemsetAsync(d_bad, 0, 8, st));
  sde_args a;
  fill_args(&a, m, o, params, x0);
  a.nsteps = nrows - 1;
  a.nonfinite = d_bad;
  const bool wdev = is_device_ptr(w), xdev = is_device_ptr(x);
  if ((rc = timed_begin())) return rc;
  if (wdev && xdev) {
    if ((rc = launch_simw(kt, m, o, a, nrows, npaths, w, x, st))) return rc;
  } else {
    const size_t wb = (size_t)nrows * MW * es, xb = (size_t)nrows * m->D * es;
    int64_t slab = (int64_t)((size_t)(128u << 20) / std::max(wb, xb));
    slab = std::max<int64_t>(32, std::min<int64_t>(slab, npaths));
    // two slabs in flight: host -> device copy and kernel of slab i + 1 (compute stream) overlap the device -> host
    // copy of slab i (copy stream); PCIe is full duplex, so host-resident simulate!(x, ..., w) moves both ways at once
    const int nb = npaths > slab ? 2 : 1;
    for (int b = 0; b < nb; b++) {
      if ((rc = g.win[b].ensure((size_t)slab * wb))) return rc;
      if ((rc = g.slab[b].ensure((size_t)slab * xb))) return rc;
    }
    bool used[2] = {false, false};
    int b = 0;
    for (int64_t s0 = 0; s0 < npaths; s0 += slab, b ^= 1) {
      const int64_t n = std::min<int64_t>(slab, npaths - s0);
      const void *wsrc = (const char *)w + (size_t)s0 * wb;
      if (used[b]) CU(cudaStreamWaitEvent(st, g.cdone[b], 0));   // slab[b] was read by the copy of slab i - 2
      CU(cudaMemcpyAsync(g.win[b].p, wsrc, (size_t)n * wb, wdev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
      if ((rc = launch_simw(kt, m, o, a, nrows, n, g.win[b].p, g.slab[b].p, st))) return rc;
      CU(cudaEventRecord(g.kdone[b], st));
      CU(cudaStreamWaitEvent(g.copy, g.kdone[b], 0));
      CU(cudaMemcpyAsync((char *)x + (size_t)s0 * xb, g.slab[b].p, (size_t)n * xb,
                         xdev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, g.copy));
      CU(cudaEventRecord(g.cdone[b], g.copy));
      used[b] = true;
    }
    CU(cudaStreamSynchronize(g.copy));
    CU(cudaStreamSynchronize(st));
  }
  if ((rc = timed_end())) return rc;
  unsigned long long bad = 0;
  if ((rc = read_nonfinite(d_bad, &bad))) return rc;
  if (bad)
    return fail(SDEB200_E_DOMAIN, "%llu of %lld trajectories ended non-finite (the reference raises DomainError, e.g. sqrt "
                "of a negative variance); the trajectories are written", bad, (long long)npaths);
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// logtransition along stored trajectories
// ---------------------------------------------------------------------------------------------------
static size_t logt_smem(int D, int precision) {
  const int ts = (SDE_SIMW_ROWBYTES / (int)esize(precision)) / D;
  const int xld = (ts * D) | 1, old = ts | 1;
  return (size_t)(SDE_BLOCK / 32) * 32 * (xld + old) * esize(precision);
}

static size_t logtt_smem() { return (size_t)(SDE_BLOCK / 32) * SDE_LOGTT_STAGES * (4096 + 8) + 1024; }
// One launch over device buffers: the TMA-load kernel K7t where its tiles apply (D*sizeof(T) | 128, 16-byte aligned input,
// paths of at least two tiles), the plain tile kernel K7 for everything else and for the npaths % g leftover paths.
static int launch_logt(const sde_ktable *kt, const sdeb200_model *m, int precision, sde_args a, int64_t nrows, int64_t npaths,
                       const void *x, void *out, cudaStream_t st) {
  const size_t es = esize(precision);
  const int D = m->D, rowb = (int)(D * es);
  const void *k = kt->logt[precision], *k_tma = kt->logtt[precision];
  int rc;
  int64_t done = 0;
  if (k_tma && g_tma_enabled && 128 % rowb == 0 && ((uintptr_t)x % SDE_SIMWT_ALIGN) == 0 && nrows >= 2 * (128 / rowb) + 4 &&
      nrows * D < (1 << 28) && npaths < ((int64_t)1 << 31)) {
    const int grp = sde_tma_group((int64_t)nrows * rowb, SDE_SIMWT_ALIGN);
    const int64_t ncov = npaths - npaths % grp;
    if (ncov >= 32) {
      a.npaths = ncov;
      a.out = out;
      a.out2 = (void *)x;
      if ((rc = make_traj_tmap(&a.tm_w, x, precision, D, nrows, ncov, grp))) return rc;
      if ((rc = launch(k_tma, cdiv(ncov, SDE_BLOCK), &a, st, logtt_smem()))) return rc;
      g.tma_launches++;
      done = ncov;
    }
  }
  if (done < npaths) {
    a.npaths = npaths - done;
    a.out = (char *)out + (size_t)done * (nrows - 1) * es;
    a.out2 = (char *)x + (size_t)done * nrows * D * es;
    if ((rc = launch(k, cdiv(npaths - done, SDE_BLOCK), &a, st, logt_smem(D, precision)))) return rc;
  }
  return 0;
}

extern "C" int sdeb200_logtransition(sdeb200_model *m, const sdeb200_opts *o, const double *params, int64_t nrows,
                                     int64_t npaths, const void *x, void *out) {
  if (!m || !o) return fail(SDEB200_E_ARG, "null argument");
  if (!params && !m->params.empty()) return fail(SDEB200_E_ARG, "params is NULL");
  if (o->scheme != SDEB200_EULER_MARUYAMA)
    return fail(SDEB200_E_UNSUPPORTED, "logtransition is implemented for EulerMaruyama (src/schemes/euler_maruyama.jl:8-13)");
  if (o->precision < 0 || o->precision > 1 || !(o->dt > 0.0)) return fail(SDEB200_E_ARG, "bad precision / Δt");
  if (nrows < 1 || npaths < 0 || ((!x || !out) && npaths > 0 && nrows > 1)) return fail(SDEB200_E_ARG, "bad nrows/npaths/x/out");
  int rc = ensure_up();
  if (rc) return rc;
  sdeb200_opts oc = *o;
  oc.math = SDEB200_MATH_FAST;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, &oc, &kt, FAM_MISC))) return rc;
  if (npaths == 0 || nrows == 1) return SDEB200_OK;
  const size_t es = esize(o->precision);
  const size_t xb = (size_t)nrows * m->D * es, ob = (size_t)(nrows - 1) * es;
  if (sharded_call(npaths) && is_host_or_null(x) && is_host_or_null(out))
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      return sdeb200_logtransition(m, o, params, nrows, n, (const char *)x + (size_t)lo * xb, (char *)out + (size_t)lo * ob);
    });
  cudaStream_t st = g.stream();
  double x0[SDE_MAXD] = {0, 0, 0, 0};
  sde_args a;
  fill_args(&a, m, &oc, params, x0);
  a.nsteps = nrows - 1;
  const bool xdev = is_device_ptr(x), odev = is_device_ptr(out);
  if ((rc = timed_begin())) return rc;
  if (xdev && odev) {
    if ((rc = launch_logt(kt, m, o->precision, a, nrows, npaths, x, out, st))) return rc;
  } else {
    int64_t slab = (int64_t)((size_t)(128u << 20) / xb);
    slab = std::max<int64_t>(32, std::min<int64_t>(slab, npaths));
    // two slabs in flight, like the host path of simulate!(x, ..., w): the copy back of slab i overlaps slab i + 1
    const int nb = npaths > slab ? 2 : 1;
    for (int i = 0; i < nb; i++) {
      if ((rc = g.win[i].ensure((size_t)slab * xb))) return rc;
      if ((rc = g.slab[i].ensure((size_t)slab * ob))) return rc;
    }
    bool used[2] = {false, false};
    int bi = 0;
    for (int64_t s0 = 0; s0 < npaths; s0 += slab, bi ^= 1) {
      const int64_t n = std::min<int64_t>(slab, npaths - s0);
      if (used[bi]) CU(cudaStreamWaitEvent(st, g.cdone[bi], 0));
      CU(cudaMemcpyAsync(g.win[bi].p, (const char *)x + (size_t)s0 * xb, (size_t)n * xb,
                         xdev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
      if ((rc = launch_logt(kt, m, o->precision, a, nrows, n, g.win[bi].p, g.slab[bi].p, st))) return rc;
      CU(cudaEventRecord(g.kdone[bi], st));
      CU(cudaStreamWaitEvent(g.copy, g.kdone[bi], 0));
      CU(cudaMemcpyAsync((char *)out + (size_t)s0 * ob, g.slab[bi].p, (size_t)n * ob,
                         odev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, g.copy));
      CU(cudaEventRecord(g.cdone[bi], g.copy));
      used[bi] = true;
    }
    CU(cudaStreamSynchronize(g.copy));
    CU(cudaStreamSynchronize(st));
  }
  return timed_end();
}

// ---------------------------------------------------------------------------------------------------
// wiener!
// ---------------------------------------------------------------------------------------------------
static int wiener_device(const sdeb200_opts *o, int dim, int64_t nrows, int64_t npaths, int64_t path0, void *dev,
                         cudaStream_t st) {
  const int slot = kslot(o);
  const void *k = o->math == SDEB200_MATH_STRICT ? sde_aot_wiener_strict(dim, slot) : sde_aot_wiener_fast(dim, slot);
  if (!k) return fail(SDEB200_E_ARG, "wiener!: dim must be 1..4");
  sde_args a;
  memset(&a, 0, sizeof a);
  a.t0 = o->t0;
  set_dt(&a, o->dt, SDEB200_EULER_MARUYAMA, o->rng);
  a.seed_lo = (uint32_t)o->seed;
  a.seed_hi = (uint32_t)(o->seed >> 32);
  a.stream = o->stream;
  a.path0 = path0;
  a.nsteps = nrows - 1;
  const void *k_tma = o->math == SDEB200_MATH_STRICT ? sde_aot_wiener_strict(16 + dim, slot) : sde_aot_wiener_fast(16 + dim, slot);
  return launch_simref(k, k_tma, dim, dim, slot, a, npaths, dev, st);
}

extern "C" int sdeb200_wiener(const sdeb200_opts *o, int dim, int64_t nrows, int64_t npaths, void *w) {
  if (!o || (!w && npaths > 0)) return fail(SDEB200_E_ARG, "null argument");
  if (nrows < 1 || npaths < 0 || dim < 1 || dim > 4) return fail(SDEB200_E_ARG, "bad nrows/npaths/dim");
  if (!(o->dt > 0.0)) return fail(SDEB200_E_ARG, "Δt must be positive");
  if (o->rng < 0 || o->rng > SDEB200_RNG_V2 || o->precision < 0 || o->precision > 1 || o->math < 0 || o->math > 1)
    return fail(SDEB200_E_ARG, "bad precision / math / rng");
  int rc = ensure_up();
  if (rc) return rc;
  if (npaths == 0) return SDEB200_OK;
  const size_t bpp = (size_t)nrows * dim * esize(o->precision);
  if (sharded_call(npaths) && is_host_or_null(w))
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      return sdeb200_wiener(&os, dim, nrows, n, (char *)w + (size_t)lo * bpp);
    });
  cudaStream_t st = g.stream();
  if ((rc = timed_begin())) return rc;
  if (is_device_ptr(w)) {
    if ((rc = wiener_device(o, dim, nrows, npaths, o->path_offset, w, st))) return rc;
  } else {
    rc = run_slabs(npaths, bpp, w, [&](int64_t s0, int64_t n, void *d, cudaStream_t s) {
      return wiener_device(o, dim, nrows, n, o->path_offset + s0, d, s);
    });
    if (rc) return rc;
  }
  return timed_end();
}

// ---------------------------------------------------------------------------------------------------
// exact accumulator, host half
// ---------------------------------------------------------------------------------------------------
extern "C" void sdeb200_bins_add(int64_t *bins, double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  const int e = (int)((b >> 52) & 0x7ff);
  uint64_t mant = b & 0xFFFFFFFFFFFFFull;
  if (e == 0x7ff || (e == 0 && mant == 0)) return;
  int sh = e - 1;
  if (e == 0) sh = 0; else mant |= (1ull << 52);
  const int bin = sh >> 5, off = sh & 31;
  const uint64_t lo64 = mant << off, hi = off ? (mant >> (64 - off)) : 0ull;
  int64_t l0 = (int64_t)(lo64 & 0xffffffffull), l1 = (int64_t)(lo64 >> 32), l2 = (int64_t)hi;
  if (b >> 63) { l0 = -l0; l1 = -l1; l2 = -l2; }
  bins[bin] += l0;
  bins[bin + 1] += l1;
  bins[bin + 2] += l2;
}

extern "C" double sdeb200_bins_to_double(const int64_t *bins) {
  uint32_t limb[SDE_NBINS];
  __int128 carry = 0;
  for (int b = 0; b < SDE_NBINS; b++) {
    __int128 t = (__int128)bins[b] + carry;
    limb[b] = (uint32_t)((unsigned __int128)t & 0xffffffffu);
    carry = t >> 32;  // arithmetic shift: floor division
  }
  bool neg = false;
  if (carry < 0) {
    if (carry != -1) return -INFINITY;
    neg = true;
    uint64_t c = 1;  // two's complement over the limbs
    for (int b = 0; b < SDE_NBINS; b++) {
      uint64_t t = (uint64_t)(uint32_t)~limb[b] + c;
      limb[b] = (uint32_t)t;
      c = t >> 32;
    }
  } else if (carry > 0) {
    return INFINITY;
  }
  int hi = SDE_NBINS - 1;
  while (hi >= 0 && limb[hi] == 0) hi--;
  if (hi < 0) return 0.0;
  double r;
  if (hi <= 1) {
    const uint64_t v = ((uint64_t)(hi == 1 ? limb[1] : 0) << 32) | limb[0];
    // <= 64 bits: exact conversion needs care above 53 bits
    if (v < (1ull << 53)) {
      r = ldexp((double)v, -1074);
    } else {
      const int L = 64 - __builtin_clzll(v), shf = L - 53;
      uint64_t mant = v >> shf, rem = v & ((1ull << shf) - 1), half = 1ull << (shf - 1);
      if (rem > half || (rem == half && (mant & 1))) mant++;
      r = ldexp((double)mant, shf - 1074);
    }
  } else {
    unsigned __int128 t = ((unsigned __int128)limb[hi] << 64) | ((unsigned __int128)limb[hi - 1] << 32) | limb[hi - 2];
    bool sticky = false;
    for (int b = hi - 3; b >= 0 && !sticky; b--) sticky = limb[b] != 0;
    int L = 96;
    while (!((t >> (L - 1)) & 1)) L--;  // L >= 65 because limb[hi] != 0
    const int shf = L - 53;
    uint64_t mant = (uint64_t)(t >> shf);
    const unsigned __int128 rem = t & (((unsigned __int128)1 << shf) - 1), half = (unsigned __int128)1 << (shf - 1);
    if (rem > half || (rem == half && (sticky || (mant & 1)))) mant++;
    r = ldexp((double)mant, shf + 32 * (hi - 2) - 1074);
  }
  return neg ? -r : r;
}

// ---------------------------------------------------------------------------------------------------
// estimators
// ---------------------------------------------------------------------------------------------------
struct RedSlot {  // device layout of one reduction record
  unsigned long long *bins, *n, *bad, *ovf;
};
static RedSlot red_slot(void *base, int idx) {
  unsigned long long *p = (unsigned long long *)base + (size_t)idx * RED_WORDS;
  RedSlot r;
  r.bins = p;
  r.n = p + 2 * SDE_MAXF * SDE_NBINS;
  r.bad = r.n + 1;
  r.ovf = r.n + 2;
  return r;
}

static int payoff_nf(const sdeb200_model *m, const sdeb200_payoff *p) {
  return p->kind == SDEB200_PAYOFF_MOMENTS ? m->D : 1;
}
static int check_payoff(const sdeb200_model *m, const sdeb200_payoff *p) {
  if (!p) return fail(SDEB200_E_ARG, "payoff is NULL");
  if (p->kind < 0 || p->kind > 3) return fail(SDEB200_E_ARG, "unknown payoff kind %d", p->kind);
  if (p->kind != SDEB200_PAYOFF_MOMENTS && (p->comp < 0 || p->comp >= m->D))
    return fail(SDEB200_E_ARG, "payoff component %d out of range (D = %d)", p->comp, m->D);
  return 0;
}

// finish nrec reduction records: (allreduce) -> host -> doubles.  nq[i] quantities per record.
// raw integer records -> doubles: res[0..nq) = the exactly accumulated sums rounded once, then n and n_nonfinite
static int round_records(int nrec, const int *nq, const unsigned long long *host, double *result, int result_stride,
                         unsigned long long *bad_total) {
  *bad_total = 0;
  for (int i = 0; i < nrec; i++) {
    const unsigned long long *h = host + (size_t)i * RED_WORDS;
    const unsigned long long *ex = h + 2 * SDE_MAXF * SDE_NBINS;
    double *res = result + (size_t)i * result_stride;
    for (int q = 0; q < nq[i]; q++) res[q] = sdeb200_bins_to_double((const int64_t *)h + q * SDE_NBINS);
    res[nq[i]] = (double)ex[0];
    res[nq[i] + 1] = (double)ex[1];
    *bad_total += ex[1];
    if (ex[2]) return fail(SDEB200_E_DOMAIN, "a partial payoff sum overflowed to inf/nan");
  }
  return 0;
}

// finish nrec reduction records: (allreduce) -> host -> doubles.  nq[i] quantities per record.  A worker of a sharded
// call (tl_raw) also leaves the raw integer records for the parent, which adds the devices' records (integer addition:
// exact, order-free) and rounds once — the same bits as one GPU.
static int finish_reductions(int nrec, const int *nq, cudaStream_t *streams, double *result, int result_stride,
                             unsigned long long *bad_total) {
  std::vector<unsigned long long> host((size_t)nrec * RED_WORDS);
  if (g.comm) {
    NC(g_nccl.GroupStart());
    for (int i = 0; i < nrec; i++) {
      RedSlot r = red_slot(g.red.p, i);
      NC(g_nccl.AllReduce(r.bins, r.bins, RED_WORDS, /*ncclInt64*/ 4, /*ncclSum*/ 0, g.comm, streams[i]));
    }
    NC(g_nccl.GroupEnd());
  }
  for (int i = 0; i < nrec; i++) {
    CU(cudaMemcpyAsync(host.data() + (size_t)i * RED_WORDS, red_slot(g.red.p, i).bins, RED_WORDS * 8,
                       cudaMemcpyDeviceToHost, streams[i]));
  }
  for (int i = 0; i < nrec; i++) CU(cudaStreamSynchronize(streams[i]));
  if (tl_raw) memcpy(tl_raw, host.data(), host.size() * 8);
  return round_records(nrec, nq, host.data(), result, result_stride, bad_total);
}

// parent of a sharded estimator: add the workers' raw records word by word
static void add_records(std::vector<unsigned long long> &total, const std::vector<std::vector<unsigned long long>> &raw) {
  for (const auto &r : raw)
    for (size_t i = 0; i < total.size(); i++) total[i] += r[i];   // two's-complement wrap-around == int64 addition
}

extern "C" int sdeb200_estimate(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                int64_t nsteps, int64_t npaths, const sdeb200_payoff *payoff, void *terminal,
                                double *result) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if ((rc = check_payoff(m, payoff))) return rc;
  if (nsteps < 0 || npaths < 0 || !result) return fail(SDEB200_E_ARG, "bad nsteps/npaths/result");
  if (terminal && !is_device_ptr(terminal) && npaths > 0)
    return fail(SDEB200_E_ARG, "estimate: `terminal` must be device memory (use sdeb200_sample for host arrays)");
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_TERMINAL))) return rc;
  const int nf = payoff_nf(m, payoff), nq = 2 * nf;
  if (sharded_call(npaths) && !terminal) {
    std::vector<std::vector<unsigned long long>> raw(g_ndev, std::vector<unsigned long long>(RED_WORDS, 0));
    rc = run_sharded(npaths, [&](int d, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      double tmp[2 * SDE_MAXF + 2];
      tl_raw = raw[d].data();
      const int r = sdeb200_estimate(m, &os, params, x0, nsteps, n, payoff, nullptr, tmp);
      tl_raw = nullptr;
      return r;
    });
    if (rc && rc != SDEB200_E_DOMAIN) return rc;
    std::vector<unsigned long long> total(RED_WORDS, 0);
    add_records(total, raw);
    unsigned long long bad = 0;
    if ((rc = round_records(1, &nq, total.data(), result, nq + 2, &bad))) return rc;
    if (bad) return fail(SDEB200_E_DOMAIN, "%llu paths ended non-finite; they are excluded from the sums", bad);
    return SDEB200_OK;
  }
  cudaStream_t st = g.stream();
  const int64_t grid = sample_grid(npaths);
  if ((rc = g.red.ensure((size_t)RED_WORDS * 8 * MAX_LEVELS))) return rc;
  if ((rc = g.partials[0].ensure((size_t)std::max<int64_t>(grid, 1) * 2 * SDE_MAXF * 8))) return rc;
  RedSlot r = red_slot(g.red.p, 0);
  if ((rc = timed_begin())) return rc;
  CU(cudaMemsetAsync(r.bins, 0, RED_WORDS * 8, st));
  unsigned long long n_local = (unsigned long long)npaths;
  CU(cudaMemcpyAsync(r.n, &n_local, 8, cudaMemcpyHostToDevice, st));
  if (npaths > 0) {
    sde_args a;
    fill_args(&a, m, o, params, x0);
    a.nsteps = nsteps;
    a.npaths = npaths;
    a.out = terminal;
    a.partials = (double *)g.partials[0].p;
    a.nonfinite = r.bad;
    a.payoff_kind = payoff->kind;
    a.payoff_comp = payoff->comp;
    a.payoff_k = payoff->strike;
    a.payoff_disc = payoff->discount;
    if ((rc = launch_sample(m, kt, o, &a, st))) return rc;
    sde_launch_accumulate((const double *)g.partials[0].p, grid, 2 * SDE_MAXF, nq, r.bins, r.ovf, st);
    g.launches++;
  }
  unsigned long long bad = 0;
  if ((rc = finish_reductions(1, &nq, &st, result, nq + 2, &bad))) return rc;
  if ((rc = timed_end())) return rc;
  if (bad) return fail(SDEB200_E_DOMAIN, "%llu paths ended non-finite; they are excluded from the sums", bad);
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// multilevel
// ---------------------------------------------------------------------------------------------------
static int mlmc_args(sde_args *a, sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                     double dt_fine, int64_t substeps, int64_t ncoarse, int64_t npaths) {
  if (substeps < 1 || ncoarse < 0 || npaths < 0) return fail(SDEB200_E_ARG, "bad substeps/ncoarse/npaths");
  if (!(dt_fine > 0.0)) return fail(SDEB200_E_ARG, "fine Δt must be positive");
  const int MW = m->M > 0 ? m->M : 1;
  if (m->M != 0 && m->D != MW)
    return fail(SDEB200_E_UNSUPPORTED,
                "the coupled level sampler needs D == M (Δw = zero(typeof(x0)) in src/multilevel.jl:79); "
                "'%s' has D=%d, M=%d", m->name.c_str(), m->D, m->M);
  fill_args(a, m, o, params, x0);
  set_dt(a, dt_fine, o->scheme, o->rng);
  a->dt_coarse = o->dt;
  a->substeps = substeps;
  a->nsteps = ncoarse * substeps;
  a->npaths = npaths;
  return 0;
}

extern "C" int sdeb200_mlmc_sample_level(sdeb200_model *m, const sdeb200_opts *o, const double *params,
                                         const double *x0, double dt_fine, int64_t substeps, int64_t ncoarse,
                                         int64_t npaths, void *xc, void *xf) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_TERMINAL))) return rc;
  sde_args a;
  if ((rc = mlmc_args(&a, m, o, params, x0, dt_fine, substeps, ncoarse, npaths))) return rc;
  if (npaths == 0) return SDEB200_OK;
  const size_t bytes = (size_t)npaths * m->D * esize(o->precision);
  if (sharded_call(npaths) && is_host_or_null(xc) && is_host_or_null(xf)) {
    const size_t bpp = (size_t)m->D * esize(o->precision);
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      return sdeb200_mlmc_sample_level(m, &os, params, x0, dt_fine, substeps, ncoarse, n, xc ? (char *)xc + lo * bpp : nullptr,
                                       xf ? (char *)xf + lo * bpp : nullptr);
    });
  }
  cudaStream_t st = g.stream();
  const bool cdev = xc && is_device_ptr(xc), fdev = xf && is_device_ptr(xf);
  if ((rc = g.red.ensure((size_t)RED_WORDS * 8 * MAX_LEVELS))) return rc;
  unsigned long long *d_bad = (unsigned long long *)g.red.p;
  CU(cudaMemsetAsync(d_bad, 0, 8, st));
  a.nonfinite = d_bad;
  if (xf) {
    if (fdev) a.out = xf;
    else { if ((rc = g.slab[0].ensure(bytes))) return rc; a.out = g.slab[0].p; }
  }
  if (xc) {
    if (cdev) a.out2 = xc;
    else { if ((rc = g.slab[1].ensure(bytes))) return rc; a.out2 = g.slab[1].p; }
  }
  if ((rc = timed_begin())) return rc;
  if ((rc = launch(kt->mlmc[o->scheme][kslot(o)], capped(cdiv(npaths, SDE_BLOCK)), &a, st))) return rc;
  if (xf && !fdev) CU(cudaMemcpyAsync(xf, a.out, bytes, cudaMemcpyDeviceToHost, st));
  if (xc && !cdev) CU(cudaMemcpyAsync(xc, a.out2, bytes, cudaMemcpyDeviceToHost, st));
  if ((rc = timed_end())) return rc;
  unsigned long long bad = 0;
  if ((rc = read_nonfinite(d_bad, &bad))) return rc;
  if (bad) return fail(SDEB200_E_DOMAIN, "%llu of %lld coupled paths ended non-finite", bad, (long long)npaths);
  return SDEB200_OK;
}

extern "C" int sdeb200_mlmc_simulate_level(sdeb200_model *m, const sdeb200_opts *o, const double *params,
                                           const double *x0, double dt_fine, int64_t substeps, int64_t ncoarse,
                                           int64_t npaths, void *xc, void *xf) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if (!xc || !xf) return fail(SDEB200_E_ARG, "xc/xf are NULL");
  if (m->D != m->M)
    return fail(SDEB200_E_UNSUPPORTED, "_multilevel_simulate needs D == M (wiener! fills an array of eltype(x0), "
                "src/multilevel.jl:94); '%s' has D=%d, M=%d", m->name.c_str(), m->D, m->M);
  if (substeps < 1 || ncoarse < 0 || npaths < 0) return fail(SDEB200_E_ARG, "bad substeps/ncoarse/npaths");
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_WIENER))) return rc;
  if (npaths == 0) return SDEB200_OK;
  const size_t es = esize(o->precision);
  const int64_t nf = ncoarse * substeps + 1, nc = ncoarse + 1;
  if (sharded_call(npaths) && is_host_or_null(xc) && is_host_or_null(xf))
    return run_sharded(npaths, [&](int, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      return sdeb200_mlmc_simulate_level(m, &os, params, x0, dt_fine, substeps, ncoarse, n, (char *)xc + (size_t)lo * nc * m->D * es,
                                         (char *)xf + (size_t)lo * nf * m->D * es);
    });
  cudaStream_t st = g.stream();
  const size_t fb = (size_t)npaths * nf * m->D * es, cb = (size_t)npaths * nc * m->D * es;
  const bool fdev = is_device_ptr(xf), cdev = is_device_ptr(xc);
  void *dxf = xf, *dxc = xc;
  if (!fdev) { if ((rc = g.slab[0].ensure(fb))) return rc; dxf = g.slab[0].p; }
  if (!cdev) { if ((rc = g.slab[1].ensure(cb))) return rc; dxc = g.slab[1].p; }
  if ((rc = timed_begin())) return rc;
  // wiener!(xf, fine.Δt); xc = xf[1:substeps:end, :]; simulate!(xf, ..., xf); simulate!(xc, ..., xc)
  sdeb200_opts of = *o;
  of.dt = dt_fine;
  if ((rc = wiener_device(&of, m->D, nf, npaths, o->path_offset, dxf, st))) return rc;
  sde_launch_subsample(o->precision, dxf, dxc, npaths, nc, nf, substeps, m->D, st);
  g.launches++;
  sde_args a;
  fill_args(&a, m, &of, params, x0);
  a.nsteps = nf - 1;
  if ((rc = launch_simw(kt, m, o, a, nf, npaths, dxf, dxf, st))) return rc;
  fill_args(&a, m, o, params, x0);
  a.nsteps = nc - 1;
  if ((rc = launch_simw(kt, m, o, a, nc, npaths, dxc, dxc, st))) return rc;
  if (!fdev) CU(cudaMemcpyAsync(xf, dxf, fb, cudaMemcpyDeviceToHost, st));
  if (!cdev) CU(cudaMemcpyAsync(xc, dxc, cb, cudaMemcpyDeviceToHost, st));
  return timed_end();
}

static int ensure_level_streams(int n) {
  for (int i = 0; i < n; i++) {
    if (!g.lvl[i]) CU(cudaStreamCreateWithFlags(&g.lvl[i], cudaStreamNonBlocking));
    if (!g.lvl_done[i]) CU(cudaEventCreateWithFlags(&g.lvl_done[i], cudaEventDisableTiming));
  }
  return 0;
}

// queue one level's kernels on `st` using reduction record `slot`
static int queue_level(sdeb200_model *m, const sde_ktable *kt, const sdeb200_opts *o, const double *params,
                       const double *x0, bool coupled, double dt_fine, int64_t substeps, int64_t ncoarse,
                       int64_t npaths, const sdeb200_payoff *payoff, int slot, cudaStream_t st) {
  int rc;
  RedSlot r = red_slot(g.red.p, slot);
  CU(cudaMemsetAsync(r.bins, 0, RED_WORDS * 8, st));
  unsigned long long n_local = (unsigned long long)npaths;
  CU(cudaMemcpyAsync(r.n, &n_local, 8, cudaMemcpyHostToDevice, st));
  if (npaths == 0) return 0;
  sde_args a;
  int64_t grid;
  const void *k;
  if (coupled) {
    if ((rc = mlmc_args(&a, m, o, params, x0, dt_fine, substeps, ncoarse, npaths))) return rc;
    grid = cdiv(npaths, SDE_BLOCK);
    k = kt->mlmc[o->scheme][kslot(o)];
  } else {
    fill_args(&a, m, o, params, x0);
    a.nsteps = ncoarse;
    a.npaths = npaths;
    grid = sample_grid(npaths);
    k = nullptr;   // launch_sample picks the form
  }
  if ((rc = g.partials[slot].ensure((size_t)grid * 2 * SDE_MAXF * 8))) return rc;
  a.partials = (double *)g.partials[slot].p;
  a.nonfinite = r.bad;
  a.payoff_kind = payoff->kind;
  a.payoff_comp = payoff->comp;
  a.payoff_k = payoff->strike;
  a.payoff_disc = payoff->discount;
  if ((rc = k ? launch(k, capped(grid), &a, st) : launch_sample(m, kt, o, &a, st))) return rc;
  sde_launch_accumulate((const double *)g.partials[slot].p, grid, 2 * SDE_MAXF, coupled ? 6 : 2, r.bins, r.ovf, st);
  g.launches++;
  return 0;
}

extern "C" int sdeb200_mlmc_estimate_level(sdeb200_model *m, const sdeb200_opts *o, const double *params,
                                           const double *x0, double dt_fine, int64_t substeps, int64_t ncoarse,
                                           int64_t npaths, const sdeb200_payoff *payoff, double *result) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if ((rc = check_payoff(m, payoff))) return rc;
  if (payoff->kind == SDEB200_PAYOFF_MOMENTS) return fail(SDEB200_E_ARG, "multilevel estimators need a scalar payoff");
  if (!result) return fail(SDEB200_E_ARG, "result is NULL");
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_TERMINAL))) return rc;
  const int nq = 6;
  if (sharded_call(npaths)) {
    std::vector<std::vector<unsigned long long>> raw(g_ndev, std::vector<unsigned long long>(RED_WORDS, 0));
    rc = run_sharded(npaths, [&](int d, int64_t lo, int64_t n) {
      sdeb200_opts os = *o;
      os.path_offset += lo;
      double tmp[8];
      tl_raw = raw[d].data();
      const int r = sdeb200_mlmc_estimate_level(m, &os, params, x0, dt_fine, substeps, ncoarse, n, payoff, tmp);
      tl_raw = nullptr;
      return r;
    });
    if (rc && rc != SDEB200_E_DOMAIN) return rc;
    std::vector<unsigned long long> total(RED_WORDS, 0);
    add_records(total, raw);
    unsigned long long bad = 0;
    if ((rc = round_records(1, &nq, total.data(), result, 8, &bad))) return rc;
    if (bad) return fail(SDEB200_E_DOMAIN, "%llu coupled paths ended non-finite; they are excluded from the sums", bad);
    return SDEB200_OK;
  }
  if ((rc = g.red.ensure((size_t)RED_WORDS * 8 * MAX_LEVELS))) return rc;
  cudaStream_t st = g.stream();
  if ((rc = timed_begin())) return rc;
  if ((rc = queue_level(m, kt, o, params, x0, true, dt_fine, substeps, ncoarse, npaths, payoff, 0, st))) return rc;
  unsigned long long bad = 0;
  if ((rc = finish_reductions(1, &nq, &st, result, 8, &bad))) return rc;
  if ((rc = timed_end())) return rc;
  if (bad) return fail(SDEB200_E_DOMAIN, "%llu coupled paths ended non-finite; they are excluded from the sums", bad);
  return SDEB200_OK;
}

static int64_t ipow64(int64_t b, int64_t e) {
  int64_t r = 1;
  while (e-- > 0) r *= b;
  return r;
}

extern "C" int sdeb200_mlmc_estimate(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                     int64_t substeps, int64_t level_first, int nlevels, int64_t nsteps,
                                     const int64_t *npaths, const sdeb200_payoff *payoff, double *result) {
  int rc = check_common(m, o, params, x0);
  if (rc) return rc;
  if ((rc = check_payoff(m, payoff))) return rc;
  if (payoff->kind == SDEB200_PAYOFF_MOMENTS) return fail(SDEB200_E_ARG, "multilevel estimators need a scalar payoff");
  if (!npaths || !result || nlevels < 1 || nlevels > MAX_LEVELS || substeps < 1 || level_first < 0 || nsteps < 0)
    return fail(SDEB200_E_ARG, "bad level specification");
  if (o->stream >= (1u << 24))
    return fail(SDEB200_E_ARG, "multilevel runs keep the level index in the top 8 bits of the Philox stream word: stream must be < 2^24");
  if ((rc = ensure_up())) return rc;
  const sde_ktable *kt;
  if ((rc = get_ktable(m, o, &kt, FAM_TERMINAL))) return rc;
  int nq[MAX_LEVELS];
  int64_t nmax = 0;
  for (int i = 0; i < nlevels; i++) nmax = std::max(nmax, npaths[i]);
  if (sharded_call(nmax)) {
    // every level is sharded over all devices: worker d runs its slice of every level concurrently
    for (int i = 0; i < nlevels; i++) nq[i] = i == 0 ? 2 : 6;
    const int nd = g_ndev;
    std::vector<std::vector<unsigned long long>> raw(nd, std::vector<unsigned long long>((size_t)nlevels * RED_WORDS, 0));
    rc = run_devices([&](int d) -> int {
      int64_t n_d[MAX_LEVELS], off_d[MAX_LEVELS], any = 0;
      for (int i = 0; i < nlevels; i++) {
        int64_t lo, hi;
        shard_range(npaths[i], d, nd, &lo, &hi);
        n_d[i] = hi - lo;
        off_d[i] = lo;
        any += n_d[i];
      }
      if (!any) return 0;
      std::vector<double> tmp((size_t)nlevels * 8);
      tl_raw = raw[d].data();
      tl_level_off = off_d;
      const int r = sdeb200_mlmc_estimate(m, o, params, x0, substeps, level_first, nlevels, nsteps, n_d, payoff, tmp.data());
      tl_raw = nullptr;
      tl_level_off = nullptr;
      return r;
    });
    if (rc && rc != SDEB200_E_DOMAIN) return rc;
    std::vector<unsigned long long> total((size_t)nlevels * RED_WORDS, 0);
    add_records(total, raw);
    std::vector<double> tmp((size_t)nlevels * 8, 0.0);
    unsigned long long bad = 0;
    if ((rc = round_records(nlevels, nq, total.data(), tmp.data(), 8, &bad))) return rc;
    for (int i = 0; i < nlevels; i++) {
      double *res = result + (size_t)i * 8, *t = tmp.data() + (size_t)i * 8;
      if (nq[i] == 2) { res[0] = t[0]; res[1] = t[1]; res[2] = t[0]; res[3] = t[1]; res[4] = 0.0; res[5] = 0.0; res[6] = t[2]; res[7] = t[3]; }
      else for (int q = 0; q < 8; q++) res[q] = t[q];
    }
    if (bad) return fail(SDEB200_E_DOMAIN, "%llu paths ended non-finite; they are excluded from the sums", bad);
    return SDEB200_OK;
  }
  if ((rc = g.red.ensure((size_t)RED_WORDS * 8 * MAX_LEVELS))) return rc;
  if ((rc = ensure_level_streams(nlevels))) return rc;
  cudaStream_t st = g.stream();
  if ((rc = timed_begin())) return rc;
  cudaStream_t streams[MAX_LEVELS];
  // finest (longest) levels first so the short ones fill the machine's tail
  for (int i = nlevels - 1; i >= 0; i--) {
    streams[i] = g.lvl[i];
    CU(cudaStreamWaitEvent(streams[i], g.ev0, 0));
    sdeb200_opts ol = *o;
    // fresh, independent noise per level pair (multilevel.jl:58-60): the pair index lives in the top byte of the Philox
    // stream word, away from the streams a caller uses for plain runs (pair 0 IS a plain `sample` on the caller's stream)
    ol.stream = SDEB200_LEVEL_STREAM(o->stream, i);
    if (tl_level_off) ol.path_offset = o->path_offset + tl_level_off[i];   // worker of a sharded call
    const int64_t li = level_first + i;
    if (i == 0) {
      // x[1] = sample(model, ml_schemes[1], t0, x0, nsteps*ml_steps[1], npaths[1])   (multilevel.jl:56)
      ol.dt = o->dt / (double)ipow64(substeps, li);
      rc = queue_level(m, kt, &ol, params, x0, false, 0.0, 1, nsteps * ipow64(substeps, li), npaths[i], payoff, i, streams[i]);
      nq[i] = 2;
    } else {
      // _multilevel_sample(model, ml_schemes[i-1], ml_schemes[i], substeps, t0, x0, nsteps*ml_steps[i-1], npaths[i])
      ol.dt = o->dt / (double)ipow64(substeps, li - 1);
      const double dt_fine = o->dt / (double)ipow64(substeps, li);
      rc = queue_level(m, kt, &ol, params, x0, true, dt_fine, substeps, nsteps * ipow64(substeps, li - 1), npaths[i],
                       payoff, i, streams[i]);
      nq[i] = 6;
    }
    if (rc) return rc;
  }
  std::vector<double> tmp((size_t)nlevels * 8, 0.0);
  unsigned long long bad = 0;
  if ((rc = finish_reductions(nlevels, nq, streams, tmp.data(), 8, &bad))) return rc;
  for (int i = 0; i < nlevels; i++) {
    double *res = result + (size_t)i * 8, *t = tmp.data() + (size_t)i * 8;
    if (nq[i] == 2) {  // level 0: Y = Pf, no coarse partner
      res[0] = t[0]; res[1] = t[1]; res[2] = t[0]; res[3] = t[1]; res[4] = 0.0; res[5] = 0.0; res[6] = t[2]; res[7] = t[3];
    } else {
      for (int q = 0; q < 8; q++) res[q] = t[q];
    }
    CU(cudaEventRecord(g.lvl_done[i], streams[i]));
    CU(cudaStreamWaitEvent(st, g.lvl_done[i], 0));
  }
  if ((rc = timed_end())) return rc;
  if (bad) return fail(SDEB200_E_DOMAIN, "%llu paths ended non-finite; they are excluded from the sums", bad);
  return SDEB200_OK;
}

// The same with one path offset PER LEVEL: rank r of a multi-process run owns [offset_l, offset_l + npaths_l) of level l
// (every level sharded over all ranks), so that the run equals the one-GPU run bit for bit.
extern "C" int sdeb200_mlmc_estimate_sharded(sdeb200_model *m, const sdeb200_opts *o, const double *params, const double *x0,
                                             int64_t substeps, int64_t level_first, int nlevels, int64_t nsteps,
                                             const int64_t *npaths, const int64_t *path_offsets, const sdeb200_payoff *payoff,
                                             double *result) {
  if (!path_offsets) return sdeb200_mlmc_estimate(m, o, params, x0, substeps, level_first, nlevels, nsteps, npaths, payoff, result);
  if (g_ndev > 1) return fail(SDEB200_E_ARG, "per-level offsets belong to the one-process-per-GPU mode");
  const int64_t *saved = tl_level_off;
  tl_level_off = path_offsets;
  const int rc = sdeb200_mlmc_estimate(m, o, params, x0, substeps, level_first, nlevels, nsteps, npaths, payoff, result);
  tl_level_off = saved;
  return rc;
}

// ---------------------------------------------------------------------------------------------------
// library-owned device buffers (hard part 4): a caller without a GPU array type of its own — plain Julia
// without CUDA.jl — keeps large results (C3: 41 GB of trajectories) on the device, passes sdeb200_buffer_ptr(b) wherever
// the drivers take a data pointer, and reads back the slices it needs.
// ---------------------------------------------------------------------------------------------------
struct sdeb200_buffer {
  void *p = nullptr;
  int device = 0, precision = 0, ndim = 0;
  int64_t shape[4] = {0, 0, 0, 0};
  size_t bytes = 0;
};
extern "C" int sdeb200_buffer_alloc(int precision, int ndim, const int64_t *shape, sdeb200_buffer **out) {
  if (!out || !shape || ndim < 1 || ndim > 4 || precision < 0 || precision > 1) return fail(SDEB200_E_ARG, "bad buffer description");
  int rc = ensure_up();
  if (rc) return rc;
  sdeb200_buffer *b = new sdeb200_buffer();
  size_t n = 1;
  for (int i = 0; i < ndim; i++) {
    if (shape[i] < 0) { delete b; return fail(SDEB200_E_ARG, "negative extent"); }
    b->shape[i] = shape[i];
    n *= (size_t)shape[i];
  }
  b->precision = precision;
  b->ndim = ndim;
  b->device = g.device;
  b->bytes = n * esize(precision);
  cudaError_t e = cudaMalloc(&b->p, std::max<size_t>(b->bytes, 16));
  if (e != cudaSuccess) { delete b; return fail(SDEB200_E_CUDA, "cudaMalloc(%zu): %s", n * esize(precision), cudaGetErrorString(e)); }
  *out = b;
  return SDEB200_OK;
}
extern "C" void *sdeb200_buffer_ptr(const sdeb200_buffer *b) { return b ? b->p : nullptr; }
extern "C" int sdeb200_buffer_info(const sdeb200_buffer *b, int *precision, int *ndim, int64_t *shape, int64_t *strides_bytes) {
  if (!b) return fail(SDEB200_E_ARG, "buffer is NULL");
  if (precision) *precision = b->precision;
  if (ndim) *ndim = b->ndim;
  int64_t st = (int64_t)esize(b->precision);
  for (int i = b->ndim - 1; i >= 0; i--) {   // C order: the last extent is contiguous (Julia reads the shape reversed)
    if (shape) shape[i] = b->shape[i];
    if (strides_bytes) strides_bytes[i] = st;
    st *= b->shape[i];
  }
  return SDEB200_OK;
}
extern "C" int sdeb200_buffer_copy_to_host(const sdeb200_buffer *b, int64_t offset_elems, int64_t count_elems, void *host) {
  if (!b || !host || offset_elems < 0 || count_elems < 0) return fail(SDEB200_E_ARG, "bad argument");
  const size_t es = esize(b->precision);
  if (((size_t)offset_elems + (size_t)count_elems) * es > b->bytes) return fail(SDEB200_E_ARG, "range exceeds the buffer");
  int rc = ensure_up();
  if (rc) return rc;
  CU(cudaSetDevice(b->device));
  CU(cudaMemcpy(host, (const char *)b->p + (size_t)offset_elems * es, (size_t)count_elems * es, cudaMemcpyDeviceToHost));
  CU(cudaSetDevice(g.device));
  return SDEB200_OK;
}
extern "C" int sdeb200_buffer_copy_from_host(sdeb200_buffer *b, int64_t offset_elems, int64_t count_elems, const void *host) {
  if (!b || !host || offset_elems < 0 || count_elems < 0) return fail(SDEB200_E_ARG, "bad argument");
  const size_t es = esize(b->precision);
  if (((size_t)offset_elems + (size_t)count_elems) * es > b->bytes) return fail(SDEB200_E_ARG, "range exceeds the buffer");
  int rc = ensure_up();
  if (rc) return rc;
  CU(cudaSetDevice(b->device));
  CU(cudaMemcpy((char *)b->p + (size_t)offset_elems * es, host, (size_t)count_elems * es, cudaMemcpyHostToDevice));
  CU(cudaSetDevice(g.device));
  return SDEB200_OK;
}
extern "C" int sdeb200_buffer_free(sdeb200_buffer *b) {
  if (!b) return SDEB200_OK;
  if (b->p) {
    cudaSetDevice(b->device);
    cudaFree(b->p);
    if (g_ctx[g_primary].up) cudaSetDevice(g_ctx[g_primary].device);
  }
  delete b;
  return SDEB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// communicator
// ---------------------------------------------------------------------------------------------------
extern "C" int sdeb200_comm_unique_id(void *id128) {
  if (!id128) return fail(SDEB200_E_ARG, "id is NULL");
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.b, 128);
  return SDEB200_OK;
}
extern "C" int sdeb200_comm_init(const void *id128, int nranks, int rank) {
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(SDEB200_E_ARG, "bad communicator arguments");
  int rc = ensure_up();
  if (rc) return rc;
  if ((rc = nccl_load())) return rc;
  if (g_ndev > 1) return fail(SDEB200_E_ARG, "sdeb200_init_devices bound %d GPUs to this process: one process per GPU and one process for all GPUs do not mix", g_ndev);
  if (g.comm) sdeb200_comm_destroy();
  nccl_uid id;
  memcpy(id.b, id128, 128);
  NC(g_nccl.CommInitRank(&g.comm, nranks, id, rank));
  g.nranks = nranks;
  g.rank = rank;
  return SDEB200_OK;
}
extern "C" int sdeb200_comm_destroy(void) {
  if (g.comm && g_nccl.CommDestroy) {
    cudaDeviceSynchronize();
    g_nccl.CommDestroy(g.comm);
  }
  g.comm = nullptr;
  g.nranks = 1;
  g.rank = 0;
  return SDEB200_OK;
}
extern "C" int sdeb200_comm_info(int *nranks, int *rank) {
  if (nranks) *nranks = g.nranks;
  if (rank) *rank = g.rank;
  return SDEB200_OK;
}
