// sde_nvrtc.cpp — run-time compilation of a generated model functor into the path kernels of
// sde_device.cuh (NVRTC -> sm_100a cubin -> cudaLibraryLoadData).  libnvrtc is resolved with dlopen so the
// C-ABI library itself has no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/sdeb200.h"
#include "sde_ktable.h"

extern const char *sde_embedded_device_cuh;
extern const char *sde_embedded_args_h;
extern const char *sde_embedded_coeffs_inc;

typedef struct _nvrtcProgram *nvrtcProgram;
struct NvrtcApi {
  void *h = nullptr;
  int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
  int (*CompileProgram)(nvrtcProgram, int, const char *const *);
  int (*GetProgramLogSize)(nvrtcProgram, size_t *);
  int (*GetProgramLog)(nvrtcProgram, char *);
  int (*GetCUBINSize)(nvrtcProgram, size_t *);
  int (*GetCUBIN)(nvrtcProgram, char *);
  int (*DestroyProgram)(nvrtcProgram *);
  const char *(*GetErrorString)(int);
};
static NvrtcApi nv;

static bool nvrtc_load(std::string *log) {
  if (nv.h) return true;
  const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
  for (const char *n : names) {
    nv.h = dlopen(n, RTLD_NOW);
    if (nv.h) break;
  }
  if (!nv.h) {
    *log = std::string("cannot dlopen libnvrtc.so.12: ") + dlerror();
    return false;
  }
#define SYM(f, name)                                                 \
  *(void **)(&nv.f) = dlsym(nv.h, name);                             \
  if (!nv.f) { *log = std::string("libnvrtc lacks ") + name; return false; }
  SYM(CreateProgram, "nvrtcCreateProgram")
  SYM(CompileProgram, "nvrtcCompileProgram")
  SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  SYM(GetProgramLog, "nvrtcGetProgramLog")
  SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  SYM(GetCUBIN, "nvrtcGetCUBIN")
  SYM(DestroyProgram, "nvrtcDestroyProgram")
  SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
  return true;
}

// ---------------------------------------------------------------------------------------------------
// On-disk cache of compiled cubins: a generated model costs 2-12 s of NVRTC per (precision, math) on first use; a
// later process with the same equations, library build and options loads the cubin back instead.
// Directory: $SDEB200_CACHE_DIR (empty string: caching off), else $XDG_CACHE_HOME/sdeb200, else $HOME/.cache/sdeb200.
// Key: two independent 64-bit FNV-1a hashes over the translation unit, the embedded headers and the options.
// ---------------------------------------------------------------------------------------------------
static uint64_t fnv1a(uint64_t h, const char *p, size_t n) {
  for (size_t i = 0; i < n; i++) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ULL; }
  return h;
}
static std::string cache_dir() {
  const char *d = getenv("SDEB200_CACHE_DIR");
  if (d) return std::string(d);   // may be empty: off
  const char *x = getenv("XDG_CACHE_HOME");
  if (x && *x) return std::string(x) + "/sdeb200";
  const char *h = getenv("HOME");
  if (h && *h) return std::string(h) + "/.cache/sdeb200";
  return std::string();
}
static std::string cache_path(const std::string &src, const std::vector<const char *> &opts) {
  const std::string dir = cache_dir();
  if (dir.empty()) return std::string();
  uint64_t a = 0xcbf29ce484222325ULL, b = 0x84222325cbf29ce4ULL;
  const char *parts[] = {src.c_str(), sde_embedded_device_cuh, sde_embedded_args_h, sde_embedded_coeffs_inc};
  for (const char *q : parts) { a = fnv1a(a, q, strlen(q)); b = fnv1a(b ^ 0x9e3779b97f4a7c15ULL, q, strlen(q)); }
  for (const char *o : opts) { a = fnv1a(a, o, strlen(o) + 1); b = fnv1a(b, o, strlen(o) + 1); }
  char name[64];
  snprintf(name, sizeof name, "/%016llx%016llx.cubin", (unsigned long long)a, (unsigned long long)b);
  return dir + name;
}
static bool cache_read(const std::string &path, std::vector<char> *cubin) {
  if (path.empty()) return false;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bool ok = n > 64;
  if (ok) {
    cubin->resize((size_t)n);
    ok = fread(cubin->data(), 1, (size_t)n, f) == (size_t)n && memcmp(cubin->data(), "\x7f" "ELF", 4) == 0;
  }
  fclose(f);
  return ok;
}
static void cache_write(const std::string &path, const std::vector<char> &cubin) {
  if (path.empty()) return;
  const std::string dir = path.substr(0, path.rfind('/'));
  for (size_t i = 1; i <= dir.size(); i++)   // mkdir -p
    if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0700);
  char tmp[32];
  snprintf(tmp, sizeof tmp, ".tmp%ld", (long)getpid());
  const std::string t = path + tmp;
  FILE *f = fopen(t.c_str(), "wb");
  if (!f) return;   // a read-only or missing cache directory only costs the next compile
  const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  fclose(f);
  if (!ok || rename(t.c_str(), path.c_str()) != 0) unlink(t.c_str());
}

// Kernel families (sde_device.cuh): a generated model is compiled lazily, one family per (slot, math) on first use.
enum { FAM_TERMINAL = 0, FAM_TRAJ = 1, FAM_WIENER = 2, FAM_MISC = 3 };
// slot: 0 = f64 on normal stream v1, 1 = f32, 2 = f64 on normal stream v2
static const char *slot_type(int slot) { return slot == 1 ? "float" : "double"; }
static const char *slot_suffix(int slot) { return slot == 1 ? "f32" : (slot == 2 ? "f64v2" : "f64"); }

// Compile only (no device needed): used by the CPU test-suite to check that generated functors build.
int sde_nvrtc_compile_only(const std::string &functor_src, int M, int slot, int math, int family, std::vector<char> *cubin,
                           std::string *log) {
  if (!nvrtc_load(log)) return SDEB200_E_NVRTC;
  std::string src = "#include \"sde_device.cuh\"\n";
  src += functor_src;
  const std::string T = slot_type(slot), SUF = slot_suffix(slot), V = slot == 2 ? "2" : "1";
  const int nsch = M == 1 ? 3 : 1;
  const char *sn[3] = {"em", "mil", "rk"};
  for (int s = 0; s < nsch; s++) {
    const std::string head = std::string("(sdeb200::GenModel, gen, ") + std::to_string(s) + ", " + sn[s] + ", " + T;
    if (family == FAM_TERMINAL) src += "\nSDE_KERNELS_TERMINAL" + head + ", " + V + ", " + SUF + ")\n";
    if (family == FAM_TRAJ) src += "\nSDE_KERNELS_TRAJ" + head + ", " + V + ", " + SUF + ")\n";
    if (family == FAM_WIENER) src += "\nSDE_KERNELS_WIENER" + head + ", " + SUF + ")\n";
  }
  if (family == FAM_MISC) {
    src += "\nSDE_KERNELS_COEF(sdeb200::GenModel, gen)\n";
    src += std::string("SDE_KERNELS_LOGT(sdeb200::GenModel, gen, ") + T + ", " + SUF + ")\n";
  }
  const char *hdr_src[] = {sde_embedded_device_cuh, sde_embedded_args_h, sde_embedded_coeffs_inc};
  const char *hdr_name[] = {"sde_device.cuh", "sde_args.h", "sde_coeffs.inc"};
  // -default-device: the generic lambdas of the device code carry no execution-space annotation (nvcc infers it, JIT mode
  // otherwise takes them for host functions)
  std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device",
                                    math ? "-DSDE_STRICT=1" : "-DSDE_STRICT=0"};
  if (math) opts.push_back("--fmad=false");
  const std::string cpath = cache_path(src, opts);
  if (cache_read(cpath, cubin)) return 0;
  nvrtcProgram prog = nullptr;
  int r = nv.CreateProgram(&prog, src.c_str(), "sde_generated_model.cu", 3, hdr_src, hdr_name);
  if (r) { *log = std::string("nvrtcCreateProgram: ") + nv.GetErrorString(r); return SDEB200_E_NVRTC; }
  r = nv.CompileProgram(prog, (int)opts.size(), opts.data());
  size_t ls = 0;
  nv.GetProgramLogSize(prog, &ls);
  if (ls > 1) {
    std::vector<char> buf(ls + 1, 0);
    nv.GetProgramLog(prog, buf.data());
    *log = buf.data();
  }
  if (r) {
    *log = std::string("nvrtcCompileProgram: ") + nv.GetErrorString(r) + "\n" + *log + "\n--- source ---\n" + functor_src;
    nv.DestroyProgram(&prog);
    return SDEB200_E_NVRTC;
  }
  size_t cs = 0;
  nv.GetCUBINSize(prog, &cs);
  cubin->resize(cs);
  nv.GetCUBIN(prog, cubin->data());
  nv.DestroyProgram(&prog);
  cache_write(cpath, *cubin);
  return 0;
}

int sde_nvrtc_compile(const std::string &functor_src, int D, int M, int slot, int math, int family, sde_ktable *kt,
                      void **library_out, std::string *log) {
  (void)D;
  std::vector<char> cubin;
  int rc = sde_nvrtc_compile_only(functor_src, M, slot, math, family, &cubin, log);
  if (rc) return rc;
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) {
    *log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e);
    return SDEB200_E_CUDA;
  }
  bool ok = true;
  auto get = [&](const std::string &nm, const void **dst) {
    cudaKernel_t k = nullptr;
    cudaError_t ee = cudaLibraryGetKernel(&k, lib, nm.c_str());
    if (ee != cudaSuccess) {
      *log = "cudaLibraryGetKernel(" + nm + "): " + cudaGetErrorString(ee);
      ok = false;
      return;
    }
    *dst = (const void *)k;
  };
  const int nsch = M == 1 ? 3 : 1;
  const char *sn[3] = {"em", "mil", "rk"};
  const std::string SUF = slot_suffix(slot);
  for (int s = 0; s < nsch && ok; s++) {
    const std::string suf = std::string("_") + sn[s] + "_" + SUF;
    if (family == FAM_TERMINAL) {
      get("sdek_gen_sample" + suf, &kt->sample[s][slot]);
      get("sdek_gen_sample1" + suf, &kt->sample1[s][slot]);
      get("sdek_gen_mlmc" + suf, &kt->mlmc[s][slot]);
    } else if (family == FAM_TRAJ) {
      get("sdek_gen_simsoa" + suf, &kt->simsoa[s][slot]);
      get("sdek_gen_simref" + suf, &kt->simref[s][slot]);
      get("sdek_gen_simrt" + suf, &kt->simrt[s][slot]);
    } else if (family == FAM_WIENER) {   // no noise: stream v2 compiles (and caches) its own copy of the f64 kernels
      get("sdek_gen_simw" + suf, &kt->simw[s][slot]);
      get("sdek_gen_simwt" + suf, &kt->simwt[s][slot]);
    }
  }
  if (family == FAM_MISC && ok) {
    get(std::string("sdek_gen_logt_") + SUF, &kt->logt[slot == 1 ? 1 : 0]);
    get(std::string("sdek_gen_logtt_") + SUF, &kt->logtt[slot == 1 ? 1 : 0]);
    get("sdek_gen_coef", &kt->coef);
  }
  if (!ok) {
    cudaLibraryUnload(lib);
    return SDEB200_E_CUDA;
  }
  *library_out = (void *)lib;
  return 0;
}

// C entry used by the CPU tests: does the functor of a model text compile for sm_100a?  (terminal-value kernels: every
// `step` form; coefficient / logtransition kernels: `coefficients`)
extern "C" int sdeb200_debug_nvrtc_check(const char *functor_src, int M, int precision, int math, char *log, int logcap) {
  std::vector<char> cubin;
  std::string lg;
  int total = 0;
  for (int family : {FAM_TERMINAL, FAM_MISC}) {
    int rc = sde_nvrtc_compile_only(functor_src, M, precision == 1 ? 1 : 0, math, family, &cubin, &lg);
    if (rc) {
      if (log && logcap > 0) { strncpy(log, lg.c_str(), logcap - 1); log[logcap - 1] = 0; }
      return rc;
    }
    total += (int)cubin.size();
  }
  if (log && logcap > 0) { strncpy(log, lg.c_str(), logcap - 1); log[logcap - 1] = 0; }
  return total;
}

// one family of one slot (CPU tests: every family must build under NVRTC, not only under nvcc)
extern "C" int sdeb200_debug_nvrtc_check_family(const char *functor_src, int M, int slot, int math, int family, char *log, int logcap) {
  std::vector<char> cubin;
  std::string lg;
  int rc = sde_nvrtc_compile_only(functor_src, M, slot, math, family, &cubin, &lg);
  if (log && logcap > 0) { strncpy(log, lg.c_str(), logcap - 1); log[logcap - 1] = 0; }
  return rc ? rc : (int)cubin.size();
}
