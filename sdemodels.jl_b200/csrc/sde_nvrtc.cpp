// sde_nvrtc.cpp — run-time compilation of a generated model functor into the path kernels of
// sde_device.cuh (NVRTC -> sm_100a cubin -> cudaLibraryLoadData).  libnvrtc is resolved with dlopen so the
// C-ABI library itself has no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>

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

// Compile only (no device needed): used by the CPU test-suite to check that generated functors build.
int sde_nvrtc_compile_only(const std::string &functor_src, int M, int precision, int math, std::vector<char> *cubin,
                           std::string *log) {
  if (!nvrtc_load(log)) return SDEB200_E_NVRTC;
  std::string src = "#include \"sde_device.cuh\"\n";
  src += functor_src;
  const char *T = precision == 0 ? "double" : "float", *P = precision == 0 ? "f64" : "f32";
  src += std::string("\nSDE_KERNELS(sdeb200::GenModel, gen, 0, em, ") + T + ", " + P + ")\n";
  if (M == 1) src += std::string("SDE_KERNELS(sdeb200::GenModel, gen, 1, mil, ") + T + ", " + P + ")\n";
  if (M == 1) src += std::string("SDE_KERNELS(sdeb200::GenModel, gen, 2, rk, ") + T + ", " + P + ")\n";
  src += "SDE_KERNELS_COEF(sdeb200::GenModel, gen)\n";
  src += std::string("SDE_KERNELS_LOGT(sdeb200::GenModel, gen, ") + T + ", " + P + ")\n";
  const char *hdr_src[] = {sde_embedded_device_cuh, sde_embedded_args_h, sde_embedded_coeffs_inc};
  const char *hdr_name[] = {"sde_device.cuh", "sde_args.h", "sde_coeffs.inc"};
  nvrtcProgram prog = nullptr;
  int r = nv.CreateProgram(&prog, src.c_str(), "sde_generated_model.cu", 3, hdr_src, hdr_name);
  if (r) { *log = std::string("nvrtcCreateProgram: ") + nv.GetErrorString(r); return SDEB200_E_NVRTC; }
  std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo",
                                    math ? "-DSDE_STRICT=1" : "-DSDE_STRICT=0"};
  if (math) opts.push_back("--fmad=false");
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
  return 0;
}

int sde_nvrtc_compile(const std::string &functor_src, int D, int M, int precision, int math, sde_ktable *kt,
                      void **library_out, std::string *log) {
  (void)D;
  std::vector<char> cubin;
  int rc = sde_nvrtc_compile_only(functor_src, M, precision, math, &cubin, log);
  if (rc) return rc;
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) {
    *log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e);
    return SDEB200_E_CUDA;
  }
  const char *P = precision == 0 ? "f64" : "f32";
  auto get = [&](const std::string &nm, const void **dst) -> bool {
    cudaKernel_t k = nullptr;
    cudaError_t ee = cudaLibraryGetKernel(&k, lib, nm.c_str());
    if (ee != cudaSuccess) {
      *log = "cudaLibraryGetKernel(" + nm + "): " + cudaGetErrorString(ee);
      return false;
    }
    *dst = (const void *)k;
    return true;
  };
  const int nsch = M == 1 ? 3 : 1;
  const char *sn[3] = {"em", "mil", "rk"};
  for (int s = 0; s < nsch; s++) {
    const std::string suf = std::string("_") + sn[s] + "_" + P;
    if (!get("sdek_gen_sample" + suf, &kt->sample[s][precision]) || !get("sdek_gen_sample1" + suf, &kt->sample1[s][precision]) || !get("sdek_gen_mlmc" + suf, &kt->mlmc[s][precision]) ||
        !get("sdek_gen_simsoa" + suf, &kt->simsoa[s][precision]) || !get("sdek_gen_simref" + suf, &kt->simref[s][precision]) ||
        !get("sdek_gen_simw" + suf, &kt->simw[s][precision]) || !get("sdek_gen_simwt" + suf, &kt->simwt[s][precision]) ||
        !get("sdek_gen_simrt" + suf, &kt->simrt[s][precision])) {
      cudaLibraryUnload(lib);
      return SDEB200_E_CUDA;
    }
  }
  if (!get(std::string("sdek_gen_logt_") + P, &kt->logt[precision]) || !get("sdek_gen_coef", &kt->coef)) {
    cudaLibraryUnload(lib);
    return SDEB200_E_CUDA;
  }
  *library_out = (void *)lib;
  return 0;
}

// C entry used by the CPU tests: does the functor of a model text compile for sm_100a?
extern "C" int sdeb200_debug_nvrtc_check(const char *functor_src, int M, int precision, int math, char *log, int logcap) {
  std::vector<char> cubin;
  std::string lg;
  int rc = sde_nvrtc_compile_only(functor_src, M, precision, math, &cubin, &lg);
  if (log && logcap > 0) {
    strncpy(log, lg.c_str(), logcap - 1);
    log[logcap - 1] = 0;
  }
  return rc ? rc : (int)cubin.size();
}
