// sde_dsl.h — the @sde_model front end re-done in C++ (what src/models/codegen.jl:87-262 computes):
// equation text -> drift vector, D x M diffusion matrix with the correlation Cholesky folded in,
// parameter order, model kind, and a CUDA device functor for NVRTC.
#pragma once
#include <string>
#include <vector>

struct sde_dsl_model {
  int D = 0, M = 0;
  int kind = 0;  // 0 AbstractSDE, 1 StateIndependentDiffusion, 2 JumpSDE
  std::vector<std::string> params, vars, wieners;
  std::vector<std::string> drift_str;      // D entries
  std::vector<std::string> diffusion_str;  // D * max(M,1), row-major
  std::string cuda_source;                 // struct GenModel { ... } in namespace sdeb200
  bool polar = false;                      // M == 2: every diffusion entry carries one sqrt(A) factor -> GenModel::step_polar is emitted
  std::string milstein_error;              // M == 1 only: non-empty when the diffusion could not be differentiated symbolically
};

bool sde_dsl_build(const std::string &name, const std::string &eqs_utf8, const std::vector<std::string> &param_names,
                   sde_dsl_model *out, std::string *err);

// Host evaluation of a symbolic expression string produced by the front end — used only by the CPU unit
// tests of the symbolic engine (tests/test_dsl.py); the product evaluates drift/diffusion on the device.
bool sde_dsl_eval(const std::string &expr, const std::vector<std::string> &names, const std::vector<double> &values,
                  double *out, std::string *err);
