#include "sde_dsl.h"
bool sde_dsl_build(const std::string &, const std::string &, const std::vector<std::string> &, sde_dsl_model *, std::string *err) {
  *err = "DSL front end not built yet";
  return false;
}
bool sde_dsl_eval(const std::string &, const std::vector<std::string> &, const std::vector<double> &, double *, std::string *err) {
  *err = "DSL front end not built yet";
  return false;
}
