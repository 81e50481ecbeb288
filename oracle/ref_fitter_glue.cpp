// ref_fitter_glue.cpp -- extern "C" access to the REFERENCE'S OWN compiled ISRSolverVCSFitFunction.
//
// ORACLE / TEST INFRASTRUCTURE (see ref_solver_glue.cpp; same shims plus a three-class Minuit2 stand-in).
// src/ISRSolverVCSFitter.cpp is compiled unmodified from /root/reference: the chi2 objective of the visible-cross-section
// fit -- N forward convolutions of a parametrised Born cross section, optionally blurred by the beam-energy spread
// (operator(), :64-93) -- and saveResults (:107-177: blurred visible cross section, radiative correction delta, Born points
// vcs / (1 + delta)).  The minimiser (Minuit2) is not part of this; parameters are given.
#include <cstddef>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "ISRSolverVCSFitter.hpp"
#include "root_shim.h"

namespace {
thread_local std::string g_error3;
template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error3 = e.what();
    return 1;
  } catch (...) {
    g_error3 = "unknown exception";
    return 1;
  }
}
typedef double (*FitFcn)(double en, const double* par, int npar, void* ctx);
}  // namespace

extern "C" {

const char* reff_last_error() { return g_error3.c_str(); }

// ISRSolverVCSFitFunction(n, threshold, energy, vis_cs, energy_err, vis_cs_err, fit_fcn) (src/ISRSolverVCSFitter.cpp:15-56);
// efficiency = 1; the points are sorted by energy inside.  energy_err != NULL also enables the energy spread.
void* reff_create(int n, double threshold, const double* ecm, const double* vcs, const double* ecm_err, const double* vcs_err,
                  FitFcn fcn, void* ctx) {
  ISRSolverVCSFitFunction* f = nullptr;
  const int rc = guarded([&]() {
    std::function<double(double, const std::vector<double>&)> ff = [fcn, ctx](double en, const std::vector<double>& par) {
      return fcn(en, par.data(), (int)par.size(), ctx);
    };
    f = new ISRSolverVCSFitFunction((std::size_t)n, threshold, const_cast<double*>(ecm), const_cast<double*>(vcs),
                                    const_cast<double*>(ecm_err), const_cast<double*>(vcs_err), ff);
    if (ecm_err) f->enableEnergySpread();
  });
  return rc ? nullptr : f;
}
void reff_free(void* vh) { delete static_cast<ISRSolverVCSFitFunction*>(vh); }
void reff_set_spread(void* vh, int on) {
  auto* f = static_cast<ISRSolverVCSFitFunction*>(vh);
  if (on) f->enableEnergySpread(); else f->disableEnergySpread();
}
// chi2(par) = operator()(par)
int reff_chi2(void* vh, const double* par, int npar, double* chi2) {
  auto* f = static_cast<ISRSolverVCSFitFunction*>(vh);
  return guarded([&]() { *chi2 = (*f)(std::vector<double>(par, par + npar)); });
}
// saveResults at the given parameters: out[4][n] = vcs_gauss_conv, rad_delta, bcs, bcs_err (in energy order)
int reff_save(void* vh, const double* par, int npar, int n, double* out) {
  auto* f = static_cast<ISRSolverVCSFitFunction*>(vh);
  return guarded([&]() {
    root_shim::registry().clear();
    ROOT::Minuit2::FunctionMinimum fm(std::vector<double>(par, par + npar));
    f->saveResults(fm, "out.root");
    auto& reg = root_shim::registry();
    auto* g1 = dynamic_cast<TGraph*>(reg.at("vcs_gauss_conv").get());
    auto* g2 = dynamic_cast<TGraph*>(reg.at("rad_delta").get());
    auto* g3 = dynamic_cast<TGraphErrors*>(reg.at("bcs").get());
    for (int i = 0; i < n; ++i) {
      out[i] = g1->GetY()[i];
      out[n + i] = g2->GetY()[i];
      out[2 * n + i] = g3->GetY()[i];
      out[3 * n + i] = g3->GetEY()[i];
    }
  });
}

}  // extern "C"
