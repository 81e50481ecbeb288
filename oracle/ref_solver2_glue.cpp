// ref_solver2_glue.cpp -- extern "C" access to the REFERENCE'S OWN compiled ISRSolverSLE2 (per-point efficiency histograms).
//
// ORACLE / TEST INFRASTRUCTURE (see ref_solver_glue.cpp; same shims).  src/BaseISRSolver2.cpp and src/ISRSolverSLE2.cpp are
// compiled unmodified from /root/reference; BaseISRSolver2.hpp and BaseISRSolver.hpp define the same exception type and
// cannot share a translation unit, hence this second file.
#include <cstddef>
#include <exception>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "ISRSolverSLE2.hpp"
#include "root_shim.h"

struct ref_eff {   // layout of orc_eff (oracle/isr_oracle.c); kind 1 = one x-histogram per energy point
  int kind;
  int nx; double xlo, xhi; const double* xedges;
  int ne; double elo, ehi; const double* eedges;
  const double* values;
};

namespace {
thread_local std::string g_error2;
template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error2 = e.what();
    return 1;
  } catch (...) {
    g_error2 = "unknown exception";
    return 1;
  }
}
}  // namespace

extern "C" {

const char* refs2_last_error() { return g_error2.c_str(); }

// ISRSolverSLE2(TGraphErrors*, std::vector<TH1D*>, threshold) (src/ISRSolverSLE2.cpp:51-57; src/BaseISRSolver2.cpp:151-158):
// histogram i holds values[i][0 .. nx + 1]; the solver owns the histograms and deletes them.
void* refs2_create(int n, const double* ecm, const double* vcs, const double* ecm_err, const double* vcs_err, double threshold,
                   const ref_eff* eff) {
  ISRSolverSLE2* s = nullptr;
  const int rc = guarded([&]() {
    TGraphErrors g(n, ecm, vcs, ecm_err, vcs_err);
    if (eff && eff->kind == 1) {
      std::vector<TH1D*> hs;
      for (int i = 0; i < n; ++i) {
        const std::string name = "eff_" + std::to_string(i);
        TH1D* h = eff->xedges ? new TH1D(name.c_str(), "", eff->nx, eff->xedges) : new TH1D(name.c_str(), "", eff->nx, eff->xlo, eff->xhi);
        for (int b = 0; b < eff->nx + 2; ++b) h->SetBinContent(b, eff->values[(std::size_t)i * (eff->nx + 2) + b]);
        hs.push_back(h);
      }
      s = new ISRSolverSLE2(&g, hs, threshold);
    } else {
      s = new ISRSolverSLE2(&g, threshold);
    }
    if (ecm_err) s->enableEnergySpread();
  });
  return rc ? nullptr : s;
}
void refs2_free(void* vh) { delete static_cast<ISRSolverSLE2*>(vh); }
int refs2_set_ranges(void* vh, int nr, const int* settings) {
  auto* s = static_cast<ISRSolverSLE2*>(vh);
  return guarded([&]() {
    std::vector<std::tuple<bool, int, int>> st;
    for (int r = 0; r < nr; ++r) st.emplace_back(settings[3 * r] != 0, settings[3 * r + 1], settings[3 * r + 2]);
    s->setRangeInterpSettings(st);
  });
}
int refs2_eval_eq_matrix(void* vh) {
  auto* s = static_cast<ISRSolverSLE2*>(vh);
  return guarded([&]() { s->evalEqMatrix(); });
}
int refs2_solve(void* vh) {
  auto* s = static_cast<ISRSolverSLE2*>(vh);
  return guarded([&]() { s->solve(); });
}
void refs2_get_bcs(void* vh, double* out) {
  auto* s = static_cast<ISRSolverSLE2*>(vh);
  for (int i = 0; i < (int)s->getN(); ++i) out[i] = s->bcs()(i);
}
void refs2_get_cov(void* vh, double* out) {
  const Eigen::MatrixXd& c = static_cast<ISRSolverSLE2*>(vh)->getBornCSCovMatrix();
  for (Eigen::Index q = 0; q < c.size(); ++q) out[q] = c.data()[q];
}
void refs2_get_matrix(void* vh, double* out) {
  const Eigen::MatrixXd& g = static_cast<ISRSolverSLE2*>(vh)->getIntegralOperatorMatrix();
  for (Eigen::Index q = 0; q < g.size(); ++q) out[q] = g.data()[q];
}
double refs2_cond(void* vh) { return static_cast<ISRSolverSLE2*>(vh)->evalConditionNumber(); }

}  // extern "C"
