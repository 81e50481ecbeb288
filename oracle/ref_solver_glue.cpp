// ref_solver_glue.cpp -- extern "C" access to the REFERENCE'S OWN compiled solver classes and toy-test drivers.
//
// ORACLE / TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load oracle/_ref/libisr_ref.so).
//
// oracle/Makefile compiles, unmodified and from where they lie under /root/reference/src:
//   BaseISRSolver.cpp  ISRSolverSLE.cpp  ISRSolverTikhonov.cpp  ISRSolverTSVD.cpp  IterISRInterpSolver.cpp  Chi2Test.cpp
//   RatioTest.cpp  Utils.cpp
// against oracle/ref_shim: a dense Eigen stand-in (textbook decompositions under Eigen's class names), ROOT containers
// backed by an in-memory registry, Boost.Format / Boost.Accumulators stand-ins, an empty nlopt.hpp.  Everything between
// the constructor and the numbers -- sorting the points, the efficiency adaptor, evalEqMatrix with the energy-spread
// product, solve() of the three solver classes, the L-curve functionals, chi2TestModel / chi2TestData / ratioTestModel
// with their per-toy solve() loops, histogram ranges and means -- is the reference's own code running.
// This file only constructs the objects, feeds gRandom's queue and copies results out.
#include <cstddef>
#include <cstdio>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "Chi2Test.hpp"
#include "ISRSolverSLE.hpp"
#include "ISRSolverTSVD.hpp"
#include "ISRSolverTikhonov.hpp"
#include "IterISRInterpSolver.hpp"
#include "RatioTest.hpp"
#include "Utils.hpp"
#include "root_shim.h"

// ---- the shim's global state (declared in ref_shim/root_shim.h)
namespace root_shim {
std::map<std::string, std::shared_ptr<TObject>>& registry() {
  static std::map<std::string, std::shared_ptr<TObject>> r;
  return r;
}
std::deque<double>& gaus_queue() {
  static std::deque<double> q;
  return q;
}
}  // namespace root_shim
static TRandom g_random_instance;
TRandom* gRandom = &g_random_instance;

// same layout as orc_eff in oracle/isr_oracle.c and ref_eff in ref_glue.cpp
struct ref_eff {
  int kind;  // 0: none; 2: TEfficiency 2-D (x, E); 3: TEfficiency 1-D (E only)
  int nx; double xlo, xhi; const double* xedges;
  int ne; double elo, ehi; const double* eedges;
  const double* values;
};

namespace {

// protected members the glue needs: the operator matrix (to inject a precomputed one), the prepared flag, D and w
struct SleX : ISRSolverSLE {
  using ISRSolverSLE::ISRSolverSLE;
  void inject(const Eigen::MatrixXd& G) {
    _getIntegralOperatorMatrix() = G;
    _isEqMatrixPrepared = true;
  }
};
struct TikX : ISRSolverTikhonov {
  using ISRSolverTikhonov::ISRSolverTikhonov;
  void inject(const Eigen::MatrixXd& G) {   // what solve() :62-67 does, with the fill loop replaced by G
    _evalDotProductOperator();
    _getIntegralOperatorMatrix() = G;
    _evalInterpPointWiseDerivativeProjector();
    _isEqMatrixPrepared = true;
  }
  const Eigen::MatrixXd& D() const { return _getInterpPointWiseDerivativeProjector(); }
  const Eigen::RowVectorXd& w() const { return _getDotProdOp(); }
};
struct TsvdX : ISRSolverTSVD {
  using ISRSolverTSVD::ISRSolverTSVD;
};
struct IterX : IterISRInterpSolver {
  using IterISRInterpSolver::IterISRInterpSolver;
  void inject(const Eigen::MatrixXd& G) {
    _getIntegralOperatorMatrix() = G;
    _isEqMatrixPrepared = true;
  }
};

struct Handle {
  int kind;  // 0 SLE, 1 Tikhonov, 2 TSVD, 3 iterative
  std::unique_ptr<ISRSolverSLE> s;
  std::string error;
};

thread_local std::string g_error;

template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  } catch (...) {
    g_error = "unknown exception";
    return 1;
  }
}

Eigen::VectorXd vec(const double* p, int n) {
  Eigen::VectorXd v(n);
  for (int i = 0; i < n; ++i) v(i) = p[i];
  return v;
}

}  // namespace

extern "C" {

const char* refs_last_error() { return g_error.c_str(); }

// kind: 0 ISRSolverSLE, 1 ISRSolverTikhonov (lambda), 2 ISRSolverTSVD (upper index), 3 IterISRInterpSolver (upper = iterations).
// ctor 0: the pointer constructor PyISR uses (include/PyISRSolver.hpp; src/BaseISRSolver.cpp:78-128; ecm_err == NULL
//         switches the energy spread off, efficiency = 1);
// ctor 1: the TGraphErrors (+ TEfficiency) constructor of the executables (src/BaseISRSolver.cpp:31-76, 135-165, 232-258);
//         the energy spread is then enabled explicitly when ecm_err != NULL, as src/isrsolver-SLE.cpp does on --enable-energy-spread.
void* refs_create(int kind, int ctor, int n, const double* ecm, const double* vcs, const double* ecm_err, const double* vcs_err,
                  double threshold, const ref_eff* eff, double lambda, int upper) {
  auto* h = new Handle();
  h->kind = kind;
  const int rc = guarded([&]() {
    if (ctor == 0) {
      std::function<double(double, double)> one = [](double, double) { return 1.; };
      double* e = const_cast<double*>(ecm);
      double* v = const_cast<double*>(vcs);
      double* ee = const_cast<double*>(ecm_err);
      double* ve = const_cast<double*>(vcs_err);
      if (kind == 0) h->s.reset(new SleX((std::size_t)n, e, v, ee, ve, threshold, one));
      if (kind == 1) {
        auto* t = new TikX((std::size_t)n, e, v, ee, ve, threshold, one);
        t->setLambda(lambda);
        h->s.reset(t);
      }
      if (kind == 2) {
        auto* t = new TsvdX((std::size_t)n, e, v, ee, ve, threshold, one);
        t->setUpperTSVDIndex(upper);
        h->s.reset(t);
      }
      if (kind == 3) h->s.reset(new IterX((std::size_t)n, e, v, ee, ve, threshold, one));
    } else {
      TGraphErrors g(n, ecm, vcs, ecm_err, vcs_err);
      std::unique_ptr<TEfficiency> te;
      if (eff && eff->kind == 2)
        te.reset(new TEfficiency(2, eff->nx, eff->xlo, eff->xhi, eff->xedges, eff->ne, eff->elo, eff->ehi, eff->eedges, eff->values));
      if (eff && eff->kind == 3)
        te.reset(new TEfficiency(1, eff->ne, eff->elo, eff->ehi, eff->eedges, 0, 0.0, 0.0, nullptr, eff->values));
      if (kind == 0) h->s.reset(te ? new SleX(&g, te.get(), threshold) : new SleX(&g, threshold));
      if (kind == 1) h->s.reset(te ? new TikX(&g, te.get(), threshold, lambda) : new TikX(&g, threshold, lambda));
      if (kind == 2) h->s.reset(te ? new TsvdX(&g, te.get(), threshold, upper) : new TsvdX(&g, threshold, upper));
      if (kind == 3) h->s.reset(te ? new IterX(&g, te.get(), threshold) : new IterX(&g, threshold));
      if (ecm_err) h->s->enableEnergySpread();
    }
  });
  if (rc || !h->s) {
    delete h;
    return nullptr;
  }
  return h;
}
void refs_free(void* vh) { delete static_cast<Handle*>(vh); }

// setRangeInterpSettings (src/ISRSolverSLE.cpp:124-127); returns 1 on InterpRangeException
int refs_set_ranges(void* vh, int nr, const int* settings) {
  auto* h = static_cast<Handle*>(vh);
  return guarded([&]() {
    std::vector<std::tuple<bool, int, int>> st;
    for (int r = 0; r < nr; ++r) st.emplace_back(settings[3 * r] != 0, settings[3 * r + 1], settings[3 * r + 2]);
    h->s->setRangeInterpSettings(st);
  });
}

// replace evalEqMatrix's N^2 adaptive integrations by a given operator matrix (column-major); not available for TSVD,
// whose solve() computes the SVD inside the same not-prepared branch (src/ISRSolverTSVD.cpp:54-62)
int refs_inject_matrix(void* vh, const double* G) {
  auto* h = static_cast<Handle*>(vh);
  const int n = (int)h->s->getN();
  Eigen::MatrixXd M(n, n);
  for (int q = 0; q < n * n; ++q) M.data()[q] = G[q];
  if (h->kind == 0) { static_cast<SleX*>(h->s.get())->inject(M); return 0; }
  if (h->kind == 1) return guarded([&]() { static_cast<TikX*>(h->s.get())->inject(M); });
  if (h->kind == 3) { static_cast<IterX*>(h->s.get())->inject(M); return 0; }
  g_error = "refs_inject_matrix: not available for ISRSolverTSVD";
  return 1;
}

int refs_eval_eq_matrix(void* vh) {
  auto* h = static_cast<Handle*>(vh);
  return guarded([&]() { h->s->evalEqMatrix(); });
}
int refs_solve(void* vh) {
  auto* h = static_cast<Handle*>(vh);
  return guarded([&]() { h->s->solve(); });
}
// resetECMErrors + enable / disableEnergySpread (src/BaseISRSolver.cpp:428-459): what src/isrsolver-condnum-test.cpp:120-128
// does per spread setting before it re-assembles the matrix and takes its condition number
void refs_reset_ecm_err(void* vh, const double* e, int enable_spread) {
  auto* h = static_cast<Handle*>(vh);
  h->s->resetECMErrors(vec(e, (int)h->s->getN()));
  if (enable_spread) h->s->enableEnergySpread(); else h->s->disableEnergySpread();
}
void refs_reset_vcs(void* vh, const double* v) {
  auto* h = static_cast<Handle*>(vh);
  h->s->resetVisibleCS(vec(v, (int)h->s->getN()));
}

// sorted inputs as the solver holds them (src/BaseISRSolver.cpp:105-128): out[4][N] = ecm, ecm_err, vcs, vcs_err
void refs_get_points(void* vh, double* out) {
  auto* h = static_cast<Handle*>(vh);
  const int n = (int)h->s->getN();
  for (int i = 0; i < n; ++i) {
    out[i] = h->s->ecm()(i);
    out[n + i] = h->s->ecmErr()(i);
    out[2 * n + i] = h->s->vcs()(i);
    out[3 * n + i] = h->s->vcsErr()(i);
  }
}
void refs_get_bcs(void* vh, double* out) {
  auto* h = static_cast<Handle*>(vh);
  for (int i = 0; i < (int)h->s->getN(); ++i) out[i] = h->s->bcs()(i);
}
void refs_get_cov(void* vh, double* out) {   // column-major
  auto* h = static_cast<Handle*>(vh);
  const Eigen::MatrixXd& c = h->s->getBornCSCovMatrix();
  for (Eigen::Index q = 0; q < c.size(); ++q) out[q] = c.data()[q];
}
void refs_get_matrix(void* vh, double* out) {   // column-major
  auto* h = static_cast<Handle*>(vh);
  const Eigen::MatrixXd& g = h->s->getIntegralOperatorMatrix();
  for (Eigen::Index q = 0; q < g.size(); ++q) out[q] = g.data()[q];
}
double refs_cond(void* vh) { return static_cast<Handle*>(vh)->s->evalConditionNumber(); }
double refs_interp_eval(void* vh, const double* y, double e) {
  auto* h = static_cast<Handle*>(vh);
  return h->s->interpEval(vec(y, (int)h->s->getN()), e);
}

// Tikhonov
void refs_tik_set(void* vh, double lambda, int deriv_reg) {
  auto* t = static_cast<TikX*>(static_cast<Handle*>(vh)->s.get());
  t->setLambda(lambda);
  if (deriv_reg) t->enableDerivNorm2Regularizator(); else t->disableDerivNorm2Regularizator();
}
// out = {evalEqNorm2, evalSmoothnessConstraintNorm2, evalLCurveCurvature, evalLCurveCurvatureDerivative} after solve()
void refs_tik_functionals(void* vh, double* out) {
  auto* t = static_cast<TikX*>(static_cast<Handle*>(vh)->s.get());
  out[0] = t->evalEqNorm2();
  out[1] = t->evalSmoothnessConstraintNorm2();
  out[2] = t->evalLCurveCurvature();
  out[3] = t->evalLCurveCurvatureDerivative();
}
// lambdaObjective (src/Utils.cpp:18-38): what NLopt calls per trial lambda; returns the curvature, *grad its derivative
double refs_lambda_objective(void* vh, double lambda, double* grad) {
  auto* t = static_cast<ISRSolverTikhonov*>(static_cast<Handle*>(vh)->s.get());
  return lambdaObjective(1, &lambda, grad, t);
}
void refs_tik_operators(void* vh, double* D, double* w) {   // D column-major
  auto* t = static_cast<TikX*>(static_cast<Handle*>(vh)->s.get());
  const int n = (int)t->getN();
  for (int q = 0; q < n * n; ++q) D[q] = t->D().data()[q];
  for (int i = 0; i < n; ++i) w[i] = t->w()(i);
}

// IterISRInterpSolver (src/IterISRInterpSolver.cpp:45-57): number of iterations; the radiative correction is private and
// leaves the object only through save() (:67-93, the "radiative_correction" graph), which is how it is read here
void refs_iter_set(void* vh, int n_iter) { static_cast<IterX*>(static_cast<Handle*>(vh)->s.get())->setNumOfIters((std::size_t)n_iter); }
int refs_iter_radcorr(void* vh, double* out) {
  auto* h = static_cast<Handle*>(vh);
  return guarded([&]() {
    root_shim::registry().clear();
    OutputOptions o;
    o.visibleCSGraphName = "vcs";
    o.bornCSGraphName = "bcs";
    h->s->save("out.root", o);
    auto* g = dynamic_cast<TGraph*>(root_shim::registry().at("radiative_correction").get());
    for (int i = 0; i < g->GetN(); ++i) out[i] = g->GetY()[i];
  });
}

// TSVD
void refs_tsvd_set(void* vh, int upper, int keep_one) {
  auto* t = static_cast<TsvdX*>(static_cast<Handle*>(vh)->s.get());
  t->setUpperTSVDIndex(upper);
  if (keep_one) t->enableKeepOne(); else t->disableKeepOne();
}

static void register_model(int n, const double* ecm, const double* model_vcs, const double* model_bcs) {
  root_shim::registry().clear();
  root_shim::registry()["vcs"] = std::make_shared<TGraphErrors>(n, ecm, model_vcs);
  root_shim::registry()["bcs"] = std::make_shared<TGraphErrors>(n, ecm, model_bcs);
}
static void queue_toys(const double* toys, long count) {
  auto& q = root_shim::gaus_queue();
  q.clear();
  for (long i = 0; i < count; ++i) q.push_back(toys[i]);
}
static int read_chi2_hist(double* chi2, double* lohi, float* counts) {
  auto it = root_shim::registry().find("chi2Hist");
  if (it == root_shim::registry().end()) { g_error = "chi2Hist was not written"; return 1; }
  auto* hst = dynamic_cast<TH1F*>(it->second.get());
  for (std::size_t i = 0; i < hst->fills().size(); ++i) chi2[i] = hst->fills()[i];
  lohi[0] = hst->lo();
  lohi[1] = hst->hi();
  for (std::size_t i = 0; i < hst->counts().size(); ++i) counts[i] = hst->counts()[i];
  return 0;
}

// chi2TestModel (src/Chi2Test.cpp:80-101 -> chi2Test :29-78) on the toys[n_toys][N] given (toy-major; they replace the
// values gRandom->Gaus would draw, in the order randomDrawVisCS asks for them, src/Utils.cpp:6-13).
// Out: chi2[n_toys] in toy order, lohi = the histogram range, counts[130] = TH1F contents incl. under/overflow.
int refs_chi2_test_model(void* vh, int n_toys, const double* toys, const double* model_vcs, const double* model_bcs, double* chi2,
                         double* lohi, float* counts) {
  auto* h = static_cast<Handle*>(vh);
  const int n = (int)h->s->getN();
  return guarded([&]() {
    register_model(n, h->s->ecm().data(), model_vcs, model_bcs);
    queue_toys(toys, (long)n_toys * n);
    Chi2TestModelArgs a;
    a.n = n_toys;
    a.initialChi2Ampl = 1.0;
    a.modelPath = "model.root";
    a.modelVCSName = "vcs";
    a.modelBCSName = "bcs";
    a.outputPath = "out.root";
    chi2TestModel(h->s.get(), a);
  }) || read_chi2_hist(chi2, lohi, counts);
}

// chi2TestData (src/Chi2Test.cpp:103-131): toys[2 n_toys][N] -- the first n_toys give the mean solution that serves as
// bcs0, the next n_toys the chi2 values.  bcs0_out[N] is recomputed here from the reference's solve() the same way.
int refs_chi2_test_data(void* vh, int n_toys, const double* toys, double* chi2, double* lohi, float* counts) {
  auto* h = static_cast<Handle*>(vh);
  const int n = (int)h->s->getN();
  return guarded([&]() {
    root_shim::registry().clear();
    queue_toys(toys, 2L * n_toys * n);
    Chi2TestArgs a;
    a.n = n_toys;
    a.initialChi2Ampl = 1.0;
    a.outputPath = "out.root";
    chi2TestData(h->s.get(), a);
  }) || read_chi2_hist(chi2, lohi, counts);
}

// ratioTestModel (src/RatioTest.cpp:13-62).  Out: means[N], mean_errors[N] (the "means" graph), counts[N][1026] and
// stats[N][3] = in-range count, sum x, sum x^2 of the bias_pt_<i> histograms.
int refs_ratio_test_model(void* vh, int n_toys, const double* toys, const double* model_vcs, const double* model_bcs, double* means,
                          double* mean_errors, float* counts, double* stats) {
  auto* h = static_cast<Handle*>(vh);
  const int n = (int)h->s->getN();
  const int rc = guarded([&]() {
    register_model(n, h->s->ecm().data(), model_vcs, model_bcs);
    queue_toys(toys, (long)n_toys * n);
    RatioTestModelArgs a;
    a.n = n_toys;
    a.modelPath = "model.root";
    a.modelVCSName = "vcs";
    a.modelBCSName = "bcs";
    a.outputPath = "out.root";
    ratioTestModel(h->s.get(), a);
  });
  if (rc) return rc;
  return guarded([&]() {
    auto& reg = root_shim::registry();
    auto* g = dynamic_cast<TGraphErrors*>(reg.at("means").get());
    for (int i = 0; i < n; ++i) {
      means[i] = g->GetY()[i];
      mean_errors[i] = g->GetEY()[i];
      auto* hst = dynamic_cast<TH1F*>(reg.at("bias_pt_" + std::to_string(i)).get());
      for (int b = 0; b < 1026; ++b) counts[(std::size_t)i * 1026 + b] = hst->counts()[(std::size_t)b];
      for (int k = 0; k < 3; ++k) stats[3 * i + k] = hst->stat(k);
    }
  });
}

// save() (src/ISRSolverSLE.cpp:97-122): what it writes, read back from the registry.  cov / inv_cov / G are ROW-major
// there (the reference transposes before SetMatrixArray), returned as written; bcs_err = sqrt(diag cov).
int refs_save(void* vh, double* bcs, double* bcs_err, double* G_rowmajor, double* cov_rowmajor, double* inv_cov_rowmajor) {
  auto* h = static_cast<Handle*>(vh);
  const int n = (int)h->s->getN();
  const int rc = guarded([&]() {
    root_shim::registry().clear();
    OutputOptions o;
    o.visibleCSGraphName = "vcs";
    o.bornCSGraphName = "bcs";
    h->s->save("out.root", o);
  });
  if (rc) return rc;
  return guarded([&]() {
    auto& reg = root_shim::registry();
    auto* g = dynamic_cast<TGraphErrors*>(reg.at("bcs").get());
    for (int i = 0; i < n; ++i) {
      bcs[i] = g->GetY()[i];
      bcs_err[i] = g->GetEY()[i];
    }
    const char* names[3] = {"intergalOperatorMatrix", "covMatrixBornCS", "invCovMatrixBornCS"};
    double* outs[3] = {G_rowmajor, cov_rowmajor, inv_cov_rowmajor};
    for (int k = 0; k < 3; ++k) {
      auto* m = dynamic_cast<TMatrixD*>(reg.at(names[k]).get());
      for (int q = 0; q < n * n; ++q) outs[k][q] = m->GetMatrixArray()[q];
    }
  });
}

}  // extern "C"
