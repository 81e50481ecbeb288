// ref_glue.cpp -- extern "C" access to the REFERENCE'S OWN compiled assembly-path code.
//
// ORACLE / TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load oracle/_ref/libisr_ref.so).
//
// oracle/Makefile compiles, unmodified and from where they lie under /root/reference/src:
//   KuraevFadin.cpp  Integration.cpp  BaseRangeInterpolator{,2}.cpp  LinearRangeInterpolator{,2}.cpp
//   CSplineRangeInterpolator{,2}.cpp  Interpolator{,2}.cpp
// against oracle/ref_shim (GSL API on the restated QUADPACK/cspline, Eigen::VectorXd as a plain vector) and
// the real nlohmann/json.hpp vendored in this image.  This file is compiled twice: once against
// Interpolator.hpp (efficiency(x, E), BaseISRSolver) and once with -DISR_REF_V2 against Interpolator2.hpp
// (efficiency(x, energyIndex), BaseISRSolver2) -- the two headers define the same exception type and
// cannot share a translation unit.
//
// Restated here, each with its reference line, because this file serves the multi-threaded row evaluation that the
// fixtures and the CPU baseline use (the reference's own versions of the same snippets run unmodified through
// ref_solver_glue.cpp / ref_solver2_glue.cpp since round 2, and tests/test_refrun_pin.py checks that both give the same
// matrices bit for bit): the efficiency adaptors (src/BaseISRSolver.cpp:232-258, src/BaseISRSolver2.cpp:231-236, ROOT bin
// arithmetic = orc_find_bin), the N x N fill loop of ISRSolverSLE::evalEqMatrix (src/ISRSolverSLE.cpp:138-145; SLE2:
// src/ISRSolverSLE2.cpp:142-149) and _energySpreadMatrix (:177-193).
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <functional>
#include <memory>
#include <thread>
#include <tuple>
#include <vector>
#include <atomic>

#include "Integration.hpp"
#include "KuraevFadin.hpp"
#ifdef ISR_REF_V2
#include "Interpolator2.hpp"
typedef Interpolator2 InterpT;
#define REFN(name) ref2_##name
#else
#include "Interpolator.hpp"
typedef Interpolator InterpT;
#define REFN(name) ref_##name
#endif

extern "C" int orc_find_bin(int nbins, double lo, double hi, const double* edges, double x);

// same layout as orc_eff in oracle/isr_oracle.c
struct ref_eff {
  int kind;  // 0: 1; 1: per-row x histogram (v2); 2: TEfficiency 2-D (x, E); 3: TEfficiency 1-D (E only)
  int nx; double xlo, xhi; const double* xedges;
  int ne; double elo, ehi; const double* eedges;
  const double* values;
};

#ifdef ISR_REF_V2
typedef std::function<double(double, std::size_t)> EffFn;
static EffFn make_eff(const ref_eff* e) {
  if (!e || e->kind == 0) return [](double, std::size_t) { return 1.; };           // src/BaseISRSolver2.cpp ctor default
  return [e](double x, std::size_t index) {                                         // :231-236
    const int bin = orc_find_bin(e->nx, e->xlo, e->xhi, e->xedges, x);              // TH1D::FindBin
    return e->values[index * (std::size_t)(e->nx + 2) + bin];                       // GetBinContent
  };
}
#else
typedef std::function<double(double, double)> EffFn;
static EffFn make_eff(const ref_eff* e) {
  if (!e || e->kind == 0) return [](double, double) { return 1.; };                 // src/BaseISRSolver.cpp:140
  if (e->kind == 3)
    return [e](double, double energy) {                                             // :236-241
      return e->values[orc_find_bin(e->ne, e->elo, e->ehi, e->eedges, energy)];
    };
  return [e](double x, double energy) {                                             // :243-250
    const int bx = orc_find_bin(e->nx, e->xlo, e->xhi, e->xedges, x);
    const int by = orc_find_bin(e->ne, e->elo, e->ehi, e->eedges, energy);
    return e->values[(std::size_t)by * (e->nx + 2) + bx];                           // global bin bx + (nx+2) by
  };
}
#endif

struct RefInterp {
  int n;
  Eigen::VectorXd ecm;
  std::unique_ptr<InterpT> interp;
};

template <typename Body>
static void parallel_for(long n, int nthreads, Body body) {
  if (nthreads < 1) nthreads = 1;
  std::atomic<long> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&]() {
      for (;;) {
        const long q = next.fetch_add(1);
        if (q >= n) break;
        body(q);
      }
    });
  for (auto& th : pool) th.join();
}

extern "C" {

#ifndef ISR_REF_V2
// kernelKuraevFadin (src/KuraevFadin.cpp:14-49)
double ref_kernel_kf(double x, double s) { return kernelKuraevFadin(x, s); }

// convolutionKuraevFadin (src/KuraevFadin.cpp:56-73) with a C callback as the Born cross section
double ref_convolution_kf(double energy, double (*fcn)(double, void*), void* ctx, double min_x, double max_x,
                          const ref_eff* eff) {
  std::function<double(double)> f = [fcn, ctx](double e) { return fcn(e, ctx); };
  return convolutionKuraevFadin(energy, f, min_x, max_x, make_eff(eff));
}

// integrate / integrateS / gaussian_conv (src/Integration.cpp:19-100)
double ref_integrate(double (*fcn)(double, void*), void* ctx, double a, double b, double* error) {
  std::function<double(double)> f = [fcn, ctx](double x) { return fcn(x, ctx); };
  return integrate(f, a, b, *error);
}
double ref_integrateS(double (*fcn)(double, void*), void* ctx, double a, double b, double* error) {
  std::function<double(double)> f = [fcn, ctx](double x) { return fcn(x, ctx); };
  return integrateS(f, a, b, *error);
}
double ref_gaussian_conv(double energy, double sigma2, double (*fcn)(double, void*), void* ctx) {
  std::function<double(double)> f = [fcn, ctx](double x) { return fcn(x, ctx); };
  return gaussian_conv(energy, sigma2, f);
}
#endif

// Interpolator(settings, cmEnergies, threshold) (src/Interpolator.cpp:34-83); nr == 0 -> the
// two-argument constructor (default settings, :23-26).  Returns NULL on InterpRangeException.
void* REFN(interp_create)(int n, const double* ecm, double threshold, int nr, const int* settings) {
  auto* h = new RefInterp();
  h->n = n;
  h->ecm = Eigen::VectorXd(n);
  for (int i = 0; i < n; ++i) h->ecm(i) = ecm[i];
  try {
    if (nr <= 0) {
      h->interp.reset(new InterpT(h->ecm, threshold));
    } else {
      std::vector<std::tuple<bool, int, int>> st;
      for (int r = 0; r < nr; ++r) st.emplace_back(settings[3 * r] != 0, settings[3 * r + 1], settings[3 * r + 2]);
      h->interp.reset(new InterpT(st, h->ecm, threshold));
    }
  } catch (const std::exception&) {
    delete h;
    return nullptr;
  }
  return h;
}
void REFN(interp_free)(void* vh) { delete static_cast<RefInterp*>(vh); }

double REFN(kf_basis_integral)(void* vh, int i, int j, const ref_eff* eff) {
  return static_cast<RefInterp*>(vh)->interp->evalKuraevFadinBasisIntegral(i, j, make_eff(eff));
}
double REFN(basis_eval)(void* vh, int j, double e) { return static_cast<RefInterp*>(vh)->interp->basisEval(j, e); }
double REFN(basis_deriv_eval)(void* vh, int j, double e) { return static_cast<RefInterp*>(vh)->interp->basisDerivEval(j, e); }
double REFN(integral_basis)(void* vh, int j) { return static_cast<RefInterp*>(vh)->interp->evalIntegralBasis(j); }
double REFN(interp_eval)(void* vh, const double* y, double e) {
  auto* h = static_cast<RefInterp*>(vh);
  Eigen::VectorXd yy(h->n);
  for (int i = 0; i < h->n; ++i) yy(i) = y[i];
  return h->interp->eval(yy, e);
}
double REFN(interp_deriv_eval)(void* vh, const double* y, double e) {
  auto* h = static_cast<RefInterp*>(vh);
  Eigen::VectorXd yy(h->n);
  for (int i = 0; i < h->n; ++i) yy(i) = y[i];
  return h->interp->derivEval(yy, e);
}
double REFN(min_energy)(void* vh) { return static_cast<RefInterp*>(vh)->interp->getMinEnergy(); }
double REFN(max_energy)(void* vh) { return static_cast<RefInterp*>(vh)->interp->getMaxEnergy(); }

// The fill loop of evalEqMatrix (src/ISRSolverSLE.cpp:141-145 / src/ISRSolverSLE2.cpp:145-149):
// G(i, j) = interp.evalKuraevFadinBasisIntegral(i, j, efficiency()); column-major like Eigen::MatrixXd.
// With the shim's stateless gsl_interp_accel the entries are independent, so rows can be evaluated on
// several threads (the reference itself is single-threaded); rows == NULL means all rows.
void REFN(eval_eq_matrix)(void* vh, const ref_eff* eff, double* G, int nrows, const int* rows, int nthreads) {
  auto* h = static_cast<RefInterp*>(vh);
  const int n = h->n;
  const EffFn ef = make_eff(eff);
  const int nr = rows ? nrows : n;
  parallel_for((long)nr * n, nthreads, [&](long q) {
    const int i = rows ? rows[q / n] : (int)(q / n);
    const int j = (int)(q % n);
    G[(std::size_t)i + (std::size_t)n * j] = h->interp->evalKuraevFadinBasisIntegral(i, j, ef);
  });
}

#ifndef ISR_REF_V2
// Same fill loop with the efficiency as a C callback eps(x, energy) -- the reference's std::function form
// (include/BaseISRSolver.hpp:36-41; a Python callable in PyISR, include/PyISRSolver.hpp:189-197).  Single-threaded:
// the callback may re-enter the interpreter.
void ref_eval_eq_matrix_cb(void* vh, double (*eff)(double, double, void*), void* ctx, double* G, int nrows, const int* rows) {
  auto* h = static_cast<RefInterp*>(vh);
  const int n = h->n;
  const EffFn ef = [eff, ctx](double x, double e) { return eff(x, e, ctx); };
  const int nr = rows ? nrows : n;
  for (int a = 0; a < nr; ++a) {
    const int i = rows ? rows[a] : a;
    for (int j = 0; j < n; ++j) G[(std::size_t)i + (std::size_t)n * j] = h->interp->evalKuraevFadinBasisIntegral(i, j, ef);
  }
}
#endif

// _energySpreadMatrix (src/ISRSolverSLE.cpp:177-193): S(i, j) = gaussian_conv(ecm_i, ecmErr_i^2, basis_j)
void REFN(energy_spread_matrix)(void* vh, const double* ecm_err, double* S) {
  auto* h = static_cast<RefInterp*>(vh);
  const int n = h->n;
  std::size_t j = 0;
  std::function<double(double)> fcn = [&j, h](double energy) { return h->interp->basisEval((int)j, energy); };
  for (int i = 0; i < n; ++i)
    for (j = 0; j < (std::size_t)n; ++j) {
      const double sigma2 = std::pow(ecm_err[i], 2);
      S[(std::size_t)i + (std::size_t)n * j] = gaussian_conv(h->ecm(i), sigma2, fcn);
    }
}

}  // extern "C"
