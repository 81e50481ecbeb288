// isr_b200.hpp -- header-only C++ mirror of the reference's solver classes for the hot path, on top of the
// C ABI (isr_b200.h).  Same class and method names as the reference (include/BaseISRSolver.hpp,
// include/ISRSolverSLE.hpp, include/ISRSolverTikhonov.hpp, include/ISRSolverTSVD.hpp, include/Chi2Test.hpp,
// include/RatioTest.hpp); Eigen/ROOT types are replaced by std::vector and the small Matrix below
// (column-major, like Eigen::MatrixXd).  Errors of the ABI are rethrown as the reference's exception types.
//
//   g++ -std=c++17 -Iinclude app.cpp -Lisrsolver_b200 -lisr_b200 -Wl,-rpath,$PWD/isrsolver_b200
#ifndef ISR_B200_HPP
#define ISR_B200_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "isr_b200.h"

namespace isrb200 {

struct InterpRangeException : std::exception {          // include/Interpolator.hpp:14-18
  const char* what() const noexcept override { return "[!] Wrong interpolation ranges.\n"; }
};
struct EfficiencyDimensionException : std::exception {  // include/BaseISRSolver.hpp:24-28
  const char* what() const noexcept override { return "[!] Wrong efficiency dimension.\n"; }
};
struct DifferentModelCrossSectionSizes : std::exception {  // include/Chi2Test.hpp:47-52
  const char* what() const noexcept override { return "[!] Model visible and Born cross sections have different sizes.\n"; }
};
struct InterpSettingsSizeException : std::exception {  // include/ISRSolverStructs.hpp:170-174 (declared there, thrown nowhere)
  const char* what() const noexcept override { return "[!] Wrong size of interpolation settings vector.\n"; }
};
struct IsrGpuError : std::runtime_error { using std::runtime_error::runtime_error; };

inline void check(int rc) {
  if (rc == ISR_OK) return;
  if (rc == ISR_ERANGE) throw InterpRangeException();
  if (rc == ISR_EEFFDIM) throw EfficiencyDimensionException();
  if (rc == ISR_EINVAL) throw std::invalid_argument(isr_last_error());
  throw IsrGpuError(isr_last_error());
}

// Column-major dense matrix (Eigen::MatrixXd layout).
struct Matrix {
  int rows = 0, cols = 0;
  std::vector<double> a;
  Matrix() = default;
  Matrix(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i + (size_t)rows * j]; }
  double operator()(int i, int j) const { return a[(size_t)i + (size_t)rows * j]; }
  double* data() { return a.data(); }
  const double* data() const { return a.data(); }
};

// Detection-efficiency table (what BaseISRSolver::_setupEfficiency wraps; src/BaseISRSolver.cpp:232-258).
struct Efficiency {
  int kind = ISR_EPS_ONE;
  int nx = 0, ne = 0;
  double xlo = 0, xhi = 1, elo = 0, ehi = 1;
  std::vector<double> xedges, eedges, values;
  isr_eps_desc desc() const {
    isr_eps_desc d{};
    d.kind = kind; d.nx = nx; d.xlo = xlo; d.xhi = xhi; d.ne = ne; d.elo = elo; d.ehi = ehi;
    d.xedges = xedges.empty() ? nullptr : xedges.data();
    d.eedges = eedges.empty() ? nullptr : eedges.data();
    d.values = values.empty() ? nullptr : values.data();
    return d;
  }
};

class Context {
 public:
  static isr_ctx* get(int device = 0) {
    static std::vector<isr_ctx*> ctxs(64, nullptr);
    if (!ctxs.at(device)) check(isr_ctx_create(device, &ctxs[device]));
    return ctxs[device];
  }
};

// One isr_group (all requested GPUs of the box, one host thread + context each, NCCL communicator) per process and size.
class GpuGroup {
 public:
  static isr_group* get(int ngpu) {
    static std::vector<isr_group*> groups(65, nullptr);
    if (ngpu < 1 || ngpu > 64) throw std::invalid_argument("GpuGroup: ngpu out of range");
    if (!groups[ngpu]) check(isr_group_create(ngpu, nullptr, &groups[ngpu]));
    return groups[ngpu];
  }
};

struct Chi2TestArgs { int n; double initialChi2Ampl; std::string outputPath; };   // include/ISRSolverStructs.hpp
struct Chi2TestResult { std::vector<double> chi2; std::vector<float> hist; double chi2Min, chi2Max; };
struct RatioTestResult { std::vector<double> means, meanErrors; std::vector<uint32_t> hist; };

class BaseISRSolver {   // include/BaseISRSolver.hpp:30-120
 public:
  // src/BaseISRSolver.cpp:77-129: copy, sort by energy through an index permutation; energyErr == nullptr
  // switches the energy spread off.
  BaseISRSolver(std::size_t numberOfPoints, const double* energy, const double* visibleCS, const double* energyErr,
                const double* visibleCSErr, double thresholdEnergy, const Efficiency& efficiency = Efficiency(),
                int device = 0)
      : _n(numberOfPoints), _energyT(thresholdEnergy), _energySpread(energyErr != nullptr), _eff(efficiency) {
    std::vector<int> index(_n);
    std::iota(index.begin(), index.end(), 0);
    std::sort(index.begin(), index.end(), [energy](int a, int b) { return energy[a] < energy[b]; });
    _ecm.resize(_n); _ecmErr.assign(_n, 0.0); _vcs.resize(_n); _vcsErr.resize(_n); _bcs.assign(_n, 0.0);
    for (std::size_t i = 0; i < _n; ++i) {
      _ecm[i] = energy[index[i]]; _vcs[i] = visibleCS[index[i]]; _vcsErr[i] = visibleCSErr[index[i]];
      if (energyErr) _ecmErr[i] = energyErr[index[i]];
    }
    _ctx = Context::get(device);
    isr_eps_desc d = _eff.desc();
    check(isr_problem_create(_ctx, (int)_n, _ecm.data(), _energyT, 0, nullptr, _eff.kind == ISR_EPS_ONE ? nullptr : &d, &_p));
  }
  // The reference's own constructor signature (include/BaseISRSolver.hpp:33-40): the efficiency is an arbitrary callable
  // eps(x, E).  It cannot run inside a kernel, so it is SAMPLED: evalEqMatrix asks the library for the quadrature nodes
  // (isr_assemble_nodes), evaluates the callable there on the host and feeds the values back (isr_assemble_sampled).
  // efficiencyBreaks: x values where the callable jumps (the panels are cut there; smooth callables need none).
  BaseISRSolver(std::size_t numberOfPoints, const double* energy, const double* visibleCS, const double* energyErr,
                const double* visibleCSErr, double thresholdEnergy, const std::function<double(double, double)>& efficiency,
                const std::vector<double>& efficiencyBreaks = {}, int device = 0)
      : BaseISRSolver(numberOfPoints, energy, visibleCS, energyErr, visibleCSErr, thresholdEnergy,
                      breaksTable(numberOfPoints, efficiencyBreaks), device) {
    _effFcn = efficiency;
  }
  BaseISRSolver(const BaseISRSolver&) = delete;
  BaseISRSolver& operator=(const BaseISRSolver&) = delete;
  virtual ~BaseISRSolver() { isr_problem_destroy(_p); }
  virtual void solve() = 0;
  std::size_t getN() const { return _n; }
  double getThresholdEnergy() const { return _energyT; }
  double getMinEnergy() const { return _ecm.front(); }
  double getMaxEnergy() const { return _ecm.back(); }
  const std::vector<double>& bcs() const { return _bcs; }
  const std::vector<double>& ecm() const { return _ecm; }
  const std::vector<double>& ecmErr() const { return _ecmErr; }
  const std::vector<double>& vcs() const { return _vcs; }
  const std::vector<double>& vcsErr() const { return _vcsErr; }
  bool isEnergySpreadEnabled() const { return _energySpread; }
  void enableEnergySpread() { _energySpread = true; invalidate(); }
  void disableEnergySpread() { _energySpread = false; invalidate(); }
  void resetVisibleCS(const std::vector<double>& v) { _vcs = v; }
  void resetVisibleCSErrors(const std::vector<double>& e) { _vcsErr = e; _factored = false; }
  void resetECMErrors(const std::vector<double>& e) { _ecmErr = e; invalidate(); }
  isr_problem* gpuProblem() { return _p; }
  const Efficiency& efficiencyTable() const { return _eff; }
  bool hasCallableEfficiency() const { return (bool)_effFcn; }

 protected:
  // a per-point table of ones whose bin edges are the callable's declared jumps (its values are ignored in sampled mode)
  static Efficiency breaksTable(std::size_t n, const std::vector<double>& breaks) {
    Efficiency e;
    if (breaks.empty()) return e;
    std::vector<double> b(breaks);
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());
    e.kind = ISR_EPS_ROW_TABLE;
    e.xedges.push_back(std::min(0.0, b.front()) - 1.0);
    for (double v : b) e.xedges.push_back(v);
    e.xedges.push_back(std::max(1.0, b.back()) + 1.0);
    e.nx = (int)e.xedges.size() - 1;
    e.xlo = e.xedges.front(); e.xhi = e.xedges.back();
    e.values.assign(n * (std::size_t)(e.nx + 2), 1.0);
    return e;
  }
  std::function<double(double, double)> _effFcn;
  virtual void invalidate() { _isEqMatrixPrepared = false; _factored = false; }
  std::size_t _n;
  double _energyT;
  bool _energySpread;
  Efficiency _eff;
  std::vector<double> _ecm, _ecmErr, _vcs, _vcsErr, _bcs;
  isr_ctx* _ctx = nullptr;
  isr_problem* _p = nullptr;
  bool _isEqMatrixPrepared = false, _factored = false;
};

class ISRSolverSLE : public BaseISRSolver {   // include/ISRSolverSLE.hpp
 public:
  using BaseISRSolver::BaseISRSolver;
  void setRangeInterpSettings(const std::vector<std::tuple<bool, int, int>>& s) {   // src/ISRSolverSLE.cpp:125-128
    std::vector<int> flat;
    for (const auto& t : s) { flat.push_back(std::get<0>(t)); flat.push_back(std::get<1>(t)); flat.push_back(std::get<2>(t)); }
    check(isr_problem_set_ranges(_p, (int)s.size(), flat.data()));
    _ranges = flat;
    invalidate();
  }
  const std::vector<int>& rangeSettings() const { return _ranges; }
  void evalEqMatrix() {   // src/ISRSolverSLE.cpp:138-149
    _integralOperatorMatrix = Matrix((int)_n, (int)_n);
    const double* spread = _energySpread ? _ecmErr.data() : nullptr;
    if (_effFcn) {      // callable efficiency: sampled at the quadrature nodes
      int64_t nn = 0;
      check(isr_assemble_nodes(_p, 0, nullptr, nullptr, &nn));
      std::vector<double> x((std::size_t)nn), eps((std::size_t)nn);
      std::vector<int> row((std::size_t)nn);
      check(isr_assemble_nodes(_p, nn, x.data(), row.data(), &nn));
      for (int64_t q = 0; q < nn; ++q) eps[(std::size_t)q] = _effFcn(x[(std::size_t)q], _ecm[(std::size_t)row[(std::size_t)q]]);
      check(isr_assemble_sampled(_p, nn, eps.data(), spread, _integralOperatorMatrix.data()));
    } else {
      check(isr_assemble(_p, spread, _integralOperatorMatrix.data()));
    }
    _isEqMatrixPrepared = true;
    _factored = false;
  }
  void solve() override {   // src/ISRSolverSLE.cpp:86-95
    ensureFactored();
    _covMatrixBornCS = Matrix((int)_n, (int)_n);
    check(isr_solve(_p, _vcs.data(), _bcs.data(), _covMatrixBornCS.data()));
  }
  const Matrix& getIntegralOperatorMatrix() const { return _integralOperatorMatrix; }
  const Matrix& getBornCSCovMatrix() const { return _covMatrixBornCS; }
  double interpEval(const std::vector<double>& y, double energy) const {   // :195-198
    double out = 0;
    check(isr_interp_eval(_p, y.data(), 1, &energy, &out));
    return out;
  }
  std::vector<double> interpEval(const std::vector<double>& y, const std::vector<double>& energies) const {
    std::vector<double> out(energies.size());
    if (!energies.empty()) check(isr_interp_eval(_p, y.data(), (int64_t)energies.size(), energies.data(), out.data()));
    return out;
  }
  // inverse covariance G^T W G (PyISR bcs_inv_cov_matrix, include/PyISRSolverSLE.hpp:72-86)
  virtual Matrix evalBornCSInvCovMatrix() {
    ensureFactored();
    Matrix m((int)_n, (int)_n);
    check(isr_get_inv_cov(_p, m.data()));
    return m;
  }
  double evalConditionNumber() {   // :200-204
    if (!_isEqMatrixPrepared) evalEqMatrix();
    double c = 0;
    check(isr_cond_number(_p, &c));
    return c;
  }
  // Condition number versus beam-energy spread: the loop of src/isrsolver-condnum-test.cpp:123-135 (sigma = 0
  // means "spread disabled") from ONE assembly.  Invalidates the cached matrix.
  std::vector<double> evalConditionNumberSpreadScan(const std::vector<double>& sigmas) {
    std::vector<double> cond(sigmas.size());
    check(isr_cond_spread_scan(_p, (int)sigmas.size(), sigmas.data(), 0, cond.data(), nullptr));
    invalidate();
    return cond;
  }
  // toy-kernel schedule (isr_set_toy_kernel): 0 default, 1 dense cross-check, 2 triangularised chi2 factor
  void setToySchedule(int variant) { check(isr_set_toy_kernel(_p, variant)); }
  // the toy loop body for T toys (toy-major V); any output may be nullptr
  virtual void toyBatch(int64_t T, const double* V, const double* bcs0, double* chi2, double* sumBcs, double* ratioStats,
                        uint32_t* ratioHist, double* bcsOut) {
    selectForToys();
    check(isr_toy_batch(_p, T, V, bcs0, chi2, sumBcs, ratioStats, ratioHist, bcsOut));
  }
  // a whole chi2 / ratio test on the selected solver in one stream-ordered call (isr_chi2_test)
  virtual void chi2TestCall(isr_chi2_test_args& a) {
    selectForToys();
    a.flags &= ~(ISR_TEST_ASSEMBLE | ISR_TEST_FACTOR);      // the solver's own state is current
    check(isr_chi2_test(_p, &a));
  }
  // true if the solver is fully described by (energies, ranges, efficiency table, spread, errors): such a solver can be
  // replicated on the GPUs of an isr_group
  virtual bool replicable() const { return !_effFcn; }

 protected:
  virtual void ensureFactored() {
    if (!_isEqMatrixPrepared) evalEqMatrix();
    if (!_factored) { check(isr_factor(_p, _vcsErr.data())); _factored = true; }
  }
  virtual void selectForToys() { ensureFactored(); }
  Matrix _integralOperatorMatrix, _covMatrixBornCS;
  std::vector<int> _ranges;
};

// IterISRInterpSolver (include/IterISRInterpSolver.hpp; src/IterISRInterpSolver.cpp:45-65): the iterative
// radiative-correction method on the same operator matrix; Cov = diag((vcsErr / radcorr)^2).
class IterISRInterpSolver : public ISRSolverSLE {
 public:
  using ISRSolverSLE::ISRSolverSLE;
  void setNumOfIters(std::size_t n) { _nIter = n; }
  std::size_t getNumOfIters() const { return _nIter; }
  const std::vector<double>& getRadCorr() const { return _radcorr; }
  void solve() override {
    if (!_isEqMatrixPrepared) evalEqMatrix();
    _radcorr.assign(_n, 0.0);
    check(isr_iterative_solve(_p, (int)_nIter, _vcs.data(), _bcs.data(), _radcorr.data()));
    _covMatrixBornCS = Matrix((int)_n, (int)_n);
    if (_nIter > 0)
      for (std::size_t i = 0; i < _n; ++i) _covMatrixBornCS((int)i, (int)i) = (_vcsErr[i] / _radcorr[i]) * (_vcsErr[i] / _radcorr[i]);
  }
  Matrix evalBornCSInvCovMatrix() override {
    Matrix m((int)_n, (int)_n);
    for (std::size_t i = 0; i < _n; ++i) m((int)i, (int)i) = 1.0 / _covMatrixBornCS((int)i, (int)i);
    return m;
  }

 private:
  std::size_t _nIter = 10;   // src/IterISRInterpSolver.cpp:19
  std::vector<double> _radcorr;
};

// ISRSolverSLE2 (include/ISRSolverSLE2.hpp; src/BaseISRSolver2.cpp:151-160, 231-236): the efficiency is one
// x-histogram per energy point, effHists[point][0..nbins+1] with ROOT's under/overflow cells.
class ISRSolverSLE2 : public ISRSolverSLE {
 public:
  ISRSolverSLE2(std::size_t n, const double* energy, const double* visibleCS, const double* energyErr,
                const double* visibleCSErr, double thresholdEnergy, const std::vector<std::vector<double>>& effHists,
                double xlo = 0.0, double xhi = 1.0, int device = 0)
      : ISRSolverSLE(n, energy, visibleCS, energyErr, visibleCSErr, thresholdEnergy,
                     rowTable(n, energy, effHists, xlo, xhi), device) {}

 private:
  static Efficiency rowTable(std::size_t n, const double* energy, const std::vector<std::vector<double>>& h, double xlo,
                             double xhi) {
    if (h.size() != n || h.empty() || h[0].size() < 3) throw std::invalid_argument("ISRSolverSLE2: one histogram per energy point");
    std::vector<int> index(n);
    std::iota(index.begin(), index.end(), 0);
    std::sort(index.begin(), index.end(), [energy](int a, int b) { return energy[a] < energy[b]; });
    Efficiency e;
    e.kind = ISR_EPS_ROW_TABLE;
    e.nx = (int)h[0].size() - 2;
    e.xlo = xlo;
    e.xhi = xhi;
    for (std::size_t i = 0; i < n; ++i) {     // histograms follow their points through the energy sort
      if (h[index[i]].size() != h[0].size()) throw std::invalid_argument("ISRSolverSLE2: histograms differ in size");
      e.values.insert(e.values.end(), h[index[i]].begin(), h[index[i]].end());
    }
    return e;
  }
};

class ISRSolverTikhonov : public ISRSolverSLE {   // include/ISRSolverTikhonov.hpp
 public:
  using ISRSolverSLE::ISRSolverSLE;
  bool replicable() const override { return false; }      // (the selected A+_lambda lives in this problem only)
  double getLambda() const { return _lambda; }
  void setLambda(double l) { _lambda = l; }
  bool isDerivNorm2RegIsEnabled() const { return _derivReg; }
  void enableDerivNorm2Regularizator() { _derivReg = true; _tik = false; }
  void disableDerivNorm2Regularizator() { _derivReg = false; _tik = false; }
  void solve() override {   // src/ISRSolverTikhonov.cpp:61-72
    ensureTik();
    _covMatrixBornCS = Matrix((int)_n, (int)_n);
    check(isr_tikhonov_solve(_p, _lambda, _vcs.data(), _bcs.data(), _covMatrixBornCS.data()));
  }
  double evalEqNorm2() { return scan1(0); }                      // :106-109
  double evalSmoothnessConstraintNorm2() { return scan1(1); }    // :111-113
  double evalLCurveCurvature() { return scan1(2); }              // :115-119
  double evalLCurveCurvatureDerivative() { return scan1(3); }    // :121-128
  // every lambda of an L-curve scan (src/isrsolver-LCurve-plot.cpp:131-151) from one decomposition
  void scanLCurve(const std::vector<double>& lambdas, std::vector<double>& x, std::vector<double>& y,
                  std::vector<double>& curv, std::vector<double>& dcurv) {
    ensureTik();
    const int64_t L = (int64_t)lambdas.size();
    x.resize(L); y.resize(L); curv.resize(L); dcurv.resize(L);
    check(isr_tikhonov_scan(_p, L, lambdas.data(), _vcs.data(), x.data(), y.data(), curv.data(), dcurv.data(), nullptr));
  }
  // chi2 of every (lambda, toy) pair, row-major L x T: the chi2Test loop of isrsolver-Tikhonov-chi2-test for all
  // lambdas at once
  std::vector<double> toyScan(const std::vector<double>& lambdas, int64_t T, const double* V, const std::vector<double>& bcs0) {
    ensureTik();
    std::vector<double> chi2(lambdas.size() * (std::size_t)T);
    check(isr_tikhonov_toy_scan(_p, (int64_t)lambdas.size(), lambdas.data(), T, V, bcs0.data(), chi2.data()));
    return chi2;
  }
  // maximise the L-curve curvature (isrsolver-Tikhonov-LCurve, src/isrsolver-Tikhonov-LCurve.cpp:111-133; NLopt
  // LD_MMA there): dense log-grid scan of the O(N)-per-lambda curvature + golden-section refinement
  double solveLCurve(double lo = 1e-12, double hi = 1.0) {
    std::vector<double> grid(2048), x, y, c, dc;
    for (std::size_t q = 0; q < grid.size(); ++q) grid[q] = lo * std::pow(hi / lo, (double)q / (grid.size() - 1));
    scanLCurve(grid, x, y, c, dc);
    const std::size_t k = std::min_element(c.begin(), c.end()) - c.begin();   // curvature is stored as -|kappa|
    double a = grid[k ? k - 1 : 0], b = grid[std::min(k + 1, grid.size() - 1)];
    const double gr = (std::sqrt(5.0) - 1.0) / 2.0;
    auto f = [this](double l) { const double keep = _lambda; _lambda = l; const double v = scan1(2); _lambda = keep; return v; };
    double p1 = b - gr * (b - a), p2 = a + gr * (b - a), f1 = f(p1), f2 = f(p2);
    while (std::fabs(b - a) > 1e-6 * std::fabs(0.5 * (a + b))) {
      if (f1 < f2) { b = p2; p2 = p1; f2 = f1; p1 = b - gr * (b - a); f1 = f(p1); }
      else { a = p1; p1 = p2; f1 = f2; p2 = a + gr * (b - a); f2 = f(p2); }
    }
    _lambda = 0.5 * (a + b);
    solve();
    return _lambda;
  }

 protected:
  void invalidate() override { ISRSolverSLE::invalidate(); _tik = false; }
  void ensureFactored() override { ensureTik(); }
  void selectForToys() override { ensureTik(); check(isr_tikhonov_select(_p, _lambda)); }
  void ensureTik() {
    if (!_isEqMatrixPrepared) { evalEqMatrix(); _tik = false; }
    if (!_tik) { check(isr_tikhonov_prepare(_p, _vcsErr.data(), _derivReg ? 1 : 0)); _tik = true; _factored = true; }
  }
  double scan1(int which) {
    ensureTik();
    double out[4];
    check(isr_tikhonov_scan(_p, 1, &_lambda, _vcs.data(), &out[0], &out[1], &out[2], &out[3], nullptr));
    return out[which];
  }
  double _lambda = 1.0;    // src/ISRSolverTikhonov.cpp:30
  bool _derivReg = true;   // :29
  bool _tik = false;
};

class ISRSolverTSVD : public ISRSolverSLE {   // include/ISRSolverTSVD.hpp
 public:
  ISRSolverTSVD(std::size_t n, const double* e, const double* v, const double* ee, const double* ve, double thr,
                const Efficiency& eff = Efficiency(), int device = 0)
      : ISRSolverSLE(n, e, v, ee, ve, thr, eff, device), _upperTSVDIndex((int)n) {}
  ISRSolverTSVD(std::size_t n, const double* e, const double* v, const double* ee, const double* ve, double thr,
                const std::function<double(double, double)>& eff, const std::vector<double>& effBreaks = {}, int device = 0)
      : ISRSolverSLE(n, e, v, ee, ve, thr, eff, effBreaks, device), _upperTSVDIndex((int)n) {}
  bool replicable() const override { return false; }
  void setUpperTSVDIndex(int k) { _upperTSVDIndex = k; }
  int getUpperTSVDIndex() const { return _upperTSVDIndex; }
  void enableKeepOne() { _keepOne = true; }
  void disableKeepOne() { _keepOne = false; }
  void solve() override {   // src/ISRSolverTSVD.cpp:53-74
    if (!_isEqMatrixPrepared) evalEqMatrix();
    const bool full = _upperTSVDIndex == (int)_n && !_keepOne;
    if (full) _covMatrixBornCS = Matrix((int)_n, (int)_n);
    check(isr_tsvd_solve(_p, _upperTSVDIndex, _keepOne ? 1 : 0, _vcs.data(), _vcsErr.data(), _bcs.data(),
                         full ? _covMatrixBornCS.data() : nullptr));
    _factored = false;
  }

 protected:
  void selectForToys() override {
    if (!_isEqMatrixPrepared) evalEqMatrix();
    check(isr_tsvd_select(_p, _upperTSVDIndex, _keepOne ? 1 : 0, _vcsErr.data()));
    _factored = false;
  }
  int _upperTSVDIndex;
  bool _keepOne = false;
};

// kernelKuraevFadin(x, s) (include/KuraevFadin.hpp:13) evaluated on the device.
inline double kernelKuraevFadin(double x, double s, int device = 0) {
  double out = 0;
  check(isr_kernel_kf(Context::get(device), 1, &x, &s, &out));
  return out;
}

// convolutionKuraevFadin for many energies at once (src/KuraevFadin.cpp:56-73 looped as
// src/isrsolver-KuraevFadin-convolution.cpp:122-134 does): the quadrature runs on the GPU, fcn is sampled on the
// host at the node energies the plan asks for, panels are bisected until every energy meets
// max(absTol, relTol |result|).  errEst (optional) receives the Gauss-Kronrod error estimates.
inline std::vector<double> convolutionKuraevFadin(const std::vector<double>& energy, const std::function<double(double)>& fcn,
                                                  const std::vector<double>& minX, const std::vector<double>& maxX,
                                                  const Efficiency& efficiency = Efficiency(), double relTol = 1e-10,
                                                  double absTol = 1e-12, std::vector<double>* errEst = nullptr,
                                                  const std::vector<double>& breakEnergies = {}, int device = 0) {
  if (minX.size() != energy.size() || maxX.size() != energy.size()) throw std::invalid_argument("convolutionKuraevFadin: sizes differ");
  isr_eps_desc d = efficiency.desc();
  isr_conv_plan* plan = nullptr;
  check(isr_conv_plan_create(Context::get(device), (int64_t)energy.size(), energy.data(), minX.data(), maxX.data(),
                             efficiency.kind == ISR_EPS_ONE ? nullptr : &d, (int)breakEnergies.size(),
                             breakEnergies.empty() ? nullptr : breakEnergies.data(), &plan));
  struct Guard { isr_conv_plan* p; ~Guard() { isr_conv_plan_destroy(p); } } guard{plan};
  std::vector<double> nodes, f;
  for (int level = 0; level < 40; ++level) {
    int64_t npan = 0, nn = 0;
    check(isr_conv_plan_size(plan, &npan, &nn));
    if (nn) {
      nodes.resize(nn); f.resize(nn);
      check(isr_conv_plan_nodes(plan, nodes.data(), nullptr, nullptr, nullptr));
      for (int64_t q = 0; q < nn; ++q) f[q] = fcn(nodes[q]);
      check(isr_conv_plan_feed(plan, f.data()));
    }
    int64_t nsplit = 0;
    check(isr_conv_plan_refine(plan, relTol, absTol, &nsplit));
    if (!nsplit) break;
  }
  int64_t npan = 0, nn = 0;
  check(isr_conv_plan_size(plan, &npan, &nn));
  if (nn) {   // level cap reached with children pending: evaluate them and report the estimate
    nodes.resize(nn); f.resize(nn);
    check(isr_conv_plan_nodes(plan, nodes.data(), nullptr, nullptr, nullptr));
    for (int64_t q = 0; q < nn; ++q) f[q] = fcn(nodes[q]);
    check(isr_conv_plan_feed(plan, f.data()));
  }
  std::vector<double> res(energy.size());
  if (errEst) errEst->resize(energy.size());
  check(isr_conv_plan_result(plan, res.data(), errEst ? errEst->data() : nullptr));
  return res;
}
// the reference's scalar signature (include/KuraevFadin.hpp:5-9)
inline double convolutionKuraevFadin(double energy, const std::function<double(double)>& fcn, double min_x, double max_x,
                                     const Efficiency& efficiency = Efficiency()) {
  return convolutionKuraevFadin(std::vector<double>{energy}, fcn, {min_x}, {max_x}, efficiency)[0];
}

// randomDrawVisCS (src/Utils.cpp:6-13) for n toys at once, toy-major n x N.  The reference draws from the
// unseeded global gRandom; here the stream is a seeded std::mt19937_64 so that campaigns are reproducible.
inline std::vector<double> randomDrawVisCS(const std::vector<double>& vcs, const std::vector<double>& vcsErr, int n,
                                           std::uint64_t seed = 20211103ull) {
  std::mt19937_64 gen(seed);
  std::normal_distribution<double> gaus(0.0, 1.0);
  const std::size_t N = vcs.size();
  std::vector<double> V((std::size_t)n * N);
  for (int k = 0; k < n; ++k)
    for (std::size_t i = 0; i < N; ++i) V[(std::size_t)k * N + i] = vcs[i] + vcsErr[i] * gaus(gen);
  return V;
}

// chi2Test (src/Chi2Test.cpp:29-104) on caller-supplied toys (toy-major n x N); the ROOT fit/file are left to
// the caller.  The histogram is TH1F(128, min, max) with under/overflow cells (130 floats).  One isr_chi2_test call:
// toys, range and cells in one stream-ordered chain.
inline Chi2TestResult chi2Test(ISRSolverSLE* solver, const Chi2TestArgs& args, const std::vector<double>& toys,
                               const std::vector<double>& bcs0) {
  Chi2TestResult r;
  r.chi2.resize(args.n);
  std::vector<int64_t> counts(130);
  double lohi[2] = {0, 0};
  isr_chi2_test_args a{};
  a.bcs0 = bcs0.data(); a.n_toys = args.n; a.V = toys.data(); a.nbins = 128;
  a.chi2_out = r.chi2.data(); a.chi2_lo_hi_out = lohi; a.chi2_counts_out = counts.data();
  solver->chi2TestCall(a);
  r.chi2Min = lohi[0]; r.chi2Max = lohi[1];
  r.hist.assign(counts.begin(), counts.end());
  return r;
}

// The same test spread over the first `ngpu` GPUs of the box from this one process (isr_group: a context, a replicated
// problem and a host thread per GPU; toys cut into contiguous slices; two small NCCL all-reduces in stream).  The
// histogram is bit-identical to the single-GPU one.  Plain SLE solvers only (ISRSolverSLE / ISRSolverSLE2 with tabulated
// efficiency): the problem is rebuilt from its description on every GPU.
inline Chi2TestResult chi2TestMultiGPU(ISRSolverSLE* solver, const Chi2TestArgs& args, const std::vector<double>& toys,
                                       const std::vector<double>& bcs0, int ngpu, std::vector<double>* sumBcs = nullptr) {
  if (!solver->replicable()) throw std::invalid_argument("chi2TestMultiGPU: only plain SLE solvers with a tabulated efficiency can be replicated");
  isr_group* g = GpuGroup::get(ngpu);
  const std::size_t N = solver->getN();
  isr_eps_desc d = solver->efficiencyTable().desc();
  isr_gproblem* gp = nullptr;
  const std::vector<int>& rg = solver->rangeSettings();
  check(isr_group_problem_create(g, (int)N, solver->ecm().data(), solver->getThresholdEnergy(), (int)rg.size() / 3,
                                 rg.empty() ? nullptr : rg.data(), solver->efficiencyTable().kind == ISR_EPS_ONE ? nullptr : &d, &gp));
  struct Guard { isr_gproblem* p; ~Guard() { isr_group_problem_destroy(p); } } guard{gp};
  Chi2TestResult r;
  r.chi2.resize(args.n);
  std::vector<int64_t> counts(130);
  double lohi[2] = {0, 0};
  isr_chi2_test_args a{};
  a.flags = ISR_TEST_ASSEMBLE | ISR_TEST_FACTOR;
  a.ecm_err = solver->isEnergySpreadEnabled() ? solver->ecmErr().data() : nullptr;
  a.vcs_err = solver->vcsErr().data();
  a.bcs0 = bcs0.data(); a.n_toys = args.n; a.V = toys.data(); a.nbins = 128;
  a.chi2_out = r.chi2.data(); a.chi2_lo_hi_out = lohi; a.chi2_counts_out = counts.data();
  if (sumBcs) { sumBcs->assign(N, 0.0); a.sum_bcs_out = sumBcs->data(); }
  check(isr_group_chi2_test(gp, &a));
  r.chi2Min = lohi[0]; r.chi2Max = lohi[1];
  r.hist.assign(counts.begin(), counts.end());
  return r;
}

// chi2TestModel (src/Chi2Test.cpp:113-151): solve, then chi2Test of toys drawn around the model visible cross
// section against the model Born cross section (the reference reads both from a ROOT file).
inline Chi2TestResult chi2TestModel(ISRSolverSLE* solver, const Chi2TestArgs& args, const std::vector<double>& modelVCS,
                                    const std::vector<double>& modelBCS, std::uint64_t seed = 20211103ull) {
  if (modelVCS.size() != modelBCS.size() || modelVCS.size() != solver->getN()) throw DifferentModelCrossSectionSizes();
  solver->solve();
  return chi2Test(solver, args, randomDrawVisCS(modelVCS, solver->vcsErr(), args.n, seed), modelBCS);
}
// chi2TestModel over `ngpu` GPUs: same toys (same seed), same histogram.
inline Chi2TestResult chi2TestModel(ISRSolverSLE* solver, const Chi2TestArgs& args, const std::vector<double>& modelVCS,
                                    const std::vector<double>& modelBCS, std::uint64_t seed, int ngpu) {
  if (modelVCS.size() != modelBCS.size() || modelVCS.size() != solver->getN()) throw DifferentModelCrossSectionSizes();
  if (ngpu <= 1) return chi2TestModel(solver, args, modelVCS, modelBCS, seed);
  solver->solve();
  return chi2TestMultiGPU(solver, args, randomDrawVisCS(modelVCS, solver->vcsErr(), args.n, seed), modelBCS, ngpu);
}
// chi2TestData (src/Chi2Test.cpp:160-195): the reference solution is the mean of the toy solutions around the data.
inline Chi2TestResult chi2TestData(ISRSolverSLE* solver, const Chi2TestArgs& args, std::uint64_t seed = 20211103ull) {
  solver->solve();
  const std::size_t N = solver->getN();
  std::vector<double> sum(N, 0.0), ones(N, 1.0);
  solver->toyBatch(args.n, randomDrawVisCS(solver->vcs(), solver->vcsErr(), args.n, seed).data(), ones.data(), nullptr,
                   sum.data(), nullptr, nullptr, nullptr);
  for (double& v : sum) v /= args.n;
  return chi2Test(solver, args, randomDrawVisCS(solver->vcs(), solver->vcsErr(), args.n, seed + 1), sum);
}

// ratioTestModel (src/RatioTest.cpp:13-62): mean and mean error of (b - b0)/b0 per point from TH1F(1024,-2,2).
inline RatioTestResult ratioTestModel(ISRSolverSLE* solver, int n, const std::vector<double>& toys,
                                      const std::vector<double>& bcs0) {
  const std::size_t N = solver->getN();
  if (bcs0.size() != N) throw DifferentModelCrossSectionSizes();
  RatioTestResult r;
  std::vector<double> stats(3 * N);
  r.hist.assign(N * 1026, 0u);
  solver->toyBatch(n, toys.data(), bcs0.data(), nullptr, nullptr, stats.data(), r.hist.data(), nullptr);
  r.means.resize(N); r.meanErrors.resize(N);
  for (std::size_t j = 0; j < N; ++j) {
    const double c = stats[3 * j], sx = stats[3 * j + 1], sx2 = stats[3 * j + 2];
    const double mean = c > 0 ? sx / c : 0.0;
    const double rms = c > 0 ? std::sqrt(std::fabs(sx2 / c - mean * mean)) : 0.0;
    r.means[j] = mean;
    r.meanErrors[j] = c > 0 ? rms / std::sqrt(c) : 0.0;
  }
  return r;
}

}  // namespace isrb200
#endif
