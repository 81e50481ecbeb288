// PyISR -- compiled CPython module over the C++ class mirror (include/isr_b200.hpp), i.e. the reference's
// src/PyISR.cpp re-backed by libisr_b200.so: same module name, class names, constructor kwargs
// (include/PyISRSolver.hpp:169-181) and method tables (include/PyISRSolverSLE.hpp:183-215,
// include/PyISRSolverTikhonov.hpp:111-146, include/PyISRSolverTSVD.hpp:30-81), plus the module functions
// kernelKuraevFadin / convolutionKuraevFadin / convolutionKuraevFadinBlurred (src/PyISR.cpp:197-207,
// include/PyKuraevFadin.hpp).  Host side: C++17 + pybind11; all numerical work goes through the C ABI.
//
// Differences from the reference module: files are .npz instead of .root (read / written
// through numpy from here); `efficiency` is an isrsolver_b200.Efficiency table or, as in the reference, a Python
// callable eps(x, energy) -- the latter is sampled at the quadrature nodes on the host (a kernel cannot call the
// interpreter); convolutionKuraevFadin keeps the callable form;
// getters return NumPy views that keep the solver alive (the reference's views dangle when it dies).
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <memory>

#include "isr_b200.hpp"

namespace py = pybind11;
using namespace isrb200;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;

static Efficiency to_efficiency(const py::object& o) {
  Efficiency e;
  if (o.is_none()) return e;
  if (!py::hasattr(o, "kind") || !py::hasattr(o, "values"))
    throw py::type_error("efficiency must be an isrsolver_b200.Efficiency table or a callable eps(x, energy)");
  const std::string kind = o.attr("kind").cast<std::string>();
  e.kind = kind == "one" ? ISR_EPS_ONE : kind == "row_table" ? ISR_EPS_ROW_TABLE : kind == "table_2d" ? ISR_EPS_TABLE_2D : ISR_EPS_E_ONLY;
  e.nx = o.attr("nx").cast<int>(); e.ne = o.attr("ne").cast<int>();
  e.xlo = o.attr("xlo").cast<double>(); e.xhi = o.attr("xhi").cast<double>();
  e.elo = o.attr("elo").cast<double>(); e.ehi = o.attr("ehi").cast<double>();
  auto grab = [](const py::object& a, std::vector<double>& out) {
    if (a.is_none()) return;
    arr v = arr::ensure(a);
    out.assign(v.data(), v.data() + v.size());
  };
  grab(o.attr("xedges"), e.xedges); grab(o.attr("eedges"), e.eedges); grab(o.attr("values"), e.values);
  return e;
}

static std::vector<double> vec(const py::object& o, std::size_t n, const char* name) {
  arr v = arr::ensure(o);
  if (!v || (std::size_t)v.size() != n) throw py::value_error(std::string(name) + ": expected " + std::to_string(n) + " values");
  return std::vector<double>(v.data(), v.data() + v.size());
}

// non-owning NumPy view on solver memory that keeps the solver object alive
template <typename Owner>
static py::array view1(const std::vector<double>& v, Owner& self) {
  return py::array_t<double>({(py::ssize_t)v.size()}, {(py::ssize_t)sizeof(double)}, v.data(), py::cast(&self));
}
// logical [i, j] view over column-major storage (include/PyISRSolverSLE.hpp:88-97 transposes the Eigen buffer)
static py::array matrix_copy(const Matrix& m) {
  py::array_t<double> out({m.rows, m.cols});
  auto r = out.mutable_unchecked<2>();
  for (int i = 0; i < m.rows; ++i)
    for (int j = 0; j < m.cols; ++j) r(i, j) = m(i, j);
  return out;
}

template <typename S>
static std::unique_ptr<S> make_solver(int n, const py::object& energy, const py::object& vcs, const py::object& energy_err,
                                      const py::object& vcs_err, double threshold, const py::object& efficiency,
                                      bool enableEnergySpread) {
  auto e = vec(energy, n, "energy"), v = vec(vcs, n, "vcs"), ve = vec(vcs_err, n, "vcs_err");
  std::vector<double> ee;
  if (!energy_err.is_none()) ee = vec(energy_err, n, "energy_err");
  // A Python callable efficiency(x, energy), as the reference's constructor takes it (include/PyISRSolver.hpp:186-197):
  // it is wrapped into the std::function overload of the C++ mirror, which samples it at the quadrature nodes on the host
  // (the reference calls it from inside every GSL integrand evaluation).  An optional attribute `breaks` on the callable
  // lists the x values where it jumps.
  std::unique_ptr<S> s;
  if (!efficiency.is_none() && !py::hasattr(efficiency, "kind") && PyCallable_Check(efficiency.ptr())) {
    std::vector<double> breaks;
    if (py::hasattr(efficiency, "breaks")) breaks = efficiency.attr("breaks").cast<std::vector<double>>();
    py::object fn = efficiency;
    std::function<double(double, double)> f = [fn](double x, double en) {
      py::gil_scoped_acquire gil;
      return fn(x, en).template cast<double>();
    };
    s = std::make_unique<S>((std::size_t)n, e.data(), v.data(), ee.empty() ? nullptr : ee.data(), ve.data(), threshold, f, breaks);
  } else {
    s = std::make_unique<S>((std::size_t)n, e.data(), v.data(), ee.empty() ? nullptr : ee.data(), ve.data(), threshold,
                            to_efficiency(efficiency));
  }
  if (enableEnergySpread) s->enableEnergySpread(); else s->disableEnergySpread();
  return s;
}

static std::pair<std::vector<double>, std::vector<double>> load_model(const std::string& path, const std::string& vname,
                                                                       const std::string& bname, std::size_t n) {
  py::object f = py::module_::import("numpy").attr("load")(path);
  auto grab = [&](const std::string& k) {
    arr a = arr::ensure(f[py::str(k)]);
    if (a.ndim() == 2) a = arr::ensure(a[py::int_(1)]);     // graphs are stored as rows (x, y, ex, ey)
    return std::vector<double>(a.data(), a.data() + a.size());
  };
  auto v = grab(vname), b = grab(bname);
  if (v.size() != b.size() || v.size() != n) throw DifferentModelCrossSectionSizes();
  return {v, b};
}

template <typename S, typename C>
static void bind_common(C& c) {
  // the long GPU calls run with the GIL released (the reference never releases it; other Python threads -- one
  // per GPU, say -- can then run); a callable efficiency re-acquires it for the duration of each call-back
  c.def("solve", [](S& s) { py::gil_scoped_release nogil; s.solve(); return 0; }, "Find solution")
      .def("bcs", [](S& s) { return view1(s.bcs(), s); }, "Get numerical solution (Born cross section)")
      .def("ecm", [](S& s) { return view1(s.ecm(), s); }, "Get center-of-mass energy")
      .def("ecm_err", [](S& s) { return view1(s.ecmErr(), s); }, "Get center-of-mass energy errors")
      .def("vcs", [](S& s) { return view1(s.vcs(), s); }, "Get visible cross section")
      .def("vcs_err", [](S& s) { return view1(s.vcsErr(), s); }, "Get visible cross section errors")
      .def("bcs_cov_matrix", [](S& s) { return matrix_copy(s.getBornCSCovMatrix()); },
           "Get covariance matrix of the numerical solution (Born cross section)")
      .def("bcs_inv_cov_matrix", [](S& s) { return matrix_copy(s.evalBornCSInvCovMatrix()); },
           "Evaluate inverse covariance matrix of the numerical solution (Born cross section)")
      .def("intop_matrix", [](S& s) { return matrix_copy(s.getIntegralOperatorMatrix()); }, "Integral operator matrix")
      .def("condnum_eval", [](S& s) { return s.evalConditionNumber(); },
           "Eval condition number of the integral operator matrix F or GF in the case, when c.m. energy spread is enabled.")
      .def("interp_eval", [](S& s, double cm_energy) { return s.interpEval(s.bcs(), cm_energy); }, py::arg("cm_energy"),
           "Eval interpolation value at certain c.m. energy point")
      .def("reset_ecm_err", [](S& s, const py::object& e) { s.resetECMErrors(vec(e, s.getN(), "energy_err")); return 0; },
           py::arg("energy_err"), "Reset c.m. energy spread")
      .def("set_interp_settings",
           [](S& s, const std::vector<std::tuple<bool, int, int>>& st) { s.setRangeInterpSettings(st); return 0; },
           "Set interpolation settings")
      .def("save",
           [](S& s, const std::string& output_path, const std::string& vcs_name, const std::string& bcs_name) {
             py::module_ np = py::module_::import("numpy");
             py::dict kw;
             const std::size_t n = s.getN();
             std::vector<double> zero(n, 0.0), berr(n, 0.0);
             const Matrix& cov = s.getBornCSCovMatrix();
             if (cov.rows == (int)n) for (std::size_t i = 0; i < n; ++i) berr[i] = std::sqrt(cov((int)i, (int)i));
             kw[py::str(vcs_name)] = np.attr("stack")(py::make_tuple(view1(s.ecm(), s), view1(s.vcs(), s), view1(s.ecmErr(), s), view1(s.vcsErr(), s)));
             kw[py::str(bcs_name)] = np.attr("stack")(py::make_tuple(view1(s.ecm(), s), view1(s.bcs(), s), py::cast(zero), py::cast(berr)));
             kw["intergalOperatorMatrix"] = matrix_copy(s.getIntegralOperatorMatrix());   // sic: the reference's object name
             if (cov.rows == (int)n) {
               kw["covMatrixBornCS"] = matrix_copy(cov);
               kw["invCovMatrixBornCS"] = matrix_copy(s.evalBornCSInvCovMatrix());
             }
             np.attr("savez")(output_path, **kw);
             return 0;
           },
           py::arg("output_path"), py::arg("vcs_name") = "vcs", py::arg("bcs_name") = "bcs", "Save results")
      .def_property_readonly("n", [](S& s) { return s.getN(); }, "Number of points")
      .def_property_readonly("energy_spread_enabled", [](S& s) { return s.isEnergySpreadEnabled(); });
}

template <typename S, typename C>
static void bind_tests(C& c) {
  c.def("eval_equation_matrix", [](S& s) { py::gil_scoped_release nogil; s.evalEqMatrix(); return 0; }, "Eval equation matrix")
      .def("chi2_test_model",
           [](S& s, int n, double initial_ampl, const std::string& model_path, const std::string& model_visible_cs_name,
              const std::string& model_born_cs_name, const py::object& output_path) {
             auto m = load_model(model_path, model_visible_cs_name, model_born_cs_name, s.getN());
             Chi2TestResult r;
             {
               py::gil_scoped_release nogil;
               r = chi2TestModel(&s, {n, initial_ampl, ""}, m.first, m.second);
             }
             py::dict out;
             out["chi2"] = py::array_t<double>(r.chi2.size(), r.chi2.data());
             out["hist"] = py::array_t<float>(r.hist.size(), r.hist.data());
             out["lo"] = r.chi2Min; out["hi"] = r.chi2Max; out["initial_ampl"] = initial_ampl;
             if (!output_path.is_none())
               py::module_::import("numpy").attr("savez")(output_path, py::arg("chi2Hist") = out["hist"], py::arg("chi2Min") = r.chi2Min,
                                                           py::arg("chi2Max") = r.chi2Max, py::arg("chi2") = out["chi2"]);
             return out;
           },
           py::arg("n"), py::arg("initial_ampl"), py::arg("model_path"), py::arg("model_visible_cs_name"),
           py::arg("model_born_cs_name"), py::arg("output_path") = py::none(),
           "Chi2 test of the numerical solution with respect to the model Born cross section.")
      .def("ratio_test_model",
           [](S& s, int n, const std::string& model_path, const std::string& model_visible_cs_name,
              const std::string& model_born_cs_name, const py::object& output_path) {
             auto m = load_model(model_path, model_visible_cs_name, model_born_cs_name, s.getN());
             RatioTestResult r;
             {
               py::gil_scoped_release nogil;
               s.solve();
               r = ratioTestModel(&s, n, randomDrawVisCS(m.first, s.vcsErr(), n), m.second);
             }
             py::dict out;
             out["means"] = py::array_t<double>(r.means.size(), r.means.data());
             out["mean_errors"] = py::array_t<double>(r.meanErrors.size(), r.meanErrors.data());
             out["hist"] = py::array_t<uint32_t>({(py::ssize_t)s.getN(), (py::ssize_t)1026}, r.hist.data());
             if (!output_path.is_none())
               py::module_::import("numpy").attr("savez")(output_path, py::arg("means") = out["means"], py::arg("mean_errors") = out["mean_errors"]);
             return out;
           },
           py::arg("n"), py::arg("model_path"), py::arg("model_visible_cs_name"), py::arg("model_born_cs_name"),
           py::arg("output_path") = py::none(), "Ratio test (test of interpolation quality).");
}

static const char* kCtorDoc = "n, energy, vcs, energy_err, vcs_err, threshold, efficiency=None, enableEnergySpread=False";
#define CTOR_ARGS py::arg("n"), py::arg("energy"), py::arg("vcs"), py::arg("energy_err"), py::arg("vcs_err"), py::arg("threshold"), \
                  py::arg("efficiency") = py::none(), py::arg("enableEnergySpread") = false

PYBIND11_MODULE(PyISR, m) {
  m.doc() = "usage: PyISR.convolutionKuraevFadin(energy, fcn, min_x, max_x)\n(B200 back-end: libisr_b200.so)";
  py::register_exception<InterpRangeException>(m, "InterpRangeException", PyExc_ValueError);
  py::register_exception<EfficiencyDimensionException>(m, "EfficiencyDimensionException", PyExc_ValueError);
  py::register_exception<DifferentModelCrossSectionSizes>(m, "DifferentModelCrossSectionSizes", PyExc_ValueError);
  py::register_exception<IsrGpuError>(m, "IsrGpuError", PyExc_RuntimeError);

  m.def("kernelKuraevFadin", [](double x, double s) { return kernelKuraevFadin(x, s); }, py::arg("x"), py::arg("s"));
  m.def("convolutionKuraevFadin",
        [](double energy, const std::function<double(double)>& fcn, double min_x, double max_x, const py::object& efficiency) {
          if (efficiency.is_none() || py::hasattr(efficiency, "kind"))
            return convolutionKuraevFadin(energy, fcn, min_x, max_x, to_efficiency(efficiency));
          // the reference's callable form eps(x, energy), sampled at the plan's nodes
          auto eff = efficiency.cast<std::function<double(double, double)>>();
          isr_conv_plan* plan = nullptr;
          check(isr_conv_plan_create(Context::get(), 1, &energy, &min_x, &max_x, nullptr, 0, nullptr, &plan));
          struct Guard { isr_conv_plan* p; ~Guard() { isr_conv_plan_destroy(p); } } guard{plan};
          std::vector<double> e, x, f;
          for (int level = 0; level < 41; ++level) {
            int64_t npan = 0, nn = 0, nsplit = 0;
            check(isr_conv_plan_size(plan, &npan, &nn));
            if (nn) {
              e.resize(nn); x.resize(nn); f.resize(nn);
              check(isr_conv_plan_nodes(plan, e.data(), x.data(), nullptr, nullptr));
              for (int64_t q = 0; q < nn; ++q) f[q] = fcn(e[q]) * eff(x[q], energy);
              check(isr_conv_plan_feed(plan, f.data()));
            }
            if (level == 40) break;
            check(isr_conv_plan_refine(plan, 1e-10, 1e-12, &nsplit));
            if (!nsplit) break;
          }
          double res = 0;
          check(isr_conv_plan_result(plan, &res, nullptr));
          return res;
        },
        py::arg("energy"), py::arg("fcn"), py::arg("min_x"), py::arg("max_x"), py::arg("efficiency") = py::none(),
        "convolutionKuraevFadin(energy, fcn, min_x, max_x, efficiency = lambda x, energy: 1.)");
  m.def("convolutionKuraevFadinBlurred",
        [](double energy, const std::function<double(double)>& fcn, double threshold_energy, double cm_energy_spread,
           const py::object& efficiency) {
          // gaussian_conv (src/Integration.cpp:86-100) of the visible cross section: 6-point Gauss-Hermite
          static const double t[6] = {-2.350604973674492222833922, -1.335849074013696949714895, -4.360774119276165086792159e-1,
                                      4.360774119276165086792159e-1, 1.335849074013696949714895, 2.350604973674492222833922};
          static const double w[6] = {4.530009905508845640857473e-3, 1.570673203228566439163116e-1, 7.246295952243925240919147e-1,
                                      7.246295952243925240919147e-1, 1.570673203228566439163116e-1, 4.530009905508845640857473e-3};
          std::vector<double> en, lo, hi;
          std::vector<int> which;
          for (int q = 0; q < 6; ++q) {
            const double e = energy + std::sqrt(2.0) * cm_energy_spread * t[q];
            if (e > threshold_energy) { en.push_back(e); lo.push_back(0.0); hi.push_back(1.0 - (threshold_energy / e) * (threshold_energy / e)); which.push_back(q); }
          }
          double res = 0;
          if (!en.empty()) {
            auto vis = convolutionKuraevFadin(en, fcn, lo, hi, to_efficiency(efficiency));
            for (std::size_t k = 0; k < en.size(); ++k) res += w[which[k]] / std::sqrt(3.14159265358979323846) * vis[k];
          }
          return res;
        },
        py::arg("energy"), py::arg("fcn"), py::arg("threshold_energy"), py::arg("cm_energy_spread"), py::arg("efficiency") = py::none());

  py::class_<ISRSolverSLE> sle(m, "ISRSolverSLE", kCtorDoc);
  sle.def(py::init(&make_solver<ISRSolverSLE>), CTOR_ARGS);
  bind_common<ISRSolverSLE>(sle);
  bind_tests<ISRSolverSLE>(sle);

  py::class_<ISRSolverTikhonov> tik(m, "ISRSolverTikhonov", kCtorDoc);
  tik.def(py::init(&make_solver<ISRSolverTikhonov>), CTOR_ARGS);
  bind_common<ISRSolverTikhonov>(tik);
  bind_tests<ISRSolverTikhonov>(tik);
  tik.def_property("reg_param", [](ISRSolverTikhonov& s) { return s.getLambda(); }, [](ISRSolverTikhonov& s, double l) { s.setLambda(l); },
                   "Regularization parameter")
      .def("make_LCurve",
           [](ISRSolverTikhonov& s, const arr& lambdas) {
             if (lambdas.ndim() != 1)
               throw py::type_error("Wrong dimension of the array with regularization parameter values. The array must be 1D array.");
             std::vector<double> lam(lambdas.data(), lambdas.data() + lambdas.size()), x, y, c, dc;
             s.scanLCurve(lam, x, y, c, dc);
             for (double& v : c) v = -v;                      // include/PyISRSolverTikhonov.hpp:66: c = -curvature
             if (!lam.empty()) { s.setLambda(lam.back()); s.solve(); }   // the reference leaves the solver at the last lambda
             py::dict out;
             out["x"] = py::array_t<double>(x.size(), x.data());
             out["y"] = py::array_t<double>(y.size(), y.data());
             out["c"] = py::array_t<double>(c.size(), c.data());
             return out;
           },
           py::arg("lambdas"), "L-curve: residual norm x, smoothness norm y and curvature c for every lambda")
      .def("solve_LCurve",
           [](ISRSolverTikhonov& s, double initial_lambda) {
             const double lam = s.solveLCurve(1e-12, std::max(10.0 * initial_lambda, 1.0));
             py::dict out;
             out["lambda"] = lam;
             out["max_curv"] = -s.evalLCurveCurvature();
             return out;
           },
           py::arg("initial_lambda"), "Find the regularization parameter of maximum L-curve curvature and solve");

  py::class_<IterISRInterpSolver> iter(m, "IterISRInterpSolver", kCtorDoc);   // include/PyIterISRInterpSolver.hpp
  iter.def(py::init(&make_solver<IterISRInterpSolver>), CTOR_ARGS);
  bind_common<IterISRInterpSolver>(iter);
  iter.def("eval_equation_matrix", [](IterISRInterpSolver& s) { py::gil_scoped_release nogil; s.evalEqMatrix(); return 0; }, "Eval equation matrix")
      .def_property("n_iter", [](IterISRInterpSolver& s) { return s.getNumOfIters(); },
                    [](IterISRInterpSolver& s, std::size_t n) { s.setNumOfIters(n); }, "Number of iterations")
      .def("radcorr", [](IterISRInterpSolver& s) { return view1(s.getRadCorr(), s); }, "Radiative correction 1 + delta");

  py::class_<ISRSolverTSVD> tsvd(m, "ISRSolverTSVD", kCtorDoc);
  tsvd.def(py::init([](int n, const py::object& energy, const py::object& vcs, const py::object& energy_err, const py::object& vcs_err,
                       double threshold, const py::object& efficiency, bool enableEnergySpread, const py::object& upper_TSVD_index) {
             auto s = make_solver<ISRSolverTSVD>(n, energy, vcs, energy_err, vcs_err, threshold, efficiency, enableEnergySpread);
             if (!upper_TSVD_index.is_none()) s->setUpperTSVDIndex(upper_TSVD_index.cast<int>());
             return s;
           }),
           CTOR_ARGS, py::arg("upper_TSVD_index") = py::none());
  bind_common<ISRSolverTSVD>(tsvd);
  tsvd.def_property("upper_TSVD_index", [](ISRSolverTSVD& s) { return s.getUpperTSVDIndex(); },
                    [](ISRSolverTSVD& s, int k) { s.setUpperTSVDIndex(k); })
      .def("enable_keepone", [](ISRSolverTSVD& s) { s.enableKeepOne(); return 0; }, "Keep just higher TSVD harmonic")
      .def("disable_keepone", [](ISRSolverTSVD& s) { s.disableKeepOne(); return 0; }, "Keep all TSVD harmonics");
}
