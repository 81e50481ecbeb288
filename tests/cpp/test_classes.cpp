// C++ class mirror over the C ABI: builds a problem from a golden fixture dumped as raw doubles, solves, runs
// toys and prints results for the Python test to compare (tests/test_gpu_cpp.py).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "isr_b200.hpp"

using namespace isrb200;

static std::vector<double> read_bin(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
  const std::streamsize sz = f.tellg();
  f.seekg(0);
  std::vector<double> v(sz / sizeof(double));
  f.read(reinterpret_cast<char*>(v.data()), sz);
  return v;
}
static void write_bin(const std::string& path, const double* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), n * sizeof(double));
}

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: test_classes <dir>\n"; return 2; }
  const std::string d = argv[1];
  auto ecm = read_bin(d + "/ecm.bin"), vcs = read_bin(d + "/vcs.bin"), err = read_bin(d + "/vcs_err.bin");
  auto toys = read_bin(d + "/toys.bin"), bcs0 = read_bin(d + "/bcs0.bin");
  const size_t n = ecm.size();
  const int T = (int)(toys.size() / n);
  try {
    ISRSolverSLE s(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827);
    try { s.setRangeInterpSettings({{false, 1, (int)n - 1}}); std::cerr << "expected InterpRangeException\n"; return 1; }
    catch (const InterpRangeException&) {}
    s.setRangeInterpSettings({{false, 0, 0}, {true, 1, (int)n - 1}});
    s.solve();
    write_bin(d + "/out_bcs.bin", s.bcs().data(), n);
    write_bin(d + "/out_G.bin", s.getIntegralOperatorMatrix().data(), n * n);
    write_bin(d + "/out_cov.bin", s.getBornCSCovMatrix().data(), n * n);
    auto r = chi2Test(&s, {T, 1e4, ""}, toys, bcs0);
    write_bin(d + "/out_chi2.bin", r.chi2.data(), T);
    auto rr = ratioTestModel(&s, T, toys, bcs0);
    write_bin(d + "/out_means.bin", rr.means.data(), n);
    std::printf("cond %.15g hist_sum %g\n", s.evalConditionNumber(), [&] { double t = 0; for (float c : r.hist) t += c; return t; }());

    ISRSolverTikhonov t(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827);
    t.setLambda(1e-3);
    t.solve();
    write_bin(d + "/out_tik_bcs.bin", t.bcs().data(), n);
    std::printf("tik x %.15g y %.15g curv %.15g\n", t.evalEqNorm2(), t.evalSmoothnessConstraintNorm2(), t.evalLCurveCurvature());

    ISRSolverTSVD v(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827);
    v.setUpperTSVDIndex((int)n / 2);
    v.solve();
    write_bin(d + "/out_tsvd_bcs.bin", v.bcs().data(), n);

    // ---- additions: free functions, SLE2, scans -----------------------------------------------------
    std::printf("kernel %.17g\n", kernelKuraevFadin(0.01, 2.25));
    auto bw = [](double e) {
      if (e <= 0.827) return 0.0;
      const double m = 1.5, g = 0.25, ps = 1.0 - 0.827 * 0.827 / (e * e), d = e * e - m * m;
      return 4.0 * m * m * g * g / (d * d + m * m * g * g) * ps * std::sqrt(ps);
    };
    std::vector<double> lo(n, 0.0), hi(n), er;
    for (size_t i = 0; i < n; ++i) hi[i] = 1.0 - (0.827 / ecm[i]) * (0.827 / ecm[i]);
    auto vis = convolutionKuraevFadin(ecm, bw, lo, hi, Efficiency(), 1e-10, 1e-12, &er);
    write_bin(d + "/out_vis.bin", vis.data(), n);
    std::printf("conv_scalar %.17g\n", convolutionKuraevFadin(ecm[3], bw, 0.0, hi[3]));
    auto cond = s.evalConditionNumberSpreadScan({0.0, 0.002});
    std::printf("condscan %.15g %.15g\n", cond[0], cond[1]);
    auto tm = chi2TestModel(&s, {256, 1e4, ""}, vcs, bcs0);
    double mean = 0; for (double c : tm.chi2) mean += c / tm.chi2.size();
    std::printf("chi2model_mean %.6g\n", mean);
    auto td = chi2TestData(&s, {256, 1e4, ""});
    std::printf("chi2data_n %zu\n", td.chi2.size());
    std::vector<std::vector<double>> hists(n, std::vector<double>(10, 0.5));
    ISRSolverSLE2 s2(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827, hists);
    s2.evalEqMatrix();
    s.setRangeInterpSettings({});
    s.evalEqMatrix();
    double worst = 0;
    for (size_t q = 0; q < n * n; ++q)
      worst = std::max(worst, std::fabs(s2.getIntegralOperatorMatrix().a[q] - 0.5 * s.getIntegralOperatorMatrix().a[q]));
    std::printf("sle2_half %.3g\n", worst);
    t.setLambda(1e-3);
    auto scan = t.toyScan({1e-4, 1e-3}, 8, toys.data(), bcs0);
    std::printf("toyscan %zu %.15g\n", scan.size(), scan[8]);
    std::printf("lcurve_lambda %.6g\n", t.solveLCurve());
  } catch (const std::exception& e) {
    std::cerr << "exception: " << e.what() << "\n";
    return 1;
  }
  std::puts("cpp ok");
  return 0;
}
