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
  } catch (const std::exception& e) {
    std::cerr << "exception: " << e.what() << "\n";
    return 1;
  }
  std::puts("cpp ok");
  return 0;
}
