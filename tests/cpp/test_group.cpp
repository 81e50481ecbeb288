// chi2TestModel of the C++ class mirror on one GPU and spread over `ngpu` GPUs of the box from this ONE process
// (isr_group: host thread + context + replicated problem per GPU, NCCL all-reduces in stream) -- the C++ caller the
// reference's drivers are (src/isrsolver-SLE-chi2-test.cpp:106-157, include/Chi2Test.hpp:18-42).  The histogram, the range
// and every chi2 value must be bit-identical.  Also: the std::function efficiency overload of the constructor
// (include/BaseISRSolver.hpp:33-40).
//   test_group <dir with ecm.bin vcs.bin vcs_err.bin bcs0.bin> <ngpu>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

int main(int argc, char** argv) {
  if (argc < 3) { std::cerr << "usage: test_group <dir> <ngpu>\n"; return 2; }
  const std::string d = argv[1];
  const int ngpu = std::atoi(argv[2]);
  auto ecm = read_bin(d + "/ecm.bin"), vcs = read_bin(d + "/vcs.bin"), err = read_bin(d + "/vcs_err.bin"), bcs0 = read_bin(d + "/bcs0.bin");
  const size_t n = ecm.size();
  try {
    ISRSolverSLE s(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827);
    s.setRangeInterpSettings({{false, 0, 0}, {true, 1, (int)n - 1}});
    const Chi2TestArgs args{30011, 1e4, ""};
    auto one = chi2TestModel(&s, args, vcs, bcs0, 7);
    auto many = chi2TestModel(&s, args, vcs, bcs0, 7, ngpu);
    if (one.chi2.size() != many.chi2.size() || std::memcmp(one.chi2.data(), many.chi2.data(), sizeof(double) * one.chi2.size()) != 0) {
      std::cerr << "chi2 values differ between 1 and " << ngpu << " GPUs\n"; return 1;
    }
    if (one.hist != many.hist || one.chi2Min != many.chi2Min || one.chi2Max != many.chi2Max) { std::cerr << "histograms differ\n"; return 1; }
    double tot = 0; for (float c : many.hist) tot += c;
    std::printf("group_hist_sum %.0f gpus %d min %.17g max %.17g\n", tot, ngpu, many.chi2Min, many.chi2Max);
    // Tikhonov / TSVD selections are not replicable: must be refused, not silently run as SLE
    ISRSolverTikhonov t(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827);
    try { chi2TestModel(&t, args, vcs, bcs0, 7, ngpu > 1 ? ngpu : 2); std::cerr << "expected a refusal\n"; return 1; }
    catch (const std::invalid_argument&) {}
    // callable efficiency (std::function overload): eps = 0.5 everywhere == half the eps = 1 matrix
    ISRSolverSLE h(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827, [](double, double) { return 0.5; });
    h.setRangeInterpSettings({{false, 0, 0}, {true, 1, (int)n - 1}});
    h.evalEqMatrix();
    s.evalEqMatrix();
    double worst = 0;
    for (size_t q = 0; q < n * n; ++q)
      worst = std::max(worst, std::fabs(h.getIntegralOperatorMatrix().a[q] - 0.5 * s.getIntegralOperatorMatrix().a[q]));
    std::printf("callable_half %.3g\n", worst);
    // ... and one that depends on x and E, written out for the Python side to compare with the ctypes mirror
    ISRSolverSLE c(n, ecm.data(), vcs.data(), nullptr, err.data(), 0.827,
                   [](double x, double e) { return 0.3 + 0.5 * std::exp(-30.0 * x) * (1.0 - 0.1 * (e - 1.5)); });
    c.setRangeInterpSettings({{false, 0, 0}, {true, 1, (int)n - 1}});
    c.evalEqMatrix();
    std::ofstream f(d + "/out_G_callable.bin", std::ios::binary);
    f.write(reinterpret_cast<const char*>(c.getIntegralOperatorMatrix().data()), n * n * sizeof(double));
  } catch (const std::exception& e) {
    std::cerr << "exception: " << e.what() << "\n";
    return 1;
  }
  std::puts("group ok");
  return 0;
}
