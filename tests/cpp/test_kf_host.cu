// CPU check of the product's own integrand source (isrsolver_b200/csrc/kf_device.cuh, panel_device.cuh), compiled for the
// host by nvcc (the functions are __host__ __device__), against the oracle's restatement of the reference
// (oracle/isr_oracle.c: orc_kernel_kf = kernelKuraevFadin src/KuraevFadin.cpp:14-49, orc_find_bin = ROOT TAxis::FindBin)
// and the known values of SURVEY.md 8c.  No GPU is needed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "panel_device.cuh"

extern "C" double orc_kernel_kf(double x, double s);
extern "C" int orc_find_bin(int nbins, double lo, double hi, const double* edges, double x);

using namespace isr;

#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      std::printf("FAILED %s (line %d): ", #cond, __LINE__); \
      std::printf(__VA_ARGS__);                            \
      std::printf("\n");                                   \
      std::exit(1);                                        \
    }                                                      \
  } while (0)

int main() {
  // ---- known answers (SURVEY.md 8c iii)
  const double known[3][3] = {{1e-6, 28082.309274929197}, {0.01, 5.283686571158812}, {0.5, 0.08908917563287333}};
  for (auto& k : known) CHECK(std::fabs(kf_plain(k[0], 2.25) / k[1] - 1.0) < 1e-13, "F(%g, 2.25) = %.17g", k[0], kf_plain(k[0], 2.25));

  // ---- kf_plain and kf_value (the form the quadrature kernels use) against the oracle on a grid
  long n = 0;
  const double energies[] = {0.9, 1.18, 1.5, 2.0, 3.1, 10.58, 91.0};
  for (double E : energies) {
    const double s = E * E;
    RowConst rc;
    row_const_init(rc, E, 0.827);
    for (int q = 0; q <= 400; ++q) {
      const double x = std::pow(10.0, -8.0 + 8.0 * q / 400.0) * 0.99;
      const double ref = orc_kernel_kf(x, s);
      const double a = kf_plain(x, s);
      CHECK(std::fabs(a - ref) <= 2e-14 * std::fabs(ref) + 1e-300, "kf_plain(%g, %g) = %.17g, oracle %.17g", x, s, a, ref);
      const double pw = x > rc.x0 ? std::exp(rc.beta * std::log(x - rc.x0)) : -1.0;
      const double b = kf_value(rc, x, std::log(x), pw);
      CHECK(std::fabs(b - ref) <= 1e-12 * std::fabs(ref), "kf_value(%g, %g) = %.17g, oracle %.17g", x, s, b, ref);
      ++n;
    }
    // grading points: ascending, inside (0, xmax), first one the pair threshold, neighbours within a factor 4 of the
    // distance to the singular point they approach
    if (rc.xmax > rc.x0) {
      row_grading(rc);
      CHECK(rc.nrb >= 1 && rc.nrb <= kMaxRowBreaks && rc.rb[0] == rc.x0, "grading of E = %g", E);
      for (int m = 1; m < rc.nrb; ++m) CHECK(rc.rb[m] > rc.rb[m - 1] && rc.rb[m] < rc.xmax, "grading order at E = %g", E);
    }
  }

  // ---- ROOT bin arithmetic, fixed and variable bins, on random points and exactly on the edges
  std::vector<double> edges;
  for (int nx : {1, 7, 64, 512, 1000}) {
    const double lo = nx == 7 ? -0.25 : 0.0, hi = nx == 7 ? 1.75 : 1.0;
    EpsDev fixed{2, nx, lo, hi, nullptr, nullptr};
    edges.resize(nx + 1);
    for (int m = 0; m <= nx; ++m) edges[m] = lo + (hi - lo) * std::pow((double)m / nx, 1.7);
    EpsDev var{2, nx, lo, hi, edges.data(), nullptr};
    unsigned long long st = 88172645463325252ull;
    for (int q = 0; q < 20000; ++q) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      double x = lo - 0.1 + (hi - lo + 0.2) * (double)(st >> 11) / 9007199254740992.0;
      if (q < 2 * (nx + 1)) x = (q & 1) ? edges[q / 2] : lo + (hi - lo) * (double)(q / 2) / nx;   // on the edges
      CHECK(eps_find_bin(fixed, x) == orc_find_bin(nx, lo, hi, nullptr, x), "fixed bins nx = %d x = %.17g", nx, x);
      CHECK(eps_find_bin(var, x) == orc_find_bin(nx, lo, hi, edges.data(), x), "variable bins nx = %d x = %.17g", nx, x);
      ++n;
    }
  }
  // ---- TH1 bin arithmetic of the chi2 / ratio histograms (th1_bin; fixed bins) against the same oracle routine
  {
    unsigned long long st = 1234567ull;
    for (int nb : {128, 1024}) {
      const double lo = nb == 1024 ? -2.0 : 151.73, hi = nb == 1024 ? 2.0 : 268.4;
      for (int q = 0; q < 50000; ++q) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        double x = lo - 0.05 * (hi - lo) + 1.1 * (hi - lo) * (double)(st >> 11) / 9007199254740992.0;
        if (q <= nb) x = lo + (hi - lo) * (double)q / nb;     // on the edges
        CHECK(th1_bin(nb, lo, hi, x) == orc_find_bin(nb, lo, hi, nullptr, x), "th1_bin nbins = %d x = %.17g", nb, x);
        ++n;
      }
    }
  }
  std::printf("integrand source ok on the host: %ld comparisons\n", n);
  return 0;
}
