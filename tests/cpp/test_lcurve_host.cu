// The per-lambda arithmetic of the Tikhonov path (isrsolver_b200/csrc/tikhonov_device.cuh: lcurve_point, the code
// lcurve_kernel runs) on the CPU.  stdin: n L / mu[n] / c[n] / lambda[L]; stdout per lambda: x y curv dcurv then the n
// coefficients c / (mu + lambda).  tests/test_host.py builds (mu, c) from a golden problem with SciPy's generalised
// eigen-solver and compares with the literal matrix formulas of src/ISRSolverTikhonov.cpp:61-128 stored in the fixture.
#include <cstdio>
#include <vector>

#include "tikhonov_device.cuh"

int main() {
  int n, L;
  if (std::scanf("%d %d", &n, &L) != 2) return 2;
  std::vector<double> mu(n), c(n), lam(L), z(n);
  for (double& v : mu) if (std::scanf("%lf", &v) != 1) return 2;
  for (double& v : c) if (std::scanf("%lf", &v) != 1) return 2;
  for (double& v : lam) if (std::scanf("%lf", &v) != 1) return 2;
  for (int l = 0; l < L; ++l) {
    double x, y, cv, dcv;
    isr::lcurve_point(n, lam[l], mu.data(), c.data(), x, y, cv, dcv, z.data());
    std::printf("%.17g %.17g %.17g %.17g", x, y, cv, dcv);
    for (int m = 0; m < n; ++m) std::printf(" %.17g", z[m]);
    std::printf("\n");
  }
  return 0;
}
