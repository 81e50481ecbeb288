// Per-lambda arithmetic of the Tikhonov path in the generalised eigen-basis (G^T W G) u = mu F u, U^T F U = I, c = U^T G^T W v
// (DESIGN K7).  __host__ __device__: tikhonov.cu's kernels call it, tests/cpp/test_lcurve_host.cu runs it on
// the CPU against the oracle's literal matrix formulas of src/ISRSolverTikhonov.cpp:61-128.
#pragma once

#include <cmath>

namespace isr {

// L-curve functionals at one lambda:
//   x     = |W^(1/2) (G b - v)|^2 = sum c^2 lambda^2 / (mu (mu + lambda)^2)      evalEqNorm2, :106-109 (cancellation-free form)
//   y     = b^T F b               = sum c^2 / (mu + lambda)^2                     evalSmoothnessConstraintNorm2, :111-113
//   curv  = -|1 / xi' / (1 + lambda^2)^(3/2)|,  xi' = -2 sum c^2 / (mu + lambda)^3                 evalLCurveCurvature, :115-119
//   dcurv = -xi'' xi'^-2 (1 + lambda^2)^(-3/2) - 3 lambda / xi' (1 + lambda^2)^(-5/2),  xi'' = 6 sum c^2 / (mu + lambda)^4    :121-128
// z (may be nullptr): coefficients c / (mu + lambda) of b_lambda in the U basis.
__host__ __device__ inline void lcurve_point(int n, double la, const double* __restrict__ mu, const double* __restrict__ c,
                                             double& x, double& y, double& curv, double& dcurv, double* __restrict__ z) {
  double sx = 0, sy = 0, s3 = 0, s4 = 0;
  for (int m = 0; m < n; ++m) {
    const double cm = c[m], d = mu[m] + la, id = 1.0 / d, c2 = cm * cm;
    const double q2 = c2 * id * id;
    sy += q2;
    sx += q2 * (la * la / mu[m]);
    s3 += q2 * id;
    s4 += q2 * id * id;
    if (z) z[m] = cm * id;
  }
  const double dksi = -2.0 * s3, d2ksi = 6.0 * s4;
  const double q = 1.0 + la * la;
  x = sx;
  y = sy;
  curv = -fabs(1.0 / dksi / pow(q, 1.5));
  dcurv = -d2ksi * (1.0 / (dksi * dksi)) * pow(q, -1.5) + -3.0 * la / dksi * pow(q, -2.5);
}

}  // namespace isr
