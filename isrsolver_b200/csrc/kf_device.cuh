// Device-side Kuraev-Fadin radiator F(x, s) and the per-row constants it needs.  (__host__ __device__: the same
// source is compiled for the host by tests/cpp/test_kf_host.cu and checked against the reference formula's known values
// without a GPU.)
// Reference formula: kernelKuraevFadin, src/KuraevFadin.cpp:14-49 (fLog :10, fBeta :12).
#pragma once

#include "common.cuh"

namespace isr {

constexpr double kApi = kAlphaQED / kPi;

// Row constants: everything in F(x, s) that depends on s only.  The reference recomputes
// fBeta(s)/fLog(s) (a log each) nine times per integrand evaluation (src/KuraevFadin.cpp:23-46).
__host__ __device__ inline void row_const_init(RowConst& rc, double E, double threshold) {
  rc.E = E;
  rc.s = E * E;
  rc.L = log(rc.s / kElectronM / kElectronM);           // fLog, :10
  rc.beta = (2 * kAlphaQED) / kPi * (rc.L - 1);          // fBeta, :12
  rc.inv_beta = 1.0 / rc.beta;
  rc.C1 = 1 + kAlphaQED / kPi * (kPi * kPi / 3 - 0.5) + 0.75 * rc.beta -
          rc.beta * rc.beta / 24 * (rc.L / 3 + 2 * kPi * kPi - 9.25);   // bracket of part1, :24-26
  rc.x0 = 4 * kElectronM / E;                            // pair threshold, :35 and :64
  rc.xin = rc.x0 * (1.0 + 1.0 / 16.0);
  const double r = threshold / E;
  rc.xmax = 1.0 - r * r;
}

// F(x, s) given ln(x) and the pair-threshold power (x - x0)^beta (pass pow_y < 0 below threshold).
// log1p(-x) replaces log(1-x): identical to rounding for the x range of the path, better for tiny x.
__host__ __device__ __forceinline__ double kf_value(const RowConst& rc, double x, double lnx, double pow_y) {
  const double b = rc.beta;
  const double mx = 1.0 - x;
  const double lnm = log1p(-x);
  const double inv_x = 1.0 / x;
  const double p1 = b * exp((b - 1.0) * lnx) * rc.C1;                                  // :23-26
  const double p2 = -b * (1.0 - 0.5 * x);                                              // :28
  const double p3 = 0.125 * b * b *
                    (-4.0 * (2.0 - x) * lnx - (1.0 + 3.0 * mx * mx) * inv_x * lnm - 6.0 + x);  // :30-31
  double res = p1 + p2 + p3;
  if (pow_y >= 0.0) {                                                                  // :35-46
    const double ss = 2.0 * lnx + rc.L - 5.0 / 3.0;
    const double sp1 = pow_y * (1.0 / 6.0) * inv_x * ss * ss * (2.0 - 2.0 * x + x * x + b / 3.0 * ss);
    const double sp2 = 0.5 * rc.L * rc.L *
                       (2.0 / 3.0 * (1.0 - mx * mx * mx) / mx + (2.0 - x) * lnm + 0.5 * x);
    res += kApi * kApi * (sp1 + sp2);
  }
  return res;
}

// Plain (x, s) evaluation with the reference's own threshold test, for isr_kernel_kf.
__host__ __device__ inline double kf_plain(double x, double s) {
  RowConst rc;
  const double E = sqrt(s);
  row_const_init(rc, E, 0.0);
  const double Ehalf = 0.5 * sqrt(s);
  const double thr = 2 * kElectronM / Ehalf;            // :35
  const double lnx = log(x);
  double pw = -1.0;
  if (x > thr) pw = exp(rc.beta * log(x - thr));
  // the reference uses log(1 - x) (:17); keep it here so spot values agree to the last digits
  const double b = rc.beta;
  const double mx = 1.0 - x;
  const double lnm = log(mx);
  const double p1 = b * pow(x, b - 1.0) * rc.C1;
  const double p2 = -b * (1.0 - 0.5 * x);
  const double p3 = 0.125 * b * b * (-4.0 * (2.0 - x) * lnx - (1.0 + 3.0 * mx * mx) / x * lnm - 6.0 + x);
  double res = p1 + p2 + p3;
  if (pw >= 0.0) {
    const double ss = 2.0 * lnx + rc.L - 5.0 / 3.0;
    const double sp1 = pw / 6.0 / x * ss * ss * (2.0 - 2.0 * x + x * x + b / 3.0 * ss);
    const double sp2 = 0.5 * rc.L * rc.L * (2.0 / 3.0 * (1.0 - mx * mx * mx) / mx + (2.0 - x) * lnm + 0.5 * x);
    res += kApi * kApi * (sp1 + sp2);
  }
  return res;
}

}  // namespace isr
