// Pure host computations of the matrix assembly: natural cubic splines, validation / ordering of the interpolation
// ranges, and the basis coefficient tables the kernels contract with.  No device code in this header: assemble.cu uses it to
// fill an isr_problem, tests/cpp/test_matrix_host.cu to assemble whole matrices on the CPU from the product's own source.
#pragma once

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace isr {

// Natural cubic spline second-derivative coefficients through (x_m, delta_{m,node}), m = 0..size-1,
// by the symmetric-tridiagonal LDL^T recurrence (the algorithm GSL's gsl_interp_cspline uses, which
// the reference instantiates once per cross-section point: src/CSplineRangeInterpolator.cpp:24-40).
inline void natural_spline_c(const std::vector<double>& x, int node, std::vector<double>& c) {
  const int size = (int)x.size();
  c.assign(size, 0.0);
  const int m = size - 2;  // interior unknowns c[1..size-2]
  if (m < 1) return;
  std::vector<double> diag(m), off(m), rhs(m);
  auto y = [node](int q) { return q == node ? 1.0 : 0.0; };
  for (int q = 0; q < m; ++q) {
    const double h0 = x[q + 1] - x[q], h1 = x[q + 2] - x[q + 1];
    const double g0 = h0 != 0.0 ? 1.0 / h0 : 0.0, g1 = h1 != 0.0 ? 1.0 / h1 : 0.0;
    off[q] = h1;
    diag[q] = 2.0 * (h1 + h0);
    rhs[q] = 3.0 * ((y(q + 2) - y(q + 1)) * g1 - (y(q + 1) - y(q)) * g0);
  }
  if (m == 1) {
    c[1] = rhs[0] / diag[0];
    return;
  }
  std::vector<double> alpha(m), gamma(m), z(m);
  alpha[0] = diag[0];
  gamma[0] = off[0] / alpha[0];
  for (int q = 1; q < m - 1; ++q) {
    alpha[q] = diag[q] - off[q - 1] * gamma[q - 1];
    gamma[q] = off[q] / alpha[q];
  }
  alpha[m - 1] = diag[m - 1] - off[m - 2] * gamma[m - 2];
  z[0] = rhs[0];
  for (int q = 1; q < m; ++q) z[q] = rhs[q] - gamma[q - 1] * z[q - 1];
  for (int q = 0; q < m; ++q) z[q] /= alpha[q];
  c[m] = z[m - 1];
  for (int q = m - 2; q >= 0; --q) c[q + 1] = z[q] - gamma[q] * c[q + 2];
}

// Validate + sort ranges exactly as Interpolator's constructor does (src/Interpolator.cpp:44-59,
// checkRangeInterpSettings :90-147).
inline int normalise_ranges(int n, int nranges, const int* ranges, std::vector<int>& out) {
  out.clear();
  if (nranges <= 0) {  // defaultInterpRangeSettings, :356-360
    out = {0, 0, n - 1};
    return ISR_OK;
  }
  std::vector<int> idx(nranges);
  for (int r = 0; r < nranges; ++r) idx[r] = r;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ranges[3 * a + 1] < ranges[3 * b + 1]; });
  if (ranges[3 * idx.front() + 1] != 0 || ranges[3 * idx.back() + 2] + 1 != n) {
    set_error("[!] Wrong interpolation ranges: must start at segment 0 and end at segment n-1");
    return ISR_ERANGE;
  }
  int pimax = -1;
  for (int r : idx) {
    const int imin = ranges[3 * r + 1], imax = ranges[3 * r + 2];
    if (imin > imax || (pimax != -1 && pimax + 1 != imin)) {
      set_error("[!] Wrong interpolation ranges: segments missed or taken twice");
      return ISR_ERANGE;
    }
    pimax = imax;
  }
  for (int r : idx) {
    out.push_back(ranges[3 * r] ? 1 : 0);
    out.push_back(ranges[3 * r + 1]);
    out.push_back(ranges[3 * r + 2]);
  }
  return ISR_OK;
}

// Ct[(k*4 + q)*n + j]: coefficient of (delta/h_k)^q of basis j on segment k = (ext[k], ext[k+1]]; klo / khi: first / last
// segment with non-zero coefficients per column; seg_range: the range that owns each segment.
// Segment k lies in exactly one range r (lo_r <= k <= hi_r), which decides the polynomial piece of every basis function
// on it: linear range -- rising edge of node j = k (delta/h), falling edge of node j = k-1 (1 - delta/h)
// (src/LinearRangeInterpolator.cpp:25-39, 70-101 incl. the (min,max] guards :74-76, :92-94); cspline range -- the cardinal
// natural spline of every node j in [max(lo-1,0), hi] (hasCSIndex, src/BaseRangeInterpolator.cpp:35-38, 80-100) restricted
// to the segment (src/CSplineRangeInterpolator.cpp:88-110).
inline void compute_basis_tables(int n, const std::vector<double>& x, const std::vector<int>& ranges, std::vector<double>& h_Ct,
                                 std::vector<int>& h_klo, std::vector<int>& h_khi, std::vector<int>& seg_range) {
  h_Ct.assign((size_t)n * 4 * n, 0.0);
  h_klo.assign(n, n);
  h_khi.assign(n, -1);
  seg_range.assign(n, 0);
  auto touch = [&](int j, int k) {
    h_klo[j] = std::min(h_klo[j], k);
    h_khi[j] = std::max(h_khi[j], k);
  };
  auto C = [&](int j, int k, int q) -> double& { return h_Ct[((size_t)k * 4 + q) * n + j]; };
  const int R = (int)ranges.size() / 3;
  std::vector<double> c;
  for (int r = 0; r < R; ++r) {
    const int cubic = ranges[3 * r], lo = ranges[3 * r + 1], hi = ranges[3 * r + 2];
    for (int k = lo; k <= hi; ++k) seg_range[k] = r;
    if (!cubic) {
      for (int k = lo; k <= hi; ++k) {
        C(k, k, 1) += 1.0;  // rising edge of node k
        touch(k, k);
        if (k >= 1) {       // falling edge of node k-1
          C(k - 1, k, 0) += 1.0;
          C(k - 1, k, 1) -= 1.0;
          touch(k - 1, k);
        }
      }
    } else {
      const int jlo = lo > 0 ? lo - 1 : 0;
      for (int j = jlo; j <= hi; ++j) {
        natural_spline_c(x, j + 1, c);
        for (int k = lo; k <= hi; ++k) {
          const double h = x[k + 1] - x[k];
          const double ylo = (k == j + 1) ? 1.0 : 0.0, yhi = (k + 1 == j + 1) ? 1.0 : 0.0;
          const double bq = (yhi - ylo) / h - h * (c[k + 1] + 2.0 * c[k]) / 3.0;
          const double dq = (c[k + 1] - c[k]) / (3.0 * h);
          C(j, k, 0) += ylo;
          C(j, k, 1) += bq * h;
          C(j, k, 2) += c[k] * h * h;
          C(j, k, 3) += dq * h * h * h;
          touch(j, k);
        }
      }
    }
  }
}

}  // namespace isr
