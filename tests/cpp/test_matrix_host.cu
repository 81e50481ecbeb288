// Whole integral-operator matrices on the CPU from the product's own source: range normalisation and basis tables
// (isrsolver_b200/csrc/basis_host.hpp, the host code isr_problem_set_ranges runs) + panel enumeration, substitutions,
// integrand (panel_device.cuh, the code the CUDA kernels run, compiled for the host) + the contraction
// G(i,j) = sum_k sum_p M[i][k][p] Ct[k][p][j] of assemble.cu's contract_kernel.  tests/test_host.py feeds it the golden
// fixtures' problems and applies the same parity gates as the GPU test (tests/test_gpu_assembly.py) -- without a GPU.
//   stdin : n  threshold  nranges  /  ecm[n]  /  nranges x (cubic lo hi)  /  nx xlo xhi  [/ n x (nx + 2) table values]
//           /  spread (0 | 1)  [/ sigma_E[n]]
//   stdout: n x n matrix, row-major, %.17g (with spread: G_spread * G, src/ISRSolverSLE.cpp:146-148)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "basis_host.hpp"
#include "panel_device.cuh"

namespace isr {   // the library's error sink (capi.cu) is not linked into this harness
void set_error(const char* fmt, ...) { std::fprintf(stderr, "%s\n", fmt); }
}
using namespace isr;

static void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n); w.resize(n);
  for (int i = 0; i < n; ++i) {
    long double z = std::cos(3.14159265358979323846L * (i + 0.75L) / (n + 0.5L)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1, p2 = 0;
      for (int j = 1; j <= n; ++j) { const long double p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1);
      const long double dz = p1 / pp;
      z -= dz;
      if (std::fabs((double)dz) < 1e-19) break;
    }
    x[n - 1 - i] = (double)z;
    w[n - 1 - i] = (double)(2 / ((1 - z * z) * pp * pp));
  }
}

int main() {
  int n, nranges;
  double thr;
  if (std::scanf("%d %lf %d", &n, &thr, &nranges) != 3) return 2;
  std::vector<double> ext(n + 1), invh(n);
  ext[0] = thr;
  for (int i = 0; i < n; ++i) if (std::scanf("%lf", &ext[i + 1]) != 1) return 2;
  std::vector<int> raw(3 * nranges), ranges;
  for (int& v : raw) if (std::scanf("%d", &v) != 1) return 2;
  if (normalise_ranges(n, nranges, raw.data(), ranges) != ISR_OK) return 3;     // ISR_ERANGE -> exit code 3
  std::vector<double> Ct;
  std::vector<int> klo, khi, seg_range;
  compute_basis_tables(n, ext, ranges, Ct, klo, khi, seg_range);
  for (int k = 0; k < n; ++k) invh[k] = 1.0 / (ext[k + 1] - ext[k]);
  std::vector<double> glx, glw;
  gauss_legendre(16, glx, glw);
  // optional per-row efficiency table (ROOT layout [n][nx + 2], fixed bins) and optional beam-energy spread
  int nx = 0, spread = 0;
  double xlo = 0.0, xhi = 1.0;
  if (std::scanf("%d %lf %lf", &nx, &xlo, &xhi) != 3) return 2;
  std::vector<double> vals((size_t)n * (nx + 2));
  if (nx > 0) for (double& v : vals) if (std::scanf("%lf", &v) != 1) return 2;
  if (std::scanf("%d", &spread) != 1) return 2;
  std::vector<double> sigE(n, 0.0);
  if (spread) for (double& v : sigE) if (std::scanf("%lf", &v) != 1) return 2;
  EpsDev eps{nx > 0 ? 2 : 0, nx, xlo, xhi, nullptr, nx > 0 ? vals.data() : nullptr};
  std::vector<double> M((size_t)n * 4), G((size_t)n * n);
  for (int i = 0; i < n; ++i) {
    RowConst rc;
    row_const_init(rc, ext[i + 1], thr);
    row_grading(rc);
    std::fill(M.begin(), M.end(), 0.0);
    for (int k = 0; k <= i; ++k) {
      const ItemCuts c = item_cuts(rc, eps, ext.data(), i, k);
      double* m = &M[(size_t)k * 4];
      walk_panels(c, rc, eps, 0, c.count(), [&](int, double a, double b) {
        const int kind = panel_kind(rc, b);
        const double ev = nx > 0 ? vals[(size_t)i * (nx + 2) + eps_find_bin(eps, 0.5 * (a + b))] : 1.0;   // as eval_kernel does
        for (int q = 0; q < 16; ++q) {
          const PanelNode pn = panel_node(rc, a, b, kind, glx[q], glw[q]);
          const double f = kf_value(rc, pn.x, pn.lnx, pn.pw) * pn.wj * ev;
          const double dn = (rc.E * std::sqrt(1.0 - pn.x) - ext[k]) * invh[k];
          m[0] += f; m[1] += f * dn; m[2] += f * dn * dn; m[3] += f * dn * dn * dn;
        }
      });
    }
    for (int j = 0; j < n; ++j) {
      double g = 0.0;
      for (int k = klo[j]; k <= khi[j] && k <= i; ++k)
        for (int p = 0; p < 4; ++p) g += M[(size_t)k * 4 + p] * Ct[((size_t)k * 4 + p) * n + j];
      G[(size_t)i * n + j] = g;
    }
  }
  if (spread) {
    // G <- S G,  S(i, j) = sum_m (w_m / sqrt(pi)) basis_j(E_i + sqrt(2) sigma_i t_m): spread_kernel of assemble.cu with the
    // 6-point Gauss-Hermite rule and the product's basis_eval_dev
    const double t[6] = {-2.350604973674492, -1.335849074013697, -0.436077411927617, 0.436077411927617, 1.335849074013697, 2.350604973674492};
    const double w[6] = {0.004530009905508846, 0.157067320322856, 0.724629595224392, 0.724629595224392, 0.157067320322856, 0.004530009905508846};
    const double sqrtpi = 1.7724538509055160273;
    std::vector<double> S((size_t)n * n), SG((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int m = 0; m < 6; ++m)
          acc += w[m] / sqrtpi * basis_eval_dev(n, ext.data(), invh.data(), Ct.data(), j, ext[i + 1] + sigE[i] * 1.4142135623730951 * t[m], 0);
        S[(size_t)i * n + j] = acc;
      }
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j) SG[(size_t)i * n + j] += S[(size_t)i * n + k] * G[(size_t)k * n + j];
    G.swap(SG);
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) std::printf("%.17g%c", G[(size_t)i * n + j], j + 1 < n ? ' ' : '\n');
  return 0;
}
