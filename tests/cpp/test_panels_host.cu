// The numerical core of the matrix assembly -- the product's own panel enumeration (item_cuts / walk_panels), panel
// kinds, singularity-removing substitutions (panel_node) and integrand (kf_value), isrsolver_b200/csrc/panel_device.cuh
// compiled for the HOST -- against the oracle's adaptive QUADPACK integration of the reference integrand
// (oracle/isr_oracle.c: orc_kernel_kf, orc_integrate = QAG 61, orc_integrateS = QAGS, the routines src/Integration.cpp
// calls).  Checked quantity: the four moments M_p = int F eps ((E' - ext[k]) / h_k)^p dx of every (row, segment) item of a
// small problem, with eps = 1 and with a tabulated (piecewise constant) efficiency.  Tolerance 1e-10 relative to the
// item's M_0 (north star: matrix entries within 1e-9).  No GPU is needed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "panel_device.cuh"

extern "C" double orc_kernel_kf(double x, double s);
typedef double (*orc_fn)(double, void*);
extern "C" double orc_integrate(orc_fn f, void* ctx, double a, double b, double* error);
extern "C" double orc_integrateS(orc_fn f, void* ctx, double a, double b, double* error);

using namespace isr;

struct Ctx { double E, s, extk, invh, eps; int p; };
static double integrand(double x, void* v) {
  const Ctx* c = (const Ctx*)v;
  const double dn = (c->E * std::sqrt(1.0 - x) - c->extk) * c->invh;
  return orc_kernel_kf(x, c->s) * c->eps * std::pow(dn, c->p);
}

// Gauss-Legendre nodes / weights by Newton iteration on P_n (independent of the product's tables)
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
  const int n = 14, nx = 16;
  const double thr = 0.827;
  std::vector<double> ext(n + 1), invh(n);
  ext[0] = thr;
  for (int i = 0; i < n; ++i) ext[i + 1] = 1.18 + (2.0 - 1.18) * i / (n - 1);
  for (int k = 0; k < n; ++k) invh[k] = 1.0 / (ext[k + 1] - ext[k]);
  std::vector<double> glx, glw;
  gauss_legendre(16, glx, glw);
  // tabulated efficiency: per-row tables of nx bins on [0, 1), under / overflow cells 0 (ROOT layout [N][nx + 2])
  std::vector<double> vals((size_t)n * (nx + 2), 0.0);
  for (int i = 0; i < n; ++i)
    for (int b = 1; b <= nx; ++b) vals[(size_t)i * (nx + 2) + b] = 0.25 + 0.5 * std::exp(-40.0 * (b - 0.5) / nx) * (1.0 - 0.1 * (ext[i + 1] - 1.5));
  long items = 0, panels = 0;
  double worst = 0.0;
  for (int table = 0; table <= 1; ++table) {
    EpsDev eps{table ? 2 : 0, table ? nx : 0, 0.0, 1.0, nullptr, table ? vals.data() : nullptr};
    for (int i = 0; i < n; ++i) {
      RowConst rc;
      row_const_init(rc, ext[i + 1], thr);
      row_grading(rc);
      for (int k = 0; k <= i; ++k) {
        const ItemCuts c = item_cuts(rc, eps, ext.data(), i, k);
        double m[4] = {0, 0, 0, 0};
        walk_panels(c, rc, eps, 0, c.count(), [&](int, double a, double b) {
          const int kind = panel_kind(rc, b);
          const double ev = table ? vals[(size_t)i * (nx + 2) + eps_find_bin(eps, 0.5 * (a + b))] : 1.0;   // as eval_kernel does
          for (int q = 0; q < 16; ++q) {
            const PanelNode pn = panel_node(rc, a, b, kind, glx[q], glw[q]);
            const double f = kf_value(rc, pn.x, pn.lnx, pn.pw) * pn.wj * ev;
            const double dn = (rc.E * std::sqrt(1.0 - pn.x) - ext[k]) * invh[k];
            m[0] += f; m[1] += f * dn; m[2] += f * dn * dn; m[3] += f * dn * dn * dn;
          }
          ++panels;
        });
        // oracle: the same integral, cut at the pair threshold x0 and at the bin edges so that every adaptive call sees
        // a smooth integrand (the "converged" evaluation of DESIGN 3)
        std::vector<double> cuts = {c.xa, c.xb};
        if (rc.x0 > c.xa && rc.x0 < c.xb) cuts.push_back(rc.x0);
        if (table) for (int e = 0; e <= nx; ++e) { const double v = (double)e / nx; if (v > c.xa && v < c.xb) cuts.push_back(v); }
        std::sort(cuts.begin(), cuts.end());
        for (int p = 0; p < 4; ++p) {
          double ref = 0.0;
          for (size_t q = 0; q + 1 < cuts.size(); ++q) {
            const double a = cuts[q], b = cuts[q + 1];
            if (!(b > a)) continue;
            Ctx cx{rc.E, rc.s, ext[k], invh[k], table ? vals[(size_t)i * (nx + 2) + eps_find_bin(eps, 0.5 * (a + b))] : 1.0, p};
            double err;
            ref += (a == 0.0) ? orc_integrateS(integrand, &cx, a, b, &err) : orc_integrate(integrand, &cx, a, b, &err);
          }
          if (p == 0 && !(c.xb > c.xa)) continue;
          static double scale;
          if (p == 0) scale = std::fabs(ref);
          const double d = std::fabs(m[p] - ref) / (scale > 0 ? scale : 1.0);
          if (d > worst) worst = d;
          if (!(d <= 1e-10)) {
            std::printf("FAILED: table %d row %d segment %d moment %d: panels %.17g, oracle %.17g (rel. to M0: %.3e)\n", table, i, k, p, m[p], ref, d);
            return 1;
          }
        }
        ++items;
      }
    }
  }
  std::printf("panel quadrature ok on the host: %ld items, %ld panels, worst deviation %.2e of M0\n", items, panels, worst);
  return 0;
}
