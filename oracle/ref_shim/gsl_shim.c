/*
 * gsl_shim.c -- the handful of GSL entry points the reference's assembly path calls, implemented on top
 * of the published-algorithm restatements in oracle/isr_oracle.c.
 *
 * ORACLE / TEST INFRASTRUCTURE.  Purpose: GSL is not installed in this image, so the reference cannot be
 * built with its own CMake recipe.  Its assembly-path SOURCES, however, compile unmodified from
 * /root/reference once <gsl/...> and <Eigen/Dense> resolve.  This file supplies the GSL side:
 *   gsl_integration_qags          -> orc_qags   (QUADPACK DQAGSE, 21-point GK + epsilon algorithm)
 *   gsl_integration_qag (key 6)   -> orc_qag61  (QUADPACK DQAGE, 61-point GK)
 *   gsl_integration_fixed_hermite -> 6-point Gauss-Hermite nodes mapped as GSL does (t/sqrt(b)+a, w/sqrt(b))
 *   gsl_spline (gsl_interp_cspline) -> natural cubic spline: tridiagonal solve orc_cspline_natural_c,
 *                                     evaluation formulas of GSL's cspline_eval / cspline_eval_deriv
 * Everything ABOVE these calls -- kernelKuraevFadin, convolutionKuraevFadin, the retry loops of
 * integrate/integrateS, gaussian_conv, the range interpolators and Interpolator -- is then the reference's
 * own compiled code (oracle/_ref/libisr_ref.so), which is what pins the CPU restatement in
 * oracle/isr_oracle.c.  The restated GSL internals themselves are pinned separately against SciPy's
 * QUADPACK and CubicSpline (tests/test_oracle.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../gh_tables.h"
#include "gsl/gsl_errno.h"
#include "gsl/gsl_integration.h"
#include "gsl/gsl_spline.h"

typedef double (*orc_fn)(double x, void* ctx);
int orc_qags(orc_fn f, void* ctx, double a, double b, double epsabs, double epsrel, size_t limit, double* result,
             double* abserr);
int orc_qag61(orc_fn f, void* ctx, double a, double b, double epsabs, double epsrel, size_t limit, double* result,
              double* abserr);
void orc_cspline_natural_c(int size, const double* xa, const double* ya, double* c);
int orc_bsearch(const double* xa, double x, int index_lo, int index_hi);

/* ---- gsl_errno ---------------------------------------------------------------------------------- */
static gsl_error_handler_t* g_handler = NULL;
gsl_error_handler_t* gsl_set_error_handler_off(void) {
  gsl_error_handler_t* old = g_handler;
  g_handler = NULL;
  return old;
}
gsl_error_handler_t* gsl_set_error_handler(gsl_error_handler_t* h) {
  gsl_error_handler_t* old = g_handler;
  g_handler = h;
  return old;
}

/* ---- adaptive integration ----------------------------------------------------------------------- */
/* The restated routines allocate their own interval lists (of `limit` entries, as GSL's workspace has),
 * so the GSL workspace object only records its size. */
gsl_integration_workspace* gsl_integration_workspace_alloc(const size_t n) {
  gsl_integration_workspace* w = (gsl_integration_workspace*)malloc(sizeof(*w));
  if (w) w->limit = n;
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace* w) { free(w); }

int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace* workspace, double* result, double* abserr) {
  (void)workspace;
  return orc_qags(f->function, f->params, a, b, epsabs, epsrel, limit, result, abserr);
}
int gsl_integration_qag(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                        int key, gsl_integration_workspace* workspace, double* result, double* abserr) {
  (void)workspace;
  if (key != 6) { /* only GSL_INTEG_GAUSS61 is used by the reference (src/Integration.cpp:66) */
    *result = NAN;
    *abserr = NAN;
    return 4; /* GSL_EINVAL */
  }
  return orc_qag61(f->function, f->params, a, b, epsabs, epsrel, limit, result, abserr);
}

/* ---- fixed Gauss-Hermite ------------------------------------------------------------------------- */
static const gsl_integration_fixed_type hermite_type = {1};
const gsl_integration_fixed_type* gsl_integration_fixed_hermite = &hermite_type;

/* weight exp(-b (x - a)^2): nodes a + t_m / sqrt(b), weights w_m / sqrt(b) (alpha = 0) */
gsl_integration_fixed_workspace* gsl_integration_fixed_alloc(const gsl_integration_fixed_type* type, const size_t n,
                                                             const double a, const double b, const double alpha,
                                                             const double beta) {
  gsl_integration_fixed_workspace* w;
  size_t m;
  (void)beta;
  if (type != gsl_integration_fixed_hermite || n != 6 || alpha != 0.0) return NULL;
  w = (gsl_integration_fixed_workspace*)malloc(sizeof(*w));
  w->n = n;
  w->x = (double*)malloc(sizeof(double) * n);
  w->weights = (double*)malloc(sizeof(double) * n);
  for (m = 0; m < n; m++) {
    w->x[m] = a + gh6_t[m] / sqrt(b);
    w->weights[m] = gh6_w[m] / sqrt(b);
  }
  return w;
}
void gsl_integration_fixed_free(gsl_integration_fixed_workspace* w) {
  if (!w) return;
  free(w->x);
  free(w->weights);
  free(w);
}
int gsl_integration_fixed(const gsl_function* func, double* result, const gsl_integration_fixed_workspace* w) {
  double sum = 0.0;
  size_t m;
  for (m = 0; m < w->n; m++) sum += w->weights[m] * func->function(w->x[m], func->params);
  *result = sum;
  return 0;
}

/* ---- natural cubic spline ------------------------------------------------------------------------ */
static const gsl_interp_type cspline_type = {1};
const gsl_interp_type* gsl_interp_cspline = &cspline_type;

gsl_interp_accel* gsl_interp_accel_alloc(void) { return (gsl_interp_accel*)calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel* a) { free(a); }

gsl_spline* gsl_spline_alloc(const gsl_interp_type* T, size_t size) {
  gsl_spline* s;
  if (T != gsl_interp_cspline) return NULL;
  s = (gsl_spline*)malloc(sizeof(*s));
  s->size = size;
  s->x = (double*)malloc(sizeof(double) * size);
  s->y = (double*)malloc(sizeof(double) * size);
  s->c = (double*)malloc(sizeof(double) * size);
  return s;
}
int gsl_spline_init(gsl_spline* s, const double xa[], const double ya[], size_t size) {
  if (size != s->size) return 4;
  memcpy(s->x, xa, sizeof(double) * size);
  memcpy(s->y, ya, sizeof(double) * size);
  orc_cspline_natural_c((int)size, s->x, s->y, s->c);
  return 0;
}
void gsl_spline_free(gsl_spline* s) {
  if (!s) return;
  free(s->x);
  free(s->y);
  free(s->c);
  free(s);
}
/* GSL cspline_eval / cspline_eval_deriv; the accelerator only caches the last interval and does not
 * change which interval is found, so it is ignored (this also makes evaluation re-entrant). */
static double spline_eval(const gsl_spline* s, double x, int deriv) {
  const int size = (int)s->size;
  int index;
  if (x < s->x[0] || x > s->x[size - 1]) return NAN; /* GSL_EDOM -> GSL_NAN */
  index = orc_bsearch(s->x, x, 0, size - 1);
  {
    const double x_hi = s->x[index + 1], x_lo = s->x[index], dx = x_hi - x_lo;
    if (dx > 0.0) {
      const double y_lo = s->y[index], y_hi = s->y[index + 1];
      const double dy = y_hi - y_lo, delx = x - x_lo;
      const double c_i = s->c[index], c_ip1 = s->c[index + 1];
      const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
      const double d_i = (c_ip1 - c_i) / (3.0 * dx);
      if (deriv) return b_i + delx * (2.0 * c_i + 3.0 * d_i * delx);
      return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
    }
    return 0.0;
  }
}
double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel* a) {
  (void)a;
  return spline_eval(s, x, 0);
}
double gsl_spline_eval_deriv(const gsl_spline* s, double x, gsl_interp_accel* a) {
  (void)a;
  return spline_eval(s, x, 1);
}
