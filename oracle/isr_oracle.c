/*
 * isr_oracle.c -- CPU restatement of the ISRSolver hot path (matrix assembly side).
 *
 * ORACLE / TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libisr_b200.so) never links or calls it.
 *
 * PINNING.  sergeigribanov/ISRSolver ships no tests, golden vectors or data files, and its own build
 * (CMake + GSL + Eigen + ROOT + Boost + NLopt) cannot run here.  Its assembly-path SOURCES, however, are
 * compiled unmodified into oracle/_ref/libisr_ref.so (oracle/Makefile) with the GSL entry points they call
 * supplied by oracle/ref_shim on top of the QUADPACK / cspline routines of THIS file; every function of the
 * restatement below (kernel, convolution split, retry loops, range bookkeeping, bases, N^2 loop, spread
 * matrix) reproduces that compiled reference BIT FOR BIT (tests/test_ref_pin.py).  The GSL internals
 * themselves, shared by both, are pinned against (i) SciPy's QUADPACK (identical evaluation counts),
 * (ii) mpmath high-precision quadrature and (iii) analytic identities -- tests/test_oracle.py.
 *
 * Third-party algorithms restated here (not in /root/reference, versions unpinned by the reference:
 * Dockerfile:1,5, docker/install.sh:14-17):
 *   GSL  gsl_integration_qags / gsl_integration_qag(key=6) == QUADPACK DQAGSE / DQAGE + DQK21 / DQK61,
 *        DQELG (epsilon algorithm), DQPSRT (ordering)            [Piessens et al., QUADPACK, 1983]
 *   GSL  gsl_interp_cspline (natural cubic spline, symmetric tridiagonal solve), gsl_interp_bsearch
 *   GSL  gsl_integration_fixed_hermite, n = 6
 *   ROOT TAxis::FindBin / FindFixBin bin arithmetic
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 *
 * build: see oracle/Makefile  (gcc -O2 -fPIC -shared -pthread; -ffp-contract=off keeps the
 *        reference's expression order free of FMA contraction)
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gh_tables.h"
#include "gk_tables.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* include/PhysicalConstants.hpp:7,11 */
#define ALPHA_QED 7.2973525698e-03
#define ELECTRON_M 5.10998928e-04

typedef double (*orc_fn)(double x, void *ctx);

static _Thread_local long g_neval = 0; /* integrand evaluations of the calling thread (statistics only) */
static _Thread_local double g_abserr_sum = 0;   /* sum of QUADPACK abserr over the calls since reset */
static _Thread_local double g_max_relerr = 0;   /* loosest epsrel a retry loop ended with since reset */
static _Thread_local double g_err_weight = 1;   /* |coefficient| the current integral is multiplied by */
long orc_get_neval(void) { return g_neval; }
void orc_reset_neval(void) { g_neval = 0; }

/* ------------------------------------------------------------------------------------------------
 * Kuraev-Fadin kernel                                                   src/KuraevFadin.cpp:10-49
 * ---------------------------------------------------------------------------------------------- */
static double fLog(double s) { return log(s / ELECTRON_M / ELECTRON_M); }             /* :10 */
static double fBeta(double s) { return (2 * ALPHA_QED) / M_PI * (fLog(s) - 1); }      /* :12 */

double orc_kernel_kf(double x, double s) {                                            /* :14-49 */
  double lnX = log(x);
  double mX = 1 - x;
  double lnMX = log(mX);
  double sM = s / ELECTRON_M / ELECTRON_M;
  double logSM = log(sM);
  double E = 0.5 * sqrt(s);
  double res = 0;
  double part1 = fBeta(s) * pow(x, fBeta(s) - 1) *
                 (1 + ALPHA_QED / M_PI * (M_PI * M_PI / 3 - 0.5) + 0.75 * fBeta(s) -
                  fBeta(s) * fBeta(s) / 24 * (fLog(s) / 3 + 2 * M_PI * M_PI - 9.25));
  double part2 = -fBeta(s) * (1 - 0.5 * x);
  double part3 = 0.125 * fBeta(s) * fBeta(s) *
                 (-4 * (2 - x) * lnX - (1 + 3 * mX * mX) / x * lnMX - 6 + x);
  res = part1 + part2 + part3;
  if (x > 2 * ELECTRON_M / E) {
    double subsubpart1 = 2 * lnX + logSM - 5. / 3;
    double subpart1 = pow(x - 2 * ELECTRON_M / E, fBeta(s)) / 6 / x * pow(subsubpart1, 2) *
                      (2 - 2 * x + x * x + fBeta(s) / 3 * subsubpart1);
    double subpart2 = 0.5 * fLog(s) * fLog(s) *
                      (2. / 3 * (1 - mX * mX * mX) / mX + (2 - x) * lnMX + 0.5 * x);
    res += (ALPHA_QED / M_PI) * (ALPHA_QED / M_PI) * (subpart1 + subpart2);
  }
  return res;
}

double orc_beta(double s) { return fBeta(s); }

/* ------------------------------------------------------------------------------------------------
 * QUADPACK restatement (what gsl_integration_qag / qags execute)
 * ---------------------------------------------------------------------------------------------- */
enum { ORC_SUCCESS = 0, ORC_EMAXITER = 11, ORC_EROUND = 18, ORC_ESING = 21, ORC_EDIVERGE = 22,
       ORC_EBADTOL = 13, ORC_EFAILED = 5 };

static double rescale_error(double err, const double result_abs, const double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    double scale = pow((200 * err / result_asc), 1.5);
    if (scale < 1) err = result_asc * scale; else err = result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

/* generic (2n+1)-point Kronrod evaluation, QUADPACK layout: xgk[0..n] descending, xgk[n]=0;
 * wgf = Gauss weights aligned with xgk (0 at Kronrod-only abscissae). */
static void qk(int n, const double *xgk, const double *wgk, const double *wgf, orc_fn f, void *ctx,
               double a, double b, double *result, double *abserr, double *resabs, double *resasc) {
  double fv1[31], fv2[31];
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  const double f_center = f(center, ctx);
  double result_gauss = f_center * wgf[n];
  double result_kronrod = f_center * wgk[n];
  double result_abs = fabs(result_kronrod);
  double result_asc, mean, err;
  int j;
  for (j = 0; j < n; j++) {
    const double abscissa = half_length * xgk[j];
    const double f1 = f(center - abscissa, ctx);
    const double f2 = f(center + abscissa, ctx);
    fv1[j] = f1; fv2[j] = f2;
    result_gauss += wgf[j] * (f1 + f2);
    result_kronrod += wgk[j] * (f1 + f2);
    result_abs += wgk[j] * (fabs(f1) + fabs(f2));
  }
  g_neval += 2 * n + 1;
  mean = result_kronrod * 0.5;
  result_asc = wgk[n] * fabs(f_center - mean);
  for (j = 0; j < n; j++) result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
  err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  *result = result_kronrod; *resabs = result_abs; *resasc = result_asc;
  *abserr = rescale_error(err, result_abs, result_asc);
}
static void qk21(orc_fn f, void *c, double a, double b, double *r, double *e, double *ra, double *rs) {
  qk(10, xgk21, wgk21, wgf21, f, c, a, b, r, e, ra, rs);
}
static void qk61(orc_fn f, void *c, double a, double b, double *r, double *e, double *ra, double *rs) {
  qk(30, xgk61, wgk61, wgf61, f, c, a, b, r, e, ra, rs);
}

typedef struct {
  size_t limit, size, nrmax, i, maximum_level;
  double *alist, *blist, *rlist, *elist;
  size_t *order, *level;
} ws_t;

static ws_t *ws_alloc(size_t n) {
  ws_t *w = (ws_t *)malloc(sizeof(ws_t));
  w->alist = (double *)malloc(n * sizeof(double));
  w->blist = (double *)malloc(n * sizeof(double));
  w->rlist = (double *)malloc(n * sizeof(double));
  w->elist = (double *)malloc(n * sizeof(double));
  w->order = (size_t *)malloc(n * sizeof(size_t));
  w->level = (size_t *)malloc(n * sizeof(size_t));
  w->size = 0; w->limit = n; w->maximum_level = 0;
  return w;
}
static void ws_free(ws_t *w) {
  free(w->alist); free(w->blist); free(w->rlist); free(w->elist); free(w->order); free(w->level); free(w);
}
static void ws_initialise(ws_t *w, double a, double b) {
  w->size = 0; w->nrmax = 0; w->i = 0;
  w->alist[0] = a; w->blist[0] = b; w->rlist[0] = 0.0; w->elist[0] = 0.0;
  w->order[0] = 0; w->level[0] = 0; w->maximum_level = 0;
}
static void ws_set_initial_result(ws_t *w, double result, double error) {
  w->size = 1; w->rlist[0] = result; w->elist[0] = error;
}
static void ws_retrieve(const ws_t *w, double *a, double *b, double *r, double *e) {
  const size_t i = w->i;
  *a = w->alist[i]; *b = w->blist[i]; *r = w->rlist[i]; *e = w->elist[i];
}
static double ws_sum_results(const ws_t *w) {
  size_t k; double s = 0;
  for (k = 0; k < w->size; k++) s += w->rlist[k];
  return s;
}
static int subinterval_too_small(double a1, double a2, double b2) {
  const double e = DBL_EPSILON, u = DBL_MIN;
  double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
  return fabs(a1) <= tmp && fabs(b2) <= tmp;
}
/* DQPSRT */
static void qpsrt(ws_t *w) {
  const size_t last = w->size - 1, limit = w->limit;
  double *elist = w->elist; size_t *order = w->order;
  double errmax, errmin; int i, k, top;
  size_t i_nrmax = w->nrmax, i_maxerr = order[i_nrmax];
  if (last < 2) { order[0] = 0; order[1] = 1; w->i = i_maxerr; return; }
  errmax = elist[i_maxerr];
  while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) { order[i_nrmax] = order[i_nrmax - 1]; i_nrmax--; }
  if (last < (limit / 2 + 2)) top = (int)last; else top = (int)(limit - last + 1);
  i = (int)i_nrmax + 1;
  while (i < top && errmax < elist[order[i]]) { order[i - 1] = order[i]; i++; }
  order[i - 1] = i_maxerr;
  errmin = elist[last];
  k = top - 1;
  while (k > i - 2 && errmin >= elist[order[k]]) { order[k + 1] = order[k]; k--; }
  order[k + 1] = last;
  i_maxerr = order[i_nrmax];
  w->i = i_maxerr; w->nrmax = i_nrmax;
}
static void ws_update(ws_t *w, double a1, double b1, double area1, double error1,
                      double a2, double b2, double area2, double error2) {
  const size_t i_max = w->i, i_new = w->size;
  const size_t new_level = w->level[i_max] + 1;
  if (error2 > error1) {
    w->alist[i_max] = a2; w->rlist[i_max] = area2; w->elist[i_max] = error2; w->level[i_max] = new_level;
    w->alist[i_new] = a1; w->blist[i_new] = b1; w->rlist[i_new] = area1; w->elist[i_new] = error1;
    w->level[i_new] = new_level;
  } else {
    w->blist[i_max] = b1; w->rlist[i_max] = area1; w->elist[i_max] = error1; w->level[i_max] = new_level;
    w->alist[i_new] = a2; w->blist[i_new] = b2; w->rlist[i_new] = area2; w->elist[i_new] = error2;
    w->level[i_new] = new_level;
  }
  w->size++;
  if (new_level > w->maximum_level) w->maximum_level = new_level;
  qpsrt(w);
}

/* DQAGE with the 61-point rule: gsl_integration_qag(..., key = 6, ...) */
int orc_qag61(orc_fn f, void *ctx, double a, double b, double epsabs, double epsrel, size_t limit,
              double *result, double *abserr) {
  ws_t *w = ws_alloc(limit);
  double area, errsum, result0, abserr0, resabs0, resasc0, tolerance, round_off;
  size_t iteration = 0; int roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0, status;
  ws_initialise(w, a, b);
  *result = 0; *abserr = 0;
  if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) { ws_free(w); return ORC_EBADTOL; }
  qk61(f, ctx, a, b, &result0, &abserr0, &resabs0, &resasc0);
  ws_set_initial_result(w, result0, abserr0);
  tolerance = fmax(epsabs, epsrel * fabs(result0));
  round_off = 50 * DBL_EPSILON * resabs0;
  if (abserr0 <= round_off && abserr0 > tolerance) {
    *result = result0; *abserr = abserr0; ws_free(w); return ORC_EROUND;
  } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
    *result = result0; *abserr = abserr0; ws_free(w); return ORC_SUCCESS;
  } else if (limit == 1) {
    *result = result0; *abserr = abserr0; ws_free(w); return ORC_EMAXITER;
  }
  area = result0; errsum = abserr0; iteration = 1;
  do {
    double a1, b1, a2, b2, a_i, b_i, r_i, e_i, area1 = 0, area2 = 0, area12 = 0;
    double error1 = 0, error2 = 0, error12 = 0, resasc1, resasc2, resabs1, resabs2;
    ws_retrieve(w, &a_i, &b_i, &r_i, &e_i);
    a1 = a_i; b1 = 0.5 * (a_i + b_i); a2 = b1; b2 = b_i;
    qk61(f, ctx, a1, b1, &area1, &error1, &resabs1, &resasc1);
    qk61(f, ctx, a2, b2, &area2, &error2, &resabs2, &resasc2);
    area12 = area1 + area2; error12 = error1 + error2;
    errsum += (error12 - e_i); area += area12 - r_i;
    if (resasc1 != error1 && resasc2 != error2) {
      double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
      if (iteration >= 10 && error12 > e_i) roundoff_type2++;
    }
    tolerance = fmax(epsabs, epsrel * fabs(area));
    if (errsum > tolerance) {
      if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
      if (subinterval_too_small(a1, a2, b2)) error_type = 3;
    }
    ws_update(w, a1, b1, area1, error1, a2, b2, area2, error2);
    ws_retrieve(w, &a_i, &b_i, &r_i, &e_i);
    iteration++;
  } while (iteration < limit && !error_type && errsum > tolerance);
  *result = ws_sum_results(w); *abserr = errsum;
  if (errsum <= tolerance) status = ORC_SUCCESS;
  else if (error_type == 2) status = ORC_EROUND;
  else if (error_type == 3) status = ORC_ESING;
  else if (iteration == limit) status = ORC_EMAXITER;
  else status = ORC_EFAILED;
  ws_free(w);
  return status;
}

/* DQELG */
typedef struct { size_t n; double rlist2[52]; size_t nres; double res3la[3]; } ext_table;
static void table_init(ext_table *t) { t->n = 0; t->nres = 0; }
static void table_append(ext_table *t, double y) { t->rlist2[t->n] = y; t->n++; }
static void qelg(ext_table *table, double *result, double *abserr) {
  double *epstab = table->rlist2, *res3la = table->res3la;
  const size_t n = table->n - 1;
  const double current = epstab[n];
  double absolute = DBL_MAX, relative = 5 * DBL_EPSILON * fabs(current);
  const size_t newelm = n / 2, n_orig = n; size_t n_final = n, i;
  const size_t nres_orig = table->nres;
  *result = current; *abserr = DBL_MAX;
  if (n < 2) { *result = current; *abserr = fmax(absolute, relative); return; }
  epstab[n + 2] = epstab[n]; epstab[n] = DBL_MAX;
  for (i = 0; i < newelm; i++) {
    double res = epstab[n - 2 * i + 2];
    double e0 = epstab[n - 2 * i - 2], e1 = epstab[n - 2 * i - 1], e2 = res;
    double e1abs = fabs(e1), delta2 = e2 - e1, err2 = fabs(delta2), tol2 = fmax(fabs(e2), e1abs) * DBL_EPSILON;
    double delta3 = e1 - e0, err3 = fabs(delta3), tol3 = fmax(e1abs, fabs(e0)) * DBL_EPSILON;
    double e3, delta1, err1, tol1, ss;
    if (err2 <= tol2 && err3 <= tol3) {
      *result = res; absolute = err2 + err3; relative = 5 * DBL_EPSILON * fabs(res);
      *abserr = fmax(absolute, relative); return;
    }
    e3 = epstab[n - 2 * i]; epstab[n - 2 * i] = e1;
    delta1 = e1 - e3; err1 = fabs(delta1); tol1 = fmax(e1abs, fabs(e3)) * DBL_EPSILON;
    if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) { n_final = 2 * i; break; }
    ss = (1 / delta1 + 1 / delta2) - 1 / delta3;
    if (fabs(ss * e1) <= 0.0001) { n_final = 2 * i; break; }
    res = e1 + 1 / ss; epstab[n - 2 * i] = res;
    { const double error = err2 + fabs(res - e2) + err3;
      if (error <= *abserr) { *abserr = error; *result = res; } }
  }
  { const size_t limexp = 50 - 1; if (n_final == limexp) n_final = 2 * (limexp / 2); }
  if (n_orig % 2 == 1) { for (i = 0; i <= newelm; i++) epstab[1 + i * 2] = epstab[i * 2 + 3]; }
  else { for (i = 0; i <= newelm; i++) epstab[i * 2] = epstab[i * 2 + 2]; }
  if (n_orig != n_final) { for (i = 0; i <= n_final; i++) epstab[i] = epstab[n_orig - n_final + i]; }
  table->n = n_final + 1;
  if (nres_orig < 3) { res3la[nres_orig] = *result; *abserr = DBL_MAX; }
  else {
    *abserr = (fabs(*result - res3la[2]) + fabs(*result - res3la[1]) + fabs(*result - res3la[0]));
    res3la[0] = res3la[1]; res3la[1] = res3la[2]; res3la[2] = *result;
  }
  table->nres = nres_orig + 1;
  *abserr = fmax(*abserr, 5 * DBL_EPSILON * fabs(*result));
}
static int large_interval(ws_t *w) { return w->level[w->i] < w->maximum_level; }
static int increase_nrmax(ws_t *w) {
  int k; int id = (int)w->nrmax; int jupbnd;
  const size_t limit = w->limit, last = w->size - 1;
  if (last > (1 + limit / 2)) jupbnd = (int)(limit + 1 - last); else jupbnd = (int)last;
  for (k = id; k <= jupbnd; k++) {
    size_t i_max = w->order[w->nrmax];
    w->i = i_max;
    if (w->level[i_max] < w->maximum_level) return 1;
    w->nrmax++;
  }
  return 0;
}
static void reset_nrmax(ws_t *w) { w->nrmax = 0; w->i = w->order[0]; }
static int test_positivity(double result, double resabs) {
  return (fabs(result) >= (1 - 50 * DBL_EPSILON) * resabs);
}

/* DQAGSE with the 21-point rule: gsl_integration_qags */
int orc_qags(orc_fn f, void *ctx, double a, double b, double epsabs, double epsrel, size_t limit,
             double *result, double *abserr) {
  ws_t *w = ws_alloc(limit);
  double area, errsum, res_ext, err_ext, result0, abserr0, resabs0, resasc0, tolerance;
  double ertest = 0, error_over_large_intervals = 0, reseps = 0, abseps = 0, correc = 0;
  size_t ktmin = 0; int roundoff_type1 = 0, roundoff_type2 = 0, roundoff_type3 = 0;
  int error_type = 0, error_type2 = 0; size_t iteration = 0;
  int positive_integrand = 0, extrapolate = 0, disallow_extrapolation = 0, status;
  ext_table table;
  ws_initialise(w, a, b);
  *result = 0; *abserr = 0;
  if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) { ws_free(w); return ORC_EBADTOL; }
  qk21(f, ctx, a, b, &result0, &abserr0, &resabs0, &resasc0);
  ws_set_initial_result(w, result0, abserr0);
  tolerance = fmax(epsabs, epsrel * fabs(result0));
  if (abserr0 <= 100 * DBL_EPSILON * resabs0 && abserr0 > tolerance) {
    *result = result0; *abserr = abserr0; ws_free(w); return ORC_EROUND;
  } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
    *result = result0; *abserr = abserr0; ws_free(w); return ORC_SUCCESS;
  } else if (limit == 1) {
    *result = result0; *abserr = abserr0; ws_free(w); return ORC_EMAXITER;
  }
  table_init(&table); table_append(&table, result0);
  area = result0; errsum = abserr0; res_ext = result0; err_ext = DBL_MAX;
  positive_integrand = test_positivity(result0, resabs0);
  iteration = 1;
  do {
    size_t current_level;
    double a1, b1, a2, b2, a_i, b_i, r_i, e_i, area1 = 0, area2 = 0, area12 = 0;
    double error1 = 0, error2 = 0, error12 = 0, resasc1, resasc2, resabs1, resabs2, last_e_i;
    ws_retrieve(w, &a_i, &b_i, &r_i, &e_i);
    current_level = w->level[w->i] + 1;
    a1 = a_i; b1 = 0.5 * (a_i + b_i); a2 = b1; b2 = b_i;
    iteration++;
    qk21(f, ctx, a1, b1, &area1, &error1, &resabs1, &resasc1);
    qk21(f, ctx, a2, b2, &area2, &error2, &resabs2, &resasc2);
    area12 = area1 + area2; error12 = error1 + error2; last_e_i = e_i;
    errsum = errsum + error12 - e_i; area = area + area12 - r_i;
    tolerance = fmax(epsabs, epsrel * fabs(area));
    if (resasc1 != error1 && resasc2 != error2) {
      double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) {
        if (!extrapolate) roundoff_type1++; else roundoff_type2++;
      }
      if (iteration > 10 && error12 > e_i) roundoff_type3++;
    }
    if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error_type = 2;
    if (roundoff_type2 >= 5) error_type2 = 1;
    if (subinterval_too_small(a1, a2, b2)) error_type = 4;
    ws_update(w, a1, b1, area1, error1, a2, b2, area2, error2);
    if (errsum <= tolerance) goto compute_result;
    if (error_type) break;
    if (iteration >= limit - 1) { error_type = 1; break; }
    if (iteration == 2) {
      error_over_large_intervals = errsum; ertest = tolerance; table_append(&table, area); continue;
    }
    if (disallow_extrapolation) continue;
    error_over_large_intervals += -last_e_i;
    if (current_level < w->maximum_level) error_over_large_intervals += error12;
    if (!extrapolate) {
      if (large_interval(w)) continue;
      extrapolate = 1; w->nrmax = 1;
    }
    if (!error_type2 && error_over_large_intervals > ertest) { if (increase_nrmax(w)) continue; }
    table_append(&table, area);
    qelg(&table, &reseps, &abseps);
    ktmin++;
    if (ktmin > 5 && err_ext < 0.001 * errsum) error_type = 5;
    if (abseps < err_ext) {
      ktmin = 0; err_ext = abseps; res_ext = reseps; correc = error_over_large_intervals;
      ertest = fmax(epsabs, epsrel * fabs(reseps));
      if (err_ext <= ertest) break;
    }
    if (table.n == 1) disallow_extrapolation = 1;
    if (error_type == 5) break;
    reset_nrmax(w); extrapolate = 0; error_over_large_intervals = errsum;
  } while (iteration < limit);
  *result = res_ext; *abserr = err_ext;
  if (err_ext == DBL_MAX) goto compute_result;
  if (error_type || error_type2) {
    if (error_type2) err_ext += correc;
    if (error_type == 0) error_type = 3;
    if (res_ext != 0.0 && area != 0.0) {
      if (err_ext / fabs(res_ext) > errsum / fabs(area)) goto compute_result;
    } else if (err_ext > errsum) goto compute_result;
    else if (area == 0.0) goto return_error;
  }
  { double max_area = fmax(fabs(res_ext), fabs(area));
    if (!positive_integrand && max_area < 0.01 * resabs0) goto return_error; }
  { double ratio = res_ext / area;
    if (ratio < 0.01 || ratio > 100.0 || errsum > fabs(area)) error_type = 6; }
  goto return_error;
compute_result:
  *result = ws_sum_results(w); *abserr = errsum;
return_error:
  if (error_type > 2) error_type--;
  if (error_type == 0) status = ORC_SUCCESS;
  else if (error_type == 1) status = ORC_EMAXITER;
  else if (error_type == 2) status = ORC_EROUND;
  else if (error_type == 3) status = ORC_ESING;
  else if (error_type == 4) status = ORC_EROUND;
  else if (error_type == 5) status = ORC_EDIVERGE;
  else status = ORC_EFAILED;
  ws_free(w);
  return status;
}

/* src/Integration.cpp:19-48 -- QAGS, epsabs 1e-12, epsrel 1e-12 loosened x10 per retry.
 * Like the reference, a 1e5-interval (resp. 1e6) workspace is allocated and freed on every call. */
#define ORC_LIMIT_S 100000   /* src/Integration.cpp:21 */
#define ORC_LIMIT 1000000    /* src/Integration.cpp:55 */

double orc_integrateS(orc_fn f, void *ctx, double a, double b, double *error) {
  double result = 0, relerr = 1.0e-12; int status = 1;
  while (status) {
    status = orc_qags(f, ctx, a, b, 1.e-12, relerr, ORC_LIMIT_S, &result, error);
    if (!status && relerr > g_max_relerr) g_max_relerr = relerr;
    if (relerr < 1.e-3) relerr *= 10; else relerr *= 1.1;
  }
  g_abserr_sum += g_err_weight * fabs(*error);
  return result;
}
/* src/Integration.cpp:53-81 -- QAG key 6 */
double orc_integrate(orc_fn f, void *ctx, double a, double b, double *error) {
  double result = 0, relerr = 1.0e-12; int status = 1;
  while (status) {
    status = orc_qag61(f, ctx, a, b, 1.e-12, relerr, ORC_LIMIT, &result, error);
    if (!status && relerr > g_max_relerr) g_max_relerr = relerr;
    if (relerr < 1.e-3) relerr *= 10; else relerr *= 1.1;
  }
  g_abserr_sum += g_err_weight * fabs(*error);
  return result;
}
/* src/Integration.cpp:86-100 -- 6-point Gauss-Hermite, weight exp(-(x-E)^2/(2 sigma2)), then
 * divided by sqrt(2 pi sigma2).  GSL's fixed rule maps t -> a + t/sqrt(b), w -> w/sqrt(b), b = 0.5/sigma2. */
double orc_gaussian_conv(double energy, double sigma2, orc_fn f, void *ctx) {
  const double b = 0.5 / sigma2;
  double result = 0; int m;
  for (m = 0; m < 6; m++) result += (gh6_w[m] / sqrt(b)) * f(energy + gh6_t[m] / sqrt(b), ctx);
  result /= sqrt(2 * M_PI * sigma2);
  return result;
}

/* ------------------------------------------------------------------------------------------------
 * Detection efficiency (ROOT bin arithmetic)    src/BaseISRSolver.cpp:232-258, BaseISRSolver2.cpp:231-236
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int kind;             /* 0: 1; 1: per-row x histogram (v2, TH1D per point); 2: TEfficiency 2-D (x,E);
                           3: TEfficiency 1-D (function of E only) */
  int nx; double xlo, xhi; const double *xedges; /* xedges NULL => fixed bins */
  int ne; double elo, ehi; const double *eedges;
  const double *values; /* kind 1: [N][nx+2]; kind 2: [(ne+2)][(nx+2)]; kind 3: [ne+2] */
} orc_eff;

/* TAxis::FindBin / FindFixBin (non-extendable axis) */
int orc_find_bin(int nbins, double lo, double hi, const double *edges, double x) {
  if (x < lo) return 0;
  if (!(x < hi)) return nbins + 1;
  if (!edges) return 1 + (int)(nbins * (x - lo) / (hi - lo));
  { /* 1 + TMath::BinarySearch: largest i with edges[i] <= x */
    int l = 0, h = nbins; /* edges has nbins+1 entries */
    while (h - l > 1) { int m = (l + h) / 2; if (edges[m] <= x) l = m; else h = m; }
    return 1 + l;
  }
}
static double eff_eval(const orc_eff *e, double x, double energy, int row) {
  if (!e || e->kind == 0) return 1.0;
  if (e->kind == 1) return e->values[(size_t)row * (e->nx + 2) + orc_find_bin(e->nx, e->xlo, e->xhi, e->xedges, x)];
  if (e->kind == 2) {
    int bx = orc_find_bin(e->nx, e->xlo, e->xhi, e->xedges, x);
    int by = orc_find_bin(e->ne, e->elo, e->ehi, e->eedges, energy);
    return e->values[(size_t)by * (e->nx + 2) + bx];
  }
  return e->values[orc_find_bin(e->ne, e->elo, e->ehi, e->eedges, energy)];
}

/* ------------------------------------------------------------------------------------------------
 * Interpolator                     src/Interpolator.cpp, src/BaseRangeInterpolator.cpp,
 *                                  src/LinearRangeInterpolator.cpp, src/CSplineRangeInterpolator.cpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n;         /* number of cross-section points */
  double *ext;   /* n+1: threshold, E_0..E_{n-1}            src/Interpolator.cpp:37-39 */
  int nr;
  int *rtype, *rbegin, *rnseg; double *rmin, *rmax;
  double **splc; /* per cspline range: [rnseg][n+1] second-derivative coefficients "c" (GSL) */
} orc_interp;

/* GSL cspline_init (natural) + solve_symm_tridiag */
static void cspline_natural_c(int size, const double *xa, const double *ya, double *c) {
  const int max_index = size - 1, sys_size = max_index - 1;
  int i;
  c[0] = 0.0; c[max_index] = 0.0;
  if (sys_size < 1) return;
  {
    double *g = (double *)malloc(sizeof(double) * sys_size), *diag = (double *)malloc(sizeof(double) * sys_size);
    double *offdiag = (double *)malloc(sizeof(double) * sys_size);
    for (i = 0; i < sys_size; i++) {
      const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
      const double ydiff_i = ya[i + 1] - ya[i], ydiff_ip1 = ya[i + 2] - ya[i + 1];
      const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
      offdiag[i] = h_ip1; diag[i] = 2.0 * (h_ip1 + h_i);
      g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
    }
    if (sys_size == 1) { c[1] = g[0] / diag[0]; }
    else {
      const int N = sys_size;
      double *gamma = (double *)malloc(sizeof(double) * N), *alpha = (double *)malloc(sizeof(double) * N);
      double *cc = (double *)malloc(sizeof(double) * N), *z = (double *)malloc(sizeof(double) * N);
      double *x = c + 1;
      alpha[0] = diag[0]; gamma[0] = offdiag[0] / alpha[0];
      for (i = 1; i < N - 1; i++) { alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1]; gamma[i] = offdiag[i] / alpha[i]; }
      if (N > 1) alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
      z[0] = g[0];
      for (i = 1; i < N; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
      for (i = 0; i < N; i++) cc[i] = z[i] / alpha[i];
      x[N - 1] = cc[N - 1];
      if (N >= 2) for (i = N - 2; i >= 0; i--) x[i] = cc[i] - gamma[i] * x[i + 1];
      free(gamma); free(alpha); free(cc); free(z);
    }
    free(g); free(diag); free(offdiag);
  }
}
/* exported for oracle/ref_shim/gsl_shim.c (the GSL-API shim the reference sources are compiled against) */
void orc_cspline_natural_c(int size, const double *xa, const double *ya, double *c) { cspline_natural_c(size, xa, ya, c); }
/* gsl_interp_bsearch(xa, x, 0, size-1) */
int orc_bsearch(const double *xa, double x, int index_lo, int index_hi) {
  int ilo = index_lo, ihi = index_hi;
  while (ihi > ilo + 1) { int i = (ihi + ilo) / 2; if (xa[i] > x) ihi = i; else ilo = i; }
  return ilo;
}
/* gsl cspline_eval / cspline_eval_deriv for the cardinal spline with y = e_{node} */
static double cspline_eval(const double *xa, const double *c, int size, int node, double x, int deriv) {
  int index;
  if (x < xa[0] || x > xa[size - 1]) return NAN; /* GSL_EDOM -> NaN */
  index = orc_bsearch(xa, x, 0, size - 1);
  {
    const double x_hi = xa[index + 1], x_lo = xa[index], dx = x_hi - x_lo;
    if (dx > 0.0) {
      const double y_lo = (index == node) ? 1.0 : 0.0, y_hi = (index + 1 == node) ? 1.0 : 0.0;
      const double dy = y_hi - y_lo, delx = x - x_lo;
      const double c_i = c[index], c_ip1 = c[index + 1];
      const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
      const double d_i = (c_ip1 - c_i) / (3.0 * dx);
      if (deriv) return b_i + delx * (2.0 * c_i + 3.0 * d_i * delx);
      return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
    }
    return 0.0;
  }
}

void orc_interp_free(orc_interp *p) {
  int r;
  if (!p) return;
  if (p->splc) { for (r = 0; r < p->nr; r++) free(p->splc[r]); free(p->splc); }
  free(p->ext); free(p->rtype); free(p->rbegin); free(p->rnseg); free(p->rmin); free(p->rmax); free(p);
}

/* settings: nr x 3 ints (cubic?, firstSegment, lastSegment).  Returns NULL where the reference
 * throws InterpRangeException (src/Interpolator.cpp:52-59, 90-147). */
orc_interp *orc_interp_create(int n, const double *ecm, double threshold, int nr, const int *settings) {
  int r, k, *idx;
  orc_interp *p;
  if (nr < 1) return NULL;
  idx = (int *)malloc(sizeof(int) * nr);
  for (r = 0; r < nr; r++) idx[r] = r;
  /* sort by first segment (src/Interpolator.cpp:44-48) */
  for (r = 1; r < nr; r++) {
    int v = idx[r]; k = r - 1;
    while (k >= 0 && settings[3 * idx[k] + 1] > settings[3 * v + 1]) { idx[k + 1] = idx[k]; k--; }
    idx[k + 1] = v;
  }
  /* checkRangeInterpSettings :90-147 */
  if (settings[3 * idx[0] + 1] != 0 || settings[3 * idx[nr - 1] + 2] + 1 != n) { free(idx); return NULL; }
  { int pimax = -1;
    for (r = 0; r < nr; r++) {
      int imin = settings[3 * idx[r] + 1], imax = settings[3 * idx[r] + 2];
      if (imin > imax) { free(idx); return NULL; }
      if (pimax != -1 && pimax + 1 != imin) { free(idx); return NULL; }
      pimax = imax;
    } }
  p = (orc_interp *)calloc(1, sizeof(orc_interp));
  p->n = n; p->nr = nr;
  p->ext = (double *)malloc(sizeof(double) * (n + 1));
  p->ext[0] = threshold; memcpy(p->ext + 1, ecm, sizeof(double) * n);
  p->rtype = (int *)malloc(sizeof(int) * nr); p->rbegin = (int *)malloc(sizeof(int) * nr);
  p->rnseg = (int *)malloc(sizeof(int) * nr);
  p->rmin = (double *)malloc(sizeof(double) * nr); p->rmax = (double *)malloc(sizeof(double) * nr);
  p->splc = (double **)calloc(nr, sizeof(double *));
  for (r = 0; r < nr; r++) {
    const int lo = settings[3 * idx[r] + 1], hi = settings[3 * idx[r] + 2];
    p->rtype[r] = settings[3 * idx[r]] ? 1 : 0;
    p->rnseg[r] = lo > 0 ? hi - lo + 2 : hi - lo + 1;   /* src/BaseRangeInterpolator.cpp:80-87 */
    p->rbegin[r] = lo > 0 ? lo - 1 : 0;                 /* :94-100 */
    p->rmin[r] = p->ext[lo]; p->rmax[r] = p->ext[hi + 1]; /* :13-19 */
    if (p->rtype[r]) {                                  /* src/CSplineRangeInterpolator.cpp:16-41 */
      double *yi = (double *)calloc(n + 1, sizeof(double));
      p->splc[r] = (double *)malloc(sizeof(double) * (size_t)p->rnseg[r] * (n + 1));
      for (k = 0; k < p->rnseg[r]; k++) {
        yi[p->rbegin[r] + k + 1] = 1.;
        cspline_natural_c(n + 1, p->ext, yi, p->splc[r] + (size_t)k * (n + 1));
        yi[p->rbegin[r] + k + 1] = 0.;
      }
      free(yi);
    }
  }
  free(idx);
  return p;
}
int orc_interp_nranges(const orc_interp *p) { return p->nr; }
void orc_interp_range(const orc_interp *p, int r, int *type, int *begin, int *nseg, double *mn, double *mx) {
  *type = p->rtype[r]; *begin = p->rbegin[r]; *nseg = p->rnseg[r]; *mn = p->rmin[r]; *mx = p->rmax[r];
}
/* raw GSL "c" coefficients of the cardinal spline owned by range r for cross-section index j */
const double *orc_interp_spline_c(const orc_interp *p, int r, int j) {
  return p->splc[r] + (size_t)(j - p->rbegin[r]) * (p->n + 1);
}

static int has_cs_index(const orc_interp *p, int r, int j) {      /* src/BaseRangeInterpolator.cpp:35-38 */
  const int index = j - p->rbegin[r];
  return (index >= 0) && (index < p->rnseg[r]);
}
static int is_energy_in_range(const orc_interp *p, int r, double e) { /* :45-47 */
  return (e > p->rmin[r]) && (e <= p->rmax[r]);
}
/* src/LinearRangeInterpolator.cpp:135-150 */
static double c00(const double *x, int j) { return x[j] / (x[j] - x[j + 1]); }
static double c01(const double *x, int j) { return 1. / (x[j + 1] - x[j]); }
static double c10(const double *x, int j) { return 1. + x[j + 1] / (x[j + 2] - x[j + 1]); }
static double c11(const double *x, int j) { return 1. / (x[j + 1] - x[j + 2]); }

static double range_basis_eval(const orc_interp *p, int r, int j, double e, int deriv) {
  const double *x = p->ext;
  if (p->rtype[r]) /* src/CSplineRangeInterpolator.cpp:64-68, 76-80 */
    return cspline_eval(x, orc_interp_spline_c(p, r, j), p->n + 1, j + 1, e, deriv);
  /* src/LinearRangeInterpolator.cpp:25-55 */
  if (e <= x[j]) return 0;
  if (e <= x[j + 1]) return deriv ? c01(x, j) : c01(x, j) * e + c00(x, j);
  if (j + 2 == p->n + 1) return deriv ? 0 : 1;
  if (e <= x[j + 2]) return deriv ? c11(x, j) : c11(x, j) * e + c10(x, j);
  return 0;
}
/* src/Interpolator.cpp:217-258 */
double orc_basis_eval(const orc_interp *p, int j, double e) {
  const double emin = p->rmin[0], emax = p->rmax[p->nr - 1];
  double result = 0; int r;
  if (e <= emin) return 0;
  if (e > emax) {
    if (has_cs_index(p, p->nr - 1, j)) return range_basis_eval(p, p->nr - 1, j, emax, 0);
    return 0;
  }
  for (r = 0; r < p->nr; r++)
    if (has_cs_index(p, r, j) && is_energy_in_range(p, r, e)) result += range_basis_eval(p, r, j, e, 0);
  return result;
}
/* src/Interpolator.cpp:266-287 */
double orc_basis_deriv_eval(const orc_interp *p, int j, double e) {
  const double emin = p->rmin[0], emax = p->rmax[p->nr - 1];
  double result = 0; int r;
  if (e <= emin || e > emax) return 0;
  for (r = 0; r < p->nr; r++)
    if (has_cs_index(p, r, j) && is_energy_in_range(p, r, e)) result += range_basis_eval(p, r, j, e, 1);
  return result;
}
/* src/Interpolator.cpp:295-323 with src/BaseRangeInterpolator.cpp:108-118 */
double orc_interp_eval(const orc_interp *p, const double *y, double e) {
  const double emin = p->rmin[0], emax = p->rmax[p->nr - 1];
  double result = 0; int r, k;
  if (e <= emin) return 0;
  if (e > emax) {
    r = p->nr - 1;
    for (k = 0; k < p->rnseg[r]; k++) result += range_basis_eval(p, r, p->rbegin[r] + k, emax, 0) * y[p->rbegin[r] + k];
    return result;
  }
  for (r = 0; r < p->nr; r++)
    if (is_energy_in_range(p, r, e))
      for (k = 0; k < p->rnseg[r]; k++) result += range_basis_eval(p, r, p->rbegin[r] + k, e, 0) * y[p->rbegin[r] + k];
  return result;
}

/* ---- Kuraev-Fadin convolution                                 src/KuraevFadin.cpp:51-92 -------- */
typedef struct {
  const orc_interp *p; const orc_eff *eff;
  double energy; int row;
  int mode;   /* 0: fcn = 1; 1: fcn = E'; 2: cspline cardinal of range r, index j; 3: user table fcn */
  int r, j;
  orc_fn user; void *user_ctx;
  const double *brk; int nbrk; /* "converged" back-end: ascending interior break points (NULL = faithful) */
} conv_ctx;

static double conv_integrand(double x, void *vctx) {
  const conv_ctx *c = (const conv_ctx *)vctx;
  const double en = c->energy * sqrt(1 - x);           /* :53 */
  double fv;
  if (c->mode == 0) fv = 1.;
  else if (c->mode == 1) fv = en;
  else if (c->mode == 2) fv = cspline_eval(c->p->ext, orc_interp_spline_c(c->p, c->r, c->j), c->p->n + 1, c->j + 1, en, 0);
  else fv = c->user(en, c->user_ctx);
  return fv * orc_kernel_kf(x, c->energy * c->energy) * eff_eval(c->eff, x, c->energy, c->row); /* :53,61 */
}
static double convolution_kf_faithful(conv_ctx *c, double min_x, double max_x);
static double convolution_piece(conv_ctx *c, double a, double b);
static double conv_integrand(double x, void *vctx);
/* "Converged" back-end (NOT what the reference does): the same integrand and the same QUADPACK
 * routines, but the range is cut at every point where the integrand is not smooth (spline knots,
 * efficiency bin edges, x0), so that every adaptive call converges at its first tolerance instead of
 * ending in the reference's loosen-until-success loop.  This is the mathematically exact value of the
 * integral the reference approximates; tests gate the CUDA path against it at 1e-9 and against the
 * faithful back-end within the latter's own error estimate. */
static double convolution_kf(conv_ctx *c, double min_x, double max_x) {
  double result = 0, lo = min_x; int q;
  if (!c->brk) return convolution_kf_faithful(c, min_x, max_x);
  if (!(max_x > min_x)) return 0.;
  for (q = 0; q < c->nbrk; q++) {
    const double v = c->brk[q];
    if (!(v > lo)) continue;
    if (!(v < max_x)) break;
    result += convolution_piece(c, lo, v);
    lo = v;
  }
  result += convolution_piece(c, lo, max_x);
  return result;
}
/* one smooth piece: x0 is a break point, so the piece lies on one side of it.  (The reference's
 * "[min,x0] forwards + [x0,max] backwards" evaluation for max < x0, src/KuraevFadin.cpp:64-68, is NOT
 * reproduced here: it would re-cross the break points the piece was cut at.) */
static double convolution_piece(conv_ctx *c, double a, double b) {
  double error;
  const double x0 = 4 * ELECTRON_M / c->energy;
  if (b <= x0) return orc_integrateS(conv_integrand, c, a, b, &error);
  return orc_integrate(conv_integrand, c, a, b, &error);
}
static double convolution_kf_faithful(conv_ctx *c, double min_x, double max_x) {   /* :56-73 */
  double error, result;
  const double x0 = 4 * ELECTRON_M / c->energy;
  if (min_x < x0) {
    result = orc_integrateS(conv_integrand, c, min_x, x0, &error) + orc_integrate(conv_integrand, c, x0, max_x, &error);
  } else {
    result = orc_integrate(conv_integrand, c, min_x, max_x, &error);
  }
  return result;
}
/* public: convolution of a user function of energy (C callback) -- used to generate sigma_vis
 * (src/isrsolver-KuraevFadin-convolution.cpp:122-134) */
double orc_convolution_kf(double energy, orc_fn fcn, void *fctx, double min_x, double max_x, const orc_eff *eff, int row) {
  conv_ctx c; memset(&c, 0, sizeof(c));
  c.energy = energy; c.eff = eff; c.row = row; c.mode = 3; c.user = fcn; c.user_ctx = fctx;
  return convolution_kf(&c, min_x, max_x);
}

/* break points of row i for the converged back-end: x0, knots 1-(ext[m]/E)^2 (m <= i), eps x-edges */
static _Thread_local const double *g_brk = NULL;
static _Thread_local int g_nbrk = 0;
static int cmp_double(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}
static double *row_breaks(const orc_interp *p, const orc_eff *eff, int i, int *count) {
  const double en = p->ext[i + 1];
  int cap = i + 4 + ((eff && (eff->kind == 1 || eff->kind == 2)) ? eff->nx + 2 : 0), n = 0, m;
  double *b = (double *)malloc(sizeof(double) * cap);
  b[n++] = 4 * ELECTRON_M / en;
  for (m = 0; m <= i; m++) b[n++] = 1 - pow(p->ext[m] / en, 2);
  if (eff && (eff->kind == 1 || eff->kind == 2))
    for (m = 0; m <= eff->nx; m++)
      b[n++] = eff->xedges ? eff->xedges[m] : (m == eff->nx ? eff->xhi : eff->xlo + m * ((eff->xhi - eff->xlo) / eff->nx));
  qsort(b, n, sizeof(double), cmp_double);
  *count = n;
  return b;
}

/* src/LinearRangeInterpolator.cpp:57-101 */
static double linear_kf_integral(const orc_interp *p, int r, int i, int j, const orc_eff *eff) {
  const double *x = p->ext;
  double result = 0;
  conv_ctx c; memset(&c, 0, sizeof(c));
  c.p = p; c.eff = eff; c.row = i; c.brk = g_brk; c.nbrk = g_nbrk;
  if (j > i) return 0;
  { /* first triangle :70-83 */
    const double enc = x[j + 1];
    if (!(enc > p->rmax[r] || enc <= p->rmin[r])) {
      const double en = x[i + 1];
      const double x0 = fmax(0., 1 - pow(enc / en, 2));
      const double x1 = 1 - pow(x[j] / en, 2);
      double i00, i01;
      c.energy = en;
      g_err_weight = fabs(c00(x, j)); c.mode = 0; i00 = convolution_kf(&c, x0, x1);
      g_err_weight = fabs(c01(x, j)); c.mode = 1; i01 = convolution_kf(&c, x0, x1);
      g_err_weight = 1;
      result = c01(x, j) * i01 + c00(x, j) * i00;
    }
  }
  { /* second triangle :85-101 */
    double second = 0;
    if (!(j + 2 >= p->n + 1)) {
      const double encp = x[j + 2];
      if (!(encp > p->rmax[r] || encp <= p->rmin[r])) {
        const double en = x[i + 1];
        const double x0 = fmax(0., 1 - pow(encp / en, 2));
        const double x1 = 1 - pow(x[j + 1] / en, 2);
        double i10, i11;
        c.energy = en;
        g_err_weight = fabs(c10(x, j)); c.mode = 0; i10 = convolution_kf(&c, x0, x1);
        g_err_weight = fabs(c11(x, j)); c.mode = 1; i11 = convolution_kf(&c, x0, x1);
        g_err_weight = 1;
        second = c11(x, j) * i11 + c10(x, j) * i10;
      }
    }
    result += second;
  }
  return result;
}
/* src/CSplineRangeInterpolator.cpp:88-110 */
static double cspline_kf_integral(const orc_interp *p, int r, int i, int j, const orc_eff *eff) {
  const double en = p->ext[i + 1];
  conv_ctx c;
  double x_min, x_max;
  if (en <= p->rmin[r]) return 0;
  memset(&c, 0, sizeof(c));
  c.p = p; c.eff = eff; c.row = i; c.energy = en; c.mode = 2; c.r = r; c.j = j; c.brk = g_brk; c.nbrk = g_nbrk;
  x_min = fmax(0., 1 - pow(p->rmax[r] / en, 2));
  x_max = 1 - pow(p->rmin[r] / en, 2);
  return convolution_kf(&c, x_min, x_max);
}
/* src/Interpolator.cpp:437-453 */
double orc_kf_basis_integral(const orc_interp *p, int i, int j, const orc_eff *eff) {
  double result = 0; int r;
  for (r = 0; r < p->nr; r++)
    if (has_cs_index(p, r, j))
      result += p->rtype[r] ? cspline_kf_integral(p, r, i, j, eff) : linear_kf_integral(p, r, i, j, eff);
  return result;
}
/* mode 0: faithful (the reference's call structure); mode 1: converged (see convolution_kf).
 * abserr: sum of the QUADPACK error estimates of every adaptive call behind the entry, each weighted
 * by the |coefficient| (c00, c01, c10, c11 of the linear basis) its integral is multiplied with;
 * relerr_used: the loosest epsrel any retry loop ended with (1e-12 = converged at first try). */
double orc_kf_basis_integral_ex(const orc_interp *p, int i, int j, const orc_eff *eff, int mode,
                                double *abserr, double *relerr_used) {
  double result, *b = NULL; int nb = 0;
  g_abserr_sum = 0; g_max_relerr = 0;
  if (mode == 1) { b = row_breaks(p, eff, i, &nb); g_brk = b; g_nbrk = nb; }
  result = orc_kf_basis_integral(p, i, j, eff);
  g_brk = NULL; g_nbrk = 0; free(b);
  if (abserr) *abserr = g_abserr_sum;
  if (relerr_used) *relerr_used = g_max_relerr;
  return result;
}

/* ---- integral of a basis function (Tikhonov weights)  src/Interpolator.cpp:459-473 ------------ */
typedef struct { const orc_interp *p; int mode, r, j; } basis_ctx;
static double basis_integrand(double e, void *v) {
  const basis_ctx *b = (const basis_ctx *)v;
  if (b->mode == 0) return 1.;
  if (b->mode == 1) return e;
  return cspline_eval(b->p->ext, orc_interp_spline_c(b->p, b->r, b->j), b->p->n + 1, b->j + 1, e, 0);
}
double orc_integral_basis(const orc_interp *p, int j) {
  double result = 0, error; int r;
  const double *x = p->ext;
  for (r = 0; r < p->nr; r++) {
    basis_ctx b; b.p = p; b.r = r; b.j = j;
    if (!has_cs_index(p, r, j)) continue;
    if (p->rtype[r]) {                /* src/CSplineRangeInterpolator.cpp:131-143 */
      b.mode = 2;
      result += orc_integrate(basis_integrand, &b, p->rmin[r], p->rmax[r], &error);
    } else {                          /* src/LinearRangeInterpolator.cpp:103-133 */
      const double enc = x[j + 1];
      double first = 0, second = 0;
      if (!(enc > p->rmax[r] || enc <= p->rmin[r])) {
        double i00, i01;
        b.mode = 0; i00 = orc_integrate(basis_integrand, &b, x[j], enc, &error);
        b.mode = 1; i01 = orc_integrate(basis_integrand, &b, x[j], enc, &error);
        first = c01(x, j) * i01 + c00(x, j) * i00;
      }
      if (!(j + 2 >= p->n + 1)) {
        const double encp = x[j + 2];
        if (!(encp > p->rmax[r] || encp <= p->rmin[r])) {
          double i10, i11;
          b.mode = 0; i10 = orc_integrate(basis_integrand, &b, enc, encp, &error);
          b.mode = 1; i11 = orc_integrate(basis_integrand, &b, enc, encp, &error);
          second = c11(x, j) * i11 + c10(x, j) * i10;
        }
      }
      result += first + second;
    }
  }
  return result;
}


/* ---- tiny pthread parallel-for (the image's default gcc has no OpenMP runtime) ----------------- */
#include <pthread.h>
typedef void (*pf_body)(long idx, void *arg);
typedef struct { pf_body body; void *arg; long n; long *next; pthread_mutex_t *mu; } pf_task;
static void *pf_worker(void *v) {
  pf_task *t = (pf_task *)v;
  for (;;) {
    long idx;
    pthread_mutex_lock(t->mu); idx = (*t->next)++; pthread_mutex_unlock(t->mu);
    if (idx >= t->n) break;
    t->body(idx, t->arg);
  }
  return NULL;
}
static void parallel_for(long n, int nthreads, pf_body body, void *arg) {
  long next = 0; int k;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pf_task t; pthread_t th[256];
  t.body = body; t.arg = arg; t.n = n; t.next = &next; t.mu = &mu;
  if (nthreads <= 1) { long i; for (i = 0; i < n; i++) body(i, arg); return; }
  if (nthreads > 256) nthreads = 256;
  for (k = 0; k < nthreads; k++) pthread_create(&th[k], NULL, pf_worker, &t);
  for (k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
}

/* ---- equation matrix                      src/ISRSolverSLE.cpp:138-149, 177-193 ----------------
 * G is written COLUMN-MAJOR (Eigen::MatrixXd storage): G[i + n*j].  ecm_err == NULL => no spread.
 * nthreads > 1 parallelises over entries (each entry is independent; results are identical). */
typedef struct { const orc_interp *p; int j; } spread_ctx;
static double spread_integrand(double e, void *v) {
  const spread_ctx *s = (const spread_ctx *)v;
  return orc_basis_eval(s->p, s->j, e);
}
void orc_energy_spread_matrix(const orc_interp *p, const double *ecm_err, double *S /* col-major */) {
  const int n = p->n; int i, j;
  for (i = 0; i < n; i++)
    for (j = 0; j < n; j++) {
      spread_ctx s; const double sigma2 = pow(ecm_err[i], 2);
      s.p = p; s.j = j;
      S[i + (size_t)n * j] = orc_gaussian_conv(p->ext[i + 1], sigma2, spread_integrand, &s);
    }
}
typedef struct { const orc_interp *p; const orc_eff *eff; double *G, *Err, *Rel; int mode; } eqm_arg;
static void eqm_body(long idx, void *v) {
  eqm_arg *a = (eqm_arg *)v; const int n = a->p->n;
  const int j = (int)(idx / n), i = (int)(idx % n);
  double e, r;
  a->G[i + (size_t)n * j] = orc_kf_basis_integral_ex(a->p, i, j, a->eff, a->mode, &e, &r);
  if (a->Err) a->Err[i + (size_t)n * j] = e;
  if (a->Rel) a->Rel[i + (size_t)n * j] = r;
}
void orc_eval_eq_matrix_ex(const orc_interp *p, const orc_eff *eff, const double *ecm_err, double *G,
                           int nthreads, int mode, double *Err, double *Rel);
void orc_eval_eq_matrix(const orc_interp *p, const orc_eff *eff, const double *ecm_err, double *G, int nthreads) {
  orc_eval_eq_matrix_ex(p, eff, ecm_err, G, nthreads, 0, NULL, NULL);
}
/* Err / Rel (optional, column-major like G) describe the un-spread matrix entries. */
void orc_eval_eq_matrix_ex(const orc_interp *p, const orc_eff *eff, const double *ecm_err, double *G,
                           int nthreads, int mode, double *Err, double *Rel) {
  const int n = p->n;
  eqm_arg a; a.p = p; a.eff = eff; a.G = G; a.Err = Err; a.Rel = Rel; a.mode = mode;
  parallel_for((long)n * n, nthreads, eqm_body, &a);
  if (ecm_err) {
    double *S = (double *)malloc(sizeof(double) * n * n), *T = (double *)calloc((size_t)n * n, sizeof(double));
    int i, j, k;
    orc_energy_spread_matrix(p, ecm_err, S);
    for (j = 0; j < n; j++)
      for (k = 0; k < n; k++) {
        const double g = G[k + (size_t)n * j];
        for (i = 0; i < n; i++) T[i + (size_t)n * j] += S[i + (size_t)n * k] * g;
      }
    memcpy(G, T, sizeof(double) * n * n);
    free(S); free(T);
  }
}
/* src/ISRSolverTikhonov.cpp:92-99 : D(i,j) = basis_j'(E_i), column-major */
void orc_deriv_projector(const orc_interp *p, double *D) {
  const int n = p->n; int i, j;
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) D[i + (size_t)n * j] = orc_basis_deriv_eval(p, j, p->ext[i + 1]);
}

/* ---- forward convolution of a Breit-Wigner Born model (bench/test input generator) ------------
 * mirrors src/isrsolver-KuraevFadin-convolution.cpp:122-134: sigma_vis(E) = conv over [0, 1-(thr/E)^2]. */
typedef struct { double m, g, s0, thr; } bw_par;
static double bw_born(double e, void *v) {
  const bw_par *q = (const bw_par *)v;
  double e2, d, ps;
  if (e <= q->thr) return 0.;
  e2 = e * e; d = (e2 - q->m * q->m); ps = 1. - q->thr * q->thr / e2;
  return q->s0 * q->m * q->m * q->g * q->g / (d * d + q->m * q->m * q->g * q->g) * ps * sqrt(ps);
}
double orc_bw_born(double e, double m, double g, double s0, double thr) {
  bw_par q; q.m = m; q.g = g; q.s0 = s0; q.thr = thr; return bw_born(e, &q);
}
typedef struct { const double *ecm; bw_par q; const orc_eff *eff; double *vcs; } bwv_arg;
static void bwv_body(long i, void *v) {
  bwv_arg *a = (bwv_arg *)v;
  a->vcs[i] = orc_convolution_kf(a->ecm[i], bw_born, &a->q, 0., 1 - pow(a->q.thr / a->ecm[i], 2), a->eff, (int)i);
}
void orc_bw_visible(int n, const double *ecm, double thr, double m, double g, double s0,
                    const orc_eff *eff, double *vcs, int nthreads) {
  bwv_arg a; a.ecm = ecm; a.q.m = m; a.q.g = g; a.q.s0 = s0; a.q.thr = thr; a.eff = eff; a.vcs = vcs;
  parallel_for(n, nthreads, bwv_body, &a);
}

/* ---- test hooks ------------------------------------------------------------------------------- */
/* integrand families for pinning the QUADPACK restatement against SciPy (tests/test_oracle.py) */
typedef struct { int kind; double p0, p1; } probe_ctx;
static double probe_fn(double x, void *v) {
  const probe_ctx *c = (const probe_ctx *)v;
  switch (c->kind) {
    case 0: return orc_kernel_kf(x, c->p0);                     /* F(x, s) */
    case 1: return c->p1 * sqrt(1 - x) * orc_kernel_kf(x, c->p0); /* E' F  */
    case 2: return pow(x, c->p0) * log(1 / x);                  /* QUADPACK book test family */
    case 3: return cos(c->p0 * x) * exp(-x * c->p1);
    default: return 0;
  }
}
int orc_probe_qags(int kind, double p0, double p1, double a, double b, double epsabs, double epsrel,
                   double *result, double *abserr, long *neval) {
  probe_ctx c; int st; long n0 = g_neval;
  c.kind = kind; c.p0 = p0; c.p1 = p1;
  st = orc_qags(probe_fn, &c, a, b, epsabs, epsrel, ORC_LIMIT_S, result, abserr);
  *neval = g_neval - n0;
  return st;
}
int orc_probe_qag61(int kind, double p0, double p1, double a, double b, double epsabs, double epsrel,
                    double *result, double *abserr, long *neval) {
  probe_ctx c; int st; long n0 = g_neval;
  c.kind = kind; c.p0 = p0; c.p1 = p1;
  st = orc_qag61(probe_fn, &c, a, b, epsabs, epsrel, ORC_LIMIT, result, abserr);
  *neval = g_neval - n0;
  return st;
}
