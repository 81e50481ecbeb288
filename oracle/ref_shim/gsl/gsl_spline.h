/* Minimal stand-in for <gsl/gsl_spline.h>: ORACLE / TEST INFRASTRUCTURE (see ../README.md).
 * Declares the GSL entry points src/CSplineRangeInterpolator{,2}.cpp call (:28-38, :67, :79, :101).
 * Implemented in ../gsl_shim.c: natural cubic spline (gsl_interp_cspline) via the tridiagonal
 * recurrence restated in oracle/isr_oracle.c. */
#ifndef ISR_REF_SHIM_GSL_SPLINE_H
#define ISR_REF_SHIM_GSL_SPLINE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t cache, miss_count, hit_count; } gsl_interp_accel;
gsl_interp_accel* gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel* a);

typedef struct { int id; } gsl_interp_type;
extern const gsl_interp_type* gsl_interp_cspline;
typedef struct { size_t size; double* x; double* y; double* c; } gsl_spline;
gsl_spline* gsl_spline_alloc(const gsl_interp_type* T, size_t size);
int gsl_spline_init(gsl_spline* spline, const double xa[], const double ya[], size_t size);
double gsl_spline_eval(const gsl_spline* spline, double x, gsl_interp_accel* a);
double gsl_spline_eval_deriv(const gsl_spline* spline, double x, gsl_interp_accel* a);
void gsl_spline_free(gsl_spline* spline);
#ifdef __cplusplus
}
#endif
#endif
