/* Minimal stand-in for <gsl/gsl_integration.h>: ORACLE / TEST INFRASTRUCTURE (see ../README.md).
 * Declares exactly the GSL entry points src/Integration.cpp calls (:23-44, :57-78, :90-97).  They are
 * implemented in ../gsl_shim.c on top of the QUADPACK restatement in oracle/isr_oracle.c. */
#ifndef ISR_REF_SHIM_GSL_INTEGRATION_H
#define ISR_REF_SHIM_GSL_INTEGRATION_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  double (*function)(double x, void* params);
  void* params;
} gsl_function;

typedef struct { size_t limit; } gsl_integration_workspace;
gsl_integration_workspace* gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace* w);
int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace* workspace, double* result, double* abserr);
int gsl_integration_qag(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                        int key, gsl_integration_workspace* workspace, double* result, double* abserr);

typedef struct { int id; } gsl_integration_fixed_type;
extern const gsl_integration_fixed_type* gsl_integration_fixed_hermite;
typedef struct { size_t n; double* x; double* weights; } gsl_integration_fixed_workspace;
gsl_integration_fixed_workspace* gsl_integration_fixed_alloc(const gsl_integration_fixed_type* type, const size_t n,
                                                             const double a, const double b, const double alpha,
                                                             const double beta);
void gsl_integration_fixed_free(gsl_integration_fixed_workspace* w);
int gsl_integration_fixed(const gsl_function* func, double* result, const gsl_integration_fixed_workspace* w);
#ifdef __cplusplus
}
#endif
#endif
