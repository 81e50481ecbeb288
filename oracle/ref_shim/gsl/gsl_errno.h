/* Minimal stand-in for <gsl/gsl_errno.h>: ORACLE / TEST INFRASTRUCTURE (see ../README.md).
 * GSL is absent from this image; the reference's src/Integration.cpp only switches the error handler
 * off and back on (src/Integration.cpp:22,45,56,79).  The shim never invokes a handler. */
#ifndef ISR_REF_SHIM_GSL_ERRNO_H
#define ISR_REF_SHIM_GSL_ERRNO_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void gsl_error_handler_t(const char* reason, const char* file, int line, int gsl_errno);
gsl_error_handler_t* gsl_set_error_handler_off(void);
gsl_error_handler_t* gsl_set_error_handler(gsl_error_handler_t* new_handler);
#ifdef __cplusplus
}
#endif
#endif
