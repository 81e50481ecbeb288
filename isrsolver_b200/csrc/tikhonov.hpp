#pragma once
#include "common.cuh"
struct isr_problem;
namespace isr {
int tikhonov_prepare(isr_problem* p, const double* vcs_err, int deriv_reg);
int tikhonov_scan(isr_problem* p, int64_t L, const double* lambda, const double* vcs, double* eqn, double* smooth,
                  double* curv, double* dcurv, double* bcs_out);
int tikhonov_solve(isr_problem* p, double la, const double* vcs, double* bcs_out, double* cov_out);
int tikhonov_select(isr_problem* p, double la);
int tikhonov_toy_scan_device(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V_dev,
                             const double* bcs0, double* chi2_dev);
int tikhonov_toy_scan_hist(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V, const double* bcs0,
                           int nbins, double* stats_out, uint32_t* counts_out);
int tikhonov_chi2_scan(isr_problem* p, const isr_tikhonov_scan_args& a);
int tsvd_solve(isr_problem* p, int k, int keep_one, const double* vcs, const double* vcs_err, double* bcs_out, double* cov_out);
int tsvd_select(isr_problem* p, int k, int keep_one, const double* vcs_err);
// E (tc x ld) = V Pe^T - shift : the tile kernel with an arbitrary operator (toy_batch.cu)
int project_device(isr_problem* p, int64_t T, const double* V_dev, const double* op_rowmajor, const double* shift_dev, double* E_dev);
}  // namespace isr
