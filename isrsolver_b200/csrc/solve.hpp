#pragma once
#include "common.cuh"
#include "toy_batch.hpp"
struct isr_problem;
namespace isr {
int factor(isr_problem* p, const double* vcs_err);
int factor_enqueue(isr_problem* p, const double* vcs_err);
int factor_finish(isr_problem* p, bool* redone);
int select_sle(isr_problem* p);
int launch_sle_operators(isr_problem* p);
int ensure_cov(isr_problem* p);
int iterative_solve(isr_problem* p, int niter, const double* vcs, double* bcs_out, double* radcorr_out);
int minmax_device(isr_ctx* c, int64_t n, const double* v_dev, double* lo, double* hi);
int histogram_device(isr_ctx* c, int64_t n, const double* v_dev, int nbins, double lo, double hi, uint32_t* counts_host);
int minmax_to_device(isr_ctx* c, int64_t n, const double* v_dev, double* out_dev);
int histogram_range_device(isr_ctx* c, int64_t n, const double* v_dev, int nbins, const double* range_dev, int* counts_dev);
int toy_batch_host(isr_problem* p, int64_t T, const double* V, const double* bcs0, double* chi2_out,
                   double* sum_bcs_out, double* ratio_stats_out, uint32_t* ratio_hist_out, double* bcs_out);
int host_pipe_begin(isr_problem* p, int64_t T, const double* V);
int host_pipe_run(isr_problem* p, int64_t T, const double* V, const double* b0_dev, const ToyOut& base, double* chi2_out,
                  double* bcs_out);
int toy_campaign_device(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* vcs,
                        const double* vcs_err, const double* bcs0, double* chi2_dev, double* sum_bcs_dev,
                        double* ratio_stats_dev, uint32_t* rhist_dev);
int draw_toys_device(isr_ctx* c, int64_t T, int n, int64_t first_toy, uint64_t seed, const double* vcs_dev,
                     const double* err_dev, double* V_dev, cudaStream_t st = nullptr);   // nullptr: the context stream
int64_t gen_chunk(int n);   // toys per generator block
int drawn_toys_run(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* b0_dev, const ToyOut& o, bool block0_drawn);
struct TestWants { bool sum = false, rhist = false, range = false; };
int chi2_test(isr_problem* p, const isr_chi2_test_args& a, TestWants w);
}  // namespace isr
