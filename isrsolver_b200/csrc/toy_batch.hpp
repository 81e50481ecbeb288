#pragma once
#include "common.cuh"
struct isr_problem;
namespace isr {
struct ToyOut {
  double* chi2 = nullptr;         // T
  double* sum_bcs = nullptr;      // N, added to
  double* ratio_stats = nullptr;  // 3N {count, sum r, sum r^2}, added to
  uint32_t* rhist = nullptr;      // N x 1026, added to
  double* bcs = nullptr;          // T x N
};
// All pointers are device pointers; launches go to the problem's context stream.
int toy_batch_device(isr_problem* p, int64_t T, const double* V_dev, const double* b0_dev, const ToyOut& out);
const char* toy_cfg_name(int n, int nstat);   // nstat: 0 none, 1 sum of solutions, 4 ratio statistics; nullptr: no tile fits
double* scratch_dfull(isr_problem* p, int64_t T);
}  // namespace isr
