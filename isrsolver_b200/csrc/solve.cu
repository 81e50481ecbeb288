// ISRSolverSLE::solve pieces that run once per problem (src/ISRSolverSLE.cpp:86-95), the chi2 / ratio
// histogramming of src/Chi2Test.cpp:64-78, and the host-buffer toy-batch pipeline.
#include "linalg.hpp"
#include "panel_device.cuh"
#include "problem.hpp"
#include "solve.hpp"
#include "toy_batch.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace isr {

__global__ void rop_kernel(int n, int ld, const double* __restrict__ G, const double* __restrict__ err, double* __restrict__ R,
                           double* __restrict__ Rop);
__global__ void sop_kernel(int n, int ld, const double* __restrict__ Ginv, double* __restrict__ Sop);

// Factor for a given error vector (W = diag(err^-2), src/BaseISRSolver.cpp:355-357):
//   Ginv = G^-1                      (b = Ginv v  ==  COD(G).solve(v) for a full-rank G)
//   R    = W^(1/2) G                 (Cov^-1 = G^T W G = R^T R)
//   Cov  = (Ginv W^(-1/2)) (Ginv W^(-1/2))^T  ==  (G^T W G)^-1, symmetric by construction
// factor_enqueue only ENQUEUES (inverse, certificate, operators) on the context stream; the certificate of the
// unpivoted inverse is read by factor_finish after the caller's next synchronisation, so that everything downstream --
// the toy kernels, the histogram, the collectives -- can be enqueued behind the factorisation without the GPU idling
// through a host wake-up (25-40 us per chi2 test in round 1).  Until then the toy schedule is chosen from what the
// host knows: a matrix known to be lower triangular (G_lower) gives exactly lower-triangular G^-1 and W^(1/2) G.
int factor_enqueue(isr_problem* p, const double* vcs_err) {
  if (!p->has_G) { set_error("isr_factor: no matrix (call isr_assemble or isr_set_matrix first)"); return ISR_ESTATE; }
  const int n = p->n, ld = (n + 1) & ~1;
  for (int i = 0; i < n; ++i)
    if (!(vcs_err[i] > 0.0) || !std::isfinite(vcs_err[i])) {
      set_error("isr_factor: vcs_err[%d] = %g must be positive and finite", i, vcs_err[i]);
      return ISR_EINVAL;
    }
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  const size_t nn = (size_t)n * n;
  ISR_TRY(p->d_werr.reserve(n));
  ISR_TRY(p->d_tmp.reserve(nn));
  ISR_TRY(p->d_Ginv.reserve(nn));
  ISR_TRY(p->d_R.reserve(nn));
  ISR_TRY(p->d_Cov.reserve(nn));
  ISR_TRY(p->d_InvCov.reserve(nn));
  ISR_TRY(p->d_Sop.reserve((size_t)n * ld));
  ISR_TRY(p->d_Rop.reserve((size_t)n * ld));
  // the Tikhonov eigen-system (d_mu, d_P, d_U) was built for the previous weights
  if (p->vcs_err.size() != (size_t)n || std::memcmp(p->vcs_err.data(), vcs_err, sizeof(double) * n) != 0) p->tik_ready = false;
  p->vcs_err.assign(vcs_err, vcs_err + n);
  ISR_CUDA(cudaMemcpyAsync(p->d_werr.p, p->vcs_err.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_TRY(p->d_cert_prod.reserve(nn * small_gemm_slices(c, n)));
  ISR_TRY(p->d_cert.reserve(4));
  if (!p->pin_cert) ISR_CUDA(cudaMallocHost((void**)&p->pin_cert, sizeof(double) * 8));
  // R = W^(1/2) G and its row-major copy do not depend on the inverse: side stream, under the inverse
  ISR_CUDA(cudaEventRecord(c->ev_fork2, st));
  ISR_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork2, 0));
  rop_kernel<<<(n * ld + 255) / 256, 256, 0, c->aux_stream>>>(n, ld, p->d_G.p, p->d_werr.p, p->d_R.p, p->d_Rop.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaEventRecord(c->ev_rop, c->aux_stream));
  bool sop_written = false;
  ISR_TRY(inverse_enqueue(c, n, p->d_G.p, p->d_tmp.p, p->d_cert_prod.p, p->d_Ginv.p, p->d_work, p->d_ipiv, p->d_info, p->G_lower,
                          p->pin_cert, p->d_cert.p, p->d_Sop.p, ld, &sop_written));
  p->cert_dev = p->d_cert.p;
  if (!sop_written) { sop_kernel<<<(n * ld + 255) / 256, 256, 0, st>>>(n, ld, p->d_Ginv.p, p->d_Sop.p); ISR_COUNT_LAUNCH(); }
  ISR_CUDA(cudaStreamWaitEvent(st, c->ev_rop, 0));
  ISR_CUDA(cudaGetLastError());
  p->factored = true;
  p->cert_pending = true;
  p->cov_ready = false;      // Cov / G^T W G are formed on first use (ensure_cov): the toy path never reads them
  p->tri_sle[0] = p->tri_sle[1] = p->G_lower ? 1 : 0;
  p->tri_sel[0] = p->tri_sle[0];
  p->tri_sel[1] = p->tri_sle[1];
  p->solver_selected = true;
  p->selected_kind = 1;
  return ISR_OK;
}

// To be called after the context stream has been synchronised.  *redone = true: the unpivoted inverse failed its
// certificate (or the triangularity the schedule assumed did not hold) and the factorisation was redone with partial
// pivoting -- results computed from the speculative operators must be recomputed.
int factor_finish(isr_problem* p, bool* redone) {
  if (redone) *redone = false;
  if (!p->cert_pending) return ISR_OK;
  p->cert_pending = false;
  isr_ctx* c = p->ctx;
  int tri[2];
  const bool ok = inverse_certified(p->pin_cert, p->n, tri);
  const bool assumed = p->tri_sle[0] || p->tri_sle[1];
  if (ok && (!assumed || (tri[0] && tri[1]))) {
    // (exactly triangular factors the host did not know about -- an injected matrix -- switch the schedule from now on)
    p->tri_sle[0] = tri[0]; p->tri_sle[1] = tri[1];
    if (p->selected_kind == 1) { p->tri_sel[0] = tri[0]; p->tri_sel[1] = tri[1]; }
    return ISR_OK;
  }
  p->factored = false;
  ISR_TRY(inverse_pivoted(c, p->n, p->d_G.p, p->d_tmp.p, p->d_Ginv.p, p->d_work, p->d_ipiv, p->d_info, p->pin_cert, p->d_cert.p, tri));
  p->tri_sle[0] = tri[0]; p->tri_sle[1] = tri[1];
  p->tri_sel[0] = tri[0]; p->tri_sel[1] = tri[1];
  ISR_TRY(launch_sle_operators(p));
  ISR_CUDA(cudaStreamSynchronize(c->stream));
  p->factored = true;
  if (redone) *redone = true;
  return ISR_OK;
}

// Synchronous form (isr_factor): enqueue, wait, certify.
int factor(isr_problem* p, const double* vcs_err) {
  ISR_TRY(factor_enqueue(p, vcs_err));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  ISR_CUDA(cudaEventSynchronize(p->ctx->ev_cert));     // the certificate ran on the side stream
  return factor_finish(p, nullptr);
}

// Cov = (Ginv W^(-1/2)) (Ginv W^(-1/2))^T and InvCov = R^T R, once per factorisation and only when asked for.
int ensure_cov(isr_problem* p) {
  if (!p->factored) { set_error("covariance: isr_factor has not been called"); return ISR_ESTATE; }
  if (p->cov_ready) return ISR_OK;
  const int n = p->n;
  isr_ctx* c = p->ctx;
  ISR_TRY(p->d_tmp.reserve((size_t)n * n));
  scale_cols(c->stream, n, p->d_Ginv.p, p->d_werr.p, 0, p->d_tmp.p);      // Ginv diag(err)
  ISR_TRY(gemm(c, false, true, n, p->d_tmp.p, p->d_tmp.p, p->d_Cov.p));   // Cov
  ISR_TRY(gemm(c, true, false, n, p->d_R.p, p->d_R.p, p->d_InvCov.p));    // G^T W G
  ISR_CUDA(cudaGetLastError());
  p->cov_ready = true;
  return ISR_OK;
}

// R = diag(1/err) G (column-major, as the covariance / Tikhonov code reads it) and the two operators the toy kernel
// consumes, Sop = G^-1 and Rop = R, row-major with the leading dimension padded to even and a zero pad column: one launch.
__global__ void sle_operators_kernel(int n, int ld, const double* __restrict__ G, const double* __restrict__ Ginv,
                                     const double* __restrict__ err, double* __restrict__ R, double* __restrict__ Sop,
                                     double* __restrict__ Rop) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ld) return;
  const int i = idx / ld, j = idx % ld;
  double sv = 0.0, rv = 0.0;
  if (j < n) {
    const size_t cm = (size_t)i + (size_t)n * j;
    sv = Ginv[cm];
    rv = G[cm] / err[i];
    R[cm] = rv;
  }
  Sop[idx] = sv;
  Rop[idx] = rv;
}

// The two halves of sle_operators_kernel, for the factorisation's two streams: Rop / R need only G and the errors, Sop only
// the inverse.
__global__ void rop_kernel(int n, int ld, const double* __restrict__ G, const double* __restrict__ err, double* __restrict__ R,
                           double* __restrict__ Rop) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ld) return;
  const int i = idx / ld, j = idx % ld;
  double rv = 0.0;
  if (j < n) {
    const size_t cm = (size_t)i + (size_t)n * j;
    rv = G[cm] / err[i];
    R[cm] = rv;
  }
  Rop[idx] = rv;
}
__global__ void sop_kernel(int n, int ld, const double* __restrict__ Ginv, double* __restrict__ Sop) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ld) return;
  const int i = idx / ld, j = idx % ld;
  Sop[idx] = j < n ? Ginv[(size_t)i + (size_t)n * j] : 0.0;
}

// Operators the toy kernel consumes: row-major, leading dimension padded to even, pad column zero.
int launch_sle_operators(isr_problem* p) {
  const int n = p->n, ld = (n + 1) & ~1;
  cudaStream_t st = p->ctx->stream;
  sle_operators_kernel<<<(n * ld + 255) / 256, 256, 0, st>>>(n, ld, p->d_G.p, p->d_Ginv.p, p->d_werr.p, p->d_R.p, p->d_Sop.p, p->d_Rop.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}
int select_sle(isr_problem* p) {
  if (!p->factored) { set_error("select: isr_factor has not been called"); return ISR_ESTATE; }
  ISR_TRY(launch_sle_operators(p));
  p->tri_sel[0] = p->tri_sle[0];
  p->tri_sel[1] = p->tri_sle[1];
  p->solver_selected = true;
  p->selected_kind = 1;
  return ISR_OK;
}

// --------------------------------------------------------------------------------------------------
// min / max and TH1-style histogram
// --------------------------------------------------------------------------------------------------
__global__ void minmax_kernel(int64_t n, const double* __restrict__ v, double* __restrict__ part) {
  __shared__ double smin[32], smax[32];
  double lo = INFINITY, hi = -INFINITY;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const double x = v[q];
    lo = fmin(lo, x);
    hi = fmax(hi, x);
  }
  for (int s = 16; s >= 1; s >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smin[w] = lo; smax[w] = hi; }
  __syncthreads();
  if (w == 0) {
    lo = l < (blockDim.x >> 5) ? smin[l] : INFINITY;
    hi = l < (blockDim.x >> 5) ? smax[l] : -INFINITY;
    for (int s = 16; s >= 1; s >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if (l == 0) { part[2 * blockIdx.x] = lo; part[2 * blockIdx.x + 1] = hi; }
  }
}

__global__ void hist_kernel(int64_t n, const double* __restrict__ v, int nbins, double lo, double hi,
                            uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t sh[];
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x) sh[q] = 0;
  __syncthreads();
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh[th1_bin(nbins, lo, hi, v[q])], 1u);
  __syncthreads();
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x)
    if (sh[q]) atomicAdd(&counts[q], sh[q]);
}

// Device-resident variants for multi-GPU pipelines (no host round trip, asynchronous on the context stream):
// out_dev[0] = min, out_dev[1] = -max, so that ONE all_reduce(MIN) over ranks yields the global range.
__global__ void minmax_final_kernel(int blocks, const double* __restrict__ part, double* __restrict__ out) {
  double lo = INFINITY, hi = -INFINITY;
  for (int q = threadIdx.x; q < blocks; q += 32) { lo = fmin(lo, part[2 * q]); hi = fmax(hi, part[2 * q + 1]); }
  for (int s = 16; s >= 1; s >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if (threadIdx.x == 0) { out[0] = lo; out[1] = -hi; }
}
int minmax_to_device(isr_ctx* c, int64_t n, const double* v_dev, double* out_dev) {
  if (n <= 0) { set_error("minmax: empty input"); return ISR_EINVAL; }
  const int blocks = (int)std::min<int64_t>(c->sm_count * 4, (n + 255) / 256);
  ISR_TRY(c->scratch_d.reserve(2 * blocks));
  minmax_kernel<<<blocks, 256, 0, c->stream>>>(n, v_dev, c->scratch_d.p); ISR_COUNT_LAUNCH();
  minmax_final_kernel<<<1, 32, 0, c->stream>>>(blocks, c->scratch_d.p, out_dev); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}
// TH1 counts with the range read from device memory ({lo, -hi} as minmax_to_device / the all_reduce leave it);
// counts_dev[nbins + 2] int32 is overwritten.
__global__ void hist_range_dev_kernel(int64_t n, const double* __restrict__ v, int nbins, const double* __restrict__ range,
                                      int* __restrict__ counts) {
  extern __shared__ uint32_t sh[];
  const double lo = range[0], hi = -range[1];
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x) sh[q] = 0;
  __syncthreads();
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh[th1_bin(nbins, lo, hi, v[q])], 1u);
  __syncthreads();
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x)
    if (sh[q]) atomicAdd(&counts[q], (int)sh[q]);
}
int histogram_range_device(isr_ctx* c, int64_t n, const double* v_dev, int nbins, const double* range_dev, int* counts_dev) {
  if (nbins < 1 || nbins > 8190) { set_error("histogram: nbins out of range"); return ISR_EINVAL; }
  ISR_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(int) * (nbins + 2), c->stream));
  if (n > 0) {
    const int blocks = (int)std::min<int64_t>(c->sm_count * 4, (n + 255) / 256);
    hist_range_dev_kernel<<<blocks, 256, sizeof(uint32_t) * (nbins + 2), c->stream>>>(n, v_dev, nbins, range_dev, counts_dev); ISR_COUNT_LAUNCH();
  }
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

int minmax_device(isr_ctx* c, int64_t n, const double* v_dev, double* lo, double* hi) {
  if (n <= 0) { set_error("minmax: empty input"); return ISR_EINVAL; }
  const int blocks = (int)std::min<int64_t>(c->sm_count * 4, (n + 255) / 256);
  ISR_TRY(c->scratch_d.reserve(2 * blocks));
  minmax_kernel<<<blocks, 256, 0, c->stream>>>(n, v_dev, c->scratch_d.p); ISR_COUNT_LAUNCH();
  std::vector<double> h(2 * blocks);
  ISR_CUDA(cudaMemcpyAsync(h.data(), c->scratch_d.p, sizeof(double) * 2 * blocks, cudaMemcpyDeviceToHost, c->stream));
  ISR_CUDA(cudaStreamSynchronize(c->stream));
  double a = INFINITY, b = -INFINITY;
  for (int q = 0; q < blocks; ++q) { a = std::min(a, h[2 * q]); b = std::max(b, h[2 * q + 1]); }
  *lo = a;
  *hi = b;
  return ISR_OK;
}

int histogram_device(isr_ctx* c, int64_t n, const double* v_dev, int nbins, double lo, double hi, uint32_t* counts_host) {
  if (nbins < 1 || nbins > 8190) { set_error("histogram: nbins out of range"); return ISR_EINVAL; }
  DevBuf<uint32_t>& cnt = c->scratch_u;
  ISR_TRY(cnt.reserve(nbins + 2));
  ISR_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(uint32_t) * (nbins + 2), c->stream));
  if (n > 0) {
    const int blocks = (int)std::min<int64_t>(c->sm_count * 4, (n + 255) / 256);
    hist_kernel<<<blocks, 256, sizeof(uint32_t) * (nbins + 2), c->stream>>>(n, v_dev, nbins, lo, hi, cnt.p); ISR_COUNT_LAUNCH();
  }
  ISR_CUDA(cudaMemcpyAsync(counts_host, cnt.p, sizeof(uint32_t) * (nbins + 2), cudaMemcpyDeviceToHost, c->stream));
  ISR_CUDA(cudaStreamSynchronize(c->stream));
  return ISR_OK;
}

// --------------------------------------------------------------------------------------------------
// Host-buffer toy batch: H2D copies of V chunks overlap the kernels of the previous chunk.
// --------------------------------------------------------------------------------------------------
// Two device slots of up to 2^17 toys; the copy of chunk k + 2 is issued as soon as the kernels of chunk k have been
// enqueued.  host_pipe_begin issues the first two copies -- behind everything the context stream holds at that moment
// (an earlier asynchronous campaign may still read d_Vh), but BEFORE the caller enqueues more work there, so that they
// run under a matrix assembly / factorisation enqueued next; host_pipe_run enqueues the rest.
static int host_pipe_copy(isr_problem* p, int64_t T, const double* V, int k) {
  const int n = p->n, q = k & 1;
  const int64_t slot = p->pipe_slot, t0 = (int64_t)k * slot, tc = std::min<int64_t>(slot, T - t0);
  if (k >= 2) ISR_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_done[q], 0));   // slot free again
  ISR_CUDA(cudaMemcpyAsync(p->d_Vh.p + (size_t)q * slot * n, V + t0 * n, sizeof(double) * tc * n, cudaMemcpyHostToDevice, p->copy_stream));
  ISR_CUDA(cudaEventRecord(p->ev_copied[q], p->copy_stream));
  return ISR_OK;
}
int host_pipe_begin(isr_problem* p, int64_t T, const double* V) {
  const int n = p->n;
  isr_ctx* c = p->ctx;
  if (T <= 0) return ISR_OK;
  p->pipe_slot = std::min<int64_t>(T, 1 << 17);
  ISR_TRY(p->d_Vh.reserve((size_t)2 * p->pipe_slot * n));
  if (!p->copy_stream) {
    ISR_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) {
      ISR_CUDA(cudaEventCreateWithFlags(&p->ev_copied[q], cudaEventDisableTiming));
      ISR_CUDA(cudaEventCreateWithFlags(&p->ev_done[q], cudaEventDisableTiming));
    }
  }
  ISR_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  ISR_CUDA(cudaStreamWaitEvent(p->copy_stream, c->ev_fork, 0));
  const int64_t nchunks = (T + p->pipe_slot - 1) / p->pipe_slot;
  for (int k = 0; k < 2 && k < nchunks; ++k) ISR_TRY(host_pipe_copy(p, T, V, k));
  return ISR_OK;
}
// chi2_dev: T values on the device (or nullptr); chi2_out / bcs_out: host destinations copied chunk by chunk (or nullptr).
// On failure the copy stream is drained before returning: it may still be reading the caller's V.
int host_pipe_run(isr_problem* p, int64_t T, const double* V, const double* b0_dev, const ToyOut& base, double* chi2_out,
                  double* bcs_out) {
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  if (T <= 0) return ISR_OK;
  const int64_t slot = p->pipe_slot, nchunks = (T + slot - 1) / slot;
  int rc = ISR_OK;
  auto body = [&]() -> int {
    if (bcs_out) ISR_TRY(p->d_Bh.reserve((size_t)2 * slot * n));
    for (int64_t k = 0; k < nchunks; ++k) {
      const int q = (int)(k & 1);
      const int64_t t0 = k * slot, tc = std::min<int64_t>(slot, T - t0);
      ISR_CUDA(cudaStreamWaitEvent(st, p->ev_copied[q], 0));
      ToyOut o = base;
      o.chi2 = base.chi2 ? base.chi2 + t0 : nullptr;
      o.bcs = bcs_out ? p->d_Bh.p + (size_t)q * slot * n : nullptr;
      ISR_TRY(toy_batch_device(p, tc, p->d_Vh.p + (size_t)q * slot * n, b0_dev, o));
      if (bcs_out) ISR_CUDA(cudaMemcpyAsync(bcs_out + t0 * n, o.bcs, sizeof(double) * tc * n, cudaMemcpyDeviceToHost, st));
      ISR_CUDA(cudaEventRecord(p->ev_done[q], st));
      if (k + 2 < nchunks) ISR_TRY(host_pipe_copy(p, T, V, (int)(k + 2)));
      // this chunk's chi2 goes back while the next chunk's toys are still coming in (the two directions have their own
      // copy engines); the values also stay on the device for isr_last_chi2_minmax / isr_last_chi2_histogram
      if (chi2_out && o.chi2) ISR_CUDA(cudaMemcpyAsync(chi2_out + t0, o.chi2, sizeof(double) * tc, cudaMemcpyDeviceToHost, st));
    }
    return ISR_OK;
  };
  rc = body();
  if (rc != ISR_OK) cudaStreamSynchronize(p->copy_stream);
  return rc;
}

int toy_batch_host(isr_problem* p, int64_t T, const double* V, const double* bcs0, double* chi2_out,
                   double* sum_bcs_out, double* ratio_stats_out, uint32_t* ratio_hist_out, double* bcs_out) {
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  if (!p->solver_selected) {
    set_error("isr_toy_batch: no solver selected (call isr_factor first)");
    return ISR_ESTATE;
  }
  if (T <= 0) return ISR_OK;
  ISR_TRY(p->d_chi2h.reserve((size_t)T));
  ISR_TRY(p->d_b0.reserve(n));
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  ISR_TRY(host_pipe_begin(p, T, V));
  p->h_b0.assign(bcs0, bcs0 + n);      // d_b0 mirrors h_b0 (isr_toy_batch_dev skips its upload on a match)
  ISR_CUDA(cudaMemcpyAsync(p->d_b0.p, p->h_b0.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemsetAsync(p->d_stat.p, 0, sizeof(double) * 4 * n, st));
  if (ratio_hist_out) {
    ISR_TRY(p->d_rhist.reserve((size_t)n * 1026));
    ISR_CUDA(cudaMemsetAsync(p->d_rhist.p, 0, sizeof(uint32_t) * (size_t)n * 1026, st));
  }
  ToyOut o;
  o.chi2 = chi2_out ? p->d_chi2h.p : nullptr;
  o.sum_bcs = sum_bcs_out ? p->d_stat.p : nullptr;
  o.ratio_stats = ratio_stats_out ? p->d_stat.p + n : nullptr;
  o.rhist = ratio_hist_out ? p->d_rhist.p : nullptr;
  ISR_TRY(host_pipe_run(p, T, V, p->d_b0.p, o, chi2_out, bcs_out));
  p->last_chi2_count = chi2_out ? T : 0;
  if (sum_bcs_out) ISR_CUDA(cudaMemcpyAsync(sum_bcs_out, p->d_stat.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (ratio_stats_out)
    ISR_CUDA(cudaMemcpyAsync(ratio_stats_out, p->d_stat.p + n, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
  if (ratio_hist_out)
    ISR_CUDA(cudaMemcpyAsync(ratio_hist_out, p->d_rhist.p, sizeof(uint32_t) * (size_t)n * 1026, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

// --------------------------------------------------------------------------------------------------
// IterISRInterpSolver::solve (src/IterISRInterpSolver.cpp:45-57): the iterative radiative-correction method on
// the assembled operator:  b <- v;  repeat nIter times:  delta = (G b) / b,  b <- v / delta;
// Cov = diag((err / delta)^2).  One block; the matrix-vector product is a warp per row (fixed order).
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) iterative_kernel(int n, int niter, const double* __restrict__ G /* col-major */,
                                                          const double* __restrict__ vcs, double* __restrict__ bcs,
                                                          double* __restrict__ radcorr) {
  extern __shared__ double sb[];      // current b
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = threadIdx.x; j < n; j += blockDim.x) sb[j] = vcs[j];
  __syncthreads();
  for (int it = 0; it < niter; ++it) {
    for (int i = w; i < n; i += nw) {
      double acc = 0.0;
      for (int j = lane; j < n; j += 32) acc = fma(G[(size_t)i + (size_t)n * j], sb[j], acc);
      for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) radcorr[i] = acc / sb[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) sb[j] = vcs[j] / radcorr[j];
    __syncthreads();
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) bcs[j] = sb[j];
}

int iterative_solve(isr_problem* p, int niter, const double* vcs, double* bcs_out, double* radcorr_out) {
  if (!p->has_G) { set_error("isr_iterative_solve: no matrix (call isr_assemble first)"); return ISR_ESTATE; }
  if (niter < 0) { set_error("isr_iterative_solve: niter < 0"); return ISR_EINVAL; }
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  double *dv = p->d_stat.p, *db = p->d_stat.p + n, *dr = p->d_stat.p + 2 * n;
  ISR_CUDA(cudaMemcpyAsync(dv, vcs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemsetAsync(dr, 0, sizeof(double) * n, st));
  iterative_kernel<<<1, 1024, sizeof(double) * n, st>>>(n, niter, p->d_G.p, dv, db, dr); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  ISR_CUDA(cudaMemcpyAsync(bcs_out, db, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (radcorr_out) ISR_CUDA(cudaMemcpyAsync(radcorr_out, dr, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

// --------------------------------------------------------------------------------------------------
// Toys drawn on the device (rng.cu), block by block through the fused kernel.
// --------------------------------------------------------------------------------------------------
// Toys per generator block: the largest power of two whose block stays under 512 MiB (2^18 toys at N = 200, 2^17 at
// N = 500).  Measured at N = 200, 2^20 toys (tools/exp_rng.py): 6.43 / 6.22 / 5.52 / 5.45 / 5.41 ms for 2^14 ... 2^18 toys per
// block against 5.38 ms with resident toys -- every block boundary costs the toy kernel a tail and a ramp, and L2
// residency of the block (2^15: 52 MB) buys nothing because the kernel is DMMA-bound.  ISR_GEN_CHUNK=<log2> overrides.
int64_t gen_chunk(int n) {
  static const int forced = [] {
    const char* e = getenv("ISR_GEN_CHUNK");
    const int lg = e ? atoi(e) : 0;
    return lg >= 10 && lg <= 22 ? lg : 0;
  }();
  if (forced) return (int64_t)1 << forced;
  int lg = 14;
  while (lg < 20 && ((int64_t)2 << lg) * n * (int64_t)sizeof(double) <= ((int64_t)512 << 20)) ++lg;
  return (int64_t)1 << lg;
}

// Blocks of gen_chunk(n) toys in two slots of d_Vh.  Block k + 1 is drawn on the generator stream while block k is being
// consumed: the generator's thread blocks move onto the SMs as the persistent toy CTAs of block k retire, so the tail of
// one kernel and the ramp of the next overlap instead of adding up (the two kernels never share an SM: a toy CTA takes
// the whole register file).  d_gen = [vcs | vcs_err] must be on the device (ordered before c->ev_fork / on c->stream).
// block0_drawn: block 0 is already in slot 0, signalled by c->ev_join on the generator stream (isr_chi2_test draws it
// under the assembly and the factorisation).
int drawn_toys_run(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* b0_dev, const ToyOut& o, bool block0_drawn) {
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  const int64_t chunk = std::min<int64_t>(T, gen_chunk(n));
  ISR_TRY(p->d_Vh.reserve((size_t)(T > chunk ? 2 : 1) * chunk * n));
  const double* gv = p->d_gen.p;
  int k = 0;
  for (int64_t t0 = 0; t0 < T; t0 += chunk, ++k) {
    const int64_t tc = std::min<int64_t>(chunk, T - t0);
    const int s = k & 1;
    if (k == 0 && !block0_drawn) {                           // block 0, behind everything enqueued so far
      ISR_CUDA(cudaEventRecord(c->ev_fork, st));
      ISR_CUDA(cudaStreamWaitEvent(c->gen_stream, c->ev_fork, 0));
      ISR_TRY(draw_toys_device(c, tc, n, first_toy, seed, gv, gv + n, p->d_Vh.p, c->gen_stream));
      ISR_CUDA(cudaEventRecord(c->ev_join, c->gen_stream));
    }
    if (t0 + chunk < T) {                                    // block k + 1 into the other slot, once block k - 1 has left it
      const int64_t t1 = t0 + chunk, tn = std::min<int64_t>(chunk, T - t1);
      if (k >= 1) ISR_CUDA(cudaStreamWaitEvent(c->gen_stream, c->ev_slot[s ^ 1], 0));
      ISR_TRY(draw_toys_device(c, tn, n, first_toy + t1, seed, gv, gv + n, p->d_Vh.p + (size_t)(s ^ 1) * chunk * n, c->gen_stream));
      ISR_CUDA(cudaEventRecord(c->ev_gen[s ^ 1], c->gen_stream));
    }
    ISR_CUDA(cudaStreamWaitEvent(st, k == 0 ? c->ev_join : c->ev_gen[s], 0));
    ToyOut oc = o;
    oc.chi2 = o.chi2 ? o.chi2 + t0 : nullptr;
    ISR_TRY(toy_batch_device(p, tc, p->d_Vh.p + (size_t)s * chunk * n, b0_dev, oc));
    ISR_CUDA(cudaEventRecord(c->ev_slot[s], st));
  }
  return ISR_OK;
}

// Toy campaign by count (isr_toy_campaign).
int toy_campaign_device(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* vcs,
                        const double* vcs_err, const double* bcs0, double* chi2_dev, double* sum_bcs_dev,
                        double* ratio_stats_dev, uint32_t* rhist_dev) {
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  if (!p->solver_selected) { set_error("isr_toy_campaign: no solver selected (call isr_factor first)"); return ISR_ESTATE; }
  if (T <= 0) return ISR_OK;
  for (int i = 0; i < n; ++i)
    if (!(vcs_err[i] >= 0.0) || !std::isfinite(vcs_err[i]) || !std::isfinite(vcs[i])) {
      set_error("isr_toy_campaign: vcs / vcs_err[%d] not finite or negative", i);
      return ISR_EINVAL;
    }
  ISR_TRY(p->d_b0.reserve(n));
  ISR_TRY(p->d_gen.reserve((size_t)2 * n));
  p->h_b0.assign(bcs0, bcs0 + n);
  p->h_gen.assign(vcs, vcs + n);
  p->h_gen.insert(p->h_gen.end(), vcs_err, vcs_err + n);
  ISR_CUDA(cudaMemcpyAsync(p->d_b0.p, p->h_b0.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(p->d_gen.p, p->h_gen.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  ToyOut o;
  o.chi2 = chi2_dev;
  o.sum_bcs = sum_bcs_dev;
  o.ratio_stats = ratio_stats_dev;
  o.rhist = rhist_dev;
  return drawn_toys_run(p, T, first_toy, seed, p->d_b0.p, o, /*block0_drawn=*/false);
}

}  // namespace isr
