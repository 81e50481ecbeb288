// A whole chi2 / ratio test in one call (isr_chi2_test): chi2TestModel / chi2Test of src/Chi2Test.cpp:29-151 and
// ratioTestModel of src/RatioTest.cpp:13-62 end to end --
//     [assemble G] -> [factor] -> T toys: b_k = S v_k, chi2_k = |R (b_k - b0)|^2 -> min / max -> TH1F cells
// -- as ONE stream-ordered chain with a single host synchronisation at its end.
//
// Round 1 drove the same chain through separate ABI calls, each of which synchronised: the GPU idled 25-40 us per call
// while the host woke up, and isr_factor waited for the certificate of its unpivoted inverse before the toy kernel could
// even be enqueued (0.925 ms for the north-star target's per-GPU share against 0.65 ms of toy kernel).  Here
//   * the certificate is DEFERRED (factor_enqueue / factor_finish, solve.cu): the toy kernels run speculatively on the
//     unpivoted inverse; the verdict travels back with the results and, if it is negative, the factorisation is redone with
//     partial pivoting and the toys are run again -- the call still only returns certified results;
//   * host toys: the first two chunks are copied while the matrix is assembled and inverted; device-drawn toys: the first
//     block is generated on a side stream meanwhile (the Gauss-Jordan cluster kernel occupies 4 of 148 SMs);
//   * with a communicator the two all-reduces of the multi-GPU test are enqueued in the same stream: MIN of {min, -max},
//     then SUM of one packed FP64 buffer [TH1F cells | sum b | ratio stats | toy count | certificate flag] -- the flag makes
//     the re-run decision collective, so that no rank can leave the others waiting in an all-reduce.
#include "comm.hpp"
#include "panel_device.cuh"
#include "problem.hpp"
#include "solve.hpp"
#include "toy_batch.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace isr {

__global__ void empty_range_kernel(double* __restrict__ range) {
  if (threadIdx.x < 2) range[threadIdx.x] = INFINITY;   // {min, -max} of nothing: neutral for the MIN all-reduce
}

// pack[0 .. nb) = TH1F cells, [nb .. nb + N) = sum of solutions, [nb + N .. nb + 4N) = ratio statistics,
// pack[nb + 4N] = toys of this rank, pack[nb + 4N + 1] = 1 if this rank's factorisation failed its certificate.
__global__ void pack_test_kernel(int nb, int n, const int* __restrict__ cells, const double* __restrict__ stat, double toys,
                                 const double* __restrict__ cert, int assumed_lower, double tol, double* __restrict__ pack) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int len = nb + 4 * n + 2;
  if (idx >= len) return;
  double v;
  if (idx < nb) v = cells ? (double)cells[idx] : 0.0;
  else if (idx < nb + 4 * n) v = stat[idx - nb];
  else if (idx == nb + 4 * n) v = toys;
  else {
    v = 0.0;
    if (cert) {
      const bool ok = cert[0] <= tol && (!assumed_lower || (cert[1] == 0.0 && cert[2] == 0.0));   // (a NaN residual fails)
      v = ok ? 0.0 : 1.0;
    }
  }
  pack[idx] = v;
}

// ---- fused tail of the test: two launches instead of six (min/max partials, final reduction, memset, TH1F, packing) --
// range_kernel: per-block min / max of the chi2 values; the block that arrives LAST (fence + arrival counter) reduces the
// partials in a fixed order to range = {min, -max} (the layout one all-reduce(MIN) over ranks needs), zeroes the TH1F
// cells for the next kernel and re-arms the counter.
// With a peer communicator (use_peer) the last block also EXCHANGES the range with the other GPUs over NVLink mailboxes
// (peer_device.cuh): the all-reduce(MIN) of the multi-GPU test costs no launch of its own.
__global__ void __launch_bounds__(256) range_kernel(int64_t n, const double* __restrict__ v, double* __restrict__ part,
                                                     unsigned* __restrict__ counter, double* range,
                                                     int* __restrict__ cells, int ncells, PeerComm pc, unsigned long long epoch, int use_peer) {
  __shared__ double smin[8], smax[8];
  __shared__ bool last;
  double lo = INFINITY, hi = -INFINITY;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const double x = v[q];
    lo = fmin(lo, x);
    hi = fmax(hi, x);
  }
  auto block_reduce = [&]() {
    for (int s = 16; s >= 1; s >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { smin[w] = lo; smax[w] = hi; }
    __syncthreads();
    if (w == 0) {
      lo = l < 8 ? smin[l] : INFINITY;
      hi = l < 8 ? smax[l] : -INFINITY;
      for (int s = 4; s >= 1; s >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
      }
    }
  };
  block_reduce();
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = lo;
    part[2 * blockIdx.x + 1] = hi;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  lo = INFINITY; hi = -INFINITY;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    lo = fmin(lo, __ldcg(&part[2 * q]));
    hi = fmax(hi, __ldcg(&part[2 * q + 1]));
  }
  __syncthreads();
  block_reduce();
  if (threadIdx.x == 0) { range[0] = lo; range[1] = -hi; *counter = 0u; }
  for (int q = threadIdx.x; q < ncells; q += blockDim.x) cells[q] = 0;
  if (use_peer) {
    __syncthreads();
    peer_allreduce_block(pc, epoch, range, range, 2, 1);      // {min, -max} over all ranks
  }
}

// cells_pack_kernel: TH1F cells of the chi2 values against the (global) range (shared-memory counters, integer atomics),
// then the block that arrives last writes the packed buffer (see pack_test_kernel for the layout).
__global__ void __launch_bounds__(256) cells_pack_kernel(int64_t n, const double* __restrict__ v, int nbins,
                                                          const double* __restrict__ range, int* __restrict__ cells,
                                                          unsigned* __restrict__ counter, int npts, const double* __restrict__ stat,
                                                          double toys, const double* __restrict__ cert, int assumed_lower, double tol,
                                                          double* pack, PeerComm pc, unsigned long long epoch, int use_peer,
                                                          double* host_out) {
  extern __shared__ uint32_t sh[];
  __shared__ bool last;
  const double lo = range[0], hi = -range[1];
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x) sh[q] = 0;
  __syncthreads();
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh[th1_bin(nbins, lo, hi, v[q])], 1u);
  __syncthreads();
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x)
    if (sh[q]) atomicAdd(&cells[q], (int)sh[q]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int nb = nbins + 2, len = nb + 4 * npts + 2;
  for (int idx = threadIdx.x; idx < len; idx += blockDim.x) {
    double val;
    if (idx < nb) val = (double)__ldcg(&cells[idx]);
    else if (idx < nb + 4 * npts) val = stat[idx - nb];
    else if (idx == nb + 4 * npts) val = toys;
    else {
      val = 0.0;
      if (cert) val = (cert[0] <= tol && (!assumed_lower || (cert[1] == 0.0 && cert[2] == 0.0))) ? 0.0 : 1.0;
    }
    pack[idx] = val;
  }
  if (threadIdx.x == 0) *counter = 0u;
  if (use_peer) {
    __syncthreads();
    peer_allreduce_block(pc, epoch, pack, pack, len, 0);      // cells, sums, toy count and certificate flag over all ranks
  }
  if (host_out) {
    // the final result goes straight into the caller-visible page-locked buffer (zero-copy stores over PCIe): no separate
    // device-to-host copy operation behind the kernel
    __syncthreads();
    if (threadIdx.x < 2) host_out[threadIdx.x] = range[threadIdx.x];
    for (int idx = threadIdx.x; idx < len; idx += blockDim.x) host_out[2 + idx] = pack[idx];
    __threadfence_system();
  }
}

static int enqueue_toys(isr_problem* p, const isr_chi2_test_args& a, const TestWants& w, bool need_chi2, bool first_attempt) {
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  const int64_t T = a.n_toys;
  const bool ratio = (a.flags & ISR_TEST_RATIO) != 0;
  ISR_CUDA(cudaMemsetAsync(p->d_stat.p, 0, sizeof(double) * 4 * n, st));
  if (w.rhist) ISR_CUDA(cudaMemsetAsync(p->d_rhist.p, 0, sizeof(uint32_t) * (size_t)n * 1026, st));
  if (T <= 0) return ISR_OK;
  ToyOut o;
  o.chi2 = need_chi2 ? p->d_chi2h.p : nullptr;
  o.sum_bcs = (w.sum || ratio) ? p->d_stat.p : nullptr;    // (the column statistics cost the kernel 3.7 %: DESIGN K5)
  o.ratio_stats = ratio ? p->d_stat.p + n : nullptr;
  o.rhist = w.rhist ? p->d_rhist.p : nullptr;
  if (a.V_dev) return toy_batch_device(p, T, a.V_dev, p->d_b0.p, o);
  if (a.V) {
    if (!first_attempt) ISR_TRY(host_pipe_begin(p, T, a.V));     // (the first attempt's copies were issued before the assembly)
    return host_pipe_run(p, T, a.V, p->d_b0.p, o, a.chi2_out, nullptr);
  }
  return drawn_toys_run(p, T, a.first_toy, a.seed, p->d_b0.p, o, /*block0_drawn=*/first_attempt);
}

// wants: what is COMPUTED (and reduced over the communicator) is decided by these flags, not by which output pointers a
// rank passes -- every rank of a communicator must take the same path through the collectives.
int chi2_test(isr_problem* p, const isr_chi2_test_args& a, TestWants w) {
  w.rhist = w.rhist && (a.flags & ISR_TEST_RATIO) != 0;
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  const int64_t T = a.n_toys;
  const bool do_asm = (a.flags & ISR_TEST_ASSEMBLE) != 0, do_fac = (a.flags & ISR_TEST_FACTOR) != 0;
  const bool drawn = !a.V && !a.V_dev;
  if (T < 0 || !a.bcs0) { set_error("isr_chi2_test: bad argument (n_toys < 0 or bcs0 == NULL)"); return ISR_EINVAL; }
  if (a.V && a.V_dev) { set_error("isr_chi2_test: give V or V_dev, not both"); return ISR_EINVAL; }
  if (do_fac && !a.vcs_err) { set_error("isr_chi2_test: ISR_TEST_FACTOR needs vcs_err"); return ISR_EINVAL; }
  if (drawn && T > 0 && (!a.vcs || !a.vcs_err)) { set_error("isr_chi2_test: toys drawn on the device need vcs and vcs_err"); return ISR_EINVAL; }
  if (a.nbins < 0 || a.nbins > 8190) { set_error("isr_chi2_test: nbins out of range"); return ISR_EINVAL; }
  if (a.comm && a.comm->ctx != c) { set_error("isr_chi2_test: the communicator belongs to another context"); return ISR_EINVAL; }
  if (!do_asm && !p->has_G) { set_error("isr_chi2_test: no matrix (set ISR_TEST_ASSEMBLE or call isr_assemble first)"); return ISR_ESTATE; }
  if (!do_fac && !p->solver_selected) { set_error("isr_chi2_test: no solver selected (set ISR_TEST_FACTOR or select one first)"); return ISR_ESTATE; }
  if (drawn && T > 0)
    for (int i = 0; i < n; ++i)
      if (!(a.vcs_err[i] >= 0.0) || !std::isfinite(a.vcs_err[i]) || !std::isfinite(a.vcs[i])) {
        set_error("isr_chi2_test: vcs / vcs_err[%d] not finite or negative", i);
        return ISR_EINVAL;
      }
  const int nb = a.nbins > 0 ? a.nbins + 2 : 0;
  const bool need_range = a.nbins > 0 || w.range;
  const bool need_chi2 = need_range || a.chi2_out;
  const int plen = nb + 4 * n + 2;

  // ---- workspaces
  ISR_TRY(p->d_b0.reserve(n));
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  if (need_chi2) ISR_TRY(p->d_chi2h.reserve((size_t)std::max<int64_t>(T, 1)));
  if (w.rhist) ISR_TRY(p->d_rhist.reserve((size_t)n * 1026));
  ISR_TRY(p->d_pack.reserve((size_t)plen + 2));          // [range (2) | pack]
  ISR_TRY(p->d_cells.reserve((size_t)std::max(nb, 1)));
  ISR_TRY(p->d_part.reserve((size_t)4 * c->sm_count));
  if (!p->d_arrive.p) {
    ISR_TRY(p->d_arrive.reserve(2));
    ISR_CUDA(cudaMemsetAsync(p->d_arrive.p, 0, 2 * sizeof(unsigned), st));     // the kernels re-arm their counters themselves
  }
  if ((size_t)plen + 2 > p->pin_test_n) {
    if (p->pin_test) cudaFreeHost(p->pin_test);
    p->pin_test = nullptr; p->pin_test_n = 0;
    ISR_CUDA(cudaMallocHost((void**)&p->pin_test, sizeof(double) * ((size_t)plen + 2)));
    p->pin_test_n = (size_t)plen + 2;
    p->pin_test_dev = nullptr;
    if (cudaHostGetDevicePointer((void**)&p->pin_test_dev, p->pin_test, 0) != cudaSuccess) { cudaGetLastError(); p->pin_test_dev = nullptr; }
  }
  double* d_range = p->d_pack.p;
  double* d_pack = p->d_pack.p + 2;

  // ---- toys that can start before the matrix exists
  if (a.V && T > 0) ISR_TRY(host_pipe_begin(p, T, a.V));
  if (drawn && T > 0) {
    const int64_t chunk = std::min<int64_t>(T, gen_chunk(n));
    ISR_TRY(p->d_Vh.reserve((size_t)(T > chunk ? 2 : 1) * chunk * n));
    ISR_TRY(p->d_gen.reserve((size_t)2 * n));
    p->h_gen.assign(a.vcs, a.vcs + n);
    p->h_gen.insert(p->h_gen.end(), a.vcs_err, a.vcs_err + n);
    ISR_CUDA(cudaMemcpyAsync(p->d_gen.p, p->h_gen.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
    ISR_CUDA(cudaEventRecord(c->ev_fork, st));
    ISR_CUDA(cudaStreamWaitEvent(c->gen_stream, c->ev_fork, 0));
    ISR_TRY(draw_toys_device(c, chunk, n, a.first_toy, a.seed, p->d_gen.p, p->d_gen.p + n, p->d_Vh.p, c->gen_stream));
    ISR_CUDA(cudaEventRecord(c->ev_join, c->gen_stream));
  }

  // ---- matrix and factorisation (enqueue only)
  int rc = ISR_OK;
  auto drain = [&](int code) {      // a failure after the pipeline started: nothing may still read the caller's buffers
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    cudaStreamSynchronize(c->aux_stream);
    cudaStreamSynchronize(c->gen_stream);
    cudaStreamSynchronize(st);
    p->cert_pending = false;
    return code;
  };
  if (!p->ev_time[0])
    for (int q = 0; q < 5; ++q) ISR_CUDA(cudaEventCreate(&p->ev_time[q]));
  cudaEventRecord(p->ev_time[0], st);
  if (do_asm && (rc = assemble(p, a.ecm_err)) != ISR_OK) return drain(rc);
  cudaEventRecord(p->ev_time[1], st);
  p->h_b0.assign(a.bcs0, a.bcs0 + n);      // (after the assembly's launches: the GPU is already busy while the host stages this)
  if (cudaMemcpyAsync(p->d_b0.p, p->h_b0.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_error("isr_chi2_test: copy of bcs0 failed (%s)", cudaGetErrorString(cudaGetLastError()));
    return drain(ISR_ECUDA);
  }
  if (do_fac && (rc = factor_enqueue(p, a.vcs_err)) != ISR_OK) return drain(rc);

  int refactored = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    cudaEventRecord(p->ev_time[2], st);
    if ((rc = enqueue_toys(p, a, w, need_chi2, attempt == 0)) != ISR_OK) return drain(rc);
    cudaEventRecord(p->ev_time[3], st);
    // ---- range, cells, packed buffer, collectives
    const bool cert_here = do_fac && attempt == 0;
    bool packed_by_peer = false, result_on_host = false;
    const int assumed = (p->tri_sle[0] || p->tri_sle[1]) ? 1 : 0;
    if (T > 0 && a.nbins > 0) {
      // the usual case: two fused launches (the stat sums of the toy kernels are complete: same stream)
      const int blocks = (int)std::min<int64_t>(c->sm_count * 2, (T + 255) / 256);
      // the two exchanges ride inside the tail kernels when the ranks have NVLink mailboxes (else: two NCCL all-reduces)
      const bool peer = a.comm && a.comm->nranks > 1 && comm_peer_fits(a.comm, plen);
      PeerComm pc{};
      unsigned long long e1 = 0, e2 = 0;
      if (peer) { pc = a.comm->peer; e1 = ++a.comm->epoch; e2 = ++a.comm->epoch; }
      range_kernel<<<blocks, 256, 0, st>>>(T, p->d_chi2h.p, p->d_part.p, p->d_arrive.p, d_range, p->d_cells.p, nb, pc, e1, peer ? 1 : 0); ISR_COUNT_LAUNCH();
      if (a.comm && !peer && (rc = comm_allreduce(a.comm, d_range, 2, 0, 1)) != ISR_OK) return drain(rc);
      // the verdict of the inverse (side stream) enters the packed buffer only when ranks have to agree on a re-run; a
      // single GPU reads it from page-locked memory after the final synchronisation, so the certificate (which cannot
      // share the SMs with the persistent toy kernel and therefore runs when that ends) overlaps the tail kernels
      if (cert_here && a.comm) cudaStreamWaitEvent(st, c->ev_cert, 0);
      cells_pack_kernel<<<blocks, 256, sizeof(uint32_t) * nb, st>>>(T, p->d_chi2h.p, a.nbins, d_range, p->d_cells.p, p->d_arrive.p + 1, n,
                                                                    p->d_stat.p, (double)T, (cert_here && a.comm) ? p->cert_dev : nullptr, assumed,
                                                                    1e-12 * n, d_pack, pc, e2, peer ? 1 : 0,
                                                                    (peer || !a.comm) ? p->pin_test_dev : nullptr); ISR_COUNT_LAUNCH();
      packed_by_peer = peer;
      result_on_host = (peer || !a.comm) && p->pin_test_dev != nullptr;
    } else {
      if (need_range) {
        if (T > 0) { if ((rc = minmax_to_device(c, T, p->d_chi2h.p, d_range)) != ISR_OK) return drain(rc); }
        else { empty_range_kernel<<<1, 32, 0, st>>>(d_range); ISR_COUNT_LAUNCH(); }
        if (a.comm && (rc = comm_allreduce(a.comm, d_range, 2, 0, 1)) != ISR_OK) return drain(rc);
        if (a.nbins > 0 && (rc = histogram_range_device(c, T, p->d_chi2h.p, a.nbins, d_range, p->d_cells.p)) != ISR_OK) return drain(rc);
      }
      if (cert_here && a.comm) cudaStreamWaitEvent(st, c->ev_cert, 0);
      pack_test_kernel<<<(plen + 255) / 256, 256, 0, st>>>(nb, n, a.nbins > 0 ? p->d_cells.p : nullptr, p->d_stat.p, (double)T,
                                                          (cert_here && a.comm) ? p->cert_dev : nullptr, assumed, 1e-12 * n, d_pack); ISR_COUNT_LAUNCH();
    }
    if (a.comm) {
      if (!packed_by_peer && (rc = comm_allreduce(a.comm, d_pack, plen, 0, 0)) != ISR_OK) return drain(rc);
      if (w.rhist && (rc = comm_allreduce(a.comm, p->d_rhist.p, (int64_t)n * 1026, 2, 0)) != ISR_OK) return drain(rc);
    }
    if (a.comm && a.comm->peer_ok) cudaMemcpyAsync(a.comm->status_pin, a.comm->peer.status, sizeof(int), cudaMemcpyDeviceToHost, st);
    const bool copy_failed = !result_on_host &&
        cudaMemcpyAsync(p->pin_test, p->d_pack.p, sizeof(double) * ((size_t)plen + 2), cudaMemcpyDeviceToHost, st) != cudaSuccess;
    cudaEventRecord(p->ev_time[4], st);
    if (copy_failed || cudaStreamSynchronize(st) != cudaSuccess || (cert_here && cudaEventSynchronize(c->ev_cert) != cudaSuccess)) {
      set_error("isr_chi2_test: CUDA error '%s'", cudaGetErrorString(cudaGetLastError()));
      return drain(ISR_ECUDA);
    }
    if (a.comm && a.comm->peer_ok && *a.comm->status_pin != 0) {
      set_error("isr_chi2_test: a rank of the communicator did not take part in the exchange (peer mailbox timed out)");
      return drain(ISR_ECUDA);
    }
    bool redone = false;
    if (do_fac && (rc = factor_finish(p, &redone)) != ISR_OK) return drain(rc);
    const bool any_failed = p->pin_test[2 + nb + 4 * n + 1] > 0.0;
    if (!redone && !any_failed) break;
    if (attempt == 1) { set_error("isr_chi2_test: factorisation failed its certificate twice"); return drain(ISR_ENUMERIC); }
    refactored = 1;      // this rank, or another one of the communicator, fell back to the pivoted inverse: run the toys again
  }
  p->last_chi2_count = need_chi2 ? T : 0;

  // ---- results
  const double* pk = p->pin_test + 2;
  if (a.chi2_lo_hi_out) { a.chi2_lo_hi_out[0] = p->pin_test[0]; a.chi2_lo_hi_out[1] = -p->pin_test[1]; }
  if (a.chi2_counts_out) for (int q = 0; q < nb; ++q) a.chi2_counts_out[q] = (int64_t)pk[q];
  if (a.sum_bcs_out) std::memcpy(a.sum_bcs_out, pk + nb, sizeof(double) * n);
  if (a.ratio_stats_out) std::memcpy(a.ratio_stats_out, pk + nb + n, sizeof(double) * 3 * n);
  if (a.total_toys_out) *a.total_toys_out = (int64_t)pk[nb + 4 * n];
  if (a.refactored_out) *a.refactored_out = refactored;
  bool more = false;
  if (a.chi2_out && T > 0 && !a.V) {      // (host toys: copied chunk by chunk inside the pipeline)
    ISR_CUDA(cudaMemcpyAsync(a.chi2_out, p->d_chi2h.p, sizeof(double) * T, cudaMemcpyDeviceToHost, st));
    more = true;
  }
  if (a.ratio_hist_out && w.rhist) {
    ISR_CUDA(cudaMemcpyAsync(a.ratio_hist_out, p->d_rhist.p, sizeof(uint32_t) * (size_t)n * 1026, cudaMemcpyDeviceToHost, st));
    more = true;
  }
  if (more) ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

}  // namespace isr

extern "C" int isr_chi2_test_timing(isr_problem* p, double* ms_out) {
  if (!p || !ms_out) { isr::set_error("isr_chi2_test_timing: NULL argument"); return ISR_EINVAL; }
  if (!p->ev_time[0]) { isr::set_error("isr_chi2_test_timing: no isr_chi2_test has run on this problem"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  for (int q = 0; q < 4; ++q) {
    float ms = 0.f;
    ISR_CUDA(cudaEventElapsedTime(&ms, p->ev_time[q], p->ev_time[q + 1]));
    ms_out[q] = ms;
  }
  return ISR_OK;
}

extern "C" int isr_chi2_test(isr_problem* p, const isr_chi2_test_args* args) {
  if (!p || !args) { isr::set_error("isr_chi2_test: NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  isr::TestWants w;
  w.sum = args->sum_bcs_out != nullptr;
  w.rhist = args->ratio_hist_out != nullptr;
  w.range = args->chi2_lo_hi_out != nullptr;
  return isr::chi2_test(p, *args, w);
}
