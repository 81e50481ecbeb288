// extern "C" surface of libisr_b200.so (declared in include/isr_b200.h).
#include <cmath>
#include <cstdarg>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>

#include "comm.hpp"
#include "linalg.hpp"
#include "problem.hpp"
#include "solve.hpp"
#include "tikhonov.hpp"
#include "toy_batch.hpp"

namespace isr {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
std::atomic<long long> g_launch_count{0};
}  // namespace isr

using namespace isr;

extern "C" {

const char* isr_last_error(void) { return isr::get_error(); }

int64_t isr_launch_count(void) { return (int64_t)isr::g_launch_count.load(); }

int isr_ctx_create(int device, isr_ctx** out) {
  if (!out) { set_error("isr_ctx_create: out == NULL"); return ISR_EINVAL; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("isr_ctx_create: no CUDA device (%s); libisr_b200 has no CPU fallback", cudaGetErrorString(e));
    return ISR_ECUDA;
  }
  if (device < 0 || device >= count) { set_error("isr_ctx_create: device %d out of range [0,%d)", device, count); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ISR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("isr_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return ISR_ECUDA;
  }
  isr_ctx* c = new (std::nothrow) isr_ctx();
  if (!c) { set_error("out of host memory"); return ISR_EINVAL; }
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  ISR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  ISR_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
  ISR_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  ISR_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (int q = 0; q < 2; ++q) {
    ISR_CUDA(cudaEventCreateWithFlags(&c->ev_gen[q], cudaEventDisableTiming));
    ISR_CUDA(cudaEventCreateWithFlags(&c->ev_slot[q], cudaEventDisableTiming));
  }
  ISR_CUDA(cudaStreamCreateWithFlags(&c->gen_stream, cudaStreamNonBlocking));
  ISR_CUDA(cudaEventCreateWithFlags(&c->ev_fork2, cudaEventDisableTiming));
  ISR_CUDA(cudaEventCreateWithFlags(&c->ev_rop, cudaEventDisableTiming));
  ISR_CUDA(cudaEventCreateWithFlags(&c->ev_inv, cudaEventDisableTiming));
  ISR_CUDA(cudaEventCreateWithFlags(&c->ev_cert, cudaEventDisableTiming));
  ISR_CUDA(cudaMallocHost((void**)&c->pin, sizeof(double) * kPinDoubles));
  ISR_CUSOLVER(cusolverDnCreate(&c->solver));
  ISR_CUSOLVER(cusolverDnSetStream(c->solver, c->stream));
  ISR_CUBLAS(cublasCreate(&c->blas));
  ISR_CUBLAS(cublasSetStream(c->blas, c->stream));
  *out = c;
  return ISR_OK;
}

void isr_ctx_destroy(isr_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->blas) cublasDestroy(c->blas);
  if (c->solver) cusolverDnDestroy(c->solver);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (int q = 0; q < 2; ++q) {
    if (c->ev_gen[q]) cudaEventDestroy(c->ev_gen[q]);
    if (c->ev_slot[q]) cudaEventDestroy(c->ev_slot[q]);
  }
  if (c->gen_stream) cudaStreamDestroy(c->gen_stream);
  if (c->ev_fork2) cudaEventDestroy(c->ev_fork2);
  if (c->ev_rop) cudaEventDestroy(c->ev_rop);
  if (c->ev_inv) cudaEventDestroy(c->ev_inv);
  if (c->ev_cert) cudaEventDestroy(c->ev_cert);
  if (c->pin) cudaFreeHost(c->pin);
  delete c;
}

void* isr_ctx_stream(isr_ctx* c) { return c ? (void*)c->stream : nullptr; }

int isr_host_alloc(isr_ctx* c, size_t bytes, void** out) {
  if (!c || !out) { set_error("isr_host_alloc: NULL argument"); return ISR_EINVAL; }
  *out = nullptr;
  if (bytes == 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(c->device));
  // the pages are allocated and pinned by THIS thread: bound, for the duration of the call, to the CPUs of the NUMA node
  // the GPU hangs off, so that the buffer is local to the device that will read it (8 ranks streaming 1.7 GB each per
  // step otherwise cross the socket interconnect: e2e efficiency 0.42 at 8 GPUs in round 1)
  ScopedNodeBinding bind(c->device);
  ISR_CUDA(cudaMallocHost(out, bytes));
  std::memset(*out, 0, bytes);      // first touch
  return ISR_OK;
}
void isr_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

int isr_ctx_synchronize(isr_ctx* c) {
  if (!c) { set_error("ctx == NULL"); return ISR_EINVAL; }
  ISR_CUDA(cudaStreamSynchronize(c->stream));
  return ISR_OK;
}

int isr_problem_create(isr_ctx* ctx, int n, const double* ecm, double threshold, int nranges,
                       const int* ranges, const isr_eps_desc* eps, isr_problem** out) {
  if (!ctx || !ecm || !out) { set_error("isr_problem_create: NULL argument"); return ISR_EINVAL; }
  *out = nullptr;
  if (n < 2) { set_error("isr_problem_create: need at least 2 energy points"); return ISR_EINVAL; }
  if (!(ecm[0] > threshold)) { set_error("isr_problem_create: first energy %.17g is not above the threshold %.17g", ecm[0], threshold); return ISR_EINVAL; }
  for (int i = 1; i < n; ++i)
    if (!(ecm[i] > ecm[i - 1])) { set_error("isr_problem_create: energies must be strictly ascending (index %d)", i); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  isr_problem* p = new (std::nothrow) isr_problem();
  if (!p) { set_error("out of host memory"); return ISR_EINVAL; }
  p->ctx = ctx;
  p->n = n;
  p->threshold = threshold;
  p->ext.resize(n + 1);
  p->ext[0] = threshold;
  std::memcpy(p->ext.data() + 1, ecm, sizeof(double) * n);
  int rc = set_ranges(p, nranges, ranges);
  if (rc == ISR_OK) rc = upload_eps(p, eps);
  if (rc != ISR_OK) { delete p; return rc; }
  *out = p;
  return ISR_OK;
}

void isr_problem_destroy(isr_problem* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  delete p;
}

int isr_problem_set_ranges(isr_problem* p, int nranges, const int* ranges) {
  if (!p) { set_error("problem == NULL"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return set_ranges(p, nranges, ranges);
}

int isr_assemble(isr_problem* p, const double* ecm_err, double* G_out) {
  if (!p) { set_error("problem == NULL"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(assemble(p, ecm_err));
  if (G_out) ISR_CUDA(cudaMemcpyAsync(G_out, p->d_G.p, sizeof(double) * p->n * p->n, cudaMemcpyDeviceToHost, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_assemble_nodes(isr_problem* p, int64_t cap, double* x_out, int* row_out, int64_t* count) {
  if (!p) { set_error("problem == NULL"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return assemble_nodes(p, cap, x_out, row_out, count);
}

int isr_assemble_sampled(isr_problem* p, int64_t n_nodes, const double* eps_values, const double* ecm_err, double* G_out) {
  if (!p || (n_nodes > 0 && !eps_values)) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(assemble_sampled(p, n_nodes, eps_values, ecm_err));
  if (G_out) ISR_CUDA(cudaMemcpyAsync(G_out, p->d_G.p, sizeof(double) * p->n * p->n, cudaMemcpyDeviceToHost, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_set_matrix(isr_problem* p, const double* G) {
  if (!p || !G) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(p->d_G.reserve((size_t)p->n * p->n));
  ISR_CUDA(cudaMemcpyAsync(p->d_G.p, G, sizeof(double) * p->n * p->n, cudaMemcpyHostToDevice, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  p->has_G = true;
  p->G_injected = true;
  p->G_lower = true;
  for (int j = 1; j < p->n && p->G_lower; ++j)
    for (int i = 0; i < j; ++i)
      if (G[(size_t)i + (size_t)p->n * j] != 0.0) { p->G_lower = false; break; }
  p->factored = p->solver_selected = p->tik_ready = p->svd_ready = false;
  return ISR_OK;
}

int isr_assemble_stats(isr_problem* p, int64_t* n_panels, int64_t* n_nodes) {
  if (!p) { set_error("problem == NULL"); return ISR_EINVAL; }
  if (!p->has_G || !p->d_off.p) { set_error("isr_assemble_stats: no assembly yet"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  const int nitems = p->n * (p->n + 1) / 2;
  int total = 0;
  ISR_CUDA(cudaMemcpyAsync(&total, p->d_off.p + nitems, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  p->last_panels = total;
  if (n_panels) *n_panels = total;
  if (n_nodes) *n_nodes = (int64_t)total * 16;
  return ISR_OK;
}

int isr_debug_panels(isr_problem* p, int64_t cap, double* abek, int* row_seg, int64_t* count) {
  if (!p) { set_error("problem == NULL"); return ISR_EINVAL; }
  int64_t total = 0;
  ISR_TRY(isr_assemble_stats(p, &total, nullptr));
  if (count) *count = total;
  if (!abek && !row_seg) return ISR_OK;
  const int64_t m = total < cap ? total : cap;
  std::vector<double> a(m), b(m), e(m);
  std::vector<int> item(m), kind(m);
  cudaStream_t st = p->ctx->stream;
  ISR_CUDA(cudaMemcpyAsync(a.data(), p->d_pa.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaMemcpyAsync(b.data(), p->d_pb.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaMemcpyAsync(e.data(), p->d_peps.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaMemcpyAsync(item.data(), p->d_pitem.p, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaMemcpyAsync(kind.data(), p->d_pkind.p, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  for (int64_t q = 0; q < m; ++q) {
    if (abek) { abek[4 * q] = a[q]; abek[4 * q + 1] = b[q]; abek[4 * q + 2] = e[q]; abek[4 * q + 3] = kind[q]; }
    if (row_seg) {
      int i = (int)((std::sqrt(8.0 * item[q] + 1.0) - 1.0) * 0.5);
      while ((long long)(i + 1) * (i + 2) / 2 <= item[q]) ++i;
      while ((long long)i * (i + 1) / 2 > item[q]) --i;
      row_seg[2 * q] = i;
      row_seg[2 * q + 1] = item[q] - i * (i + 1) / 2;
    }
  }
  return ISR_OK;
}

int isr_debug_moments(isr_problem* p, double* M_out) {
  if (!p || !M_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (!p->has_G || p->G_injected || !p->d_pm.p) { set_error("isr_debug_moments: no assembly yet"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(reduce_moments(p));
  ISR_CUDA(cudaMemcpyAsync(M_out, p->d_M.p, sizeof(double) * p->n * p->n * 4, cudaMemcpyDeviceToHost, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_kernel_kf(isr_ctx* ctx, int64_t n, const double* x, const double* s, double* out) {
  if (!ctx || !x || !s || !out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (n <= 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> dx, ds, dout;
  ISR_TRY(dx.reserve(n)); ISR_TRY(ds.reserve(n)); ISR_TRY(dout.reserve(n));
  ISR_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  ISR_CUDA(cudaMemcpyAsync(ds.p, s, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  launch_kf_plain(n, dx.p, ds.p, dout.p, ctx->stream);
  ISR_CUDA(cudaGetLastError());
  ISR_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(ctx->stream));
  return ISR_OK;
}

int isr_basis_eval(isr_problem* p, int64_t m, const int* cs_index, const double* energy, int deriv, double* out) {
  if (!p || !cs_index || !energy || !out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (m <= 0) return ISR_OK;
  for (int64_t q = 0; q < m; ++q)
    if (cs_index[q] < 0 || cs_index[q] >= p->n) { set_error("isr_basis_eval: cs_index[%lld] out of range", (long long)q); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  DevBuf<int> dj; DevBuf<double> de, dout;
  ISR_TRY(dj.reserve(m)); ISR_TRY(de.reserve(m)); ISR_TRY(dout.reserve(m));
  ISR_CUDA(cudaMemcpyAsync(dj.p, cs_index, sizeof(int) * m, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(de.p, energy, sizeof(double) * m, cudaMemcpyHostToDevice, st));
  launch_basis_eval(p, m, dj.p, de.p, deriv, dout.p);
  ISR_CUDA(cudaGetLastError());
  ISR_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int isr_basis_integrals(isr_problem* p, double* w_out) {
  if (!p || !w_out) { set_error("NULL argument"); return ISR_EINVAL; }
  // exact integral of the per-segment polynomials: w_j = sum_k h_k sum_q C[j][k][q] / (q+1)
  // (Interpolator::evalIntegralBasis, src/Interpolator.cpp:459-473, integrates the same pieces with QAG)
  const int n = p->n;
  for (int j = 0; j < n; ++j) {
    double w = 0.0;
    for (int k = p->h_klo[j]; k <= p->h_khi[j]; ++k) {
      const double h = p->ext[k + 1] - p->ext[k];
      double s = 0.0;
      for (int q = 0; q < 4; ++q) s += p->h_Ct[((size_t)k * 4 + q) * n + j] / (q + 1);
      w += h * s;
    }
    w_out[j] = w;
  }
  return ISR_OK;
}

int isr_interp_eval(isr_problem* p, const double* y, int64_t m, const double* energy, double* out) {
  if (!p || !y || !energy || !out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (m <= 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  DevBuf<double> dy, de, dout;
  ISR_TRY(dy.reserve(p->n)); ISR_TRY(de.reserve(m)); ISR_TRY(dout.reserve(m));
  ISR_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(double) * p->n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(de.p, energy, sizeof(double) * m, cudaMemcpyHostToDevice, st));
  launch_interp_eval(p, dy.p, m, de.p, dout.p);
  ISR_CUDA(cudaGetLastError());
  ISR_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int isr_factor(isr_problem* p, const double* vcs_err) {
  if (!p || !vcs_err) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(factor(p, vcs_err));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_solve(isr_problem* p, const double* vcs, double* bcs_out, double* cov_out) {
  if (!p || !vcs || !bcs_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (!p->factored) { set_error("isr_solve: call isr_factor first"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  ISR_TRY(p->d_vin.reserve(n));
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  ISR_CUDA(cudaMemcpyAsync(p->d_vin.p, vcs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_TRY(gemv(p->ctx, false, n, p->d_Ginv.p, p->d_vin.p, p->d_stat.p));
  ISR_CUDA(cudaMemcpyAsync(bcs_out, p->d_stat.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (cov_out) {
    ISR_TRY(ensure_cov(p));
    ISR_CUDA(cudaMemcpyAsync(cov_out, p->d_Cov.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
  }
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int isr_iterative_solve(isr_problem* p, int niter, const double* vcs, double* bcs_out, double* radcorr_out) {
  if (!p || !vcs || !bcs_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return iterative_solve(p, niter, vcs, bcs_out, radcorr_out);
}

static int copy_out(isr_problem* p, const double* dev, double* host, const char* what) {
  if (!p || !host) { set_error("NULL argument"); return ISR_EINVAL; }
  if (!p->factored) { set_error("%s: call isr_factor first", what); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * p->n * p->n, cudaMemcpyDeviceToHost, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}
int isr_get_inverse(isr_problem* p, double* out) { return copy_out(p, p ? p->d_Ginv.p : nullptr, out, "isr_get_inverse"); }
int isr_get_inv_cov(isr_problem* p, double* out) {
  if (p && p->factored) {
    ISR_CUDA(cudaSetDevice(p->ctx->device));
    ISR_TRY(ensure_cov(p));
  }
  return copy_out(p, p ? p->d_InvCov.p : nullptr, out, "isr_get_inv_cov");
}

int isr_cond_number(isr_problem* p, double* cond_out) {
  if (!p || !cond_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(ensure_svd(p));
  *cond_out = p->h_sv.front() / p->h_sv.back();
  return ISR_OK;
}

// Condition number of G_spread(sigma_k) G for K beam-energy-spread settings: the loop of
// src/isrsolver-condnum-test.cpp:88-140 (resetECMErrors + evalEqMatrix + JacobiSVD per point; also
// notebooks/condnum_enspread.ipynb).  The reference re-integrates the whole matrix for every sigma; here the
// unspread matrix is assembled once and only the 6-point Gauss-Hermite spread product and the SVD are redone.
int isr_cond_spread_scan(isr_problem* p, int K, const double* sigma, int per_point, double* cond_out, double* sv_out) {
  if (!p || !sigma || !cond_out || K < 0) { set_error("isr_cond_spread_scan: bad argument"); return ISR_EINVAL; }
  if (K == 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  const int n = p->n;
  const size_t nn = (size_t)n * n;
  cudaStream_t st = p->ctx->stream;
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < (per_point ? n : 1); ++i) {
      const double v = sigma[per_point ? (size_t)k * n + i : k];
      if (!(v >= 0.0) || !std::isfinite(v) || (per_point && v == 0.0)) {
        set_error("isr_cond_spread_scan: sigma[%d] = %g (per-point spreads must be positive, common ones non-negative)", k, v);
        return ISR_EINVAL;
      }
    }
  if (!p->G_injected) ISR_TRY(assemble(p, nullptr));          // the unspread matrix
  else if (!p->has_G) { set_error("isr_cond_spread_scan: no matrix"); return ISR_ESTATE; }
  ISR_TRY(p->d_G0.reserve(nn));
  ISR_CUDA(cudaMemcpyAsync(p->d_G0.p, p->d_G.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
  p->G_lower = false;
  // The K settings are independent: ONE batched launch each for the 6-point Gauss-Hermite spread matrices, the products
  // S_k G0 and the singular values (one-sided Jacobi, one CTA per matrix: svd_jacobi.cu), in chunks of two waves of CTAs.
  if (n > 512) { set_error("isr_cond_spread_scan: N = %d exceeds the batched Jacobi kernel's 512", n); return ISR_EINVAL; }
  const int chunk = std::min(K, 2 * p->ctx->sm_count);
  // grow-only scratch of the context: a scan allocates its 2 x K x N^2 doubles once, not per call (cudaMalloc / cudaFree of
  // tens of MB cost more than the scan itself)
  DevBuf<double>&dS = p->ctx->scratch_batch_a, &dA = p->ctx->scratch_batch_b, &dsig = p->ctx->scratch_batch_sig, &dsv = p->ctx->scratch_batch_sv;
  DevBuf<int>& dsw = p->ctx->scratch_batch_i;
  const size_t per = per_point ? (size_t)n : 1;
  ISR_TRY(dS.reserve((size_t)chunk * nn)); ISR_TRY(dA.reserve((size_t)chunk * nn));
  ISR_TRY(dsig.reserve((size_t)K * per)); ISR_TRY(dsv.reserve((size_t)K * n)); ISR_TRY(dsw.reserve(K));
  ISR_CUDA(cudaMemcpyAsync(dsig.p, sigma, sizeof(double) * K * per, cudaMemcpyHostToDevice, st));
  for (int k0 = 0; k0 < K; k0 += chunk) {
    const int kc = std::min(chunk, K - k0);
    ISR_TRY(spread_batch(p, st, kc, dsig.p + (size_t)k0 * per, per_point, dS.p));
    ISR_TRY(small_gemm_batch(p->ctx, st, n, kc, dS.p, p->d_G0.p, dA.p));
    ISR_TRY(jacobi_singular_values_batch(p->ctx, st, n, kc, dA.p, dsv.p + (size_t)k0 * n, dsw.p + k0));
  }
  std::vector<double> h_sv((size_t)K * n);
  std::vector<int> h_sw(K);
  ISR_CUDA(cudaMemcpyAsync(h_sv.data(), dsv.p, sizeof(double) * K * n, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaMemcpyAsync(h_sw.data(), dsw.p, sizeof(int) * K, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  if (std::getenv("ISR_DEBUG_SWEEPS")) { for (int k = 0; k < K; ++k) fprintf(stderr, "%d ", h_sw[k]); fprintf(stderr, "<- Jacobi sweeps per setting\n"); }
  for (int k = 0; k < K; ++k) {
    if (h_sw[k] <= 0) { set_error("isr_cond_spread_scan: the Jacobi sweeps of setting %d did not converge (%d sweeps)", k, -h_sw[k]); return ISR_ENUMERIC; }
    cond_out[k] = h_sv[(size_t)k * n] / h_sv[(size_t)k * n + n - 1];
  }
  if (sv_out) std::memcpy(sv_out, h_sv.data(), sizeof(double) * K * n);
  p->svd_ready = false;    // d_sv no longer belongs to the problem's own matrix
  return ISR_OK;
}

const char* isr_toy_tile_name(int n, int nstat) {
  if (n < 1 || (nstat != 0 && nstat != 1 && nstat != 4)) return nullptr;
  return toy_cfg_name(n, nstat);
}

int isr_set_toy_kernel(isr_problem* p, int variant) {
  if (!p || variant < 0 || variant > 2) { set_error("isr_set_toy_kernel: bad argument"); return ISR_EINVAL; }
  p->toy_variant = variant;
  return ISR_OK;
}

int isr_toy_batch(isr_problem* p, int64_t T, const double* V, const double* bcs0, double* chi2_out,
                  double* sum_bcs_out, double* ratio_stats_out, uint32_t* ratio_hist_out, double* bcs_out) {
  if (!p || (T > 0 && (!V || !bcs0))) { set_error("NULL argument"); return ISR_EINVAL; }
  if (T < 0) { set_error("isr_toy_batch: T < 0"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  if (T == 0) {
    if (sum_bcs_out) std::memset(sum_bcs_out, 0, sizeof(double) * p->n);
    if (ratio_stats_out) std::memset(ratio_stats_out, 0, sizeof(double) * 3 * p->n);
    if (ratio_hist_out) std::memset(ratio_hist_out, 0, sizeof(uint32_t) * (size_t)p->n * 1026);
    return ISR_OK;
  }
  return toy_batch_host(p, T, V, bcs0, chi2_out, sum_bcs_out, ratio_stats_out, ratio_hist_out, bcs_out);
}

int isr_toy_batch_dev(isr_problem* p, int64_t T, const double* V_dev, const double* bcs0, double* chi2_dev,
                      double* sum_bcs_dev, double* ratio_stats_dev, uint32_t* ratio_hist_dev, double* bcs_dev) {
  if (!p || (T > 0 && (!V_dev || !bcs0))) { set_error("NULL argument"); return ISR_EINVAL; }
  if (T <= 0) return T == 0 ? ISR_OK : ISR_EINVAL;
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  if (p->h_b0.size() != (size_t)p->n || std::memcmp(p->h_b0.data(), bcs0, sizeof(double) * p->n) != 0) {
    ISR_TRY(p->d_b0.reserve(p->n));
    // a stream-ordered copy from a stable host mirror: earlier launches that still read d_b0 finish first
    p->h_b0.assign(bcs0, bcs0 + p->n);
    ISR_CUDA(cudaMemcpyAsync(p->d_b0.p, p->h_b0.data(), sizeof(double) * p->n, cudaMemcpyHostToDevice, p->ctx->stream));
    ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  }
  ToyOut o;
  o.chi2 = chi2_dev; o.sum_bcs = sum_bcs_dev; o.ratio_stats = ratio_stats_dev; o.rhist = ratio_hist_dev; o.bcs = bcs_dev;
  return toy_batch_device(p, T, V_dev, p->d_b0.p, o);
}

// ---- toy campaigns by count (toys drawn on the device) ----------------------------------------------
int isr_draw_toys(isr_ctx* ctx, int64_t T, int n, int64_t first_toy, uint64_t seed, const double* vcs, const double* vcs_err,
                  double* V_out) {
  if (!ctx || !vcs || !vcs_err || !V_out || n < 1 || T < 0) { set_error("isr_draw_toys: bad argument"); return ISR_EINVAL; }
  if (T == 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> dg, dV;
  ISR_TRY(dg.reserve((size_t)2 * n));
  ISR_TRY(dV.reserve((size_t)T * n));
  ISR_CUDA(cudaMemcpyAsync(dg.p, vcs, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  ISR_CUDA(cudaMemcpyAsync(dg.p + n, vcs_err, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  ISR_TRY(draw_toys_device(ctx, T, n, first_toy, seed, dg.p, dg.p + n, dV.p));
  ISR_CUDA(cudaMemcpyAsync(V_out, dV.p, sizeof(double) * T * n, cudaMemcpyDeviceToHost, ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(ctx->stream));
  return ISR_OK;
}

int isr_toy_campaign_dev(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* vcs, const double* vcs_err,
                         const double* bcs0, double* chi2_dev, double* sum_bcs_dev, double* ratio_stats_dev,
                         uint32_t* ratio_hist_dev) {
  if (!p || !vcs || !vcs_err || !bcs0 || T < 0) { set_error("isr_toy_campaign_dev: bad argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return toy_campaign_device(p, T, first_toy, seed, vcs, vcs_err, bcs0, chi2_dev, sum_bcs_dev, ratio_stats_dev, ratio_hist_dev);
}

int isr_toy_campaign(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* vcs, const double* vcs_err,
                     const double* bcs0, double* chi2_out, double* sum_bcs_out, double* ratio_stats_out,
                     uint32_t* ratio_hist_out, int nbins, double* chi2_lo_hi_out, uint32_t* chi2_counts_out) {
  if (!p || !vcs || !vcs_err || !bcs0 || T < 0) { set_error("isr_toy_campaign: bad argument"); return ISR_EINVAL; }
  if ((chi2_lo_hi_out || chi2_counts_out) && (nbins < 1 || nbins > 8190)) { set_error("isr_toy_campaign: nbins out of range"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  if (T == 0) {
    if (sum_bcs_out) std::memset(sum_bcs_out, 0, sizeof(double) * n);
    if (ratio_stats_out) std::memset(ratio_stats_out, 0, sizeof(double) * 3 * n);
    if (ratio_hist_out) std::memset(ratio_hist_out, 0, sizeof(uint32_t) * (size_t)n * 1026);
    return ISR_OK;
  }
  const bool want_chi2 = chi2_out || chi2_lo_hi_out || chi2_counts_out;
  if (want_chi2) ISR_TRY(p->d_chi2h.reserve((size_t)T));
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  ISR_CUDA(cudaMemsetAsync(p->d_stat.p, 0, sizeof(double) * 4 * n, st));
  if (ratio_hist_out) {
    ISR_TRY(p->d_rhist.reserve((size_t)n * 1026));
    ISR_CUDA(cudaMemsetAsync(p->d_rhist.p, 0, sizeof(uint32_t) * (size_t)n * 1026, st));
  }
  p->last_chi2_count = want_chi2 ? T : 0;
  ISR_TRY(toy_campaign_device(p, T, first_toy, seed, vcs, vcs_err, bcs0, want_chi2 ? p->d_chi2h.p : nullptr,
                              sum_bcs_out ? p->d_stat.p : nullptr, ratio_stats_out ? p->d_stat.p + n : nullptr,
                              ratio_hist_out ? p->d_rhist.p : nullptr));
  if (chi2_out) ISR_CUDA(cudaMemcpyAsync(chi2_out, p->d_chi2h.p, sizeof(double) * T, cudaMemcpyDeviceToHost, st));
  if (sum_bcs_out) ISR_CUDA(cudaMemcpyAsync(sum_bcs_out, p->d_stat.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (ratio_stats_out) ISR_CUDA(cudaMemcpyAsync(ratio_stats_out, p->d_stat.p + n, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
  if (ratio_hist_out) ISR_CUDA(cudaMemcpyAsync(ratio_hist_out, p->d_rhist.p, sizeof(uint32_t) * (size_t)n * 1026, cudaMemcpyDeviceToHost, st));
  if (chi2_lo_hi_out || chi2_counts_out) {
    double lo = 0, hi = 0;
    ISR_TRY(minmax_device(p->ctx, T, p->d_chi2h.p, &lo, &hi));
    if (chi2_lo_hi_out) { chi2_lo_hi_out[0] = lo; chi2_lo_hi_out[1] = hi; }
    if (chi2_counts_out) ISR_TRY(histogram_device(p->ctx, T, p->d_chi2h.p, nbins, lo, hi, chi2_counts_out));
  }
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

// TH1F(nbins, lo, hi) of the chi2 values the LAST isr_toy_campaign / isr_toy_batch (host API) of this problem left
// on the device -- the multi-GPU re-binning against the global range, without moving chi2.
int isr_last_chi2_histogram(isr_problem* p, int nbins, double lo, double hi, uint32_t* counts_out) {
  if (!p || !counts_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (p->last_chi2_count <= 0 || !p->d_chi2h.p) { set_error("isr_last_chi2_histogram: no chi2 values on the device"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return histogram_device(p->ctx, p->last_chi2_count, p->d_chi2h.p, nbins, lo, hi, counts_out);
}

// min / max of the same device-resident chi2 values: with isr_last_chi2_histogram the whole chi2 test of a host-fed
// batch is summarised without uploading the chi2 array again.
int isr_last_chi2_minmax(isr_problem* p, double* min_out, double* max_out) {
  if (!p || !min_out || !max_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (p->last_chi2_count <= 0 || !p->d_chi2h.p) { set_error("isr_last_chi2_minmax: no chi2 values on the device"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return minmax_device(p->ctx, p->last_chi2_count, p->d_chi2h.p, min_out, max_out);
}

int isr_minmax_dev(isr_ctx* ctx, int64_t n, const double* v_dev, double* min_out, double* max_out) {
  if (!ctx || !v_dev || !min_out || !max_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  return minmax_device(ctx, n, v_dev, min_out, max_out);
}

int isr_minmax_to_dev(isr_ctx* ctx, int64_t n, const double* v_dev, double* out_dev) {
  if (!ctx || !v_dev || !out_dev) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  return minmax_to_device(ctx, n, v_dev, out_dev);
}

int isr_histogram_range_dev(isr_ctx* ctx, int64_t n, const double* v_dev, int nbins, const double* range_dev, int32_t* counts_dev) {
  if (!ctx || (n > 0 && !v_dev) || !range_dev || !counts_dev) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  return histogram_range_device(ctx, n, v_dev, nbins, range_dev, counts_dev);
}

int isr_histogram_dev(isr_ctx* ctx, int64_t n, const double* v_dev, int nbins, double lo, double hi, uint32_t* counts_out) {
  if (!ctx || (n > 0 && !v_dev) || !counts_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  return histogram_device(ctx, n, v_dev, nbins, lo, hi, counts_out);
}

int isr_histogram(isr_ctx* ctx, int64_t n, const double* v, int nbins, double lo, double hi, uint32_t* counts_out) {
  if (!ctx || (n > 0 && !v) || !counts_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double>& d = ctx->scratch_v;
  ISR_TRY(d.reserve(n > 0 ? n : 1));
  if (n > 0) ISR_CUDA(cudaMemcpyAsync(d.p, v, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  return histogram_device(ctx, n, d.p, nbins, lo, hi, counts_out);
}

int isr_chi2_histogram(isr_ctx* ctx, int64_t n, const double* chi2, int nbins, double* lo_out, double* hi_out, uint32_t* counts_out) {
  if (!ctx || !chi2 || !lo_out || !hi_out || !counts_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (n <= 0) { set_error("isr_chi2_histogram: empty input"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double>& d = ctx->scratch_v;
  ISR_TRY(d.reserve(n));
  ISR_CUDA(cudaMemcpyAsync(d.p, chi2, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  ISR_TRY(minmax_device(ctx, n, d.p, lo_out, hi_out));
  return histogram_device(ctx, n, d.p, nbins, *lo_out, *hi_out, counts_out);
}

int isr_tikhonov_prepare(isr_problem* p, const double* vcs_err, int deriv_reg) {
  if (!p || !vcs_err) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return tikhonov_prepare(p, vcs_err, deriv_reg);
}

int isr_tikhonov_scan(isr_problem* p, int64_t L, const double* lambda, const double* vcs, double* eq_norm2_out,
                      double* smooth_norm2_out, double* curv_out, double* dcurv_out, double* bcs_out) {
  if (!p || (L > 0 && (!lambda || !vcs))) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return tikhonov_scan(p, L, lambda, vcs, eq_norm2_out, smooth_norm2_out, curv_out, dcurv_out, bcs_out);
}

int isr_tikhonov_solve(isr_problem* p, double lambda, const double* vcs, double* bcs_out, double* cov_out) {
  if (!p || !vcs || !bcs_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return tikhonov_solve(p, lambda, vcs, bcs_out, cov_out);
}

int isr_tikhonov_toy_scan_dev(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V_dev,
                              const double* bcs0, double* chi2_dev) {
  if (!p || !lambda || !V_dev || !bcs0 || !chi2_dev) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(tikhonov_toy_scan_device(p, L, lambda, T, V_dev, bcs0, chi2_dev));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_tikhonov_toy_scan(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V,
                          const double* bcs0, double* chi2_out) {
  if (!p || !lambda || !V || !bcs0 || !chi2_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (L <= 0 || T <= 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  DevBuf<double>&dV = p->d_scan_V, &dchi = p->d_scan_chi;
  ISR_TRY(dV.reserve((size_t)T * p->n));
  ISR_TRY(dchi.reserve((size_t)L * T));
  ISR_CUDA(cudaMemcpyAsync(dV.p, V, sizeof(double) * T * p->n, cudaMemcpyHostToDevice, st));
  ISR_TRY(tikhonov_toy_scan_device(p, L, lambda, T, dV.p, bcs0, dchi.p));
  ISR_CUDA(cudaMemcpyAsync(chi2_out, dchi.p, sizeof(double) * L * T, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int isr_tikhonov_toy_scan_hist(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V,
                               const double* bcs0, int nbins, double* stats_out, uint32_t* counts_out) {
  if (!p || !lambda || !V || !bcs0) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return tikhonov_toy_scan_hist(p, L, lambda, T, V, bcs0, nbins, stats_out, counts_out);
}

int isr_tikhonov_chi2_scan(isr_problem* p, const isr_tikhonov_scan_args* args) {
  if (!p || !args) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return tikhonov_chi2_scan(p, *args);
}

int isr_set_scan_kernel(isr_problem* p, int variant) {
  if (!p || variant < 0 || variant > 2) { set_error("isr_set_scan_kernel: bad argument"); return ISR_EINVAL; }
  p->scan_variant = variant;
  return ISR_OK;
}

int isr_tikhonov_select(isr_problem* p, double lambda) {
  if (!p) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(tikhonov_select(p, lambda));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_get_regulariser(isr_problem* p, double* F_out) {
  if (!p || !F_out) { set_error("NULL argument"); return ISR_EINVAL; }
  if (!p->tik_ready) { set_error("isr_get_regulariser: call isr_tikhonov_prepare first"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_CUDA(cudaMemcpyAsync(F_out, p->d_F.p, sizeof(double) * p->n * p->n, cudaMemcpyDeviceToHost, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

int isr_tsvd_solve(isr_problem* p, int k, int keep_one, const double* vcs, const double* vcs_err, double* bcs_out,
                   double* cov_out) {
  if (!p || !vcs || !bcs_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  return tsvd_solve(p, k, keep_one, vcs, vcs_err, bcs_out, cov_out);
}

int isr_singular_values(isr_problem* p, double* sv_out) {
  if (!p || !sv_out) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(ensure_svd(p));
  std::memcpy(sv_out, p->h_sv.data(), sizeof(double) * p->n);
  return ISR_OK;
}

int isr_tsvd_select(isr_problem* p, int k, int keep_one, const double* vcs_err) {
  if (!p || !vcs_err) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(tsvd_select(p, k, keep_one, vcs_err));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return ISR_OK;
}

}  // extern "C"
