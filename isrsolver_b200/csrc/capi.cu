// extern "C" surface of libisr_b200.so (declared in include/isr_b200.h).
#include <cstdarg>
#include <cstring>
#include <new>

#include "problem.hpp"

namespace isr {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace isr

using namespace isr;

extern "C" {

const char* isr_last_error(void) { return isr::get_error(); }

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
  delete c;
}

void* isr_ctx_stream(isr_ctx* c) { return c ? (void*)c->stream : nullptr; }

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

int isr_set_matrix(isr_problem* p, const double* G) {
  if (!p || !G) { set_error("NULL argument"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
  ISR_TRY(p->d_G.reserve((size_t)p->n * p->n));
  ISR_CUDA(cudaMemcpyAsync(p->d_G.p, G, sizeof(double) * p->n * p->n, cudaMemcpyHostToDevice, p->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(p->ctx->stream));
  p->has_G = true;
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
  if (!p->has_G || !p->d_M.p) { set_error("isr_debug_moments: no assembly yet"); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(p->ctx->device));
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

}  // extern "C"
