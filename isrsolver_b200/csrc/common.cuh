// Shared declarations for libisr_b200.so (host + device).  sm_100a only, FP64 throughout.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include "../../include/isr_b200.h"

namespace isr {

// ---- error plumbing: no exceptions cross the C ABI ---------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define ISR_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      isr::set_error("CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(e__), __FILE__,     \
                     __LINE__, #expr);                                                       \
      return ISR_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define ISR_CUSOLVER(expr)                                                                   \
  do {                                                                                       \
    cusolverStatus_t s__ = (expr);                                                           \
    if (s__ != CUSOLVER_STATUS_SUCCESS) {                                                    \
      isr::set_error("cuSOLVER status %d at %s:%d (%s)", (int)s__, __FILE__, __LINE__, #expr); \
      return ISR_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define ISR_CUBLAS(expr)                                                                     \
  do {                                                                                       \
    cublasStatus_t s__ = (expr);                                                             \
    if (s__ != CUBLAS_STATUS_SUCCESS) {                                                      \
      isr::set_error("cuBLAS status %d at %s:%d (%s)", (int)s__, __FILE__, __LINE__, #expr); \
      return ISR_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define ISR_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != ISR_OK) return rc__; \
  } while (0)

// ---- launch accounting (bench.py reports it as gpu_launches) -----------------------------------------
extern std::atomic<long long> g_launch_count;   // problems on different host threads may launch concurrently
#define ISR_COUNT_LAUNCH() (isr::g_launch_count.fetch_add(1, std::memory_order_relaxed))

// ---- physical constants (include/PhysicalConstants.hpp:7,11) -------------------------------------
constexpr double kAlphaQED = 7.2973525698e-03;
constexpr double kElectronM = 5.10998928e-04;
constexpr double kPi = 3.14159265358979323846;

// ---- device buffer with RAII (host side only) ----------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  // grow-only allocation
  int reserve(size_t count) {
    if (count <= n && p) return ISR_OK;
    release();
    if (count == 0) count = 1;
    ISR_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
    return ISR_OK;
  }
};

constexpr int kMaxRowBreaks = 24;

// Per-row constants of the Kuraev-Fadin kernel and the row's quadrature grading points.
struct RowConst {
  double E, s, L, beta, inv_beta, C1, x0, xin, xmax;
  int nrb;
  double rb[kMaxRowBreaks];  // ascending: x0, x0(1+4^m), 1-(1-xmax)4^m
};

// Efficiency in the per-row form the device consumes.
struct EpsDev {
  int kind;             // 0: none, 1: per-row scale only, 2: per-row x table
  int nx;
  double xlo, xhi;
  const double* xedges; // device, nx+1 entries, or nullptr for fixed-width bins
  const double* values; // device: kind 1: [N]; kind 2: [N][nx+2]
};

}  // namespace isr

// Small page-locked staging area of a context: certificates, ranges, counters and other O(N) read-backs go through it,
// so that a cudaMemcpyAsync(DeviceToHost) really is asynchronous (into pageable memory it returns only after the copy).
constexpr size_t kPinDoubles = 1 << 16;   // 512 KB
// slots (doubles) of a problem's page-locked certificate block: {max|AX-I|, X upper != 0, A upper != 0}, cuSOLVER info (2 ints)
constexpr int kPinCert = 0, kPinInfo = 4;

struct isr_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux_stream = nullptr;     // side work that may overlap the main stream (first toy block, eigen-system)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_gen[2] = {nullptr, nullptr}, ev_slot[2] = {nullptr, nullptr};   // device toy generator: block drawn / slot consumed
  cudaEvent_t ev_fork2 = nullptr, ev_rop = nullptr;  // factorisation: fork to the side stream / chi2 factor ready there
  cudaStream_t gen_stream = nullptr;                 // toy generator of the first block (chi2_test)
  cudaEvent_t ev_inv = nullptr, ev_cert = nullptr;   // inverse enqueued (main stream) / its certificate complete (side stream)
  double* pin = nullptr;                 // kPinDoubles page-locked doubles (layout owned by the call that uses it)
  cusolverDnHandle_t solver = nullptr;
  cublasHandle_t blas = nullptr;
  // grow-only scratch for the reductions (min/max partials, histogram counters, staged chi2)
  isr::DevBuf<double> scratch_d;
  isr::DevBuf<uint32_t> scratch_u;
  isr::DevBuf<double> scratch_v;
  isr::DevBuf<double> scratch_t;   // N x N scratch of the blocked triangular inverse
  isr::DevBuf<double> scratch_batch_a, scratch_batch_b, scratch_batch_sig, scratch_batch_sv;   // batched spread scan
  isr::DevBuf<int> scratch_batch_i;
};
