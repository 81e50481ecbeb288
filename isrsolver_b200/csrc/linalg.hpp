// Declarations for linalg.cu (one-off N x N library calls, column-major).
#pragma once
#include "common.cuh"

#include <functional>

namespace isr {
void colmajor_to_rowmajor(cudaStream_t st, int n, const double* A, const double* rowscale, int invert_scale,
                          double* out, int ld);
void scale_cols(cudaStream_t st, int n, const double* A, const double* s, int invert, double* B);
void scale_rows(cudaStream_t st, int n, const double* A, const double* s, int invert, double* B);
int gemm(isr_ctx* c, bool ta, bool tb, int n, const double* A, const double* B, double* C);
int gemv(isr_ctx* c, bool ta, int n, const double* A, const double* x, double* y);
int inverse_enqueue(isr_ctx* c, int n, const double* A, double* lu, double* prod, double* Ainv, DevBuf<double>& work,
                    DevBuf<int>& ipiv, DevBuf<int>& info, bool lower_hint, double* cert_host, double* cert_dev,
                    double* Ainv_rowmajor = nullptr, int ld = 0, bool* wrote_rowmajor = nullptr);
bool inverse_certified(const double* cert_host, int n, int* tri);
int inverse_pivoted(isr_ctx* c, int n, const double* A, double* lu, double* Ainv, DevBuf<double>& work, DevBuf<int>& ipiv,
                    DevBuf<int>& info, double* cert_host, double* cert_dev, int* tri);
// X = L^-1 for a lower-triangular L (column-major, N x N), blocked: diagonal blocks by substitution, then the
// off-diagonal block columns by products (tri_inverse.cu).  Entries above the diagonal of X are exact zeros.
int tri_inverse_lower(isr_ctx* c, int n, const double* L, double* X);
// C = A B, n x n column-major, hand-written 32 x 32 tiles; lower: both factors (hence the product) are lower triangular.
// slices_out != nullptr: K is split into *slices_out slices whose partial products lie n^2 apart in C (the caller adds
// them); C then holds small_gemm_slices(c, n) * n^2 doubles.
int small_gemm(isr_ctx* c, int n, bool lower, const double* A, const double* B, double* C, cudaStream_t st = nullptr,
               int* slices_out = nullptr);
int small_gemm_slices(isr_ctx* c, int n);
// C_k = A_k B for k < K (n x n column-major; A_k, C_k back to back with stride n^2, B shared): one launch
int small_gemm_batch(isr_ctx* c, cudaStream_t st, int n, int K, const double* A_batch, const double* B, double* C_batch);
// singular values (descending) of K n x n matrices by one-sided Jacobi, one CTA per matrix (svd_jacobi.cu); A destroyed
int jacobi_singular_values_batch(isr_ctx* c, cudaStream_t st, int n, int K, double* A_batch, double* sv_batch, int* sweeps_dev);
int ql_factor_rowmajor(isr_ctx* c, int n, int ld, double* A, double* scratch, DevBuf<double>& work, DevBuf<int>& info);
int svd(isr_ctx* c, int n, double* A, double* U, double* S, double* VT, DevBuf<double>& work, DevBuf<int>& info);
int svd_values_async(cusolverDnHandle_t h, int n, double* A, double* S, DevBuf<double>& work, DevBuf<int>& info);
int sygvd(isr_ctx* c, int n, double* A, double* B, double* W, DevBuf<double>& work, DevBuf<int>& info);
}  // namespace isr
