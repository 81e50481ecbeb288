// Probe: time to invert one N x N FP64 matrix on the device -- cuSOLVER getrf+getrs (what isr_factor uses) against
// cuBLAS getrfBatched+getriBatched with batch = 1.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/inv_probe.cu -lcusolver -lcublas -o tools/inv_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do { auto e_ = (x); if (e_ != 0) { fprintf(stderr, "error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); exit(1);} } while (0)
int main() {
  cusolverDnHandle_t sh; cublasHandle_t bh; cudaStream_t st;
  CK(cudaStreamCreate(&st)); CK(cusolverDnCreate(&sh)); CK(cusolverDnSetStream(sh, st)); CK(cublasCreate(&bh)); CK(cublasSetStream(bh, st));
  for (int n : {50, 100, 200, 500}) {
    std::vector<double> h((size_t)n * n);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) h[i + (size_t)n * j] = (i >= j ? 0.5 * exp(-0.05 * (i - j)) : 0.01 * sin(i + 2.0 * j)) + (i == j);
    double *A, *B, *C, *work; int *ipiv, *info; double **Ap, **Cp;
    CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
    CK(cudaMalloc(&ipiv, sizeof(int) * n)); CK(cudaMalloc(&info, sizeof(int))); CK(cudaMalloc(&Ap, sizeof(double*))); CK(cudaMalloc(&Cp, sizeof(double*)));
    CK(cudaMemcpy(Ap, &B, sizeof(double*), cudaMemcpyHostToDevice)); CK(cudaMemcpy(Cp, &C, sizeof(double*), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(A, h.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    int lwork = 0; CK(cusolverDnDgetrf_bufferSize(sh, n, n, B, n, &lwork)); CK(cudaMalloc(&work, sizeof(double) * lwork));
    std::vector<double> eye((size_t)n * n, 0.0); for (int i = 0; i < n; ++i) eye[i + (size_t)n * i] = 1.0;
    double* I; CK(cudaMalloc(&I, sizeof(double) * n * n)); CK(cudaMemcpy(I, eye.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms1 = 0, ms2 = 0;
    for (int rep = 0; rep < 12; ++rep) {
      if (rep == 2) cudaEventRecord(e0, st);
      CK(cudaMemcpyAsync(B, A, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(C, I, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
      CK(cusolverDnDgetrf(sh, n, n, B, n, work, ipiv, info));
      CK(cusolverDnDgetrs(sh, CUBLAS_OP_N, n, n, B, n, ipiv, C, n, info));
    }
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1); ms1 /= 10;
    std::vector<double> r1((size_t)n * n); CK(cudaMemcpy(r1.data(), C, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    for (int rep = 0; rep < 12; ++rep) {
      if (rep == 2) cudaEventRecord(e0, st);
      CK(cudaMemcpyAsync(B, A, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
      CK(cublasDgetrfBatched(bh, n, Ap, n, ipiv, info, 1));
      CK(cublasDgetriBatched(bh, n, (const double* const*)Ap, n, ipiv, Cp, n, info, 1));
    }
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms2, e0, e1); ms2 /= 10;
    std::vector<double> r2((size_t)n * n); CK(cudaMemcpy(r2.data(), C, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    float ms3 = 0;
    for (int rep = 0; rep < 12; ++rep) {
      if (rep == 2) cudaEventRecord(e0, st);
      CK(cudaMemcpyAsync(B, A, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(C, I, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
      CK(cusolverDnDgetrf(sh, n, n, B, n, work, nullptr, info));                 // no pivoting
      CK(cusolverDnDgetrs(sh, CUBLAS_OP_N, n, n, B, n, nullptr, C, n, info));
    }
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms3, e0, e1); ms3 /= 10;
    std::vector<double> r3((size_t)n * n); CK(cudaMemcpy(r3.data(), C, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    double d3 = 0; for (size_t q = 0; q < r1.size(); ++q) d3 = fmax(d3, fabs(r1[q] - r3[q]));
    printf("n=%d  cusolver getrf+getrs WITHOUT pivoting %.3f ms  max|diff| %.2e\n", n, ms3, d3);
    double d = 0, s = 0; for (size_t q = 0; q < r1.size(); ++q) { d = fmax(d, fabs(r1[q] - r2[q])); s = fmax(s, fabs(r1[q])); }
    printf("n=%d  cusolver getrf+getrs %.3f ms   cublas getrfBatched+getriBatched %.3f ms   max|diff|/max %.2e\n", n, ms1, ms2, d / s);
  }
  return 0;
}
