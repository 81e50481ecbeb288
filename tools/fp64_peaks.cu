// FP64 peak probes for B200 (sm_100a): DFMA vector pipe, DMMA (mma.sync f64) tensor pipe and
// cuBLAS DGEMM.  MEASURED_PEAKS.json carries only HBM and bf16 peaks; the toy-batch kernel is
// FP64-compute bound, so its roofline denominator is measured here.
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/fp64_peaks.cu -lcublas -o tools/fp64_peaks
// run  : tools/fp64_peaks > gpurun_out/fp64_peaks.json
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- DFMA
template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- DMMA m8n8k4
template <int ILP>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters) {
  double c0[ILP], c1[ILP];
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- DMMA m16n8k8 (sm_90+ shapes)
template <int ILP>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters) {
  double c[ILP][4];
  double a0 = threadIdx.x * 1e-6, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double b0 = 1.0 + threadIdx.x * 1e-7, b1 = b0 * 0.5;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- DMMA m16n8k16
template <int ILP>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters) {
  double c[ILP][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-6 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-7 + i;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F&& launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1 << 22));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

  for (int wpsm = 8; wpsm <= 32; wpsm *= 2) {   // resident warps per SM
    int blocks = sms * wpsm / 8;
    {
      constexpr int ILP = 16;
      float ms = time_ms([&] { dfma_kernel<ILP><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * ILP * iters * 256.0 * blocks;
      printf(", \"dfma_tflops_w%d\": %.3f", wpsm, fl / ms * 1e-9);
    }
    {
      constexpr int ILP = 8;
      float ms = time_ms([&] { dmma884_kernel<ILP><<<blocks, 256>>>(out, iters); }, 5);
      double fl = 2.0 * 8 * 8 * 4 * ILP * (double)iters * 8.0 * blocks;
      printf(", \"dmma_m8n8k4_tflops_w%d\": %.3f", wpsm, fl / ms * 1e-9);
    }
    {
      constexpr int ILP = 8;
      float ms = time_ms([&] { dmma1688_kernel<ILP><<<blocks, 256>>>(out, iters); }, 5);
      double fl = 2.0 * 16 * 8 * 8 * ILP * (double)iters * 8.0 * blocks;
      printf(", \"dmma_m16n8k8_tflops_w%d\": %.3f", wpsm, fl / ms * 1e-9);
    }
    {
      constexpr int ILP = 8;
      float ms = time_ms([&] { dmma16816_kernel<ILP><<<blocks, 256>>>(out, iters); }, 5);
      double fl = 2.0 * 16 * 8 * 16 * ILP * (double)iters * 8.0 * blocks;
      printf(", \"dmma_m16n8k16_tflops_w%d\": %.3f", wpsm, fl / ms * 1e-9);
    }
  }

  // cuBLAS DGEMM: square and the toy-batch shape (T x N x N)
  cublasHandle_t h; cublasCreate(&h);
  {
    const int n = 4096;
    double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
    CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
    const double one = 1, zero = 0;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
    printf(", \"cublas_dgemm_4096_tflops\": %.3f", 2.0 * n * n * (double)n / ms * 1e-9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  for (int n : {100, 200, 500}) {
    const int T = 1 << 18;
    double *S, *V, *Bm;
    CK(cudaMalloc(&S, sizeof(double) * n * n)); CK(cudaMalloc(&V, sizeof(double) * (size_t)n * T)); CK(cudaMalloc(&Bm, sizeof(double) * (size_t)n * T));
    CK(cudaMemset(S, 0, sizeof(double) * n * n)); CK(cudaMemset(V, 0, sizeof(double) * (size_t)n * T));
    const double one = 1, zero = 0;
    // B (n x T, col-major; toy k is column k) = S (n x n) * V (n x T)
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, T, n, &one, S, n, V, n, &zero, Bm, n); }, 5);
    printf(", \"cublas_dgemm_toyshape_n%d_tflops\": %.3f", n, 2.0 * n * n * (double)T / ms * 1e-9);
    cudaFree(S); cudaFree(V); cudaFree(Bm);
  }
  // HBM copy for reference
  {
    size_t bytes = (size_t)1 << 31;
    char *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    float ms = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 5);
    printf(", \"d2d_copy_gbs\": %.1f", 2.0 * bytes / ms * 1e-6);
    cudaFree(a); cudaFree(b);
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf(", \"sm_clock_khz_attr\": %d}\n", clk);
  return 0;
}
