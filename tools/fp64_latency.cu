// FP64 dependent-issue latency and single-warp throughput on sm_100a (B200): the numbers the critical chain of the cluster
// Gauss-Jordan inverse (linalg.cu) is made of.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_latency.cu -o tools/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, long long* cyc, double m, int iters) {
  double a[ILP];
  for (int q = 0; q < ILP; ++q) a[q] = threadIdx.x + q;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int q = 0; q < ILP; ++q) a[q] = fma(-m, a[q], a[q]);
  long long t1 = clock64();
  double s = 0; for (int q = 0; q < ILP; ++q) s += a[q];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_chain(double* out, long long* cyc, int iters) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31) + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput of independent double shuffles (2 SHFL.IDX each) with a run-time source lane, per SM
__global__ void shfl_tput(double* out, long long* cyc, int iters, int src) {
  double a[8];
  for (int q = 0; q < 8; ++q) a[q] = threadIdx.x + q;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = __shfl_sync(0xffffffffu, a[q], (src + q + it) & 31);
  long long t1 = clock64();
  double s = 0; for (int q = 0; q < 8; ++q) s += a[q];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// LDS.128 throughput: every lane reads its own 16 bytes (512 B per warp instruction), 4 independent loads per step
__global__ void lds_tput(double* out, long long* cyc, int iters) {
  __shared__ double2 buf[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
  __syncthreads();
  double s = 0;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { const double2 v = buf[(lane + 32 * q + 128 * (it & 15)) & 2047]; s += v.x; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  long long h[8]; const int iters = 4096;
#define RUN(ILP, THREADS) chain<ILP><<<1, THREADS>>>(out, cyc, 1e-9, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DFMA ILP %2d, %4d threads/CTA: %.2f cycles per dependent step, %.2f cycles per warp-instruction\n", ILP, THREADS, (double)h[0] / iters, (double)h[0] / iters / ILP);
  RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(16, 32) RUN(32, 32)
  RUN(8, 128) RUN(8, 256) RUN(32, 128) RUN(32, 256)
  shfl_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("double shuffle + DADD dependent step: %.2f cycles\n", (double)h[0] / iters);
  for (int th : {32, 128, 256, 512}) {
    shfl_tput<<<1, th>>>(out, cyc, iters, 3); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("double shuffles, run-time lane, %3d threads/SM: %.2f cycles per double shuffle per warp, %.2f per SM\n", th, (double)h[0] / iters / 8, (double)h[0] / iters / 8 / (th / 32));
  }
  for (int th : {32, 128, 256}) {
    lds_tput<<<1, th>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.128 (512 B per warp), %3d threads/SM: %.2f cycles per load per warp, %.2f per SM\n", th, (double)h[0] / iters / 4, (double)h[0] / iters / 4 / (th / 32));
  }
  return 0;
}
