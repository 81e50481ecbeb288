// Can a small kernel's blocks become resident on an SM that a persistent, large-shared-memory kernel occupies?
// A: 148 CTAs x THREADS_A threads, SMEM_A bytes of dynamic shared memory, REGS via launch bounds; spins for ~SPIN_US.
// B: many blocks of 64 threads launched on a second stream while A runs; each records (smid, start, end) by globaltimer.
// Prints how many B blocks started while A was running.   nvcc -O2 -arch=sm_100a -o coresident_probe coresident_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) kernelA(unsigned long long spin_ns, unsigned long long* win, double* sink, int regs_pad) {
  extern __shared__ double sm[];
  const unsigned long long t0 = gtime();
  double acc[40];
#pragma unroll
  for (int i = 0; i < 40; ++i) acc[i] = threadIdx.x + i;
  while (gtime() - t0 < spin_ns) {
#pragma unroll
    for (int i = 0; i < 40; ++i) acc[i] = fma(acc[i], 1.0000001, 0.5);
    sm[threadIdx.x] = acc[regs_pad & 31];
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 40; ++i) s += acc[i];
  if (s == 12345.678) sink[0] = s;
  if (threadIdx.x == 0) { win[2 * blockIdx.x] = t0; win[2 * blockIdx.x + 1] = gtime(); }
}
__global__ void __launch_bounds__(64, 32) kernelB(unsigned long long* rec, float* sink) {
  const unsigned long long t0 = gtime();
  float x = threadIdx.x;
  for (int i = 0; i < 2000; ++i) x = fmaf(x, 1.0001f, 0.25f);
  if (x == 1234.5f) sink[0] = x;
  if (threadIdx.x == 0) { rec[3 * blockIdx.x] = smid(); rec[3 * blockIdx.x + 1] = t0; rec[3 * blockIdx.x + 2] = gtime(); }
}
template <int THREADS>
void run(int smem_a, int prio) {
  cudaStream_t sa, sb;
  int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, prio ? hi : lo);
  cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo);
  const int nb = 148 * 8;
  unsigned long long *win, *rec; double* sink; float* sinkf;
  cudaMalloc(&win, 2 * 148 * 8); cudaMalloc(&rec, 3 * nb * 8); cudaMalloc(&sink, 8); cudaMalloc(&sinkf, 4);
  cudaFuncSetAttribute(kernelA<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kernelA<THREADS>);
  kernelA<THREADS><<<148, THREADS, smem_a, sa>>>(2000000ull, win, sink, 3);
  kernelB<<<nb, 64, 0, sb>>>(rec, sinkf);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> w(2 * 148), r(3 * nb);
  cudaMemcpy(w.data(), win, w.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r.data(), rec, r.size() * 8, cudaMemcpyDeviceToHost);
  unsigned long long a0 = ~0ull, a1 = 0;
  for (int i = 0; i < 148; ++i) { if (w[2 * i] < a0) a0 = w[2 * i]; if (w[2 * i + 1] > a1) a1 = w[2 * i + 1]; }
  int inside = 0, before = 0, after = 0;
  for (int i = 0; i < nb; ++i) { const unsigned long long s = r[3 * i + 1]; if (s < a0) ++before; else if (s > a1 - 100000) ++after; else ++inside; }
  printf("A: %d threads x %d regs, %d KB smem, prio %s (%s): B blocks started before A %d, while A ran %d, after A %d\n", THREADS, fa.numRegs,
         smem_a / 1024, prio ? "high" : "same", cudaGetErrorString(e), before, inside, after);
}
int main() {
  run<640>(170 * 1024, 1);
  run<640>(170 * 1024, 0);
  run<640>(100 * 1024, 1);
  run<512>(208 * 1024, 1);
  run<256>(100 * 1024, 1);
  return 0;
}
