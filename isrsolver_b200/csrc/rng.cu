// Device-side toy generator: V[k][i] = vcs[i] + vcs_err[i] * z(seed, k, i),  z ~ N(0, 1).
//
// Stands in for randomDrawVisCS (src/Utils.cpp:6-13: gRandom->Gaus(vcs_i, vcsErr_i) per point and per toy) when
// the caller asks for a toy campaign by COUNT, as the reference's chi2Test / ratioTestModel signatures do
// (src/Chi2Test.cpp:29-60, src/RatioTest.cpp:13-45), instead of supplying the toy vectors.  ROOT's unseeded
// global TRandom3 stream cannot be reproduced without ROOT, so the stream is defined here:
//   counter-based Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit seed, counter = (toy index k,
//   pair index i / 2); the four 32-bit outputs make two 53-bit uniforms u1 in (0, 1], u2 in [0, 1) and one
//   Box-Muller pair  z0 = sqrt(-2 ln u1) cos(2 pi u2),  z1 = sqrt(-2 ln u1) sin(2 pi u2)  for columns 2j, 2j+1.
// z depends only on (seed, k, i): the same toy is produced whatever the chunking, the grid or the number of
// GPUs a campaign is sharded over.  isr_draw_toys exposes the generator so that tests can run the CPU oracle
// on exactly the vectors the fused campaign used.
#include "problem.hpp"
#include "rng_device.cuh"

namespace isr {

// one thread per (toy, column pair)
__global__ void __launch_bounds__(256) draw_toys_kernel(int64_t T, int n, int64_t first_toy, uint64_t seed,
                                                         const double* __restrict__ vcs, const double* __restrict__ err,
                                                         double* __restrict__ V) {
  const int npair = (n + 1) >> 1;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * npair) return;
  const int64_t t = idx / npair;
  const int j = (int)(idx % npair);
  const uint64_t k = (uint64_t)(first_toy + t);
  uint32_t r[4];
  philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)j, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  // 53-bit uniforms: u1 in (0, 1] (safe for the logarithm), u2 in [0, 1)
  const uint64_t a = ((uint64_t)r[0] << 21) ^ (uint64_t)(r[1] >> 11), b = ((uint64_t)r[2] << 21) ^ (uint64_t)(r[3] >> 11);
  const double u1 = ((double)(a & ((1ull << 53) - 1)) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = (double)(b & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  const int i0 = 2 * j;
  double* row = V + t * n;
  row[i0] = fma(err[i0], rad * cs, vcs[i0]);
  if (i0 + 1 < n) row[i0 + 1] = fma(err[i0 + 1], rad * sn, vcs[i0 + 1]);
}

int draw_toys_device(isr_ctx* c, int64_t T, int n, int64_t first_toy, uint64_t seed, const double* vcs_dev,
                     const double* err_dev, double* V_dev, cudaStream_t st) {
  if (T <= 0) return ISR_OK;
  const int64_t work = T * ((n + 1) >> 1);
  draw_toys_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st ? st : c->stream>>>(T, n, first_toy, seed, vcs_dev, err_dev, V_dev); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

}  // namespace isr
