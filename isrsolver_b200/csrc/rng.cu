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
//
// Two normal transforms over the same Philox stream family:
//   * default (round 2): the Box-Muller pair is formed in SINGLE precision on the special-function unit (MUFU.LG2,
//     MUFU.SQRT, MUFU.SIN / COS) and only the final v = vcs + err * z is an FP64 FMA.  The FP64 transform below costs
//     ~250 FP64 instructions per pair on the very pipe the toy kernel's DMMA saturates (a campaign by count ran at 0.79
//     of the rate with resident toys); this one issues 2 FP64 FMAs + 2 conversions per pair.  One Philox call serves two
//     pairs (radius from 32 bits: |z| <= 6.76, P = 1.4e-11 beyond; angle from 23 bits).  z carries ~1e-6 of absolute
//     error -- a distortion of N(0, 1) far below what any campaign can resolve (and of the order of ROOT's own
//     TRandom3::Gaus acceptance tables); the toys are exactly reproducible on sm_100a all the same;
//   * ISR_TOY_RNG=fp64 in the environment: the round-1 transform, 53-bit uniforms and FP64 log / sqrt / sincospi.
#include <cstdlib>
#include <cstring>

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

// one thread per column quad, looping over kFastIters toy groups: one Philox call -> two single-precision Box-Muller
// pairs -> four columns of one toy.  err / vcs of the quad stay in registers over the loop (the first version re-loaded
// them per toy and stalled on the load / store queue: 33 % issue-active, ncu lg_throttle + drain).
constexpr int kFastIters = 8;

// float -> double by bit arithmetic on the integer pipe (F2F.F64.F32 shares the 16-lane special-function unit with the four
// MUFU operations of a pair, the busiest pipe of this kernel).  Exact for normal floats; the inputs are flushed to zero
// below 2^-126, and +-0 becomes +-2^-127, which vcs + err * z rounds away.
__device__ __forceinline__ double widen(float f) {
  const uint32_t b = __float_as_uint(f);
  const uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
  return __hiloint2double((int)hi, (int)(b << 29));
}

__device__ __forceinline__ void fast_normals(uint64_t k, unsigned q, uint64_t seed, double z[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), q, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // radius: u = (r + 0.5) 2^-32 in (0, 1]; -2 ln u = -2 ln2 lg2(u)
    const float u = fmaf(__uint2float_rn(r[2 * h]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float lg, rad;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(fmaxf(-1.3862943611198906f * lg, 0.0f)));
    // angle: 23 bits -> f in [1, 2) -> 2 pi (f - 1.5) in [-pi, pi)
    const float ang = 6.283185307179586f * (__uint_as_float(0x3f800000u | (r[2 * h + 1] >> 9)) - 1.5f);
    float sn, cs;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
    z[2 * h] = widen(rad * cs);
    z[2 * h + 1] = widen(rad * sn);
  }
}

__global__ void __launch_bounds__(256) draw_toys_fast_kernel(int64_t T, int n, int64_t first_toy, uint64_t seed,
                                                              const double* __restrict__ vcs, const double* __restrict__ err,
                                                              double* __restrict__ V, unsigned nquad, unsigned tpb, unsigned magic) {
  // a block covers kFastIters groups of `tpb` whole toys (tpb * nquad <= 256 threads busy); magic = ceil(2^16 / nquad).
  // Thread q of a toy owns the column pairs q and q + nquad (columns 2q, 2q+1 and 2(q+nquad), 2(q+nquad)+1), so that each
  // of its two 16-byte stores is contiguous across the lanes of the warp (one pair per lane: 512 B per store instruction;
  // adjacent pairs per thread made every store instruction touch twice the L1 wavefronts: 82 % L1 throughput at 33 % issue).
  // N > 1024 (nquad > 256): tpb = 1 and the thread walks the quads tid, tid + 256, ...
  const unsigned lt = (threadIdx.x * magic) >> 16, q0 = threadIdx.x - lt * nquad;
  if (lt >= tpb) return;
  const bool even = (n & 1) == 0;                          // even N: every pair is 16-byte aligned
  const int64_t tb = (int64_t)blockIdx.x * kFastIters * tpb + lt;
  for (unsigned q = q0; q < nquad; q += 256) {
    const int col[4] = {2 * (int)q, 2 * (int)q + 1, 2 * (int)(q + nquad), 2 * (int)(q + nquad) + 1};
    double e[4], c[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const bool in = col[a] < n;
      e[a] = in ? err[col[a]] : 0.0;
      c[a] = in ? vcs[col[a]] : 0.0;
    }
#pragma unroll 1
    for (int it = 0; it < kFastIters; ++it) {
      const int64_t t = tb + (int64_t)it * tpb;
      if (t >= T) break;
      double z[4];
      fast_normals((uint64_t)(first_toy + t), q, seed, z);
      double* row = V + t * n;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double v0 = fma(e[2 * h], z[2 * h], c[2 * h]), v1 = fma(e[2 * h + 1], z[2 * h + 1], c[2 * h + 1]);
        if (even && col[2 * h + 1] < n) {
          *reinterpret_cast<double2*>(row + col[2 * h]) = make_double2(v0, v1);
        } else {
          if (col[2 * h] < n) row[col[2 * h]] = v0;
          if (col[2 * h + 1] < n) row[col[2 * h + 1]] = v1;
        }
      }
    }
  }
}

static bool rng_fp64() {
  static const bool v = [] {
    const char* e = getenv("ISR_TOY_RNG");
    return e && (strcmp(e, "fp64") == 0 || strcmp(e, "FP64") == 0);
  }();
  return v;
}

int draw_toys_device(isr_ctx* c, int64_t T, int n, int64_t first_toy, uint64_t seed, const double* vcs_dev,
                     const double* err_dev, double* V_dev, cudaStream_t st) {
  if (T <= 0) return ISR_OK;
  if (!rng_fp64()) {
    const unsigned nquad = (unsigned)(n + 3) >> 2, tpb = nquad < 256u ? 256u / nquad : 1u, magic = (65536u + nquad - 1) / nquad;
    draw_toys_fast_kernel<<<(unsigned)((T + (int64_t)tpb * kFastIters - 1) / ((int64_t)tpb * kFastIters)), 256, 0, st ? st : c->stream>>>(T, n, first_toy, seed, vcs_dev, err_dev, V_dev, nquad, tpb, magic); ISR_COUNT_LAUNCH();
    ISR_CUDA(cudaGetLastError());
    return ISR_OK;
  }
  const int64_t work = T * ((n + 1) >> 1);
  draw_toys_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st ? st : c->stream>>>(T, n, first_toy, seed, vcs_dev, err_dev, V_dev); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

}  // namespace isr
