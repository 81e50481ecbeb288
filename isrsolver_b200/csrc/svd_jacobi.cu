// Batched singular values by one-sided (Hestenes) Jacobi -- the hot loop of the condition-number-versus-energy-spread
// scan (src/isrsolver-condnum-test.cpp:88-140; ISRSolverSLE::evalConditionNumber, src/ISRSolverSLE.cpp:200-204, is a
// JacobiSVD of the operator matrix per spread setting; notebooks/condnum_enspread.ipynb scans 100 settings).
//
// Round 1 ran one cuSOLVER gesvd per setting on 8 streams: about 11 ms each, 1.1 s for 100 settings at N = 200, the batch
// axis unused.  Here ONE launch handles the whole batch, one CTA (1024 threads) per matrix:
//   * columns a_p, a_q of a pair are rotated so that they become orthogonal (gamma = a_p . a_q -> 0); after convergence
//     the singular values are the column norms.  The squared norms are cached in shared memory and updated analytically
//     (alpha' = alpha - t gamma, beta' = beta + t gamma), so a pair costs ONE dot product; they are recomputed from the
//     columns at the start of every sweep;
//   * round-robin ("circle") ordering: N - 1 rounds of N / 2 disjoint pairs per sweep, pair i of round r computed in
//     closed form; a warp owns pairs i = w, w + 32, ...; one block barrier per round;
//   * the matrix (N^2 doubles: 320 kB at N = 200, more than shared memory) stays in global memory, i.e. in L2: a warp
//     streams two 1.6 kB columns per pair with coalesced loads (32 warps per CTA hide the L2 latency of each other).
//     One CTA per SM, so a batch of up to 148 matrices runs concurrently.
// Bound: L2 / barrier latency (199 rounds x ~8 sweeps per matrix), not FP64 throughput; the batch axis is what fills
// the machine.  One-sided Jacobi is backward stable and computes small singular values at least as accurately as the
// bidiagonalisation-based gesvd (Demmel & Veselic), which matters for the ill-conditioned spread products.
#include "linalg.hpp"

namespace isr {

// Template: PER = column elements per lane (N <= 32 PER), THREADS per CTA.  A lane keeps 2 PER doubles of the current pair
// in registers, so 1024 threads (64 registers each) go with PER <= 8 (N <= 256) and 512 threads with PER = 16 (N <= 512).

__device__ __forceinline__ void circle_pair(int M, int r, int i, int& p, int& q) {
  // M even players, round r in [0, M - 1), table i in [0, M / 2): player M - 1 stays, the others rotate
  if (i == 0) { p = M - 1; q = r; return; }
  int a = r + i, b = r - i;
  if (a >= M - 1) a -= M - 1;
  if (b < 0) b += M - 1;
  p = a; q = b;
}

// sv[k][0 .. n) = singular values of A_k, descending; sweeps[k] = sweeps used (negative: not converged)
template <int kJacMaxPerLane, int kJacThreads>
__global__ void __launch_bounds__(kJacThreads, 1)
jacobi_sv_kernel(int n, double* Abatch, double* __restrict__ svbatch, int* __restrict__ sweeps, int max_sweeps, double tol) {
  extern __shared__ double sm_norm[];      // [M] squared column norms
  __shared__ int s_rot;
  double* A = Abatch + (size_t)blockIdx.x * n * n;
  double* sv = svbatch + (size_t)blockIdx.x * n;
  constexpr int kJacWarps = kJacThreads / 32;
  const int M = (n + 1) & ~1;              // odd n: a phantom column that is never rotated
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + 31) / 32;           // elements of a column per lane
  int sweep = 0;
  bool converged = false;
  for (; sweep < max_sweeps && !converged; ++sweep) {
    // squared norms from the columns (fixed order: lane-strided partial sums + xor tree)
    for (int j = warp; j < M; j += kJacWarps) {
      double s = 0.0;
      if (j < n) {
        const double* col = A + (size_t)j * n;
        for (int e = lane; e < n; e += 32) { const double v = __ldcg(col + e); s = fma(v, v, s); }   // (L2: other warps' st.cg)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      if (lane == 0) sm_norm[j] = s;
    }
    if (tid == 0) s_rot = 0;
    __syncthreads();
    int rotations = 0;
    for (int r = 0; r < M - 1; ++r) {
      for (int i = warp; i < M / 2; i += kJacWarps) {
        int p, q;
        circle_pair(M, r, i, p, q);
        if (p >= n || q >= n) continue;    // phantom column
        double* cp = A + (size_t)p * n;
        double* cq = A + (size_t)q * n;
        double ap[kJacMaxPerLane], aq[kJacMaxPerLane];
        double g = 0.0;
#pragma unroll
        for (int e = 0; e < kJacMaxPerLane; ++e) {
          if (e < per) {
            const int idx = lane + 32 * e;
            ap[e] = idx < n ? __ldcg(cp + idx) : 0.0;
            aq[e] = idx < n ? __ldcg(cq + idx) : 0.0;
            g = fma(ap[e], aq[e], g);
          }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        const double alpha = sm_norm[p], beta = sm_norm[q];
        if (g * g > tol * tol * alpha * beta && alpha > 0.0 && beta > 0.0) {
          const double zeta = (beta - alpha) / (2.0 * g);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = rsqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int e = 0; e < kJacMaxPerLane; ++e) {
            if (e < per) {
              const int idx = lane + 32 * e;
              if (idx < n) {
                __stcg(cp + idx, fma(c, ap[e], -s * aq[e]));
                __stcg(cq + idx, fma(s, ap[e], c * aq[e]));
              }
            }
          }
          if (lane == 0) {
            sm_norm[p] = alpha - t * g;
            sm_norm[q] = beta + t * g;
          }
          ++rotations;
        }
      }
      __syncthreads();       // the pairs of a round are disjoint; the next round must see this round's columns
    }
    if (lane == 0 && rotations) atomicAdd(&s_rot, rotations);
    __syncthreads();
    converged = s_rot == 0;
    __syncthreads();
  }
  // singular values = column norms (recomputed), sorted descending by rank
  for (int j = warp; j < n; j += kJacWarps) {
    const double* col = A + (size_t)j * n;
    double s = 0.0;
    for (int e = lane; e < n; e += 32) { const double v = __ldcg(col + e); s = fma(v, v, s); }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sm_norm[j] = sqrt(s);
  }
  __syncthreads();
  for (int j = tid; j < n; j += kJacThreads) {
    const double v = sm_norm[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const double u = sm_norm[k];
      rank += (u > v) || (u == v && k < j);
    }
    sv[rank] = v;
  }
  if (tid == 0) sweeps[blockIdx.x] = converged ? sweep : -sweep;
}

// Singular values of K matrices (n x n column-major, back to back in A_batch, destroyed), descending in sv_batch[K][n].
// Asynchronous on `st`; sweeps_dev[K] receives the number of sweeps (negative if max_sweeps did not suffice).
int jacobi_singular_values_batch(isr_ctx* c, cudaStream_t st, int n, int K, double* A_batch, double* sv_batch, int* sweeps_dev) {
  if (n > 512) { set_error("batched Jacobi SVD: N = %d exceeds 512", n); return ISR_EINVAL; }
  if (K <= 0) return ISR_OK;
  const int M = (n + 1) & ~1;
  cudaStream_t s = st ? st : c->stream;
  const size_t smem = sizeof(double) * M;
  const int max_sweeps = 40;
  // a pair counts as orthogonal when |a_p . a_q| <= N eps |a_p| |a_q| (LAPACK dgesvj's default CTOL = M): the dot product
  // of two columns carries about sqrt(N) eps |a_p| |a_q| of rounding noise, so a tighter test (1e-15 was tried: twice the
  // sweeps) keeps rotating noise
  const double tol = n * 1.1102230246251565e-16;
  if (n <= 128) { jacobi_sv_kernel<4, 1024><<<K, 1024, smem, s>>>(n, A_batch, sv_batch, sweeps_dev, max_sweeps, tol); }
  else if (n <= 256) { jacobi_sv_kernel<8, 1024><<<K, 1024, smem, s>>>(n, A_batch, sv_batch, sweeps_dev, max_sweeps, tol); }
  else { jacobi_sv_kernel<16, 512><<<K, 512, smem, s>>>(n, A_batch, sv_batch, sweeps_dev, max_sweeps, tol); }
  ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

}  // namespace isr
