// Blocked inverse of a lower-triangular matrix and the small FP64 tile product it is built from.
//
// Replaces, for N > 240 (BASELINE config C5: N = 500, linear interpolation, no energy spread => G lower triangular),
// the cuBLAS triangular solve L X = I behind ISRSolverSLE::solve's inverse (src/ISRSolverSLE.cpp:91-94), and the cuBLAS
// product of the a-posteriori certificate max |A X - I| for every N.
//
// Recursive doubling on 32 x 32 blocks:
//   level 0   the N/32 diagonal blocks are inverted independently (one warp each: column j of a block's inverse is a
//             32-step forward substitution held by lane j);
//   level l   (s = 32 * 2^l)  [ A 0 ; C B ]^-1 = [ A^-1 0 ; -B^-1 C A^-1  B^-1 ] for every aligned pair of s x s
//             diagonal blocks whose inverses level l-1 produced: two launches, T = C A^-1 and X_C = -B^-1 T, each a
//             batch (blockIdx.z = pair) of 32 x 32 output tiles that visit only the K range where the triangular
//             factor is non-zero.
// N = 500: 1 + 2 * 4 launches, about N^3 / 3 flop in all.  Entries above the diagonal of X are written as exact zeros,
// so the toy kernel's triangular schedule may be selected from the host-side knowledge alone.
#include "linalg.hpp"

#include <algorithm>

namespace isr {

constexpr int kTb = 32;   // block / tile edge

// One warp per diagonal block b: X_bb = (L_bb)^-1.  Lane j holds column j of the inverse; the block of L sits in shared
// memory.  Rows / columns beyond n are treated as the identity.  Also zero-fills the part of the tile row above the
// diagonal (blocks (b, b' > b)) so that X is exactly lower triangular whatever the buffer held before.
__global__ void __launch_bounds__(32) tri_diag_inverse_kernel(int n, const double* __restrict__ L, double* __restrict__ X) {
  __shared__ double sL[kTb][kTb + 1];
  const int b = blockIdx.x, r0 = b * kTb, j = threadIdx.x;
  for (int i = 0; i < kTb; ++i) {
    const int gi = r0 + i, gj = r0 + j;
    sL[i][j] = (gi < n && gj < n) ? L[(size_t)gi + (size_t)n * gj] : (i == j ? 1.0 : 0.0);
  }
  __syncwarp();
  double x[kTb];
#pragma unroll
  for (int i = 0; i < kTb; ++i) {
    double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) acc = fma(-sL[i][k], x[k], acc);   // x[k] == 0 for k < j: exact zeros above the diagonal
    x[i] = (i >= j) ? acc / sL[i][i] : 0.0;
  }
  const int gj = r0 + j;
  if (gj < n) {
#pragma unroll
    for (int i = 0; i < kTb; ++i)
      if (r0 + i < n) X[(size_t)(r0 + i) + (size_t)n * gj] = x[i];
  }
  // blocks to the right of the diagonal block in this block row: zeros
  if (r0 + j < n)
    for (int gc = r0 + kTb; gc < n; ++gc) X[(size_t)(r0 + j) + (size_t)n * gc] = 0.0;   // lanes along a column: coalesced
}

// C_tile(32 x 32) = alpha * sum_k A(i, k) B(k, j) over k in [k_lo, k_hi) (multiples of 32 relative to the operands'
// origins), all operands column-major sub-matrices of n x n buffers with leading dimension n.  256 threads: thread
// (tx, ty) owns row tx and the four columns 4 ty .. 4 ty + 3; the B tile is stored k-major so that the four columns are
// one 32-byte broadcast load.
struct TileJob {
  const double* A; const double* B; double* C;   // origins of the sub-matrices
  int rows, cols, kdim;                          // extents (ragged edges are zero-filled / not written)
};
__device__ __forceinline__ void tile_product(const TileJob& jb, int ld, int tr, int tc, int k_lo, int k_hi, double alpha,
                                             double (*sA)[kTb + 1], double (*sB)[kTb + 4]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int row0 = tr * kTb, col0 = tc * kTb;
  double acc[4] = {0, 0, 0, 0};
  for (int k0 = k_lo; k0 < k_hi; k0 += kTb) {
    __syncthreads();
    for (int q = ty; q < kTb; q += 8) {
      const int ar = row0 + tx, ac = k0 + q;
      sA[q][tx] = (ar < jb.rows && ac < jb.kdim) ? jb.A[(size_t)ar + (size_t)ld * ac] : 0.0;   // sA[k][row]
      const int br = k0 + tx, bc = col0 + q;
      sB[tx][q] = (br < jb.kdim && bc < jb.cols) ? jb.B[(size_t)br + (size_t)ld * bc] : 0.0;   // sB[k][col]
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kTb; ++k) {
      const double a = sA[k][tx];
      const double4 b = *reinterpret_cast<const double4*>(&sB[k][ty * 4]);
      acc[0] = fma(a, b.x, acc[0]); acc[1] = fma(a, b.y, acc[1]); acc[2] = fma(a, b.z, acc[2]); acc[3] = fma(a, b.w, acc[3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = row0 + tx, col = col0 + ty * 4 + r;
    if (row < jb.rows && col < jb.cols) jb.C[(size_t)row + (size_t)ld * col] = alpha * acc[r];
  }
}

// mode 0: T = C A^-1   (pair z: C = L[lo + s .., lo ..], A^-1 = X[lo .., lo ..] lower triangular => k >= column tile)
// mode 1: X_C = -B^-1 T (B^-1 = X[lo + s .., lo + s ..] lower triangular => k <= row tile)
__global__ void __launch_bounds__(256) tri_level_kernel(int n, int s, int mode, const double* __restrict__ L,
                                                         double* __restrict__ X, double* __restrict__ T) {
  __shared__ double sA[kTb][kTb + 1];
  __shared__ __align__(32) double sB[kTb][kTb + 4];
  const int lo = blockIdx.z * 2 * s;           // first row / column of the pair
  const int hi = lo + s;                       // first row of the lower block
  if (hi >= n) return;                         // the pair has no lower block: nothing to do (whole block exits together)
  const int rows = min(s, n - hi), cols = s;   // the upper block is always complete when a lower block exists
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tr * kTb >= rows) return;
  TileJob jb;
  if (mode == 0) {
    jb.A = L + (size_t)hi + (size_t)n * lo;    // C block of L
    jb.B = X + (size_t)lo + (size_t)n * lo;    // A^-1
    jb.C = T + (size_t)hi + (size_t)n * lo;
    jb.rows = rows; jb.cols = cols; jb.kdim = s;
    tile_product(jb, n, tr, tc, tc * kTb, s, 1.0, sA, sB);
  } else {
    jb.A = X + (size_t)hi + (size_t)n * hi;    // B^-1
    jb.B = T + (size_t)hi + (size_t)n * lo;
    jb.C = X + (size_t)hi + (size_t)n * lo;
    jb.rows = rows; jb.cols = cols; jb.kdim = rows;
    tile_product(jb, n, tr, tc, 0, min(rows, (tr + 1) * kTb), -1.0, sA, sB);
  }
}

int tri_inverse_lower(isr_ctx* c, int n, const double* L, double* X) {
  cudaStream_t st = c->stream;
  ISR_TRY(c->scratch_t.reserve((size_t)n * n));
  const int nb = (n + kTb - 1) / kTb;
  tri_diag_inverse_kernel<<<nb, 32, 0, st>>>(n, L, X); ISR_COUNT_LAUNCH();
  for (int s = kTb; s < n; s *= 2) {
    const int pairs = (n + 2 * s - 1) / (2 * s), t = s / kTb;
    dim3 grid(t, t, pairs);
    tri_level_kernel<<<grid, 256, 0, st>>>(n, s, 0, L, X, c->scratch_t.p); ISR_COUNT_LAUNCH();
    tri_level_kernel<<<grid, 256, 0, st>>>(n, s, 1, L, X, c->scratch_t.p); ISR_COUNT_LAUNCH();
  }
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

// C = A B (n x n, column-major) for the inverse's certificate, K split into gridDim.z slices: slice z writes its partial
// product to C + z n^2 (the consumer adds the slices in order).  A 7 x 7 grid of tiles leaves two thirds of the machine
// idle at N = 200 and walks seven dependent K tiles per block (24 us); four slices make it 196 blocks of two.
// lower != 0: both factors are lower triangular, the product is too -- tiles above the diagonal are written as zeros and
// the others visit K in [column tile, row tile] only.
__global__ void __launch_bounds__(256) small_gemm_kernel(int n, int lower, const double* __restrict__ A,
                                                          const double* __restrict__ B, double* __restrict__ C) {
  __shared__ double sA[kTb][kTb + 1];
  __shared__ __align__(32) double sB[kTb][kTb + 4];
  const int tr = blockIdx.y, tc = blockIdx.x;
  TileJob jb{A, B, C + (size_t)blockIdx.z * n * n, n, n, n};
  if (lower && tc > tr) { tile_product(jb, n, tr, tc, 0, 0, 0.0, sA, sB); return; }
  const int k_lo = lower ? tc * kTb : 0, k_hi = lower ? min(n, (tr + 1) * kTb) : n;
  // this slice's share of the K tiles [k_lo, k_hi)
  const int tiles = (k_hi - k_lo + kTb - 1) / kTb, per = (tiles + (int)gridDim.z - 1) / (int)gridDim.z;
  const int a = k_lo + (int)blockIdx.z * per * kTb, b = min(k_hi, a + per * kTb);
  tile_product(jb, n, tr, tc, a, max(a, b), 1.0, sA, sB);
}
int small_gemm_slices(isr_ctx* c, int n) {
  const int t = (n + kTb - 1) / kTb;
  return std::max(1, std::min({8, t, (2 * c->sm_count + t * t - 1) / (t * t)}));
}
// C must hold small_gemm_slices(c, n) * n^2 doubles; *slices_out tells how many partial products to add.
int small_gemm(isr_ctx* c, int n, bool lower, const double* A, const double* B, double* C, cudaStream_t st, int* slices_out) {
  const int t = (n + kTb - 1) / kTb;
  const int nz = slices_out ? small_gemm_slices(c, n) : 1;
  small_gemm_kernel<<<dim3(t, t, nz), 256, 0, st ? st : c->stream>>>(n, lower ? 1 : 0, A, B, C); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  if (slices_out) *slices_out = nz;
  return ISR_OK;
}

__global__ void __launch_bounds__(256) small_gemm_batch_kernel(int n, const double* __restrict__ A, const double* __restrict__ B,
                                                                double* __restrict__ C) {
  __shared__ double sA[kTb][kTb + 1];
  __shared__ __align__(32) double sB[kTb][kTb + 4];
  const size_t off = (size_t)blockIdx.z * n * n;
  TileJob jb{A + off, B, C + off, n, n, n};
  tile_product(jb, n, blockIdx.y, blockIdx.x, 0, n, 1.0, sA, sB);
}
int small_gemm_batch(isr_ctx* c, cudaStream_t st, int n, int K, const double* A_batch, const double* B, double* C_batch) {
  if (K <= 0) return ISR_OK;
  const int t = (n + kTb - 1) / kTb;
  small_gemm_batch_kernel<<<dim3(t, t, K), 256, 0, st ? st : c->stream>>>(n, A_batch, B, C_batch); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

}  // namespace isr
