// FP64 tensor-pipe (DMMA, mma.sync.m8n8k4.f64) tile kernel behind the toy batch:
//     C[t, n] = sum_k A[t, k] * B[n, k]          A: T x K (toy-major), B: N x K (row-major operator)
// with fused epilogues.  The toy batch calls it twice (toy_batch.cu): pass 1 with B = S (solve
// operator, b = S v), epilogue "solve"; pass 2 with B = R (chi2 factor) on D = b - b0, epilogue
// "chi2".  North-star rule: DMMA is used because the probe in tools/fp64_peaks.cu measured 37.1
// TFLOP/s for the tensor pipe against 33.3 TFLOP/s for DFMA on this B200 (profiles/fp64_peaks_r01.json);
// the DFMA variant below stays selectable (isr_set_toy_kernel) for the ncu comparison.
//
// Tiling: one CTA = WM x WN warps, warp tile = (AT*8 toys) x (BT*8 columns); K streamed in chunks of
// 24 doubles through a cp.async (LDGSTS) ring of STAGES buffers.  Rows are 24 doubles = 192 B, i.e.
// stride = 8 (mod 16) doubles, which makes the 16-byte fragment loads below bank-conflict free with
// no padding.  Each 16-byte load feeds two DMMAs: the k-permutation {2c, 2c+1} is applied to A and B
// alike, so the contraction is unchanged.
#pragma once

#include "common.cuh"

namespace isr {

constexpr int kKC = 24;  // K chunk (doubles); 24 = 8 (mod 16) keeps LDS.128 conflict-free

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

enum { EPI_SOLVE = 0, EPI_CHI2 = 1 };

struct GemmArgs {
  const double* A;   // T x lda
  const double* B;   // N x ldb (operator, row-major, K contiguous)
  int64_t T;
  int N, K, lda, ldb;
  // EPI_SOLVE
  const double* b0;  // N
  double* D;         // T x ldd : b - b0 (input of pass 2)
  int ldd;
  double* Bout;      // T x N or nullptr : b
  double* part;      // [gridDim.x][4][Npad_total] per-CTA partial sums {sum b, count, sum r, sum r^2} or nullptr
  int part_ld;       // Npad_total
  int want_ratio;    // accumulate ratio statistics
  uint32_t* rhist;   // N x 1026 or nullptr
  // EPI_CHI2
  double* chi2;      // T (single column block) or partial [nblk][T]
};

template <int WM, int WN, int AT, int BT, int STAGES, int EPI>
__global__ void __launch_bounds__(WM* WN * 32, 1) toy_gemm_kernel(GemmArgs g) {
  constexpr int TM = WM * AT * 8;
  constexpr int NP = WN * BT * 8;
  constexpr int NTHREADS = WM * WN * 32;
  constexpr int ROWS = TM + NP;
  constexpr int PIECES = ROWS * (kKC / 2);  // 16-byte pieces per stage
  extern __shared__ __align__(16) double smem[];
  // layout: STAGES x [ROWS][kKC], then 4 x NP running column totals (EPI_SOLVE).  The epilogues
  // reuse the stage buffers as scratch once the K loop has drained.
  double* stage0 = smem;
  double* scratch = smem + (size_t)STAGES * ROWS * kKC;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int gq = lane >> 2, c = lane & 3;
  const int n0 = blockIdx.y * NP;
  const int nk = (g.K + kKC - 1) / kKC;
  const int64_t ntiles = (g.T + TM - 1) / TM;

  // per-CTA running column statistics (EPI_SOLVE): scratch[4][NP] totals + [WM][4][NP] per-tile partials
  double* sTot = scratch;
  double* sPart = stage0;   // [WM][4][NP], valid only inside the epilogue
  if (EPI == EPI_SOLVE && g.part) {
    for (int q = tid; q < 4 * NP; q += NTHREADS) sTot[q] = 0.0;
  }
  __syncthreads();

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = tile * TM;
    double acc[AT][BT][2];
#pragma unroll
    for (int i = 0; i < AT; ++i)
#pragma unroll
      for (int j = 0; j < BT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_chunk = [&](int kc, int st) {
      double* dst = stage0 + (size_t)st * ROWS * kKC;
      const int k0 = kc * kKC;
      for (int q = tid; q < PIECES; q += NTHREADS) {
        const int row = q / (kKC / 2), pc = q % (kKC / 2);
        const int k = k0 + pc * 2;
        const double* src;
        int bytes = 0;
        if (row < TM) {
          const int64_t t = t0 + row;
          src = g.A;
          if (t < g.T && k < g.K) { src = g.A + t * g.lda + k; bytes = 16; }
        } else {
          const int n = n0 + row - TM;
          src = g.B;
          if (n < g.N && k < g.K) { src = g.B + (size_t)n * g.ldb + k; bytes = 16; }
        }
        cp_async16(dst + row * kKC + pc * 2, src, bytes);
      }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nk) load_chunk(s, s);
      cp_async_commit();
    }
    for (int kc = 0; kc < nk; ++kc) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {  // prefetch chunk kc + STAGES - 1 into the buffer freed by iteration kc - 1
        const int nxt = kc + STAGES - 1;
        if (nxt < nk) load_chunk(nxt, nxt % STAGES);
        cp_async_commit();
      }
      const double* sA = stage0 + (size_t)(kc % STAGES) * ROWS * kKC;
      const double* sB = sA + TM * kKC;
      const int ksteps = min(kKC, g.K - kc * kKC + 7) / 8;  // 8 physical k per double-step
#pragma unroll
      for (int kp = 0; kp < kKC / 8; ++kp) {
        if (kp < ksteps) {
          double2 a[AT], b[BT];
#pragma unroll
          for (int i = 0; i < AT; ++i)
            a[i] = *reinterpret_cast<const double2*>(sA + ((wm * AT + i) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
          for (int j = 0; j < BT; ++j)
            b[j] = *reinterpret_cast<const double2*>(sB + ((wn * BT + j) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
          for (int i = 0; i < AT; ++i)
#pragma unroll
            for (int j = 0; j < BT; ++j) {
              dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
              dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            }
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the last stage: the stage buffers become scratch

    // ------------------------------------------------------------------ epilogues
    if (EPI == EPI_SOLVE) {
      const bool stats = g.part != nullptr;
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        const int col = n0 + (wn * BT + j) * 8 + 2 * c;
        const bool ok0 = col < g.N, ok1 = col + 1 < g.N;
        const double b00 = ok0 ? g.b0[col] : 0.0, b01 = ok1 ? g.b0[col + 1] : 0.0;
        double cs[2] = {0, 0}, cn[2] = {0, 0}, cr[2] = {0, 0}, cq[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < AT; ++i) {
          const int64_t t = t0 + (wm * AT + i) * 8 + gq;
          if (t < g.T) {
            const double v0 = acc[i][j][0], v1 = acc[i][j][1];
            const double d0 = v0 - b00, d1 = v1 - b01;
            if (ok1) {
              *reinterpret_cast<double2*>(g.D + t * g.ldd + col) = make_double2(d0, d1);
            } else if (ok0) {
              g.D[t * g.ldd + col] = d0;
            }
            if (g.Bout) {
              if (ok0) g.Bout[t * g.N + col] = v0;
              if (ok1) g.Bout[t * g.N + col + 1] = v1;
            }
            cs[0] += v0;
            cs[1] += v1;
            if (g.want_ratio) {
              // TH1F(1024, -2, 2)::Fill((b - b0)/b0): statistics of in-range entries only
              // (src/RatioTest.cpp:36,43,51-52)
              const double r0 = d0 / b00, r1 = d1 / b01;
              if (ok0 && r0 >= -2.0 && r0 < 2.0) { cn[0] += 1.0; cr[0] += r0; cq[0] += r0 * r0; }
              if (ok1 && r1 >= -2.0 && r1 < 2.0) { cn[1] += 1.0; cr[1] += r1; cq[1] += r1 * r1; }
              if (g.rhist) {
                if (ok0) {
                  int bin = r0 < -2.0 ? 0 : (!(r0 < 2.0) ? 1025 : 1 + (int)__ddiv_rn(__dmul_rn(1024.0, __dsub_rn(r0, -2.0)), 4.0));
                  atomicAdd(&g.rhist[(size_t)col * 1026 + bin], 1u);
                }
                if (ok1) {
                  int bin = r1 < -2.0 ? 0 : (!(r1 < 2.0) ? 1025 : 1 + (int)__ddiv_rn(__dmul_rn(1024.0, __dsub_rn(r1, -2.0)), 4.0));
                  atomicAdd(&g.rhist[(size_t)(col + 1) * 1026 + bin], 1u);
                }
              }
            }
          }
        }
        if (stats) {
          // fixed-order reduction: the 8 toy lanes by shuffles here, the WM warps through smem below
#pragma unroll
          for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int sft = 4; sft <= 16; sft <<= 1) {
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], sft);
              cn[e] += __shfl_xor_sync(0xffffffffu, cn[e], sft);
              cr[e] += __shfl_xor_sync(0xffffffffu, cr[e], sft);
              cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], sft);
            }
            if (gq == 0) {
              const int lc = (wn * BT + j) * 8 + 2 * c + e;
              double* pp = sPart + (size_t)wm * 4 * NP;
              pp[0 * NP + lc] = cs[e];
              pp[1 * NP + lc] = cn[e];
              pp[2 * NP + lc] = cr[e];
              pp[3 * NP + lc] = cq[e];
            }
          }
        }
      }
      if (stats) {
        __syncthreads();
        for (int q = tid; q < 4 * NP; q += NTHREADS) {
          double s = sTot[q];
#pragma unroll
          for (int w = 0; w < WM; ++w) s += sPart[(size_t)w * 4 * NP + q];
          sTot[q] = s;
        }
      }
    } else {
      // chi2_t = sum_n C[t, n]^2 : thread-local, then the 4 column lanes, then the WN warps (fixed order)
      double* sRow = stage0;  // [WN][TM]
#pragma unroll
      for (int i = 0; i < AT; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < BT; ++j) s += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (c == 0) sRow[wn * TM + (wm * AT + i) * 8 + gq] = s;
      }
      __syncthreads();
      for (int r = tid; r < TM; r += NTHREADS) {
        const int64_t t = t0 + r;
        if (t < g.T) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < WN; ++w) s += sRow[w * TM + r];
          g.chi2[(size_t)blockIdx.y * g.T + t] = s;
        }
      }
    }
    __syncthreads();  // smem stages and scratch are reused by the next tile
  }

  if (EPI == EPI_SOLVE && g.part) {
    for (int q = tid; q < 4 * NP; q += NTHREADS) {
      const int stat = q / NP, lc = q % NP;
      g.part[((size_t)blockIdx.x * 4 + stat) * g.part_ld + n0 + lc] = sTot[q];
    }
  }
}

template <int WM, int WN, int AT, int BT, int STAGES>
constexpr size_t toy_gemm_smem() {
  constexpr int TM = WM * AT * 8, NP = WN * BT * 8;
  static_assert((size_t)WM * 4 * NP <= (size_t)STAGES * (TM + NP) * kKC, "epilogue scratch must fit the stage buffers");
  static_assert((size_t)WN * TM <= (size_t)STAGES * (TM + NP) * kKC, "epilogue scratch must fit the stage buffers");
  return (size_t)STAGES * (TM + NP) * kKC * sizeof(double) + (size_t)4 * NP * sizeof(double);
}

}  // namespace isr
