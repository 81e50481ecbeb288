// Fused FP64 tensor-pipe (DMMA, mma.sync.m8n8k4.f64) kernel of the toy batch.  Per toy k:
//     pass 1   b_k    = S v_k               (S: N x N solve operator)        2 N^2 flop
//     pass 2   chi2_k = | R (b_k - b0) |^2  (R: N x N chi2 factor)           2 N^2 flop
// Both passes are the same tile product  C[t, n] = sum_k A[t, k] * B[n, k]  (A toy-major, B row-major,
// K contiguous in both) with different epilogues.  ONE persistent launch does both: a CTA owns tiles of
// TM toys; for each tile it runs pass 1 (epilogue: d = b - b0 into a small per-CTA ring in global memory
// that stays in L2, optional b output, column statistics) and later pass 2 on the same tile (epilogue:
// row sums of squares -> chi2).  The passes of consecutive tiles are interleaved
//     P1(0) P1(1) P2(0) P1(2) P2(1) ... P2(m-1)
// so that the d-tile a pass 2 streams back was stored a whole tile-pass earlier.
//
// TMA pipeline without a block-wide barrier: every K chunk is two cp.async.bulk.tensor (TMA) box loads
// (A box TM x KC doubles, B box NP x KC doubles) into a ring of STAGES shared-memory buffers, completion
// signalled on a "full" mbarrier per stage.  All WM x WN warps are consumers; when a warp is done with a
// stage it bumps a shared counter, and the LAST warp to finish a stage immediately issues the TMA loads
// of the chunk STAGES ahead into it (so there is no dedicated producer warp -- 21 warps would cap the
// register file at 80 registers/thread on one SM sub-partition -- and no per-chunk __syncthreads: the
// warps drift apart and the DMMA pipe stays fed while others sit in an epilogue).  Out-of-range rows /
// columns are zero-filled by the TMA unit (no bounds code in the loop).
//
// North-star rule: DMMA is used because tools/fp64_peaks.cu measured 37.1 TFLOP/s for the tensor pipe
// against 33.3 TFLOP/s for DFMA on this B200 (profiles/fp64_peaks_r01.json).
//
// Tiling: warp tile = (AT*8 toys) x (BT*8 columns), K chunks of KC = 24 / 40 / 56 doubles.  Rows of KC doubles
// (192 / 320 / 448 B) have stride 8 (mod 16) doubles, which makes the 16-byte fragment loads conflict free without padding
// or swizzle; each 16-byte load feeds two DMMAs (the k-permutation {2c, 2c+1} is applied to A and B
// alike, so the contraction is unchanged).  Operators wider than one CTA tile (N > NP) are processed as
// column blocks by the same CTA.  The BT column tiles of a warp are dealt cyclically (tile j of warp wn = j WN + wn).
//
// Two instantiations per tile shape: TRI = false, the dense kernel; TRI = true for operator pairs of which at least
// one is EXACTLY lower triangular (linear interpolation without energy spread: G^-1 and W^(1/2) G; or a chi2 factor
// replaced by its QL factor) -- column blocks then load only the K chunks below their last row and a warp skips the
// column tiles entirely above the diagonal, through branches into chunk bodies specialised for the number of skipped
// tiles.  The skipped products are exact zeros, so both instantiations return the same bits.
// History (profiles/ncu_r01_toy_v*.txt): v1 two launches per chunk, per-tile drain: DMMA pipe 71 % busy
// inside the kernels, 18.7 TFLOP/s overall; v2 fused cp.async version with a block barrier per chunk: 58 %
// (every warp issued its share of the loads at the same moment, leaving the pipe idle).
#pragma once

#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace isr {

// K chunk KC (doubles, template parameter): rows of KC doubles must have stride 8 (mod 16) doubles -- 24, 40, 56 --
// to keep the 16-byte fragment loads conflict-free.  Fewer, fuller chunks are faster: every chunk boundary costs a
// barrier wait, a stage hand-over and a TMA issue (measured at N = 200: KC = 24 -> 9 chunks, the last one a third
// full, 31.8 TFLOP/s; KC = 40 -> 5 full chunks, 33.4 TFLOP/s), but the stage ring has to fit 227 KB.
constexpr int kRingSlots = 3;  // d-tiles alive per CTA (2 by schedule, +1 slack)

// L2 eviction-priority policies for TMA loads (same encodings CUTLASS uses for SM90+ cache hints)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // toys: streamed once
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;    // operators and the d ring: reused by every tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// TAG: the chunk body specialised for TAG skipped column tiles carries it as a PTX comment, which keeps the compiler
// from merging the tails of the specialised bodies into one another (that splits the k loop of a body in two and
// re-loads the A fragments: measured 33.4 -> 30.7 TFLOP/s on the dense schedule)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
template <int TAG>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1}; // body %4\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b), "n"(TAG));
}

struct ToyArgs {
  const double* b0;   // N
  int64_t T;
  int N, ld, nblk, do2;
  double* Dring;      // [gridDim.x][kRingSlots][TM][ld]   (do2)
  double* Dfull;      // T x ld                            (!do2: projection mode)
  double* Bout;       // T x N or nullptr
  double* chi2;       // T or nullptr
  double* part;       // [gridDim.x][4][part_ld] per-CTA {sum b, count, sum r, sum r^2} or nullptr
  int part_ld, nstat; // nstat = 1 (sum only) or 4
  uint32_t* rhist;    // N x 1026 or nullptr
  int tri;            // bit 0 / bit 1: the pass-1 / pass-2 operator is exactly lower triangular (B[n][k] == 0 for k > n)
};

// Position in the CTA's chunk stream.  Units: U1(i) = pass 1 of local tile i over all column blocks,
// U2(i) = pass 2.  Interleaved order: U1(0) U1(1) U2(0) U1(2) U2(1) ... U2(m-1).  A d-tile must be complete
// before the first loads of its pass 2 are issued, and loads run STAGES chunks ahead of the slowest warp:
// at least STAGES-1 chunks must lie between the end of U1(i) and the start of U2(i).  When that does not
// hold (a single tile, or an operator so small that a unit is shorter than STAGES-1 chunks) the
// sequential order U1(i) NOP U2(i) is used instead, NOP being STAGES-1 empty chunks.
enum { PH_P1 = 0, PH_P2 = 1, PH_NOP = 2 };
// (The schedule structs are __host__ __device__: tests/cpp/test_sched.cu walks them on the CPU and checks the
// invariants stated above for every tile shape, chunk count and triangularity.)
// Dense schedule: every column block has all nk chunks.
struct SchedDense {
  int m, u;        // local tiles, unit index
  int nblk, nk, nop, do2, seq;
  int cb, kc;
  __host__ __device__ void init(int m_, int nblk_, int nk_, int nop_, int do2_, int, int) {
    m = m_; nblk = nblk_; nk = nk_; nop = nop_; do2 = do2_; u = 0; cb = 0; kc = 0;
    seq = (m == 1) || (nk * nblk < nop);
  }
  __host__ __device__ int units() const { return do2 ? (seq ? 3 * m : 2 * m) : m; }
  __host__ __device__ bool done() const { return u >= units(); }
  __host__ __device__ int kind() const {
    if (!do2) return PH_P1;
    if (seq) { const int r = u % 3; return r == 0 ? PH_P1 : (r == 1 ? PH_NOP : PH_P2); }
    if (u == 0) return PH_P1;
    if (u == 2 * m - 1) return PH_P2;
    return (u & 1) ? PH_P1 : PH_P2;
  }
  __host__ __device__ int tile() const {   // local tile index
    if (!do2) return u;
    if (seq) return u / 3;
    if (u == 0) return 0;
    if (u == 2 * m - 1) return m - 1;
    return (u & 1) ? (u + 1) / 2 : u / 2 - 1;
  }
  __host__ __device__ bool last_chunk() const { return kc == nk - 1; }   // of the current column block
  __host__ __device__ void advance() {
    if (kind() == PH_NOP) {
      if (++kc == nop) { kc = 0; ++u; }
      return;
    }
    if (++kc == nk) {
      kc = 0;
      if (++cb == nblk) { cb = 0; ++u; }
    }
  }
};
// Triangular schedule (at least one operator exactly lower triangular): the chunk count depends on the pass and the
// column block.
template <int NP_, int KC_>
struct SchedTri {
  static constexpr int np = NP_, kcs = KC_;   // compile-time: the chunk count of a triangular column block is a division
  int m, u;        // local tiles, unit index
  int nblk, nk, nop, do2, seq;
  int cb, kc;
  int tri, ld;   // tri: bit 0 = pass-1 operator lower triangular, bit 1 = pass-2 operator.  Column block cb of a
                          // triangular operator only has non-zero K below (cb + 1) * np: the chunks beyond are not loaded
  __host__ __device__ int nk_tri(int b, int pass2) const {
    if (!((tri >> pass2) & 1)) return nk;
    const int lim = min(ld, (b + 1) * np);
    return (lim + kcs - 1) / kcs;
  }
  int knd, cnk;   // kind and chunk count of the current (unit, column block): refreshed on transitions only
  __host__ __device__ void refresh() {
    knd = kind_of_unit();
    cnk = knd == PH_NOP ? nop : nk_tri(cb, knd == PH_P2);
  }
  __host__ __device__ void init(int m_, int nblk_, int nk_, int nop_, int do2_, int tri_, int ld_) {
    m = m_; nblk = nblk_; nk = nk_; nop = nop_; do2 = do2_; u = 0; cb = 0; kc = 0;
    tri = tri_; ld = ld_;
    int t1 = 0, t2 = 0;
    for (int b = 0; b < nblk; ++b) { t1 += nk_tri(b, 0); t2 += nk_tri(b, 1); }
    seq = (m == 1) || (min(t1, t2) < nop);
    refresh();
  }
  __host__ __device__ int units() const { return do2 ? (seq ? 3 * m : 2 * m) : m; }
  __host__ __device__ bool done() const { return u >= units(); }
  __host__ __device__ int kind() const { return knd; }
  __host__ __device__ int kind_of_unit() const {
    if (!do2) return PH_P1;
    if (seq) { const int r = u % 3; return r == 0 ? PH_P1 : (r == 1 ? PH_NOP : PH_P2); }
    if (u == 0) return PH_P1;
    if (u == 2 * m - 1) return PH_P2;
    return (u & 1) ? PH_P1 : PH_P2;
  }
  __host__ __device__ int tile() const {   // local tile index
    if (!do2) return u;
    if (seq) return u / 3;
    if (u == 0) return 0;
    if (u == 2 * m - 1) return m - 1;
    return (u & 1) ? (u + 1) / 2 : u / 2 - 1;
  }
  __host__ __device__ bool last_chunk() const { return kc == cnk - 1; }   // of the current column block
  __host__ __device__ void advance() {
    if (++kc == cnk) {
      kc = 0;
      if (knd == PH_NOP || ++cb == nblk) { cb = 0; ++u; }
      refresh();
    }
  }
};

// Triangular schedule: first column tile j of warp wn (tile j covers columns (j WN + wn) 8 .. + 8 of the column block) that
// has a non-zero in a chunk whose first K index lies k_rel columns past the block's first column: B[n][k] == 0 for k > n,
// so the tile is live iff its last column (j WN + wn) 8 + 7 >= k_rel.  __host__ __device__: tests/cpp/test_sched.cu checks
// on the CPU that exactly the all-zero tiles are skipped.
template <int WN>
__host__ __device__ __forceinline__ int tri_first_live_tile(int k_rel, int wn) {
  const int need = k_rel / 8 - wn;
  return need <= 0 ? 0 : (need + WN - 1) / WN;
}

template <int J> struct IntC { static constexpr int value = J; };
// f(IntC<jmin>) for a run-time jmin in [0, BT): a chain of compares ending in a compile-time specialised body
template <int J, int BT, class F>
__device__ __forceinline__ void dispatch_jmin(int jmin, F&& f) {
  if constexpr (J < BT) {
    if (jmin == J) f(IntC<J>{});
    else dispatch_jmin<J + 1, BT>(jmin, f);
  }
}

// TRI = false is the dense kernel (every structure of the triangular schedule compiled out: it costs the dense schedule
// 1.6 % when merely present -- measured, tools/exp_time.py); TRI = true is launched when an operator is triangular.
template <int WM, int WN, int AT, int BT, int STAGES, int KC, bool TRI = false>
__global__ void __launch_bounds__(WM * WN * 32, 1)
toy_fused_kernel(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapS,
                 const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapD, ToyArgs g) {
  constexpr int kKC = KC;
  static_assert(KC % 16 == 8, "KC must be 8 (mod 16) doubles");
  constexpr int TM = WM * AT * 8;
  constexpr int NP = WN * BT * 8;
  constexpr int NCW = WM * WN;            // warps (all consumers)
  constexpr int NCT = NCW * 32;
  constexpr int ROWS = TM + NP;
  constexpr uint32_t STAGE_BYTES = ROWS * kKC * sizeof(double);
  static_assert((TM * kKC * sizeof(double)) % 128 == 0 && STAGE_BYTES % 128 == 0, "TMA destinations must be 128-byte aligned");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: STAGES x [ROWS][kKC] | full[STAGES] | done[STAGES] (int, padded) | sRow[2][WN][TM] | sAcc[WM][nstat][nblk*NP]
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  int* done = reinterpret_cast<int*>(bars + STAGES);
  double* sRowBase = reinterpret_cast<double*>(bars + 2 * STAGES + kRingSlots + 1);
  double* sAcc = sRowBase + 2 * WN * TM;

  const int tid = threadIdx.x, lane = tid & 31;
  // TRI: broadcast from lane 0 so that the compiler knows the warp index -- and the triangular-schedule branches that
  // depend on it -- to be warp-uniform (otherwise every mma.sync behind such a branch gets a WARPSYNC)
  const int warp = TRI ? __shfl_sync(0xffffffffu, tid >> 5, 0) : tid >> 5;
  const int nk = (g.ld + kKC - 1) / kKC;
  const int64_t ntiles = (g.T + TM - 1) / TM;
  const bool stats = g.part != nullptr;
  const int accw = g.nblk * NP;

  const int m = (int)(ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const int ring_row0 = blockIdx.x * kRingSlots * TM;

  // issue the TMA loads of the chunk at schedule position `p` into stage st (one thread)
  using SchedT = std::conditional_t<TRI, SchedTri<NP, kKC>, SchedDense>;
  auto issue = [&](const SchedT& p, int st) {
    const int kind = p.kind();
    if (kind == PH_NOP) { mbar_arrive(&full[st]); return; }   // keeps the stage's phase in step
    mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
    double* dst = stage0 + (size_t)st * ROWS * kKC;
    const int k0 = p.kc * kKC;
    const int lt = p.tile();
    if (kind == PH_P1) {
      const int64_t t0 = ((int64_t)blockIdx.x + (int64_t)lt * gridDim.x) * TM;
      tma_load_2d(dst, &mapV, &full[st], k0, (int)t0, kEvictFirst);
      tma_load_2d(dst + TM * kKC, &mapS, &full[st], k0, p.cb * NP, kEvictLast);
    } else {
      tma_load_2d(dst, &mapD, &full[st], k0, ring_row0 + (lt % kRingSlots) * TM, kEvictLast);
      tma_load_2d(dst + TM * kKC, &mapR, &full[st], k0, p.cb * NP, kEvictLast);
    }
  };

  SchedT pf;   // schedule position of the chunk STAGES ahead of the one this warp consumes
  pf.init(m, g.nblk, nk, STAGES - 1, g.do2, g.tri, g.ld);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (stats) {
    for (int q = tid; q < WM * g.nstat * accw; q += NCT) sAcc[q] = 0.0;
  }
  __syncthreads();   // the only block-wide barrier
  for (int s = 0; s < STAGES; ++s) {   // prologue: chunks 0 .. STAGES-1
    if (!pf.done()) {
      if (tid == 0) issue(pf, s);
      pf.advance();
    }
  }

  const int wm = warp / WN, wn = warp % WN;
  const int gq = lane >> 2, c = lane & 3;
  // column tile j of this warp within the CTA's column block: dealt cyclically, so that under the triangular schedule
  // every warp keeps the same share of the work below the diagonal (the dense kernel uses the same map: the two
  // schedules then sum in the same order and agree bit for bit)
  auto ctile = [&](int j) { return j * WN + wn; };
  double* ring = g.do2 ? g.Dring + (size_t)blockIdx.x * kRingSlots * TM * g.ld : nullptr;
  auto consumer_barrier = [&]() { asm volatile("bar.sync 1, %0;\n" ::"n"(NCT) : "memory"); };

  if (m > 0) {
    SchedT q;
    q.init(m, g.nblk, nk, STAGES - 1, g.do2, g.tri, g.ld);
    int s = 0, parity = 0;
    uint32_t ph = 0;
    double acc[AT][BT][2];
    while (!q.done()) {
      const int kind = q.kind();
      mbar_wait(&full[s], ph);
      if (kind != PH_NOP) {
      if (q.kc == 0) {
#pragma unroll
        for (int i = 0; i < AT; ++i)
#pragma unroll
          for (int j = 0; j < BT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      }
      const double* sA = stage0 + (size_t)s * ROWS * kKC;
      const double* sB = sA + TM * kKC;
      const int ksteps = min(kKC, g.ld - q.kc * kKC + 7) / 8;  // 8 physical k per double-step
      if constexpr (!TRI) {
#pragma unroll
        for (int kp = 0; kp < kKC / 8; ++kp) {
          if (kp < ksteps) {
            double2 a[AT], b[BT];
#pragma unroll
            for (int i = 0; i < AT; ++i)
              a[i] = *reinterpret_cast<const double2*>(sA + ((wm * AT + i) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
            for (int j = 0; j < BT; ++j)
              b[j] = *reinterpret_cast<const double2*>(sB + ((j * WN + wn) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
            for (int i = 0; i < AT; ++i)
#pragma unroll
              for (int j = 0; j < BT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
            for (int i = 0; i < AT; ++i)
#pragma unroll
              for (int j = 0; j < BT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
          }
        }
      } else {
        // Lower-triangular operator (B[n][k] == 0 for k > n): column tile j of this warp, columns
        // n0 + (j * WN + wn) * 8 .. + 8, has non-zeros in this chunk (k >= kc * KC) only if j * WN + wn >= kt0, the
        // chunk's first k-tile counted from column n0.  The fragment loads and DMMAs of the tiles j < jm are skipped -- by
        // real branches into a body specialised for jm: a predicated-off DMMA still occupies the pipe (measured, no gain)
        // and a per-k-step test costs more instructions than the zeros it saves (ncu: 60 % DMMA-busy against 75 %).
        int jm = 0;
        if ((g.tri >> (kind == PH_P2)) & 1) jm = tri_first_live_tile<WN>(q.kc * kKC - q.cb * NP, wn);
        auto body = [&](auto jmc) {
          constexpr int JM = decltype(jmc)::value;
#pragma unroll
          for (int kp = 0; kp < kKC / 8; ++kp) {
            if (kp < ksteps) {
              double2 a[AT], b[BT];
#pragma unroll
              for (int i = 0; i < AT; ++i)
                a[i] = *reinterpret_cast<const double2*>(sA + ((wm * AT + i) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
              for (int j = JM; j < BT; ++j)
                b[j] = *reinterpret_cast<const double2*>(sB + ((j * WN + wn) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
              for (int i = 0; i < AT; ++i)
#pragma unroll
                for (int j = JM; j < BT; ++j) dmma884<JM>(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
              for (int i = 0; i < AT; ++i)
#pragma unroll
                for (int j = JM; j < BT; ++j) dmma884<JM>(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            }
          }
        };
        // The chunk that CROSSES the diagonal of a column tile is where chunk-granular skipping wastes work: tile j is
        // live for the chunk's first 8 k but dead for its later ones (k > column).  For such chunks -- the first live tile
        // differs between the chunk's first and last k-step -- the k-steps are grouped by their first live tile and each
        // group runs the body specialised for it, over a run-time range of k-steps (round 2; round 1 executed 0.60 of the
        // dense DMMA count at N = 200 where 0.52 is needed).  All other chunks keep the fully unrolled body above.
        int jm_last = jm;
        const int kt0 = (q.kc * kKC - q.cb * NP) / 8;        // first k-tile of the chunk, counted from the block's first column (multiple-of-8 arithmetic: exact)
        if ((g.tri >> (kind == PH_P2)) & 1) jm_last = tri_first_live_tile<WN>(q.kc * kKC - q.cb * NP + 8 * (ksteps - 1), wn);
        if (jm_last == jm) {
          if (jm == 0) body(IntC<0>{});   // a dense pass of a half-triangular pair never reaches the jump table below
          else dispatch_jmin<1, BT>(jm, body);
        } else {
          auto body_range = [&](auto jmc, int kp_lo, int kp_hi) {
            constexpr int JM = decltype(jmc)::value;
            for (int kp = kp_lo; kp < kp_hi; ++kp) {
              double2 a[AT], b[BT];
#pragma unroll
              for (int i = 0; i < AT; ++i)
                a[i] = *reinterpret_cast<const double2*>(sA + ((wm * AT + i) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
              for (int j = JM; j < BT; ++j)
                b[j] = *reinterpret_cast<const double2*>(sB + ((j * WN + wn) * 8 + gq) * kKC + kp * 8 + 2 * c);
#pragma unroll
              for (int i = 0; i < AT; ++i)
#pragma unroll
                for (int j = JM; j < BT; ++j) dmma884<JM + 64>(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
              for (int i = 0; i < AT; ++i)
#pragma unroll
                for (int j = JM; j < BT; ++j) dmma884<JM + 64>(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            }
          };
          int kp = 0;
          for (int J = jm; J <= jm_last && J < BT && kp < ksteps; ++J) {
            // k-steps whose first live tile is J: k-tile kt0 + kp needs tile index j WN + wn >= kt0 + kp
            const int kp_hi = min(ksteps, J * WN + wn - kt0 + 1);
            if (kp_hi > kp) {
              if (J == 0) body_range(IntC<0>{}, kp, kp_hi);
              else dispatch_jmin<1, BT>(J, [&](auto jc) { body_range(jc, kp, kp_hi); });
              kp = kp_hi;
            }
          }
        }
      }

      if (q.last_chunk()) {
        const int lt = q.tile();
        const int64_t t0 = ((int64_t)blockIdx.x + (int64_t)lt * gridDim.x) * TM;
        const int n0 = q.cb * NP;
        if (kind == PH_P1) {
          // ------------------------------------------------------------ pass-1 epilogue
          double* Dt = g.do2 ? ring + (size_t)(lt % kRingSlots) * TM * g.ld : g.Dfull + t0 * g.ld;
#pragma unroll
          for (int j = 0; j < BT; ++j) {
            const int col = n0 + ctile(j) * 8 + 2 * c;
            const bool ok0 = col < g.N, ok1 = col + 1 < g.N;
            const double b00 = ok0 ? g.b0[col] : 0.0, b01 = ok1 ? g.b0[col + 1] : 0.0;
            double cs[2] = {0, 0}, cn[2] = {0, 0}, cr[2] = {0, 0}, cq[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < AT; ++i) {
              const int r = (wm * AT + i) * 8 + gq;
              const int64_t t = t0 + r;
              if (t < g.T) {
                const double v0 = acc[i][j][0], v1 = acc[i][j][1];
                const double d0 = v0 - b00, d1 = v1 - b01;
                if (ok1) {
                  *reinterpret_cast<double2*>(Dt + (size_t)r * g.ld + col) = make_double2(d0, d1);
                } else if (ok0) {
                  Dt[(size_t)r * g.ld + col] = d0;
                }
                if (g.Bout) {
                  if (ok0) g.Bout[t * g.N + col] = v0;
                  if (ok1) g.Bout[t * g.N + col + 1] = v1;
                }
                cs[0] += v0;
                cs[1] += v1;
                if (g.nstat == 4) {
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
              // Fixed-order reduction over the 8 toy lanes (gq = lane bits 2..4) as a REDUCE-SCATTER: the 2 (nstat = 1) or
              // 8 (nstat = 4) per-thread partial sums {stat, column e} are halved at every butterfly stage, so that after
              // the xor-16 / xor-8 / xor-4 stages each of the 8 lanes holds ONE complete sum -- 7 shuffles instead of 24
              // (3 instead of 6 for nstat = 1; the replicated all-reduce form cost the C5 test 9 %).  Lane gq owns slot
              // (wm, stat, column) of value index gq: no other thread of the CTA ever touches it, so the running sum
              // needs no barrier or atomic, and the summation order is fixed (bit-reproducible).
              const bool h4 = (lane & 16) != 0, h2 = (lane & 8) != 0, h1 = (lane & 4) != 0;
              double* pp = sAcc + (size_t)wm * g.nstat * accw;
              const int lc0 = n0 + ctile(j) * 8 + 2 * c;
              if (g.nstat == 4) {
                // v[k]: k = 4 e + stat
                double v[8] = {cs[0], cn[0], cr[0], cq[0], cs[1], cn[1], cr[1], cq[1]};
                double w[4], x[2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const double send = h4 ? v[k] : v[k + 4], keep = h4 ? v[k + 4] : v[k];
                  w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const double send = h2 ? w[k] : w[k + 2], keep = h2 ? w[k + 2] : w[k];
                  x[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                const double send = h1 ? x[0] : x[1], keep = h1 ? x[1] : x[0];
                const double y = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                const int e = h4 ? 1 : 0, stat = (h2 ? 2 : 0) + (h1 ? 1 : 0);
                pp[stat * accw + lc0 + e] += y;
              } else {
                const double send = h4 ? cs[0] : cs[1], keep = h4 ? cs[1] : cs[0];
                double y = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                y += __shfl_xor_sync(0xffffffffu, y, 8);
                y += __shfl_xor_sync(0xffffffffu, y, 4);
                if ((lane & 12) == 0) pp[lc0 + (h4 ? 1 : 0)] += y;
              }
            }
          }
          // order this warp's generic-proxy stores of d before the TMA (async proxy) reads of pass 2; the
          // stage hand-over below (fence + counter) then publishes them to whichever thread issues those loads
          if (g.do2) asm volatile("fence.proxy.async;\n" ::: "memory");
        } else {
          // ------------------------------------------------------------ pass-2 epilogue
          // chi2_t = sum_n C[t, n]^2: thread-local, then the 4 column lanes, then the WN warps (fixed order)
          double* sRow = sRowBase + (size_t)parity * WN * TM;
#pragma unroll
          for (int i = 0; i < AT; ++i) {
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < BT; ++j) sum += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (c == 0) sRow[wn * TM + (wm * AT + i) * 8 + gq] = sum;
          }
          consumer_barrier();
          for (int r = tid; r < TM; r += NCT) {
            const int64_t t = t0 + r;
            if (t < g.T) {
              double sum = 0.0;
#pragma unroll
              for (int w = 0; w < WN; ++w) sum += sRow[w * TM + r];
              if (q.cb == 0) g.chi2[t] = sum; else g.chi2[t] += sum;   // column blocks accumulate in order
            }
          }
          parity ^= 1;   // the next tile writes the other buffer: no second barrier needed
        }
      }
      }  // kind != PH_NOP

      // hand the stage back: the last warp to finish it refills it with the chunk STAGES ahead
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&done[s], 1) == NCW - 1) {
          done[s] = 0;
          __threadfence_block();
          if (!pf.done()) issue(pf, s);
        }
      }
      pf.advance();
      q.advance();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  }

  if (stats) {
    consumer_barrier();
    for (int q = tid; q < g.nstat * accw; q += NCT) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < WM; ++w) sum += sAcc[(size_t)w * g.nstat * accw + q];
      g.part[((size_t)blockIdx.x * 4 + q / accw) * g.part_ld + q % accw] = sum;
    }
  }
}

template <int WM, int WN, int AT, int BT, int STAGES, int KC>
constexpr size_t toy_fused_smem_base() {
  constexpr int kKC = KC;
  constexpr int TM = WM * AT * 8, NP = WN * BT * 8;
  return (size_t)STAGES * (TM + NP) * kKC * sizeof(double) + (2 * STAGES + kRingSlots + 1) * sizeof(uint64_t) +
         (size_t)2 * WN * TM * sizeof(double);
}

}  // namespace isr
