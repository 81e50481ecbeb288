// Toy Monte-Carlo batch: the loop bodies of chi2Test (src/Chi2Test.cpp:38-60), chi2TestData (:177-183)
// and ratioTestModel (src/RatioTest.cpp:38-45) for T toys at once.
//
// The reference re-runs solve() -- COD + (G^T W G)^-1, O(N^3) -- for every toy.  Here the operator is
// factored once (isr_factor) and a toy costs 4 N^2 flop:
//   pass 1  b_k = Sop v_k                 (DMMA tile kernel, epilogue: d = b - b0, column statistics)
//   pass 2  chi2_k = | Rop d_k |^2        (same tile kernel, epilogue: row sum of squares)
// Toys are processed in chunks sized so that the intermediate D = b - b0 stays L2-resident between the
// two passes (126 MB L2) and each launch is a whole number of waves over the SMs.
#include "problem.hpp"
#include "toy_batch.hpp"
#include "tikhonov.hpp"
#include "toy_gemm.cuh"

#include <cstdlib>

namespace isr {

// out[stat][j] (+)= sum over CTAs, fixed order => run-to-run deterministic
__global__ void reduce_part_kernel(int nctas, int ld, int n, const double* __restrict__ part,
                                   double* __restrict__ sum_bcs, double* __restrict__ ratio_stats) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s[4] = {0, 0, 0, 0};
  for (int cta = 0; cta < nctas; ++cta)
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] += part[((size_t)cta * 4 + q) * ld + j];
  if (sum_bcs) sum_bcs[j] += s[0];
  if (ratio_stats) {
    ratio_stats[3 * j + 0] += s[1];
    ratio_stats[3 * j + 1] += s[2];
    ratio_stats[3 * j + 2] += s[3];
  }
}

__global__ void sum_blocks_kernel(int64_t T, int nblk, const double* __restrict__ part, double* __restrict__ chi2) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[(size_t)b * T + t];
  chi2[t] = s;
}

// repack T x N (odd N) into T x ld with a zero pad column
__global__ void repack_kernel(int64_t T, int n, int ld, const double* __restrict__ in, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * ld) return;
  const int64_t t = idx / ld;
  const int j = (int)(idx % ld);
  out[idx] = j < n ? in[t * n + j] : 0.0;
}

struct ToyCfg {
  int tm, np, threads;
  size_t smem;
  void (*solve)(GemmArgs);
  void (*chi2)(GemmArgs);
  const char* name;
};

template <int WM, int WN, int AT, int BT, int ST>
static ToyCfg make_cfg(const char* name) {
  ToyCfg c;
  c.tm = WM * AT * 8;
  c.np = WN * BT * 8;
  c.threads = WM * WN * 32;
  c.smem = toy_gemm_smem<WM, WN, AT, BT, ST>();
  c.solve = toy_gemm_kernel<WM, WN, AT, BT, ST, EPI_SOLVE>;
  c.chi2 = toy_gemm_kernel<WM, WN, AT, BT, ST, EPI_CHI2>;
  c.name = name;
  return c;
}

static const std::vector<ToyCfg>& configs() {
  static const std::vector<ToyCfg> v = {
      make_cfg<4, 2, 4, 4, 3>("m128n64w8"),   make_cfg<8, 1, 2, 13, 3>("m128n104w8"),
      make_cfg<4, 4, 2, 4, 3>("m64n128w16"),  make_cfg<4, 5, 2, 5, 3>("m64n200w20"),
      make_cfg<4, 2, 2, 13, 3>("m64n208w8"),  make_cfg<4, 4, 2, 8, 3>("m64n256w16"),
      make_cfg<2, 8, 2, 8, 2>("m32n512w16"),
  };
  return v;
}

static const ToyCfg* pick_cfg(int n) {
  const char* force = std::getenv("ISR_TOY_CFG");
  const ToyCfg* best = nullptr;
  double best_cost = 1e300;
  for (const ToyCfg& c : configs()) {
    if (force && std::string(force) == c.name) return &c;
    const int nblk = (n + c.np - 1) / c.np;
    // padded work, a penalty per extra column block (V is re-read), a mild preference for tall tiles
    const double cost = (double)nblk * c.np * (1.0 + 0.05 * (nblk - 1)) * (1.0 + 2.0 / c.tm);
    if (cost < best_cost) { best_cost = cost; best = &c; }
  }
  return best;
}

const char* toy_cfg_name(int n) { return pick_cfg(n)->name; }

int toy_batch_device(isr_problem* p, int64_t T, const double* V_dev, const double* b0_dev, const ToyOut& out) {
  if (!p->solver_selected) {
    set_error("toy batch: no solver selected (call isr_factor / isr_tikhonov_select / isr_tsvd_select first)");
    return ISR_ESTATE;
  }
  if (T <= 0) return ISR_OK;
  const int n = p->n, ld = (n + 1) & ~1;
  cudaStream_t st = p->ctx->stream;
  const ToyCfg* cfg = pick_cfg(n);
  static thread_local const ToyCfg* configured = nullptr;
  if (configured != cfg) {
    ISR_CUDA(cudaFuncSetAttribute(cfg->solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
    ISR_CUDA(cudaFuncSetAttribute(cfg->chi2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
    configured = cfg;
  }
  const int nblk = (n + cfg->np - 1) / cfg->np;
  const int ctas = p->ctx->sm_count;
  // chunk: whole waves, D chunk <= 48 MB so that it survives in L2 between the passes
  int64_t wave = (int64_t)ctas * cfg->tm;
  int64_t waves = (48ll << 20) / (wave * ld * 8);
  if (waves < 1) waves = 1;
  const int64_t Tc = wave * waves;
  const bool want_stats = out.sum_bcs || out.ratio_stats || out.rhist;
  const size_t d_elems = (size_t)std::min<int64_t>(Tc, T) * ld;
  if (p->d_B.n < d_elems) {
    ISR_TRY(p->d_B.reserve(d_elems));
    ISR_CUDA(cudaMemsetAsync(p->d_B.p, 0, d_elems * sizeof(double), st));   // pad column stays zero
  }
  if (want_stats) ISR_TRY(p->d_acc.reserve((size_t)ctas * 4 * nblk * cfg->np));
  if (nblk > 1) ISR_TRY(p->d_chi2.reserve((size_t)nblk * std::min<int64_t>(Tc, T)));
  if (n != ld) ISR_TRY(p->d_V.reserve((size_t)std::min<int64_t>(Tc, T) * ld));

  for (int64_t t0 = 0; t0 < T; t0 += Tc) {
    const int64_t tc = std::min<int64_t>(Tc, T - t0);
    const double* A = V_dev + t0 * n;
    if (n != ld) {
      repack_kernel<<<(unsigned)((tc * ld + 255) / 256), 256, 0, st>>>(tc, n, ld, A, p->d_V.p); ISR_COUNT_LAUNCH();
      A = p->d_V.p;
    }
    const int64_t ntiles = (tc + cfg->tm - 1) / cfg->tm;
    dim3 grid((unsigned)std::min<int64_t>(ctas, ntiles), nblk);
    GemmArgs g{};
    g.A = A; g.lda = ld; g.T = tc; g.B = p->d_Sop.p; g.ldb = ld; g.N = n; g.K = ld;
    g.b0 = b0_dev; g.D = p->d_B.p; g.ldd = ld;
    g.Bout = out.bcs ? out.bcs + t0 * n : nullptr;
    g.part = want_stats ? p->d_acc.p : nullptr;
    g.part_ld = nblk * cfg->np;
    g.want_ratio = (out.ratio_stats || out.rhist) ? 1 : 0;
    g.rhist = out.rhist;
    cfg->solve<<<grid, cfg->threads, cfg->smem, st>>>(g); ISR_COUNT_LAUNCH();
    if (want_stats)
      reduce_part_kernel<<<(n + 127) / 128, 128, 0, st>>>((int)grid.x, g.part_ld, n, p->d_acc.p, out.sum_bcs, out.ratio_stats); ISR_COUNT_LAUNCH();
    if (out.chi2) {
      GemmArgs h{};
      h.A = p->d_B.p; h.lda = ld; h.T = tc; h.B = p->d_Rop.p; h.ldb = ld; h.N = n; h.K = ld;
      h.chi2 = nblk > 1 ? p->d_chi2.p : out.chi2 + t0;
      cfg->chi2<<<grid, cfg->threads, cfg->smem, st>>>(h); ISR_COUNT_LAUNCH();
      if (nblk > 1)
        sum_blocks_kernel<<<(unsigned)((tc + 255) / 256), 256, 0, st>>>(tc, nblk, p->d_chi2.p, out.chi2 + t0); ISR_COUNT_LAUNCH();
    }
  }
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

// E (T x ld) = V op^T - shift, op row-major [n][ld]: the tile kernel with a caller-supplied operator
// (used by the Tikhonov toy scan to project toys onto the generalised eigenbasis).
int project_device(isr_problem* p, int64_t T, const double* V_dev, const double* op, const double* shift_dev, double* E_dev) {
  const int n = p->n, ld = (n + 1) & ~1;
  cudaStream_t st = p->ctx->stream;
  const ToyCfg* cfg = pick_cfg(n);
  ISR_CUDA(cudaFuncSetAttribute(cfg->solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
  const int nblk = (n + cfg->np - 1) / cfg->np;
  const double* A = V_dev;
  if (n != ld) {
    ISR_TRY(p->d_V.reserve((size_t)T * ld));
    repack_kernel<<<(unsigned)((T * ld + 255) / 256), 256, 0, st>>>(T, n, ld, V_dev, p->d_V.p); ISR_COUNT_LAUNCH();
    A = p->d_V.p;
    ISR_CUDA(cudaMemsetAsync(E_dev, 0, sizeof(double) * T * ld, st));   // zero pad column
  }
  const int64_t ntiles = (T + cfg->tm - 1) / cfg->tm;
  dim3 grid((unsigned)std::min<int64_t>(p->ctx->sm_count, ntiles), nblk);
  GemmArgs g{};
  g.A = A; g.lda = ld; g.T = T; g.B = op; g.ldb = ld; g.N = n; g.K = ld;
  g.b0 = shift_dev; g.D = E_dev; g.ldd = ld;
  cfg->solve<<<grid, cfg->threads, cfg->smem, st>>>(g); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

}  // namespace isr
