// Toy Monte-Carlo batch: the loop bodies of chi2Test (src/Chi2Test.cpp:38-60), chi2TestData (:177-183)
// and ratioTestModel (src/RatioTest.cpp:38-45) for T toys at once.
//
// The reference re-runs solve() -- COD + (G^T W G)^-1, O(N^3) -- for every toy.  Here the operator is
// factored once (isr_factor) and a toy costs 4 N^2 flop in ONE fused launch (toy_gemm.cuh):
//   pass 1  b_k = Sop v_k ; d_k = b_k - b0 ; column statistics
//   pass 2  chi2_k = | Rop d_k |^2
// (about 2 N^2 when both operators are exactly lower triangular: the triangular schedule skips the structural zeros).
#include "problem.hpp"
#include "tikhonov.hpp"
#include "toy_batch.hpp"
#include "linalg.hpp"
#include "toy_gemm.cuh"

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace isr {

// out[stat][j] += sum over CTAs.  One warp per column; lane l sums CTAs l, l+32, ... in order and the
// lanes are combined by a fixed xor tree => run-to-run deterministic for a fixed grid.
__global__ void reduce_part_kernel(int nctas, int ld, int n, int nstat, const double* __restrict__ part,
                                   double* __restrict__ sum_bcs, double* __restrict__ ratio_stats) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  double s[4] = {0, 0, 0, 0};
  for (int cta = lane; cta < nctas; cta += 32)
    for (int q = 0; q < nstat; ++q) s[q] += part[((size_t)cta * 4 + q) * ld + j];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], sft);
  if (lane == 0) {
    if (sum_bcs) sum_bcs[j] += s[0];
    if (ratio_stats && nstat == 4) {
      ratio_stats[3 * j + 0] += s[1];
      ratio_stats[3 * j + 1] += s[2];
      ratio_stats[3 * j + 2] += s[3];
    }
  }
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
  int tm, np, threads, wm, kc, stages;
  size_t smem_base;   // stages + barriers + chi2 row scratch; statistics scratch is added per launch
  void (*kernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, ToyArgs);
  void (*kernel_tri)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, ToyArgs);   // triangular schedule
  const char* name;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D FP64 map over a row-major [rows][ld] matrix, box = box_rows x kc, no swizzle, zero OOB fill
static int make_map(CUtensorMap* m, const double* base, int64_t rows, int ld, int box_rows, int kc) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ISR_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %d matrix, box %d x %d", (int)r, (long long)rows, ld, box_rows, kc);
    return ISR_ECUDA;
  }
  return ISR_OK;
}

template <int WM, int WN, int AT, int BT, int ST, int KC = 24>
static ToyCfg make_cfg(const char* name) {
  ToyCfg c;
  c.tm = WM * AT * 8;
  c.np = WN * BT * 8;
  c.threads = WM * WN * 32;
  c.wm = WM;
  c.kc = KC;
  c.stages = ST;
  c.smem_base = toy_fused_smem_base<WM, WN, AT, BT, ST, KC>();
  c.kernel = toy_fused_kernel<WM, WN, AT, BT, ST, KC, false>;
  c.kernel_tri = toy_fused_kernel<WM, WN, AT, BT, ST, KC, true>;
  c.name = name;
  return c;
}

static const std::vector<ToyCfg>& configs() {
  // name = m<TM>n<NP>w<warps>[s<stages>][k<KC>]; defaults 3 stages, KC = 24.  The k40 variants trade the third
  // pipeline stage for 40-double chunks (fewer chunk boundaries); they are preferred whenever they fit.
  static const std::vector<ToyCfg> v = {
      make_cfg<4, 2, 4, 4, 3>("m128n64w8"),    make_cfg<8, 1, 2, 13, 3>("m128n104w8"),
      make_cfg<4, 4, 2, 4, 3>("m64n128w16"),   make_cfg<4, 5, 2, 5, 3>("m64n200w20"),
      make_cfg<4, 2, 2, 13, 3>("m64n208w8"),   make_cfg<4, 4, 2, 8, 3>("m64n256w16"),
      make_cfg<4, 4, 2, 8, 2>("m64n256w16s2"), make_cfg<8, 2, 2, 7, 3>("m128n112w16"),
      make_cfg<4, 5, 2, 5, 2>("m64n200w20s2"),         make_cfg<4, 4, 2, 4, 2>("m64n128w16s2"),   // smallest ring: large N with ratio statistics
      make_cfg<4, 5, 2, 5, 2, 40>("m64n200w20s2k40"),  make_cfg<4, 4, 2, 8, 2, 40>("m64n256w16s2k40"),
      make_cfg<8, 2, 2, 7, 2, 40>("m128n112w16s2k40"), make_cfg<4, 4, 2, 4, 2, 40>("m64n128w16s2k40"),
      make_cfg<4, 4, 2, 4, 3, 40>("m64n128w16k40"),   make_cfg<16, 1, 1, 13, 2, 40>("m128n104w16s2k40"),
      make_cfg<16, 1, 1, 13, 3, 40>("m128n104w16k40"),
      // column widths between 128 and 200 and 2 x 144 for N up to 288 (the N sweep of round 2 showed N = 136 ... 168
      // paying for 200 columns: 21.7 TFLOP/s at N = 136)
      make_cfg<4, 3, 2, 6, 2, 40>("m64n144w12s2k40"), make_cfg<4, 4, 2, 5, 2, 40>("m64n160w16s2k40"),
      make_cfg<8, 2, 2, 11, 2, 40>("m128n176w16s2k40"), make_cfg<4, 4, 2, 6, 2, 40>("m64n192w16s2k40"),
      // (a 64-toy tile for N <= 112, <8, 2, 1, 7>, was measured for short batches -- C2's 10^5 toys are 5.3 waves of
      // 128-toy tiles -- and is 9 % slower per tile: 173.7 against 171.1 us, no gain)
  };
  return v;
}

constexpr size_t kSmemLimit = 227 * 1024;

static size_t cfg_smem(const ToyCfg& c, int n, int nstat) {
  const int nblk = (n + c.np - 1) / c.np;
  return c.smem_base + (size_t)c.wm * nstat * nblk * c.np * sizeof(double);
}

// The choice depends on N and the requested statistics ONLY -- never on the batch size: a toy's chi2 must not depend on
// how the campaign was chunked or over how many GPUs it was sharded (different tile shapes sum in different orders), or
// the multi-GPU histograms would stop being bit-identical to the single-GPU one.  (A batch-size-aware chooser was tried
// for C2 -- 10^5 toys in 128-toy tiles are 5.3 waves of CTAs, paid as 6 -- and bought nothing: the 64-toy tile that
// quantises better is 9 % slower per tile.)
static const ToyCfg* pick_cfg(int n, int nstat) {
  const char* force = std::getenv("ISR_TOY_CFG");
  const ToyCfg* best = nullptr;
  double best_cost = 1e300;
  const int ld = (n + 1) & ~1;
  for (const ToyCfg& c : configs()) {
    if (cfg_smem(c, n, nstat) > kSmemLimit) continue;
    if (force && std::string(force) == c.name) return &c;
    const int nblk = (n + c.np - 1) / c.np;
    // cost model (fitted to the B200 measurements in DESIGN.md K5): DMMA work of the padded tile in 8-deep k-steps
    // plus ~0.3 k-step per chunk boundary; a mild preference for tall tiles; 8-warp tiles hide latency worse
    const int nk = (ld + c.kc - 1) / c.kc;
    double cost = (double)nblk * c.np * ((ld + 7) / 8 + 0.31 * nk) * (1.0 + 2.0 / c.tm);
    if (c.threads < 384) cost *= 1.10;        // 8 warps; measured: N=100 25.8 vs 24.9 TFLOP/s
    else if (c.threads < 512) cost *= 1.04;   // 12 warps; measured: N=144 30.3 vs N=160 (16 warps) 31.7 TFLOP/s, both exact fits
    if (c.stages < 3 && c.kc == 24) cost *= 1.001;   // tie-break: the deeper ring when nothing else differs
    if (cost < best_cost) { best_cost = cost; best = &c; }
  }
  return best;
}

const char* toy_cfg_name(int n, int nstat) {
  const ToyCfg* c = pick_cfg(n, nstat);
  return c ? c->name : nullptr;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) and size: it is a driver call on the path
// between the factorisation's certificate and the toy launch, where the GPU waits for the host
static int ensure_dynamic_smem(int device, const void* kernel, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[{device, kernel}];
  if (smem > have) {
    ISR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  return ISR_OK;
}

static int launch_fused(isr_problem* p, int64_t T, const double* V_dev, const double* Sop, const double* Rop,
                        const double* b0_dev, double* Dfull, const ToyOut& out) {
  const int n = p->n, ld = (n + 1) & ~1;
  cudaStream_t st = p->ctx->stream;
  const bool want_stats = out.sum_bcs || out.ratio_stats || out.rhist;
  const int nstat = (out.ratio_stats || out.rhist) ? 4 : 1;
  const ToyCfg* cfg = pick_cfg(n, want_stats ? nstat : 0);
  if (!cfg) { set_error("toy batch: N = %d does not fit any kernel configuration", n); return ISR_EINVAL; }
  const size_t smem = cfg_smem(*cfg, n, want_stats ? nstat : 0);
  const int nblk = (n + cfg->np - 1) / cfg->np;
  const int64_t ntiles = (T + cfg->tm - 1) / cfg->tm;
  const int ctas = (int)std::min<int64_t>(p->ctx->sm_count, ntiles);
  const double* A = V_dev;
  if (n != ld) {
    ISR_TRY(p->d_V.reserve((size_t)T * ld));
    repack_kernel<<<(unsigned)((T * ld + 255) / 256), 256, 0, st>>>(T, n, ld, V_dev, p->d_V.p); ISR_COUNT_LAUNCH();
    A = p->d_V.p;
  }
  ToyArgs g{};
  g.b0 = b0_dev; g.T = T; g.N = n; g.ld = ld; g.nblk = nblk; g.do2 = Rop ? 1 : 0;
  CUtensorMap mapV, mapS, mapR, mapD;
  ISR_TRY(make_map(&mapV, A, T, ld, cfg->tm, cfg->kc));
  ISR_TRY(make_map(&mapS, Sop, n, ld, cfg->np, cfg->kc));
  if (Rop) {
    const size_t ring = (size_t)p->ctx->sm_count * kRingSlots * cfg->tm * ld;
    if (p->d_B.n < ring) {
      ISR_TRY(p->d_B.reserve(ring));
      ISR_CUDA(cudaMemsetAsync(p->d_B.p, 0, ring * sizeof(double), st));   // pad column stays zero
    }
    g.Dring = p->d_B.p;
    ISR_TRY(make_map(&mapR, Rop, n, ld, cfg->np, cfg->kc));
    ISR_TRY(make_map(&mapD, p->d_B.p, (int64_t)p->ctx->sm_count * kRingSlots * cfg->tm, ld, cfg->tm, cfg->kc));
  } else {
    g.Dfull = Dfull;
    mapR = mapS;
    mapD = mapV;
  }
  g.Bout = out.bcs;
  g.chi2 = out.chi2;
  if (want_stats) {
    ISR_TRY(p->d_acc.reserve((size_t)p->ctx->sm_count * 4 * nblk * cfg->np));
    g.part = p->d_acc.p;
    g.part_ld = nblk * cfg->np;
  }
  g.nstat = nstat;
  g.rhist = out.rhist;
  // a kernel PARAMETER, so that the schedule arithmetic stays in the uniform datapath (read from device memory it cost
  // the dense schedule 1.2 %); the flags describe the selected pair only
  g.tri = 0;
  if (p->toy_variant != 1) g.tri = (Sop == p->d_Sop.p ? p->tri_sel[0] : 0) | ((Rop && Rop == p->d_Rop.p) ? p->tri_sel[1] << 1 : 0);
  auto kernel = g.tri ? cfg->kernel_tri : cfg->kernel;
  ISR_TRY(ensure_dynamic_smem(p->ctx->device, (const void*)kernel, smem));
  kernel<<<ctas, cfg->threads, smem, st>>>(mapV, mapS, mapR, mapD, g); ISR_COUNT_LAUNCH();
  if (want_stats) {
    reduce_part_kernel<<<(n + 3) / 4, 128, 0, st>>>(ctas, g.part_ld, n, nstat, p->d_acc.p, out.sum_bcs, out.ratio_stats); ISR_COUNT_LAUNCH();
  }
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

// Schedule 2 (isr_set_toy_kernel): chi2 = |R d|^2 only depends on R^T R, so a dense chi2 factor is replaced -- once
// per selection, at the first toy batch that asks for chi2 -- by the lower-triangular factor L of its QL decomposition
// (|L d| == |R d| for every d): pass 2 then skips the structural zeros, N^2 instead of 2 N^2 flop per toy.  Already
// triangular factors (linear interpolation without energy spread) are left as they are.  Opt-in, because cuSOLVER's
// Householder QR costs 0.6 ms at N = 200 (1.8 ms at N = 500): it pays from about 10^6 toys per factorisation.
static int ensure_rop_triangular(isr_problem* p) {
  const int n = p->n, ld = (n + 1) & ~1;
  if (p->tri_sel[1] || n < 16) return ISR_OK;
  ISR_TRY(p->d_tmp.reserve((size_t)n * n));
  ISR_TRY(ql_factor_rowmajor(p->ctx, n, ld, p->d_Rop.p, p->d_tmp.p, p->d_work, p->d_info));
  p->tri_sel[1] = 1;
  return ISR_OK;
}

int toy_batch_device(isr_problem* p, int64_t T, const double* V_dev, const double* b0_dev, const ToyOut& out) {
  if (!p->solver_selected) {
    set_error("toy batch: no solver selected (call isr_factor / isr_tikhonov_select / isr_tsvd_select first)");
    return ISR_ESTATE;
  }
  if (T <= 0) return ISR_OK;
  if (out.chi2 && p->toy_variant == 2) ISR_TRY(ensure_rop_triangular(p));
  // pass 2 is skipped when no chi2 is requested (e.g. the mean-solution loop of chi2TestData)
  double* dfull = nullptr;
  if (!out.chi2) {
    dfull = scratch_dfull(p, T);
    if (!dfull) return ISR_ECUDA;      // (the allocation's message is in isr_last_error)
  }
  return launch_fused(p, T, V_dev, p->d_Sop.p, out.chi2 ? p->d_Rop.p : nullptr, b0_dev, dfull, out);
}

// pass-1-only launches still write d = b - b0 somewhere: a scratch T x ld buffer
double* scratch_dfull(isr_problem* p, int64_t T) {
  const int ld = (p->n + 1) & ~1;
  if (p->d_E.reserve((size_t)T * ld) != ISR_OK) return nullptr;
  return p->d_E.p;
}

// E (T x ld) = V op^T - shift, op row-major [n][ld]: the tile kernel with a caller-supplied operator
// (used by the Tikhonov toy scan to project toys onto the generalised eigenbasis).
int project_device(isr_problem* p, int64_t T, const double* V_dev, const double* op, const double* shift_dev, double* E_dev) {
  const int n = p->n, ld = (n + 1) & ~1;
  if (n != ld) ISR_CUDA(cudaMemsetAsync(E_dev, 0, sizeof(double) * T * ld, p->ctx->stream));   // zero pad column
  ToyOut none;
  return launch_fused(p, T, V_dev, op, nullptr, shift_dev, E_dev, none);
}

}  // namespace isr
