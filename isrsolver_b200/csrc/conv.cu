// Forward convolution service: sigma_vis(E) = int_{min_x}^{max_x} F(x, E^2) eps(x, E) f(E sqrt(1 - x)) dx
// for MANY energies at once and an arbitrary Born cross section f.
//
// Replaces convolutionKuraevFadin / convolutionKuraevFadin_1d (src/KuraevFadin.cpp:56-92; PyISR
// "convolutionKuraevFadin", include/PyKuraevFadin.hpp:19-62) and the callers that loop it over energies
// (src/isrsolver-KuraevFadin-convolution.cpp:122-134, src/ISRSolverVCSFitter.cpp:64-93,
// src/isrsolver-radcorr-function.cpp:152-155).
//
// f lives on the host (a Python / C++ callable), so the work is split at the only place it can be:
//   1. the plan cuts every row's [min_x, max_x] exactly as the matrix assembly does (x0, geometric grading
//      towards 0, x0 and 1, efficiency bin edges, plus caller-supplied break energies) and places a
//      15-point Gauss-Kronrod rule on every panel in the singularity-removing variable (x, x^beta or
//      (x - x0)^beta);
//   2. isr_conv_plan_nodes hands the node energies E' = E sqrt(1 - x) of the PENDING panels to the caller,
//      who samples f there (one vectorised call);
//   3. isr_conv_plan_feed multiplies by F eps dx/dv on the device and stores, per panel, the Kronrod value
//      and the QUADPACK-style error estimate from the embedded 7-point Gauss rule;
//   4. isr_conv_plan_refine bisects (in the integration variable) the panels whose error exceeds their
//      share of the row tolerance -- narrow resonances of f are resolved adaptively while smooth regions
//      stay at one panel -- and the loop returns to 2 with only the new panels pending.
// All per-node and per-panel work (transcendentals, reductions, compaction) runs on the GPU; the host sees
// node energies in and sums out.  Summation orders are fixed, so results are run-to-run reproducible.
#include <algorithm>
#include <cmath>
#include <new>

#include "gl_tables.cuh"
#include "panel_device.cuh"
#include "problem.hpp"

namespace isr {

struct ConvPanel {
  double lo, hi;          // bounds in the integration variable v (x, t = x^beta or u = (x - x0)^beta)
  double eps;             // efficiency on the panel (piecewise constant tables) or the row scale
  double val, err, aval;  // Kronrod estimate, error estimate, integral of |integrand|
  int row, kind, state, pad;   // kind as in assemble.cu; state 0 = pending, 1 = evaluated
};

}  // namespace isr

struct isr_conv_plan {
  isr_ctx* ctx = nullptr;
  int64_t M = 0;
  std::vector<double> energy;
  isr::DevBuf<isr::RowConst> d_row;
  isr::DevBuf<double> d_minx, d_brk;
  int nbrk = 0;
  isr::EpsDev eps{};
  isr::DevBuf<double> d_eps_vals, d_eps_edges;
  isr::DevBuf<isr::ConvPanel> d_pan[2];
  int cur = 0;
  int64_t npan = 0, npending = 0;
  isr::DevBuf<int> d_cnt, d_off, d_pend, d_rowoff;
  isr::DevBuf<double> d_nodeE, d_nodeX, d_f, d_w, d_res;
  isr::DevBuf<int> d_nodeRow;
};

namespace isr {

constexpr int kGK = 15;    // Kronrod nodes per panel; 16 node slots are exposed (the last one duplicates node 14)
constexpr int kSlots = 16;

// ---- setup ------------------------------------------------------------------------------------------
__global__ void conv_row_setup_kernel(int64_t M, const double* __restrict__ energy, const double* __restrict__ max_x,
                                      RowConst* __restrict__ rows) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  RowConst rc;
  row_const_init(rc, energy[m], 0.0);
  rc.xmax = max_x[m];
  row_grading(rc);
  rows[m] = rc;
}

// Walk the panels of row m: cuts at the grading points, the efficiency bin edges and the break energies
// (descending in E, i.e. ascending in x = 1 - (E_b / E)^2).
template <typename Emit>
__device__ __forceinline__ int conv_walk(const RowConst& rc, const EpsDev& eps, double xa, double xb, int nbrk,
                                         const double* __restrict__ brk, Emit emit) {
  if (!(xb > xa)) return 0;
  int count = 0;
  int ir = 0;
  while (ir < rc.nrb && !(rc.rb[ir] > xa)) ++ir;
  int ie = 0x7fffffff;
  if (eps.kind == 2) ie = eps_first_edge_above(eps, xa);
  auto brk_x = [&](int q) {
    const double r = brk[q] / rc.E;
    return 1.0 - r * r;
  };
  int ib = 0;
  while (ib < nbrk && !(brk_x(ib) > xa)) ++ib;
  double cur = xa;
  for (;;) {
    const double nr = ir < rc.nrb ? rc.rb[ir] : 2.0;
    const double ne = (eps.kind == 2 && ie <= eps.nx) ? eps_edge(eps, ie) : 2.0;
    const double nb = ib < nbrk ? brk_x(ib) : 2.0;
    const double nxt = fmin(nr, fmin(ne, nb));
    if (!(nxt < xb)) {
      emit(cur, xb);
      ++count;
      break;
    }
    if (nxt > cur) {
      emit(cur, nxt);
      ++count;
      cur = nxt;
    }
    if (nr <= nxt) ++ir;
    if (ne <= nxt) ++ie;
    while (ib < nbrk && !(brk_x(ib) > nxt)) ++ib;
  }
  return count;
}

__global__ void conv_count_kernel(int64_t M, const RowConst* __restrict__ rows, EpsDev eps,
                                  const double* __restrict__ min_x, int nbrk, const double* __restrict__ brk,
                                  int* __restrict__ cnt) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const RowConst& rc = rows[m];
  cnt[m] = conv_walk(rc, eps, fmax(min_x[m], 0.0), rc.xmax, nbrk, brk, [](double, double) {});
}

__global__ void conv_fill_kernel(int64_t M, const RowConst* __restrict__ rows, EpsDev eps,
                                 const double* __restrict__ min_x, int nbrk, const double* __restrict__ brk,
                                 const int* __restrict__ off, ConvPanel* __restrict__ pan) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const RowConst& rc = rows[m];
  int o = off[m];
  double scale = 1.0;
  if (eps.kind == 1) scale = eps.values[m];
  conv_walk(rc, eps, fmax(min_x[m], 0.0), rc.xmax, nbrk, brk, [&](double a, double b) {
    ConvPanel q;
    q.eps = scale;
    if (eps.kind == 2) q.eps = eps.values[(size_t)m * (eps.nx + 2) + eps_find_bin(eps, 0.5 * (a + b))];
    if (!(b > rc.x0)) {
      q.kind = 1;
      q.lo = a > 0.0 ? exp(rc.beta * log(a)) : 0.0;
      q.hi = exp(rc.beta * log(b));
    } else if (!(b > rc.xin)) {
      q.kind = 2;
      const double ya = a - rc.x0, yb = b - rc.x0;
      q.lo = ya > 0.0 ? exp(rc.beta * log(ya)) : 0.0;
      q.hi = exp(rc.beta * log(yb));
    } else {
      q.kind = 0;
      q.lo = a;
      q.hi = b;
    }
    q.val = q.err = q.aval = 0.0;
    q.row = (int)m;
    q.state = 0;
    q.pad = 0;
    pan[o++] = q;
  });
}

__global__ void iota_kernel(int64_t n, int* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = (int)q;
}

// ---- nodes ------------------------------------------------------------------------------------------
// x, ln x, the pair-threshold power and dx/dv (times the half width) at Kronrod node `lane` of a panel.
struct NodeGeom { double x, lnx, pw, jac; };
__device__ __forceinline__ NodeGeom conv_node(const RowConst& rc, const ConvPanel& p, double xi) {
  NodeGeom g;
  const double half = 0.5 * (p.hi - p.lo);
  const double v = fma(half, xi, 0.5 * (p.lo + p.hi));
  if (p.kind == 0) {
    g.x = v;
    g.lnx = log(v);
    g.pw = exp(rc.beta * log(v - rc.x0));
    g.jac = half;
  } else if (p.kind == 1) {
    g.lnx = log(v) * rc.inv_beta;
    g.x = exp(g.lnx);
    g.pw = -1.0;
    g.jac = half * rc.inv_beta * g.x / v;
  } else {
    const double y = exp(log(v) * rc.inv_beta);
    g.x = rc.x0 + y;
    g.lnx = log(g.x);
    g.pw = v;
    g.jac = half * rc.inv_beta * y / v;
  }
  return g;
}

// One 16-lane group per pending panel: node energies, x, row, and the full node weights
// w = wk * F(x, s) * eps * dx/dv * half  (so that  sum_nodes w f(E')  is the Kronrod value).
__global__ void __launch_bounds__(256)
conv_nodes_kernel(int64_t npend, const int* __restrict__ pend, const ConvPanel* __restrict__ pan,
                  const RowConst* __restrict__ rows, double* __restrict__ nodeE, double* __restrict__ nodeX,
                  int* __restrict__ nodeRow, double* __restrict__ nodeW) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pi = slot / kSlots;
  if (pi >= npend) return;
  const int lane = (int)(slot % kSlots);
  const int q = lane < kGK ? lane : kGK - 1;
  const ConvPanel p = pan[pend[pi]];
  const RowConst& rc = rows[p.row];
  const NodeGeom g = conv_node(rc, p, kGK15_x[q]);
  if (nodeE) nodeE[slot] = rc.E * sqrt(1.0 - g.x);
  if (nodeX) nodeX[slot] = g.x;
  if (nodeRow) nodeRow[slot] = p.row;
  if (nodeW) nodeW[slot] = lane < kGK ? kGK15_wk[q] * kf_value(rc, g.x, g.lnx, g.pw) * g.jac * p.eps : 0.0;
}

// QUADPACK's rescale of |Kronrod - Gauss| (qk.c: rescale_error)
__device__ __forceinline__ double rescale_error(double err, double resabs, double resasc) {
  err = fabs(err);
  if (resasc != 0.0 && err != 0.0) {
    const double scale = pow(200.0 * err / resasc, 1.5);
    err = scale < 1.0 ? resasc * scale : resasc;
  }
  if (resabs > 2.2250738585072014e-308 / (50.0 * 2.220446049250313e-16))
    err = fmax(err, 50.0 * 2.220446049250313e-16 * resabs);
  return err;
}

// Evaluate the pending panels from the sampled Born values f[slot].
__global__ void __launch_bounds__(256)
conv_feed_kernel(int64_t npend, const int* __restrict__ pend, ConvPanel* __restrict__ pan,
                 const RowConst* __restrict__ rows, const double* __restrict__ f) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pi = slot / kSlots;
  const bool live = pi < npend;
  const int lane = (int)(slot % kSlots);
  const int q = lane < kGK ? lane : kGK - 1;
  double gk = 0.0, gg = 0.0, ga = 0.0, gv = 0.0;
  int idx = 0;
  if (live) {
    idx = pend[pi];
    const ConvPanel p = pan[idx];
    const RowConst& rc = rows[p.row];
    const NodeGeom g = conv_node(rc, p, kGK15_x[q]);
    if (lane < kGK) {
      gv = kf_value(rc, g.x, g.lnx, g.pw) * g.jac * p.eps * f[slot];   // integrand * half width
      gk = kGK15_wk[q] * gv;
      gg = kGK15_wg[q] * gv;
      ga = kGK15_wk[q] * fabs(gv);
    }
  }
#pragma unroll
  for (int s = kSlots / 2; s >= 1; s >>= 1) {
    gk += __shfl_xor_sync(0xffffffffu, gk, s);
    gg += __shfl_xor_sync(0xffffffffu, gg, s);
    ga += __shfl_xor_sync(0xffffffffu, ga, s);
  }
  // resasc = int |g - mean|, mean = K / (b - a)  ->  in half-width units: |gv - gk / 2|
  double gs = (live && lane < kGK) ? kGK15_wk[q] * fabs(gv - 0.5 * gk) : 0.0;
#pragma unroll
  for (int s = kSlots / 2; s >= 1; s >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, s);
  if (live && lane == 0) {
    pan[idx].val = gk;
    pan[idx].aval = ga;
    pan[idx].err = rescale_error(gk - gg, ga, gs);
    pan[idx].state = 1;
  }
}

// ---- per-row reduction --------------------------------------------------------------------------------
// rowoff[m] = first panel of row m (panels are row-contiguous and rows ascend), rowoff[M] = npan
__global__ void conv_rowoff_kernel(int64_t npan, int64_t M, const ConvPanel* __restrict__ pan, int* __restrict__ rowoff) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > npan) return;
  const int prev = i == 0 ? -1 : pan[i - 1].row;
  const int here = i == npan ? (int)M : pan[i].row;
  for (int m = prev + 1; m <= here; ++m) rowoff[m] = (int)i;
}
// res[m] = {sum val, sum err, sum aval, n panels}: one warp per row, lane-strided partials + fixed xor tree
__global__ void conv_rowsum_kernel(int64_t M, const int* __restrict__ rowoff, const ConvPanel* __restrict__ pan,
                                   double* __restrict__ res) {
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  double v = 0, e = 0, a = 0;
  const int b = rowoff[m], en = rowoff[m + 1];
  for (int q = b + lane; q < en; q += 32) {
    v += pan[q].val;
    e += pan[q].err;
    a += pan[q].aval;
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, s);
    e += __shfl_xor_sync(0xffffffffu, e, s);
    a += __shfl_xor_sync(0xffffffffu, a, s);
  }
  if (lane == 0) {
    res[4 * m + 0] = v;
    res[4 * m + 1] = e;
    res[4 * m + 2] = a;
    res[4 * m + 3] = (double)(en - b);
  }
}

// ---- refinement ---------------------------------------------------------------------------------------
// A panel is bisected when its error exceeds its share of the row tolerance
//   tol_m = max(abs_tol, rel_tol |I_m|),  share = max(1 / n_m, |I_p| / sum_p |I_p|)
// (so the accepted errors of a row add up to at most 2 tol_m) and it is still wider than rounding.
__device__ __forceinline__ bool conv_wants_split(const ConvPanel& p, const double* __restrict__ res, double rel_tol,
                                                 double abs_tol) {
  if (p.state != 1) return false;
  const double* r = res + 4 * (size_t)p.row;
  const double tol = fmax(abs_tol, rel_tol * fabs(r[0]));
  const double share = fmax(1.0 / r[3], r[2] > 0.0 ? p.aval / r[2] : 0.0);
  if (!(p.err > tol * share)) return false;
  return (p.hi - p.lo) > 1e-13 * fmax(fabs(p.lo), fabs(p.hi));
}
__global__ void conv_mark_kernel(int64_t npan, const ConvPanel* __restrict__ pan, const double* __restrict__ res,
                                 double rel_tol, double abs_tol, int* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npan) return;
  cnt[i] = conv_wants_split(pan[i], res, rel_tol, abs_tol) ? 2 : 1;
}
__global__ void conv_split_kernel(int64_t npan, const ConvPanel* __restrict__ pan, const int* __restrict__ cnt,
                                  const int* __restrict__ off, ConvPanel* __restrict__ out, int* __restrict__ pend) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npan) return;
  const ConvPanel p = pan[i];
  const int o = off[i];
  if (cnt[i] == 1) {
    out[o] = p;
    return;
  }
  const double mid = 0.5 * (p.lo + p.hi);
  ConvPanel a = p, b = p;
  a.hi = mid;
  b.lo = mid;
  a.val = a.err = a.aval = b.val = b.err = b.aval = 0.0;
  a.state = b.state = 0;
  out[o] = a;
  out[o + 1] = b;
  const int ps = 2 * (o - (int)i);   // two pending entries per earlier split, in panel order
  pend[ps] = o;
  pend[ps + 1] = o + 1;
}
__global__ void conv_reset_kernel(int64_t npan, ConvPanel* __restrict__ pan, int* __restrict__ pend) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npan) return;
  pan[i].state = 0;
  pan[i].val = pan[i].err = pan[i].aval = 0.0;
  pend[i] = (int)i;
}

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static int conv_row_sums(isr_conv_plan* c) {
  cudaStream_t st = c->ctx->stream;
  ISR_TRY(c->d_rowoff.reserve(c->M + 1));
  ISR_TRY(c->d_res.reserve(4 * (size_t)c->M));
  conv_rowoff_kernel<<<blocks_for(c->npan + 1, 256), 256, 0, st>>>(c->npan, c->M, c->d_pan[c->cur].p, c->d_rowoff.p); ISR_COUNT_LAUNCH();
  conv_rowsum_kernel<<<blocks_for(c->M * 32, 256), 256, 0, st>>>(c->M, c->d_rowoff.p, c->d_pan[c->cur].p, c->d_res.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

}  // namespace isr

using namespace isr;

extern "C" {

int isr_conv_plan_create(isr_ctx* ctx, int64_t M, const double* energy, const double* min_x, const double* max_x,
                         const isr_eps_desc* eps, int nbreaks, const double* break_energy, isr_conv_plan** out) {
  if (!ctx || !energy || !min_x || !max_x || !out) { set_error("isr_conv_plan_create: NULL argument"); return ISR_EINVAL; }
  *out = nullptr;
  if (M < 1 || M > (1 << 24)) { set_error("isr_conv_plan_create: M = %lld out of range", (long long)M); return ISR_EINVAL; }
  if (nbreaks < 0 || (nbreaks > 0 && !break_energy)) { set_error("isr_conv_plan_create: bad break list"); return ISR_EINVAL; }
  for (int64_t m = 0; m < M; ++m) {
    if (!(energy[m] > 0.0) || !std::isfinite(energy[m])) { set_error("isr_conv_plan_create: energy[%lld] = %g", (long long)m, energy[m]); return ISR_EINVAL; }
    if (!(min_x[m] >= 0.0) || !(max_x[m] < 1.0) || !(max_x[m] >= min_x[m])) {
      set_error("isr_conv_plan_create: need 0 <= min_x <= max_x < 1 (row %lld: %g, %g)", (long long)m, min_x[m], max_x[m]);
      return ISR_EINVAL;
    }
  }
  if (eps && eps->kind == ISR_EPS_ROW_TABLE && M > 0x7fffffff / (eps->nx + 2)) { set_error("efficiency table too large"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(ctx->device));
  isr_conv_plan* c = new (std::nothrow) isr_conv_plan();
  if (!c) { set_error("out of host memory"); return ISR_EINVAL; }
  c->ctx = ctx;
  c->M = M;
  c->energy.assign(energy, energy + M);
  cudaStream_t st = ctx->stream;
  auto fail = [&](int rc) { delete c; return rc; };
  int rc = build_eps_rows(st, (int)M, energy, eps, c->d_eps_vals, c->d_eps_edges, c->eps);
  if (rc != ISR_OK) return fail(rc);
  DevBuf<double> dE, dmax;
  if ((rc = dE.reserve(M)) || (rc = dmax.reserve(M)) || (rc = c->d_minx.reserve(M)) || (rc = c->d_row.reserve(M)) ||
      (rc = c->d_cnt.reserve(M)) || (rc = c->d_off.reserve(M + 1)))
    return fail(rc);
  std::vector<double> brk(break_energy, break_energy + nbreaks);
  std::sort(brk.begin(), brk.end(), [](double a, double b) { return a > b; });   // descending E = ascending x
  c->nbrk = nbreaks;
  if ((rc = c->d_brk.reserve(std::max(nbreaks, 1)))) return fail(rc);
#define CONV_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("CUDA error '%s' (%s)", cudaGetErrorString(e__), #expr); return fail(ISR_ECUDA); } } while (0)
  CONV_CUDA(cudaMemcpyAsync(dE.p, energy, sizeof(double) * M, cudaMemcpyHostToDevice, st));
  CONV_CUDA(cudaMemcpyAsync(dmax.p, max_x, sizeof(double) * M, cudaMemcpyHostToDevice, st));
  CONV_CUDA(cudaMemcpyAsync(c->d_minx.p, min_x, sizeof(double) * M, cudaMemcpyHostToDevice, st));
  if (nbreaks) CONV_CUDA(cudaMemcpyAsync(c->d_brk.p, brk.data(), sizeof(double) * nbreaks, cudaMemcpyHostToDevice, st));
  conv_row_setup_kernel<<<blocks_for(M, 128), 128, 0, st>>>(M, dE.p, dmax.p, c->d_row.p); ISR_COUNT_LAUNCH();
  conv_count_kernel<<<blocks_for(M, 128), 128, 0, st>>>(M, c->d_row.p, c->eps, c->d_minx.p, c->nbrk, c->d_brk.p, c->d_cnt.p); ISR_COUNT_LAUNCH();
  if ((rc = exclusive_scan_int(st, (int)M, c->d_cnt.p, c->d_off.p))) return fail(rc);
  int total = 0;
  CONV_CUDA(cudaMemcpyAsync(&total, c->d_off.p + M, sizeof(int), cudaMemcpyDeviceToHost, st));
  CONV_CUDA(cudaStreamSynchronize(st));
  c->npan = total;
  if ((rc = c->d_pan[0].reserve(std::max<int64_t>(total, 1))) || (rc = c->d_pend.reserve(std::max<int64_t>(total, 1)))) return fail(rc);
  conv_fill_kernel<<<blocks_for(M, 128), 128, 0, st>>>(M, c->d_row.p, c->eps, c->d_minx.p, c->nbrk, c->d_brk.p, c->d_off.p, c->d_pan[0].p); ISR_COUNT_LAUNCH();
  if (total) { iota_kernel<<<blocks_for(total, 256), 256, 0, st>>>(total, c->d_pend.p); ISR_COUNT_LAUNCH(); }
  CONV_CUDA(cudaGetLastError());
  CONV_CUDA(cudaStreamSynchronize(st));   // dE / dmax go out of scope
#undef CONV_CUDA
  c->cur = 0;
  c->npending = total;
  *out = c;
  return ISR_OK;
}

void isr_conv_plan_destroy(isr_conv_plan* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  delete c;
}

int isr_conv_plan_size(isr_conv_plan* c, int64_t* n_panels, int64_t* n_pending_nodes) {
  if (!c) { set_error("plan == NULL"); return ISR_EINVAL; }
  if (n_panels) *n_panels = c->npan;
  if (n_pending_nodes) *n_pending_nodes = c->npending * kSlots;
  return ISR_OK;
}

int isr_conv_plan_nodes(isr_conv_plan* c, double* energy_out, double* x_out, int* row_out, double* weight_out) {
  if (!c) { set_error("plan == NULL"); return ISR_EINVAL; }
  const int64_t nn = c->npending * kSlots;
  if (nn == 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t st = c->ctx->stream;
  ISR_TRY(c->d_nodeE.reserve(nn));
  if (x_out) ISR_TRY(c->d_nodeX.reserve(nn));
  if (row_out) ISR_TRY(c->d_nodeRow.reserve(nn));
  if (weight_out) ISR_TRY(c->d_w.reserve(nn));
  conv_nodes_kernel<<<blocks_for(nn, 256), 256, 0, st>>>(c->npending, c->d_pend.p, c->d_pan[c->cur].p, c->d_row.p, c->d_nodeE.p,
                                                         x_out ? c->d_nodeX.p : nullptr, row_out ? c->d_nodeRow.p : nullptr,
                                                         weight_out ? c->d_w.p : nullptr); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  if (energy_out) ISR_CUDA(cudaMemcpyAsync(energy_out, c->d_nodeE.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
  if (x_out) ISR_CUDA(cudaMemcpyAsync(x_out, c->d_nodeX.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
  if (row_out) ISR_CUDA(cudaMemcpyAsync(row_out, c->d_nodeRow.p, sizeof(int) * nn, cudaMemcpyDeviceToHost, st));
  if (weight_out) ISR_CUDA(cudaMemcpyAsync(weight_out, c->d_w.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

static int conv_feed_device(isr_conv_plan* c) {
  cudaStream_t st = c->ctx->stream;
  const int64_t nn = c->npending * kSlots;
  conv_feed_kernel<<<blocks_for(nn, 256), 256, 0, st>>>(c->npending, c->d_pend.p, c->d_pan[c->cur].p, c->d_row.p, c->d_f.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  c->npending = 0;
  return conv_row_sums(c);
}

int isr_conv_plan_feed(isr_conv_plan* c, const double* fvals) {
  if (!c || !fvals) { set_error("NULL argument"); return ISR_EINVAL; }
  const int64_t nn = c->npending * kSlots;
  if (nn == 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(c->ctx->device));
  ISR_TRY(c->d_f.reserve(nn));
  ISR_CUDA(cudaMemcpyAsync(c->d_f.p, fvals, sizeof(double) * nn, cudaMemcpyHostToDevice, c->ctx->stream));
  ISR_TRY(conv_feed_device(c));
  ISR_CUDA(cudaStreamSynchronize(c->ctx->stream));
  return ISR_OK;
}

// f = the interpolated Born cross section of a solver (Interpolator::eval, src/Interpolator.cpp:295-323),
// sampled on the device: no host round trip (radiative-correction and iterative use).
int isr_conv_plan_feed_interp(isr_conv_plan* c, isr_problem* p, const double* bcs) {
  if (!c || !p || !bcs) { set_error("NULL argument"); return ISR_EINVAL; }
  if (p->ctx != c->ctx) { set_error("isr_conv_plan_feed_interp: plan and problem live on different contexts"); return ISR_EINVAL; }
  const int64_t nn = c->npending * kSlots;
  if (nn == 0) return ISR_OK;
  ISR_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t st = c->ctx->stream;
  ISR_TRY(c->d_f.reserve(nn));
  ISR_TRY(c->d_nodeE.reserve(nn));
  ISR_TRY(p->d_b0.reserve(p->n));
  p->h_b0.clear();   // d_b0 no longer mirrors the toy reference solution
  ISR_CUDA(cudaMemcpyAsync(p->d_b0.p, bcs, sizeof(double) * p->n, cudaMemcpyHostToDevice, st));
  conv_nodes_kernel<<<blocks_for(nn, 256), 256, 0, st>>>(c->npending, c->d_pend.p, c->d_pan[c->cur].p, c->d_row.p, c->d_nodeE.p,
                                                         nullptr, nullptr, nullptr); ISR_COUNT_LAUNCH();
  launch_interp_eval(p, p->d_b0.p, nn, c->d_nodeE.p, c->d_f.p);
  ISR_TRY(conv_feed_device(c));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int isr_conv_plan_refine(isr_conv_plan* c, double rel_tol, double abs_tol, int64_t* n_split) {
  if (!c) { set_error("plan == NULL"); return ISR_EINVAL; }
  if (c->npending) { set_error("isr_conv_plan_refine: %lld panels still pending (call isr_conv_plan_feed first)", (long long)c->npending); return ISR_ESTATE; }
  if (n_split) *n_split = 0;
  if (c->npan == 0) return ISR_OK;
  if (c->npan > 0x3fffffff) { set_error("isr_conv_plan_refine: panel list too long"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t st = c->ctx->stream;
  ISR_TRY(c->d_cnt.reserve(c->npan));
  ISR_TRY(c->d_off.reserve(c->npan + 1));
  conv_mark_kernel<<<blocks_for(c->npan, 256), 256, 0, st>>>(c->npan, c->d_pan[c->cur].p, c->d_res.p, rel_tol, abs_tol, c->d_cnt.p); ISR_COUNT_LAUNCH();
  ISR_TRY(exclusive_scan_int(st, (int)c->npan, c->d_cnt.p, c->d_off.p));
  int total = 0;
  ISR_CUDA(cudaMemcpyAsync(&total, c->d_off.p + c->npan, sizeof(int), cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  const int64_t nsplit = total - c->npan;
  if (n_split) *n_split = nsplit;
  if (nsplit == 0) return ISR_OK;
  const int nxt = c->cur ^ 1;
  ISR_TRY(c->d_pan[nxt].reserve((size_t)total + total / 2));
  ISR_TRY(c->d_pend.reserve((size_t)total + total / 2));
  conv_split_kernel<<<blocks_for(c->npan, 256), 256, 0, st>>>(c->npan, c->d_pan[c->cur].p, c->d_cnt.p, c->d_off.p, c->d_pan[nxt].p, c->d_pend.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  c->cur = nxt;
  c->npan = total;
  c->npending = 2 * nsplit;
  return ISR_OK;
}

int isr_conv_plan_result(isr_conv_plan* c, double* result, double* err_est) {
  if (!c || !result) { set_error("NULL argument"); return ISR_EINVAL; }
  if (c->npending) { set_error("isr_conv_plan_result: %lld panels still pending", (long long)c->npending); return ISR_ESTATE; }
  ISR_CUDA(cudaSetDevice(c->ctx->device));
  if (!c->d_res.p) ISR_TRY(conv_row_sums(c));
  std::vector<double> h(4 * (size_t)c->M);
  ISR_CUDA(cudaMemcpyAsync(h.data(), c->d_res.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->ctx->stream));
  ISR_CUDA(cudaStreamSynchronize(c->ctx->stream));
  for (int64_t m = 0; m < c->M; ++m) {
    result[m] = h[4 * m];
    if (err_est) err_est[m] = h[4 * m + 1];
  }
  return ISR_OK;
}

int isr_conv_plan_reset(isr_conv_plan* c) {
  if (!c) { set_error("plan == NULL"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(c->ctx->device));
  if (c->npan) {
    ISR_TRY(c->d_pend.reserve(c->npan));
    conv_reset_kernel<<<blocks_for(c->npan, 256), 256, 0, c->ctx->stream>>>(c->npan, c->d_pan[c->cur].p, c->d_pend.p); ISR_COUNT_LAUNCH();
    ISR_CUDA(cudaGetLastError());
  }
  c->npending = c->npan;
  return ISR_OK;
}

}  // extern "C"
