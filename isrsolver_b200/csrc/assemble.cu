// Matrix assembly: G(i,j) = int F(x, s_i) * eps_i(x) * basis_j(E_i sqrt(1-x)) dx.
//
// Replaces ISRSolverSLE::evalEqMatrix (src/ISRSolverSLE.cpp:138-149) and the N^2 adaptive GSL
// integrations below it (src/Interpolator.cpp:437-453 -> src/LinearRangeInterpolator.cpp:57-101 /
// src/CSplineRangeInterpolator.cpp:88-110 -> src/KuraevFadin.cpp:56-73 -> src/Integration.cpp:19-81).
//
// Design (DESIGN.md "Assembly"): per (row i, energy segment k <= i) the x-interval
// [1-(ext[k+1]/E_i)^2, 1-(ext[k]/E_i)^2] is cut at the pair threshold x0 = 4 m_e / E_i, at the
// efficiency bin edges and at a few geometric grading points; every resulting panel gets one
// 16-point Gauss-Legendre rule -- in x (regular panels), in t = x^beta (panels in [0, x0]: removes
// the x^(beta-1) and log(x) endpoint singularities) or in u = (x-x0)^beta (the panel touching x0 from
// above: removes the (x-x0)^beta infinite slope of the pair term).  Each panel yields the four
// normalised moments  M_p = int F eps ((E'-ext[k])/h_k)^p dx,  p = 0..3;  G is then the small dense
// contraction  G(i,j) = sum_{k<=i} sum_p M[i,k,p] * C[j,k,p]  with C the per-segment polynomial
// coefficients of basis j (hat functions or GSL-natural cardinal cubic splines).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "gl_tables.cuh"
#include "kf_device.cuh"
#include "panel_device.cuh"
#include "basis_host.hpp"
#include "problem.hpp"

namespace isr {

// =================================================================================================
// Host: basis coefficient tables
// =================================================================================================

int set_ranges(isr_problem* p, int nranges, const int* ranges) {
  std::vector<int> sorted;
  ISR_TRY(normalise_ranges(p->n, nranges, ranges, sorted));
  p->ranges = sorted;
  return build_basis_tables(p);
}

// Segment k = (ext[k], ext[k+1]] lies in exactly one range r (lo_r <= k <= hi_r).  That range decides
// the polynomial piece of every basis function on the segment:
//   linear range : rising edge of node j = k (delta/h), falling edge of node j = k-1 (1 - delta/h)
//                  (src/LinearRangeInterpolator.cpp:25-39, 70-101 incl. the (min,max] guards :74-76, :92-94)
//   cspline range: the cardinal natural spline of every node j in [max(lo-1,0), hi] -- hasCSIndex,
//                  src/BaseRangeInterpolator.cpp:35-38, 80-100 -- restricted to the segment
//                  (src/CSplineRangeInterpolator.cpp:88-110).
int build_basis_tables(isr_problem* p) {
  const int n = p->n;
  const std::vector<double>& x = p->ext;
  compute_basis_tables(n, x, p->ranges, p->h_Ct, p->h_klo, p->h_khi, p->seg_range);
  std::vector<double> invh(n);
  for (int k = 0; k < n; ++k) invh[k] = 1.0 / (x[k + 1] - x[k]);
  cudaStream_t st = p->ctx->stream;
  ISR_TRY(p->d_ext.reserve(n + 1));
  ISR_TRY(p->d_invh.reserve(n));
  ISR_TRY(p->d_Ct.reserve((size_t)n * 4 * n));
  ISR_TRY(p->d_klo.reserve(n));
  ISR_TRY(p->d_khi.reserve(n));
  ISR_CUDA(cudaMemcpyAsync(p->d_ext.p, x.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(p->d_invh.p, invh.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(p->d_Ct.p, p->h_Ct.data(), sizeof(double) * p->h_Ct.size(), cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(p->d_klo.p, p->h_klo.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(p->d_khi.p, p->h_khi.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  p->has_G = false;
  p->factored = false;
  p->solver_selected = false;
  p->tik_ready = false;
  p->svd_ready = false;
  return ISR_OK;
}

// =================================================================================================
// Host: efficiency -> per-row device form
// =================================================================================================

// TAxis::FindBin on a non-extendable axis (ROOT): under/overflow aware; fixed or variable bins.
static int host_find_bin(int nbins, double lo, double hi, const double* edges, double v) {
  if (v < lo) return 0;
  if (!(v < hi)) return nbins + 1;
  if (!edges) return 1 + (int)(nbins * (v - lo) / (hi - lo));
  int l = 0, h = nbins;
  while (h - l > 1) {
    const int m = (l + h) / 2;
    if (edges[m] <= v) l = m; else h = m;
  }
  return 1 + l;
}

// Reduce an efficiency descriptor to the per-row device form for `n` energies (rows): used by the
// problem (rows = the energy points) and by the convolution plans (rows = the requested energies).
int build_eps_rows(cudaStream_t st, int n, const double* energy, const isr_eps_desc* e, DevBuf<double>& d_vals,
                   DevBuf<double>& d_edges, EpsDev& out) {
  out = EpsDev{};
  if (!e || e->kind == ISR_EPS_ONE) return ISR_OK;
  if (e->kind < 0 || e->kind > ISR_EPS_E_ONLY) {
    set_error("[!] Wrong efficiency dimension.");   // EfficiencyDimensionException
    return ISR_EEFFDIM;
  }
  if (!e->values) { set_error("efficiency values == NULL"); return ISR_EINVAL; }
  double elo = e->elo, ehi = e->ehi, xlo = e->xlo, xhi = e->xhi;
  if (e->eedges) { elo = e->eedges[0]; ehi = e->eedges[e->ne]; }
  if (e->xedges) { xlo = e->xedges[0]; xhi = e->xedges[e->nx]; }
  if (e->kind == ISR_EPS_E_ONLY) {
    // TEfficiency 1-D: efficiency(x, E) = eff[FindFixBin(E)] (src/BaseISRSolver.cpp:236-241)
    if (e->ne < 1) { set_error("efficiency: ne < 1"); return ISR_EINVAL; }
    std::vector<double> scale(n);
    for (int i = 0; i < n; ++i) scale[i] = e->values[host_find_bin(e->ne, elo, ehi, e->eedges, energy[i])];
    ISR_TRY(d_vals.reserve(n));
    ISR_CUDA(cudaMemcpyAsync(d_vals.p, scale.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    ISR_CUDA(cudaStreamSynchronize(st));
    out.kind = 1;
    out.values = d_vals.p;
    return ISR_OK;
  }
  if (e->nx < 1 || !(xhi > xlo)) { set_error("efficiency: bad x axis"); return ISR_EINVAL; }
  const int w = e->nx + 2;
  std::vector<double> vals((size_t)n * w);
  if (e->kind == ISR_EPS_ROW_TABLE) {
    // v2: eff(x, index) = hist[index]->GetBinContent(FindBin(x)) (src/BaseISRSolver2.cpp:231-236)
    std::memcpy(vals.data(), e->values, sizeof(double) * vals.size());
  } else {
    // TEfficiency 2-D: global bin = bx + (nx+2)*by (src/BaseISRSolver.cpp:243-250)
    if (e->ne < 1) { set_error("efficiency: ne < 1"); return ISR_EINVAL; }
    for (int i = 0; i < n; ++i) {
      const int by = host_find_bin(e->ne, elo, ehi, e->eedges, energy[i]);
      std::memcpy(&vals[(size_t)i * w], &e->values[(size_t)by * w], sizeof(double) * w);
    }
  }
  ISR_TRY(d_vals.reserve(vals.size()));
  ISR_CUDA(cudaMemcpyAsync(d_vals.p, vals.data(), sizeof(double) * vals.size(), cudaMemcpyHostToDevice, st));
  out.kind = 2;
  out.nx = e->nx;
  out.xlo = xlo;
  out.xhi = xhi;
  out.values = d_vals.p;
  out.xedges = nullptr;
  if (e->xedges) {
    ISR_TRY(d_edges.reserve(e->nx + 1));
    ISR_CUDA(cudaMemcpyAsync(d_edges.p, e->xedges, sizeof(double) * (e->nx + 1), cudaMemcpyHostToDevice, st));
    out.xedges = d_edges.p;
  }
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int upload_eps(isr_problem* p, const isr_eps_desc* e) {
  p->panels_ready = false;
  return build_eps_rows(p->ctx->stream, p->n, p->ext.data() + 1, e, p->d_eps_vals, p->d_eps_edges, p->eps);
}

// =================================================================================================
// Device: panel enumeration
// =================================================================================================

__global__ void row_setup_kernel(int n, const double* __restrict__ ext, RowConst* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RowConst rc;
  row_const_init(rc, ext[i + 1], ext[0]);
  row_grading(rc);
  rows[i] = rc;
}

// Panel count of every item + exclusive scan WITHIN the block (local[]) + the block's total (blocksum[]): the
// global offsets follow from the few hundred block totals (every fill_kernel block sums the prefix it needs), so no
// kernel ever walks all N(N+1)/2 counts serially.
constexpr int kCountThreads = 256;
__global__ void __launch_bounds__(kCountThreads)
count_kernel(int nitems, const RowConst* __restrict__ rows, EpsDev eps, const double* __restrict__ ext,
             int* __restrict__ cnt, int* __restrict__ local, int* __restrict__ blocksum) {
  __shared__ int wsum[kCountThreads / 32];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int c = 0;
  if (idx < nitems) {
    int i, k;
    item_decode(idx, i, k);
    c = item_cuts(rows[i], eps, ext, i, k).count();
    cnt[idx] = c;
  }
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int q = 0; q < kCountThreads / 32; ++q) before += q < w ? wsum[q] : 0;
  if (idx < nitems) local[idx] = before + incl - c;
  if (threadIdx.x == kCountThreads - 1) blocksum[blockIdx.x] = before + incl;
}
// Single-block exclusive scan of cnt[0..n) into off[0..n], off[n] = total.  Tiles of 8 x 1024 elements go
// through shared memory: coalesced loads, an 8-element serial scan per thread out of shared memory, a
// shuffle-based block scan of the 1024 thread totals, coalesced stores (no dependent global-load chains).
constexpr int kScanPer = 8, kScanTile = kScanPer * 1024;
__global__ void __launch_bounds__(1024) scan_kernel(int n, const int* __restrict__ cnt, int* __restrict__ off) {
  __shared__ int tile[kScanTile];
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kScanTile) {
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) {
      const int idx = base + q * 1024 + t;
      tile[q * 1024 + t] = idx < n ? cnt[idx] : 0;
    }
    __syncthreads();
    int v[kScanPer], sum = 0;
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) { v[q] = tile[t * kScanPer + q]; sum += v[q]; }
    int incl = sum;                      // inclusive scan of the thread totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += u;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int x = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += u;
      }
      wsum[lane] = x;                    // inclusive scan of the warp totals
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + (w ? wsum[w - 1] : 0) + incl - sum;   // exclusive prefix of this thread's first element
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) { tile[t * kScanPer + q] = run; run += v[q]; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) {
      const int idx = base + q * 1024 + t;
      if (idx < n) off[idx] = tile[q * 1024 + t];
    }
    if (t == 1023) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (t == 0) off[n] = carry_s;
}

int exclusive_scan_int(cudaStream_t st, int n, const int* cnt_dev, int* off_dev) {
  scan_kernel<<<1, 1024, 0, st>>>(n, cnt_dev, off_dev); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

constexpr int kHeavyItem = 32;   // items with more panels than this are filled by a whole warp
// Thread per item: global offset off[idx] = local prefix + block offset (written for EVERY item), then the item's
// panels by a serial merge -- one to three for most items.  Items with more than kHeavyItem panels (the wide
// segments of a row under a finely binned efficiency) are left to fill_heavy_kernel.
__global__ void __launch_bounds__(kCountThreads)
fill_kernel(int nitems, const RowConst* __restrict__ rows, EpsDev eps, const double* __restrict__ ext,
            const int* __restrict__ cnt, const int* __restrict__ local, const int* __restrict__ blocksum, int* __restrict__ off,
            double* __restrict__ pa, double* __restrict__ pb, int* __restrict__ pitem, int* __restrict__ pkind, int skip_heavy) {
  // offset of this block's items = sum of the totals of the blocks before it: every block adds up that prefix of the
  // few hundred block totals itself (integer sum, any order), which saves the separate scan launch
  __shared__ int s_part[kCountThreads / 32];
  __shared__ int s_blockoff;
  {
    int acc = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kCountThreads) acc += blocksum[b];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
#pragma unroll
      for (int q = 0; q < kCountThreads / 32; ++q) t += s_part[q];
      s_blockoff = t;
      if (blockIdx.x == gridDim.x - 1) off[nitems] = t + blocksum[blockIdx.x];   // grand total
    }
    __syncthreads();
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nitems) return;
  const int o = local[idx] + s_blockoff;
  off[idx] = o;
  const int count = cnt[idx];
  if (count == 0 || (skip_heavy && count > kHeavyItem)) return;
  int i, k;
  item_decode(idx, i, k);
  const RowConst& rc = rows[i];
  const ItemCuts c = item_cuts(rc, eps, ext, i, k);
  walk_panels(c, rc, eps, 0, count, [&](int q, double a, double b) {
    pa[o + q] = a;
    pb[o + q] = b;
    pitem[o + q] = idx;
    pkind[o + q] = panel_kind(rc, b);
  });
}
// Warp per item, only for the heavy ones: each lane fills a contiguous run of panels, starting from the merge
// state selected directly at its first cut.
__global__ void __launch_bounds__(256)
fill_heavy_kernel(int nitems, const RowConst* __restrict__ rows, EpsDev eps, const double* __restrict__ ext,
                  const int* __restrict__ cnt, const int* __restrict__ off, double* __restrict__ pa, double* __restrict__ pb,
                  int* __restrict__ pitem, int* __restrict__ pkind) {
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= nitems) return;
  const int count = cnt[idx];
  if (count <= kHeavyItem) return;
  int i, k;
  item_decode(idx, i, k);
  const RowConst& rc = rows[i];
  const ItemCuts c = item_cuts(rc, eps, ext, i, k);
  const int o = off[idx], per = (count + 31) / 32;
  walk_panels(c, rc, eps, lane * per, min(count, (lane + 1) * per), [&](int q, double a, double b) {
    pa[o + q] = a;
    pb[o + q] = b;
    pitem[o + q] = idx;
    pkind[o + q] = panel_kind(rc, b);
  });
}

// =================================================================================================
// Device: panel evaluation (the hot kernel of the assembly)
// =================================================================================================
// One 16-lane group per panel, lane = Gauss-Legendre node; four moments reduced with warp shuffles.
constexpr int kQ = 16;
constexpr int kEvalThreads = 256;

// node_eps != nullptr: efficiency sampled by the host at every node (ISR_EPS_SAMPLED), [npanels][kQ]
__global__ void __launch_bounds__(kEvalThreads)
eval_kernel(const int* __restrict__ off_total, const RowConst* __restrict__ rows, EpsDev eps,
            const double* __restrict__ ext, const double* __restrict__ invh,
            const double* __restrict__ pa, const double* __restrict__ pb, double* __restrict__ peps,
            const int* __restrict__ pitem, const int* __restrict__ pkind, const double* __restrict__ node_eps,
            double* __restrict__ pm) {
  __shared__ double s_x[kQ], s_w[kQ];
  if (threadIdx.x < kQ) {
    s_x[threadIdx.x] = kGL16_x[threadIdx.x];
    s_w[threadIdx.x] = kGL16_w[threadIdx.x];
  }
  __syncthreads();
  const int npanels = *off_total;
  const int lane = threadIdx.x & (kQ - 1);
  const int groups_per_block = kEvalThreads / kQ;
  const int group = blockIdx.x * groups_per_block + (threadIdx.x / kQ);
  const int ngroups = gridDim.x * groups_per_block;
  const double t = s_x[lane], w = s_w[lane];
  for (int pidx = group; pidx < npanels; pidx += ngroups) {
    const double a = pa[pidx], b = pb[pidx];
    const int item = pitem[pidx], kind = pkind[pidx];
    int i, k;
    item_decode(item, i, k);
    const RowConst& rc = rows[i];
    // efficiency: constant on a panel (ROOT bin of its midpoint / the row's scale), or sampled per node by the host
    double ev = 1.0;
    if (eps.kind == 2) ev = eps.values[(size_t)i * (eps.nx + 2) + eps_find_bin(eps, 0.5 * (a + b))];
    else if (eps.kind == 1) ev = eps.values[i];
    if (lane == 0) peps[pidx] = ev;      // kept for isr_debug_panels
    if (node_eps) ev = node_eps[(size_t)pidx * kQ + lane];
    const PanelNode q = panel_node(rc, a, b, kind, t, w);
    const double x = q.x;
    const double f = kf_value(rc, x, q.lnx, q.pw) * q.wj * ev;
    const double dn = (rc.E * sqrt(1.0 - x) - ext[k]) * invh[k];
    double m0 = f, m1 = f * dn, m2 = m1 * dn, m3 = m2 * dn;
#pragma unroll
    for (int sft = kQ / 2; sft >= 1; sft >>= 1) {
      m0 += __shfl_xor_sync(0xffffffffu, m0, sft);
      m1 += __shfl_xor_sync(0xffffffffu, m1, sft);
      m2 += __shfl_xor_sync(0xffffffffu, m2, sft);
      m3 += __shfl_xor_sync(0xffffffffu, m3, sft);
    }
    if (lane == 0) {
      double4 v = make_double4(m0, m1, m2, m3);
      reinterpret_cast<double4*>(pm)[pidx] = v;
    }
  }
}

// x and row index of every quadrature node of the current panel list (for host-sampled efficiencies)
__global__ void __launch_bounds__(kEvalThreads)
nodes_kernel(const int* __restrict__ off_total, const RowConst* __restrict__ rows, const double* __restrict__ pa,
             const double* __restrict__ pb, const int* __restrict__ pitem, const int* __restrict__ pkind,
             double* __restrict__ node_x, int* __restrict__ node_row) {
  const int npanels = *off_total;
  const int lane = threadIdx.x & (kQ - 1);
  const int pidx = blockIdx.x * (kEvalThreads / kQ) + (threadIdx.x / kQ);
  if (pidx >= npanels) return;
  int i, k;
  item_decode(pitem[pidx], i, k);
  const PanelNode q = panel_node(rows[i], pa[pidx], pb[pidx], pkind[pidx], kGL16_x[lane], kGL16_w[lane]);
  node_x[(size_t)pidx * kQ + lane] = q.x;
  node_row[(size_t)pidx * kQ + lane] = i;
}

// M[i][k][0..3] = sum of the item's panel moments; zero for k > i.  One warp per (i, k): lanes stride over the
// panels, fixed xor tree => deterministic whatever the panel count.
__global__ void __launch_bounds__(256) reduce_kernel(int n, const int* __restrict__ off, const double* __restrict__ pm,
                                                      double* __restrict__ M) {
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // over n*n (i,k)
  const int lane = threadIdx.x & 31;
  if (idx >= n * n) return;
  const int i = idx / n, k = idx % n;
  double4 acc = make_double4(0, 0, 0, 0);
  if (k <= i) {
    const int item = i * (i + 1) / 2 + k;
    for (int q = off[item] + lane; q < off[item + 1]; q += 32) {
      const double4 v = reinterpret_cast<const double4*>(pm)[q];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, s);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, s);
    }
  }
  if (lane == 0) reinterpret_cast<double4*>(M)[idx] = acc;
}

// G(i,j) = sum_{k = klo_j}^{min(khi_j, i)} sum_p M[i][k][p] * Ct[k][p][j]   (column-major output), with the panel
// reduction fused in: M[i][k][p] = sum of the item's panel moments pm[q][p], q in [off[item], off[item+1]) in panel
// order (deterministic).  One block = 16 rows x 64 columns, 256 threads (thread = column x 4 rows).  Segments are
// processed in chunks of 16: the chunk's moments (16 x 64 doubles) and coefficients (64 x 64 doubles) are staged in
// shared memory with independent coalesced loads, then a fully unrolled 64-deep FMA loop runs out of shared memory
// (Ct is zero outside a column's band, so no per-column bounds are needed inside the chunk).  Only the segment range
// the block's columns actually touch ([min klo, min(max khi, last row)]) is visited: a purely linear matrix costs two
// chunks per block, a cubic one N/16 -- split over blockIdx.z (contract_slices() slices of the segment range, partial sums
// combined in fixed order by contract_sum_kernel) so that N = 200 still fills the machine.
constexpr int kCtRows = 16, kCtCols = 64, kCtK = 16, kCtMaxSplit = 16;
// Slices of the segment range (gridDim.z): about four blocks per SM in all, at most one slice per 16-segment chunk --
// a block that handles a single chunk pays the latency of its dependent loads (panel offsets -> moments, coefficients)
// once, and the slices run side by side (N = 200 cubic: 12 slices, 624 blocks; 4 slices took 20 us).
static int contract_slices(int n, int sm_count, bool all_linear) {
  // a block of 64 hat-function columns touches 65 segments (5 chunks), one with a spline column all of them
  const int tiles = ((n + kCtCols - 1) / kCtCols) * ((n + kCtRows - 1) / kCtRows);
  const int nchunks = all_linear ? std::min(5, (n + kCtK - 1) / kCtK) : (n + kCtK - 1) / kCtK;
  return std::max(1, std::min({kCtMaxSplit, nchunks, (4 * sm_count + tiles - 1) / tiles}));
}
// FUSED = false: the moments were reduced beforehand by reduce_kernel into pm = M[i][k][4] (tabulated efficiencies:
// dozens of panels per item, which every column tile would otherwise re-reduce).
template <bool FUSED>
__global__ void __launch_bounds__(256)
contract_kernel(int n, const int* __restrict__ off, const double* __restrict__ pm, const double* __restrict__ Ct,
                const int* __restrict__ klo, const int* __restrict__ khi, double* __restrict__ Gpart) {
  __shared__ double sM[kCtRows][kCtK * 4];
  __shared__ double sC[kCtK * 4][kCtCols];
  __shared__ int s_lo, s_hi;
  const int tx = threadIdx.x & (kCtCols - 1);
  const int j = blockIdx.x * kCtCols + tx;
  const int rsub = threadIdx.x / kCtCols;  // 0..3, each thread does 4 rows
  const int i0 = blockIdx.y * kCtRows;
  const int imax = min(n - 1, i0 + kCtRows - 1);
  if (threadIdx.x == 0) { s_lo = n; s_hi = -1; }
  __syncthreads();
  if (j < n) { atomicMin(&s_lo, klo[j]); atomicMax(&s_hi, khi[j]); }
  __syncthreads();
  // this block's slice of the block-wide segment range, in whole chunks
  int blo = s_lo, bhi = min(s_hi, imax);
  {
    const int nchunks = bhi >= blo ? (bhi - blo) / kCtK + 1 : 0, per = (nchunks + (int)gridDim.z - 1) / (int)gridDim.z;
    const int a = blo + (int)blockIdx.z * per * kCtK;
    bhi = min(bhi, a + per * kCtK - 1);
    blo = a;
  }
  double acc[4] = {0, 0, 0, 0};
  for (int k0 = blo; k0 <= bhi; k0 += kCtK) {
    __syncthreads();
    // moments of (row, segment) pairs of this chunk: 16 rows x 16 segments, one thread each
    {
      const int r = threadIdx.x / kCtK, kk = threadIdx.x % kCtK;
      const int i = i0 + r, k = k0 + kk;
      double4 m = make_double4(0, 0, 0, 0);
      if (i < n && k <= i && k <= bhi) {
        if (FUSED) {
          const int item = i * (i + 1) / 2 + k;
          for (int q = off[item]; q < off[item + 1]; ++q) {
            const double4 v = reinterpret_cast<const double4*>(pm)[q];
            m.x += v.x; m.y += v.y; m.z += v.z; m.w += v.w;
          }
        } else {
          m = reinterpret_cast<const double4*>(pm)[(size_t)i * n + k];
        }
      }
      sM[r][kk * 4 + 0] = m.x; sM[r][kk * 4 + 1] = m.y; sM[r][kk * 4 + 2] = m.z; sM[r][kk * 4 + 3] = m.w;
    }
    // coefficients: rows q = (k - k0) * 4 + p of Ct, 64 columns (coalesced along j)
#pragma unroll
    for (int it = 0; it < (kCtK * 4 * kCtCols) / 256; ++it) {
      const int e = it * 256 + threadIdx.x;
      const int q = e / kCtCols, c = e % kCtCols;
      const int k = k0 + q / 4, jj = blockIdx.x * kCtCols + c;
      sC[q][c] = (k <= bhi && jj < n) ? Ct[((size_t)k * 4 + (q & 3)) * n + jj] : 0.0;
    }
    __syncthreads();
#pragma unroll 16
    for (int q = 0; q < kCtK * 4; ++q) {
      const double cv = sC[q][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fma(sM[rsub * 4 + r][q], cv, acc[r]);
    }
  }
  if (j < n) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + rsub * 4 + r;
      if (i < n) Gpart[(size_t)blockIdx.z * n * n + (size_t)i + (size_t)n * j] = acc[r];
    }
  }
}
// G = sum of the slices' partial results, in slice order.  (Letting the slice block that arrives last at a tile do
// this -- __threadfence + arrival counter -- was measured: the fences make the contraction 18 us slower than this
// 4 us launch.)
__global__ void contract_sum_kernel(int nn, int nslices, const double* __restrict__ Gpart, double* __restrict__ G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nn) return;
  double s = 0.0;
  for (int z = 0; z < nslices; ++z) s += Gpart[(size_t)z * nn + idx];   // fixed order
  G[idx] = s;
}

// =================================================================================================
// Device: energy spread  G <- G_spread * G   (src/ISRSolverSLE.cpp:146-148, 177-193)
// =================================================================================================
// S(i,j) = sum_m (w_m/sqrt(pi)) * basis_j(E_i + sqrt(2) sigma_i t_m): 6-point Gauss-Hermite
// (gaussian_conv, src/Integration.cpp:86-100).  Column-major output.
__global__ void spread_kernel(int n, const double* __restrict__ ext, const double* __restrict__ invh,
                              const double* __restrict__ Ct, const double* __restrict__ sigE,
                              double* __restrict__ S) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int j = idx % n, i = idx / n;
  const double e0 = ext[i + 1], sg = sigE[i] * 1.4142135623730951;
  double acc = 0.0;
#pragma unroll
  for (int m = 0; m < 6; ++m)
    acc += kGH6_w_over_sqrtpi[m] * basis_eval_dev(n, ext, invh, Ct, j, e0 + sg * kGH6_t[m], 0);
  S[(size_t)i + (size_t)n * j] = acc;
}

// The same for a batch of spread settings (condition-number scan): S_k from sigma_k (one common value per setting, or
// one per point), matrices back to back.
__global__ void spread_batch_kernel(int n, const double* __restrict__ ext, const double* __restrict__ invh,
                                    const double* __restrict__ Ct, const double* __restrict__ sig, int per_point,
                                    double* __restrict__ Sbatch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int k = blockIdx.y;
  const int j = idx % n, i = idx / n;
  const double sigma = per_point ? sig[(size_t)k * n + i] : sig[k];
  const double e0 = ext[i + 1], sg = sigma * 1.4142135623730951;
  double acc = 0.0;
  if (sigma == 0.0) {
    acc = (i == j) ? 1.0 : 0.0;      // spread disabled for this setting: S = I (the product below returns G0 itself)
  } else {
#pragma unroll
    for (int m = 0; m < 6; ++m)
      acc += kGH6_w_over_sqrtpi[m] * basis_eval_dev(n, ext, invh, Ct, j, e0 + sg * kGH6_t[m], 0);
  }
  Sbatch[(size_t)k * n * n + (size_t)i + (size_t)n * j] = acc;
}
int spread_batch(isr_problem* p, cudaStream_t st, int K, const double* sig_dev, int per_point, double* Sbatch) {
  const int n = p->n;
  spread_batch_kernel<<<dim3((n * n + 127) / 128, K), 128, 0, st>>>(n, p->d_ext.p, p->d_invh.p, p->d_Ct.p, sig_dev, per_point, Sbatch); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

// C = A * B, all n x n column-major, FP64; 32x32 tiles.  Used for the one-off spread product.
__global__ void __launch_bounds__(256) gemm_nn_kernel(int n, const double* __restrict__ A,
                                                       const double* __restrict__ B, double* __restrict__ C) {
  __shared__ double sA[32][33], sB[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7, 4 rows each
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  double acc[4] = {0, 0, 0, 0};
  for (int k0 = 0; k0 < n; k0 += 32) {
    for (int q = ty; q < 32; q += 8) {
      const int ar = row0 + tx, ac = k0 + q;
      sA[q][tx] = (ar < n && ac < n) ? A[(size_t)ar + (size_t)n * ac] : 0.0;   // sA[k][row]
      const int br = k0 + tx, bc = col0 + q;
      sB[q][tx] = (br < n && bc < n) ? B[(size_t)br + (size_t)n * bc] : 0.0;   // sB[col][k]
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const double a = sA[k][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fma(a, sB[ty * 4 + r][k], acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = row0 + tx, col = col0 + ty * 4 + r;
    if (row < n && col < n) C[(size_t)row + (size_t)n * col] = acc[r];
  }
}

// Gout = G_spread(sigE) * G0 (all column-major n x n on the device); S is scratch.
int apply_spread(isr_problem* p, cudaStream_t st, const double* sigE_dev, const double* G0, double* S, double* Gout) {
  const int n = p->n;
  spread_kernel<<<(n * n + 127) / 128, 128, 0, st>>>(n, p->d_ext.p, p->d_invh.p, p->d_Ct.p, sigE_dev, S); ISR_COUNT_LAUNCH();
  dim3 g((n + 31) / 32, (n + 31) / 32);
  gemm_nn_kernel<<<g, 256, 0, st>>>(n, S, G0, Gout); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

// =================================================================================================
// Host driver
// =================================================================================================
static int64_t panel_upper_bound(const isr_problem* p) {
  const int64_t n = p->n;
  int64_t per_row_extra = kMaxRowBreaks + 2;
  if (p->eps.kind == 2) per_row_extra += p->eps.nx + 2;
  return n * (n + 1) / 2 + n * per_row_extra;
}

// Per-(row, segment) moments M[i][k][4] = sum of the item's panel moments: used when items have many panels (tabulated
// efficiency) and by isr_debug_moments; otherwise the contraction sums the panels on the fly.
int reduce_moments(isr_problem* p) {
  const int n = p->n;
  ISR_TRY(p->d_M.reserve((size_t)n * n * 4));
  reduce_kernel<<<(n * n + 7) / 8, 256, 0, p->ctx->stream>>>(n, p->d_off.p, p->d_pm.p, p->d_M.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;
}

// Panel enumeration (row constants, count, scan, fill): everything before the integrand is touched.
// The panel set depends only on the energies and on the efficiency's bin edges -- both fixed when the problem is created
// (the interpolation ranges only enter the contraction) -- so, like the reference's Interpolator, which is built in the
// solver's constructor and not in evalEqMatrix (src/ISRSolverSLE.cpp:54-58 vs :138-149), it is enumerated once per problem;
// every isr_assemble still evaluates the integrand at every node and redoes the contraction.
static int enumerate_panels(isr_problem* p) {
  if (p->panels_ready) return ISR_OK;
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  const int nitems = n * (n + 1) / 2;
  const int64_t cap = panel_upper_bound(p);
  if (cap > p->max_panels || !p->d_pa.p) {
    ISR_TRY(p->d_pa.reserve(cap));
    ISR_TRY(p->d_pb.reserve(cap));
    ISR_TRY(p->d_peps.reserve(cap));
    ISR_TRY(p->d_pitem.reserve(cap));
    ISR_TRY(p->d_pkind.reserve(cap));
    ISR_TRY(p->d_pm.reserve(cap * 4));
    p->max_panels = cap;
  }
  const int nb = (nitems + kCountThreads - 1) / kCountThreads;
  ISR_TRY(p->d_row.reserve(n));
  ISR_TRY(p->d_cnt.reserve(nitems));
  ISR_TRY(p->d_off.reserve(nitems + 1));
  ISR_TRY(p->d_local.reserve(nitems));
  ISR_TRY(p->d_blocksum.reserve(nb));
  row_setup_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, p->d_ext.p, p->d_row.p); ISR_COUNT_LAUNCH();
  count_kernel<<<nb, kCountThreads, 0, st>>>(nitems, p->d_row.p, p->eps, p->d_ext.p, p->d_cnt.p, p->d_local.p, p->d_blocksum.p); ISR_COUNT_LAUNCH();
  const int heavy = p->eps.kind == 2;     // only a tabulated efficiency can give an item more than 1 + kMaxRowBreaks panels
  fill_kernel<<<nb, kCountThreads, 0, st>>>(nitems, p->d_row.p, p->eps, p->d_ext.p, p->d_cnt.p, p->d_local.p, p->d_blocksum.p,
                                            p->d_off.p, p->d_pa.p, p->d_pb.p, p->d_pitem.p, p->d_pkind.p, heavy); ISR_COUNT_LAUNCH();
  if (heavy) {
    fill_heavy_kernel<<<(nitems + 7) / 8, 256, 0, st>>>(nitems, p->d_row.p, p->eps, p->d_ext.p, p->d_cnt.p, p->d_off.p,
                                                        p->d_pa.p, p->d_pb.p, p->d_pitem.p, p->d_pkind.p); ISR_COUNT_LAUNCH();
  }
  ISR_CUDA(cudaGetLastError());
  p->panels_ready = true;
  return ISR_OK;
}

// Integrand evaluation, reduction, basis contraction and the optional spread product.
static int finish_assembly(isr_problem* p, const double* ecm_err, const double* node_eps_dev) {
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  const int nitems = n * (n + 1) / 2;
  ISR_TRY(p->d_G.reserve((size_t)n * n));
  bool all_linear = true;
  for (size_t r = 0; r + 2 < p->ranges.size(); r += 3) all_linear = all_linear && p->ranges[r] == 0;
  const int nslices = contract_slices(n, p->ctx->sm_count, all_linear);
  ISR_TRY(p->d_Gpart.reserve((size_t)nslices * n * n));
  const int eval_blocks = p->ctx->sm_count * 8;
  eval_kernel<<<eval_blocks, kEvalThreads, 0, st>>>(p->d_off.p + nitems, p->d_row.p, p->eps, p->d_ext.p, p->d_invh.p,
                                                     p->d_pa.p, p->d_pb.p, p->d_peps.p, p->d_pitem.p,
                                                     p->d_pkind.p, node_eps_dev, p->d_pm.p); ISR_COUNT_LAUNCH();
  dim3 cgrid((n + kCtCols - 1) / kCtCols, (n + kCtRows - 1) / kCtRows, nslices);
  double* part = nslices == 1 ? p->d_G.p : p->d_Gpart.p;   // a single slice is the result itself
  if (p->eps.kind == 2) {   // many panels per item: reduce once, then contract
    ISR_TRY(reduce_moments(p));
    contract_kernel<false><<<cgrid, 256, 0, st>>>(n, p->d_off.p, p->d_M.p, p->d_Ct.p, p->d_klo.p, p->d_khi.p, part); ISR_COUNT_LAUNCH();
  } else {
    contract_kernel<true><<<cgrid, 256, 0, st>>>(n, p->d_off.p, p->d_pm.p, p->d_Ct.p, p->d_klo.p, p->d_khi.p, part); ISR_COUNT_LAUNCH();
  }
  if (nslices > 1) { contract_sum_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(n * n, nslices, p->d_Gpart.p, p->d_G.p); ISR_COUNT_LAUNCH(); }
  if (ecm_err) {
    for (int i = 0; i < n; ++i)
      if (!(ecm_err[i] > 0.0)) {
        set_error("energy spread enabled but ecm_err[%d] = %g is not positive (src/Integration.cpp:89 divides by it)", i, ecm_err[i]);
        return ISR_EINVAL;
      }
    ISR_TRY(p->d_sigE.reserve(n));
    ISR_TRY(p->d_S.reserve((size_t)n * n));
    ISR_TRY(p->d_tmp.reserve((size_t)n * n));
    ISR_CUDA(cudaMemcpyAsync(p->d_sigE.p, ecm_err, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    ISR_TRY(apply_spread(p, st, p->d_sigE.p, p->d_G.p, p->d_S.p, p->d_tmp.p));
    ISR_CUDA(cudaMemcpyAsync(p->d_G.p, p->d_tmp.p, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
  }
  ISR_CUDA(cudaGetLastError());
  p->has_G = true;
  p->G_injected = false;
  p->G_lower = !ecm_err;   // linear ranges only (LinearRangeInterpolator returns 0 for j > i, src/LinearRangeInterpolator.cpp:60-62), no spread product
  for (size_t r = 0; r + 2 < p->ranges.size() && p->G_lower; r += 3) p->G_lower = p->ranges[r] == 0;
  p->factored = false;
  p->solver_selected = false;
  p->tik_ready = false;
  p->svd_ready = false;
  p->last_panels = -1;  // fetched lazily by isr_assemble_stats
  return ISR_OK;
}

int assemble(isr_problem* p, const double* ecm_err) {
  ISR_TRY(enumerate_panels(p));
  return finish_assembly(p, ecm_err, nullptr);
}

// ---- host-sampled efficiency (ISR_EPS_SAMPLED): eps(x, E_i) is an arbitrary host callable ------------------
// Phase 1: enumerate the panels (cut at x0 and the grading points, plus the bin edges of a table given at problem
// creation -- the way to tell the quadrature where a sampled efficiency jumps) and return x and the row of every node.  Phase 2: the caller's values enter the integrand node by node.
int assemble_nodes(isr_problem* p, int64_t cap, double* x_out, int* row_out, int64_t* count) {
  const int n = p->n;
  cudaStream_t st = p->ctx->stream;
  const int nitems = n * (n + 1) / 2;
  // a tabulated efficiency may be present: its bin edges then serve as break points of the sampled one (the sampled
  // values replace the table's; pass a table of ones whose edges are the callable's discontinuities)
  ISR_TRY(enumerate_panels(p));
  int total = 0;
  ISR_CUDA(cudaMemcpyAsync(&total, p->d_off.p + nitems, sizeof(int), cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  const int64_t nn = (int64_t)total * kQ;
  p->sampled_panels = total;
  if (count) *count = nn;
  if (!x_out && !row_out) return ISR_OK;
  if (cap < nn) { set_error("isr_assemble_nodes: buffer holds %lld nodes, %lld needed", (long long)cap, (long long)nn); return ISR_EINVAL; }
  ISR_TRY(p->d_node_x.reserve(nn));
  ISR_TRY(p->d_node_row.reserve(nn));
  nodes_kernel<<<(total + kEvalThreads / kQ - 1) / (kEvalThreads / kQ), kEvalThreads, 0, st>>>(
      p->d_off.p + nitems, p->d_row.p, p->d_pa.p, p->d_pb.p, p->d_pitem.p, p->d_pkind.p, p->d_node_x.p, p->d_node_row.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  if (x_out) ISR_CUDA(cudaMemcpyAsync(x_out, p->d_node_x.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
  if (row_out) ISR_CUDA(cudaMemcpyAsync(row_out, p->d_node_row.p, sizeof(int) * nn, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int assemble_sampled(isr_problem* p, int64_t nn, const double* eps_values, const double* ecm_err) {
  if (p->sampled_panels < 0 || nn != (int64_t)p->sampled_panels * kQ) {
    set_error("isr_assemble_sampled: %lld values given, the node list of the last isr_assemble_nodes has %lld",
              (long long)nn, (long long)(p->sampled_panels < 0 ? 0 : (int64_t)p->sampled_panels * kQ));
    return ISR_ESTATE;
  }
  ISR_TRY(p->d_node_eps.reserve(nn > 0 ? nn : 1));
  ISR_CUDA(cudaMemcpyAsync(p->d_node_eps.p, eps_values, sizeof(double) * nn, cudaMemcpyHostToDevice, p->ctx->stream));
  return finish_assembly(p, ecm_err, p->d_node_eps.p);
}

// ---- device helpers exported to capi.cu ---------------------------------------------------------
__global__ void kf_plain_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ s,
                                double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = kf_plain(x[q], s[q]);
}
void launch_kf_plain(int64_t n, const double* x, const double* s, double* out, cudaStream_t st) {
  kf_plain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, x, s, out); ISR_COUNT_LAUNCH();
}

__global__ void basis_eval_kernel(int n, const double* __restrict__ ext, const double* __restrict__ invh,
                                  const double* __restrict__ Ct, int64_t m, const int* __restrict__ js,
                                  const double* __restrict__ es, int deriv, double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < m) out[q] = basis_eval_dev(n, ext, invh, Ct, js[q], es[q], deriv);
}
void launch_basis_eval(const isr_problem* p, int64_t m, const int* js, const double* es, int deriv, double* out) {
  basis_eval_kernel<<<(unsigned)((m + 255) / 256), 256, 0, p->ctx->stream>>>(p->n, p->d_ext.p, p->d_invh.p,
                                                                             p->d_Ct.p, m, js, es, deriv, out); ISR_COUNT_LAUNCH();
}

// Interpolator::eval (src/Interpolator.cpp:295-323): sum_j y_j basis_j(E); the value at E_max beyond it.
__global__ void interp_eval_kernel(int n, const double* __restrict__ ext, const double* __restrict__ invh,
                                   const double* __restrict__ Ct, const double* __restrict__ y, int64_t m,
                                   const double* __restrict__ es, double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const double e = es[q];
  double acc = 0.0;
  if (e > ext[0]) {
    if (e > ext[n]) {
      acc = y[n - 1];
    } else {
      for (int j = 0; j < n; ++j) acc += y[j] * basis_eval_dev(n, ext, invh, Ct, j, e, 0);
    }
  }
  out[q] = acc;
}
void launch_interp_eval(const isr_problem* p, const double* y, int64_t m, const double* es, double* out) {
  interp_eval_kernel<<<(unsigned)((m + 127) / 128), 128, 0, p->ctx->stream>>>(p->n, p->d_ext.p, p->d_invh.p, p->d_Ct.p,
                                                                               y, m, es, out); ISR_COUNT_LAUNCH();
}

}  // namespace isr
