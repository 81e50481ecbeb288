// Device helpers shared by the matrix assembly (assemble.cu) and the forward-convolution service (conv.cu):
// grading points of a row's x-interval, ROOT bin arithmetic for the efficiency tables, the panels of a (row, segment)
// item and the quadrature nodes of a panel in its singularity-removing variable.
#pragma once

#include "kf_device.cuh"

namespace isr {

// Round-to-nearest products / quotients that must not be contracted into FMAs (bit-exact ROOT bin arithmetic): the
// intrinsics on the device, plain operators in the host compilation of the same source (tests/cpp/test_kf_host.cu).
#ifdef __CUDA_ARCH__
#define ISR_DMUL_RN(a, b) __dmul_rn((a), (b))
#define ISR_DDIV_RN(a, b) __ddiv_rn((a), (b))
#define ISR_DSUB_RN(a, b) __dsub_rn((a), (b))
#else
#define ISR_DMUL_RN(a, b) ((a) * (b))
#define ISR_DDIV_RN(a, b) ((a) / (b))
#define ISR_DSUB_RN(a, b) ((a) - (b))
#endif

// TH1::Fill bin arithmetic (TAxis::FindBin, fixed bins): 0 = underflow, nbins+1 = overflow.  __host__ __device__ (the
// round-to-nearest macros of panel_device.cuh): tests/cpp/test_kf_host.cu checks it against the oracle on the CPU.
__host__ __device__ __forceinline__ int th1_bin(int nbins, double lo, double hi, double x) {
  if (x < lo) return 0;
  if (!(x < hi)) return nbins + 1;
  return 1 + (int)ISR_DDIV_RN(ISR_DMUL_RN((double)nbins, ISR_DSUB_RN(x, lo)), ISR_DSUB_RN(hi, lo));
}

// Grading points of the x-interval (0, rc.xmax] of a row whose constants are already set: x0, then
// x0 (1 + 4^m) for m = -2, -1, 0, 1, ... (ratio <= 4 to both singular points 0 and x0), then
// 1 - (1 - xmax) 4^m towards the pole of the pair term at x = 1.  Every panel between consecutive points is
// then analytic with Bernstein parameter rho >= 3 in its integration variable.
__host__ __device__ __forceinline__ void row_grading(RowConst& rc) {
  int nb = 0;
  rc.rb[nb++] = rc.x0;
  double f = 1.0 / 16.0;
  while (nb < kMaxRowBreaks - 6) {
    const double v = rc.x0 * (1.0 + f);
    if (!(v < rc.xmax)) break;
    rc.rb[nb++] = v;
    f *= 4.0;
  }
  const double d = 1.0 - rc.xmax;
  double g = 4.0;
  double tmp[6];
  int nt = 0;
  while (nt < 6) {
    const double v = 1.0 - d * g;
    if (!(v > 0.25) || !(v > rc.rb[nb - 1])) break;
    tmp[nt++] = v;
    g *= 4.0;
  }
  // tmp is descending; merge so rb stays ascending
  for (int q = nt - 1; q >= 0; --q) rc.rb[nb++] = tmp[q];
  // insertion sort (the two families can interleave)
  for (int a = 1; a < nb; ++a) {
    const double v = rc.rb[a];
    int b = a - 1;
    while (b >= 0 && rc.rb[b] > v) { rc.rb[b + 1] = rc.rb[b]; --b; }
    rc.rb[b + 1] = v;
  }
  rc.nrb = nb;
}

// ROOT TAxis::FindBin arithmetic (bit-exact: no FMA contraction in the index expression).
__host__ __device__ __forceinline__ int eps_find_bin(const EpsDev& e, double x) {
  if (x < e.xlo) return 0;
  if (!(x < e.xhi)) return e.nx + 1;
  if (!e.xedges) {
    const double t = ISR_DDIV_RN(ISR_DMUL_RN((double)e.nx, ISR_DSUB_RN(x, e.xlo)), ISR_DSUB_RN(e.xhi, e.xlo));
    return 1 + (int)t;
  }
  int l = 0, h = e.nx;
  while (h - l > 1) {
    const int m = (l + h) >> 1;
    if (e.xedges[m] <= x) l = m; else h = m;
  }
  return 1 + l;
}
__host__ __device__ __forceinline__ double eps_edge(const EpsDev& e, int m) {
  if (e.xedges) return e.xedges[m];
  if (m == e.nx) return e.xhi;
  return e.xlo + (double)m * ((e.xhi - e.xlo) / (double)e.nx);
}
// smallest edge index m in [0, nx] with edge(m) > v, or nx+1
__host__ __device__ __forceinline__ int eps_first_edge_above(const EpsDev& e, double v) {
  int m;
  if (e.xedges) {
    int l = -1, h = e.nx + 1;  // edge(l) <= v < edge(h)
    while (h - l > 1) {
      const int q = (l + h) >> 1;
      if (e.xedges[q] <= v) l = q; else h = q;
    }
    return h;
  }
  const double w = (e.xhi - e.xlo) / (double)e.nx;
  const double t = floor((v - e.xlo) / w) + 1.0;
  m = t < 0.0 ? 0 : (t > (double)(e.nx + 1) ? e.nx + 1 : (int)t);
  while (m > 0 && eps_edge(e, m - 1) > v) --m;
  while (m <= e.nx && !(eps_edge(e, m) > v)) ++m;
  return m;
}

// smallest edge index m in [0, nx] with edge(m) >= v, or nx+1
__host__ __device__ __forceinline__ int eps_first_edge_ge(const EpsDev& e, double v) {
  int m = eps_first_edge_above(e, v);          // edge(m) > v, edge(m-1) <= v
  if (m > 0 && eps_edge(e, m - 1) == v) --m;   // an edge exactly at v
  return m;
}

// ---- panel enumeration and quadrature nodes (used by assemble.cu; __host__ __device__ so that tests/cpp/test_panels_host.cu
// can integrate matrix moments with this very source on the CPU and compare them with the oracle) ----------------------

// x-coordinate of energy node m seen from row i: x = 1 - (ext[m]/E_i)^2, exactly 0 at m = i+1.
__host__ __device__ __forceinline__ double knot_x(const double* __restrict__ ext, double E, int i, int m) {
  if (m == i + 1) return 0.0;
  const double r = ext[m] / E;
  return fmax(0.0, 1.0 - r * r);
}

// Panels of item (row i, segment k): the x-interval (xa, xb) of the segment seen from row i, cut at the row's
// grading points R = {rb[r] in (xa, xb)} and at the efficiency bin edges E = {edge(m) in (xa, xb)}.  The interior
// cuts are the stable merge of the two ascending sets (R first on a tie; a tie leaves a zero-width panel, which
// integrates to zero), so the panel count is analytic -- 1 + |R| + |E| -- and the q-th cut can be selected directly:
// no serial walk over hundreds of bins for the items that span most of a tabulated efficiency.
struct ItemCuts {
  double xa, xb;
  int r0, nR;     // rb[r0 .. r0 + nR) lie strictly inside (xa, xb)
  int e0, nE;     // edges e0 .. e0 + nE - 1 lie strictly inside (xa, xb)
  __host__ __device__ int count() const { return xb > xa ? 1 + nR + nE : 0; }
};
__host__ __device__ __forceinline__ ItemCuts item_cuts(const RowConst& rc, const EpsDev& eps, const double* __restrict__ ext,
                                              int i, int k) {
  ItemCuts c;
  c.xa = knot_x(ext, rc.E, i, k + 1);
  c.xb = knot_x(ext, rc.E, i, k);
  c.r0 = c.nR = c.e0 = c.nE = 0;
  if (!(c.xb > c.xa)) return c;
  int r = 0;
  while (r < rc.nrb && !(rc.rb[r] > c.xa)) ++r;
  c.r0 = r;
  while (r < rc.nrb && rc.rb[r] < c.xb) ++r;
  c.nR = r - c.r0;
  if (eps.kind == 2) {
    c.e0 = eps_first_edge_above(eps, c.xa);
    const int e1 = min(eps_first_edge_ge(eps, c.xb), eps.nx + 1);
    c.nE = max(0, e1 - c.e0);
  }
  return c;
}
// Merge state after the first t interior cuts have been consumed: ir grading points and ie = t - ir edges.
__host__ __device__ __forceinline__ void cuts_state_at(const ItemCuts& c, const RowConst& rc, const EpsDev& eps, int t, int& ir, int& ie) {
  int used = 0;                 // grading points whose merged position is below t
  for (int r = 0; r < c.nR; ++r) {
    int before = 0;             // edges of E strictly below rb[r] (they precede it; equal ones follow)
    if (c.nE > 0) before = min(max(eps_first_edge_ge(eps, rc.rb[c.r0 + r]) - c.e0, 0), c.nE);
    if (r + before < t) ++used; else break;
  }
  ir = used;
  ie = t - used;
}
// Emit panels [q0, q1) of the item in ascending x: incremental merge from the state at cut q0 - 1.
template <typename Emit>
__host__ __device__ __forceinline__ void walk_panels(const ItemCuts& c, const RowConst& rc, const EpsDev& eps, int q0, int q1, Emit emit) {
  const int count = c.count();
  if (q0 >= q1 || q0 >= count) return;
  int ir = 0, ie = 0;
  double cur = c.xa;
  if (q0 > 0) {
    cuts_state_at(c, rc, eps, q0 - 1, ir, ie);       // state before cut number q0 - 1 ...
    const double nr = ir < c.nR ? rc.rb[c.r0 + ir] : 2.0, ne = ie < c.nE ? eps_edge(eps, c.e0 + ie) : 2.0;
    if (nr <= ne) { cur = nr; ++ir; } else { cur = ne; ++ie; }   // ... and past it: cur = left edge of panel q0
  }
  for (int q = q0; q < q1; ++q) {
    double nxt = c.xb;
    if (q < count - 1) {
      const double nr = ir < c.nR ? rc.rb[c.r0 + ir] : 2.0, ne = ie < c.nE ? eps_edge(eps, c.e0 + ie) : 2.0;
      if (nr <= ne) { nxt = nr; ++ir; } else { nxt = ne; ++ie; }
    }
    emit(q, cur, nxt);
    cur = nxt;
  }
}

__host__ __device__ __forceinline__ void item_decode(int idx, int& i, int& k) {
  int r = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
  while ((long long)(r + 1) * (r + 2) / 2 <= idx) ++r;
  while ((long long)r * (r + 1) / 2 > idx) --r;
  i = r;
  k = idx - r * (r + 1) / 2;
}

// kind: 0 = regular (x), 1 = t = x^beta, 2 = u = (x - x0)^beta
__host__ __device__ __forceinline__ int panel_kind(const RowConst& rc, double b) {
  if (!(b > rc.x0)) return 1;
  if (!(b > rc.xin)) return 2;
  return 0;
}
// Node `lane` of a panel: x, ln x, the pair-threshold power (x - x0)^beta (-1 below x0) and the quadrature
// weight including dx/d(variable).  kind 0: x itself, 1: t = x^beta, 2: u = (x - x0)^beta.
struct PanelNode { double x, lnx, pw, wj; };
__host__ __device__ __forceinline__ PanelNode panel_node(const RowConst& rc, double a, double b, int kind, double t, double w) {
  PanelNode q;
  q.pw = -1.0;
  if (kind == 0) {
    const double half = 0.5 * (b - a);
    q.x = fma(half, t, 0.5 * (a + b));
    q.wj = half * w;
    q.lnx = log(q.x);
    q.pw = exp(rc.beta * log(q.x - rc.x0));
  } else if (kind == 1) {
    const double ta = a > 0.0 ? exp(rc.beta * log(a)) : 0.0;
    const double tb = exp(rc.beta * log(b));
    const double half = 0.5 * (tb - ta);
    const double tt = fma(half, t, 0.5 * (ta + tb));
    q.lnx = log(tt) * rc.inv_beta;
    q.x = exp(q.lnx);
    q.wj = half * w * rc.inv_beta * q.x / tt;
  } else {
    const double ya = a - rc.x0, yb = b - rc.x0;
    const double ua = ya > 0.0 ? exp(rc.beta * log(ya)) : 0.0;
    const double ub = exp(rc.beta * log(yb));
    const double half = 0.5 * (ub - ua);
    const double u = fma(half, t, 0.5 * (ua + ub));
    const double y = exp(log(u) * rc.inv_beta);
    q.x = rc.x0 + y;
    q.lnx = log(q.x);
    q.wj = half * w * rc.inv_beta * y / u;
    q.pw = u;
  }
  return q;
}

// ---- basis functions from the coefficient tables (energy spread, isr_basis_eval, isr_interp_eval) ---------------------
// Segment lookup: k with ext[k] < E <= ext[k+1] (the (min, max] convention of
// BaseRangeInterpolator::isEnergyInRange, src/BaseRangeInterpolator.cpp:45-47).
__host__ __device__ __forceinline__ int segment_of(const double* __restrict__ ext, int n, double e) {
  int l = 0, h = n;  // ext[l] < e <= ext[h]
  while (h - l > 1) {
    const int m = (l + h) >> 1;
    if (ext[m] < e) l = m; else h = m;
  }
  return l;
}
// Interpolator::basisEval (src/Interpolator.cpp:217-258) / basisDerivEval (:266-287) from the tables.
__host__ __device__ inline double basis_eval_dev(int n, const double* __restrict__ ext, const double* __restrict__ invh,
                                 const double* __restrict__ Ct, int j, double e, int deriv) {
  if (e <= ext[0]) return 0.0;
  if (e > ext[n]) {
    if (deriv) return 0.0;
    return j == n - 1 ? 1.0 : 0.0;  // value at E_max of the last range's basis: delta_{j, n-1}
  }
  const int k = segment_of(ext, n, e);
  const double dn = (e - ext[k]) * invh[k];
  const double c0 = Ct[((size_t)k * 4 + 0) * n + j], c1 = Ct[((size_t)k * 4 + 1) * n + j];
  const double c2 = Ct[((size_t)k * 4 + 2) * n + j], c3 = Ct[((size_t)k * 4 + 3) * n + j];
  if (deriv) return (c1 + dn * (2.0 * c2 + 3.0 * c3 * dn)) * invh[k];
  return c0 + dn * (c1 + dn * (c2 + dn * c3));
}

}  // namespace isr
