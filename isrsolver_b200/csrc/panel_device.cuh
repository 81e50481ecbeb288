// Device helpers shared by the matrix assembly (assemble.cu) and the forward-convolution service (conv.cu):
// grading points of a row's x-interval, ROOT bin arithmetic for the efficiency tables.
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

}  // namespace isr
