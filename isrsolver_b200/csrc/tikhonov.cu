// Tikhonov lambda-scan / L-curve and TSVD from ONE decomposition.
//
// Reference: every ISRSolverTikhonov::solve() call rebuilds F, T = G^T W G + lambda F,
// L = F^-1 G^T W G + lambda I and two FullPivLU factorisations (src/ISRSolverTikhonov.cpp:139-155),
// i.e. ~15 N^3 flop per lambda and per toy.  Here the generalised symmetric eigen-system
//     (G^T W G) u_m = mu_m F u_m ,   U^T F U = I
// is computed once (isr_tikhonov_prepare).  With  P = U^T G^T W  and  c = P v :
//     b_lambda   = U (c / (mu + lambda))                                   (solve, :61-72)
//     |W^1/2 (G b - v)|^2 = sum_m c_m^2 lambda^2 / (mu_m (mu_m + lambda)^2)  (evalEqNorm2, :106-109)
//     b^T F b    = sum_m c_m^2 / (mu_m + lambda)^2                          (evalSmoothnessConstraintNorm2)
//     xi'  = 2 b^T F b'  = -2 sum_m c_m^2 / (mu_m + lambda)^3 ,  b' = -L^-1 b   (:115-128)
//     xi'' = 2 b'^T F b' + 2 b^T F b'' = 6 sum_m c_m^2 / (mu_m + lambda)^4 ,  b'' = -2 L^-1 b'
//     Cov_lambda = A+ W^-1 A+^T = U diag(mu / (mu + lambda)^2) U^T
//     chi2(lambda, toy k) = (b_k - b0)^T Cov_lambda^-1 (b_k - b0) = sum_m (e_km - lambda f_m)^2,
//         e_km = (c_km - g_m mu_m) / sqrt(mu_m),  f_m = g_m / sqrt(mu_m),  g = U^T F b0
// so every lambda is a diagonal rescaling: O(N) per lambda and per (lambda, toy) pair.
#include "comm.hpp"
#include "linalg.hpp"
#include "problem.hpp"
#include "solve.hpp"
#include "tikhonov.hpp"
#include "tikhonov_device.cuh"
#include "toy_batch.hpp"
#include "toy_gemm.cuh"

#include <algorithm>
#include <cmath>

namespace isr {

int ensure_svd(isr_problem* p) {
  if (p->svd_ready) return ISR_OK;
  if (!p->has_G) { set_error("SVD: no matrix (call isr_assemble first)"); return ISR_ESTATE; }
  const int n = p->n;
  const size_t nn = (size_t)n * n;
  cudaStream_t st = p->ctx->stream;
  ISR_TRY(p->d_tmp.reserve(nn));
  ISR_TRY(p->d_svU.reserve(nn));
  ISR_TRY(p->d_svVT.reserve(nn));
  ISR_TRY(p->d_sv.reserve(n));
  ISR_CUDA(cudaMemcpyAsync(p->d_tmp.p, p->d_G.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
  ISR_TRY(svd(p->ctx, n, p->d_tmp.p, p->d_svU.p, p->d_sv.p, p->d_svVT.p, p->d_work, p->d_info));
  p->h_sv.resize(n);
  ISR_CUDA(cudaMemcpyAsync(p->h_sv.data(), p->d_sv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  p->svd_ready = true;
  return ISR_OK;
}

// D(i,j) = basis_j'(E_i) (column-major): the point-wise derivative projector,
// src/ISRSolverTikhonov.cpp:92-99.
__global__ void deriv_pairs_kernel(int n, const double* __restrict__ ext, int* __restrict__ js, double* __restrict__ es) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int i = idx % n, j = idx / n;  // column-major position idx = i + n j
  js[idx] = j;
  es[idx] = ext[i + 1];
}

int tikhonov_prepare(isr_problem* p, const double* vcs_err, int deriv_reg) {
  if (!p->has_G) { set_error("isr_tikhonov_prepare: no matrix (call isr_assemble first)"); return ISR_ESTATE; }
  const int n = p->n;
  const size_t nn = (size_t)n * n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  bool same = p->factored && p->vcs_err.size() == (size_t)n;
  if (same) for (int i = 0; i < n; ++i) same = same && p->vcs_err[i] == vcs_err[i];
  if (!same) ISR_TRY(factor(p, vcs_err));
  ISR_TRY(ensure_cov(p));      // G^T W G below
  // w_j = int basis_j dE (src/ISRSolverSLE.cpp:165-171), exact from the polynomial pieces
  std::vector<double> w(n);
  ISR_TRY(isr_basis_integrals(p, w.data()));
  ISR_TRY(p->d_w.reserve(n));
  ISR_CUDA(cudaMemcpyAsync(p->d_w.p, w.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_TRY(p->d_F.reserve(nn));
  ISR_TRY(p->d_U.reserve(nn));
  ISR_TRY(p->d_mu.reserve(n));
  ISR_TRY(p->d_P.reserve(nn));
  ISR_TRY(p->d_FU.reserve(nn));
  ISR_TRY(p->d_tmp.reserve(nn));
  ISR_TRY(p->d_tmp2.reserve(nn));
  if (deriv_reg) {
    // F = D^T diag(w) D  (:141-145)
    DevBuf<int> js;
    DevBuf<double> es;
    ISR_TRY(js.reserve(nn));
    ISR_TRY(es.reserve(nn));
    deriv_pairs_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(n, p->d_ext.p, js.p, es.p); ISR_COUNT_LAUNCH();
    launch_basis_eval(p, (int64_t)nn, js.p, es.p, 1, p->d_tmp.p);           // tmp = D
    scale_rows(st, n, p->d_tmp.p, p->d_w.p, 0, p->d_tmp2.p);                 // tmp2 = diag(w) D
    ISR_TRY(gemm(c, true, false, n, p->d_tmp.p, p->d_tmp2.p, p->d_F.p));
    ISR_CUDA(cudaStreamSynchronize(st));                                      // js/es go out of scope
  } else {
    // F = diag(w)  (:146-148)
    ISR_CUDA(cudaMemsetAsync(p->d_F.p, 0, sizeof(double) * nn, st));
    ISR_CUDA(cudaMemcpy2DAsync(p->d_F.p, sizeof(double) * (n + 1), p->d_w.p, sizeof(double), sizeof(double), n,
                               cudaMemcpyDeviceToDevice, st));
  }
  // generalised eigen-system: U <- eigenvectors of (G^T W G, F), mu ascending
  ISR_CUDA(cudaMemcpyAsync(p->d_U.p, p->d_InvCov.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(p->d_tmp.p, p->d_F.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
  ISR_TRY(sygvd(c, n, p->d_U.p, p->d_tmp.p, p->d_mu.p, p->d_work, p->d_info));
  p->h_mu.resize(n);
  ISR_CUDA(cudaMemcpyAsync(p->h_mu.data(), p->d_mu.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  // P = U^T G^T W = U^T (W G)^T ;  W G = diag(err^-2) G = diag(1/err) R
  scale_rows(st, n, p->d_R.p, p->d_werr.p, 1, p->d_tmp.p);
  ISR_TRY(gemm(c, true, true, n, p->d_U.p, p->d_tmp.p, p->d_P.p));
  // FU = F U  (U^-1 = U^T F = (F U)^T)
  ISR_TRY(gemm(c, false, false, n, p->d_F.p, p->d_U.p, p->d_FU.p));
  ISR_CUDA(cudaStreamSynchronize(st));
  for (int m = 0; m < n; ++m)
    if (!(p->h_mu[m] > 0.0)) {
      set_error("isr_tikhonov_prepare: generalised eigenvalue mu[%d] = %g is not positive", m, p->h_mu[m]);
      return ISR_ENUMERIC;
    }
  p->tik_ready = true;
  p->tik_deriv = deriv_reg;
  return ISR_OK;
}

// One thread per lambda: the L-curve functionals (tikhonov_device.cuh).
__global__ void lcurve_kernel(int n, int64_t L, const double* __restrict__ lam, const double* __restrict__ mu,
                              const double* __restrict__ c, double* __restrict__ eqn, double* __restrict__ smooth,
                              double* __restrict__ curv, double* __restrict__ dcurv, double* __restrict__ Z) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double x, y, cv, dcv;
  lcurve_point(n, lam[l], mu, c, x, y, cv, dcv, Z ? Z + (size_t)l * n : nullptr);
  if (eqn) eqn[l] = x;
  if (smooth) smooth[l] = y;
  if (curv) curv[l] = cv;
  if (dcurv) dcurv[l] = dcv;
}

int tikhonov_scan(isr_problem* p, int64_t L, const double* lambda, const double* vcs, double* eqn, double* smooth,
                  double* curv, double* dcurv, double* bcs_out) {
  if (!p->tik_ready) { set_error("isr_tikhonov_scan: call isr_tikhonov_prepare first"); return ISR_ESTATE; }
  if (L <= 0) return ISR_OK;
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  DevBuf<double> dlam, dv, dc, dout, dZ, dB;
  ISR_TRY(dlam.reserve(L)); ISR_TRY(dv.reserve(n)); ISR_TRY(dc.reserve(n)); ISR_TRY(dout.reserve(4 * L));
  ISR_CUDA(cudaMemcpyAsync(dlam.p, lambda, sizeof(double) * L, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(dv.p, vcs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_TRY(gemv(c, false, n, p->d_P.p, dv.p, dc.p));
  if (bcs_out) { ISR_TRY(dZ.reserve((size_t)L * n)); ISR_TRY(dB.reserve((size_t)L * n)); }
  lcurve_kernel<<<(unsigned)((L + 127) / 128), 128, 0, st>>>(n, L, dlam.p, p->d_mu.p, dc.p, dout.p, dout.p + L,
                                                              dout.p + 2 * L, dout.p + 3 * L, bcs_out ? dZ.p : nullptr); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  if (bcs_out) {
    // B (row-major L x N) = Z U^T  <=>  column-major (N x L): B^T = U Z^T
    const double one = 1.0, zero = 0.0;
    ISR_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, n, (int)L, n, &one, p->d_U.p, n, dZ.p, n, &zero, dB.p, n));
    ISR_CUDA(cudaMemcpyAsync(bcs_out, dB.p, sizeof(double) * L * n, cudaMemcpyDeviceToHost, st));
  }
  if (eqn) ISR_CUDA(cudaMemcpyAsync(eqn, dout.p, sizeof(double) * L, cudaMemcpyDeviceToHost, st));
  if (smooth) ISR_CUDA(cudaMemcpyAsync(smooth, dout.p + L, sizeof(double) * L, cudaMemcpyDeviceToHost, st));
  if (curv) ISR_CUDA(cudaMemcpyAsync(curv, dout.p + 2 * L, sizeof(double) * L, cudaMemcpyDeviceToHost, st));
  if (dcurv) ISR_CUDA(cudaMemcpyAsync(dcurv, dout.p + 3 * L, sizeof(double) * L, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

// out(i, m) = U(i, m) * f(mu_m, lambda)   column-major; mode 0: 1/(mu+la), 1: sqrt(mu)/(mu+la)
__global__ void scale_cols_mu_kernel(int n, const double* __restrict__ U, const double* __restrict__ mu, double la,
                                     int mode, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int m = idx / n;
  const double d = mu[m] + la;
  out[idx] = U[idx] * (mode == 0 ? 1.0 / d : sqrt(mu[m]) / d);
}
// Rop[m][i] (row-major, ld) = ((mu_m + la)/sqrt(mu_m)) * FU(i, m)
__global__ void tik_rop_kernel(int n, int ld, const double* __restrict__ FU, const double* __restrict__ mu, double la,
                               double* __restrict__ R) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ld) return;
  const int m = idx / ld, i = idx % ld;
  R[idx] = i < n ? (mu[m] + la) / sqrt(mu[m]) * FU[(size_t)i + (size_t)n * m] : 0.0;
}

// A+ = U diag(1/(mu+la)) P (column-major into out)
static int tik_operator(isr_problem* p, double la, double* out) {
  const int n = p->n;
  scale_cols_mu_kernel<<<(n * n + 255) / 256, 256, 0, p->ctx->stream>>>(n, p->d_U.p, p->d_mu.p, la, 0, p->d_tmp.p); ISR_COUNT_LAUNCH();
  return gemm(p->ctx, false, false, n, p->d_tmp.p, p->d_P.p, out);
}

int tikhonov_solve(isr_problem* p, double la, const double* vcs, double* bcs_out, double* cov_out) {
  if (!p->tik_ready) { set_error("isr_tikhonov_solve: call isr_tikhonov_prepare first"); return ISR_ESTATE; }
  const int n = p->n;
  const size_t nn = (size_t)n * n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  ISR_TRY(p->d_tmp.reserve(nn));
  ISR_TRY(p->d_tmp2.reserve(nn));
  ISR_TRY(p->d_vin.reserve(n));
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  ISR_CUDA(cudaMemcpyAsync(p->d_vin.p, vcs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_TRY(tik_operator(p, la, p->d_tmp2.p));
  ISR_TRY(gemv(c, false, n, p->d_tmp2.p, p->d_vin.p, p->d_stat.p));
  ISR_CUDA(cudaMemcpyAsync(bcs_out, p->d_stat.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (cov_out) {
    scale_cols_mu_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(n, p->d_U.p, p->d_mu.p, la, 1, p->d_tmp.p); ISR_COUNT_LAUNCH();
    ISR_TRY(gemm(c, false, true, n, p->d_tmp.p, p->d_tmp.p, p->d_tmp2.p));
    ISR_CUDA(cudaMemcpyAsync(cov_out, p->d_tmp2.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
  }
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

int tikhonov_select(isr_problem* p, double la) {
  if (!p->tik_ready) { set_error("isr_tikhonov_select: call isr_tikhonov_prepare first"); return ISR_ESTATE; }
  const int n = p->n, ld = (n + 1) & ~1;
  cudaStream_t st = p->ctx->stream;
  ISR_TRY(p->d_tmp2.reserve((size_t)n * n));
  ISR_TRY(tik_operator(p, la, p->d_tmp2.p));
  colmajor_to_rowmajor(st, n, p->d_tmp2.p, nullptr, 0, p->d_Sop.p, ld);
  tik_rop_kernel<<<(n * ld + 255) / 256, 256, 0, st>>>(n, ld, p->d_FU.p, p->d_mu.p, la, p->d_Rop.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  p->tri_sel[0] = p->tri_sel[1] = 0;
  p->solver_selected = true;
  p->selected_kind = 2;
  return ISR_OK;
}

// Pe[m][i] (row-major, ld) = P(m, i) / sqrt(mu_m) ;  gs[m] = g_m sqrt(mu_m) ; f[m] = g_m / sqrt(mu_m)
__global__ void tik_proj_kernel(int n, int ld, const double* __restrict__ P, const double* __restrict__ mu,
                                const double* __restrict__ g, double* __restrict__ Pe, double* __restrict__ gs,
                                double* __restrict__ f) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n * ld) {
    const int m = idx / ld, i = idx % ld;
    Pe[idx] = i < n ? P[(size_t)m + (size_t)n * i] / sqrt(mu[m]) : 0.0;
  }
  if (idx < n) {
    const double r = sqrt(mu[idx]);
    gs[idx] = g[idx] * r;
    f[idx] = g[idx] / r;
  }
}

// chi2[l][k] = sum_m (e[k][m] - lambda_l f[m])^2.  One CTA: 32 toys (e tile in shared memory, odd
// leading dimension => conflict-free), 8 warps striding over lambda, 4 lambdas per thread per sweep.
constexpr int kScanToys = 32, kScanLam = 4;
__global__ void __launch_bounds__(256) lambda_scan_kernel(int n, int ld, int64_t T, const double* __restrict__ E,
                                                           const double* __restrict__ f, int64_t L,
                                                           const double* __restrict__ lam, double* __restrict__ chi2,
                                                           int64_t stride) {
  extern __shared__ double sh[];
  const int lds = n | 1;
  double* se = sh;                    // [32][lds]
  double* sf = sh + kScanToys * lds;  // [n]
  const int64_t t0 = (int64_t)blockIdx.x * kScanToys;
  for (int q = threadIdx.x; q < kScanToys * n; q += blockDim.x) {
    const int r = q / n, m = q % n;
    se[r * lds + m] = (t0 + r < T) ? E[(t0 + r) * ld + m] : 0.0;
  }
  for (int q = threadIdx.x; q < n; q += blockDim.x) sf[q] = f[q];
  __syncthreads();
  const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double* e = se + t * lds;
  for (int64_t l0 = (int64_t)w * kScanLam; l0 < L; l0 += 8 * kScanLam) {
    double la[kScanLam], acc[kScanLam];
#pragma unroll
    for (int q = 0; q < kScanLam; ++q) { la[q] = (l0 + q < L) ? lam[l0 + q] : 0.0; acc[q] = 0.0; }
    for (int m = 0; m < n; ++m) {
      const double ev = e[m], fv = sf[m];
#pragma unroll
      for (int q = 0; q < kScanLam; ++q) {
        const double d = fma(-la[q], fv, ev);
        acc[q] = fma(d, d, acc[q]);
      }
    }
    if (t0 + t < T) {
#pragma unroll
      for (int q = 0; q < kScanLam; ++q)
        if (l0 + q < L) chi2[(size_t)(l0 + q) * stride + t0 + t] = acc[q];
    }
  }
}

// The (lambda, toy) scan collapses further: e is fixed per toy and f per problem, so
//     chi2(lambda, k) = |e_k - lambda f|^2 = A_k - 2 lambda B_k + lambda^2 C,   A_k = |e_k|^2, B_k = e_k . f, C = |f|^2
// -- two dot products per toy, then O(1) per pair: the scan is bound by WRITING chi2 (8 B per pair).
// No catastrophic cancellation: e_k is unit-variance noise in every eigen-direction (the signal part
// cancels exactly: U^T G^T W G b0 = mu (.) U^T F b0), so e_k is never aligned with the fixed vector f and
// |e_k - lambda f|^2 >= (|e_k| - lambda |f|)^2 stays comparable to the larger of the two terms.
// One warp per toy, lane-strided partial sums + fixed xor tree (deterministic).
__global__ void __launch_bounds__(256) scan_coeff_kernel(int n, int ld, int64_t T, const double* __restrict__ E,
                                                          const double* __restrict__ f, double* __restrict__ AB) {
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= T) return;
  const double* e = E + k * ld;
  double a = 0.0, b = 0.0;
  for (int m = lane; m < n; m += 32) {
    const double ev = e[m];
    a = fma(ev, ev, a);
    b = fma(ev, f[m], b);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, s);
    b += __shfl_xor_sync(0xffffffffu, b, s);
  }
  if (lane == 0) {
    AB[2 * k] = a;
    AB[2 * k + 1] = b;
  }
}
// C = |f|^2 (one warp)
__global__ void scan_fnorm_kernel(int n, const double* __restrict__ f, double* __restrict__ out) {
  double c = 0.0;
  for (int m = threadIdx.x; m < n; m += 32) c = fma(f[m], f[m], c);
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
  if (threadIdx.x == 0) *out = c;
}
// chi2[l][t0 + k] = A_k - lambda_l (2 B_k - lambda_l C): thread = toy (coalesced row writes), 8 lambdas per thread
constexpr int kPairLam = 8;
__global__ void __launch_bounds__(256) scan_pair_kernel(int64_t T, const double* __restrict__ AB, const double* __restrict__ C,
                                                         int64_t L, const double* __restrict__ lam, double* __restrict__ chi2,
                                                         int64_t stride) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t l0 = (int64_t)blockIdx.y * kPairLam;
  if (k >= T) return;
  const double a = AB[2 * k], b2 = 2.0 * AB[2 * k + 1], c = *C;
#pragma unroll
  for (int q = 0; q < kPairLam; ++q) {
    if (l0 + q < L) {
      const double la = lam[l0 + q];
      // streaming store: chi2 is written once and read by nobody on the device
      __stcs(&chi2[(size_t)(l0 + q) * stride + k], fma(-la, fma(-la, c, b2), a));
    }
  }
}

int tikhonov_toy_scan_device(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V_dev,
                             const double* bcs0, double* chi2_dev) {
  if (!p->tik_ready) { set_error("isr_tikhonov_toy_scan: call isr_tikhonov_prepare first"); return ISR_ESTATE; }
  if (L <= 0 || T <= 0) return ISR_OK;
  const int n = p->n, ld = (n + 1) & ~1;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  // persistent scratch (grow-only): repeated scans do not pay cudaMalloc / cudaFree
  DevBuf<double>&dlam = p->d_scan_lam, &db0 = p->d_scan_b0, &dg = p->d_scan_g, &dPe = p->d_scan_Pe, &dAB = p->d_scan_AB, &dC = p->d_scan_C;
  ISR_TRY(dlam.reserve(L)); ISR_TRY(db0.reserve(n)); ISR_TRY(dg.reserve(3 * (size_t)n));
  ISR_CUDA(cudaMemcpyAsync(dlam.p, lambda, sizeof(double) * L, cudaMemcpyHostToDevice, st));
  ISR_CUDA(cudaMemcpyAsync(db0.p, bcs0, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  ISR_TRY(gemv(c, true, n, p->d_FU.p, db0.p, dg.p));               // g = (F U)^T b0 = U^T F b0
  // temporarily route the tile kernel: Sop <- Pe, "b0" <- g sqrt(mu); D then holds e
  ISR_TRY(dPe.reserve((size_t)n * ld));
  tik_proj_kernel<<<(n * ld + 255) / 256, 256, 0, st>>>(n, ld, p->d_P.p, p->d_mu.p, dg.p, dPe.p, dg.p + n, dg.p + 2 * n); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  const size_t smem = sizeof(double) * ((size_t)kScanToys * (n | 1) + n);
  const bool direct = p->scan_variant == 1;    // 1: O(N)-per-pair reference kernel, 0: quadratic form (default)
  if (direct) ISR_CUDA(cudaFuncSetAttribute(lambda_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!direct) {
    ISR_TRY(dC.reserve(1));
    scan_fnorm_kernel<<<1, 32, 0, st>>>(n, dg.p + 2 * n, dC.p); ISR_COUNT_LAUNCH();
  }
  // chunk so that E (T x ld) stays modest; the scan reads E once per chunk
  const int64_t Tc = std::min<int64_t>(T, 1 << 18);
  ISR_TRY(p->d_E2.reserve((size_t)Tc * ld));
  for (int64_t t0 = 0; t0 < T; t0 += Tc) {
    const int64_t tc = std::min<int64_t>(Tc, T - t0);
    ISR_TRY(project_device(p, tc, V_dev + t0 * n, dPe.p, dg.p + n, p->d_E2.p));
    // output rows (one per lambda) have the full stride T; this chunk starts at column t0
    if (direct) {
      lambda_scan_kernel<<<(unsigned)((tc + kScanToys - 1) / kScanToys), 256, smem, st>>>(
          n, ld, tc, p->d_E2.p, dg.p + 2 * n, L, dlam.p, chi2_dev + t0, T); ISR_COUNT_LAUNCH();
    } else {
      ISR_TRY(dAB.reserve((size_t)2 * Tc));
      scan_coeff_kernel<<<(unsigned)((tc + 7) / 8), 256, 0, st>>>(n, ld, tc, p->d_E2.p, dg.p + 2 * n, dAB.p); ISR_COUNT_LAUNCH();
      dim3 grid((unsigned)((tc + 255) / 256), (unsigned)((L + kPairLam - 1) / kPairLam));
      scan_pair_kernel<<<grid, 256, 0, st>>>(tc, dAB.p, dC.p, L, dlam.p, chi2_dev + t0, T); ISR_COUNT_LAUNCH();
    }
  }
  ISR_CUDA(cudaGetLastError());
  return ISR_OK;    // asynchronous on the context stream; lambda / bcs0 were staged from pageable memory (copied at call time)
}

// Per-lambda chi2 test summary without moving the L x T chi2 matrix to the host: one CTA per lambda computes
// min / max / mean over the T toys (fixed-order tree: deterministic) and then fills TH1F(nbins, min, max)
// with ROOT's bin arithmetic (src/Chi2Test.cpp:64-78 -- the maximum lands in the overflow cell).
__global__ void __launch_bounds__(256) scan_hist_kernel(int64_t T, const double* __restrict__ chi2, int nbins,
                                                         double* __restrict__ stats, uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t sh_cnt[];
  __shared__ double s_lo[8], s_hi[8], s_sum[8];
  __shared__ double b_lo, b_hi;
  const double* row = chi2 + (size_t)blockIdx.x * T;
  double lo = INFINITY, hi = -INFINITY, sum = 0.0;
  for (int64_t k = threadIdx.x; k < T; k += blockDim.x) {
    const double v = row[k];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
    sum += v;
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    sum += __shfl_xor_sync(0xffffffffu, sum, s);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; s_sum[w] = sum; }
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x) sh_cnt[q] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = s_lo[0], b = s_hi[0], c = s_sum[0];
    for (int q = 1; q < 8; ++q) { a = fmin(a, s_lo[q]); b = fmax(b, s_hi[q]); c += s_sum[q]; }
    b_lo = a; b_hi = b;
    stats[3 * blockIdx.x + 0] = a;
    stats[3 * blockIdx.x + 1] = b;
    stats[3 * blockIdx.x + 2] = c / (double)T;
  }
  __syncthreads();
  const double a = b_lo, b = b_hi;
  for (int64_t k = threadIdx.x; k < T; k += blockDim.x) {
    const double x = row[k];
    int bin;
    if (x < a) bin = 0;
    else if (!(x < b)) bin = nbins + 1;
    else bin = 1 + (int)__ddiv_rn(__dmul_rn((double)nbins, __dsub_rn(x, a)), __dsub_rn(b, a));
    atomicAdd(&sh_cnt[bin], 1u);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nbins + 2; q += blockDim.x) counts[(size_t)blockIdx.x * (nbins + 2) + q] = sh_cnt[q];
}

int tikhonov_toy_scan_hist(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V, const double* bcs0,
                           int nbins, double* stats_out, uint32_t* counts_out) {
  if (nbins < 1 || nbins > 8190) { set_error("isr_tikhonov_toy_scan_hist: nbins out of range"); return ISR_EINVAL; }
  if (L <= 0 || T <= 0) return ISR_OK;
  cudaStream_t st = p->ctx->stream;
  DevBuf<double>&dV = p->d_scan_V, &dchi = p->d_scan_chi, &dstats = p->d_scan_stats;
  DevBuf<uint32_t>& dcnt = p->d_scan_cnt;
  ISR_TRY(dV.reserve((size_t)T * p->n));
  ISR_TRY(dchi.reserve((size_t)L * T));
  ISR_TRY(dstats.reserve(3 * (size_t)L));
  ISR_TRY(dcnt.reserve((size_t)L * (nbins + 2)));
  ISR_CUDA(cudaMemcpyAsync(dV.p, V, sizeof(double) * T * p->n, cudaMemcpyHostToDevice, st));
  ISR_TRY(tikhonov_toy_scan_device(p, L, lambda, T, dV.p, bcs0, dchi.p));
  scan_hist_kernel<<<(unsigned)L, 256, sizeof(uint32_t) * (nbins + 2), st>>>(T, dchi.p, nbins, dstats.p, dcnt.p); ISR_COUNT_LAUNCH();
  ISR_CUDA(cudaGetLastError());
  if (stats_out) ISR_CUDA(cudaMemcpyAsync(stats_out, dstats.p, sizeof(double) * 3 * L, cudaMemcpyDeviceToHost, st));
  if (counts_out) ISR_CUDA(cudaMemcpyAsync(counts_out, dcnt.p, sizeof(uint32_t) * L * (nbins + 2), cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  return ISR_OK;
}

// ---- the same summary with the TOYS sharded over the ranks of a communicator ----------------------------------------
// Round 1 sharded BASELINE config C4 (1024 lambdas x 10^4 toys) over the lambda grid: the H2D of the toys and the 4 N^2 T
// projection were then replicated on every GPU and 8 GPUs bought 1.05x.  Sharding the toys cuts both 8 ways; the per-lambda
// range needs ONE all-reduce(MIN) of {min_l, -max_l} (2 L doubles) before the cells can be filled, and cells + sums travel
// in ONE all-reduce(SUM) of a packed FP64 buffer [L x (nbins + 2) cells | L sums | toys] (integers are exact in FP64).
__global__ void __launch_bounds__(256) scan_stats_kernel(int64_t T, const double* __restrict__ chi2, double* __restrict__ range,
                                                          double* __restrict__ sums) {
  __shared__ double s_lo[8], s_hi[8], s_sum[8];
  const double* row = chi2 + (size_t)blockIdx.x * T;
  double lo = INFINITY, hi = -INFINITY, sum = 0.0;
  for (int64_t k = threadIdx.x; k < T; k += blockDim.x) {
    const double v = row[k];
    lo = fmin(lo, v); hi = fmax(hi, v); sum += v;
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    sum += __shfl_xor_sync(0xffffffffu, sum, s);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; s_sum[w] = sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = s_lo[0], b = s_hi[0], c = s_sum[0];
    for (int q = 1; q < 8; ++q) { a = fmin(a, s_lo[q]); b = fmax(b, s_hi[q]); c += s_sum[q]; }
    range[2 * blockIdx.x] = a;
    range[2 * blockIdx.x + 1] = -b;
    sums[blockIdx.x] = c;
  }
}
// cells[l][q] = TH1F cells of lambda l against range[l] = {min, -max}; sums[L] = toys of this rank (rides in the sums' all-reduce)
__global__ void __launch_bounds__(256) scan_cells_kernel(int64_t T, const double* __restrict__ chi2, int nbins,
                                                          const double* __restrict__ range, double* __restrict__ sums,
                                                          int64_t L, double toys, uint32_t* __restrict__ cells) {
  extern __shared__ uint32_t sh_cnt[];
  const int nb = nbins + 2;
  const double a = range[2 * blockIdx.x], b = -range[2 * blockIdx.x + 1];
  for (int q = threadIdx.x; q < nb; q += blockDim.x) sh_cnt[q] = 0;
  __syncthreads();
  const double* row = chi2 + (size_t)blockIdx.x * T;
  for (int64_t k = threadIdx.x; k < T; k += blockDim.x) {
    const double x = row[k];
    int bin;
    if (x < a) bin = 0;
    else if (!(x < b)) bin = nbins + 1;
    else bin = 1 + (int)__ddiv_rn(__dmul_rn((double)nbins, __dsub_rn(x, a)), __dsub_rn(b, a));
    atomicAdd(&sh_cnt[bin], 1u);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nb; q += blockDim.x) cells[(size_t)blockIdx.x * nb + q] = sh_cnt[q];
  if (threadIdx.x == 0 && blockIdx.x == 0) sums[L] = toys;
}
__global__ void scan_empty_stats_kernel(int64_t L, double* __restrict__ range, double* __restrict__ sums) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L) { range[2 * l] = INFINITY; range[2 * l + 1] = INFINITY; sums[l] = 0.0; }
}

int tikhonov_chi2_scan(isr_problem* p, const isr_tikhonov_scan_args& a) {
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  const int64_t L = a.n_lambda, T = a.n_toys;
  if (L <= 0 || !a.lambda || T < 0 || !a.bcs0) { set_error("isr_tikhonov_chi2_scan: bad argument"); return ISR_EINVAL; }
  if (a.V && a.V_dev) { set_error("isr_tikhonov_chi2_scan: give V or V_dev, not both"); return ISR_EINVAL; }
  if (T > 0 && !a.V && !a.V_dev) { set_error("isr_tikhonov_chi2_scan: no toys"); return ISR_EINVAL; }
  if (a.nbins < 1 || a.nbins > 8190) { set_error("isr_tikhonov_chi2_scan: nbins out of range"); return ISR_EINVAL; }
  if (a.comm && a.comm->ctx != c) { set_error("isr_tikhonov_chi2_scan: the communicator belongs to another context"); return ISR_EINVAL; }
  const int nb = a.nbins + 2;
  DevBuf<double>&dV = p->d_scan_V, &dchi = p->d_scan_chi, &dstats = p->d_scan_stats;
  DevBuf<uint32_t>& dcnt = p->d_scan_cnt;
  ISR_TRY(dchi.reserve((size_t)L * std::max<int64_t>(T, 1)));
  ISR_TRY(dstats.reserve(3 * (size_t)L + 1));
  ISR_TRY(dcnt.reserve((size_t)L * nb));
  // page-locked staging for the O(L) read-back (a copy into pageable memory would block the host)
  const size_t hlen = 3 * (size_t)L + 1;
  if (hlen > p->pin_scan_n) {
    if (p->pin_scan) cudaFreeHost(p->pin_scan);
    p->pin_scan = nullptr; p->pin_scan_n = 0;
    ISR_CUDA(cudaMallocHost((void**)&p->pin_scan, sizeof(double) * hlen));
    p->pin_scan_n = hlen;
  }
  // the toys go up on the copy stream while the eigen-system is (re)built on the main one
  const double* Vdev = a.V_dev;
  if (a.V && T > 0) {
    ISR_TRY(dV.reserve((size_t)T * n));
    if (!p->copy_stream) {
      ISR_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
      for (int q = 0; q < 2; ++q) {
        ISR_CUDA(cudaEventCreateWithFlags(&p->ev_copied[q], cudaEventDisableTiming));
        ISR_CUDA(cudaEventCreateWithFlags(&p->ev_done[q], cudaEventDisableTiming));
      }
    }
    ISR_CUDA(cudaEventRecord(c->ev_fork, st));
    ISR_CUDA(cudaStreamWaitEvent(p->copy_stream, c->ev_fork, 0));
    ISR_CUDA(cudaMemcpyAsync(dV.p, a.V, sizeof(double) * T * n, cudaMemcpyHostToDevice, p->copy_stream));
    ISR_CUDA(cudaEventRecord(p->ev_copied[0], p->copy_stream));
    Vdev = dV.p;
  }
  int rc = ISR_OK;
  auto drain = [&](int code) { if (p->copy_stream) cudaStreamSynchronize(p->copy_stream); cudaStreamSynchronize(st); return code; };
  if ((a.flags & ISR_TEST_ASSEMBLE) && (rc = assemble(p, a.ecm_err)) != ISR_OK) return drain(rc);
  if (a.flags & ISR_TEST_FACTOR) {
    if (!a.vcs_err) { set_error("isr_tikhonov_chi2_scan: ISR_TEST_FACTOR needs vcs_err"); return drain(ISR_EINVAL); }
    if ((rc = tikhonov_prepare(p, a.vcs_err, a.deriv_reg)) != ISR_OK) return drain(rc);
  }
  if (!p->tik_ready) { set_error("isr_tikhonov_chi2_scan: no eigen-system (set ISR_TEST_FACTOR or call isr_tikhonov_prepare first)"); return drain(ISR_ESTATE); }
  if (a.V && T > 0) ISR_CUDA(cudaStreamWaitEvent(st, p->ev_copied[0], 0));
  double* d_range = dstats.p;          // [L][2]
  double* d_sums = dstats.p + 2 * L;   // [L] + toy count
  if (T > 0) {
    if ((rc = tikhonov_toy_scan_device(p, L, a.lambda, T, Vdev, a.bcs0, dchi.p)) != ISR_OK) return drain(rc);
    scan_stats_kernel<<<(unsigned)L, 256, 0, st>>>(T, dchi.p, d_range, d_sums); ISR_COUNT_LAUNCH();
  } else {
    scan_empty_stats_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(L, d_range, d_sums); ISR_COUNT_LAUNCH();
  }
  if (a.comm && (rc = comm_allreduce(a.comm, d_range, 2 * L, 0, 1)) != ISR_OK) return drain(rc);
  scan_cells_kernel<<<(unsigned)L, 256, sizeof(uint32_t) * nb, st>>>(T, dchi.p, a.nbins, d_range, d_sums, L, (double)T, dcnt.p); ISR_COUNT_LAUNCH();
  if (a.comm) {
    if ((rc = comm_allreduce(a.comm, dcnt.p, (int64_t)L * nb, 2, 0)) != ISR_OK) return drain(rc);
    if ((rc = comm_allreduce(a.comm, d_sums, L + 1, 0, 0)) != ISR_OK) return drain(rc);
  }
  bool bad = cudaMemcpyAsync(p->pin_scan, dstats.p, sizeof(double) * hlen, cudaMemcpyDeviceToHost, st) != cudaSuccess;
  if (a.counts_out) bad = bad || cudaMemcpyAsync(a.counts_out, dcnt.p, sizeof(uint32_t) * (size_t)L * nb, cudaMemcpyDeviceToHost, st) != cudaSuccess;
  if (bad || cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("isr_tikhonov_chi2_scan: CUDA error '%s'", cudaGetErrorString(cudaGetLastError()));
    return drain(ISR_ECUDA);
  }
  const double* h = p->pin_scan;
  const double total = h[3 * L];
  if (a.total_toys_out) *a.total_toys_out = (int64_t)total;
  if (a.stats_out)
    for (int64_t l = 0; l < L; ++l) {
      a.stats_out[3 * l + 0] = h[2 * l];
      a.stats_out[3 * l + 1] = -h[2 * l + 1];
      a.stats_out[3 * l + 2] = total > 0 ? h[2 * L + l] / total : 0.0;
    }
  return ISR_OK;
}

// ---- TSVD (src/ISRSolverTSVD.cpp:53-74) -----------------------------------------------------------
// z_m = (U^T v)_m / sigma_m for m in [first, first + count), 0 otherwise
__global__ void tsvd_filter_kernel(int n, int first, int count, const double* __restrict__ sv, double* __restrict__ z) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  z[m] = (m >= first && m < first + count) ? z[m] / sv[m] : 0.0;
}
// out(i, m) = U(i, m) * (m in window ? scale_m : 0), column-major; mode 0: 1/sigma, mode 1: sigma
__global__ void tsvd_window_kernel(int n, int first, int count, const double* __restrict__ U, const double* __restrict__ sv,
                                   int mode, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int m = idx / n;
  const bool in = m >= first && m < first + count;
  out[idx] = in ? U[idx] * (mode == 0 ? 1.0 / sv[m] : sv[m]) : 0.0;
}

static int tsvd_window(isr_problem* p, int k, int keep_one, int* first, int* count) {
  if (k < 1 || k > p->n) { set_error("TSVD: upper index %d out of range [1, %d]", k, p->n); return ISR_EINVAL; }
  *first = keep_one ? k - 1 : 0;   // :63-68
  *count = keep_one ? 1 : k;
  return ISR_OK;
}

int tsvd_solve(isr_problem* p, int k, int keep_one, const double* vcs, const double* vcs_err, double* bcs_out,
               double* cov_out) {
  int first, count;
  ISR_TRY(tsvd_window(p, k, keep_one, &first, &count));
  ISR_TRY(ensure_svd(p));
  const int n = p->n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  ISR_TRY(p->d_vin.reserve(n));
  ISR_TRY(p->d_stat.reserve((size_t)4 * n));
  ISR_CUDA(cudaMemcpyAsync(p->d_vin.p, vcs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  // b = V_s Sigma_s^-1 U_s^T v  (minimum-norm least squares of K b = v, what COD(K).solve returns)
  ISR_TRY(gemv(c, true, n, p->d_svU.p, p->d_vin.p, p->d_stat.p));
  tsvd_filter_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, first, count, p->d_sv.p, p->d_stat.p); ISR_COUNT_LAUNCH();
  ISR_TRY(gemv(c, true, n, p->d_svVT.p, p->d_stat.p, p->d_stat.p + n));
  ISR_CUDA(cudaMemcpyAsync(bcs_out, p->d_stat.p + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  if (cov_out) {
    if (keep_one || k != n) {
      set_error("isr_tsvd_solve: the covariance (K^T W K)^-1 is singular for a truncated K; request it only for k == N");
      return ISR_ENUMERIC;
    }
    if (!vcs_err) { set_error("isr_tsvd_solve: vcs_err needed for the covariance"); return ISR_EINVAL; }
    // (K^T W K)^-1 for the full-rank K = G is the SLE covariance: it needs the SLE factorisation, whose operator set-up
    // overwrites d_Sop / d_Rop.  A Tikhonov / TSVD selection made earlier is therefore DROPPED, explicitly: the next toy
    // batch fails with ISR_ESTATE until the caller selects again (it used to run silently on the SLE operators).
    const int kind_before = p->selected_kind;
    bool same = p->factored && p->selected_kind == 1 && p->vcs_err.size() == (size_t)n;
    if (same) for (int i = 0; i < n; ++i) same = same && p->vcs_err[i] == vcs_err[i];
    if (!same) ISR_TRY(factor(p, vcs_err));
    if (kind_before == 2 || kind_before == 3) { p->solver_selected = false; p->selected_kind = 0; }
    ISR_TRY(ensure_cov(p));
    ISR_CUDA(cudaMemcpyAsync(cov_out, p->d_Cov.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
    ISR_CUDA(cudaStreamSynchronize(st));
  }
  return ISR_OK;
}

int tsvd_select(isr_problem* p, int k, int keep_one, const double* vcs_err) {
  int first, count;
  ISR_TRY(tsvd_window(p, k, keep_one, &first, &count));
  ISR_TRY(ensure_svd(p));
  const int n = p->n, ld = (n + 1) & ~1;
  const size_t nn = (size_t)n * n;
  isr_ctx* c = p->ctx;
  cudaStream_t st = c->stream;
  ISR_TRY(p->d_tmp.reserve(nn));
  ISR_TRY(p->d_tmp2.reserve(nn));
  ISR_TRY(p->d_werr.reserve(n));
  ISR_TRY(p->d_Sop.reserve((size_t)n * ld));
  ISR_TRY(p->d_Rop.reserve((size_t)n * ld));
  ISR_CUDA(cudaMemcpyAsync(p->d_werr.p, vcs_err, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  // Sop = K^+ = V_s Sigma_s^-1 U_s^T = VT^T (U_window / sigma)^T
  tsvd_window_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(n, first, count, p->d_svU.p, p->d_sv.p, 0, p->d_tmp.p); ISR_COUNT_LAUNCH();
  ISR_TRY(gemm(c, true, true, n, p->d_svVT.p, p->d_tmp.p, p->d_tmp2.p));
  colmajor_to_rowmajor(st, n, p->d_tmp2.p, nullptr, 0, p->d_Sop.p, ld);
  // Rop = W^(1/2) K = diag(1/err) U_window Sigma VT   (chi2 = d^T K^T W K d)
  tsvd_window_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(n, first, count, p->d_svU.p, p->d_sv.p, 1, p->d_tmp.p); ISR_COUNT_LAUNCH();
  ISR_TRY(gemm(c, false, false, n, p->d_tmp.p, p->d_svVT.p, p->d_tmp2.p));
  colmajor_to_rowmajor(st, n, p->d_tmp2.p, p->d_werr.p, 1, p->d_Rop.p, ld);
  ISR_CUDA(cudaGetLastError());
  p->tri_sel[0] = p->tri_sel[1] = 0;
  p->factored = false;   // Sop/Rop no longer belong to the SLE factorisation
  p->solver_selected = true;
  p->selected_kind = 3;
  return ISR_OK;
}

}  // namespace isr
