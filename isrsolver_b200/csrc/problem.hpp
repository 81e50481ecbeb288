// isr_problem: one solver instance (host mirror + device residency).
#pragma once

#include "common.cuh"

struct isr_problem {
  isr_ctx* ctx = nullptr;
  int n = 0;
  double threshold = 0;
  std::vector<double> ext;   // n+1: threshold, E_0..E_{n-1}        (src/Interpolator.cpp:37-39)
  std::vector<int> ranges;   // sorted by first segment, R x {cubic, lo, hi}
  std::vector<int> seg_range;  // per segment k: index of the range that contains it

  // ---- basis tables --------------------------------------------------------------------------
  // Ct[(k*4 + p)*n + j]: coefficient of (delta/h_k)^p of basis j on segment k = (ext[k], ext[k+1]].
  std::vector<double> h_Ct;
  std::vector<int> h_klo, h_khi;  // per column j: first/last segment with non-zero coefficients
  isr::DevBuf<double> d_ext, d_invh, d_Ct;
  isr::DevBuf<int> d_klo, d_khi;

  // ---- efficiency ----------------------------------------------------------------------------
  isr::EpsDev eps{};
  isr::DevBuf<double> d_eps_vals, d_eps_edges;

  // ---- assembly workspaces -------------------------------------------------------------------
  isr::DevBuf<isr::RowConst> d_row;
  isr::DevBuf<int> d_cnt;        // panels per (row, segment) item, then exclusive scan (+1 entry)
  isr::DevBuf<int> d_off;
  isr::DevBuf<int> d_local, d_blocksum;   // block-local prefixes and block totals of the panel counts
  isr::DevBuf<double> d_pa, d_pb, d_peps, d_pm;  // panel bounds, eps value, 4 moments per panel
  isr::DevBuf<int> d_pitem, d_pkind;
  isr::DevBuf<double> d_M;       // [i][k][4] normalised moments (filled on demand by isr_debug_moments)
  isr::DevBuf<double> d_G;       // column-major N x N
  isr::DevBuf<double> d_Gpart;   // split-K partials of the contraction
  isr::DevBuf<double> d_S, d_tmp, d_sigE;
  int64_t max_panels = 0;
  bool panels_ready = false;          // the panel set of this problem has been enumerated (it depends on energies and efficiency edges only)
  int sampled_panels = -1;            // panel count of the last isr_assemble_nodes (host-sampled efficiency)
  isr::DevBuf<double> d_node_x, d_node_eps;
  isr::DevBuf<int> d_node_row;
  int64_t last_panels = 0;
  bool has_G = false;
  bool G_injected = false;   // matrix came from isr_set_matrix (cannot be re-assembled)
  bool G_lower = false;      // known on the host to be lower triangular (all ranges linear and no spread, or an injected
                             // matrix that was checked): lets isr_factor pick the triangular solve for N > 240
  isr::DevBuf<double> d_G0;   // condition-number scan: the unspread matrix

  // ---- factorisation / selected solver -------------------------------------------------------
  // All N x N, ROW-major on the device: Sop (solve operator: b = Sop v) and Rop (chi2 factor:
  // chi2 = |Rop (b - b0)|^2).
  isr::DevBuf<double> d_Ginv, d_R, d_Cov, d_InvCov, d_Sop, d_Rop, d_werr;
  isr::DevBuf<double> d_work;
  isr::DevBuf<double> d_cert;        // the certificate's three scalars
  isr::DevBuf<double> d_cert_prod;   // G X of the inverse's certificate (side stream: nobody else may touch it)
  isr::DevBuf<int> d_ipiv, d_info;
  int tri_sel[2] = {0, 0};     // {S, R}: d_Sop / d_Rop of the current selection is exactly lower triangular (the toy kernel
                               // skips the zeros); known on the host without a synchronisation of its own (lu_inverse)
  int tri_sle[2] = {0, 0};     // the same for the SLE factorisation {Ginv, R}, restored by select_sle
  std::vector<double> vcs_err, h_b0, h_gen;   // h_gen: {vcs, vcs_err} of the device toy generator (stable host mirror)
  isr::DevBuf<double> d_gen;
  bool factored = false;
  bool cov_ready = false;      // d_Cov / d_InvCov match the current factorisation
  bool solver_selected = false;
  int selected_kind = 0;       // whose operators d_Sop / d_Rop hold: 0 none, 1 SLE (isr_factor), 2 Tikhonov, 3 TSVD
  bool cert_pending = false;   // factor_enqueue ran; factor_finish has not read the inverse's certificate yet
  double* cert_dev = nullptr;  // device copy of the certificate scalars (= d_cert.p once a factorisation was enqueued)
  double* pin_cert = nullptr;  // page-locked: certificate scalars + cuSOLVER info of the pending factorisation (8 doubles)
  int toy_variant = 0;
  int64_t last_chi2_count = 0;   // chi2 values the last host-API batch / campaign left in d_chi2h

  // ---- toy batch workspaces ------------------------------------------------------------------
  isr::DevBuf<double> d_V, d_chi2, d_acc, d_b0, d_B;
  isr::DevBuf<double> d_vin;     // right-hand side of the single solves (isr_solve / isr_tikhonov_solve / isr_tsvd_solve): NOT d_b0, whose
                                 // content isr_toy_batch_dev caches against h_b0
  isr::DevBuf<uint32_t> d_rhist;
  // isr_chi2_test: [range (2) | packed result] on the device, TH1F cells, and the page-locked copy of the former
  isr::DevBuf<double> d_pack;
  isr::DevBuf<int> d_cells;
  isr::DevBuf<double> d_part;        // per-block {min, max} partials of the range kernel
  isr::DevBuf<unsigned> d_arrive;    // arrival counters of the two fused tail kernels
  double* pin_test = nullptr;
  double* pin_test_dev = nullptr;     // the same buffer as the device sees it (zero-copy result store of the last kernel)
  cudaEvent_t ev_time[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // start | assembled | factor enqueued | toys enqueued | results copied
  size_t pin_test_n = 0;
  // host-buffer pipeline (isr_toy_batch): two V slots, chi2 / statistics staging, copy stream
  isr::DevBuf<double> d_Vh, d_chi2h, d_stat, d_Bh;
  cudaStream_t copy_stream = nullptr;
  int64_t pipe_slot = 0;       // toys per slot of the pipeline in flight (host_pipe_begin)
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  ~isr_problem() {
    for (int q = 0; q < 2; ++q) {
      if (ev_copied[q]) cudaEventDestroy(ev_copied[q]);
      if (ev_done[q]) cudaEventDestroy(ev_done[q]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (int q = 0; q < 5; ++q) if (ev_time[q]) cudaEventDestroy(ev_time[q]);
    if (pin_cert) cudaFreeHost(pin_cert);
    if (pin_scan) cudaFreeHost(pin_scan);
    if (pin_test) cudaFreeHost(pin_test);
  }

  // ---- Tikhonov ------------------------------------------------------------------------------
  bool tik_ready = false;
  int tik_deriv = 1;
  int scan_variant = 0;   // (lambda, toy) scan: 0 = quadratic form in lambda, 1 = direct O(N)-per-pair kernel
  isr::DevBuf<double> d_F, d_U, d_mu, d_P, d_FU, d_w, d_tmp2, d_E, d_E2;
  std::vector<double> h_mu;
  // (lambda, toy) scan scratch, grow-only
  isr::DevBuf<double> d_scan_lam, d_scan_b0, d_scan_g, d_scan_Pe, d_scan_AB, d_scan_C, d_scan_V, d_scan_chi, d_scan_stats;
  isr::DevBuf<uint32_t> d_scan_cnt;
  double* pin_scan = nullptr;           // isr_tikhonov_chi2_scan: page-locked [L ranges (2) | L sums | toys]
  size_t pin_scan_n = 0;

  // ---- SVD -----------------------------------------------------------------------------------
  bool svd_ready = false;
  isr::DevBuf<double> d_svU, d_svVT, d_sv;  // column-major U, V^T; singular values descending
  std::vector<double> h_sv;
};

namespace isr {
int build_basis_tables(isr_problem* p);
int set_ranges(isr_problem* p, int nranges, const int* ranges);
void launch_kf_plain(int64_t n, const double* x, const double* s, double* out, cudaStream_t st);
void launch_basis_eval(const isr_problem* p, int64_t m, const int* js, const double* es, int deriv, double* out);
void launch_interp_eval(const isr_problem* p, const double* y, int64_t m, const double* es, double* out);
int ensure_svd(isr_problem* p);
int upload_eps(isr_problem* p, const isr_eps_desc* eps);
int build_eps_rows(cudaStream_t st, int n, const double* energy, const isr_eps_desc* e, DevBuf<double>& d_vals,
                   DevBuf<double>& d_edges, EpsDev& out);
int exclusive_scan_int(cudaStream_t st, int n, const int* cnt_dev, int* off_dev);   // off[n] = total
int assemble(isr_problem* p, const double* ecm_err);
int reduce_moments(isr_problem* p);
int assemble_nodes(isr_problem* p, int64_t cap, double* x_out, int* row_out, int64_t* count);
int assemble_sampled(isr_problem* p, int64_t nn, const double* eps_values, const double* ecm_err);
int spread_batch(isr_problem* p, cudaStream_t st, int K, const double* sig_dev, int per_point, double* Sbatch);
int apply_spread(isr_problem* p, cudaStream_t st, const double* sigE_dev, const double* G0, double* S, double* Gout);
}  // namespace isr
