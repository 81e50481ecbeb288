/*
 * isr_b200.h -- C ABI of libisr_b200.so: the B200-native (sm_100a, FP64) implementation of
 * ISRSolver's data-parallel hot path.
 *
 * The reference (sergeigribanov/ISRSolver) has no C ABI; its boundary for this path is the C++
 * class API plus the PyISR CPython module.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference tree).  All pointers are HOST pointers unless the
 * parameter name ends in _dev.  Matrices cross the ABI COLUMN-MAJOR (Eigen::MatrixXd storage, what
 * extractIntOpMatrix / extractBCSCovMatrix hand to NumPy: src/ISRSolverSLE.cpp:19-25).  Every call
 * returns 0 on success or an ISR_E* code; isr_last_error() gives the message.  No exceptions cross
 * the ABI.  Calls on one isr_problem must not run concurrently; different problems may.
 *
 * There is no CPU fallback: without a CUDA device isr_ctx_create fails with ISR_ECUDA.
 */
#ifndef ISR_B200_H
#define ISR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISR_OK 0
#define ISR_EINVAL 1      /* bad argument */
#define ISR_ERANGE 2      /* InterpRangeException (include/Interpolator.hpp:14-18) */
#define ISR_ECUDA 3       /* CUDA / cuSOLVER / cuBLAS failure */
#define ISR_ESTATE 4      /* call order violated (e.g. toy batch before factor) */
#define ISR_ENUMERIC 5    /* singular matrix / non-SPD regulariser */
#define ISR_EEFFDIM 6     /* EfficiencyDimensionException (include/BaseISRSolver.hpp:24-28) */

typedef struct isr_ctx isr_ctx;         /* one CUDA device + stream + library handles */
typedef struct isr_problem isr_problem; /* one solver instance: energies, basis, epsilon, G, factors */

/* Detection efficiency.  Replaces std::function<double(double,double)> efficiency
 * (include/BaseISRSolver.hpp:36-41) as built by BaseISRSolver::_setupEfficiency
 * (src/BaseISRSolver.cpp:232-258) and BaseISRSolver2::_setupEfficiency (src/BaseISRSolver2.cpp:231-236).
 * `values` include ROOT's underflow (index 0) and overflow (index nbins+1) cells. */
enum {
  ISR_EPS_ONE = 0,       /* efficiency == 1 (src/BaseISRSolver.cpp:140) */
  ISR_EPS_ROW_TABLE = 1, /* v2: one TH1D in x per energy point; values[N][nx+2] */
  ISR_EPS_TABLE_2D = 2,  /* TEfficiency 2-D in (x, E); values[(ne+2)][(nx+2)], global bin bx+(nx+2)*by */
  ISR_EPS_E_ONLY = 3     /* TEfficiency 1-D: a function of E only; values[ne+2] */
};
typedef struct {
  int kind;
  int nx; double xlo, xhi; const double* xedges; /* xedges == NULL: nx fixed-width bins on [xlo,xhi) */
  int ne; double elo, ehi; const double* eedges; /* same for the energy axis */
  const double* values;
} isr_eps_desc;

/* ---- context ---------------------------------------------------------------------------------- */
int isr_ctx_create(int device, isr_ctx** out);
void isr_ctx_destroy(isr_ctx* ctx);
const char* isr_last_error(void);
/* cudaStream_t used by all launches of this context (as void*), for event timing by the caller. */
void* isr_ctx_stream(isr_ctx* ctx);
/* Page-locked host memory for the arrays the host-buffer calls move in bulk (the toys of isr_toy_batch, the chi2 array):
 * from pageable memory such a copy runs at a fraction of the PCIe rate.  Plain cudaMallocHost / cudaFreeHost. */
int isr_host_alloc(isr_ctx* ctx, size_t bytes, void** out);
void isr_host_free(void* ptr);
int isr_ctx_synchronize(isr_ctx* ctx);
/* Number of kernels of THIS library launched by this process so far (diagnostic; bench.py's gpu_launches). */
int64_t isr_launch_count(void);

/* ---- problem ---------------------------------------------------------------------------------- */
/* Replaces BaseISRSolver(n, energy, ...) + Interpolator(settings, ecm, threshold)
 * (src/BaseISRSolver.cpp:77-129; src/Interpolator.cpp:34-83).  `ecm` must already be sorted ascending
 * (the host class applies the reference's sort permutation).  ranges: nranges x {cubic?, firstSeg,
 * lastSeg} as in the interpolation-settings JSON (README.md:177-188); nranges == 0 selects the default
 * single linear range (src/Interpolator.cpp:356-360).  Invalid ranges -> ISR_ERANGE. */
int isr_problem_create(isr_ctx* ctx, int n, const double* ecm, double threshold, int nranges,
                       const int* ranges, const isr_eps_desc* eps, isr_problem** out);
void isr_problem_destroy(isr_problem* p);
/* ISRSolverSLE::setRangeInterpSettings (src/ISRSolverSLE.cpp:125-136) */
int isr_problem_set_ranges(isr_problem* p, int nranges, const int* ranges);

/* ISRSolverSLE::evalEqMatrix (src/ISRSolverSLE.cpp:138-149): G(i,j) = integral of F * basis_j * eps,
 * optionally G <- G_spread * G (src/ISRSolverSLE.cpp:177-193) when ecm_err != NULL.
 * G_out (N*N, column-major) may be NULL (matrix stays on the device). */
int isr_assemble(isr_problem* p, const double* ecm_err, double* G_out);
/* Host-sampled efficiency ("SAMPLED"): eps(x, E) is an arbitrary host callable, as the reference's
 * std::function / Python callable is (include/BaseISRSolver.hpp:36-41, include/PyISRSolver.hpp:189-197).
 * isr_assemble_nodes enumerates the quadrature panels of a problem and
 * returns, for every node, x and the energy-point index i (evaluate eps(x, ecm[i]) there; pass NULL buffers to
 * query *count only); isr_assemble_sampled then assembles G with those values entering the integrand node by
 * node.  The callable must be smooth between break points: besides the kernel's own (x0, grading), the bin edges of an
 * ISR_EPS_ROW_TABLE / ISR_EPS_TABLE_2D given at problem creation are honoured (its values are then ignored) -- pass a
 * table whose edges are the callable's discontinuities. */
int isr_assemble_nodes(isr_problem* p, int64_t cap, double* x_out, int* row_out, int64_t* count);
int isr_assemble_sampled(isr_problem* p, int64_t n_nodes, const double* eps_values, const double* ecm_err, double* G_out);
/* Inject a matrix instead of assembling it (tests of the solve path; column-major). */
int isr_set_matrix(isr_problem* p, const double* G);
/* Number of quadrature panels / integrand nodes of the last isr_assemble (roofline unit counts). */
int isr_assemble_stats(isr_problem* p, int64_t* n_panels, int64_t* n_nodes);
/* Debug/inspection: copy out the panel list of the last assembly.  Each panel = 4 doubles
 * {a, b, eps, kind} and 2 ints {row, segment}.  Pass NULL to query the count only. */
int isr_debug_panels(isr_problem* p, int64_t cap, double* ab_eps_kind, int* row_seg, int64_t* count);
/* Debug/inspection: per-(row, segment) normalised moments M[i][k][0..3] (N*N*4, row-major). */
int isr_debug_moments(isr_problem* p, double* M_out);

/* kernelKuraevFadin(x, s) evaluated on the device (include/KuraevFadin.hpp:13; PyISR
 * "kernelKuraevFadin", include/PyKuraevFadin.hpp:9-17). */
int isr_kernel_kf(isr_ctx* ctx, int64_t n, const double* x, const double* s, double* out);

/* Basis queries (Interpolator::basisEval / basisDerivEval / evalIntegralBasis,
 * src/Interpolator.cpp:217-287, 459-473), evaluated from the device-resident coefficient tables. */
int isr_basis_eval(isr_problem* p, int64_t n, const int* cs_index, const double* energy, int deriv, double* out);
int isr_basis_integrals(isr_problem* p, double* w_out /* N */);
/* ISRSolverSLE::interpEval (src/ISRSolverSLE.cpp:195-198) */
int isr_interp_eval(isr_problem* p, const double* y, int64_t n, const double* energy, double* out);

/* ---- solve ------------------------------------------------------------------------------------ */
/* One-time factorisation for a given error vector: S = G^-1, R = W^(1/2) G, Cov = S W^-1 S^T
 * (replaces the per-call COD + (G^T W G)^-1 of ISRSolverSLE::solve, src/ISRSolverSLE.cpp:91-94). */
int isr_factor(isr_problem* p, const double* vcs_err);
/* ISRSolverSLE::solve: bcs = G^-1 vcs; cov_out (N*N col-major) may be NULL. */
int isr_solve(isr_problem* p, const double* vcs, double* bcs_out, double* cov_out);
/* IterISRInterpSolver::solve (src/IterISRInterpSolver.cpp:45-57) on the assembled operator: b <- vcs, then niter
 * times  radcorr = (G b) / b,  b <- vcs / radcorr.  radcorr_out (N) may be NULL; the covariance of that method is
 * diag((vcs_err / radcorr)^2), formed by the caller. */
int isr_iterative_solve(isr_problem* p, int niter, const double* vcs, double* bcs_out, double* radcorr_out);
int isr_get_inverse(isr_problem* p, double* Ginv_out);         /* column-major */
int isr_get_inv_cov(isr_problem* p, double* inv_cov_out);      /* G^T W G, column-major */
/* ISRSolverSLE::evalConditionNumber (src/ISRSolverSLE.cpp:200-204) */
int isr_cond_number(isr_problem* p, double* cond_out);

/* Condition number versus beam-energy spread (src/isrsolver-condnum-test.cpp:88-140: resetECMErrors +
 * evalEqMatrix + evalConditionNumber per point).  sigma: K common spreads (per_point == 0; 0 means "spread
 * disabled") or K x N per-point spreads (row-major).  cond_out[K]; sv_out (K x N singular values,
 * descending) may be NULL.  Leaves the problem holding the UNSPREAD matrix. */
int isr_cond_spread_scan(isr_problem* p, int K, const double* sigma, int per_point, double* cond_out, double* sv_out);

/* ---- toy Monte-Carlo batch -------------------------------------------------------------------- */
/* The loop bodies of chi2Test (src/Chi2Test.cpp:38-60), chi2TestData (:177-183) and ratioTestModel
 * (src/RatioTest.cpp:38-45) over T toys at once:
 *   b_k = S v_k ;  chi2_k = || R (b_k - bcs0) ||^2  (== (b_k-b0)^T Cov^-1 (b_k-b0))
 *   sum_bcs[j] = sum_k b_k[j] ;  ratio_stats[j] = {count, sum r, sum r^2} of r = (b_k[j]-b0[j])/b0[j]
 *   restricted to -2 <= r < 2 (TH1F(1024,-2,2) in-range statistics) ; ratio_hist[j][0..1025] counts.
 * V is toy-major T x N.  Any output pointer may be NULL.  Uses the solver set by the last
 * isr_factor / isr_tikhonov_select / isr_tsvd_select. */
int isr_toy_batch(isr_problem* p, int64_t T, const double* V, const double* bcs0, double* chi2_out,
                  double* sum_bcs_out, double* ratio_stats_out, uint32_t* ratio_hist_out, double* bcs_out);
/* Same with device-resident V / outputs (accumulators are ADDED to, caller zeroes them). */
int isr_toy_batch_dev(isr_problem* p, int64_t T, const double* V_dev, const double* bcs0,
                      double* chi2_dev, double* sum_bcs_dev, double* ratio_stats_dev,
                      uint32_t* ratio_hist_dev, double* bcs_dev);
/* Toy-kernel schedule.  0 (default): operators that are exactly lower triangular -- G^-1 and W^(1/2) G of linear
 * interpolation without energy spread -- skip their structural zeros (bit-identical results, about 2 N^2 instead of
 * 4 N^2 flop per toy).  1: dense schedule on the operators as given (cross-check).  2: as 0, and a dense chi2 factor R is
 * replaced at the first chi2 batch by the lower-triangular factor L of its QL decomposition (L^T L = R^T R, so
 * |L d|^2 = |R d|^2 = d^T Cov^-1 d of src/Chi2Test.cpp:51-55 up to rounding): 3 N^2 flop per toy, for campaigns long
 * enough (about 10^6 toys) to repay the Householder QR. */
int isr_set_toy_kernel(isr_problem* p, int variant);
/* Name of the tile shape the fused toy kernel uses at N points -- m<toys per tile>n<columns per block>w<warps>[s<stages>]
 * [k<K chunk>] -- with no column statistics (nstat 0), the sum of solutions (1) or the ratio statistics (4); NULL if no
 * shape fits the 227 KB of shared memory.  Pure host logic (no device needed): diagnostics and bench labels. */
const char* isr_toy_tile_name(int n, int nstat);

/* Toy campaigns by COUNT, as the reference's chi2Test / chi2TestData / ratioTestModel take them (src/Chi2Test.cpp:29-60,
 * 160-195; src/RatioTest.cpp:13-45): toys first_toy .. first_toy + T - 1 of the stream `seed` are drawn on the device,
 *   V[k][i] = vcs[i] + vcs_err[i] * z(seed, k, i)      (randomDrawVisCS, src/Utils.cpp:6-13)
 * with z from counter-based Philox4x32-10 + Box-Muller (isrsolver_b200/csrc/rng.cu; the pair is formed in single
 * precision on the special-function unit, |z| <= 6.76, unless ISR_TOY_RNG=fp64 is set) -- a toy depends only on
 * (seed, k, i), so a campaign sharded over ranks by first_toy reproduces the single-GPU campaign exactly.
 * ROOT's unseeded global TRandom3 stream cannot be reproduced; isr_draw_toys returns the very vectors a campaign
 * uses, for checking it against any other implementation.  Outputs as isr_toy_batch; additionally the
 * TH1F(nbins, min, max) chi2 histogram of src/Chi2Test.cpp:64-78 (chi2_lo_hi_out[2], chi2_counts_out[nbins + 2])
 * computed without moving chi2 to the host.  Any output pointer may be NULL. */
int isr_toy_campaign(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* vcs, const double* vcs_err,
                     const double* bcs0, double* chi2_out, double* sum_bcs_out, double* ratio_stats_out,
                     uint32_t* ratio_hist_out, int nbins, double* chi2_lo_hi_out, uint32_t* chi2_counts_out);
/* Same with device-resident outputs (accumulators are ADDED to); asynchronous on the context stream. */
int isr_toy_campaign_dev(isr_problem* p, int64_t T, int64_t first_toy, uint64_t seed, const double* vcs, const double* vcs_err,
                         const double* bcs0, double* chi2_dev, double* sum_bcs_dev, double* ratio_stats_dev,
                         uint32_t* ratio_hist_dev);
int isr_draw_toys(isr_ctx* ctx, int64_t T, int n, int64_t first_toy, uint64_t seed, const double* vcs, const double* vcs_err,
                  double* V_out /* T x n toy-major, host */);

/* chi2 histogram, src/Chi2Test.cpp:64-78: TH1F(128, min, max) filled with every chi2 (ROOT bin
 * arithmetic; the maximum lands in the overflow cell).  counts_out has nbins+2 entries. */
int isr_minmax_dev(isr_ctx* ctx, int64_t n, const double* v_dev, double* min_out, double* max_out);
/* Fully device-resident, asynchronous forms for multi-GPU pipelines: out_dev = {min, -max}, so that ONE
 * all_reduce(MIN) over the ranks yields the global range in the same layout; isr_histogram_range_dev bins
 * against a range stored that way in device memory into int32 counters (overwritten) ready for an all_reduce(SUM). */
int isr_minmax_to_dev(isr_ctx* ctx, int64_t n, const double* v_dev, double* out_dev);
int isr_histogram_range_dev(isr_ctx* ctx, int64_t n, const double* v_dev, int nbins, const double* range_dev, int32_t* counts_dev);
int isr_histogram_dev(isr_ctx* ctx, int64_t n, const double* v_dev, int nbins, double lo, double hi,
                      uint32_t* counts_out);
/* TH1F(nbins, lo, hi) of the chi2 values the last isr_toy_batch / isr_toy_campaign of this problem left on the
 * device (multi-GPU: re-bin against the GLOBAL range without moving chi2). */
int isr_last_chi2_histogram(isr_problem* p, int nbins, double lo, double hi, uint32_t* counts_out);
/* min / max of the same device-resident values (the range src/Chi2Test.cpp:64-71 takes from the toy loop). */
int isr_last_chi2_minmax(isr_problem* p, double* min_out, double* max_out);
/* TH1F(nbins, lo, hi) of a host array with a caller-supplied range (multi-GPU: the GLOBAL min / max). */
int isr_histogram(isr_ctx* ctx, int64_t n, const double* v, int nbins, double lo, double hi, uint32_t* counts_out);
int isr_chi2_histogram(isr_ctx* ctx, int64_t n, const double* chi2, int nbins, double* lo_out,
                       double* hi_out, uint32_t* counts_out);

/* ---- multi-GPU: communicators ------------------------------------------------------------------ */
/* The path shards with no data-path collective (toys, lambdas, spread settings are independent); the only exchange of a
 * chi2 test is two small all-reduces over NVLink (NCCL, loaded from libnccl.so.2 when the first communicator is made):
 *   MIN over ranks of {min chi2, -max chi2}   -> the global TH1F range of src/Chi2Test.cpp:64-71, and
 *   SUM over ranks of ONE packed FP64 buffer  [TH1F cells | sum of solutions (N) | ratio statistics (3N) | toy count | flag]
 * (integer counters are exact in FP64 up to 2^53), plus the N x 1026 ratio cells when ratio_hist_out is requested.
 * A communicator binds one context, i.e. one GPU; it can be made
 *   - for several contexts of ONE process (isr_comm_init_all; isr_group below wraps this with one host thread per GPU), or
 *   - one per process (isr_comm_unique_id on rank 0, distributed by the caller, then isr_comm_init_rank everywhere). */
typedef struct isr_comm isr_comm;
#define ISR_COMM_ID_BYTES 128
int isr_comm_unique_id(void* id_out /* ISR_COMM_ID_BYTES */);
int isr_comm_init_rank(isr_ctx* ctx, int nranks, int rank, const void* id, isr_comm** out);
int isr_comm_init_all(int n, isr_ctx* const* ctxs, isr_comm** out /* n */);
void isr_comm_destroy(isr_comm* comm);
/* 1 if the small exchanges of this communicator go over NVLink peer-memory mailboxes (fused into the kernels that
 * produce / consume them; rank-ordered, hence bit-reproducible sums), 0 if they are NCCL all-reduces (peers not mappable,
 * more than 16 ranks, or ISR_NO_P2P=1 in the environment). */
int isr_comm_uses_peer_memory(const isr_comm* comm);
int isr_comm_rank(const isr_comm* comm);
int isr_comm_size(const isr_comm* comm);
/* In-place all-reduce of a device buffer on the context stream (asynchronous).  dtype: 0 double, 1 int32, 2 uint32,
 * 3 int64; op: 0 sum, 1 min, 2 max. */
int isr_comm_allreduce_dev(isr_comm* comm, void* buf_dev, int64_t count, int dtype, int op);
/* Bind the calling host thread to the CPUs of the NUMA node the context's GPU hangs off (PCI topology from sysfs);
 * returns the node through *node_out (-1: unknown, nothing done).  isr_host_alloc binds temporarily by itself, so that
 * page-locked toy buffers are NUMA-local to the GPU that reads them. */
int isr_ctx_bind_host_thread(isr_ctx* ctx, int* node_out);

/* ---- a whole chi2 / ratio test in ONE call ------------------------------------------------------ */
/* chi2TestModel / chi2Test (src/Chi2Test.cpp:113-151, 29-104) and ratioTestModel (src/RatioTest.cpp:13-62) end to end:
 * [assemble] -> [factor] -> toys -> min / max -> TH1F cells, everything enqueued on the context stream with ONE
 * synchronisation at the end: the certificate of the unpivoted inverse is read there (if it failed the factorisation is
 * redone with partial pivoting and the toys are re-run; *refactored_out tells), the first toys are copied / drawn while
 * the matrix is still being assembled, and with comm != NULL the two all-reduces above run in the same stream. */
#define ISR_TEST_ASSEMBLE 1 /* (re)assemble the matrix first (isr_assemble with ecm_err) */
#define ISR_TEST_FACTOR 2   /* factor for vcs_err (isr_factor); without it the currently selected solver is used */
#define ISR_TEST_RATIO 4    /* ratio statistics (and the ratio cells if ratio_hist_out is given) */
typedef struct {
  int flags;
  const double* ecm_err; /* N or NULL: beam-energy spread of the assembly */
  const double* vcs_err; /* N: weights of the factorisation / widths of device-drawn toys */
  const double* bcs0;    /* N: reference solution */
  int64_t n_toys;        /* toys of THIS call (this rank's share when comm != NULL; may be 0) */
  const double* V;       /* host toys, toy-major n_toys x N, or NULL */
  const double* V_dev;   /* device toys, or NULL.  V == V_dev == NULL: toys first_toy .. of stream `seed` are drawn on the device */
  const double* vcs;     /* N: centre of device-drawn toys */
  uint64_t seed;
  int64_t first_toy;
  int nbins;             /* TH1F cells (128 in src/Chi2Test.cpp:72); 0: no chi2 histogram */
  isr_comm* comm;        /* NULL: this GPU only */
  /* outputs, host memory, any may be NULL; with comm != NULL range, cells and sums are GLOBAL (identical on every rank) */
  double* chi2_out;          /* n_toys: chi2 of this call's toys */
  double* chi2_lo_hi_out;    /* 2 */
  int64_t* chi2_counts_out;  /* nbins + 2 (under / overflow cells included) */
  double* sum_bcs_out;       /* N */
  double* ratio_stats_out;   /* 3N: {count, sum r, sum r^2} per point */
  uint32_t* ratio_hist_out;  /* N x 1026 */
  int64_t* total_toys_out;   /* toys over all ranks */
  int* refactored_out;       /* 1: certificate failed, pivoted factorisation used */
} isr_chi2_test_args;
int isr_chi2_test(isr_problem* p, const isr_chi2_test_args* args);
/* CUDA-event times (ms, on the context stream) of the phases of the last isr_chi2_test of this problem:
 * ms_out[0] assembly, [1] factorisation (inverse + operators; the certificate runs on a side stream), [2] toy kernels,
 * [3] range + collectives + TH1F cells + copy of the packed result. */
int isr_chi2_test_timing(isr_problem* p, double* ms_out /* 4 */);

/* ---- several GPUs of one box from ONE process ---------------------------------------------------- */
/* What a C++ caller of chi2TestModel / ratioTestModel (include/Chi2Test.hpp:18-42, include/RatioTest.hpp:14-15) uses to
 * spread a test over the GPUs of a box: one context, one replicated problem and one host thread per device, one
 * communicator over all of them.  Problems are replicated (assembly + factorisation are << 1 ms and deterministic);
 * toys are cut into contiguous slices.  Calls on one group must not run concurrently. */
typedef struct isr_group isr_group;
typedef struct isr_gproblem isr_gproblem;
int isr_group_create(int ngpu, const int* devices /* NULL: 0 .. ngpu-1 */, isr_group** out);
void isr_group_destroy(isr_group* g);
int isr_group_size(const isr_group* g);
isr_ctx* isr_group_ctx(isr_group* g, int rank);
int isr_group_problem_create(isr_group* g, int n, const double* ecm, double threshold, int nranges, const int* ranges,
                             const isr_eps_desc* eps, isr_gproblem** out);
void isr_group_problem_destroy(isr_gproblem* gp);
isr_problem* isr_group_problem_member(isr_gproblem* gp, int rank);
/* Page-locked buffer for the toys of isr_group_chi2_test (toy-major n_toys x n doubles) whose slice for each GPU is
 * first-touched by that GPU's host thread, i.e. NUMA-local to the device that will read it. */
int isr_group_host_alloc(isr_group* g, int64_t n_toys, int n, double** out);
void isr_group_host_free(isr_group* g, double* ptr);
/* isr_chi2_test over all GPUs of the group: args->n_toys is the TOTAL, args->V (if given) the whole host array, args->comm
 * and args->V_dev must be NULL; chi2_out receives all n_toys values in toy order, the other outputs are global.
 * Histogram cells are bit-identical to the single-GPU test (every rank bins against the global range). */
int isr_group_chi2_test(isr_gproblem* gp, const isr_chi2_test_args* args);
/* isr_cond_spread_scan with the K spread settings cut into one contiguous slice per GPU (no exchange at all). */
int isr_group_cond_spread_scan(isr_gproblem* gp, int K, const double* sigma, int per_point, double* cond_out, double* sv_out);

/* ---- Tikhonov --------------------------------------------------------------------------------- */
/* One decomposition shared by every lambda (replaces _evalProblemMatrices + two FullPivLU per lambda,
 * src/ISRSolverTikhonov.cpp:139-155): F = D^T diag(w) D (deriv_reg != 0) or diag(w);
 * generalised eigen-system (G^T W G) u = mu F u, U^T F U = I.  Needs isr_factor's vcs_err. */
int isr_tikhonov_prepare(isr_problem* p, const double* vcs_err, int deriv_reg);
/* For each lambda: solve (:61-72) + evalEqNorm2 (:106-109) + evalSmoothnessConstraintNorm2 (:111-113)
 * + evalLCurveCurvature (:115-119) + evalLCurveCurvatureDerivative (:121-128).
 * bcs_out: L x N row-major, may be NULL (as any other output). */
int isr_tikhonov_scan(isr_problem* p, int64_t L, const double* lambda, const double* vcs,
                      double* eq_norm2_out, double* smooth_norm2_out, double* curv_out,
                      double* dcurv_out, double* bcs_out);
/* Full solution for one lambda: bcs and Cov = A+ W^-1 A+^T (col-major; may be NULL). */
int isr_tikhonov_solve(isr_problem* p, double lambda, const double* vcs, double* bcs_out, double* cov_out);
/* chi2 of every (lambda, toy) pair: chi2_out[l*T + k]; the chi2Test loop of
 * isrsolver-Tikhonov-chi2-test run for L lambdas at once.  V toy-major T x N (host). */
int isr_tikhonov_toy_scan(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V,
                          const double* bcs0, double* chi2_out);
int isr_tikhonov_toy_scan_dev(isr_problem* p, int64_t L, const double* lambda, int64_t T,
                              const double* V_dev, const double* bcs0, double* chi2_dev);
/* The chi2 test of isrsolver-Tikhonov-chi2-test (src/Chi2Test.cpp:38-78) for L lambdas at once, summarised on
 * the device: stats_out[l] = {min, max, mean} of chi2 over the T toys, counts_out[l][0..nbins+1] = the
 * TH1F(nbins, min_l, max_l) cells (under/overflow included).  The L x T chi2 matrix never leaves the GPU. */
int isr_tikhonov_toy_scan_hist(isr_problem* p, int64_t L, const double* lambda, int64_t T, const double* V,
                               const double* bcs0, int nbins, double* stats_out, uint32_t* counts_out);
/* The same per-lambda chi2 test with the TOYS sharded over the ranks of a communicator (BASELINE config C4 on 2/4/8 GPUs):
 * every rank passes its slice of the toys (host V or device V_dev; n_toys may be 0); the per-lambda {min, max, mean} and
 * TH1F cells returned are GLOBAL and identical on every rank -- one all-reduce(MIN) of 2 L doubles for the ranges, one
 * all-reduce(SUM) of a packed FP64 buffer for cells and sums.  flags: ISR_TEST_ASSEMBLE (re-assemble with ecm_err first),
 * ISR_TEST_FACTOR (isr_tikhonov_prepare for vcs_err / deriv_reg first; the toys are copied meanwhile). */
typedef struct {
  int flags;
  const double* ecm_err;   /* N or NULL */
  const double* vcs_err;   /* N (ISR_TEST_FACTOR) */
  int deriv_reg;
  int64_t n_lambda;
  const double* lambda;
  int64_t n_toys;          /* this rank's share */
  const double* V;         /* host, toy-major n_toys x N, or NULL */
  const double* V_dev;     /* device, or NULL */
  const double* bcs0;      /* N */
  int nbins;
  isr_comm* comm;          /* NULL: this GPU only */
  double* stats_out;       /* L x 3 {min, max, mean}, or NULL */
  uint32_t* counts_out;    /* L x (nbins + 2), or NULL */
  int64_t* total_toys_out; /* or NULL */
} isr_tikhonov_scan_args;
int isr_tikhonov_chi2_scan(isr_problem* p, const isr_tikhonov_scan_args* args);
/* The same over all GPUs of an isr_group from one process: args->n_toys is the TOTAL, args->V the whole host array;
 * args->comm and args->V_dev must be NULL. */
int isr_group_tikhonov_chi2_scan(isr_gproblem* gp, const isr_tikhonov_scan_args* args);
/* (lambda, toy) scan variant: 0 (default) = chi2 as a quadratic form in lambda from two dot products per toy
 * (O(1) per pair, bound by writing the output); 1 = the direct O(N)-per-pair kernel (kept as a cross-check). */
int isr_set_scan_kernel(isr_problem* p, int variant);
/* Make isr_toy_batch use the Tikhonov operator A+_lambda and its covariance. */
int isr_tikhonov_select(isr_problem* p, double lambda);
int isr_get_regulariser(isr_problem* p, double* F_out); /* column-major */

/* ---- TSVD ------------------------------------------------------------------------------------- */
/* ISRSolverTSVD::solve (src/ISRSolverTSVD.cpp:53-74): one SVD, rank-k (or keep-one) pseudo-inverse.
 * cov_out is defined only for k == N and !keep_one (the reference inverts a singular matrix otherwise). */
int isr_tsvd_solve(isr_problem* p, int k, int keep_one, const double* vcs, const double* vcs_err,
                   double* bcs_out, double* cov_out);
int isr_singular_values(isr_problem* p, double* sv_out /* N, descending */);
int isr_tsvd_select(isr_problem* p, int k, int keep_one, const double* vcs_err);

/* ---- forward convolution service ---------------------------------------------------------------- */
/* convolutionKuraevFadin / convolutionKuraevFadin_1d (src/KuraevFadin.cpp:56-92; PyISR
 * "convolutionKuraevFadin", include/PyKuraevFadin.hpp:19-62) for M energies at once:
 *   result[m] = int_{min_x[m]}^{max_x[m]} F(x, E_m^2) eps(x, E_m) f(E_m sqrt(1 - x)) dx .
 * The Born cross section f is a host callable, so the plan exposes its quadrature nodes: sample f at the
 * node energies (isr_conv_plan_nodes), hand the values back (isr_conv_plan_feed), bisect the panels whose
 * Gauss-Kronrod error estimate is too large (isr_conv_plan_refine) and repeat until no panel is split.
 * Each pending panel exposes 16 node slots (15 Kronrod nodes + one duplicate with zero weight).
 * eps may be NULL (== 1); for ISR_EPS_ROW_TABLE the table has one row per requested energy.
 * break_energy: optional energies where f has a kink or a narrow peak (panels are cut there). */
typedef struct isr_conv_plan isr_conv_plan;
int isr_conv_plan_create(isr_ctx* ctx, int64_t M, const double* energy, const double* min_x, const double* max_x,
                         const isr_eps_desc* eps, int nbreaks, const double* break_energy, isr_conv_plan** out);
void isr_conv_plan_destroy(isr_conv_plan* plan);
int isr_conv_plan_size(isr_conv_plan* plan, int64_t* n_panels, int64_t* n_pending_nodes);
/* Nodes of the pending panels, n_pending_nodes entries each (any pointer may be NULL): E' = E sqrt(1 - x),
 * x, the row index m, and the complete node weight w (so that sum w f(E') is the panel's Kronrod value). */
int isr_conv_plan_nodes(isr_conv_plan* plan, double* energy_out, double* x_out, int* row_out, double* weight_out);
int isr_conv_plan_feed(isr_conv_plan* plan, const double* fvals /* n_pending_nodes */);
/* f = the interpolated Born cross section of a solver (Interpolator::eval, src/Interpolator.cpp:295-323),
 * sampled on the device without a host round trip. */
int isr_conv_plan_feed_interp(isr_conv_plan* plan, isr_problem* p, const double* bcs /* N */);
/* Bisect every evaluated panel whose error estimate exceeds its share of max(abs_tol, rel_tol |result|);
 * *n_split panels were replaced by two pending children each. */
int isr_conv_plan_refine(isr_conv_plan* plan, double rel_tol, double abs_tol, int64_t* n_split);
int isr_conv_plan_result(isr_conv_plan* plan, double* result /* M */, double* err_est /* M or NULL */);
/* Keep the refined panel set, mark every panel pending again (same grid, new f). */
int isr_conv_plan_reset(isr_conv_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* ISR_B200_H */
