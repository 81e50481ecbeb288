"""CPU oracle for the ISRSolver hot path -- ORACLE / TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (isrsolver_b200 -> libisr_b200.so) never touches it.

PINNING.  The assembly layer (isr_oracle.c) reproduces the reference's own compiled sources bit for bit
(oracle/_ref, oracle/isr_ref.py, tests/test_ref_pin.py).  The linear-algebra layer in THIS file (solve,
covariance, toy loops, histograms, Tikhonov, TSVD) is pinned by RUNS OF THE REFERENCE ITSELF: its solver
classes and toy-test drivers (BaseISRSolver{,2}, ISRSolverSLE{,2}, ISRSolverTikhonov, ISRSolverTSVD,
Chi2Test, RatioTest, Utils) compile unmodified against oracle/ref_shim (dense Eigen stand-in, ROOT
containers over an in-memory registry), their outputs on the fixtures' inputs are committed as
tests/golden/refrun_*.npz (tests/golden/make_refrun_golden.py), and tests/test_refrun_pin.py requires
this file to reproduce them to 1e-12 (histogram cells bit for bit).  Second anchor: the same formulas in
60-digit mpmath arithmetic (test_linear_algebra_layer_against_exact_arithmetic), LAPACK identities and
the eigen-form identities (tests/test_oracle.py).

Two layers:
  * oracle/isr_oracle.c   kernel, QUADPACK QAGS/QAG61, interpolation bases, efficiency lookups and
                          the N^2 loop of ISRSolverSLE::evalEqMatrix (src/ISRSolverSLE.cpp:138-149)
  * this file             the dense linear algebra of solve()/chi2Test/ratioTest/Tikhonov/TSVD,
                          restated with NumPy/SciPy LAPACK calls in place of Eigen:
                            completeOrthogonalDecomposition().solve -> LAPACK xGELSY (a COD)
                            .inverse()                              -> xGETRF/xGETRI
                            FullPivLU::solve                        -> xGESV
                            JacobiSVD                               -> xGESDD
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALPHA_QED = 7.2973525698e-03      # include/PhysicalConstants.hpp:7
ELECTRON_M = 5.10998928e-04       # include/PhysicalConstants.hpp:11

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile oracle/libisr_oracle.so (gcc, seconds)."""
    so = os.path.join(HERE, "libisr_oracle.so")
    src = os.path.join(HERE, "isr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libisr_oracle.so"])
    return so


class _Eff(C.Structure):
    _fields_ = [("kind", C.c_int), ("nx", C.c_int), ("xlo", C.c_double), ("xhi", C.c_double),
                ("xedges", _dp), ("ne", C.c_int), ("elo", C.c_double), ("ehi", C.c_double),
                ("eedges", _dp), ("values", _dp)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_kernel_kf.restype = C.c_double
        L.orc_kernel_kf.argtypes = [C.c_double, C.c_double]
        L.orc_beta.restype = C.c_double
        L.orc_beta.argtypes = [C.c_double]
        L.orc_interp_create.restype = C.c_void_p
        L.orc_interp_create.argtypes = [C.c_int, _dp, C.c_double, C.c_int, _ip]
        L.orc_interp_free.argtypes = [C.c_void_p]
        L.orc_interp_nranges.argtypes = [C.c_void_p]
        L.orc_interp_range.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _ip, _dp, _dp]
        L.orc_interp_spline_c.restype = _dp
        L.orc_interp_spline_c.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for nm in ("orc_basis_eval", "orc_basis_deriv_eval"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_interp_eval.restype = C.c_double
        L.orc_interp_eval.argtypes = [C.c_void_p, _dp, C.c_double]
        L.orc_kf_basis_integral.restype = C.c_double
        L.orc_kf_basis_integral.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_integral_basis.restype = C.c_double
        L.orc_integral_basis.argtypes = [C.c_void_p, C.c_int]
        L.orc_eval_eq_matrix.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int]
        L.orc_eval_eq_matrix_ex.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
        L.orc_kf_basis_integral_ex.restype = C.c_double
        L.orc_kf_basis_integral_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, _dp, _dp]
        L.orc_energy_spread_matrix.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_deriv_projector.argtypes = [C.c_void_p, _dp]
        L.orc_find_bin.restype = C.c_int
        L.orc_find_bin.argtypes = [C.c_int, C.c_double, C.c_double, _dp, C.c_double]
        L.orc_bsearch.restype = C.c_int
        L.orc_bsearch.argtypes = [_dp, C.c_double, C.c_int, C.c_int]
        L.orc_bw_born.restype = C.c_double
        L.orc_bw_born.argtypes = [C.c_double] * 5
        L.orc_bw_visible.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_void_p, _dp, C.c_int]
        for nm in ("orc_probe_qags", "orc_probe_qag61"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_double, _dp, _dp, C.POINTER(C.c_long)]
        L.orc_get_neval.restype = C.c_long
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


# ------------------------------------------------------------------------------------------------
# kernel / quadrature probes
# ------------------------------------------------------------------------------------------------
def kernel_kf(x: float, s: float) -> float:
    """kernelKuraevFadin(x, s) -- src/KuraevFadin.cpp:14-49."""
    return lib().orc_kernel_kf(float(x), float(s))


def beta(s: float) -> float:
    return lib().orc_beta(float(s))


def probe_quad(which: str, kind: int, p0: float, p1: float, a: float, b: float,
               epsabs: float = 1e-12, epsrel: float = 1e-12):
    """Run the QUADPACK restatement on a built-in integrand family; returns (result, abserr, neval, status)."""
    r, e, n = C.c_double(), C.c_double(), C.c_long()
    fn = lib().orc_probe_qags if which == "qags" else lib().orc_probe_qag61
    st = fn(kind, p0, p1, a, b, epsabs, epsrel, C.byref(r), C.byref(e), C.byref(n))
    return r.value, e.value, n.value, st


# ------------------------------------------------------------------------------------------------
# efficiency descriptors (ROOT histogram semantics; src/BaseISRSolver.cpp:232-258, BaseISRSolver2.cpp:231-236)
# ------------------------------------------------------------------------------------------------
class Efficiency:
    """kind: 'one' | 'row_table' (v2: one TH1D per energy point) | 'table_2d' (TEfficiency 2-D in (x, E))
    | 'e_only' (TEfficiency 1-D: function of E).  `values` include under/overflow bins."""

    def __init__(self, kind="one", values=None, nx=0, xlo=0.0, xhi=1.0, xedges=None,
                 ne=0, elo=0.0, ehi=1.0, eedges=None):
        self.kind = kind
        self._keep = []
        s = _Eff()
        s.kind = {"one": 0, "row_table": 1, "table_2d": 2, "e_only": 3}[kind]
        s.nx, s.xlo, s.xhi = int(nx), float(xlo), float(xhi)
        s.ne, s.elo, s.ehi = int(ne), float(elo), float(ehi)
        if xedges is not None:
            a, p = _d(xedges); self._keep.append(a); s.xedges = p
            s.nx, s.xlo, s.xhi = len(a) - 1, float(a[0]), float(a[-1])
        if eedges is not None:
            a, p = _d(eedges); self._keep.append(a); s.eedges = p
            s.ne, s.elo, s.ehi = len(a) - 1, float(a[0]), float(a[-1])
        if values is not None:
            a, p = _d(values); self._keep.append(a); s.values = p
            self.values = a
        self.nx, self.xlo, self.xhi, self.ne, self.elo, self.ehi = s.nx, s.xlo, s.xhi, s.ne, s.elo, s.ehi
        self.xedges = None if xedges is None else np.asarray(xedges, dtype=np.float64)
        self.eedges = None if eedges is None else np.asarray(eedges, dtype=np.float64)
        self._s = s

    @property
    def ptr(self):
        return C.cast(C.pointer(self._s), C.c_void_p)


def find_bin(nbins, lo, hi, edges, x) -> int:
    """TAxis::FindBin for a non-extendable axis."""
    if edges is None:
        return lib().orc_find_bin(nbins, lo, hi, None, x)
    a, p = _d(edges)
    return lib().orc_find_bin(len(a) - 1, a[0], a[-1], p, x)


# ------------------------------------------------------------------------------------------------
# interpolator
# ------------------------------------------------------------------------------------------------
class InterpRangeException(Exception):
    """include/Interpolator.hpp:14-18"""


class Interp:
    def __init__(self, ecm, threshold, settings=None):
        ecm = np.ascontiguousarray(ecm, dtype=np.float64)
        n = len(ecm)
        if settings is None:
            settings = [(False, 0, n - 1)]      # src/Interpolator.cpp:356-360
        st = np.ascontiguousarray([[int(bool(a)), int(b), int(c)] for a, b, c in settings], dtype=np.int32)
        self.n, self.ecm, self.threshold = n, ecm, float(threshold)
        self.settings = [(bool(a), int(b), int(c)) for a, b, c in settings]
        self._h = lib().orc_interp_create(n, ecm.ctypes.data_as(_dp), float(threshold), len(st),
                                          st.ctypes.data_as(_ip))
        if not self._h:
            raise InterpRangeException("[!] Wrong interpolation ranges")
        self.ext = np.concatenate([[threshold], ecm])

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.orc_interp_free(self._h)
            self._h = None

    def ranges(self):
        out = []
        for r in range(lib().orc_interp_nranges(self._h)):
            t, b, s = C.c_int(), C.c_int(), C.c_int()
            mn, mx = C.c_double(), C.c_double()
            lib().orc_interp_range(self._h, r, C.byref(t), C.byref(b), C.byref(s), C.byref(mn), C.byref(mx))
            out.append((t.value, b.value, s.value, mn.value, mx.value))
        return out

    def spline_c(self, r, j):
        p = lib().orc_interp_spline_c(self._h, r, j)
        return np.ctypeslib.as_array(p, shape=(self.n + 1,)).copy()

    def basis_eval(self, j, e):
        return lib().orc_basis_eval(self._h, j, float(e))

    def basis_deriv_eval(self, j, e):
        return lib().orc_basis_deriv_eval(self._h, j, float(e))

    def eval(self, y, e):
        a, p = _d(y)
        return lib().orc_interp_eval(self._h, p, float(e))

    def kf_basis_integral(self, i, j, eff: Efficiency | None = None):
        return lib().orc_kf_basis_integral(self._h, i, j, eff.ptr if eff else None)

    def eval_rows(self, rows, eff: Efficiency | None = None, mode: str = "converged"):
        """Selected rows of the equation matrix (full-size spot checks): returns (G[rows, :], Err[rows, :])."""
        m = {"faithful": 0, "converged": 1}[mode]
        G = np.zeros((len(rows), self.n)); E = np.zeros((len(rows), self.n))
        for a, i in enumerate(rows):
            for j in range(self.n):
                e, r = C.c_double(), C.c_double()
                G[a, j] = lib().orc_kf_basis_integral_ex(self._h, int(i), j, eff.ptr if eff else None, m, C.byref(e), C.byref(r))
                E[a, j] = e.value
        return G, E

    def integral_basis(self):
        """_evalDotProductOperator -- src/ISRSolverSLE.cpp:165-171."""
        return np.array([lib().orc_integral_basis(self._h, j) for j in range(self.n)])

    def deriv_projector(self):
        """D[i, j] = basis_j'(E_i) -- src/ISRSolverTikhonov.cpp:92-99."""
        D = np.empty((self.n, self.n), order="F")
        lib().orc_deriv_projector(self._h, D.ctypes.data_as(_dp))
        return np.ascontiguousarray(D)

    def energy_spread_matrix(self, ecm_err):
        a, p = _d(ecm_err)
        S = np.empty((self.n, self.n), order="F")
        lib().orc_energy_spread_matrix(self._h, p, S.ctypes.data_as(_dp))
        return np.ascontiguousarray(S)

    def eval_eq_matrix(self, eff: Efficiency | None = None, ecm_err=None, nthreads: int = 0,
                       mode: str = "faithful", with_errors: bool = False):
        """ISRSolverSLE::evalEqMatrix -- src/ISRSolverSLE.cpp:138-149.  Returns G[i, j] (logical).

        mode="faithful": the reference's call structure (one adaptive call per range, split at x0 only,
        loosen-until-success retries).  mode="converged": same integrand and QUADPACK routines with the
        range additionally cut at spline knots / efficiency bin edges, so every call converges at 1e-12.
        with_errors=True also returns (sum of QUADPACK abserr, loosest epsrel used) per entry."""
        if nthreads <= 0:
            nthreads = os.cpu_count() or 1
        G = np.empty((self.n, self.n), order="F")
        Err = np.zeros((self.n, self.n), order="F")
        Rel = np.zeros((self.n, self.n), order="F")
        ep = None
        if ecm_err is not None:
            ea, ep = _d(ecm_err)
        lib().orc_eval_eq_matrix_ex(self._h, eff.ptr if eff else None, ep, G.ctypes.data_as(_dp), nthreads,
                                    {"faithful": 0, "converged": 1}[mode],
                                    Err.ctypes.data_as(_dp), Rel.ctypes.data_as(_dp))
        if with_errors:
            return np.ascontiguousarray(G), np.ascontiguousarray(Err), np.ascontiguousarray(Rel)
        return np.ascontiguousarray(G)


# ------------------------------------------------------------------------------------------------
# synthetic Breit-Wigner workload
# ------------------------------------------------------------------------------------------------
BW_M, BW_G, BW_S0 = 1.50, 0.25, 4.0
E_MIN, E_MAX, E_THR = 1.18, 2.00, 0.827
TOY_SEED = 20211103


def bw_born(e, thr=E_THR):
    return np.array([lib().orc_bw_born(float(x), BW_M, BW_G, BW_S0, thr) for x in np.atleast_1d(e)])


def bw_visible(ecm, thr=E_THR, eff: Efficiency | None = None, nthreads: int = 0):
    """sigma_vis,i = int_0^{1-thr^2/E_i^2} F eps sigma_B(E_i sqrt(1-x)) dx, as
    src/isrsolver-KuraevFadin-convolution.cpp:122-134 does with the reference's adaptive quadrature."""
    a, p = _d(ecm)
    out = np.empty(len(a))
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    lib().orc_bw_visible(len(a), p, thr, BW_M, BW_G, BW_S0, eff.ptr if eff else None,
                         out.ctypes.data_as(_dp), nthreads)
    return out


def draw_toys(vcs, vcs_err, ntoys, seed=TOY_SEED):
    """Host-generated toy visible cross sections, toy-major T x N (stands in for
    randomDrawVisCS/gRandom->Gaus, src/Utils.cpp:6-13, which is unseeded and ROOT-only)."""
    z = np.random.Generator(np.random.PCG64(seed)).standard_normal((ntoys, len(vcs)))
    return np.asarray(vcs)[None, :] + np.asarray(vcs_err)[None, :] * z


def row_table_efficiency(ecm, nbins=512):
    """Config C3 efficiency: per-row table of `nbins` uniform x-bins on [0,1)."""
    xc = (np.arange(nbins) + 0.5) / nbins
    vals = np.zeros((len(ecm), nbins + 2))
    for i, e in enumerate(ecm):
        vals[i, 1:-1] = 0.25 + 0.5 * np.exp(-40.0 * xc) * (1.0 - 0.1 * (e - 1.5))
    return Efficiency("row_table", values=vals, nx=nbins, xlo=0.0, xhi=1.0)


# ------------------------------------------------------------------------------------------------
# ISRSolverSLE::solve and the toy loops
# ------------------------------------------------------------------------------------------------
def sle_solve(G, vcs, vcs_err):
    """src/ISRSolverSLE.cpp:86-95: bcs = COD(G).solve(vcs); Cov = (G^T W G)^-1, W = diag(err^-2)
    (src/BaseISRSolver.cpp:355-357)."""
    n = G.shape[0]
    bcs = sla.lstsq(G, vcs, cond=n * np.finfo(float).eps, lapack_driver="gelsy")[0]
    W = np.diag(np.asarray(vcs_err, dtype=float) ** -2.0)
    cov = np.linalg.inv(G.T @ W @ G)
    return bcs, cov


def chi2_test_loop(G, toys, bcs0, vcs_err):
    """src/Chi2Test.cpp:37-60 literally: LU of Cov once, then per toy a full solve() (COD + G^T W G +
    inverse -- O(N^3) each) and chi2 = d . Cov^-1 d.  Returns (chi2[T], bcs[T, N])."""
    _, cov = sle_solve(G, toys[0] * 0 + 1.0, vcs_err)   # Cov does not depend on vcs
    lu = sla.lu_factor(cov)
    chi2 = np.empty(len(toys))
    bs = np.empty_like(toys)
    for k, v in enumerate(toys):
        b, _cov_k = sle_solve(G, v, vcs_err)
        d = b - bcs0
        chi2[k] = d @ sla.lu_solve(lu, d)
        bs[k] = b
    return chi2, bs


def chi2_test_fair(G, toys, bcs0, vcs_err):
    """Same numbers with the redundancy removed (factor once, one GEMM over all toys): the
    'algorithmically fair' CPU baseline of BASELINE.md 3 and the expected values for large T."""
    Ginv = np.linalg.inv(G)
    B = toys @ Ginv.T
    D = B - bcs0[None, :]
    R = G / np.asarray(vcs_err)[:, None]        # W^(1/2) G ; Cov^-1 = R^T R
    Q = D @ R.T
    return np.einsum("ij,ij->i", Q, Q), B


def th1_fill(values, nbins, lo, hi):
    """TH1F::Fill over `values`: returns (counts[nbins+2] float32 incl. under/overflow, n_in, sum_x, sum_x2)
    with ROOT's bin arithmetic; statistics accumulate in-range entries only."""
    counts = np.zeros(nbins + 2, dtype=np.float32)
    v = np.asarray(values, dtype=np.float64)
    idx = np.empty(len(v), dtype=np.int64)
    below, above = v < lo, ~(v < hi)
    mid = ~(below | above)
    idx[below] = 0
    idx[above] = nbins + 1
    idx[mid] = 1 + (nbins * (v[mid] - lo) / (hi - lo)).astype(np.int64)
    np.add.at(counts, idx, 1.0)
    vin = v[mid]
    return counts, idx, len(vin), vin.sum(), (vin * vin).sum()


def chi2_histogram(chi2):
    """src/Chi2Test.cpp:64-78: TH1F(128, min, max), every value filled (max lands in overflow)."""
    lo, hi = float(np.min(chi2)), float(np.max(chi2))
    counts, idx, *_ = th1_fill(chi2, 128, lo, hi)
    return counts, idx, lo, hi


def ratio_test(B, bcs0):
    """src/RatioTest.cpp:33-54: per point TH1F(1024,-2,2) of (b-b0)/b0; GetMean / GetMeanError."""
    n = B.shape[1]
    means, errs = np.empty(n), np.empty(n)
    counts = np.zeros((n, 1026), dtype=np.float32)
    stats = np.zeros((n, 3))
    for j in range(n):
        r = (B[:, j] - bcs0[j]) / bcs0[j]
        c, _, cnt, sx, sx2 = th1_fill(r, 1024, -2.0, 2.0)
        counts[j] = c
        stats[j] = (cnt, sx, sx2)
        mean = sx / cnt if cnt else 0.0
        rms = np.sqrt(abs(sx2 / cnt - mean * mean)) if cnt else 0.0
        means[j], errs[j] = mean, (rms / np.sqrt(cnt) if cnt else 0.0)
    return means, errs, counts, stats


# ------------------------------------------------------------------------------------------------
# Tikhonov (src/ISRSolverTikhonov.cpp) and TSVD (src/ISRSolverTSVD.cpp)
# ------------------------------------------------------------------------------------------------
def tikhonov_F(D, w, deriv_reg=True):
    """_evalProblemMatrices :139-148: F = D^T diag(w) D, or diag(w)."""
    return D.T @ (D * w[:, None]) if deriv_reg else np.diag(w)


def tikhonov_solve(G, vcs, vcs_err, F, lam):
    """solve() :61-72 with _evalProblemMatrices :139-155 (T, L, two LUs, A+ = T^-1 G^T W)."""
    n = G.shape[0]
    W = np.diag(np.asarray(vcs_err, dtype=float) ** -2.0)
    Winv = np.linalg.inv(W)
    mAt = G.T
    mT = mAt @ W @ G + lam * F
    mL = np.linalg.inv(F) @ mAt @ W @ G + lam * np.eye(n)
    mAp = np.linalg.solve(mT, mAt @ W)
    bcs = mAp @ vcs
    cov = mAp @ Winv @ mAp.T
    return dict(bcs=bcs, cov=cov, T=mT, L=mL, Ap=mAp, W=W)


def tikhonov_lcurve(G, vcs, vcs_err, F, lam):
    """evalEqNorm2 :106-109, evalSmoothnessConstraintNorm2 :111-113, evalLCurveCurvature :115-119,
    evalLCurveCurvatureDerivative :121-128."""
    s = tikhonov_solve(G, vcs, vcs_err, F, lam)
    b, L, W = s["bcs"], s["L"], s["W"]
    dv = G @ b - vcs
    x = dv @ (W @ dv)
    y = b @ (F @ b)
    ds = -np.linalg.solve(L, b)
    d2s = -2.0 * np.linalg.solve(L, ds)
    dksi = 2.0 * b @ (F @ ds)
    d2ksi = 2.0 * ds @ (F @ ds) + 2.0 * b @ (F @ d2s)
    curv = -abs(1.0 / dksi / (1.0 + lam * lam) ** 1.5)
    dcurv = -d2ksi * dksi ** -2.0 * (1.0 + lam * lam) ** -1.5 + -3.0 * lam / dksi * (1.0 + lam * lam) ** -2.5
    return dict(x=x, y=y, curv=curv, dcurv=dcurv, **s)


def tsvd_solve(G, vcs, vcs_err, k, keep_one=False):
    """src/ISRSolverTSVD.cpp:53-74."""
    U, S, Vt = np.linalg.svd(G)
    f, n = (k - 1, 1) if keep_one else (0, k)
    K = U[:, f:f + n] @ np.diag(S[f:f + n]) @ Vt[f:f + n, :]
    N = G.shape[0]
    bcs = sla.lstsq(K, vcs, cond=N * np.finfo(float).eps, lapack_driver="gelsy")[0]
    cov = None
    if not keep_one and k == N:
        W = np.diag(np.asarray(vcs_err, dtype=float) ** -2.0)
        cov = np.linalg.inv(K.T @ W @ K)
    return bcs, cov, (U, S, Vt)


def iterative_solve(G, vcs, vcs_err, n_iter=10):
    """IterISRInterpSolver::solve -- src/IterISRInterpSolver.cpp:45-57: b <- vcs; repeat: radcorr = (G b) / b, b = vcs / radcorr,
    Cov = diag((vcs_err / radcorr)^2).  Returns (bcs, radcorr, cov)."""
    vcs = np.asarray(vcs, dtype=float)
    b, rad = vcs.copy(), np.zeros(len(vcs))
    cov = np.zeros((len(vcs), len(vcs)))
    for _ in range(int(n_iter)):
        rad = (G @ b) / b
        b = vcs / rad
        cov = np.diag((np.asarray(vcs_err, dtype=float) / rad) ** 2)
    return b, rad, cov


_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def convolution_kf(energy, fcn, min_x, max_x):
    """convolutionKuraevFadin (src/KuraevFadin.cpp:56-73) of a Python callable of the energy, efficiency = 1."""
    L = lib()
    L.orc_convolution_kf.restype = C.c_double
    L.orc_convolution_kf.argtypes = [C.c_double, _FN, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int]
    cb = _FN(lambda e, ctx: float(fcn(e)))
    return L.orc_convolution_kf(float(energy), cb, None, float(min_x), float(max_x), None, 0)


def gaussian_conv(energy, sigma2, fcn):
    """gaussian_conv (src/Integration.cpp:86-100): 6-point Gauss-Hermite average of fcn around `energy`."""
    L = lib()
    L.orc_gaussian_conv.restype = C.c_double
    L.orc_gaussian_conv.argtypes = [C.c_double, C.c_double, _FN, C.c_void_p]
    cb = _FN(lambda e, ctx: float(fcn(e)))
    return L.orc_gaussian_conv(float(energy), float(sigma2), cb, None)


def vcs_fit_model(ecm, threshold, born, ecm_err=None):
    """The model side of ISRSolverVCSFitFunction (src/ISRSolverVCSFitter.cpp:64-88, 122-150): sigma_vis of the Born
    function `born(E)` at every energy -- 0 at or below threshold, convolutionKuraevFadin over [0, 1 - thr^2 / E^2], and
    its Gaussian average over the beam-energy spread when `ecm_err` is given."""
    def vis(e):
        if e * e <= threshold * threshold:
            return 0.0
        return convolution_kf(e, born, 0.0, 1.0 - threshold * threshold / (e * e))
    if ecm_err is None:
        return np.array([vis(float(e)) for e in ecm])
    return np.array([gaussian_conv(float(e), float(s) ** 2, vis) for e, s in zip(ecm, ecm_err)])


def vcs_fit_chi2(ecm, vcs, vcs_err, threshold, born, ecm_err=None):
    """ISRSolverVCSFitFunction::operator() (src/ISRSolverVCSFitter.cpp:64-93): sum of ((model - vcs) / vcs_err)^2 in point
    order."""
    m = vcs_fit_model(ecm, threshold, born, ecm_err)
    chi2 = 0.0
    for mi, vi, ei in zip(m, vcs, vcs_err):
        d = mi - vi
        chi2 += d * d / ei / ei
    return chi2


def cond_number(G):
    """evalConditionNumber -- src/ISRSolverSLE.cpp:200-204."""
    s = np.linalg.svd(G, compute_uv=False)
    return s[0] / s[-1]


def sort_points(energy, vcs, energy_err, vcs_err):
    """BaseISRSolver ctor -- src/BaseISRSolver.cpp:105-128: permutation that sorts by energy."""
    idx = np.argsort(np.asarray(energy), kind="stable")
    f = lambda a: None if a is None else np.asarray(a, dtype=float)[idx]
    return idx, f(energy), f(vcs), f(energy_err), f(vcs_err)
