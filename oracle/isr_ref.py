"""ctypes access to oracle/_ref/libisr_ref.so -- the REFERENCE'S OWN sources for this path, compiled here.

ORACLE / TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

libisr_ref.so holds, compiled unmodified from /root/reference/src by oracle/Makefile:
    KuraevFadin.cpp, Integration.cpp, {Base,Linear,CSpline}RangeInterpolator{,2}.cpp, Interpolator{,2}.cpp   (assembly)
    BaseISRSolver{,2}.cpp, ISRSolverSLE{,2}.cpp, ISRSolverTikhonov.cpp, ISRSolverTSVD.cpp, IterISRInterpSolver.cpp,
    Chi2Test.cpp, RatioTest.cpp, Utils.cpp
linked against oracle/ref_shim (the GSL entry points they call, implemented on the QUADPACK / cspline
restatement of oracle/isr_oracle.c; a dense Eigen stand-in with textbook decompositions; ROOT containers over an
in-memory registry; Boost / NLopt stand-ins) and oracle/ref_glue.cpp + oracle/ref_solver_glue.cpp (extern "C" access).
It is built in the container that has /root/reference (`make -C oracle`), is git-ignored, and travels to
the GPU box as a prebuilt file; `available()` is False where it is missing, and the tests that need it skip.

This is what pins oracle/isr_oracle.c and the linear-algebra layer of oracle/isr_oracle.py: the same inputs through the
restatement and through the reference's compiled code must agree (tests/test_ref_pin.py), and tests/golden/*.npz carry
its outputs (`G_ref`; `refrun_*.npz` for the solver classes and the toy tests).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import isr_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libisr_ref.so")
REFERENCE = "/root/reference"
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
CALLBACK = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
EFF_CALLBACK = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_void_p)
FIT_CALLBACK = C.CFUNCTYPE(C.c_double, C.c_double, _dp, C.c_int, C.c_void_p)


def build() -> bool:
    """Build oracle/_ref/libisr_ref.so when the reference tree is present; returns availability."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-C", HERE, "-s", "_ref/libisr_ref.so"])
    return os.path.exists(SO)


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libisr_ref.so is missing (built only where /root/reference exists)")
        L = C.CDLL(SO)
        L.ref_kernel_kf.restype = C.c_double
        L.ref_kernel_kf.argtypes = [C.c_double, C.c_double]
        L.ref_convolution_kf.restype = C.c_double
        L.ref_convolution_kf.argtypes = [C.c_double, CALLBACK, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        for nm in ("ref_integrate", "ref_integrateS"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [CALLBACK, C.c_void_p, C.c_double, C.c_double, _dp]
        L.ref_gaussian_conv.restype = C.c_double
        L.ref_gaussian_conv.argtypes = [C.c_double, C.c_double, CALLBACK, C.c_void_p]
        L.ref_eval_eq_matrix_cb.argtypes = [C.c_void_p, EFF_CALLBACK, C.c_void_p, _dp, C.c_int, _ip]
        for pre in ("ref_", "ref2_"):
            f = lambda nm: getattr(L, pre + nm)
            f("interp_create").restype = C.c_void_p
            f("interp_create").argtypes = [C.c_int, _dp, C.c_double, C.c_int, _ip]
            f("interp_free").argtypes = [C.c_void_p]
            f("kf_basis_integral").restype = C.c_double
            f("kf_basis_integral").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            for nm in ("basis_eval", "basis_deriv_eval"):
                f(nm).restype = C.c_double
                f(nm).argtypes = [C.c_void_p, C.c_int, C.c_double]
            f("integral_basis").restype = C.c_double
            f("integral_basis").argtypes = [C.c_void_p, C.c_int]
            for nm in ("interp_eval", "interp_deriv_eval"):
                f(nm).restype = C.c_double
                f(nm).argtypes = [C.c_void_p, _dp, C.c_double]
            for nm in ("min_energy", "max_energy"):
                f(nm).restype = C.c_double
                f(nm).argtypes = [C.c_void_p]
            f("eval_eq_matrix").argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_int, _ip, C.c_int]
            f("energy_spread_matrix").argtypes = [C.c_void_p, _dp, _dp]
        # the solver classes and the toy-test drivers (oracle/ref_solver_glue.cpp)
        vp, fp = C.c_void_p, C.POINTER(C.c_float)
        L.refs_last_error.restype = C.c_char_p
        L.refs_create.restype = vp
        L.refs_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_double, vp, C.c_double, C.c_int]
        L.refs_free.argtypes = [vp]
        L.refs_set_ranges.argtypes = [vp, C.c_int, _ip]
        L.refs_inject_matrix.argtypes = [vp, _dp]
        L.refs_eval_eq_matrix.argtypes = [vp]
        L.refs_solve.argtypes = [vp]
        L.refs_reset_vcs.argtypes = [vp, _dp]
        L.refs_reset_ecm_err.argtypes = [vp, _dp, C.c_int]
        for nm in ("refs_get_points", "refs_get_bcs", "refs_get_cov", "refs_get_matrix", "refs_tik_functionals"):
            getattr(L, nm).argtypes = [vp, _dp]
        L.refs_cond.restype = C.c_double
        L.refs_cond.argtypes = [vp]
        L.refs_interp_eval.restype = C.c_double
        L.refs_interp_eval.argtypes = [vp, _dp, C.c_double]
        L.refs_tik_set.argtypes = [vp, C.c_double, C.c_int]
        L.refs_lambda_objective.restype = C.c_double
        L.refs_lambda_objective.argtypes = [vp, C.c_double, _dp]
        L.refs_tik_operators.argtypes = [vp, _dp, _dp]
        L.refs_tsvd_set.argtypes = [vp, C.c_int, C.c_int]
        L.refs_iter_set.argtypes = [vp, C.c_int]
        L.refs_iter_radcorr.argtypes = [vp, _dp]
        L.refs_chi2_test_model.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp, _dp, fp]
        L.refs_chi2_test_data.argtypes = [vp, C.c_int, _dp, _dp, _dp, fp]
        L.refs_ratio_test_model.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp, _dp, fp, _dp]
        L.refs_save.argtypes = [vp, _dp, _dp, _dp, _dp, _dp]
        L.reff_last_error.restype = C.c_char_p
        L.reff_create.restype = vp
        L.reff_create.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, _dp, FIT_CALLBACK, vp]
        L.reff_free.argtypes = [vp]
        L.reff_set_spread.argtypes = [vp, C.c_int]
        L.reff_chi2.argtypes = [vp, _dp, C.c_int, _dp]
        L.reff_save.argtypes = [vp, _dp, C.c_int, C.c_int, _dp]
        L.refs2_last_error.restype = C.c_char_p
        L.refs2_create.restype = vp
        L.refs2_create.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, vp]
        L.refs2_free.argtypes = [vp]
        L.refs2_set_ranges.argtypes = [vp, C.c_int, _ip]
        L.refs2_eval_eq_matrix.argtypes = [vp]
        L.refs2_solve.argtypes = [vp]
        for nm in ("refs2_get_bcs", "refs2_get_cov", "refs2_get_matrix"):
            getattr(L, nm).argtypes = [vp, _dp]
        L.refs2_cond.restype = C.c_double
        L.refs2_cond.argtypes = [vp]
        _LIB = L
    return _LIB


def kernel_kf(x: float, s: float) -> float:
    """The reference's compiled kernelKuraevFadin (src/KuraevFadin.cpp:14-49)."""
    return lib().ref_kernel_kf(float(x), float(s))


def convolution_kf(energy, fcn, min_x, max_x, eff: O.Efficiency | None = None) -> float:
    """The reference's compiled convolutionKuraevFadin (src/KuraevFadin.cpp:56-73); fcn(E) is a Python callable."""
    cb = CALLBACK(lambda e, _ctx: float(fcn(e)))
    return lib().ref_convolution_kf(float(energy), cb, None, float(min_x), float(max_x), eff.ptr if eff else None)


def integrate(fcn, a, b, singular=False):
    """integrate / integrateS (src/Integration.cpp:19-81): returns (result, error)."""
    cb = CALLBACK(lambda x, _ctx: float(fcn(x)))
    err = C.c_double()
    f = lib().ref_integrateS if singular else lib().ref_integrate
    return f(cb, None, float(a), float(b), C.byref(err)), err.value


def gaussian_conv(energy, sigma2, fcn) -> float:
    cb = CALLBACK(lambda x, _ctx: float(fcn(x)))
    return lib().ref_gaussian_conv(float(energy), float(sigma2), cb, None)


class RefInterp:
    """The reference's Interpolator (v2=False; efficiency(x, E)) or Interpolator2 (v2=True;
    efficiency(x, energyIndex)) object, constructed by the reference's own constructor."""

    def __init__(self, ecm, threshold, settings=None, v2: bool = False):
        self._pre = "ref2_" if v2 else "ref_"
        ecm = np.ascontiguousarray(ecm, dtype=np.float64)
        self.n, self.ecm = len(ecm), ecm
        if settings is None:
            st, nr = None, 0
        else:
            sa = np.ascontiguousarray([[int(bool(a)), int(b), int(c)] for a, b, c in settings], dtype=np.int32)
            st, nr = sa.ctypes.data_as(_ip), len(sa)
            self._keep = sa
        self._h = self._f("interp_create")(self.n, ecm.ctypes.data_as(_dp), float(threshold), nr, st)
        if not self._h:
            raise O.InterpRangeException("[!] Wrong interpolation range.")

    def _f(self, name):
        return getattr(lib(), self._pre + name)

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            self._f("interp_free")(self._h)
            self._h = None

    def kf_basis_integral(self, i, j, eff: O.Efficiency | None = None):
        return self._f("kf_basis_integral")(self._h, int(i), int(j), eff.ptr if eff else None)

    def basis_eval(self, j, e): return self._f("basis_eval")(self._h, int(j), float(e))
    def basis_deriv_eval(self, j, e): return self._f("basis_deriv_eval")(self._h, int(j), float(e))
    def integral_basis(self): return np.array([self._f("integral_basis")(self._h, j) for j in range(self.n)])
    def min_energy(self): return self._f("min_energy")(self._h)
    def max_energy(self): return self._f("max_energy")(self._h)

    def eval(self, y, e):
        a = np.ascontiguousarray(y, dtype=np.float64)
        return self._f("interp_eval")(self._h, a.ctypes.data_as(_dp), float(e))

    def deriv_eval(self, y, e):
        a = np.ascontiguousarray(y, dtype=np.float64)
        return self._f("interp_deriv_eval")(self._h, a.ctypes.data_as(_dp), float(e))

    def eval_eq_matrix(self, eff: O.Efficiency | None = None, ecm_err=None, rows=None, nthreads: int = 0):
        """evalEqMatrix (src/ISRSolverSLE.cpp:138-149): the fill loop over the reference's own
        evalKuraevFadinBasisIntegral, then the optional spread product.  Logical G[i, j]; with `rows`
        only those rows are filled (the others are zero)."""
        if nthreads <= 0:
            nthreads = os.cpu_count() or 1
        G = np.zeros((self.n, self.n), order="F")
        if rows is None:
            self._f("eval_eq_matrix")(self._h, eff.ptr if eff else None, G.ctypes.data_as(_dp), 0, None, nthreads)
        else:
            ra = np.ascontiguousarray(rows, dtype=np.int32)
            self._f("eval_eq_matrix")(self._h, eff.ptr if eff else None, G.ctypes.data_as(_dp), len(ra),
                                      ra.ctypes.data_as(_ip), nthreads)
        G = np.ascontiguousarray(G)
        if ecm_err is not None:
            G = self.energy_spread_matrix(ecm_err) @ G
        return G

    def eval_eq_matrix_callable(self, eff, rows=None):
        """evalEqMatrix fill loop with the efficiency as a Python callable eps(x, energy), as PyISR passes it
        (include/PyISRSolver.hpp:189-197).  v1 interpolator only; single-threaded."""
        assert self._pre == "ref_"
        cb = EFF_CALLBACK(lambda x, e, _ctx: float(eff(x, e)))
        G = np.zeros((self.n, self.n), order="F")
        ra = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        lib().ref_eval_eq_matrix_cb(self._h, cb, None, G.ctypes.data_as(_dp), 0 if ra is None else len(ra),
                                    None if ra is None else ra.ctypes.data_as(_ip))
        return np.ascontiguousarray(G)

    def energy_spread_matrix(self, ecm_err):
        a = np.ascontiguousarray(ecm_err, dtype=np.float64)
        S = np.empty((self.n, self.n), order="F")
        self._f("energy_spread_matrix")(self._h, a.ctypes.data_as(_dp), S.ctypes.data_as(_dp))
        return np.ascontiguousarray(S)


# ------------------------------------------------------------------------------------------------
# the reference's solver classes and toy-test drivers, compiled (oracle/ref_solver_glue.cpp)
# ------------------------------------------------------------------------------------------------
def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefSolver:
    """One of the reference's compiled ISRSolverSLE / ISRSolverTikhonov / ISRSolverTSVD objects.

    kind: 'sle' | 'tikhonov' | 'tsvd' | 'iterative' (IterISRInterpSolver).  ctor 'pointers' = the constructor PyISR uses (src/BaseISRSolver.cpp:78-128),
    'graph' = the TGraphErrors (+ TEfficiency) constructor of the executables (:135-165).  Matrices cross this boundary
    row-major [i][j] like the rest of the oracle (the library is column-major inside)."""

    KINDS = {"sle": 0, "tikhonov": 1, "tsvd": 2, "iterative": 3}

    def __init__(self, kind, ecm, vcs, vcs_err, threshold, ecm_err=None, eff: O.Efficiency | None = None, settings=None,
                 lam=1.0, upper=None, ctor="graph"):
        self._L = lib()
        self.n = len(ecm)
        self.kind = kind
        self._eff = eff
        e, v, ve = _c(ecm), _c(vcs), _c(vcs_err)
        ee = None if ecm_err is None else _c(ecm_err)
        if eff is not None and eff.kind not in ("one", "table_2d", "e_only"):
            raise ValueError("the v1 solver classes take a TEfficiency (2-D or E-only) or none")
        if eff is not None and eff.kind != "one" and ctor != "graph":
            raise ValueError("a TEfficiency goes with the graph constructor")
        self._h = self._L.refs_create(self.KINDS[kind], 0 if ctor == "pointers" else 1, self.n, _p(e), _p(v),
                                      None if ee is None else _p(ee), _p(ve), float(threshold),
                                      eff.ptr if (eff is not None and eff.kind != "one") else None, float(lam),
                                      int(self.n if upper is None else upper))
        if not self._h:
            raise RuntimeError("reference constructor failed: " + self._L.refs_last_error().decode())
        if settings is not None:
            st = np.ascontiguousarray(np.asarray(settings, dtype=np.int32).reshape(-1, 3))
            if self._L.refs_set_ranges(self._h, len(st), st.ctypes.data_as(_ip)):
                raise O.InterpRangeException(self._L.refs_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.refs_free(self._h)
            self._h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError("reference call failed: " + self._L.refs_last_error().decode())

    def inject_matrix(self, G):
        """Skip the N^2 adaptive integrations of evalEqMatrix: hand the solver a precomputed operator matrix."""
        self._check(self._L.refs_inject_matrix(self._h, _p(_c(np.asarray(G).T))))

    def eval_eq_matrix(self):
        self._check(self._L.refs_eval_eq_matrix(self._h))
        return self.matrix()

    def solve(self):
        self._check(self._L.refs_solve(self._h))
        return self.bcs(), self.cov()

    def cond_spread_scan(self, sigmas):
        """The loop of src/isrsolver-condnum-test.cpp:84-128: per spread setting resetECMErrors(sigma * ones), evalEqMatrix,
        JacobiSVD -> sigma_max / sigma_min (sigma = 0: the spread switched off)."""
        out = []
        for sg in sigmas:
            self._L.refs_reset_ecm_err(self._h, _p(_c(np.full(self.n, float(sg)))), int(sg > 0))
            self._check(self._L.refs_eval_eq_matrix(self._h))
            out.append(self.cond())
        return np.array(out)

    def reset_vcs(self, v):
        self._L.refs_reset_vcs(self._h, _p(_c(v)))

    def points(self):
        out = np.empty((4, self.n))
        self._L.refs_get_points(self._h, _p(out))
        return dict(ecm=out[0], ecm_err=out[1], vcs=out[2], vcs_err=out[3])

    def bcs(self):
        out = np.empty(self.n)
        self._L.refs_get_bcs(self._h, _p(out))
        return out

    def cov(self):
        out = np.empty((self.n, self.n))
        self._L.refs_get_cov(self._h, _p(out))
        return out.T.copy()

    def matrix(self):
        out = np.empty((self.n, self.n))
        self._L.refs_get_matrix(self._h, _p(out))
        return out.T.copy()

    def cond(self):
        return self._L.refs_cond(self._h)

    def interp_eval(self, y, e):
        return self._L.refs_interp_eval(self._h, _p(_c(y)), float(e))

    # ---- Tikhonov
    def tik_set(self, lam, deriv_reg=True):
        self._L.refs_tik_set(self._h, float(lam), int(bool(deriv_reg)))

    def tik_functionals(self):
        """evalEqNorm2, evalSmoothnessConstraintNorm2, evalLCurveCurvature, evalLCurveCurvatureDerivative (after solve)."""
        out = np.empty(4)
        self._L.refs_tik_functionals(self._h, _p(out))
        return dict(x=out[0], y=out[1], curv=out[2], dcurv=out[3])

    def lambda_objective(self, lam):
        g = C.c_double(0.0)
        val = self._L.refs_lambda_objective(self._h, float(lam), C.byref(g))
        return val, g.value

    def tik_operators(self):
        D, w = np.empty((self.n, self.n)), np.empty(self.n)
        self._L.refs_tik_operators(self._h, _p(D), _p(w))
        return D.T.copy(), w

    # ---- IterISRInterpSolver
    def iter_set(self, n_iter):
        self._L.refs_iter_set(self._h, int(n_iter))

    def radcorr(self):
        out = np.empty(self.n)
        self._check(self._L.refs_iter_radcorr(self._h, _p(out)))
        return out

    # ---- TSVD
    def tsvd_set(self, upper, keep_one=False):
        self._L.refs_tsvd_set(self._h, int(upper), int(bool(keep_one)))

    # ---- the toy-test drivers; `toys` replace what gRandom->Gaus would draw
    def chi2_test_model(self, toys, model_vcs, model_bcs):
        toys = _c(toys)
        T = len(toys)
        chi2, lohi, counts = np.empty(T), np.empty(2), np.empty(130, dtype=np.float32)
        self._check(self._L.refs_chi2_test_model(self._h, T, _p(toys), _p(_c(model_vcs)), _p(_c(model_bcs)), _p(chi2), _p(lohi),
                                                 counts.ctypes.data_as(C.POINTER(C.c_float))))
        return chi2, counts, lohi[0], lohi[1]

    def chi2_test_data(self, toys):
        toys = _c(toys)
        T = len(toys) // 2
        chi2, lohi, counts = np.empty(T), np.empty(2), np.empty(130, dtype=np.float32)
        self._check(self._L.refs_chi2_test_data(self._h, T, _p(toys), _p(chi2), _p(lohi), counts.ctypes.data_as(C.POINTER(C.c_float))))
        return chi2, counts, lohi[0], lohi[1]

    def ratio_test_model(self, toys, model_vcs, model_bcs):
        toys = _c(toys)
        T, n = len(toys), self.n
        means, errs, counts, stats = np.empty(n), np.empty(n), np.empty((n, 1026), dtype=np.float32), np.empty((n, 3))
        self._check(self._L.refs_ratio_test_model(self._h, T, _p(toys), _p(_c(model_vcs)), _p(_c(model_bcs)), _p(means), _p(errs),
                                                  counts.ctypes.data_as(C.POINTER(C.c_float)), _p(stats)))
        return means, errs, counts, stats

    def save(self):
        """What save() writes (src/ISRSolverSLE.cpp:97-122): bcs, bcs_err, G, cov, inv_cov (row-major as written)."""
        n = self.n
        b, be = np.empty(n), np.empty(n)
        G, cv, ic = np.empty((n, n)), np.empty((n, n)), np.empty((n, n))
        self._check(self._L.refs_save(self._h, _p(b), _p(be), _p(G), _p(cv), _p(ic)))
        return dict(bcs=b, bcs_err=be, G=G, cov=cv, inv_cov=ic)


class RefSolver2:
    """The reference's compiled ISRSolverSLE2: one efficiency histogram in x per energy point (`eff` of kind 'row_table')."""

    def __init__(self, ecm, vcs, vcs_err, threshold, ecm_err=None, eff: O.Efficiency | None = None, settings=None):
        self._L = lib()
        self.n = len(ecm)
        self._eff = eff
        if eff is not None and eff.kind not in ("one", "row_table"):
            raise ValueError("ISRSolverSLE2 takes per-point histograms ('row_table') or none")
        ee = None if ecm_err is None else _c(ecm_err)
        self._h = self._L.refs2_create(self.n, _p(_c(ecm)), _p(_c(vcs)), None if ee is None else _p(ee), _p(_c(vcs_err)),
                                       float(threshold), eff.ptr if (eff is not None and eff.kind != "one") else None)
        if not self._h:
            raise RuntimeError("reference constructor failed: " + self._L.refs2_last_error().decode())
        if settings is not None:
            st = np.ascontiguousarray(np.asarray(settings, dtype=np.int32).reshape(-1, 3))
            if self._L.refs2_set_ranges(self._h, len(st), st.ctypes.data_as(_ip)):
                raise O.InterpRangeException(self._L.refs2_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.refs2_free(self._h)
            self._h = None

    def _mat(self, fn):
        out = np.empty((self.n, self.n))
        fn(self._h, _p(out))
        return out.T.copy()

    def eval_eq_matrix(self):
        if self._L.refs2_eval_eq_matrix(self._h):
            raise RuntimeError(self._L.refs2_last_error().decode())
        return self._mat(self._L.refs2_get_matrix)

    def solve(self):
        if self._L.refs2_solve(self._h):
            raise RuntimeError(self._L.refs2_last_error().decode())
        b = np.empty(self.n)
        self._L.refs2_get_bcs(self._h, _p(b))
        return b, self._mat(self._L.refs2_get_cov)

    def matrix(self):
        return self._mat(self._L.refs2_get_matrix)

    def cond(self):
        return self._L.refs2_cond(self._h)


class RefVCSFitFunction:
    """The reference's compiled ISRSolverVCSFitFunction (src/ISRSolverVCSFitter.cpp): chi2 of a parametrised Born cross
    section `fcn(energy, par)` against the visible data through N forward convolutions (blurred by the energy spread when
    `ecm_err` is given), and what saveResults writes at given parameters."""

    def __init__(self, ecm, vcs, vcs_err, threshold, fcn, ecm_err=None):
        self._L = lib()
        self.n = len(ecm)
        self._cb = FIT_CALLBACK(lambda en, par, npar, ctx: float(fcn(en, [par[i] for i in range(npar)])))
        ee = None if ecm_err is None else _c(ecm_err)
        self._h = self._L.reff_create(self.n, float(threshold), _p(_c(ecm)), _p(_c(vcs)), None if ee is None else _p(ee),
                                      _p(_c(vcs_err)), self._cb, None)
        if not self._h:
            raise RuntimeError("reference constructor failed: " + self._L.reff_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.reff_free(self._h)
            self._h = None

    def set_spread(self, on):
        self._L.reff_set_spread(self._h, int(bool(on)))

    def chi2(self, par):
        par = _c(par)
        out = C.c_double(0.0)
        if self._L.reff_chi2(self._h, _p(par), len(par), C.byref(out)):
            raise RuntimeError(self._L.reff_last_error().decode())
        return out.value

    def save_results(self, par):
        par = _c(par)
        out = np.empty((4, self.n))
        if self._L.reff_save(self._h, _p(par), len(par), self.n, _p(out)):
            raise RuntimeError(self._L.reff_last_error().decode())
        return dict(vcs_gauss_conv=out[0], rad_delta=out[1], bcs=out[2], bcs_err=out[3])
