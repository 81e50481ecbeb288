"""ctypes access to oracle/_ref/libisr_ref.so -- the REFERENCE'S OWN assembly-path sources, compiled here.

ORACLE / TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

libisr_ref.so holds, compiled unmodified from /root/reference/src by oracle/Makefile:
    KuraevFadin.cpp, Integration.cpp, {Base,Linear,CSpline}RangeInterpolator{,2}.cpp, Interpolator{,2}.cpp
linked against oracle/ref_shim (the GSL entry points they call, implemented on the QUADPACK / cspline
restatement of oracle/isr_oracle.c; Eigen::VectorXd as a plain vector) and oracle/ref_glue.cpp (extern "C"
access + the three ROOT/Eigen-bound snippets that cannot be compiled, restated with their line numbers).
It is built in the container that has /root/reference (`make -C oracle`), is git-ignored, and travels to
the GPU box as a prebuilt file; `available()` is False where it is missing, and the tests that need it skip.

This is what pins oracle/isr_oracle.c: the same inputs through the restatement and through the reference's
compiled code must agree (tests/test_ref_pin.py), and tests/golden/*.npz carry its outputs (`G_ref`).
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
