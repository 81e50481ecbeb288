"""Host-side mirror of the reference's PyISR module for the hot path, on top of the C ABI.

Reference surface mirrored here (paths relative to the reference tree):
  module functions  kernelKuraevFadin                      src/PyISR.cpp:197-207, include/PyKuraevFadin.hpp:9-17
  ISRSolverSLE      ctor kwargs + methods table            include/PyISRSolver.hpp:156-213, include/PyISRSolverSLE.hpp:183-215
  ISRSolverSLE2     per-point efficiency histograms        src/BaseISRSolver2.cpp:151-160, 196-204, 231-236
  ISRSolverTikhonov reg_param, make_LCurve, solve_LCurve   include/PyISRSolverTikhonov.hpp:44-146
  ISRSolverTSVD     upper_TSVD_index, enable/disable_keepone include/PyISRSolverTSVD.hpp:30-81
Method names, argument names and meaning follow the reference.  Differences forced by the environment
(documented in INTEGRATION.md): ROOT files are replaced by .npz files holding arrays under the same object
names; a Python-callable efficiency is replaced by an `Efficiency` table (the GPU cannot call back into
the interpreter per integrand node as include/PyISRSolver.hpp:189-197 does); the toy generator is a
seeded NumPy PCG64 stream instead of the unseeded global gRandom (src/Utils.cpp:6-13).

All numerical work goes through libisr_b200.so; nothing here computes on the CPU beyond O(N)
bookkeeping, and there is no fallback when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Sequence

import numpy as np

from . import _lib as L
from ._lib import (EfficiencyDimensionException, InterpRangeException, IsrGpuError)  # noqa: F401 (re-export)

TOY_SEED = 20211103

_CTX = {}


_GROUPS = {}


def _group(ngpu: int):
    """One isr_group per (process, size): contexts, host threads and the communicator over the first `ngpu` GPUs."""
    if ngpu not in _GROUPS:
        g = C.c_void_p()
        L.check(L.load().isr_group_create(int(ngpu), None, C.byref(g)))
        _GROUPS[ngpu] = g
    return _GROUPS[ngpu]


def _ctx(device: int | None = None):
    """One isr_ctx per (process, device); device defaults to LOCAL_RANK (one process per GPU) or 0."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _CTX:
        h = C.c_void_p()
        L.check(L.load().isr_ctx_create(device, C.byref(h)))
        _CTX[device] = h
    return _CTX[device]


def _f64(a, n=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        raise ValueError(f"{name}: expected shape ({n},), got {a.shape}")
    return a


class DifferentModelCrossSectionSizes(ValueError):
    """include/Chi2Test.hpp:47-52"""


# ---------------------------------------------------------------------------------------------------
# module-level functions
# ---------------------------------------------------------------------------------------------------
def kernelKuraevFadin(x, s, device: int | None = None):
    """Kuraev-Fadin radiator F(x, s) evaluated on the GPU (kernelKuraevFadin, src/KuraevFadin.cpp:14-49).
    Scalars or arrays (broadcast)."""
    xa, sa = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(s, dtype=np.float64))
    xa, sa = np.ascontiguousarray(xa).ravel(), np.ascontiguousarray(sa).ravel()
    out = np.empty_like(xa)
    L.check(L.load().isr_kernel_kf(_ctx(device), xa.size, L.dptr(xa), L.dptr(sa), L.dptr(out)))
    return float(out[0]) if np.ndim(x) == 0 and np.ndim(s) == 0 else out.reshape(np.broadcast(x, s).shape)


def _call_vectorised(fcn, arr):
    """Call a user function on an array; fall back to one call per element for scalar-only callables."""
    try:
        out = np.asarray(fcn(arr), dtype=np.float64)
        if out.shape == arr.shape:
            return np.ascontiguousarray(out)
    except Exception:
        pass
    return np.ascontiguousarray([float(fcn(float(v))) for v in arr], dtype=np.float64)


class ConvolutionPlan:
    """Forward Kuraev-Fadin convolution of an arbitrary Born cross section for many energies at once
    (convolutionKuraevFadin, src/KuraevFadin.cpp:56-73, looped over energies as
    src/isrsolver-KuraevFadin-convolution.cpp:122-134 does).  The quadrature lives on the GPU; the Born
    function is sampled on the host at the node energies the plan asks for.

        plan = ConvolutionPlan(energies, threshold=0.827)         # min_x = 0, max_x = 1 - (thr/E)^2
        vcs, err = plan.convolve(born)                            # adaptive, refines around narrow peaks
        vcs2, _ = plan.convolve(born2, refine=False)              # same grid, another function
    """

    def __init__(self, energy, threshold=None, min_x=None, max_x=None, efficiency=None, breaks=None,
                 device: int | None = None):
        e = np.ascontiguousarray(np.atleast_1d(energy), dtype=np.float64).ravel()
        self._m = e.size
        if max_x is None:
            if threshold is None:
                raise ValueError("give either threshold or max_x")
            max_x = 1.0 - (float(threshold) / e) ** 2
        mn = np.ascontiguousarray(np.broadcast_to(0.0 if min_x is None else min_x, e.shape), dtype=np.float64)
        mx = np.ascontiguousarray(np.broadcast_to(max_x, e.shape), dtype=np.float64)
        self._eff_call = None
        desc = None
        if efficiency is not None:
            if isinstance(efficiency, Efficiency):
                desc = efficiency.desc(self._m)
                self._keep_eff = efficiency
            elif callable(efficiency):
                self._eff_call = efficiency       # eps(x, E): sampled on the host at the plan's nodes
            else:
                raise TypeError("efficiency must be an Efficiency table or a callable eps(x, energy)")
        br = None if breaks is None else np.ascontiguousarray(np.atleast_1d(breaks), dtype=np.float64)
        self._lib = L.load()
        self._p = C.c_void_p()
        self._energy = e
        L.check(self._lib.isr_conv_plan_create(_ctx(device), e.size, L.dptr(e), L.dptr(mn), L.dptr(mx),
                                               C.byref(desc) if desc is not None else None,
                                               0 if br is None else br.size, L.dptr(br), C.byref(self._p)))

    def __del__(self):
        p = getattr(self, "_p", None)
        if p:
            try:
                self._lib.isr_conv_plan_destroy(p)
            except Exception:
                pass
            self._p = None

    def size(self):
        a, b = C.c_int64(), C.c_int64()
        L.check(self._lib.isr_conv_plan_size(self._p, C.byref(a), C.byref(b)))
        return a.value, b.value

    def nodes(self, want_weights=False):
        """(E', x, row[, weight]) of the pending panels' node slots."""
        n = self.size()[1]
        e, x, r = np.empty(n), np.empty(n), np.empty(n, dtype=np.int32)
        w = np.empty(n) if want_weights else None
        L.check(self._lib.isr_conv_plan_nodes(self._p, L.dptr(e), L.dptr(x), L.iptr(r), L.dptr(w)))
        return (e, x, r, w) if want_weights else (e, x, r)

    def _sample(self, fcn):
        e, x, r = self.nodes()
        f = _call_vectorised(fcn, e)
        if self._eff_call is not None:
            en = self._energy[r]
            try:
                ev = np.asarray(self._eff_call(x, en), dtype=np.float64)
                assert ev.shape == x.shape
            except Exception:
                ev = np.array([float(self._eff_call(float(a), float(b))) for a, b in zip(x, en)])
            f = f * ev
        return np.ascontiguousarray(f)

    def convolve(self, fcn, rel_tol=1e-10, abs_tol=1e-12, refine=True, max_levels=40):
        """Returns (result[M], error_estimate[M]).  refine=True bisects panels until every row meets
        max(abs_tol, rel_tol |result|); refine=False evaluates f once on the current panel set."""
        if self.size()[1] == 0:
            L.check(self._lib.isr_conv_plan_reset(self._p))
        for _ in range(max_levels):
            if self.size()[1]:
                L.check(self._lib.isr_conv_plan_feed(self._p, L.dptr(self._sample(fcn))))
            if not refine:
                break
            ns = C.c_int64()
            L.check(self._lib.isr_conv_plan_refine(self._p, float(rel_tol), float(abs_tol), C.byref(ns)))
            if ns.value == 0:
                break
        if self.size()[1]:      # max_levels exhausted with children pending: evaluate them, report the error
            L.check(self._lib.isr_conv_plan_feed(self._p, L.dptr(self._sample(fcn))))
        res, err = np.empty(self._m), np.empty(self._m)
        L.check(self._lib.isr_conv_plan_result(self._p, L.dptr(res), L.dptr(err)))
        return res, err

    def convolve_interp(self, solver, bcs=None):
        """Convolve the interpolated Born cross section of `solver` (its bcs() by default) entirely on the
        device: sigma_vis at the plan's energies, as ISRSolverVCSFitFunction / the radiative-correction
        tools evaluate it."""
        y = _f64(solver.bcs() if bcs is None else bcs, solver.n, "bcs")
        L.check(self._lib.isr_conv_plan_reset(self._p))
        L.check(self._lib.isr_conv_plan_feed_interp(self._p, solver._p, L.dptr(y)))
        res, err = np.empty(self._m), np.empty(self._m)
        L.check(self._lib.isr_conv_plan_result(self._p, L.dptr(res), L.dptr(err)))
        return res, err


def convolutionKuraevFadin(energy, fcn, min_x, max_x, efficiency=None, rel_tol=1e-10, abs_tol=1e-12, breaks=None,
                           device: int | None = None):
    """convolutionKuraevFadin(energy, fcn, min_x, max_x, efficiency) -- include/PyKuraevFadin.hpp:19-62,
    src/KuraevFadin.cpp:56-73.  `energy`, `min_x`, `max_x` may be arrays (one integral per entry); `fcn` is
    called with arrays of energies when it accepts them.  `efficiency` is an Efficiency table or a callable
    eps(x, energy) as in the reference."""
    scalar = np.ndim(energy) == 0
    plan = ConvolutionPlan(energy, min_x=min_x, max_x=max_x, efficiency=efficiency, breaks=breaks, device=device)
    res, _ = plan.convolve(fcn, rel_tol=rel_tol, abs_tol=abs_tol)
    return float(res[0]) if scalar else res.reshape(np.shape(energy))


def convolutionKuraevFadinBlurred(energy, fcn, threshold_energy, cm_energy_spread, efficiency=None, rel_tol=1e-10,
                                  abs_tol=1e-12, breaks=None, device: int | None = None):
    """convolutionKuraevFadinBlurred (include/PyKuraevFadin.hpp:64-116): the visible cross section
    convolved with a Gaussian beam-energy spread by the 6-point Gauss-Hermite rule of gaussian_conv
    (src/Integration.cpp:86-100)."""
    t = np.array([-2.350604973674492222833922, -1.335849074013696949714895, -4.360774119276165086792159e-1,
                  4.360774119276165086792159e-1, 1.335849074013696949714895, 2.350604973674492222833922])
    w = np.array([4.530009905508845640857473e-3, 1.570673203228566439163116e-1, 7.246295952243925240919147e-1,
                  7.246295952243925240919147e-1, 1.570673203228566439163116e-1, 4.530009905508845640857473e-3])
    scalar = np.ndim(energy) == 0
    e0 = np.atleast_1d(np.asarray(energy, dtype=np.float64)).ravel()
    sg = np.broadcast_to(np.asarray(cm_energy_spread, dtype=np.float64), e0.shape)
    en = (e0[:, None] + np.sqrt(2.0) * sg[:, None] * t[None, :]).ravel()
    live = en > threshold_energy            # below threshold max_x < 0: the reference integrates an empty range
    vis = np.zeros_like(en)
    if live.any():
        plan = ConvolutionPlan(en[live], threshold=threshold_energy, efficiency=efficiency, breaks=breaks, device=device)
        vis[live] = plan.convolve(fcn, rel_tol=rel_tol, abs_tol=abs_tol)[0]
    res = (vis.reshape(-1, 6) * (w / np.sqrt(np.pi))[None, :]).sum(1)
    return float(res[0]) if scalar else res.reshape(np.shape(energy))


# ---------------------------------------------------------------------------------------------------
# efficiency tables
# ---------------------------------------------------------------------------------------------------
class Efficiency:
    """Detection efficiency with ROOT histogram semantics (values include under/overflow cells).

    kind = 'one'        efficiency = 1                         (src/BaseISRSolver.cpp:140)
           'row_table'  one TH1D in x per energy point (v2)    (src/BaseISRSolver2.cpp:231-236)
           'table_2d'   TEfficiency 2-D in (x, E)              (src/BaseISRSolver.cpp:243-250)
           'e_only'     TEfficiency 1-D, a function of E only  (src/BaseISRSolver.cpp:236-241)
    """
    KINDS = {"one": 0, "row_table": 1, "table_2d": 2, "e_only": 3}

    def __init__(self, kind="one", values=None, nx=0, xlo=0.0, xhi=1.0, xedges=None,
                 ne=0, elo=0.0, ehi=1.0, eedges=None, dimension=None):
        if dimension is not None and dimension > 2:
            raise EfficiencyDimensionException("[!] Wrong efficiency dimension.")
        if kind not in self.KINDS:
            raise ValueError(f"unknown efficiency kind {kind!r}")
        self.kind = kind
        self.values = None if values is None else np.ascontiguousarray(values, dtype=np.float64)
        self.xedges = None if xedges is None else _f64(xedges)
        self.eedges = None if eedges is None else _f64(eedges)
        self.nx = int(nx) if self.xedges is None else len(self.xedges) - 1
        self.ne = int(ne) if self.eedges is None else len(self.eedges) - 1
        self.xlo, self.xhi = (float(xlo), float(xhi)) if self.xedges is None else (float(self.xedges[0]), float(self.xedges[-1]))
        self.elo, self.ehi = (float(elo), float(ehi)) if self.eedges is None else (float(self.eedges[0]), float(self.eedges[-1]))

    def desc(self, n_points: int) -> L.EpsDesc:
        d = L.EpsDesc()
        d.kind = self.KINDS[self.kind]
        d.nx, d.xlo, d.xhi, d.ne, d.elo, d.ehi = self.nx, self.xlo, self.xhi, self.ne, self.elo, self.ehi
        d.xedges, d.eedges, d.values = L.dptr(self.xedges), L.dptr(self.eedges), L.dptr(self.values)
        if self.kind == "row_table" and self.values.shape != (n_points, self.nx + 2):
            raise ValueError(f"row_table efficiency: values must be ({n_points}, {self.nx + 2})")
        if self.kind == "table_2d" and self.values.shape != (self.ne + 2, self.nx + 2):
            raise ValueError(f"table_2d efficiency: values must be ({self.ne + 2}, {self.nx + 2})")
        if self.kind == "e_only" and self.values.shape != (self.ne + 2,):
            raise ValueError(f"e_only efficiency: values must be ({self.ne + 2},)")
        return d


def pinned_empty(shape, device: int | None = None):
    """An uninitialised float64 array in page-locked host memory (isr_host_alloc): what isr_toy_batch copies from at
    the PCIe rate.  Freed when the array (and every view of it) is gone."""
    import weakref
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    count = int(np.prod(shape))
    if count == 0:
        return np.empty(shape)
    lib = L.load()
    ptr = C.c_void_p()
    L.check(lib.isr_host_alloc(_ctx(device), count * 8, C.byref(ptr)))
    buf = (C.c_double * count).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.float64).reshape(shape)
    weakref.finalize(buf, lib.isr_host_free, C.c_void_p(ptr.value))   # arr.base chain keeps buf alive
    return arr


def draw_toys(vcs, vcs_err, n: int, seed: int = TOY_SEED, pinned: bool = False):
    """n toy visible cross sections, toy-major (n, N): vcs + vcs_err * z, z ~ N(0,1) from PCG64(seed).
    Stands in for randomDrawVisCS (src/Utils.cpp:6-13).  pinned=True: the same values in page-locked memory."""
    gen = np.random.Generator(np.random.PCG64(seed))
    if not pinned:
        z = gen.standard_normal((int(n), len(vcs)))
        return np.asarray(vcs)[None, :] + np.asarray(vcs_err)[None, :] * z
    V = pinned_empty((int(n), len(vcs)))
    gen.standard_normal(out=V)
    V *= np.asarray(vcs_err)[None, :]
    V += np.asarray(vcs)[None, :]
    return V


def draw_toys_device(vcs, vcs_err, n: int, seed: int = TOY_SEED, first_toy: int = 0, device: int | None = None):
    """The toys a device-side campaign uses (isr_toy_campaign): toy k of stream `seed` is
    vcs + vcs_err * z(seed, k, .), z from Philox4x32-10 + Box-Muller (csrc/rng.cu).  Returned toy-major (n, N)."""
    vcs, vcs_err = _f64(vcs), _f64(vcs_err, len(vcs), "vcs_err")
    V = np.empty((int(n), len(vcs)))
    L.check(L.load().isr_draw_toys(_ctx(device), int(n), len(vcs), int(first_toy), int(seed), L.dptr(vcs), L.dptr(vcs_err), L.dptr(V)))
    return V


# ---------------------------------------------------------------------------------------------------
# solvers
# ---------------------------------------------------------------------------------------------------
class BaseISRSolver:
    """BaseISRSolver (include/BaseISRSolver.hpp:30-120; array constructor src/BaseISRSolver.cpp:77-129)."""

    def __init__(self, n, energy, vcs, energy_err, vcs_err, threshold, efficiency=None,
                 enableEnergySpread=False, device: int | None = None, efficiency_breaks=None):
        n = int(n)
        energy, vcs, vcs_err = _f64(energy, n, "energy"), _f64(vcs, n, "vcs"), _f64(vcs_err, n, "vcs_err")
        energy_err = np.zeros(n) if energy_err is None else _f64(energy_err, n, "energy_err")
        # sort by energy through an index permutation (src/BaseISRSolver.cpp:105-128)
        idx = np.argsort(energy, kind="stable")
        if np.any(np.diff(energy[idx]) <= 0):
            raise ValueError("duplicate centre-of-mass energies (the reference's unstable std::sort makes their order undefined)")
        self._perm = idx
        self._ecm, self._vcs = energy[idx].copy(), vcs[idx].copy()
        self._ecm_err, self._vcs_err = energy_err[idx].copy(), vcs_err[idx].copy()
        self._n, self._thr = n, float(threshold)
        self._spread = bool(enableEnergySpread)
        # efficiency: an Efficiency table (binned, resolved exactly at its bin edges) or, as in the reference
        # (include/PyISRSolver.hpp:189-197), a callable eps(x, energy) -- sampled by the host at the quadrature
        # nodes the GPU asks for (isr_assemble_nodes / isr_assemble_sampled); it must be smooth in x
        self._eff_call = None
        if efficiency is not None and not isinstance(efficiency, Efficiency):
            if not callable(efficiency):
                raise TypeError("efficiency must be an isrsolver_b200.Efficiency table or a callable eps(x, energy)")
            self._eff_call, efficiency = efficiency, None
            if efficiency_breaks is not None and len(efficiency_breaks):
                # x values where the callable jumps: handed to the quadrature as the edges of a table of ones
                br = np.unique(np.clip(np.asarray(efficiency_breaks, dtype=np.float64), 0.0, 1.0))
                edges = np.unique(np.concatenate([[0.0], br, [1.0]]))
                efficiency = Efficiency("row_table", values=np.ones((n, len(edges) + 1)), xedges=edges)
        self._eff = efficiency
        self._bcs = np.zeros(n)
        self._lib = L.load()
        self._ctxh = _ctx(device)
        self._p = C.c_void_p()
        self._settings = None
        desc = efficiency.desc(n) if efficiency is not None else None
        L.check(self._lib.isr_problem_create(self._ctxh, n, L.dptr(self._ecm), self._thr, 0, None,
                                             C.byref(desc) if desc is not None else None, C.byref(self._p)))

    def __del__(self):
        p = getattr(self, "_p", None)
        if p:
            try:
                self._lib.isr_problem_destroy(p)
            except Exception:
                pass
            self._p = None

    # getters (include/PyISRSolver.hpp:66-99 return views on solver-owned memory; so do these)
    n = property(lambda self: self._n)
    energy_spread_enabled = property(lambda self: self._spread)

    def getN(self): return self._n
    def ecm(self): return self._ecm
    def ecm_err(self): return self._ecm_err
    def vcs(self): return self._vcs
    def vcs_err(self): return self._vcs_err
    def bcs(self): return self._bcs
    def getThresholdEnergy(self): return self._thr
    def getMinEnergy(self): return float(self._ecm[0])
    def getMaxEnergy(self): return float(self._ecm[-1])
    def isEnergySpreadEnabled(self): return self._spread
    def enableEnergySpread(self): self._spread = True; self._invalidate()
    def disableEnergySpread(self): self._spread = False; self._invalidate()

    def resetVisibleCS(self, vcs): self._vcs = _f64(vcs, self._n, "vcs").copy()
    def resetVisibleCSErrors(self, vcs_err): self._vcs_err = _f64(vcs_err, self._n, "vcs_err").copy(); self._factored = False

    def reset_ecm_err(self, energy_err):
        """resetECMErrors (src/BaseISRSolver.cpp:448-450).  Unlike the reference (see its TO-DO at
        include/BaseISRSolver.hpp:85) the cached matrix is invalidated."""
        self._ecm_err = _f64(energy_err, self._n, "energy_err").copy()
        self._invalidate()
    resetECMErrors = reset_ecm_err

    def _invalidate(self):
        pass


class ISRSolverSLE(BaseISRSolver):
    """ISRSolverSLE -- the naive method (include/ISRSolverSLE.hpp; src/ISRSolverSLE.cpp)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._eq_ready = False      # _isEqMatrixPrepared
        self._factored = False
        self._G = None              # column-major storage like Eigen; intop_matrix() returns the transposed view
        self._cov = None

    def _invalidate(self):
        self._eq_ready = False
        self._factored = False

    # -- interpolation settings (src/ISRSolverSLE.cpp:125-136; JSON format README.md:177-188)
    def set_interp_settings(self, settings: Sequence[Sequence]):
        st = np.ascontiguousarray([[int(bool(a)), int(b), int(c)] for a, b, c in settings], dtype=np.int32)
        L.check(self._lib.isr_problem_set_ranges(self._p, len(st), L.iptr(st)))
        self._settings = [tuple(r) for r in st.tolist()]
        self._invalidate()
    setRangeInterpSettings = set_interp_settings

    def setRangeInterpSettingsJSON(self, path_or_obj):
        obj = path_or_obj
        if isinstance(path_or_obj, (str, os.PathLike)):
            with open(path_or_obj) as f:
                obj = json.load(f)
        self.set_interp_settings(obj)

    # -- matrix
    def eval_equation_matrix(self):
        """evalEqMatrix (src/ISRSolverSLE.cpp:138-149)."""
        G = np.empty((self._n, self._n))
        spread = L.dptr(self._ecm_err) if self._spread else None
        if self._eff_call is None:
            L.check(self._lib.isr_assemble(self._p, spread, L.dptr(G)))
        else:
            cnt = C.c_int64()
            L.check(self._lib.isr_assemble_nodes(self._p, 0, None, None, C.byref(cnt)))
            x, row = np.empty(cnt.value), np.empty(cnt.value, dtype=np.int32)
            L.check(self._lib.isr_assemble_nodes(self._p, cnt.value, L.dptr(x), L.iptr(row), C.byref(cnt)))
            en = self._ecm[row]
            try:
                ev = np.asarray(self._eff_call(x, en), dtype=np.float64)
                assert ev.shape == x.shape
            except Exception:
                ev = np.array([float(self._eff_call(float(a), float(b))) for a, b in zip(x, en)])
            L.check(self._lib.isr_assemble_sampled(self._p, cnt.value, L.dptr(np.ascontiguousarray(ev)), spread, L.dptr(G)))
        self._G = G
        self._eq_ready = True
        self._factored = False
        return 0
    evalEqMatrix = eval_equation_matrix

    def intop_matrix(self):
        """getIntegralOperatorMatrix: logical [i, j] view over column-major storage
        (include/PyISRSolverSLE.hpp:88-97)."""
        return self._G.T

    def set_matrix(self, G):
        """Inject an operator matrix (logical [i, j]) instead of assembling it."""
        self._G = np.ascontiguousarray(np.asarray(G, dtype=np.float64).T)
        L.check(self._lib.isr_set_matrix(self._p, L.dptr(self._G)))
        self._eq_ready = True
        self._factored = False

    def _ensure_factored(self):
        if not self._eq_ready:
            self.eval_equation_matrix()
        if not self._factored:
            L.check(self._lib.isr_factor(self._p, L.dptr(self._vcs_err)))
            self._factored = True

    def solve(self):
        """solve (src/ISRSolverSLE.cpp:86-95)."""
        self._ensure_factored()
        cov = np.empty((self._n, self._n))
        L.check(self._lib.isr_solve(self._p, L.dptr(self._vcs), L.dptr(self._bcs), L.dptr(cov)))
        self._cov = cov
        return 0

    def bcs_cov_matrix(self): return self._cov.T
    getBornCSCovMatrix = bcs_cov_matrix

    def bcs_inv_cov_matrix(self):
        """Inverse covariance G^T W G (the reference inverts Cov again: include/PyISRSolverSLE.hpp:72-86)."""
        self._ensure_factored()
        out = np.empty((self._n, self._n))
        L.check(self._lib.isr_get_inv_cov(self._p, L.dptr(out)))
        return out.T

    def bcs_err(self): return np.sqrt(np.diag(self._cov))

    def condnum_eval(self):
        """evalConditionNumber (src/ISRSolverSLE.cpp:200-204)."""
        if not self._eq_ready:
            self.eval_equation_matrix()
        c = C.c_double()
        L.check(self._lib.isr_cond_number(self._p, C.byref(c)))
        return c.value
    evalConditionNumber = condnum_eval

    def condnum_spread_scan(self, sigmas, want_singular_values=False):
        """Condition number of the operator matrix versus the beam-energy spread: the loop of
        src/isrsolver-condnum-test.cpp:123-135 (`sigmas`: K common spreads, 0 = disabled) or K x N per-point
        spreads.  One assembly + K Gauss-Hermite spread products and SVDs instead of K full assemblies.
        The solver's cached matrix is invalidated."""
        sg = np.ascontiguousarray(sigmas, dtype=np.float64)
        per_point = sg.ndim == 2
        if per_point and sg.shape[1] != self._n:
            raise ValueError(f"per-point spreads must be (K, {self._n})")
        K = sg.shape[0]
        cond = np.empty(K)
        sv = np.empty((K, self._n)) if want_singular_values else None
        L.check(self._lib.isr_cond_spread_scan(self._p, K, L.dptr(sg), int(per_point), L.dptr(cond), L.dptr(sv)))
        self._invalidate()
        return (cond, sv) if want_singular_values else cond

    def interp_eval(self, cm_energy):
        """interpEval(bcs, energy) (src/ISRSolverSLE.cpp:195-198)."""
        e = np.atleast_1d(np.asarray(cm_energy, dtype=np.float64)).ravel().copy()
        out = np.empty_like(e)
        L.check(self._lib.isr_interp_eval(self._p, L.dptr(self._bcs), e.size, L.dptr(e), L.dptr(out)))
        return float(out[0]) if np.ndim(cm_energy) == 0 else out.reshape(np.shape(cm_energy))

    def basis_eval(self, cs_index, energy, deriv=False):
        j = np.ascontiguousarray(np.atleast_1d(cs_index), dtype=np.int32)
        e = np.ascontiguousarray(np.atleast_1d(energy), dtype=np.float64)
        j, e = np.broadcast_arrays(j, e)
        j, e = np.ascontiguousarray(j), np.ascontiguousarray(e)
        out = np.empty(e.shape)
        L.check(self._lib.isr_basis_eval(self._p, e.size, L.iptr(j), L.dptr(e), int(deriv), L.dptr(out)))
        return out

    def save(self, output_path, vcs_name="vcs", bcs_name="bcs"):
        """save (src/ISRSolverSLE.cpp:97-123) with the ROOT objects replaced by arrays of the same names."""
        np.savez(output_path, **{
            vcs_name: np.stack([self._ecm, self._vcs, self._ecm_err, self._vcs_err]),
            bcs_name: np.stack([self._ecm, self._bcs, np.zeros(self._n), self.bcs_err()]),
            "intergalOperatorMatrix": self.intop_matrix(),   # sic, the reference's object name
            "covMatrixBornCS": self.bcs_cov_matrix(),
            "invCovMatrixBornCS": self.bcs_inv_cov_matrix(),
        })
        return 0

    # -- toy Monte-Carlo validation -------------------------------------------------------------------
    def _select_for_toys(self):
        self._ensure_factored()

    def set_toy_schedule(self, variant: int):
        """0 (default): exactly lower-triangular operators skip their zeros; 1: dense schedule on the operators as
        given (cross-check); 2: as 0, and a dense chi2 factor is replaced by its triangular QL factor (long campaigns).
        See isr_set_toy_kernel."""
        L.check(self._lib.isr_set_toy_kernel(self._p, int(variant)))

    def toy_batch(self, V, bcs0, want_chi2=True, want_sum=False, want_ratio=False, want_hist=False, want_bcs=False):
        """The loop body of chi2Test / ratioTestModel over all rows of V (toy-major (T, N), host)."""
        V = np.ascontiguousarray(V, dtype=np.float64)
        T = V.shape[0]
        if V.ndim != 2 or V.shape[1] != self._n:
            raise ValueError(f"V must be (T, {self._n})")
        bcs0 = _f64(bcs0, self._n, "bcs0")
        self._select_for_toys()
        out = {}
        chi2 = np.empty(T) if want_chi2 else None
        sb = np.zeros(self._n) if want_sum else None
        rs = np.zeros(3 * self._n) if want_ratio else None
        rh = np.zeros((self._n, 1026), dtype=np.uint32) if want_hist else None
        B = np.empty((T, self._n)) if want_bcs else None
        L.check(self._lib.isr_toy_batch(self._p, T, L.dptr(V), L.dptr(bcs0), L.dptr(chi2), L.dptr(sb), L.dptr(rs),
                                        L.uptr(rh), L.dptr(B)))
        if want_chi2: out["chi2"] = chi2
        if want_sum: out["sum_bcs"] = sb
        if want_ratio: out["ratio_stats"] = rs.reshape(self._n, 3)
        if want_hist: out["ratio_hist"] = rh
        if want_bcs: out["bcs"] = B
        return out

    def toy_campaign(self, n, vcs, vcs_err, bcs0, seed=TOY_SEED, first_toy=0, want_chi2=True, want_sum=False,
                     want_ratio=False, want_hist=False, want_chi2_hist=True, nbins=128):
        """A toy campaign by count with the toys drawn on the GPU (isr_toy_campaign): nothing but the results
        crosses PCIe.  Toy k depends only on (seed, k): shard a campaign over ranks with `first_toy`."""
        vcs, vcs_err, bcs0 = _f64(vcs, self._n, "vcs"), _f64(vcs_err, self._n, "vcs_err"), _f64(bcs0, self._n, "bcs0")
        self._select_for_toys()
        n = int(n)
        chi2 = np.empty(n) if want_chi2 else None
        sb = np.zeros(self._n) if want_sum else None
        rs = np.zeros(3 * self._n) if want_ratio else None
        rh = np.zeros((self._n, 1026), dtype=np.uint32) if want_hist else None
        lohi = np.zeros(2) if want_chi2_hist else None
        cnt = np.zeros(nbins + 2, dtype=np.uint32) if want_chi2_hist else None
        L.check(self._lib.isr_toy_campaign(self._p, n, int(first_toy), int(seed), L.dptr(vcs), L.dptr(vcs_err), L.dptr(bcs0),
                                           L.dptr(chi2), L.dptr(sb), L.dptr(rs), L.uptr(rh), int(nbins), L.dptr(lohi), L.uptr(cnt)))
        out = {}
        if want_chi2: out["chi2"] = chi2
        if want_sum: out["sum_bcs"] = sb
        if want_ratio: out["ratio_stats"] = rs.reshape(self._n, 3)
        if want_hist: out["ratio_hist"] = rh
        if want_chi2_hist: out.update(hist=cnt.astype(np.float32), lo=float(lohi[0]), hi=float(lohi[1]))
        return out

    def chi2_test_call(self, bcs0, n=None, V=None, vcs=None, vcs_err=None, seed=TOY_SEED, first_toy=0, nbins=128,
                       want_chi2=True, want_sum=False, want_ratio=False, want_hist=False, ngpu=1):
        """chi2Test / ratioTestModel as ONE stream-ordered C-ABI call (isr_chi2_test): toys -> range -> TH1F cells with a
        single synchronisation.  V (host, (T, N)): the toys; or n + vcs (+ vcs_err): toys drawn on the device.
        ngpu > 1 (plain SLE solvers with tabulated efficiency): the test is spread over the first `ngpu` GPUs of the box
        from this process (isr_group: a thread, a context and a replicated problem per GPU, two small all-reduces); the
        histogram is bit-identical to the single-GPU one."""
        bcs0 = _f64(bcs0, self._n, "bcs0")
        a = L.Chi2TestArgs()
        keep = [bcs0]
        if V is not None:
            V = np.ascontiguousarray(V, dtype=np.float64)
            if V.ndim != 2 or V.shape[1] != self._n:
                raise ValueError(f"V must be (T, {self._n})")
            T = V.shape[0]
            a.V = L.dptr(V)
        else:
            T = int(n)
            vcs = _f64(vcs, self._n, "vcs"); keep.append(vcs)
            a.vcs = L.dptr(vcs)
        err = _f64(self._vcs_err if vcs_err is None else vcs_err, self._n, "vcs_err"); keep.append(err)
        a.vcs_err, a.bcs0, a.n_toys, a.seed, a.first_toy, a.nbins = L.dptr(err), L.dptr(bcs0), T, int(seed), int(first_toy), int(nbins)
        a.flags = L.ISR_TEST_RATIO if (want_ratio or want_hist) else 0
        chi2 = np.empty(T) if want_chi2 else None
        lohi, cnt = np.zeros(2), np.zeros(nbins + 2, dtype=np.int64)
        sb = np.zeros(self._n) if want_sum else None
        rs = np.zeros(3 * self._n) if (want_ratio or want_hist) else None
        rh = np.zeros((self._n, 1026), dtype=np.uint32) if want_hist else None
        total = C.c_int64()
        a.chi2_out, a.chi2_lo_hi_out, a.chi2_counts_out = L.dptr(chi2), L.dptr(lohi), cnt.ctypes.data_as(C.POINTER(C.c_int64))
        a.sum_bcs_out, a.ratio_stats_out, a.ratio_hist_out, a.total_toys_out = L.dptr(sb), L.dptr(rs), L.uptr(rh), C.pointer(total)
        if ngpu > 1:
            gp = self._group_problem(int(ngpu))
            a.flags |= L.ISR_TEST_ASSEMBLE | L.ISR_TEST_FACTOR
            if self._spread:
                ee = _f64(self._ecm_err, self._n, "energy_err"); keep.append(ee)
                a.ecm_err = L.dptr(ee)
            L.check(self._lib.isr_group_chi2_test(gp, C.byref(a)))
        else:
            self._select_for_toys()
            L.check(self._lib.isr_chi2_test(self._p, C.byref(a)))
        out = {"hist": cnt.astype(np.float32), "lo": float(lohi[0]), "hi": float(lohi[1]), "n": int(total.value)}
        if want_chi2: out["chi2"] = chi2
        if want_sum: out["sum_bcs"] = sb
        if want_ratio or want_hist: out["ratio_stats"] = rs.reshape(self._n, 3)
        if want_hist: out["ratio_hist"] = rh
        return out

    def _group_problem(self, ngpu):
        """The replicated problem of this solver on an isr_group of `ngpu` GPUs (cached per solver and size)."""
        if type(self)._select_for_toys is not ISRSolverSLE._select_for_toys or self._eff_call is not None:
            raise ValueError("multi-GPU tests need a plain SLE solver with a tabulated efficiency (the problem is rebuilt "
                             "from its description on every GPU)")
        cache = self.__dict__.setdefault("_gproblems", {})
        if ngpu not in cache:
            grp = _group(ngpu)
            gp = C.c_void_p()
            st = np.ascontiguousarray(self._settings, dtype=np.int32).reshape(-1, 3) if self._settings is not None else np.zeros((0, 3), dtype=np.int32)
            desc = self._eff.desc(self._n) if self._eff is not None and self._eff.kind != "one" else None
            L.check(self._lib.isr_group_problem_create(grp, self._n, L.dptr(self._ecm), self._thr, len(st),
                                                       L.iptr(st) if len(st) else None, C.byref(desc) if desc is not None else None,
                                                       C.byref(gp)))
            cache[ngpu] = gp
        return cache[ngpu]

    def chi2_test_sharded(self, n, vcs, bcs0, vcs_err, seed=TOY_SEED, nbins=128):
        """chi2Test over `n` toys sharded across the ranks of an initialised torch.distributed group (one
        process per GPU): rank r draws and processes toys [begin_r, end_r) of the device stream `seed`
        (toy k depends only on (seed, k)), then the TH1F(nbins, min, max) is combined with two O(N)-byte
        collectives (dist.histogram_sharded) -- chi2 never leaves the GPUs.  Every rank returns the same
        result; the histogram and its range are bit-identical to toy_campaign(n) on one GPU (sum_bcs / ratio_stats agree
        to rounding: float sums are re-associated across ranks).  Ranks without toys (n < world) are legal."""
        import torch.distributed as td
        from . import dist as D
        rank, world = td.get_rank(), td.get_world_size()
        b, e = D.shard_bounds(int(n), world, rank)
        if e > b:
            r = self.toy_campaign(e - b, vcs, vcs_err, bcs0, seed, first_toy=b, want_chi2=False, want_sum=True,
                                  want_ratio=True, want_chi2_hist=True, nbins=nbins)
        else:
            r = {"lo": np.inf, "hi": -np.inf, "sum_bcs": np.zeros(self._n), "ratio_stats": np.zeros((self._n, 3))}
        counts = np.zeros(nbins + 2, dtype=np.uint32)

        def rebin(lo, hi):
            L.check(self._lib.isr_last_chi2_histogram(self._p, int(nbins), lo, hi, L.uptr(counts)))
            return counts
        dev = None
        if td.get_backend() == "nccl":
            import torch
            dev = torch.device("cuda", torch.cuda.current_device())
        cnt, glo, ghi, sb, rs = D.histogram_sharded(np.empty(0), r["sum_bcs"], r["ratio_stats"], nbins,
                                                    count_fn=rebin if e > b else (lambda lo, hi: counts), device=dev,
                                                    local_range=(r["lo"], r["hi"]))
        return {"hist": cnt.astype(np.float32), "lo": glo, "hi": ghi, "sum_bcs": sb, "ratio_stats": rs, "n": int(n)}

    def chi2_test(self, n, vcs, bcs0, vcs_err, seed=TOY_SEED, toys=None, rng="host"):
        """chi2Test (src/Chi2Test.cpp:29-104) without the ROOT fit/file: returns the chi2 values and the
        TH1F(128, min, max) counts (under/overflow included).  rng="host": toys from the seeded NumPy stream
        (or `toys`), copied to the GPU; rng="device": toys drawn on the GPU (toy_campaign), no bulk transfer."""
        if rng == "device" and toys is None:
            return self.toy_campaign(n, vcs, vcs_err, bcs0, seed)
        V = draw_toys(vcs, vcs_err, n, seed, pinned=True) if toys is None else toys   # pinned: H2D at the PCIe rate
        chi2 = self.toy_batch(V, bcs0)["chi2"]
        lo, hi = C.c_double(), C.c_double()
        counts = np.zeros(130, dtype=np.uint32)
        # range and TH1F from the chi2 values the batch left on the device (no second upload)
        L.check(self._lib.isr_last_chi2_minmax(self._p, C.byref(lo), C.byref(hi)))
        L.check(self._lib.isr_last_chi2_histogram(self._p, 128, lo.value, hi.value, L.uptr(counts)))
        return {"chi2": chi2, "hist": counts.astype(np.float32), "lo": lo.value, "hi": hi.value}

    @staticmethod
    def _load_model(model_path, vcs_name, bcs_name):
        with np.load(model_path) as f:
            v, b = np.asarray(f[vcs_name], dtype=np.float64), np.asarray(f[bcs_name], dtype=np.float64)
        v = v[1] if v.ndim == 2 else v      # graphs are stored as rows (x, y, ex, ey)
        b = b[1] if b.ndim == 2 else b
        if v.shape != b.shape:
            raise DifferentModelCrossSectionSizes("[!] Model visible and Born cross sections have different sizes.")
        return v, b

    def chi2_test_model(self, n, initial_ampl, model_path, model_visible_cs_name, model_born_cs_name,
                        output_path=None, seed=TOY_SEED, rng="host"):
        """chi2TestModel (src/Chi2Test.cpp:113-151); kwargs as include/PyISRSolverSLE.hpp:125-153."""
        self.solve()
        vcs, bcs0 = self._load_model(model_path, model_visible_cs_name, model_born_cs_name)
        res = self.chi2_test(n, vcs, bcs0, self._vcs_err, seed, rng=rng)
        res["initial_ampl"] = float(initial_ampl)
        if output_path:
            np.savez(output_path, chi2Hist=res["hist"], chi2Min=res["lo"], chi2Max=res["hi"], chi2=res["chi2"])
        return res

    def chi2_test_data(self, n, initial_ampl=1e4, output_path=None, seed=TOY_SEED, rng="host"):
        """chi2TestData (src/Chi2Test.cpp:160-195): reference = mean solution over n toys, then chi2Test."""
        self.solve()
        vcs = self._vcs.copy()
        if rng == "device":
            s = self.toy_campaign(n, vcs, self._vcs_err, np.ones(self._n), seed, want_chi2=False, want_sum=True,
                                  want_chi2_hist=False)["sum_bcs"]
        else:
            V = draw_toys(vcs, self._vcs_err, n, seed)
            s = self.toy_batch(V, np.ones(self._n), want_chi2=False, want_sum=True)["sum_bcs"]
        bcs0 = s / n
        res = self.chi2_test(n, vcs, bcs0, self._vcs_err, seed + 1, rng=rng)
        res["bcs0"] = bcs0
        if output_path:
            np.savez(output_path, chi2Hist=res["hist"], chi2Min=res["lo"], chi2Max=res["hi"], chi2=res["chi2"])
        return res

    def ratio_test_model(self, n, model_path, model_visible_cs_name, model_born_cs_name, output_path=None,
                         seed=TOY_SEED, toys=None, rng="host"):
        """ratioTestModel (src/RatioTest.cpp:13-62): per point TH1F(1024,-2,2) of (b - b0)/b0, GetMean and
        GetMeanError of the in-range entries."""
        self.solve()
        vcs, bcs0 = self._load_model(model_path, model_visible_cs_name, model_born_cs_name)
        if rng == "device" and toys is None:
            r = self.toy_campaign(n, vcs, self._vcs_err, bcs0, seed, want_chi2=False, want_ratio=True, want_hist=True,
                                  want_chi2_hist=False)
        else:
            V = draw_toys(vcs, self._vcs_err, n, seed) if toys is None else toys
            r = self.toy_batch(V, bcs0, want_chi2=False, want_ratio=True, want_hist=True)
        cnt, sx, sx2 = r["ratio_stats"].T
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = np.where(cnt > 0, sx / cnt, 0.0)
            rms = np.sqrt(np.abs(np.where(cnt > 0, sx2 / cnt, 0.0) - mean * mean))
            err = np.where(cnt > 0, rms / np.sqrt(cnt), 0.0)
        res = {"means": mean, "mean_errors": err, "hist": r["ratio_hist"].astype(np.float32), "ecm": self._ecm}
        if output_path:
            np.savez(output_path, means=np.stack([self._ecm, mean, np.zeros(self._n), err]),
                     **{f"bias_pt_{i}": res["hist"][i] for i in range(self._n)})
        return res


class IterISRInterpSolver(ISRSolverSLE):
    """IterISRInterpSolver (include/IterISRInterpSolver.hpp; src/IterISRInterpSolver.cpp:45-65; PyISR
    "IterISRInterpSolver", include/PyIterISRInterpSolver.hpp): the iterative radiative-correction method on the
    same operator matrix.  bcs_cov_matrix() is diag((vcs_err / radcorr)^2) as in the reference."""

    def __init__(self, *args, **kwargs):
        self._n_iter = int(kwargs.pop("n_iter", 10))      # _nIter(10), src/IterISRInterpSolver.cpp:19
        super().__init__(*args, **kwargs)
        self._radcorr = np.zeros(self._n)

    n_iter = property(lambda self: self._n_iter, lambda self, v: self.setNumOfIters(v))
    def getNumOfIters(self): return self._n_iter
    def setNumOfIters(self, n): self._n_iter = int(n)
    def radcorr(self): return self._radcorr

    def solve(self):
        if not self._eq_ready:
            self.eval_equation_matrix()
        L.check(self._lib.isr_iterative_solve(self._p, self._n_iter, L.dptr(self._vcs), L.dptr(self._bcs), L.dptr(self._radcorr)))
        self._cov = np.diag((self._vcs_err / self._radcorr) ** 2) if self._n_iter > 0 else np.zeros((self._n, self._n))
        return 0

    def bcs_inv_cov_matrix(self):
        return np.diag(1.0 / np.diag(self._cov))

    def save(self, output_path, vcs_name="vcs", bcs_name="bcs"):
        np.savez(output_path, **{
            vcs_name: np.stack([self._ecm, self._vcs, self._ecm_err, self._vcs_err]),
            bcs_name: np.stack([self._ecm, self._bcs, np.zeros(self._n), self.bcs_err()]),
            "radiative_correction": np.stack([self._ecm, self._radcorr]),
            "intergalOperatorMatrix": self.intop_matrix(), "covMatrixBornCS": self.bcs_cov_matrix(),
            "invCovMatrixBornCS": self.bcs_inv_cov_matrix()})
        return 0


class ISRSolverSLE2(ISRSolverSLE):
    """ISRSolverSLE2: efficiency eps(x, energyIndex) from one x-histogram per energy point
    (src/BaseISRSolver2.cpp:151-160, 196-204, 231-236; src/ISRSolverSLE2.cpp:142-153)."""

    def __init__(self, n, energy, vcs, energy_err, vcs_err, threshold, eff_hists=None, nbins=None, xlo=0.0, xhi=1.0,
                 xedges=None, enableEnergySpread=False, device=None):
        eff = None
        if eff_hists is not None:
            vals = np.ascontiguousarray(eff_hists, dtype=np.float64)
            order = np.argsort(np.asarray(energy, dtype=np.float64), kind="stable")
            vals = vals[order]     # histograms follow their energy points through the sort
            eff = Efficiency("row_table", values=vals, nx=(vals.shape[1] - 2 if nbins is None else nbins),
                             xlo=xlo, xhi=xhi, xedges=xedges)
        super().__init__(n, energy, vcs, energy_err, vcs_err, threshold, eff, enableEnergySpread, device)


class ISRSolverTikhonov(ISRSolverSLE):
    """ISRSolverTikhonov (include/ISRSolverTikhonov.hpp; src/ISRSolverTikhonov.cpp)."""

    def __init__(self, *args, **kwargs):
        lam = kwargs.pop("reg_param", 1.0)      # _lambda(1.), src/ISRSolverTikhonov.cpp:30
        super().__init__(*args, **kwargs)
        self._lambda = float(lam)
        self._deriv_reg = True                  # _enabledDerivNorm2Reg(true), :29
        self._tik_ready = False

    def _invalidate(self):
        super()._invalidate()
        self._tik_ready = False

    reg_param = property(lambda self: self._lambda, lambda self, v: self.setLambda(v))
    def getLambda(self): return self._lambda
    def setLambda(self, lam): self._lambda = float(lam)
    def isDerivNorm2RegIsEnabled(self): return self._deriv_reg
    def enableDerivNorm2Regularizator(self): self._deriv_reg = True; self._tik_ready = False
    def disableDerivNorm2Regularizator(self): self._deriv_reg = False; self._tik_ready = False

    def _ensure_tik(self):
        if not self._eq_ready:
            self.eval_equation_matrix()
            self._tik_ready = False
        if not self._tik_ready:
            L.check(self._lib.isr_tikhonov_prepare(self._p, L.dptr(self._vcs_err), int(self._deriv_reg)))
            self._tik_ready = True
            self._factored = True

    def resetVisibleCSErrors(self, vcs_err):
        super().resetVisibleCSErrors(vcs_err)
        self._tik_ready = False

    def solve(self):
        """solve (src/ISRSolverTikhonov.cpp:61-72)."""
        self._ensure_tik()
        cov = np.empty((self._n, self._n))
        L.check(self._lib.isr_tikhonov_solve(self._p, self._lambda, L.dptr(self._vcs), L.dptr(self._bcs), L.dptr(cov)))
        self._cov = cov
        return 0

    def bcs_inv_cov_matrix(self):
        return np.linalg.inv(self.bcs_cov_matrix())

    def _scan(self, lambdas, want_bcs=False):
        self._ensure_tik()
        lam = _f64(np.atleast_1d(lambdas))
        Lm = lam.size
        x, y, c, dc = (np.empty(Lm) for _ in range(4))
        B = np.empty((Lm, self._n)) if want_bcs else None
        L.check(self._lib.isr_tikhonov_scan(self._p, Lm, L.dptr(lam), L.dptr(self._vcs), L.dptr(x), L.dptr(y),
                                            L.dptr(c), L.dptr(dc), L.dptr(B)))
        return x, y, c, dc, B

    def evalEqNorm2(self): return float(self._scan([self._lambda])[0][0])
    def evalSmoothnessConstraintNorm2(self): return float(self._scan([self._lambda])[1][0])
    def evalLCurveCurvature(self): return float(self._scan([self._lambda])[2][0])
    def evalLCurveCurvatureDerivative(self): return float(self._scan([self._lambda])[3][0])

    def make_LCurve(self, lambdas):
        """make_LCurve (include/PyISRSolverTikhonov.hpp:44-79): x = |G b - v|^2_W, y = b^T F b, c = -curvature,
        for every lambda, from one decomposition instead of one solve() per lambda.  Like the reference,
        the solver is left at the last lambda."""
        lam = np.asarray(lambdas, dtype=np.float64)
        if lam.ndim != 1:
            raise TypeError("Wrong dimension of the array with regularization parameter values. The array must be 1D array.")
        x, y, c, _, _ = self._scan(lam)
        if lam.size:
            self._lambda = float(lam[-1])
            self.solve()
        return {"x": x, "y": y, "c": -c}

    def solve_LCurve(self, initial_lambda, lo=1e-12, hi=None):
        """solve_LCurve (include/PyISRSolverTikhonov.hpp:81-109): maximise the L-curve curvature over
        lambda >= 0.  NLopt LD_MMA (absent here) is replaced by a dense log-grid scan of the O(N)-per-lambda
        curvature followed by golden-section refinement to xtol_rel 1e-6."""
        hi = max(10.0 * initial_lambda, 1.0) if hi is None else hi
        grid = np.geomspace(lo, hi, 2048)
        obj = self._scan(grid)[2]          # curvature = -|...| <= 0 ; NLopt minimises it
        k = int(np.argmin(obj))
        a, b = grid[max(k - 1, 0)], grid[min(k + 1, len(grid) - 1)]
        gr = (np.sqrt(5.0) - 1.0) / 2.0
        c_, d_ = b - gr * (b - a), a + gr * (b - a)
        fc, fd = self._scan([c_])[2][0], self._scan([d_])[2][0]
        while abs(b - a) > 1e-6 * abs(0.5 * (a + b)):
            if fc < fd:
                b, d_, fd = d_, c_, fc
                c_ = b - gr * (b - a)
                fc = self._scan([c_])[2][0]
            else:
                a, c_, fc = c_, d_, fd
                d_ = a + gr * (b - a)
                fd = self._scan([d_])[2][0]
        lam = 0.5 * (a + b)
        self._lambda = float(lam)
        self.solve()
        return {"lambda": float(lam), "max_curv": float(-self._scan([lam])[2][0])}

    def _select_for_toys(self):
        self._ensure_tik()
        L.check(self._lib.isr_tikhonov_select(self._p, self._lambda))

    def _ensure_factored(self):
        self._ensure_tik()

    def toy_scan(self, lambdas, V, bcs0):
        """chi2 of every (lambda, toy) pair, (L, T): the chi2Test loop of isrsolver-Tikhonov-chi2-test for
        all lambdas at once."""
        self._ensure_tik()
        lam = _f64(np.atleast_1d(lambdas))
        V = np.ascontiguousarray(V, dtype=np.float64)
        out = np.empty((lam.size, V.shape[0]))
        L.check(self._lib.isr_tikhonov_toy_scan(self._p, lam.size, L.dptr(lam), V.shape[0], L.dptr(V),
                                                L.dptr(_f64(bcs0, self._n)), L.dptr(out)))
        return out

    def toy_scan_hist(self, lambdas, V, bcs0, nbins=128):
        """The chi2 test (chi2Test, src/Chi2Test.cpp:38-78) for every lambda at once, summarised on the GPU:
        returns {"min", "max", "mean"} (L,) and "hist" (L, nbins + 2) float32 TH1F cells per lambda."""
        self._ensure_tik()
        lam = _f64(np.atleast_1d(lambdas))
        V = np.ascontiguousarray(V, dtype=np.float64)
        stats = np.empty((lam.size, 3))
        counts = np.zeros((lam.size, nbins + 2), dtype=np.uint32)
        L.check(self._lib.isr_tikhonov_toy_scan_hist(self._p, lam.size, L.dptr(lam), V.shape[0], L.dptr(V),
                                                     L.dptr(_f64(bcs0, self._n)), int(nbins), L.dptr(stats), L.uptr(counts)))
        return {"min": stats[:, 0].copy(), "max": stats[:, 1].copy(), "mean": stats[:, 2].copy(),
                "hist": counts.astype(np.float32)}

    def toy_scan_hist_sharded(self, lambdas, V, bcs0, nbins=128):
        """BASELINE config C4: the lambda grid sharded over the ranks of an initialised torch.distributed group
        (lambdas are independent given the eigen-system, src/isrsolver-LCurve-plot.cpp:131-151); every rank
        scans its contiguous slice against all toys and ONE all_gather of the per-lambda summaries
        [min, max, mean, TH1F cells] follows.  Every rank returns the full grid."""
        import torch
        import torch.distributed as td
        from . import dist as D
        lam = _f64(np.atleast_1d(lambdas))
        rank, world = td.get_rank(), td.get_world_size()
        b, e = D.shard_bounds(lam.size, world, rank)
        width = 3 + nbins + 2
        per = max(D.shard_bounds(lam.size, world, r)[1] - D.shard_bounds(lam.size, world, r)[0] for r in range(world))
        mine = np.zeros((per, width))
        if e > b:
            r = self.toy_scan_hist(lam[b:e], V, bcs0, nbins)
            mine[: e - b, 0], mine[: e - b, 1], mine[: e - b, 2] = r["min"], r["max"], r["mean"]
            mine[: e - b, 3:] = r["hist"]          # integer counts < 2^53: exact in float64
        t = torch.from_numpy(mine)
        if td.get_backend() == "nccl":
            t = t.to(torch.device("cuda", torch.cuda.current_device()))
        out = torch.empty((world,) + tuple(t.shape), dtype=torch.float64, device=t.device)
        td.all_gather_into_tensor(out.view(-1), t.view(-1))        # the single collective
        out = out.cpu().numpy()
        rows = [out[r, : D.shard_bounds(lam.size, world, r)[1] - D.shard_bounds(lam.size, world, r)[0]] for r in range(world)]
        full = np.concatenate(rows)
        return {"min": full[:, 0].copy(), "max": full[:, 1].copy(), "mean": full[:, 2].copy(),
                "hist": full[:, 3:].astype(np.float32)}

    def regulariser(self):
        self._ensure_tik()
        F = np.empty((self._n, self._n))
        L.check(self._lib.isr_get_regulariser(self._p, L.dptr(F)))
        return F.T


class ISRSolverTSVD(ISRSolverSLE):
    """ISRSolverTSVD (include/ISRSolverTSVD.hpp; src/ISRSolverTSVD.cpp)."""

    def __init__(self, *args, **kwargs):
        k = kwargs.pop("upper_TSVD_index", None)
        super().__init__(*args, **kwargs)
        self._k = self._n if k is None else int(k)   # _upperTSVDIndex(numberOfPoints), src/ISRSolverTSVD.cpp:19
        self._keep_one = False

    upper_TSVD_index = property(lambda self: self._k, lambda self, v: self.setUpperTSVDIndex(v))
    def getUpperTSVDIndex(self): return self._k
    def setUpperTSVDIndex(self, k): self._k = int(k)
    def enable_keepone(self): self._keep_one = True
    def disable_keepone(self): self._keep_one = False
    enableKeepOne, disableKeepOne = enable_keepone, disable_keepone

    def solve(self):
        """solve (src/ISRSolverTSVD.cpp:53-74).  Cov = (K^T W K)^-1 exists only for k == N; otherwise the
        reference inverts a singular matrix and this implementation leaves bcs_cov_matrix() unset."""
        if not self._eq_ready:
            self.eval_equation_matrix()
        full = (self._k == self._n) and not self._keep_one
        cov = np.empty((self._n, self._n)) if full else None
        L.check(self._lib.isr_tsvd_solve(self._p, self._k, int(self._keep_one), L.dptr(self._vcs),
                                         L.dptr(self._vcs_err), L.dptr(self._bcs), L.dptr(cov)))
        self._cov = cov
        self._factored = False
        return 0

    def singular_values(self):
        if not self._eq_ready:
            self.eval_equation_matrix()
        s = np.empty(self._n)
        L.check(self._lib.isr_singular_values(self._p, L.dptr(s)))
        return s

    def _select_for_toys(self):
        if not self._eq_ready:
            self.eval_equation_matrix()
        L.check(self._lib.isr_tsvd_select(self._p, self._k, int(self._keep_one), L.dptr(self._vcs_err)))
        self._factored = False
