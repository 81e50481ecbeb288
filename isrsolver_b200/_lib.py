"""ctypes binding of libisr_b200.so (the C ABI in include/isr_b200.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible when a context is
created, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISR_B200_LIB") or os.path.join(_HERE, "libisr_b200.so")   # override: kernel experiments only

ISR_OK, ISR_EINVAL, ISR_ERANGE, ISR_ECUDA, ISR_ESTATE, ISR_ENUMERIC, ISR_EEFFDIM = range(7)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint32)
_i64 = C.c_int64
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p


class IsrGpuError(RuntimeError):
    """CUDA / cuSOLVER / state errors reported by libisr_b200."""


class InterpRangeException(ValueError):
    """include/Interpolator.hpp:14-18 -- wrong interpolation ranges."""


class EfficiencyDimensionException(ValueError):
    """include/BaseISRSolver.hpp:24-28 -- wrong efficiency dimension."""


class EpsDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("nx", C.c_int), ("xlo", C.c_double), ("xhi", C.c_double),
                ("xedges", _dp), ("ne", C.c_int), ("elo", C.c_double), ("ehi", C.c_double),
                ("eedges", _dp), ("values", _dp)]


class Chi2TestArgs(C.Structure):
    """isr_chi2_test_args (include/isr_b200.h): one whole chi2 / ratio test."""
    _fields_ = [("flags", C.c_int), ("ecm_err", _dp), ("vcs_err", _dp), ("bcs0", _dp), ("n_toys", _i64),
                ("V", _dp), ("V_dev", _vp), ("vcs", _dp), ("seed", C.c_uint64), ("first_toy", _i64),
                ("nbins", C.c_int), ("comm", _vp),
                ("chi2_out", _dp), ("chi2_lo_hi_out", _dp), ("chi2_counts_out", _i64p), ("sum_bcs_out", _dp),
                ("ratio_stats_out", _dp), ("ratio_hist_out", _up), ("total_toys_out", _i64p),
                ("refactored_out", C.POINTER(C.c_int))]


class TikhonovScanArgs(C.Structure):
    """isr_tikhonov_scan_args (include/isr_b200.h): per-lambda chi2 test, toys sharded over a communicator."""
    _fields_ = [("flags", C.c_int), ("ecm_err", _dp), ("vcs_err", _dp), ("deriv_reg", C.c_int), ("n_lambda", _i64),
                ("lambda_", _dp), ("n_toys", _i64), ("V", _dp), ("V_dev", _vp), ("bcs0", _dp), ("nbins", C.c_int),
                ("comm", _vp), ("stats_out", _dp), ("counts_out", _up), ("total_toys_out", _i64p)]


ISR_TEST_ASSEMBLE, ISR_TEST_FACTOR, ISR_TEST_RATIO = 1, 2, 4
ISR_COMM_ID_BYTES = 128

# name -> (restype, argtypes); every symbol include/isr_b200.h declares
SIGNATURES = {
    "isr_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "isr_ctx_destroy": (None, [_vp]),
    "isr_last_error": (C.c_char_p, []),
    "isr_ctx_stream": (_vp, [_vp]),
    "isr_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(C.c_void_p)]),
    "isr_host_free": (None, [C.c_void_p]),
    "isr_ctx_synchronize": (C.c_int, [_vp]),
    "isr_launch_count": (_i64, []),
    "isr_problem_create": (C.c_int, [_vp, C.c_int, _dp, C.c_double, C.c_int, _ip, C.POINTER(EpsDesc), C.POINTER(_vp)]),
    "isr_problem_destroy": (None, [_vp]),
    "isr_problem_set_ranges": (C.c_int, [_vp, C.c_int, _ip]),
    "isr_assemble": (C.c_int, [_vp, _dp, _dp]),
    "isr_assemble_nodes": (C.c_int, [_vp, _i64, _dp, _ip, _i64p]),
    "isr_assemble_sampled": (C.c_int, [_vp, _i64, _dp, _dp, _dp]),
    "isr_set_matrix": (C.c_int, [_vp, _dp]),
    "isr_assemble_stats": (C.c_int, [_vp, _i64p, _i64p]),
    "isr_debug_panels": (C.c_int, [_vp, _i64, _dp, _ip, _i64p]),
    "isr_debug_moments": (C.c_int, [_vp, _dp]),
    "isr_kernel_kf": (C.c_int, [_vp, _i64, _dp, _dp, _dp]),
    "isr_basis_eval": (C.c_int, [_vp, _i64, _ip, _dp, C.c_int, _dp]),
    "isr_basis_integrals": (C.c_int, [_vp, _dp]),
    "isr_interp_eval": (C.c_int, [_vp, _dp, _i64, _dp, _dp]),
    "isr_factor": (C.c_int, [_vp, _dp]),
    "isr_solve": (C.c_int, [_vp, _dp, _dp, _dp]),
    "isr_iterative_solve": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "isr_get_inverse": (C.c_int, [_vp, _dp]),
    "isr_get_inv_cov": (C.c_int, [_vp, _dp]),
    "isr_cond_number": (C.c_int, [_vp, _dp]),
    "isr_cond_spread_scan": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _dp, _dp]),
    "isr_toy_batch": (C.c_int, [_vp, _i64, _dp, _dp, _dp, _dp, _dp, _up, _dp]),
    "isr_toy_batch_dev": (C.c_int, [_vp, _i64, _vp, _dp, _vp, _vp, _vp, _vp, _vp]),
    "isr_toy_campaign": (C.c_int, [_vp, _i64, _i64, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp, _up, C.c_int, _dp, _up]),
    "isr_toy_campaign_dev": (C.c_int, [_vp, _i64, _i64, C.c_uint64, _dp, _dp, _dp, _vp, _vp, _vp, _vp]),
    "isr_draw_toys": (C.c_int, [_vp, _i64, C.c_int, _i64, C.c_uint64, _dp, _dp, _dp]),
    "isr_set_toy_kernel": (C.c_int, [_vp, C.c_int]),
    "isr_toy_tile_name": (C.c_char_p, [C.c_int, C.c_int]),
    "isr_minmax_dev": (C.c_int, [_vp, _i64, _vp, _dp, _dp]),
    "isr_minmax_to_dev": (C.c_int, [_vp, _i64, _vp, _vp]),
    "isr_histogram_range_dev": (C.c_int, [_vp, _i64, _vp, C.c_int, _vp, _vp]),
    "isr_histogram_dev": (C.c_int, [_vp, _i64, _vp, C.c_int, C.c_double, C.c_double, _up]),
    "isr_last_chi2_histogram": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _up]),
    "isr_last_chi2_minmax": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "isr_histogram": (C.c_int, [_vp, _i64, _dp, C.c_int, C.c_double, C.c_double, _up]),
    "isr_chi2_histogram": (C.c_int, [_vp, _i64, _dp, C.c_int, _dp, _dp, _up]),
    "isr_comm_unique_id": (C.c_int, [_vp]),
    "isr_comm_init_rank": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "isr_comm_init_all": (C.c_int, [C.c_int, C.POINTER(_vp), C.POINTER(_vp)]),
    "isr_comm_destroy": (None, [_vp]),
    "isr_comm_uses_peer_memory": (C.c_int, [_vp]),
    "isr_comm_rank": (C.c_int, [_vp]),
    "isr_comm_size": (C.c_int, [_vp]),
    "isr_comm_allreduce_dev": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int]),
    "isr_ctx_bind_host_thread": (C.c_int, [_vp, _ip]),
    "isr_chi2_test": (C.c_int, [_vp, C.POINTER(Chi2TestArgs)]),
    "isr_chi2_test_timing": (C.c_int, [_vp, _dp]),
    "isr_group_create": (C.c_int, [C.c_int, _ip, C.POINTER(_vp)]),
    "isr_group_destroy": (None, [_vp]),
    "isr_group_size": (C.c_int, [_vp]),
    "isr_group_ctx": (_vp, [_vp, C.c_int]),
    "isr_group_problem_create": (C.c_int, [_vp, C.c_int, _dp, C.c_double, C.c_int, _ip, C.POINTER(EpsDesc), C.POINTER(_vp)]),
    "isr_group_problem_destroy": (None, [_vp]),
    "isr_group_problem_member": (_vp, [_vp, C.c_int]),
    "isr_group_host_alloc": (C.c_int, [_vp, _i64, C.c_int, C.POINTER(_dp)]),
    "isr_group_host_free": (None, [_vp, _dp]),
    "isr_group_chi2_test": (C.c_int, [_vp, C.POINTER(Chi2TestArgs)]),
    "isr_group_cond_spread_scan": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _dp, _dp]),
    "isr_group_tikhonov_chi2_scan": (C.c_int, [_vp, C.POINTER(TikhonovScanArgs)]),
    "isr_tikhonov_prepare": (C.c_int, [_vp, _dp, C.c_int]),
    "isr_tikhonov_scan": (C.c_int, [_vp, _i64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "isr_tikhonov_solve": (C.c_int, [_vp, C.c_double, _dp, _dp, _dp]),
    "isr_tikhonov_toy_scan": (C.c_int, [_vp, _i64, _dp, _i64, _dp, _dp, _dp]),
    "isr_tikhonov_toy_scan_dev": (C.c_int, [_vp, _i64, _dp, _i64, _vp, _dp, _vp]),
    "isr_tikhonov_toy_scan_hist": (C.c_int, [_vp, _i64, _dp, _i64, _dp, _dp, C.c_int, _dp, _up]),
    "isr_tikhonov_chi2_scan": (C.c_int, [_vp, C.POINTER(TikhonovScanArgs)]),
    "isr_set_scan_kernel": (C.c_int, [_vp, C.c_int]),
    "isr_tikhonov_select": (C.c_int, [_vp, C.c_double]),
    "isr_get_regulariser": (C.c_int, [_vp, _dp]),
    "isr_tsvd_solve": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp]),
    "isr_singular_values": (C.c_int, [_vp, _dp]),
    "isr_tsvd_select": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "isr_conv_plan_create": (C.c_int, [_vp, _i64, _dp, _dp, _dp, C.POINTER(EpsDesc), C.c_int, _dp, C.POINTER(_vp)]),
    "isr_conv_plan_destroy": (None, [_vp]),
    "isr_conv_plan_size": (C.c_int, [_vp, _i64p, _i64p]),
    "isr_conv_plan_nodes": (C.c_int, [_vp, _dp, _dp, _ip, _dp]),
    "isr_conv_plan_feed": (C.c_int, [_vp, _dp]),
    "isr_conv_plan_feed_interp": (C.c_int, [_vp, _vp, _dp]),
    "isr_conv_plan_refine": (C.c_int, [_vp, C.c_double, C.c_double, _i64p]),
    "isr_conv_plan_result": (C.c_int, [_vp, _dp, _dp]),
    "isr_conv_plan_reset": (C.c_int, [_vp]),
}

_LIB = None


def load():
    """dlopen libisr_b200.so and bind every declared symbol.  Raises if the library is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise IsrGpuError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  isrsolver_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(rc: int):
    if rc == ISR_OK:
        return
    msg = load().isr_last_error().decode("utf-8", "replace")
    if rc == ISR_ERANGE:
        raise InterpRangeException(msg)
    if rc == ISR_EEFFDIM:
        raise EfficiencyDimensionException(msg)
    if rc == ISR_EINVAL:
        raise ValueError(msg)
    raise IsrGpuError(f"[isr rc={rc}] {msg}")


def dptr(a):
    """double* of a C-contiguous float64 ndarray (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def uptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_up)
