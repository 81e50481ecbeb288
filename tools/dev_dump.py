import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from isrsolver_b200 import _lib as L
lib = C.CDLL(L.LIB_PATH)
for name in ["isr_ctx_create","isr_problem_create","isr_assemble","isr_assemble_stats","isr_last_error","isr_debug_moments","isr_debug_panels"]:
    fn = getattr(lib, name); fn.restype, fn.argtypes = L.SIGNATURES[name]
ctx = C.c_void_p(); lib.isr_ctx_create(0, C.byref(ctx))
out = {}
for n in (20, 50):
    ecm = np.linspace(1.18, 2.0, n)
    for nm, sts in (("linear", [(0,0,n-1)]), ("cspline", [(1,0,n-1)])):
        st = np.ascontiguousarray(sts, dtype=np.int32); p = C.c_void_p()
        assert lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, len(st), L.iptr(st), None, C.byref(p)) == 0
        G = np.empty((n, n)); assert lib.isr_assemble(p, None, L.dptr(G)) == 0
        M = np.empty((n, n, 4)); assert lib.isr_debug_moments(p, L.dptr(M)) == 0
        cnt = C.c_int64(); lib.isr_debug_panels(p, 0, None, None, C.byref(cnt))
        ab = np.empty((cnt.value, 4)); rs = np.empty((cnt.value, 2), dtype=np.int32)
        lib.isr_debug_panels(p, cnt.value, L.dptr(ab), L.iptr(rs), C.byref(cnt))
        out[f"G_{n}_{nm}"] = G.T.copy(); out[f"M_{n}_{nm}"] = M; out[f"ab_{n}_{nm}"] = ab; out[f"rs_{n}_{nm}"] = rs
np.savez("gpurun_out/dump.npz", **out)
