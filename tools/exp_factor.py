"""Timing and accuracy of isr_factor (inverse + certificate + operator set-up) at several N, cubic and linear.
Usage: python tools/exp_factor.py [N ...]"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
for n in [int(a) for a in sys.argv[1:]] or [7, 50, 100, 200, 256, 257, 500]:
  for kind in ("linear", "cubic"):
    ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
    st = np.ascontiguousarray([[0, 0, n - 1]] if kind == "linear" else [[0, 0, 0], [1, 1, n - 1]], dtype=np.int32)
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, len(st), L.iptr(st), None, C.byref(p)))
    G = np.empty((n, n)); L.check(lib.isr_assemble(p, None, L.dptr(G))); G = G.T.copy()
    bcs0 = np.ones(n); vcs = G @ bcs0; err = 0.05 * vcs
    with torch.cuda.stream(stream):
        for _ in range(3): L.check(lib.isr_factor(p, L.dptr(err)))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20): L.check(lib.isr_factor(p, L.dptr(err)))
        e1.record(stream)
    e1.synchronize(); ms = e0.elapsed_time(e1) / 20
    b = np.empty(n); cov = np.empty((n, n))
    L.check(lib.isr_solve(p, L.dptr(vcs), L.dptr(b), L.dptr(cov)))
    print(f"N={n} {kind}: factor {ms*1e3:.1f} us   max |b - 1| = {np.abs(b - 1).max():.2e}  cond {np.linalg.cond(G):.2f}", flush=True)
    lib.isr_problem_destroy(p)
