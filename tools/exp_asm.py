"""Assemble the bench matrix (N = 200, [[linear,0,0],[cubic,1,N-1]]) a few times: for ncu launch lists of the assembly kernels.
Usage: python tools/exp_asm.py [N] [linear|cubic]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kind = sys.argv[2] if len(sys.argv) > 2 else "cubic"
ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
st = np.ascontiguousarray([[0, 0, n - 1]] if kind == "linear" else [[0, 0, 0], [1, 1, n - 1]], dtype=np.int32)
L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, len(st), L.iptr(st), None, C.byref(p)))
with torch.cuda.stream(stream):
    for _ in range(3): L.check(lib.isr_assemble(p, None, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20): L.check(lib.isr_assemble(p, None, None))
    e1.record(stream)
e1.synchronize()
print(f"N={n} {kind}: assembly {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
