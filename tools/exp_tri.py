"""A/B timing of the toy kernel's triangular schedule (isr_set_toy_kernel 0) against the dense schedule on the operators
as given (variant 1): lower-triangular operators of linear interpolation, and dense (cubic) operators whose chi2 factor
is replaced by its QL factor.  Usage: python tools/exp_tri.py [N ...]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
for n in [int(a) for a in sys.argv[1:]] or [100, 200, 500]:
  for kind in ("linear", "cubic"):
    ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
    st = np.ascontiguousarray([[0, 0, n - 1]] if kind == "linear" else [[0, 0, 0], [1, 1, n - 1]], dtype=np.int32)
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, len(st), L.iptr(st), None, C.byref(p)))
    G = np.empty((n, n)); L.check(lib.isr_assemble(p, None, L.dptr(G))); G = G.T.copy()
    bcs0 = np.ones(n); vcs = G @ bcs0; err = 0.05 * vcs
    T = 1 << 19
    Vd = torch.from_numpy(vcs)[None, :].cuda() + torch.from_numpy(err)[None, :].cuda() * torch.randn(T, n, dtype=torch.float64, device='cuda')
    res = {}
    for tag, variant in (("auto", 0), ("dense", 1), ("ql", 2)):
        L.check(lib.isr_set_toy_kernel(p, variant))
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            chi2d = torch.empty(T, dtype=torch.float64, device='cuda')
            L.check(lib.isr_factor(p, L.dptr(err)))
            L.check(lib.isr_toy_batch_dev(p, 64, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), None, None, None, None))
            torch.cuda.synchronize()
            f0.record(stream)
            L.check(lib.isr_factor(p, L.dptr(err)))
            L.check(lib.isr_toy_batch_dev(p, 64, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), None, None, None, None))
            f1.record(stream)
            for _ in range(2): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), None, None, None, None))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), None, None, None, None))
            e1.record(stream)
        e1.synchronize(); ms = e0.elapsed_time(e1) / 5
        res[tag] = chi2d.clone()
        print(f"N={n} {kind} {tag}: {ms:.3f} ms  {T/ms*1e3:.3e} toys/s  dense-equivalent {4*n*n*T/ms*1e-9:.2f} TFLOP/s  factor + 64 toys {f0.elapsed_time(f1):.3f} ms  mean chi2 {chi2d.mean().item():.4f}", flush=True)
    for tag in ("auto", "ql"):
        d = ((res[tag] - res["dense"]).abs() / res["dense"]).max().item()
        print(f"N={n} {kind} max rel |chi2 {tag} - dense| = {d:.3e}")
        assert d <= 1e-11
    lib.isr_problem_destroy(p)
