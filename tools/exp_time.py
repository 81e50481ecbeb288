import ctypes as C, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
for n in [int(a) for a in os.environ.get("N", "200").split(",")]:
    ecm = np.linspace(1.18, 2.0, n); st = np.ascontiguousarray([[0,0,0],[1,1,n-1]], dtype=np.int32); p = C.c_void_p()
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, 2, L.iptr(st), None, C.byref(p)))
    G = np.empty((n,n)); L.check(lib.isr_assemble(p, None, L.dptr(G))); G = G.T.copy()
    bcs0 = np.ones(n); vcs = G @ bcs0; err = 0.05*vcs
    L.check(lib.isr_factor(p, L.dptr(err)))
    T = int(os.environ.get("T", 1 << 20))
    Vd = torch.from_numpy(vcs)[None,:].cuda() + torch.from_numpy(err)[None,:].cuda()*torch.randn(T, n, dtype=torch.float64, device='cuda')
    chi2d = torch.empty(T, dtype=torch.float64, device='cuda')
    sbd = torch.zeros(n, dtype=torch.float64, device='cuda')
    sb = sbd.data_ptr() if os.environ.get('SUM') else None
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(2): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), sb, None, None, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), sb, None, None, None))
        e1.record(stream)
    e1.synchronize(); ms = e0.elapsed_time(e1)/5
    print(f"{os.environ.get('TAG','')} N={n}: {ms:.3f} ms {4*n*n*T/ms*1e-9:.2f} TFLOP/s mean chi2 {chi2d.mean().item():.3f}")
