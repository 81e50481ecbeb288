"""Where C5's toy time goes (N = 500, linear, triangular schedule): chi2 only / + sum of solutions / + ratio statistics
/ + ratio cells, and the tile shapes they run on.  Usage: python tools/exp_c5.py [N] [tile ...]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
T = 1 << 18
ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
st = np.ascontiguousarray([[0, 0, n - 1]], dtype=np.int32)
L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, 1, L.iptr(st), None, C.byref(p)))
G = np.empty((n, n)); L.check(lib.isr_assemble(p, None, L.dptr(G))); G = G.T.copy()
bcs0 = 1.0 + np.arange(n) / n; vcs = G @ bcs0; err = 0.05 * vcs
L.check(lib.isr_factor(p, L.dptr(err)))
Vd = torch.from_numpy(vcs)[None, :].cuda() + torch.from_numpy(err)[None, :].cuda() * torch.randn(T, n, dtype=torch.float64, device='cuda')
chi2 = torch.empty(T, dtype=torch.float64, device='cuda'); acc = torch.zeros(4 * n, dtype=torch.float64, device='cuda')
rh = torch.zeros(n * 1026, dtype=torch.int32, device='cuda')
def run(tag, sumb, ratio, hist, variant=0):
    L.check(lib.isr_set_toy_kernel(p, variant))
    args = (chi2.data_ptr(), acc.data_ptr() if sumb else None, acc[n:].data_ptr() if ratio else None, rh.data_ptr() if hist else None, None)
    with torch.cuda.stream(stream):
        for _ in range(2): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), *args))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(4): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), *args))
        e1.record(stream)
    e1.synchronize(); ms = e0.elapsed_time(e1) / 4
    nstat = 4 if (ratio or hist) else (1 if sumb else 0)
    print(f"N={n} {tag:28s} tile {lib.isr_toy_tile_name(n, nstat).decode():20s} {ms:7.3f} ms  algorithmic(2N^2) {2*n*n*T/ms*1e-9:6.2f} TFLOP/s = {2*n*n*T/ms*1e-9/37.1:.3f}", flush=True)
for tile in (sys.argv[2:] or [None]):
    if tile: os.environ["ISR_TOY_CFG"] = tile
    print("forced tile:", tile)
    run("chi2 only, triangular", False, False, False)
    run("+ sum of solutions", True, False, False)
    run("+ ratio statistics", True, True, False)
    run("+ ratio cells", True, True, True)
    run("chi2 only, dense schedule", False, False, False, 1)
