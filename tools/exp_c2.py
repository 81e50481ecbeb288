"""C2's toy phase (N = 100, dense, 10^5 toys, chi2 only) on the candidate tiles: wave quantisation against per-tile efficiency.
Usage: python tools/exp_c2.py [N] [T] [tile ...]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rng = np.random.default_rng(1)
ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
st = np.ascontiguousarray([[0, 0, n - 1]], dtype=np.int32)
L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, 1, L.iptr(st), None, C.byref(p)))
G = rng.uniform(-0.2, 0.2, (n, n)) + np.eye(n)
L.check(lib.isr_set_matrix(p, L.dptr(np.ascontiguousarray(G.T))))
err = np.ones(n); b0 = np.ones(n)
L.check(lib.isr_factor(p, L.dptr(err)))
V = torch.randn((T, n), dtype=torch.float64, device="cuda"); chi2 = torch.empty(T, dtype=torch.float64, device="cuda")
for tile in (sys.argv[3:] or [""]):
    if tile: os.environ["ISR_TOY_CFG"] = tile
    elif "ISR_TOY_CFG" in os.environ: del os.environ["ISR_TOY_CFG"]
    with torch.cuda.stream(stream):
        for _ in range(3): L.check(lib.isr_toy_batch_dev(p, T, V.data_ptr(), L.dptr(b0), chi2.data_ptr(), None, None, None, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): L.check(lib.isr_toy_batch_dev(p, T, V.data_ptr(), L.dptr(b0), chi2.data_ptr(), None, None, None, None))
        e1.record(stream)
    e1.synchronize(); ms = e0.elapsed_time(e1) / 10
    print(f"N={n} T={T} tile {tile or '(chooser)':22s} {ms*1e3:8.1f} us  {4.0*n*n*T/ms*1e-9:6.2f} TFLOP/s", flush=True)
