"""TFLOP/s of the fused toy kernel (chi2 only, dense schedule) over N = 96 ... 304 step 8 (plus 500): shows whether the
tile / chunk chooser (pick_cfg, toy_batch.cu) has a cliff between the sizes it was fitted at.
Usage: python tools/exp_nsweep.py > profiles/toy_nsweep_r02.json"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
T = 1 << 19
rows = []
rng = np.random.default_rng(1)
for n in list(range(96, 305, 8)) + [100, 150, 250, 500]:
    ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
    st = np.ascontiguousarray([[0, 0, n - 1]], dtype=np.int32)
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, 1, L.iptr(st), None, C.byref(p)))
    G = rng.uniform(-0.2, 0.2, (n, n)) + np.eye(n)             # dense operator: the dense schedule whatever N
    L.check(lib.isr_set_matrix(p, L.dptr(np.ascontiguousarray(G.T))))
    err = np.ones(n); b0 = np.ones(n)
    L.check(lib.isr_factor(p, L.dptr(err)))
    V = torch.randn((T, n), dtype=torch.float64, device="cuda")
    chi2 = torch.empty(T, dtype=torch.float64, device="cuda")
    with torch.cuda.stream(stream):
        for _ in range(2):
            L.check(lib.isr_toy_batch_dev(p, T, V.data_ptr(), L.dptr(b0), chi2.data_ptr(), None, None, None, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            L.check(lib.isr_toy_batch_dev(p, T, V.data_ptr(), L.dptr(b0), chi2.data_ptr(), None, None, None, None))
        e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tf = 4.0 * n * n * T / (ms * 1e-3) * 1e-12
    rows.append({"n": n, "tile": (lib.isr_toy_tile_name(n, 0) or b"?").decode(), "ms": ms, "tflops": tf, "frac_of_37.1": tf / 37.1})
    print(f"N={n:4d} {rows[-1]['tile']:22s} {ms:8.3f} ms  {tf:6.2f} TFLOP/s", file=sys.stderr, flush=True)
    del V, chi2
    lib.isr_problem_destroy(p)
rows.sort(key=lambda r: r["n"])
print(json.dumps({"toys": T, "what": "toy_fused_kernel, chi2 only (pass 1 + pass 2), dense operators, CUDA events, mean of 5", "rows": rows}))
