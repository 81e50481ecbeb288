"""Campaign by count (toys drawn on the device): timing of isr_chi2_test with n toys requested by count, for ncu launch
lists of draw_toys_fast_kernel / toy_fused_kernel.  N=... T=... REPS=... in the environment."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isrsolver_b200 as isr  # noqa: E402

n = int(os.environ.get("N", 200))
T = int(os.environ.get("T", 1 << 20))
reps = int(os.environ.get("REPS", 5))
ecm = np.linspace(1.18, 2.0, n)
s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
s.eval_equation_matrix()
G = s.intop_matrix()
b0 = 1.0 + 0.3 * np.sin(5 * ecm)
vcs = G @ b0
err = 0.05 * vcs
s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
# the same test on resident toys (V_dev): the floor a campaign by count can reach
import ctypes as C
import torch
from isrsolver_b200 import _lib as L
lib = L.load()
Vd = torch.from_numpy(isr.draw_toys_device(vcs, err, 1, seed=1)).cuda().repeat(T, 1).contiguous()
a = L.Chi2TestArgs()
lohi, cnt, total = np.zeros(2), np.zeros(130, dtype=np.int64), C.c_int64()
errc, b0c = np.ascontiguousarray(err), np.ascontiguousarray(b0)
a.vcs_err, a.bcs0, a.n_toys, a.nbins, a.V_dev = L.dptr(errc), L.dptr(b0c), T, 128, Vd.data_ptr()
a.chi2_lo_hi_out, a.chi2_counts_out, a.total_toys_out = L.dptr(lohi), cnt.ctypes.data_as(C.POINTER(C.c_int64)), C.pointer(total)
s._select_for_toys()
for k in range(reps + 2):
    t0 = time.perf_counter()
    L.check(lib.isr_chi2_test(s._p, C.byref(a)))
    dt = time.perf_counter() - t0
print(f"resident toys: {dt * 1e3:.3f} ms  {T / dt:.4g} toys/s", flush=True)
for k in range(reps + 2):
    t0 = time.perf_counter()
    o = s.chi2_test_call(b0, n=T, vcs=vcs, vcs_err=err, seed=7 + k, want_chi2=False)
    dt = time.perf_counter() - t0
    print(f"rep {k}: {dt * 1e3:.3f} ms  {T / dt:.4g} toys/s  mean chi2 bin {np.argmax(o['hist'])}", flush=True)
