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
for k in range(reps + 2):
    t0 = time.perf_counter()
    o = s.chi2_test_call(b0, n=T, vcs=vcs, vcs_err=err, seed=7 + k, want_chi2=False)
    dt = time.perf_counter() - t0
    print(f"rep {k}: {dt * 1e3:.3f} ms  {T / dt:.4g} toys/s  mean chi2 bin {np.argmax(o['hist'])}", flush=True)
