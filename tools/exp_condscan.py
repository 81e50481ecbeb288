"""isr_cond_spread_scan, N = 200, 100 settings (for ncu captures of the batched Jacobi kernel and for timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import isrsolver_b200 as isr
n = 200
ecm = np.linspace(1.18, 2.0, n)
s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
sig = np.linspace(0.0, 0.005, 100)
for _ in range(2):
    t0 = time.perf_counter(); cond = s.condnum_spread_scan(sig); dt = time.perf_counter() - t0
print(f"100 settings: {dt*1e3:.1f} ms; cond[0] {cond[0]:.4f} cond[50] {cond[50]:.2f}")
t0 = time.perf_counter(); cond = s.condnum_spread_scan(sig[:1]); dt = time.perf_counter() - t0
print(f"1 setting: {dt*1e3:.1f} ms")
