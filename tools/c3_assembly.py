"""C3 (N = 200, 512-bin efficiency histogram per point) assembly only, a few repetitions: launch-list / ncu target."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isrsolver_b200 as isr
n = int(os.environ.get("ISR_N", "200")); nbins = int(os.environ.get("ISR_NBINS", "512"))
ecm = np.linspace(1.18, 2.0, n)
xc = (np.arange(max(nbins, 1)) + 0.5) / max(nbins, 1)
vals = np.zeros((n, max(nbins, 1) + 2))
vals[:, 1:-1] = 0.25 + 0.5 * np.exp(-40.0 * xc)[None, :] * (1.0 - 0.1 * (ecm[:, None] - 1.5))
if nbins:
    s = isr.ISRSolverSLE2(n, ecm, np.ones(n), None, np.ones(n), 0.827, eff_hists=vals, xlo=0.0, xhi=1.0)
else:
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
if os.environ.get("ISR_CUBIC"):
    s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
for _ in range(int(os.environ.get("ISR_REPS", "4"))):
    s.eval_equation_matrix()
print("ok", s.intop_matrix()[5, 3])
