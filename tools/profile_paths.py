#!/usr/bin/env python3
"""Exercise the non-headline kernels at BASELINE.json sizes and print one JSON line of timings:
C1 (N=50 linear) / bench (N=200 cubic) / C3 (N=200, eps 512 x N) assembly, C4 (1024 lambdas x 10k toys)
Tikhonov scan, the forward convolution service and the condition-number scan.  Run plain for timings and
under `ncu --set full -k regex:...` for the profiles summarised in profiles/ (tools/ncu_summary.py).
Times are wall clock around the synchronous C-ABI calls (each ends with a stream synchronize)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isrsolver_b200 as isr  # noqa: E402
from isrsolver_b200 import _lib as L  # noqa: E402
from bench import bw_born  # noqa: E402

THR = 0.827
REP = int(os.environ.get("ISR_PROFILE_REPS", "20"))


def timed(fn, reps=REP):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def stats(s):
    a, b = C.c_int64(), C.c_int64()
    L.check(L.load().isr_assemble_stats(s._p, C.byref(a), C.byref(b)))
    return a.value, b.value


def assembly_case(n, settings, nbins=None, spread=False):
    ecm = np.linspace(1.18, 2.0, n)
    if nbins:
        xc = (np.arange(nbins) + 0.5) / nbins
        vals = np.zeros((n, nbins + 2))
        vals[:, 1:-1] = 0.25 + 0.5 * np.exp(-40.0 * xc)[None, :] * (1.0 - 0.1 * (ecm[:, None] - 1.5))
        s = isr.ISRSolverSLE2(n, ecm, np.ones(n), np.full(n, 0.002), np.ones(n), THR, eff_hists=vals, xlo=0.0, xhi=1.0,
                              enableEnergySpread=spread)
    else:
        s = isr.ISRSolverSLE(n, ecm, np.ones(n), np.full(n, 0.002), np.ones(n), THR, enableEnergySpread=spread)
    if settings:
        s.set_interp_settings(settings)
    lib = L.load()
    err = L.dptr(s._ecm_err) if spread else None
    ms = timed(lambda: L.check(lib.isr_assemble(s._p, err, None)))      # device-resident: no D2H of G
    pan, nodes = stats(s)
    return {"n": n, "ms": ms, "panels": pan, "nodes": nodes, "nodes_per_s": nodes / (ms * 1e-3)}


def main():
    out = {}
    out["assembly_c1_n50_linear"] = assembly_case(50, None)
    out["assembly_c2_n100_cubic"] = assembly_case(100, [(False, 0, 0), (True, 1, 99)])
    out["assembly_bench_n200_cubic"] = assembly_case(200, [(False, 0, 0), (True, 1, 199)])
    out["assembly_c3_n200_eps512"] = assembly_case(200, None, nbins=512)
    out["assembly_c4_n200_cubic_spread"] = assembly_case(200, [(False, 0, 0), (True, 1, 199)], spread=True)
    out["assembly_c5_n500_linear"] = assembly_case(500, None)

    # C4: Tikhonov lambda x toy scan
    n, T, Lm = 200, 10000, 1024
    ecm = np.linspace(1.18, 2.0, n)
    s = isr.ISRSolverTikhonov(n, ecm, np.ones(n), np.full(n, 0.002), np.ones(n), THR, enableEnergySpread=True)
    s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    s.eval_equation_matrix()
    G = s.intop_matrix().copy()
    b0 = bw_born(ecm); vcs = G @ b0; err = 0.05 * vcs
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    lam = 1e-9 * (1e9) ** (np.arange(Lm) / (Lm - 1))
    V = isr.draw_toys(vcs, err, T)
    t_prep = timed(lambda: (setattr(s, "_tik_ready", False), s._ensure_tik()), reps=5)
    t_scan = timed(lambda: s.toy_scan(lam, V, b0), reps=5)
    t_hist = timed(lambda: s.toy_scan_hist(lam, V, b0), reps=5)
    t_lc = timed(lambda: s.make_LCurve(lam), reps=5)
    out["c4_tikhonov"] = {"n": n, "lambdas": Lm, "toys": T, "prepare_ms": t_prep, "toy_scan_host_api_ms": t_scan, "toy_scan_hist_host_api_ms": t_hist,
                          "pairs_per_s": Lm * T / (t_scan * 1e-3), "lcurve_1024_ms": t_lc}

    # forward convolution: 1000 energies, Breit-Wigner
    ens = np.linspace(0.9, 3.0, 1000)
    bw = lambda e: bw_born(e)
    plan = isr.ConvolutionPlan(ens, threshold=THR)
    t0 = time.perf_counter(); res, er = plan.convolve(bw); t_first = (time.perf_counter() - t0) * 1e3
    t_again = timed(lambda: plan.convolve(bw, refine=False), reps=10)
    out["convolution_1000_energies"] = {"adaptive_ms": t_first, "fixed_grid_ms": t_again, "panels": plan.size()[0],
                                        "max_rel_err_est": float((er / np.abs(res)).max())}
    # condition number vs spread, 100 points (notebooks/condnum_enspread.ipynb cell 9)
    s2 = isr.ISRSolverSLE(n, ecm, vcs, None, err, THR)
    s2.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    sig = np.linspace(0.0, 0.01, 100)
    t0 = time.perf_counter(); cond = s2.condnum_spread_scan(sig); t_c = (time.perf_counter() - t0) * 1e3
    out["condnum_scan_100_sigmas"] = {"n": n, "ms": t_c, "cond_first": float(cond[0]), "cond_last": float(cond[-1])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
