#!/usr/bin/env python3
"""Generate tests/golden/refrun_*.npz: outputs of the REFERENCE'S OWN solver classes and toy-test drivers.

oracle/_ref/libisr_ref.so holds src/BaseISRSolver{,2}.cpp, ISRSolverSLE{,2}.cpp, ISRSolverTikhonov.cpp, ISRSolverTSVD.cpp,
IterISRInterpSolver.cpp, Chi2Test.cpp, RatioTest.cpp and Utils.cpp compiled unmodified from /root/reference against oracle/ref_shim (dense Eigen
stand-in, ROOT containers over an in-memory registry; oracle/Makefile).  This script runs them on the inputs of the
existing fixtures (tests/golden/c*.npz: same points, errors, toys, model) and stores what they return under ref_* keys:

  injected   the solver is handed the fixture's operator matrix (G_converged / G_spread), so that only the linear-algebra
             layer differs from the restatement: solve(), save(), evalConditionNumber, chi2TestModel, chi2TestData,
             ratioTestModel, IterISRInterpSolver::solve (10 and 3 iterations); for c4 also ISRSolverTikhonov (12 lambda, both regularisers, the L-curve functionals,
             lambdaObjective, chi2TestModel with the Tikhonov covariance)
  native     the solver runs its own evalEqMatrix (N^2 adaptive integrations, the efficiency adaptor, the energy-spread
             product) and then solve(); for c4 also ISRSolverTSVD (whose SVD cannot be separated from evalEqMatrix)
  vcsfit     ISRSolverVCSFitFunction (src/ISRSolverVCSFitter.cpp): the chi2 objective of the visible-cross-section fit at four
             parameter sets (N forward convolutions each, with and without energy spread) and saveResults
  e2e        full-size runs with nothing injected: chi2TestModel + ratioTestModel at N = 200 cubic spline on 32 toys
             (the north-star geometry), N = 100 cubic spline on 64 toys (C2) and N = 500 linear on 16 toys (C5: a
             triangular operator), ISRSolverSLE2 at N = 200 with 512-bin efficiency histograms (C3), and N = 200 cubic spline with energy spread, ISRSolverTikhonov at 4 lambda
             with chi2TestModel on 16 toys each (configs[3] geometry).

tests/test_ref_pin.py checks that oracle/isr_oracle.py reproduces these numbers (the pin of its linear-algebra layer) and
that a fresh run reproduces the files; tests/test_gpu_refrun.py compares the CUDA path with them through the C ABI.
Run in the container that has /root/reference:  python tests/golden/make_refrun_golden.py   (about ten minutes, most of
it the N = 500 run: 125 000 adaptive integrations and an O(N^3) solve per toy on one core).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import isr_oracle as O  # noqa: E402
from oracle import isr_ref as R     # noqa: E402

TIK_COV_AT = (0, 3, 8)          # lambda indices whose covariance is stored
TIK_CHI2_AT = (3, 9)            # lambda indices with a chi2TestModel run
TIK_NOREG_AT = (3, 6, 9)        # lambda indices also solved with the L2-norm regulariser


def load(name):
    with np.load(os.path.join(HERE, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def fixture_matrix(g):
    return g["G_spread"] if "G_spread" in g else g["G_converged"]


def run_v1(g):
    """Everything the v1 classes give for one small fixture."""
    thr, st = float(g["threshold"]), g["settings"]
    ee = g.get("ecm_err")
    G = fixture_matrix(g)
    out = {}
    S = R.RefSolver("sle", g["ecm"], g["vcs"], g["vcs_err"], thr, ecm_err=ee, settings=st)
    S.inject_matrix(G)
    b, cov = S.solve()
    out.update(ref_bcs=b, ref_cov=cov, ref_cond=S.cond())
    sv = S.save()
    out.update(ref_bcs_err=sv["bcs_err"], ref_inv_cov=sv["inv_cov"])
    # interpEval (src/ISRSolverSLE.cpp:195-198 -> Interpolator::eval): below threshold, inside every range, between nodes, beyond E_max
    ei = np.concatenate([[thr - 0.01, thr, thr + 1e-3], np.linspace(thr, g["ecm"][-1], 41)[1:-1] + 1e-4, g["ecm"][[0, 5, -1]], [g["ecm"][-1] + 0.3]])
    out.update(ref_interp_E=ei, ref_interp_val=np.array([S.interp_eval(b, float(e)) for e in ei]))
    assert np.array_equal(sv["bcs"], b) and np.array_equal(sv["cov"], cov) and np.array_equal(sv["G"], G)
    chi2, counts, lo, hi = S.chi2_test_model(g["toys"], g["vcs"], g["bcs0"])
    out.update(ref_chi2=chi2, ref_chi2_hist=counts, ref_chi2_lo=lo, ref_chi2_hi=hi)
    means, errs, rc, stats = S.ratio_test_model(g["toys"], g["vcs"], g["bcs0"])
    out.update(ref_ratio_means=means, ref_ratio_mean_errors=errs, ref_ratio_counts=rc.astype(np.uint16), ref_ratio_stats=stats)
    # chi2TestData: the solver's own visible cross section is the centre of the toys; first half -> mean solution
    S.reset_vcs(g["vcs"])
    chi2d, cd, lod, hid = S.chi2_test_data(g["toys"])
    out.update(ref_chi2_data=chi2d, ref_chi2_data_hist=cd, ref_chi2_data_lo=lod, ref_chi2_data_hi=hid)
    # IterISRInterpSolver on the same operator: default 10 iterations, and 3
    It = R.RefSolver("iterative", g["ecm"], g["vcs"], g["vcs_err"], thr, ecm_err=ee, settings=st)
    It.inject_matrix(G)
    for tag, n_iter in (("", 10), ("3", 3)):
        It.iter_set(n_iter)
        bi, ci = It.solve()
        assert np.count_nonzero(ci - np.diag(np.diag(ci))) == 0
        out.update({"ref_iter%s_bcs" % tag: bi, "ref_iter%s_radcorr" % tag: It.radcorr(), "ref_iter%s_cov_diag" % tag: np.diag(ci).copy()})
    # native: both constructors, nothing injected
    for ctor in ("graph", "pointers"):
        Sn = R.RefSolver("sle", g["ecm"], g["vcs"], g["vcs_err"], thr, ecm_err=ee, settings=st, ctor=ctor)
        bn, cn = Sn.solve()
        if ctor == "graph":
            out.update(ref_G_native=Sn.matrix(), ref_bcs_native=bn, ref_cov_native=cn)
        else:
            assert np.array_equal(Sn.matrix(), out["ref_G_native"]) and np.array_equal(bn, out["ref_bcs_native"])
    Gn = out["ref_G_native"]
    assert np.array_equal(Gn, g["G_ref"]) if ee is None else np.allclose(Gn, g["S_ref"] @ g["G_ref"], rtol=1e-13, atol=1e-17)
    return out


def run_tikhonov(g):
    thr, st, ee = float(g["threshold"]), g["settings"], g.get("ecm_err")
    G = fixture_matrix(g)
    lam = g["tik_lambda"]
    S = R.RefSolver("tikhonov", g["ecm"], g["vcs"], g["vcs_err"], thr, ecm_err=ee, settings=st)
    S.inject_matrix(G)
    D, w = S.tik_operators()
    assert np.array_equal(D, g["D_ref"]) and np.array_equal(w, g["w_ref"])
    rows, covs, noreg, chi2s, obj = [], [], [], [], []
    for k, l in enumerate(lam):
        S.tik_set(l, True)
        b, cov = S.solve()
        f = S.tik_functionals()
        rows.append((b, f["x"], f["y"], f["curv"], f["dcurv"]))
        if k in TIK_COV_AT:
            covs.append(cov)
        if k in TIK_CHI2_AT:
            chi2s.append(S.chi2_test_model(g["toys"], g["vcs"], g["bcs0"])[0])
            S.reset_vcs(g["vcs"])
        if k in TIK_NOREG_AT:
            S.tik_set(l, False)
            bn, cn = S.solve()
            fn = S.tik_functionals()
            noreg.append((bn, fn["x"], fn["y"], fn["curv"], fn["dcurv"]))
    S.tik_set(1.0, True)
    for l in lam[[2, 7]]:
        obj.append(S.lambda_objective(l))
    out = dict(ref_tik_bcs=np.array([r[0] for r in rows]), ref_tik_x=np.array([r[1] for r in rows]), ref_tik_y=np.array([r[2] for r in rows]),
               ref_tik_curv=np.array([r[3] for r in rows]), ref_tik_dcurv=np.array([r[4] for r in rows]),
               ref_tik_cov=np.array(covs), ref_tik_cov_at=np.array(TIK_COV_AT), ref_tik_chi2=np.array(chi2s), ref_tik_chi2_at=np.array(TIK_CHI2_AT),
               ref_tik_noreg_at=np.array(TIK_NOREG_AT), ref_tik_noreg_bcs=np.array([r[0] for r in noreg]),
               ref_tik_noreg_xycd=np.array([r[1:] for r in noreg]), ref_lambda_objective=np.array(obj), ref_lambda_objective_at=np.array([2, 7]))
    # condition number versus beam-energy spread (src/isrsolver-condnum-test.cpp:84-128), the reference's own re-assembly per setting
    sig = np.array([0.0, 0.0005, 0.001, 0.002, 0.004, 0.008, 0.016])
    Sc = R.RefSolver("sle", g["ecm"], g["vcs"], g["vcs_err"], thr, settings=st)
    out.update(ref_condscan_sigma=sig, ref_condscan_cond=Sc.cond_spread_scan(sig))
    # TSVD: native only
    k = int(g["tsvd_k"])
    n = len(g["ecm"])
    T = R.RefSolver("tsvd", g["ecm"], g["vcs"], g["vcs_err"], thr, ecm_err=ee, settings=st, upper=k)
    bk, _ = T.solve()
    T.tsvd_set(k, True)
    b1, _ = T.solve()
    T.tsvd_set(n, False)
    bf, cf = T.solve()
    out.update(ref_tsvd_G=T.matrix(), ref_tsvd_bcs=bk, ref_tsvd_bcs_keepone=b1, ref_tsvd_bcs_full=bf, ref_tsvd_cov_full=cf)
    return out


def run_v2(g):
    eff = O.Efficiency("row_table", values=g["eff_values"], nx=int(g["eff_nx"]), xlo=float(g["eff_xlo"]), xhi=float(g["eff_xhi"]))
    S = R.RefSolver2(g["ecm"], g["vcs"], g["vcs_err"], float(g["threshold"]), eff=eff, settings=g["settings"])
    b, cov = S.solve()
    assert np.array_equal(S.matrix(), g["G_ref"])
    return dict(ref_G_native=S.matrix(), ref_bcs_native=b, ref_cov_native=cov, ref_cond_native=S.cond())


def e2e_inputs(n, toys, spread, cubic=True):
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    vcs = O.bw_visible(ecm)
    vcs_err = 0.05 * vcs
    settings = [(0, 0, 0), (1, 1, n - 1)] if cubic else [(0, 0, n - 1)]
    out = dict(ecm=ecm, threshold=O.E_THR, settings=np.array(settings, dtype=np.int32), vcs=vcs, vcs_err=vcs_err,
               bcs0=O.bw_born(ecm), toys=O.draw_toys(vcs, vcs_err, toys))
    if spread:
        out["ecm_err"] = np.full(n, 0.002)
    return out


def run_e2e_sle(n=200, toys=32, cubic=True):
    g = e2e_inputs(n, toys, False, cubic)
    S = R.RefSolver("sle", g["ecm"], g["vcs"], g["vcs_err"], float(g["threshold"]), settings=g["settings"])
    chi2, counts, lo, hi = S.chi2_test_model(g["toys"], g["vcs"], g["bcs0"])      # runs evalEqMatrix + solve() itself
    S.reset_vcs(g["vcs"])
    b, cov = S.solve()
    means, errs, rc, stats = S.ratio_test_model(g["toys"], g["vcs"], g["bcs0"])
    rows = np.array([0, n // 2, n - 1])
    g.update(ref_bcs=b, ref_cov_diag=np.diag(cov).copy(), ref_cov_rows_at=rows, ref_cov_rows=cov[rows], ref_cond=S.cond(), ref_chi2=chi2,
             ref_chi2_hist=counts, ref_chi2_lo=lo, ref_chi2_hi=hi, ref_ratio_means=means, ref_ratio_mean_errors=errs,
             ref_ratio_counts=rc.astype(np.uint16), ref_ratio_stats=stats, ref_G_rows=S.matrix()[rows])
    return g


def run_e2e_sle2(n=200, nbins=512):
    """BASELINE configs C3 geometry: ISRSolverSLE2, linear interpolation, one 512-bin efficiency histogram per energy point.
    The reference's adaptive quadrature meets the bin edges inside its integration ranges and its retry loop
    (src/Integration.cpp:28-42) ends at epsrel up to 1e-3 there: G_err_rows carries the error estimates it returned."""
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    eff = O.row_table_efficiency(ecm, nbins)
    vcs = O.bw_visible(ecm, eff=eff)
    vcs_err = 0.05 * vcs
    S = R.RefSolver2(ecm, vcs, vcs_err, O.E_THR, eff=eff)
    b, cov = S.solve()
    rows = np.array([0, n // 2, n - 1])
    I = O.Interp(ecm, O.E_THR)
    Gf, Ef, Rf = I.eval_eq_matrix(eff=eff, with_errors=True)
    assert np.array_equal(Gf, S.matrix()), "the faithful restatement differs from the compiled ISRSolverSLE2"
    return dict(ecm=ecm, threshold=O.E_THR, settings=np.array([(0, 0, n - 1)], dtype=np.int32), eff_nbins=nbins, vcs=vcs, vcs_err=vcs_err,
                ref_bcs=b, ref_cov_diag=np.diag(cov).copy(), ref_cond=S.cond(), ref_G_rows_at=rows, ref_G_rows=S.matrix()[rows],
                ref_G_err_rows=Ef[rows], ref_G_relerr_max=Rf.max())


def run_e2e_tikhonov(n=200, toys=16, lam=(1e-6, 1e-4, 1e-2, 1.0)):
    g = e2e_inputs(n, toys, True)
    S = R.RefSolver("tikhonov", g["ecm"], g["vcs"], g["vcs_err"], float(g["threshold"]), ecm_err=g["ecm_err"], settings=g["settings"])
    bs, fs, chi2s = [], [], []
    for l in lam:
        S.tik_set(l, True)
        S.reset_vcs(g["vcs"])
        b, _ = S.solve()
        f = S.tik_functionals()
        bs.append(b)
        fs.append((f["x"], f["y"], f["curv"], f["dcurv"]))
        chi2s.append(S.chi2_test_model(g["toys"], g["vcs"], g["bcs0"])[0])
    rows = np.array([0, n // 2, n - 1])
    g.update(tik_lambda=np.array(lam), ref_tik_bcs=np.array(bs), ref_tik_xycd=np.array(fs), ref_tik_chi2=np.array(chi2s),
             ref_G_rows_at=rows, ref_G_rows=S.matrix()[rows])
    return g


VCSFIT_PARS = np.array([[4.2, 1.52, 0.26], [3.7, 1.47, 0.22], [4.0, 1.50, 0.30], [4.4, 1.55, 0.25]])


def bw_model(e, par, thr=O.E_THR):
    """The parametrised Born cross section of the fit: relativistic Breit-Wigner x threshold factor, par = (sigma0, M, Gamma)."""
    s0, M, Gm = par
    if e <= thr:
        return 0.0
    return s0 * M * M * Gm * Gm / ((e * e - M * M) ** 2 + M * M * Gm * Gm) * (1.0 - thr * thr / (e * e)) ** 1.5


def run_vcsfit(n=24):
    """ISRSolverVCSFitFunction of the reference: chi2 at four parameter sets away from the truth, without and with a 2 MeV
    energy spread, and saveResults (blurred model, radiative correction, Born points) at the first set."""
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    vcs = O.bw_visible(ecm)
    vcs_err = 0.05 * vcs
    ee = np.full(n, 0.002)
    out = dict(ecm=ecm, threshold=O.E_THR, vcs=vcs, vcs_err=vcs_err, ecm_err=ee, pars=VCSFIT_PARS)
    F = R.RefVCSFitFunction(ecm, vcs, vcs_err, O.E_THR, bw_model)
    out["ref_chi2"] = np.array([F.chi2(p) for p in VCSFIT_PARS])
    out["ref_chi2_true"] = F.chi2([O.BW_S0, O.BW_M, O.BW_G])
    sv = F.save_results(VCSFIT_PARS[0])
    out.update({"ref_save_" + k: v for k, v in sv.items()})
    Fs = R.RefVCSFitFunction(ecm, vcs, vcs_err, O.E_THR, bw_model, ecm_err=ee)
    out["ref_chi2_spread"] = np.array([Fs.chi2(p) for p in VCSFIT_PARS])
    sv = Fs.save_results(VCSFIT_PARS[0])
    out.update({"ref_save_spread_" + k: v for k, v in sv.items()})
    return out


def save(name, out):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written:", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "kB", flush=True)


SMALL = {"c1_sle_n50_linear": (run_v1,), "c2_sle_n40_cspline": (run_v1,), "c3_sle2_n24_eps64": (run_v2,),
         "c4_tikhonov_n32_spread": (run_v1, run_tikhonov), "c5_mixed3_n36": (run_v1,)}


def small_case(name):
    g = load(name)
    out = {}
    for f in SMALL[name]:
        out.update(f(g))
    return out


if __name__ == "__main__":
    assert R.build(), "oracle/_ref/libisr_ref.so cannot be built here (no /root/reference)"
    t0 = time.time()
    for name in SMALL:
        save("refrun_" + name, small_case(name))
    save("refrun_vcsfit_n24", run_vcsfit())
    save("refrun_e2e_sle_n200_cspline", run_e2e_sle())
    save("refrun_e2e_sle_n100_cspline", run_e2e_sle(100, 64))                  # BASELINE configs C2 geometry
    save("refrun_e2e_sle_n500_linear", run_e2e_sle(500, 16, cubic=False))      # C5 geometry: triangular operator, ratio + chi2
    save("refrun_e2e_sle2_n200_eps512", run_e2e_sle2())                        # C3 geometry
    save("refrun_e2e_tikhonov_n200_spread", run_e2e_tikhonov())
    print("done in %.0f s" % (time.time() - t0))
