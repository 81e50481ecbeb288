"""Pin the linear-algebra layer of the oracle (oracle/isr_oracle.py: solve, covariance, chi2 / ratio toy tests, Tikhonov,
TSVD, condition number) against OUTPUTS OF THE REFERENCE ITSELF (no GPU).

tests/golden/refrun_*.npz were produced by tests/golden/make_refrun_golden.py from the reference's own solver classes and
toy-test drivers -- src/BaseISRSolver{,2}.cpp, ISRSolverSLE{,2}.cpp, ISRSolverTikhonov.cpp, ISRSolverTSVD.cpp, Chi2Test.cpp,
RatioTest.cpp, Utils.cpp compiled unmodified against oracle/ref_shim -- on the inputs of tests/golden/c*.npz (same
points, errors, toys, model).  Here the NumPy restatement runs on those inputs and must reproduce the reference's numbers:
1e-12 relative where only rounding differs (the decompositions are different implementations of the same factorisations),
bit-exact for histogram contents and ranges derived from identical chi2 values.  The second half needs
oracle/_ref/libisr_ref.so and checks that a fresh reference run reproduces the committed files bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden
from oracle import isr_oracle as O
from oracle import isr_ref as R

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
V1 = ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c4_tikhonov_n32_spread", "c5_mixed3_n36"]
RT = 1e-12


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


def fixture_matrix(g):
    return g["G_spread"] if "G_spread" in g else g["G_converged"]


@pytest.mark.parametrize("name", V1)
def test_sle_solve_covariance_condition_number(name):
    """ISRSolverSLE::solve (COD solve + (G^T W G)^-1), save()'s bcs errors / inverse covariance, evalConditionNumber."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    G = fixture_matrix(g)
    b, cov = O.sle_solve(G, g["vcs"], g["vcs_err"])
    assert rel(b, r["ref_bcs"]) < RT and rel(cov, r["ref_cov"]) < RT
    assert rel(np.sqrt(np.diag(cov)), r["ref_bcs_err"]) < RT
    assert rel(np.linalg.inv(cov), r["ref_inv_cov"]) < 1e-11          # save() inverts the covariance once more
    assert abs(O.cond_number(G) / float(r["ref_cond"]) - 1.0) < RT
    I = O.Interp(g["ecm"], float(g["threshold"]), [tuple(x) for x in g["settings"]])
    np.testing.assert_array_equal([I.eval(r["ref_bcs"], float(e)) for e in r["ref_interp_E"]], r["ref_interp_val"])   # interpEval
    # the fixtures the GPU tests use carry the restatement's numbers: the same
    assert rel(g["bcs"], r["ref_bcs"]) < RT and rel(g["cov"], r["ref_cov"]) < RT and abs(g["cond"] / r["ref_cond"] - 1) < RT


@pytest.mark.parametrize("name", V1)
def test_chi2_test_model(name):
    """chi2TestModel -> chi2Test (src/Chi2Test.cpp:29-101): per-toy solve(), chi2 through FullPivLU of the covariance, range,
    TH1F.  The restatement's literal loop and its factor-once form give the reference's chi2; from the reference's chi2
    values the restated TH1F arithmetic gives the reference's histogram bit for bit."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    G = fixture_matrix(g)
    loop, _ = O.chi2_test_loop(G, g["toys"], g["bcs0"], g["vcs_err"])
    fair, _ = O.chi2_test_fair(G, g["toys"], g["bcs0"], g["vcs_err"])
    assert rel(loop, r["ref_chi2"]) < RT and rel(fair, r["ref_chi2"]) < RT and rel(g["chi2"], r["ref_chi2"]) < RT
    counts, _, lo, hi = O.chi2_histogram(r["ref_chi2"])
    assert (lo, hi) == (float(r["ref_chi2_lo"]), float(r["ref_chi2_hi"]))
    np.testing.assert_array_equal(counts, r["ref_chi2_hist"])
    assert counts[0] == 0 and counts[-1] == 1 and counts.sum() == len(g["toys"])     # the maximum lands in the overflow bin
    # and the committed histogram of the restatement's own chi2 is the same histogram
    np.testing.assert_array_equal(g["chi2_hist"], r["ref_chi2_hist"])


@pytest.mark.parametrize("name", V1)
def test_chi2_test_data(name):
    """chi2TestData (src/Chi2Test.cpp:103-131): bcs0 = mean solution over the first n toys (Boost mean = sum / count)."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    G = fixture_matrix(g)
    T = len(g["toys"]) // 2
    _, B = O.chi2_test_fair(G, g["toys"][:T], g["bcs0"], g["vcs_err"])
    bcs0 = B.sum(0) / T
    chi2, _ = O.chi2_test_fair(G, g["toys"][T:], bcs0, g["vcs_err"])
    assert rel(chi2, r["ref_chi2_data"]) < 1e-11
    counts, _, lo, hi = O.chi2_histogram(r["ref_chi2_data"])
    assert (lo, hi) == (float(r["ref_chi2_data_lo"]), float(r["ref_chi2_data_hi"]))
    np.testing.assert_array_equal(counts, r["ref_chi2_data_hist"])


@pytest.mark.parametrize("name", V1)
def test_ratio_test_model(name):
    """ratioTestModel (src/RatioTest.cpp:13-62): TH1F(1024, -2, 2) of (b - b0) / b0 per point, GetMean / GetMeanError."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    _, B = O.chi2_test_fair(fixture_matrix(g), g["toys"], g["bcs0"], g["vcs_err"])
    means, errs, counts, stats = O.ratio_test(B, g["bcs0"])
    np.testing.assert_array_equal(counts.astype(np.uint16), r["ref_ratio_counts"])
    np.testing.assert_allclose(stats, r["ref_ratio_stats"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(means, r["ref_ratio_means"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(errs, r["ref_ratio_mean_errors"], rtol=1e-10, atol=1e-15)


@pytest.mark.parametrize("name", V1)
def test_iterative_solver(name):
    """IterISRInterpSolver::solve (src/IterISRInterpSolver.cpp:45-57): radiative-correction iteration, diagonal covariance."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    for tag, n_iter in (("", 10), ("3", 3)):
        b, rad, cov = O.iterative_solve(fixture_matrix(g), g["vcs"], g["vcs_err"], n_iter)
        assert rel(b, r["ref_iter%s_bcs" % tag]) < RT and rel(rad, r["ref_iter%s_radcorr" % tag]) < RT
        assert rel(np.diag(cov), r["ref_iter%s_cov_diag" % tag]) < RT


@pytest.mark.parametrize("name", V1 + ["c3_sle2_n24_eps64"])
def test_native_run_matrix_and_solution(name):
    """Nothing injected: the reference's own constructor (sorting, efficiency adaptor), evalEqMatrix and solve()."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    Gn = r["ref_G_native"]
    if "S_ref" in g:
        assert rel(Gn, g["S_ref"] @ g["G_ref"]) < 1e-14               # the energy-spread product, shim order vs BLAS
    else:
        np.testing.assert_array_equal(Gn, g["G_ref"])                 # and G_ref == G_faithful (test_ref_pin.py)
    b, cov = O.sle_solve(Gn, g["vcs"], g["vcs_err"])
    assert rel(b, r["ref_bcs_native"]) < RT and rel(cov, r["ref_cov_native"]) < RT
    if "ref_cond_native" in r:
        assert abs(O.cond_number(Gn) / float(r["ref_cond_native"]) - 1.0) < RT


def test_tikhonov_solve_and_lcurve():
    """ISRSolverTikhonov::solve + evalEqNorm2 / evalSmoothnessConstraintNorm2 / evalLCurveCurvature(+Derivative) at 12 lambda,
    both regularisers, lambdaObjective (what NLopt calls), and chi2TestModel with the Tikhonov covariance."""
    g, r = load_golden("c4_tikhonov_n32_spread"), load_golden("refrun_c4_tikhonov_n32_spread")
    G, lam = g["G_spread"], g["tik_lambda"]
    F = O.tikhonov_F(g["D"], g["w"], True)
    # |G b - v|^2_W is a difference of nearly equal vectors for small lambda: its rounding noise is eps * |v|^2_W
    x_atol = 1e-13 * float(g["vcs"] @ (g["vcs"] / g["vcs_err"] ** 2))
    for k, l in enumerate(lam):
        o = O.tikhonov_lcurve(G, g["vcs"], g["vcs_err"], F, l)
        assert rel(o["bcs"], r["ref_tik_bcs"][k]) < RT
        assert abs(o["x"] - r["ref_tik_x"][k]) < x_atol + RT * r["ref_tik_x"][k]
        assert abs(o["y"] / r["ref_tik_y"][k] - 1) < RT
        assert abs(o["curv"] / r["ref_tik_curv"][k] - 1) < 1e-11 and abs(o["dcurv"] / r["ref_tik_dcurv"][k] - 1) < 1e-11
        if k in r["ref_tik_cov_at"]:
            assert rel(o["cov"], r["ref_tik_cov"][list(r["ref_tik_cov_at"]).index(k)]) < RT
        # the committed fixture (what the GPU tests compare with) carries the same numbers
        assert rel(g["tik_bcs"][k], r["ref_tik_bcs"][k]) < RT and abs(g["tik_curv"][k] / r["ref_tik_curv"][k] - 1) < 1e-11
    Fn = O.tikhonov_F(g["D"], g["w"], False)
    for q, k in enumerate(r["ref_tik_noreg_at"]):
        o = O.tikhonov_lcurve(G, g["vcs"], g["vcs_err"], Fn, lam[k])
        assert rel(o["bcs"], r["ref_tik_noreg_bcs"][q]) < RT
        ref = r["ref_tik_noreg_xycd"][q]
        assert abs(o["x"] - ref[0]) < x_atol + RT * ref[0] and abs(o["y"] / ref[1] - 1) < RT
        assert abs(o["curv"] / ref[2] - 1) < 1e-11 and abs(o["dcurv"] / ref[3] - 1) < 1e-11
    for q, k in enumerate(r["ref_lambda_objective_at"]):
        o = O.tikhonov_lcurve(G, g["vcs"], g["vcs_err"], F, lam[k])
        assert abs(o["curv"] / r["ref_lambda_objective"][q][0] - 1) < 1e-11 and abs(o["dcurv"] / r["ref_lambda_objective"][q][1] - 1) < 1e-11
    # chi2 with the Tikhonov covariance: d^T Cov^-1 d, Cov = A+ W^-1 A+^T
    for q, k in enumerate(r["ref_tik_chi2_at"]):
        s = O.tikhonov_solve(G, g["vcs"], g["vcs_err"], F, lam[k])
        D = g["toys"] @ s["Ap"].T - g["bcs0"][None, :]
        chi2 = np.einsum("ti,ti->t", D, np.linalg.solve(s["cov"], D.T).T)
        assert rel(chi2, r["ref_tik_chi2"][q]) < 1e-10


def test_condition_number_versus_spread():
    """The loop of src/isrsolver-condnum-test.cpp:84-128 run on the reference's own classes (resetECMErrors, evalEqMatrix,
    JacobiSVD per setting) against the restated spread product and LAPACK singular values."""
    g, r = load_golden("c4_tikhonov_n32_spread"), load_golden("refrun_c4_tikhonov_n32_spread")
    I = O.Interp(g["ecm"], float(g["threshold"]), [tuple(s) for s in g["settings"]])
    G = I.eval_eq_matrix(mode="faithful")
    n = len(g["ecm"])
    for sg, ref in zip(r["ref_condscan_sigma"], r["ref_condscan_cond"]):
        c = O.cond_number(G if sg == 0 else I.energy_spread_matrix(np.full(n, sg)) @ G)
        assert abs(c / ref - 1.0) < RT
    assert np.all(np.diff(r["ref_condscan_cond"]) > 0)


def test_tsvd():
    """ISRSolverTSVD::solve: JacobiSVD, truncation to k, keep-one, full rank; minimum-norm COD solve of the rank-k matrix."""
    g, r = load_golden("c4_tikhonov_n32_spread"), load_golden("refrun_c4_tikhonov_n32_spread")
    G, k, n = r["ref_tsvd_G"], int(g["tsvd_k"]), len(g["ecm"])
    assert rel(O.tsvd_solve(G, g["vcs"], g["vcs_err"], k)[0], r["ref_tsvd_bcs"]) < 1e-11
    assert rel(O.tsvd_solve(G, g["vcs"], g["vcs_err"], k, True)[0], r["ref_tsvd_bcs_keepone"]) < 1e-11
    b, cov, _ = O.tsvd_solve(G, g["vcs"], g["vcs_err"], n)
    assert rel(b, r["ref_tsvd_bcs_full"]) < 1e-11 and rel(cov, r["ref_tsvd_cov_full"]) < 1e-11


def _bw(par, thr):
    s0, M, Gm = par
    def f(e):
        if e <= thr:
            return 0.0
        return s0 * M * M * Gm * Gm / ((e * e - M * M) ** 2 + M * M * Gm * Gm) * (1.0 - thr * thr / (e * e)) ** 1.5
    return f


def test_vcs_fit_function():
    """ISRSolverVCSFitFunction (src/ISRSolverVCSFitter.cpp:64-93, 107-177) run by the reference itself: the restated forward
    convolution + Gauss-Hermite blur + chi2 sum reproduces its objective BIT FOR BIT, with and without energy spread, and
    what saveResults writes (blurred model, radiative correction delta = model / Born - 1, Born points vcs / (1 + delta))."""
    r = load_golden("refrun_vcsfit_n24")
    thr = float(r["threshold"])
    for k, par in enumerate(r["pars"]):
        assert O.vcs_fit_chi2(r["ecm"], r["vcs"], r["vcs_err"], thr, _bw(par, thr)) == r["ref_chi2"][k]
        if k < 2:
            assert O.vcs_fit_chi2(r["ecm"], r["vcs"], r["vcs_err"], thr, _bw(par, thr), r["ecm_err"]) == r["ref_chi2_spread"][k]
    assert float(r["ref_chi2_true"]) < 1e-18                              # the data are the model at the true parameters
    f = _bw(r["pars"][0], thr)
    born = np.array([f(float(e)) for e in r["ecm"]])
    for tag, ee in (("", None), ("spread_", r["ecm_err"])):
        m = O.vcs_fit_model(r["ecm"], thr, f, ee)
        np.testing.assert_allclose(m, r["ref_save_%svcs_gauss_conv" % tag], rtol=1e-10)   # (TF1::Eval in the unblurred branch)
        delta = m / born - 1.0
        np.testing.assert_allclose(delta, r["ref_save_%srad_delta" % tag], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r["vcs"] / (delta + 1.0), r["ref_save_%sbcs" % tag], rtol=1e-10)
        np.testing.assert_allclose(r["vcs_err"] / (delta + 1.0), r["ref_save_%sbcs_err" % tag], rtol=1e-10)


@pytest.mark.parametrize("name", ["refrun_e2e_sle_n100_cspline", "refrun_e2e_sle_n200_cspline", "refrun_e2e_sle_n500_linear"])
def test_full_size_sle_runs_end_to_end(name):
    """BASELINE geometries (C2: N = 100 cubic; north star: N = 200 cubic; C5: N = 500 linear) run by the reference itself
    with nothing injected -- its own N^2 adaptive integrations, solve(), chi2TestModel, ratioTestModel.  The oracle from the
    same inputs: matrix rows bit for bit (faithful back-end), everything downstream to 1e-12, histograms bit for bit."""
    r = load_golden(name)
    I = O.Interp(r["ecm"], float(r["threshold"]), [tuple(s) for s in r["settings"]])
    G = I.eval_eq_matrix(mode="faithful")
    rows = list(r["ref_cov_rows_at"])
    np.testing.assert_array_equal(G[rows], r["ref_G_rows"])
    b, cov = O.sle_solve(G, r["vcs"], r["vcs_err"])
    assert rel(b, r["ref_bcs"]) < RT and rel(np.diag(cov), r["ref_cov_diag"]) < RT and rel(cov[rows], r["ref_cov_rows"]) < RT
    assert abs(O.cond_number(G) / float(r["ref_cond"]) - 1.0) < RT
    chi2, B = O.chi2_test_fair(G, r["toys"], r["bcs0"], r["vcs_err"])
    assert rel(chi2, r["ref_chi2"]) < RT
    counts, _, lo, hi = O.chi2_histogram(r["ref_chi2"])
    assert (lo, hi) == (float(r["ref_chi2_lo"]), float(r["ref_chi2_hi"]))
    np.testing.assert_array_equal(counts, r["ref_chi2_hist"])
    means, errs, rc, stats = O.ratio_test(B, r["bcs0"])
    np.testing.assert_array_equal(rc.astype(np.uint16), r["ref_ratio_counts"])
    np.testing.assert_allclose(means, r["ref_ratio_means"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(errs, r["ref_ratio_mean_errors"], rtol=1e-9, atol=1e-14)


def test_full_size_tikhonov_run_against_long_double():
    """The N = 200 reference run (cubic spline + energy spread, 4 lambda, 16 toys each): its per-toy chi2 against the
    80-bit evaluation of the same formula on the same operator -- how much of the reference's own double-precision route
    (two FullPivLU, an explicit covariance) is rounding noise at full size."""
    import hp_reference as H
    r = load_golden("refrun_e2e_tikhonov_n200_spread")
    n = len(r["ecm"])
    I = O.Interp(r["ecm"], float(r["threshold"]), [tuple(s) for s in r["settings"]])
    rows = list(r["ref_G_rows_at"])
    G = I.energy_spread_matrix(r["ecm_err"]) @ I.eval_eq_matrix(mode="converged")
    assert rel(G[rows], r["ref_G_rows"]) < 1e-9                       # the reference's own adaptive assembly, 3 rows
    F = O.tikhonov_F(I.deriv_projector(), I.integral_basis(), True)
    for q, l in enumerate(r["tik_lambda"]):
        hp, B = H.ld_tikhonov_chi2(G, r["vcs_err"], F, l, r["toys"], r["bcs0"])
        assert rel(hp, r["ref_tik_chi2"][q]) < 1e-7, (l, rel(hp, r["ref_tik_chi2"][q]))
        o = O.tikhonov_solve(G, r["vcs"], r["vcs_err"], F, l)
        assert rel(o["bcs"], r["ref_tik_bcs"][q]) < 1e-8


# ---------------------------------------------------------------------------------------------------------------------
# the committed files are what the compiled reference produces today (needs oracle/_ref)
# ---------------------------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libisr_ref.so not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("name", ["c2_sle_n40_cspline", "c3_sle2_n24_eps64", "c4_tikhonov_n32_spread"])
def test_refrun_fixtures_are_reference_output(name):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_refrun_golden as M
    fresh, stored = M.small_case(name), load_golden("refrun_" + name)
    assert set(fresh) == set(stored)
    for k in stored:
        np.testing.assert_array_equal(np.asarray(fresh[k]), stored[k], err_msg=k)


@needs_ref
def test_refrun_vcsfit_fixture_is_reference_output():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_refrun_golden as M
    fresh, stored = M.run_vcsfit(), load_golden("refrun_vcsfit_n24")
    assert set(fresh) == set(stored)
    for k in stored:
        np.testing.assert_array_equal(np.asarray(fresh[k]), stored[k], err_msg=k)


@needs_ref
def test_reference_constructor_sorts_points_and_applies_efficiency():
    """BaseISRSolver: both constructors sort by energy (src/BaseISRSolver.cpp:31-76, 105-128); the TEfficiency adaptor
    (:232-258) through the reference's own evalEqMatrix equals the restatement's matrix bit for bit."""
    n = 10
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    perm = np.random.default_rng(5).permutation(n)
    vcs, err, ee = 1.0 + np.arange(n), 0.1 + 0.01 * np.arange(n), 0.001 * (1 + np.arange(n))
    for ctor in ("graph", "pointers"):
        S = R.RefSolver("sle", ecm[perm], vcs[perm], err[perm], O.E_THR, ecm_err=ee[perm], ctor=ctor)
        p = S.points()
        np.testing.assert_array_equal(p["ecm"], ecm)
        np.testing.assert_array_equal(p["vcs"], vcs)
        np.testing.assert_array_equal(p["vcs_err"], err)
        np.testing.assert_array_equal(p["ecm_err"], ee)
        idx = O.sort_points(ecm[perm], vcs[perm], ee[perm], err[perm])[0]
        np.testing.assert_array_equal(idx, np.argsort(perm))
    I = O.Interp(ecm, O.E_THR)
    xe = np.array([0.0, 0.01, 0.05, 0.2, 0.5, 1.0])
    ee_edges = np.linspace(0.7, 2.2, 7)
    vals = 0.3 + 0.6 * np.random.default_rng(7).random((len(ee_edges) + 1, len(xe) + 1))
    for eff in (O.Efficiency("table_2d", values=vals, xedges=xe, eedges=ee_edges),
                O.Efficiency("table_2d", values=vals, nx=5, xlo=0.0, xhi=1.0, ne=6, elo=0.7, ehi=2.2),
                O.Efficiency("e_only", values=vals[:, 0].copy(), ne=6, elo=0.7, ehi=2.2)):
        S = R.RefSolver("sle", ecm, vcs, err, O.E_THR, eff=eff)
        np.testing.assert_array_equal(S.eval_eq_matrix(), I.eval_eq_matrix(eff=eff))


@needs_ref
def test_reference_rank_deficient_cod_is_minimum_norm():
    """completeOrthogonalDecomposition on a rank-deficient operator (what ISRSolverTSVD hands it): minimum-norm solution."""
    rng = np.random.default_rng(11)
    n, k = 12, 7
    A = rng.standard_normal((n, k)) @ rng.standard_normal((k, n))
    v = rng.standard_normal(n)
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    S = R.RefSolver("sle", ecm, v, np.ones(n), O.E_THR)
    S.inject_matrix(A)
    # solve() also inverts the singular G^T W G (inf / nan there is the reference's behaviour); only bcs is compared
    S._L.refs_solve(S._h)
    np.testing.assert_allclose(S.bcs(), np.linalg.pinv(A) @ v, rtol=1e-9, atol=1e-11)
