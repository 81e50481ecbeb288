"""The CUDA path against OUTPUTS OF THE REFERENCE ITSELF, through the C ABI (the ctypes mirror calls nothing else).

tests/golden/refrun_*.npz hold what the reference's own compiled classes and drivers returned (ISRSolverSLE / SLE2 /
Tikhonov / TSVD ::solve, evalConditionNumber, save, the L-curve functionals, chi2TestModel, chi2TestData, ratioTestModel;
tests/golden/make_refrun_golden.py) on the inputs of tests/golden/c*.npz, plus two full-size runs at N = 200 with nothing
injected.  /root/reference is not needed here.  North-star tolerances: 1e-9 matrix entries, 1e-8 for solutions,
covariances, chi2 and the Tikhonov functionals; histogram cells are compared bit for bit wherever the filled value is not
within rounding of a cell edge (the reference's chi2 and the GPU's agree to ~1e-13, so at most a couple of the
range-dependent chi2 cells may move; the fixed-range ratio cells are identical)."""
import numpy as np
import pytest

from conftest import assert_matrix_parity, load_golden
from test_gpu_assembly import make_solver

pytestmark = pytest.mark.gpu
RT = 1e-8
V1 = ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c4_tikhonov_n32_spread", "c5_mixed3_n36"]


def close(a, b, rt=RT):
    """|a - b| <= rt * max|b|: vectors and matrices are gated relative to their scale, like the reference's own noise."""
    a, b = np.asarray(a), np.asarray(b)
    d = float(np.abs(a - b).max())
    assert d <= rt * float(np.abs(b).max()), (d, float(np.abs(b).max()))


def fixture_matrix(g):
    return g["G_spread"] if "G_spread" in g else g["G_converged"]


@pytest.mark.parametrize("name", V1)
def test_sle_solution_covariance_condition_number(name):
    g, r = load_golden(name), load_golden("refrun_" + name)
    s = make_solver(g, spread="ecm_err" in g)
    s.set_matrix(fixture_matrix(g))
    s.solve()
    close(s.bcs(), r["ref_bcs"])
    close(s.bcs_cov_matrix(), r["ref_cov"])
    close(s.bcs_err(), r["ref_bcs_err"])
    close(s.bcs_inv_cov_matrix(), r["ref_inv_cov"])
    assert s.condnum_eval() == pytest.approx(float(r["ref_cond"]), rel=RT)
    # interpEval of the reference at 47 energies: 0 at and below threshold, every range type, the clamp beyond E_max
    got = s.interp_eval(r["ref_interp_E"])
    np.testing.assert_allclose(got, r["ref_interp_val"], rtol=RT, atol=RT * np.abs(r["ref_bcs"]).max())
    assert got[0] == 0.0 and got[1] == 0.0 and got[-1] == s.bcs()[-1]


@pytest.mark.parametrize("name", V1)
def test_chi2_and_ratio_tests(name):
    """isr_chi2_test on the toys the reference's chi2TestModel / ratioTestModel were fed."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    n, T = len(g["ecm"]), len(g["toys"])
    s = make_solver(g, spread="ecm_err" in g)
    s.set_matrix(fixture_matrix(g))
    o = s.chi2_test_call(g["bcs0"], V=g["toys"], want_sum=True, want_hist=True)
    np.testing.assert_allclose(o["chi2"], r["ref_chi2"], rtol=RT)
    assert o["lo"] == pytest.approx(float(r["ref_chi2_lo"]), rel=RT) and o["hi"] == pytest.approx(float(r["ref_chi2_hi"]), rel=RT)
    assert o["n"] == T and o["hist"].sum() == T
    assert np.abs(o["hist"] - r["ref_chi2_hist"]).sum() <= 4           # a value within rounding of a cell edge may move
    np.testing.assert_array_equal(o["ratio_hist"], r["ref_ratio_counts"].astype(np.uint32))
    np.testing.assert_allclose(o["ratio_stats"], r["ref_ratio_stats"], rtol=RT, atol=1e-9)
    cnt, sx, sx2 = o["ratio_stats"].T
    mean = sx / cnt
    np.testing.assert_allclose(mean, r["ref_ratio_means"], rtol=RT, atol=1e-10)
    np.testing.assert_allclose(np.sqrt(np.abs(sx2 / cnt - mean * mean)) / np.sqrt(cnt), r["ref_ratio_mean_errors"], rtol=1e-7, atol=1e-10)
    # chi2TestData: the mean solution of the first half is the model of the second half
    h = T // 2
    first = s.chi2_test_call(g["bcs0"], V=g["toys"][:h], want_sum=True)
    second = s.chi2_test_call(first["sum_bcs"] / h, V=g["toys"][h:])
    np.testing.assert_allclose(second["chi2"], r["ref_chi2_data"], rtol=RT)
    assert second["lo"] == pytest.approx(float(r["ref_chi2_data_lo"]), rel=RT)
    assert np.abs(second["hist"] - r["ref_chi2_data_hist"]).sum() <= 4


@pytest.mark.parametrize("name", V1)
def test_iterative_solver(name):
    """isr_iterative_solve against IterISRInterpSolver::solve of the reference itself (10 and 3 iterations)."""
    import isrsolver_b200 as isr
    g, r = load_golden(name), load_golden("refrun_" + name)
    for tag, n_iter in (("", 10), ("3", 3)):
        s = make_solver(g, cls=isr.IterISRInterpSolver, spread="ecm_err" in g)
        s.setNumOfIters(n_iter)
        s.set_matrix(fixture_matrix(g))
        s.solve()
        close(s.bcs(), r["ref_iter%s_bcs" % tag])
        close(s.radcorr(), r["ref_iter%s_radcorr" % tag])
        close(np.diag(s.bcs_cov_matrix()), r["ref_iter%s_cov_diag" % tag])
        assert np.count_nonzero(s.bcs_cov_matrix() - np.diag(np.diag(s.bcs_cov_matrix()))) == 0


@pytest.mark.parametrize("name", V1)
def test_native_assembly_and_solve(name):
    """Nothing injected on either side: the reference's evalEqMatrix + solve() against isr_assemble + isr_factor + isr_solve."""
    g, r = load_golden(name), load_golden("refrun_" + name)
    s = make_solver(g, spread="ecm_err" in g)
    s.eval_equation_matrix()
    err = (32 if "ecm_err" in g else 1) * 3 * g["G_faithful_err"]
    assert_matrix_parity(s.intop_matrix(), r["ref_G_native"], err.max() if "ecm_err" in g else err, row_rel=1e-9, what=name + " vs the reference run")
    s.solve()
    close(s.bcs(), r["ref_bcs_native"])
    close(s.bcs_cov_matrix(), r["ref_cov_native"])


def test_sle2_native_assembly_and_solve():
    """ISRSolverSLE2 (one efficiency histogram per energy point) run by the reference itself."""
    import isrsolver_b200 as isr
    g, r = load_golden("c3_sle2_n24_eps64"), load_golden("refrun_c3_sle2_n24_eps64")
    s = isr.ISRSolverSLE2(len(g["ecm"]), g["ecm"], g["vcs"], None, g["vcs_err"], float(g["threshold"]), eff_hists=g["eff_values"], xlo=0.0, xhi=1.0)
    s.set_interp_settings([tuple(x) for x in g["settings"]])
    s.eval_equation_matrix()
    assert_matrix_parity(s.intop_matrix(), r["ref_G_native"], 3 * g["G_faithful_err"], row_rel=1e-9, what="SLE2 vs the reference run")
    s.solve()
    # the binned efficiency makes the integrands piecewise: the reference's adaptive quadrature is converged to ~1e-9 of
    # the row scale there (G_faithful_relerr), which the solution inherits; on the reference's own matrix the gate is 1e-8
    close(s.bcs(), r["ref_bcs_native"], 1e-7)
    s.set_matrix(r["ref_G_native"])
    s.solve()
    close(s.bcs(), r["ref_bcs_native"])
    close(s.bcs_cov_matrix(), r["ref_cov_native"])
    assert s.condnum_eval() == pytest.approx(float(r["ref_cond_native"]), rel=RT)


def _x_atol(g):
    # evalEqNorm2 = |G b - v|^2_W is formed in double precision by the reference: rounding noise eps * |v|^2_W, which at
    # lambda = 1e-9 is 1e-5 of the value itself (tests/test_gpu_tikhonov.py gates the GPU against 50-digit arithmetic there)
    return 1e-13 * float(g["vcs"] @ (g["vcs"] / g["vcs_err"] ** 2))


def test_tikhonov_scan_solve_and_toys():
    import isrsolver_b200 as isr
    g, r = load_golden("c4_tikhonov_n32_spread"), load_golden("refrun_c4_tikhonov_n32_spread")
    lam = g["tik_lambda"]
    s = make_solver(g, cls=isr.ISRSolverTikhonov, spread=True)
    s.set_matrix(g["G_spread"])
    x, y, c, dc, B = s._scan(lam, want_bcs=True)
    close(B, r["ref_tik_bcs"])
    assert np.all(np.abs(x - r["ref_tik_x"]) <= _x_atol(g) + RT * r["ref_tik_x"])
    np.testing.assert_allclose(y, r["ref_tik_y"], rtol=RT)
    np.testing.assert_allclose(c, r["ref_tik_curv"], rtol=RT)
    np.testing.assert_allclose(dc, r["ref_tik_dcurv"], rtol=RT)
    for q, k in enumerate(r["ref_tik_cov_at"]):
        s.reg_param = float(lam[k])
        s.solve()
        close(s.bcs(), r["ref_tik_bcs"][k])
        close(s.bcs_cov_matrix(), r["ref_tik_cov"][q])
    for q, k in enumerate(r["ref_lambda_objective_at"]):           # what NLopt evaluates per trial lambda
        s.reg_param = float(lam[k])
        assert s.evalLCurveCurvature() == pytest.approx(r["ref_lambda_objective"][q][0], rel=RT)
        assert s.evalLCurveCurvatureDerivative() == pytest.approx(r["ref_lambda_objective"][q][1], rel=RT)
    chi2 = s.toy_scan(lam[list(r["ref_tik_chi2_at"])], g["toys"], g["bcs0"])      # chi2TestModel with a Tikhonov solver
    np.testing.assert_allclose(chi2, r["ref_tik_chi2"], rtol=RT)
    s.disableDerivNorm2Regularizator()
    at = list(r["ref_tik_noreg_at"])
    x, y, c, dc, B = s._scan(lam[at], want_bcs=True)
    close(B, r["ref_tik_noreg_bcs"])
    ref = r["ref_tik_noreg_xycd"]
    assert np.all(np.abs(x - ref[:, 0]) <= _x_atol(g) + RT * ref[:, 0])
    np.testing.assert_allclose(y, ref[:, 1], rtol=RT)
    np.testing.assert_allclose(c, ref[:, 2], rtol=RT)
    np.testing.assert_allclose(dc, ref[:, 3], rtol=RT)


def test_condition_number_versus_spread_scan():
    """isr_cond_spread_scan (one assembly, batched spread products and Jacobi singular values) against the reference's own
    loop, which re-assembles the matrix and runs a JacobiSVD per setting (src/isrsolver-condnum-test.cpp:84-128)."""
    g, r = load_golden("c4_tikhonov_n32_spread"), load_golden("refrun_c4_tikhonov_n32_spread")
    s = make_solver(g, spread=False)
    cond = s.condnum_spread_scan(r["ref_condscan_sigma"])
    np.testing.assert_allclose(cond, r["ref_condscan_cond"], rtol=RT)


def test_tsvd():
    import isrsolver_b200 as isr
    g, r = load_golden("c4_tikhonov_n32_spread"), load_golden("refrun_c4_tikhonov_n32_spread")
    n, k = len(g["ecm"]), int(g["tsvd_k"])
    s = make_solver(g, cls=isr.ISRSolverTSVD, spread=True)
    s.set_matrix(r["ref_tsvd_G"])
    s.setUpperTSVDIndex(k)
    s.solve()
    close(s.bcs(), r["ref_tsvd_bcs"])
    s.enable_keepone()
    s.solve()
    close(s.bcs(), r["ref_tsvd_bcs_keepone"])
    s.disable_keepone()
    s.setUpperTSVDIndex(n)
    s.solve()
    close(s.bcs(), r["ref_tsvd_bcs_full"])
    close(s.bcs_cov_matrix(), r["ref_tsvd_cov_full"])


def test_vcs_fit_objective_through_the_convolution_service():
    """The chi2 objective of the reference's visible-cross-section fit (ISRSolverVCSFitFunction::operator(): N forward
    convolutions of a parametrised Born function, optionally blurred) and what its saveResults writes, against the GPU
    convolution service -- one plan per data set, one convolve() per parameter set."""
    import isrsolver_b200 as isr
    r = load_golden("refrun_vcsfit_n24")
    thr, ecm = float(r["threshold"]), r["ecm"]

    def bw(par):
        s0, M, Gm = par
        def f(e):
            e = np.asarray(e, dtype=np.float64)
            tf = np.where(e > thr, 1.0 - thr * thr / (e * e), 0.0)
            return np.where(e > thr, s0 * M * M * Gm * Gm / ((e * e - M * M) ** 2 + M * M * Gm * Gm) * tf ** 1.5, 0.0)
        return f
    plan = isr.ConvolutionPlan(ecm, threshold=thr)
    for k, par in enumerate(r["pars"]):
        vis, _ = plan.convolve(bw(par))
        chi2 = float((((vis - r["vcs"]) / r["vcs_err"]) ** 2).sum())
        assert chi2 == pytest.approx(float(r["ref_chi2"][k]), rel=RT)
        blurred = isr.convolutionKuraevFadinBlurred(ecm, bw(par), thr, r["ecm_err"])
        chi2s = float((((blurred - r["vcs"]) / r["vcs_err"]) ** 2).sum())
        assert chi2s == pytest.approx(float(r["ref_chi2_spread"][k]), rel=RT)
        if k == 0:
            born = bw(par)(ecm)
            np.testing.assert_allclose(vis, r["ref_save_vcs_gauss_conv"], rtol=1e-10)
            np.testing.assert_allclose(blurred, r["ref_save_spread_vcs_gauss_conv"], rtol=1e-10)
            np.testing.assert_allclose(vis / born - 1.0, r["ref_save_rad_delta"], rtol=RT, atol=1e-11)
            np.testing.assert_allclose(r["vcs"] / (blurred / born), r["ref_save_spread_bcs"], rtol=1e-10)
    # at the true parameters the data ARE the model: chi2 collapses to rounding
    vis, _ = plan.convolve(bw((4.0, 1.50, 0.25)))
    assert float((((vis - r["vcs"]) / r["vcs_err"]) ** 2).sum()) < 1e-12


@pytest.mark.parametrize("name", ["refrun_e2e_sle_n200_cspline", "refrun_e2e_sle_n100_cspline", "refrun_e2e_sle_n500_linear"])
def test_full_size_sle_runs(name):
    """BASELINE geometries end to end, nothing injected on either side: the reference assembled the matrix with its N^2
    adaptive integrations, solved, and ran chi2TestModel + ratioTestModel (a full O(N^3) solve per toy); the GPU does the
    same from the same inputs.  N = 200 cubic spline (the north-star geometry), N = 100 cubic spline (C2), N = 500 linear
    (C5: a lower-triangular operator -- the blocked triangular inverse and the triangular toy schedule)."""
    import isrsolver_b200 as isr
    r = load_golden(name)
    n, T = len(r["ecm"]), len(r["toys"])
    s = isr.ISRSolverSLE(n, r["ecm"], r["vcs"], None, r["vcs_err"], float(r["threshold"]))
    s.set_interp_settings([tuple(x) for x in r["settings"]])
    s.eval_equation_matrix()
    rows = list(r["ref_cov_rows_at"])
    G = s.intop_matrix()
    scale = np.abs(r["ref_G_rows"]).max(axis=1, keepdims=True)
    assert np.all(np.abs(G[rows] - r["ref_G_rows"]) <= 1e-9 * scale + 2e-12)
    if "linear" in name:
        assert np.all(np.triu(G, 1) == 0) and np.all(np.triu(r["ref_G_rows"][0:1], 1) == 0)
    s.solve()
    close(s.bcs(), r["ref_bcs"])
    cov = s.bcs_cov_matrix()
    close(np.diag(cov), r["ref_cov_diag"])
    close(cov[rows], r["ref_cov_rows"])
    assert s.condnum_eval() == pytest.approx(float(r["ref_cond"]), rel=RT)
    o = s.chi2_test_call(r["bcs0"], V=r["toys"], want_hist=True)
    np.testing.assert_allclose(o["chi2"], r["ref_chi2"], rtol=RT)
    assert o["lo"] == pytest.approx(float(r["ref_chi2_lo"]), rel=RT) and o["hi"] == pytest.approx(float(r["ref_chi2_hi"]), rel=RT)
    assert np.abs(o["hist"] - r["ref_chi2_hist"]).sum() <= 4
    assert np.abs(o["ratio_hist"].astype(np.int64) - r["ref_ratio_counts"].astype(np.int64)).sum() <= 4
    cnt, sx, sx2 = o["ratio_stats"].T
    # a ratio (b - b0) / b0 inherits the 1e-9 relative difference of the two assemblies as an ABSOLUTE 1e-9 (b / b0 ~ 1)
    np.testing.assert_allclose(sx / cnt, r["ref_ratio_means"], rtol=RT, atol=RT)


def test_full_size_sle2_run_n200_eps512(oracle):
    """BASELINE configs C3 end to end: ISRSolverSLE2 at N = 200 with one 512-bin efficiency histogram per energy point, run
    by the reference itself (its TH1D adaptor, its adaptive quadrature across the bin edges).  Here the REFERENCE is the
    less accurate side -- its retry loop ends at epsrel up to 1e-3 on these piecewise integrands (the fixture records the
    error estimates it returned; the converged oracle differs from it by up to 9e-6) -- so the matrix gate is the
    reference's own error estimate, and the solution is gated at 1e-8 against the converged oracle, at the matrix
    difference against the reference."""
    import isrsolver_b200 as isr
    r = load_golden("refrun_e2e_sle2_n200_eps512")
    n, nb = len(r["ecm"]), int(r["eff_nbins"])
    effo = oracle.row_table_efficiency(r["ecm"], nb)
    s = isr.ISRSolverSLE2(n, r["ecm"], r["vcs"], None, r["vcs_err"], float(r["threshold"]), eff_hists=effo.values, xlo=0.0, xhi=1.0)
    s.eval_equation_matrix()
    G = s.intop_matrix()
    rows = list(r["ref_G_rows_at"])
    scale = np.abs(r["ref_G_rows"]).max(axis=1, keepdims=True)
    assert np.all(np.abs(G[rows] - r["ref_G_rows"]) <= 3 * r["ref_G_err_rows"] + 1e-9 * scale + 2e-12)
    assert np.all(np.triu(G, 1) == 0)
    Gc, Ec = oracle.Interp(r["ecm"], float(r["threshold"])).eval_rows(rows, eff=effo, mode="converged")
    assert np.all(np.abs(G[rows] - Gc) <= 1e-9 * np.abs(Gc).max(axis=1, keepdims=True) + Ec + 2e-12)
    s.solve()
    # against the reference: its matrix is up to 9e-6 of the row scale away from the converged integrals (measured when the
    # fixture was made: max |G_faithful - G_converged| = 8.8e-6), and the solution inherits that
    close(s.bcs(), r["ref_bcs"], 2e-5)
    close(np.diag(s.bcs_cov_matrix()), r["ref_cov_diag"], 2e-5)
    assert s.condnum_eval() == pytest.approx(float(r["ref_cond"]), rel=2e-5)
    # against the converged oracle (the exact value of the integrals the reference approximates): the 1e-8 gate
    Gfull = oracle.Interp(r["ecm"], float(r["threshold"])).eval_eq_matrix(eff=effo, mode="converged")
    b, cov = oracle.sle_solve(Gfull, r["vcs"], r["vcs_err"])
    close(s.bcs(), b)
    close(np.diag(s.bcs_cov_matrix()), np.diag(cov))


def test_full_size_tikhonov_run_n200():
    """BASELINE configs[3] geometry end to end: N = 200, cubic spline, 2 MeV energy spread; ISRSolverTikhonov at 4 lambda
    with the reference's chi2TestModel on 16 toys each (its own assembly, spread product, two FullPivLU per solve)."""
    import isrsolver_b200 as isr
    r = load_golden("refrun_e2e_tikhonov_n200_spread")
    n = len(r["ecm"])
    lam = r["tik_lambda"]
    s = isr.ISRSolverTikhonov(n, r["ecm"], r["vcs"], r["ecm_err"], r["vcs_err"], float(r["threshold"]), enableEnergySpread=True)
    s.set_interp_settings([tuple(x) for x in r["settings"]])
    s.eval_equation_matrix()
    rows = list(r["ref_G_rows_at"])
    scale = np.abs(r["ref_G_rows"]).max(axis=1, keepdims=True)
    assert np.all(np.abs(s.intop_matrix()[rows] - r["ref_G_rows"]) <= 1e-9 * scale + 2e-12)
    x, y, c, dc, B = s._scan(lam, want_bcs=True)
    close(B, r["ref_tik_bcs"])
    ref = r["ref_tik_xycd"]
    # x: the reference's residual norm carries eps |v|^2_W of cancellation noise plus the 1e-9-level difference of the two
    # assemblies seen through a small residual (|G b - v| << |v|): both scale with |v|^2_W, not with x
    vw = float(r["vcs"] @ (r["vcs"] / r["vcs_err"] ** 2))
    assert np.all(np.abs(x - ref[:, 0]) <= 1e-9 * vw + RT * ref[:, 0])
    np.testing.assert_allclose(y, ref[:, 1], rtol=RT)
    np.testing.assert_allclose(c, ref[:, 2], rtol=RT)
    np.testing.assert_allclose(dc, ref[:, 3], rtol=1e-7)              # second derivative through two solves with L on either side
    chi2 = s.toy_scan(lam, r["toys"], r["bcs0"])
    np.testing.assert_allclose(chi2, r["ref_tik_chi2"], rtol=RT)
