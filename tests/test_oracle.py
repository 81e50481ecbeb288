"""The oracle against independent anchors (no GPU).  The reference has no tests or golden vectors
(SURVEY.md 4, F5), so the restatement is pinned here against SciPy's QUADPACK, mpmath, analytic
identities and the survey's probed known-answer values."""
import numpy as np
import pytest
from scipy import integrate
from scipy.interpolate import CubicSpline

from conftest import load_golden


def test_kernel_known_answers(oracle):
    O = oracle
    assert O.beta(2.25) == pytest.approx(0.0695415742432671, rel=1e-14)
    assert O.kernel_kf(1e-6, 2.25) == pytest.approx(28082.309274929197, rel=1e-13)
    assert O.kernel_kf(0.01, 2.25) == pytest.approx(5.283686571158812, rel=1e-13)
    assert O.kernel_kf(0.5, 2.25) == pytest.approx(0.08908917563287333, rel=1e-13)


def test_kernel_against_mpmath(oracle):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    AL, ME = mp.mpf(oracle.ALPHA_QED), mp.mpf(oracle.ELECTRON_M)

    def F(x, s):
        x, s = mp.mpf(x), mp.mpf(s)
        L = mp.log(s / ME / ME); b = 2 * AL / mp.pi * (L - 1)
        lnx, mx = mp.log(x), 1 - x; lnm = mp.log(mx)
        r = b * x ** (b - 1) * (1 + AL / mp.pi * (mp.pi ** 2 / 3 - mp.mpf(1) / 2) + mp.mpf(3) / 4 * b
                                - b * b / 24 * (L / 3 + 2 * mp.pi ** 2 - mp.mpf(37) / 4))
        r += -b * (1 - x / 2) + b * b / 8 * (-4 * (2 - x) * lnx - (1 + 3 * mx * mx) / x * lnm - 6 + x)
        x0 = 4 * ME / mp.sqrt(s)
        if x > x0:
            ss = 2 * lnx + L - mp.mpf(5) / 3
            r += (AL / mp.pi) ** 2 * ((x - x0) ** b / 6 / x * ss ** 2 * (2 - 2 * x + x * x + b / 3 * ss)
                                      + L * L / 2 * (mp.mpf(2) / 3 * (1 - mx ** 3) / mx + (2 - x) * lnm + x / 2))
        return r
    for s in (1.18 ** 2, 2.25, 4.0):
        for x in (1e-8, 1e-4, 0.00137, 0.002, 0.05, 0.4, 0.8):
            assert oracle.kernel_kf(x, s) == pytest.approx(float(F(x, s)), rel=2e-12)


def test_gk21_table_matches_scipy():
    """The generated 21-point Kronrod rule against the copy inside SciPy."""
    from scipy.integrate import _quad_vec
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "oracle", "gk_tables.h")).read()
    def arr(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", hdr, re.S).group(1)
        return np.array([float(v) for v in body.split(",")])
    x, w = arr("xgk21"), arr("wgk21")
    # scipy stores all 21 nodes (descending), the 10-point Gauss weights "w" and the 21-point weights "v"
    src = open(_quad_vec.__file__).read()
    body = src[src.index("def _quadrature_gk21"):]
    body = body[:body.index("return _quadrature_gk")]
    def tup(tag):
        m = re.search(tag + r" = \((.*?)\)", body, re.S)
        return np.array([float(t) for t in m.group(1).replace("\n", " ").split(",") if t.strip()])
    sx, sg, sv = tup("x"), tup("w"), tup("v")
    np.testing.assert_allclose(x, sx[:11], rtol=0, atol=2e-16)
    np.testing.assert_allclose(w, sv[:11], rtol=0, atol=2e-16)
    # embedded 10-point Gauss weights sit on the odd positions of the Kronrod node list
    g = arr("wgf21")
    np.testing.assert_allclose(g[1::2], sg[:5], rtol=0, atol=2e-16)
    assert np.all(g[0::2] == 0)


@pytest.mark.parametrize("alpha", [-0.9, -0.5, 0.0, 2.6])
def test_qags_against_scipy_quadpack(oracle, alpha):
    """QUADPACK book family x^alpha log(1/x) on [0,1]: same evaluation count as SciPy's DQAGSE and the
    exact answer 1/(alpha+1)^2 within the requested tolerance."""
    r, e, neval, st = oracle.probe_quad("qags", 2, alpha, 0.0, 0.0, 1.0, 0.0, 1e-10)
    s = integrate.quad(lambda x: x ** alpha * np.log(1 / x), 0, 1, epsabs=0, epsrel=1e-10, limit=100000, full_output=1)
    exact = 1.0 / (alpha + 1) ** 2
    assert st == 0
    assert neval == s[2]["neval"]
    assert r == pytest.approx(exact, rel=1e-9)
    assert r == pytest.approx(s[0], rel=1e-9)


def test_kernel_integrals_known_answers(oracle):
    """integrals of F at E = 1.5 GeV (mpmath 50-digit values)."""
    E = 1.5; x0 = 4 * oracle.ELECTRON_M / E
    r, e, n, st = oracle.probe_quad("qags", 0, E * E, 0.0, 0.0, x0)
    assert st == 0 and r == pytest.approx(0.667004136029696732, rel=2e-13)
    r2, e2, n2, st2 = oracle.probe_quad("qag61", 0, E * E, 0.0, x0, 1 - (0.827 / 1.5) ** 2)
    assert st2 == 0 and r2 == pytest.approx(0.32585579807259912, rel=2e-13)
    # analytic identity: the leading term C beta x^(beta-1) integrates to C x0^beta
    s = integrate.quad(lambda x: oracle.kernel_kf(x, E * E), 0, x0, epsabs=1e-12, epsrel=1e-12, limit=100000)
    assert r == pytest.approx(s[0], rel=1e-12)


def test_qag61_smooth(oracle):
    r, e, n, st = oracle.probe_quad("qag61", 3, 7.0, 0.3, 0.0, 5.0)
    exact = integrate.quad(lambda x: np.cos(7 * x) * np.exp(-0.3 * x), 0, 5, epsabs=1e-14, epsrel=1e-14)[0]
    assert st == 0 and r == pytest.approx(exact, abs=1e-13)


def test_cspline_against_scipy(oracle):
    """GSL natural cubic spline restatement == scipy CubicSpline(bc_type='natural')."""
    rng = np.random.default_rng(1)
    ecm = np.sort(1.0 + rng.random(17))
    I = oracle.Interp(ecm, 0.9, [(True, 0, 16)])
    for j in (0, 5, 16):
        y = np.zeros(18); y[j + 1] = 1
        cs = CubicSpline(I.ext, y, bc_type="natural")
        for e in np.linspace(0.9001, ecm[-1], 57):
            assert I.basis_eval(j, e) == pytest.approx(float(cs(e)), abs=1e-13)
            assert I.basis_deriv_eval(j, e) == pytest.approx(float(cs(e, 1)), abs=1e-11)


def test_linear_basis_and_ranges(oracle):
    ecm = np.linspace(1.0, 2.0, 6)
    I = oracle.Interp(ecm, 0.8)
    assert I.ranges() == [(0, 0, 6, 0.8, 2.0)]
    assert I.basis_eval(0, 0.8) == 0 and I.basis_eval(0, 0.7) == 0
    assert I.basis_eval(0, 0.9) == pytest.approx(0.5)
    assert I.basis_eval(2, ecm[2]) == pytest.approx(1.0)
    assert I.basis_eval(5, 2.5) == pytest.approx(1.0)          # clamped beyond E_max
    assert I.basis_eval(3, 2.5) == 0
    # partition of unity on (E_0, E_max]
    for e in np.linspace(1.0001, 2.0, 13):
        assert sum(I.basis_eval(j, e) for j in range(6)) == pytest.approx(1.0, abs=1e-14)
    # mixed ranges: begin index / number of segments (src/BaseRangeInterpolator.cpp:80-100)
    I2 = oracle.Interp(ecm, 0.8, [(True, 2, 5), (False, 0, 1)])
    assert I2.ranges() == [(0, 0, 2, 0.8, ecm[1]), (1, 1, 5, ecm[1], 2.0)]
    for bad in ([(False, 1, 5)], [(False, 0, 4)], [(False, 0, 2), (True, 4, 5)], [(False, 0, 3), (True, 3, 5)],
                [(False, 3, 2), (True, 0, 5)]):
        with pytest.raises(oracle.InterpRangeException):
            oracle.Interp(ecm, 0.8, bad)


def test_find_bin_root_semantics(oracle):
    fb = oracle.find_bin
    assert fb(10, 0.0, 1.0, None, -1e-9) == 0
    assert fb(10, 0.0, 1.0, None, 0.0) == 1
    assert fb(10, 0.0, 1.0, None, 0.0999999) == 1
    assert fb(10, 0.0, 1.0, None, 0.1) == 1 + int(10 * 0.1 / 1.0)
    assert fb(10, 0.0, 1.0, None, 1.0) == 11
    edges = [0.0, 0.1, 0.5, 2.0]
    assert [fb(3, 0, 2, edges, v) for v in (-1, 0.0, 0.05, 0.1, 0.49, 0.5, 1.99, 2.0, 3)] == [0, 1, 1, 2, 2, 3, 3, 4, 4]


def test_matrix_anchor_n20(oracle):
    """SURVEY.md 8c(iv) four-digit anchors + structure."""
    ecm = np.linspace(1.18, 2.0, 20)
    G = oracle.Interp(ecm, 0.827).eval_eq_matrix()
    assert G[0, 0] == pytest.approx(0.9346, abs=5e-5) and G[1, 0] == pytest.approx(0.1288, abs=5e-5)
    assert G[1, 1] == pytest.approx(0.8232, abs=5e-5)
    assert np.all(np.triu(G, 1) == 0)
    assert oracle.cond_number(G) == pytest.approx(1.388, abs=1e-3)
    G2 = oracle.Interp(ecm, 0.827, [(True, 0, 19)]).eval_eq_matrix()
    assert G2[0, 0] == pytest.approx(1.149, abs=5e-4) and G2[0, 1] == pytest.approx(-0.248, abs=5e-4)
    assert oracle.cond_number(G2) == pytest.approx(1.751, abs=1e-3)


def test_matrix_entry_against_mpmath(oracle):
    """One cubic-spline entry integrated piecewise with mpmath (30 digits): the converged back-end agrees
    to 1e-12; the faithful back-end (reference call structure) within 1e-9."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 30
    n = 12; ecm = np.linspace(1.18, 2.0, n); thr = 0.827
    I = oracle.Interp(ecm, thr, [(True, 0, n - 1)])
    i, j = 9, 4
    E = I.ext[i + 1]; s = E * E
    y = np.zeros(n + 1); y[j + 1] = 1
    cs = CubicSpline(I.ext, y, bc_type="natural")
    AL, ME = mp.mpf(oracle.ALPHA_QED), mp.mpf(oracle.ELECTRON_M)
    L = mp.log(mp.mpf(s) / ME / ME); b = 2 * AL / mp.pi * (L - 1); x0 = 4 * ME / mp.mpf(E)

    def F(x):
        lnx, mx = mp.log(x), 1 - x; lnm = mp.log(mx)
        r = b * x ** (b - 1) * (1 + AL / mp.pi * (mp.pi ** 2 / 3 - mp.mpf(1) / 2) + mp.mpf(3) / 4 * b
                                - b * b / 24 * (L / 3 + 2 * mp.pi ** 2 - mp.mpf(37) / 4))
        r += -b * (1 - x / 2) + b * b / 8 * (-4 * (2 - x) * lnx - (1 + 3 * mx * mx) / x * lnm - 6 + x)
        if x > x0:
            ss = 2 * lnx + L - mp.mpf(5) / 3
            r += (AL / mp.pi) ** 2 * ((x - x0) ** b / 6 / x * ss ** 2 * (2 - 2 * x + x * x + b / 3 * ss)
                                      + L * L / 2 * (mp.mpf(2) / 3 * (1 - mx ** 3) / mx + (2 - x) * lnm + x / 2))
        return r
    knots = sorted(set([0.0, float(x0)] + [1 - (I.ext[m] / E) ** 2 for m in range(0, i + 1)]))
    tot = mp.mpf(0)
    for a, c in zip(knots[:-1], knots[1:]):
        mid = 0.5 * (a + c); k = int(np.searchsorted(I.ext, E * np.sqrt(1 - mid)) - 1)
        co = cs.c[:, k]; xk = I.ext[k]
        f = lambda x: F(x) * sum(mp.mpf(float(co[3 - q])) * (mp.mpf(E) * mp.sqrt(1 - x) - xk) ** q for q in range(4))
        if a == 0:
            g = lambda t: f(t ** (1 / b)) * (1 / b) * t ** (1 / b - 1)
            tot += mp.quad(g, [0, mp.mpf(c) ** b])
        else:
            tot += mp.quad(f, [a, c])
    truth = float(tot)
    assert I.kf_basis_integral(i, j) == pytest.approx(truth, abs=1e-9)
    Gc = I.eval_eq_matrix(mode="converged", nthreads=4)
    assert Gc[i, j] == pytest.approx(truth, abs=1e-12)


def test_faithful_vs_converged_backends(oracle):
    """Smooth integrands: both back-ends agree to ~1e-10.  Tabulated efficiency: the reference's
    loosen-until-success loop (src/Integration.cpp:28-42) ends at epsrel up to 1e-3, and its own error
    estimate covers the difference to the converged value."""
    g = load_golden("c1_sle_n50_linear")
    assert np.abs(g["G_faithful"] - g["G_converged"]).max() < 1e-10
    g3 = load_golden("c3_sle2_n24_eps64")
    assert g3["G_faithful_relerr"].max() > 1e-9
    d = np.abs(g3["G_faithful"] - g3["G_converged"])
    # QUADPACK's abserr is an estimate, not a bound: allow 3x
    assert np.all(d <= 1e-9 * np.abs(g3["G_converged"]) + 3 * (g3["G_faithful_err"] + g3["G_converged_err"]) + 1e-13)


@pytest.mark.parametrize("name", ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c3_sle2_n24_eps64", "c5_mixed3_n36"])
def test_golden_reproducible(oracle, name):
    """The committed fixtures are what the oracle produces today (guards against silent oracle drift)."""
    g = load_golden(name)
    eff = None
    if "eff_values" in g:
        eff = oracle.Efficiency("row_table", values=g["eff_values"], nx=int(g["eff_nx"]), xlo=float(g["eff_xlo"]), xhi=float(g["eff_xhi"]))
    I = oracle.Interp(g["ecm"], float(g["threshold"]), [tuple(r) for r in g["settings"]])
    G = I.eval_eq_matrix(eff=eff, mode="converged")
    np.testing.assert_allclose(G, g["G_converged"], rtol=1e-13, atol=1e-15)
    b, cov = oracle.sle_solve(G, g["vcs"], g["vcs_err"])
    np.testing.assert_allclose(b, g["bcs"], rtol=1e-11)


def test_solve_chi2_loop_equals_fair(oracle):
    """The reference's per-toy O(N^3) loop and the factor-once restatement give the same chi2 and b."""
    g = load_golden("c2_sle_n40_cspline")
    np.testing.assert_allclose(g["chi2_loop"], g["chi2"], rtol=1e-10)
    # identity A.7 of the survey: chi2_k = |W^(1/2) (v_k - G b0)|^2
    G = g["G_converged"]
    r = (g["toys"] - (G @ g["bcs0"])[None, :]) / g["vcs_err"][None, :]
    np.testing.assert_allclose((r * r).sum(1), g["chi2"], rtol=1e-10)


def test_tikhonov_eigenform_identities(oracle):
    """Appendix A.8: the diagonal-rescaling formulas the CUDA scan uses, against the literal matrix
    implementation of src/ISRSolverTikhonov.cpp."""
    import scipy.linalg as sla
    g = load_golden("c4_tikhonov_n32_spread")
    G, F, v, err = g["G_spread"], g["tik_F"], g["vcs"], g["vcs_err"]
    W = np.diag(err ** -2.0)
    A = G.T @ W @ G
    mu, U = sla.eigh(A, F)
    c = U.T @ G.T @ W @ v
    for q, lam in enumerate(g["tik_lambda"]):
        d = mu + lam
        np.testing.assert_allclose(U @ (c / d), g["tik_bcs"][q], rtol=1e-6)
        assert np.sum(c * c / d ** 2) == pytest.approx(g["tik_y"][q], rel=1e-6)
        assert np.sum(c * c * lam * lam / (mu * d ** 2)) == pytest.approx(g["tik_x"][q], rel=1e-5, abs=1e-9)
        dksi = -2 * np.sum(c * c / d ** 3)
        assert -abs(1 / dksi / (1 + lam * lam) ** 1.5) == pytest.approx(g["tik_curv"][q], rel=1e-5)


def test_histogram_semantics(oracle):
    v = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.0])
    counts, idx, lo, hi = oracle.chi2_histogram(v)
    assert (lo, hi) == (0.0, 2.0)
    assert counts[129] == 2 and counts[0] == 0 and counts.sum() == 6      # the maximum lands in overflow
    assert idx[1] == 1 + int(128 * 0.5 / 2.0)


def test_sort_points(oracle):
    e = np.array([1.5, 1.2, 1.9]); idx, es, vs, ee, ve = oracle.sort_points(e, [1, 2, 3], None, [4, 5, 6])
    assert list(idx) == [1, 0, 2] and list(vs) == [2, 1, 3] and ee is None and list(ve) == [5, 4, 6]


def test_linear_algebra_layer_against_exact_arithmetic(oracle):
    """The linear-algebra layer of the oracle (NumPy/LAPACK in place of Eigen -- the part of the reference that
    cannot be compiled here) against the SAME formulas evaluated in 60-digit mpmath arithmetic: solve and
    covariance (src/ISRSolverSLE.cpp:91-94), per-toy chi2 (src/Chi2Test.cpp:51-55), the Tikhonov matrices,
    solution, covariance and L-curve functionals (src/ISRSolverTikhonov.cpp:61-72, 106-128, 139-155) and the TSVD
    solution (src/ISRSolverTSVD.cpp:63-73)."""
    import mpmath as mp
    O = oracle
    mp.mp.dps = 60
    n = 10
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    I = O.Interp(ecm, O.E_THR, [(False, 0, 0), (True, 1, n - 1)])
    G = I.eval_eq_matrix(mode="converged")
    b0 = O.bw_born(ecm); vcs = G @ b0; err = 0.05 * vcs
    V = O.draw_toys(vcs, err, 4)
    mG, mW = mp.matrix(G.tolist()), mp.diag([mp.mpf(float(e)) ** -2 for e in err])
    f = lambda M: np.array(M.tolist(), dtype=float).reshape(M.rows, M.cols)
    # --- SLE
    mb = mp.lu_solve(mG, mp.matrix(vcs.tolist()))
    mCovInv = mG.T * mW * mG
    mCov = mp.inverse(mCovInv)
    b, cov = O.sle_solve(G, vcs, err)
    np.testing.assert_allclose(b, f(mb).ravel(), rtol=1e-11)
    np.testing.assert_allclose(cov, f(mCov), rtol=1e-9, atol=1e-12 * np.abs(cov).max())
    c2_loop, _ = O.chi2_test_loop(G, V, b0, err)
    c2_fair, _ = O.chi2_test_fair(G, V, b0, err)
    for k in range(len(V)):
        d = mp.lu_solve(mG, mp.matrix(V[k].tolist())) - mp.matrix(b0.tolist())
        exact = float((d.T * mCovInv * d)[0])
        assert c2_loop[k] == pytest.approx(exact, rel=1e-9) and c2_fair[k] == pytest.approx(exact, rel=1e-11)
    # --- Tikhonov
    D, w = I.deriv_projector(), I.integral_basis()
    F = O.tikhonov_F(D, w, True)
    mF = mp.matrix(D.tolist()).T * mp.diag([mp.mpf(float(x)) for x in w]) * mp.matrix(D.tolist())
    np.testing.assert_allclose(F, f(mF), rtol=1e-12, atol=1e-13 * np.abs(F).max())
    lam = mp.mpf("1e-3")
    mT = mCovInv + lam * mF
    mL = mp.inverse(mF) * mCovInv + lam * mp.eye(n)
    mAp = mp.inverse(mT) * mG.T * mW      # A+ = T^-1 G^T W
    mbt = mAp * mp.matrix(vcs.tolist())
    mCt = mAp * mp.inverse(mW) * mAp.T
    dv = mG * mbt - mp.matrix(vcs.tolist())
    x, y = (dv.T * mW * dv)[0], (mbt.T * mF * mbt)[0]
    ds = -mp.lu_solve(mL, mbt)
    d2s = -2 * mp.lu_solve(mL, ds)
    dksi = 2 * (mbt.T * mF * ds)[0]
    d2ksi = 2 * (ds.T * mF * ds)[0] + 2 * (mbt.T * mF * d2s)[0]
    curv = -abs(1 / dksi / (1 + lam * lam) ** mp.mpf("1.5"))
    dcurv = -d2ksi / dksi ** 2 * (1 + lam * lam) ** mp.mpf("-1.5") - 3 * lam / dksi * (1 + lam * lam) ** mp.mpf("-2.5")
    r = O.tikhonov_lcurve(G, vcs, err, F, 1e-3)
    np.testing.assert_allclose(r["bcs"], f(mbt).ravel(), rtol=1e-8)
    np.testing.assert_allclose(r["cov"], f(mCt), rtol=1e-7, atol=1e-10 * np.abs(r["cov"]).max())
    for got, exact, tol in ((r["x"], x, 1e-7), (r["y"], y, 1e-8), (r["curv"], curv, 1e-7), (r["dcurv"], dcurv, 1e-6)):
        assert got == pytest.approx(float(exact), rel=tol)
    # --- TSVD (k = n - 2): minimum-norm least squares through the exact SVD
    U, S, Vt = mp.svd_r(mG)
    k = n - 2
    z = U.T * mp.matrix(vcs.tolist())
    bt = mp.matrix(n, 1)
    for m in range(k):
        for i in range(n):
            bt[i] += Vt[m, i] * z[m] / S[m]
    np.testing.assert_allclose(O.tsvd_solve(G, vcs, err, k)[0], f(bt).ravel(), rtol=1e-8, atol=1e-10)
    mp.mp.dps = 15


def test_device_quadrature_tables_against_numpy_and_scipy():
    """The constants the CUDA kernels integrate with (isrsolver_b200/csrc/gl_tables.cuh, generated at 80 digits) parsed
    back from the header and checked independently of the generator: Gauss-Legendre nodes / weights against
    numpy.polynomial.legendre.leggauss, the 6-point Gauss-Hermite rule against numpy.polynomial.hermite.hermgauss, the
    15-point Kronrod rule with its embedded 7-point Gauss weights against the exactness conditions (Kronrod: degree 22,
    Gauss: degree 13) -- every table to 1e-15."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "isrsolver_b200", "csrc", "gl_tables.cuh")).read()
    tabs = {m.group(1): np.array([float(v) for v in m.group(3).replace("\n", " ").split(",") if v.strip()])
            for m in re.finditer(r"double (k\w+)\[(\d+)\] = \{(.*?)\};", hdr, re.S)}
    assert tabs, "no tables parsed"
    for n in (8, 16, 32):
        x, w = np.polynomial.legendre.leggauss(n)
        np.testing.assert_allclose(tabs[f"kGL{n}_x"], x, rtol=0, atol=1e-15)
        np.testing.assert_allclose(tabs[f"kGL{n}_w"], w, rtol=2e-13, atol=1e-16)   # (numpy's own weights are good to ~1e-13)
        for p in range(0, 2 * n, max(1, n // 4)):      # exactness through degree 2n - 1: the sharper check
            exact = 0.0 if p % 2 else 2.0 / (p + 1)
            assert abs(np.dot(tabs[f"kGL{n}_w"], tabs[f"kGL{n}_x"] ** p) - exact) < 2e-15, (n, p)
    t, w = np.polynomial.hermite.hermgauss(6)                  # physicists' rule, as gsl_integration_fixed_hermite
    order = np.argsort(tabs["kGH6_t"])
    np.testing.assert_allclose(tabs["kGH6_t"][order], t, rtol=0, atol=1e-15)
    np.testing.assert_allclose(tabs["kGH6_w_over_sqrtpi"][order], w / np.sqrt(np.pi), rtol=1e-14)
    x, wk, wg = tabs["kGK15_x"], tabs["kGK15_wk"], tabs["kGK15_wg"]
    assert len(x) == 15 and np.count_nonzero(wg) == 7
    for p in range(0, 23):          # Kronrod: exact through degree 22
        exact = 0.0 if p % 2 else 2.0 / (p + 1)
        assert abs(np.dot(wk, x ** p) - exact) < 2e-15, p
    for p in range(0, 14):          # embedded 7-point Gauss rule: exact through degree 13
        exact = 0.0 if p % 2 else 2.0 / (p + 1)
        assert abs(np.dot(wg, x ** p) - exact) < 2e-15, p
    xg, wgl = np.polynomial.legendre.leggauss(7)
    np.testing.assert_allclose(np.sort(x[wg != 0]), xg, rtol=0, atol=1e-15)


def test_high_precision_yardsticks_agree(golden, oracle):
    """tests/hp_reference.py: the long-double evaluation used as the gate at full size (N = 200) against the literal
    formulas of src/ISRSolverTikhonov.cpp:61-72, 139-155 / src/Chi2Test.cpp:51-55 in 50-digit arithmetic, and both
    against the NumPy literal-matrix oracle."""
    from hp_reference import ld_solve, ld_tikhonov_chi2, mp_tikhonov
    g = golden("c4_tikhonov_n32_spread")
    G, err, F, V, b0 = g["G_spread"], g["vcs_err"], g["tik_F"], g["toys"][:5], g["bcs0"]
    for lam in (1e-9, 1e-3, 0.3):
        m = mp_tikhonov(G, g["vcs"], err, F, lam, V, b0)
        c, B = ld_tikhonov_chi2(G, err, F, lam, V, b0)
        np.testing.assert_allclose(c, m["chi2"], rtol=1e-13)
        np.testing.assert_allclose(B, m["B"], rtol=1e-13, atol=1e-15)
        r = oracle.tikhonov_lcurve(G, g["vcs"], err, F, lam)
        np.testing.assert_allclose(r["bcs"], m["bcs"], rtol=1e-10)
        assert r["y"] == pytest.approx(m["y"], rel=1e-10) and r["curv"] == pytest.approx(m["curv"], rel=1e-9)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((40, 40)) + 5 * np.eye(40); Bm = rng.standard_normal((40, 3))
    np.testing.assert_allclose(np.asarray(ld_solve(A, Bm), dtype=float), np.linalg.solve(A, Bm), rtol=1e-12)
