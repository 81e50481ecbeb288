"""Tikhonov lambda-scan / L-curve and TSVD parity on the B200 (through the C ABI).

The oracle is the literal matrix implementation of src/ISRSolverTikhonov.cpp / src/ISRSolverTSVD.cpp in
NumPy.  Every gate is the north-star 1e-8; where the double-precision literal route is itself noisier than that
(the residual norm at tiny lambda, the curvature derivative) the yardstick is the same formula in 50-digit mpmath
arithmetic (tests/hp_reference.py), so the GPU is measured against the true value and not against the oracle's rounding."""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_assembly import make_solver

pytestmark = pytest.mark.gpu


def tik_solver(g):
    import isrsolver_b200 as isr
    s = make_solver(g, cls=isr.ISRSolverTikhonov, spread=True)
    s.set_matrix(g["G_spread"])
    return s


def test_regulariser_and_basis_tables():
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    n = len(g["ecm"])
    np.testing.assert_allclose(s.regulariser(), g["tik_F"], rtol=1e-8, atol=1e-10 * np.abs(g["tik_F"]).max())
    # D(i,j) = basis_j'(E_i) and w_j = int basis_j
    jj, ii = np.meshgrid(np.arange(n), np.arange(n))
    D = s.basis_eval(jj.ravel(), g["ecm"][ii.ravel()], deriv=True).reshape(n, n)
    np.testing.assert_allclose(D, g["D"], rtol=1e-9, atol=1e-9 * np.abs(g["D"]).max())
    import ctypes as C
    from isrsolver_b200 import _lib as L
    w = np.empty(n); L.check(L.load().isr_basis_integrals(s._p, L.dptr(w)))
    np.testing.assert_allclose(w, g["w"], rtol=1e-10)


def test_lcurve_scan_vs_literal_matrices(oracle):
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    lam = g["tik_lambda"]
    res = s.make_LCurve(lam)
    np.testing.assert_allclose(res["y"], g["tik_y"], rtol=1e-8)
    np.testing.assert_allclose(res["c"], -g["tik_curv"], rtol=1e-8)
    x, y, c, dc, B = s._scan(lam, want_bcs=True)
    np.testing.assert_allclose(B, g["tik_bcs"], rtol=1e-8, atol=1e-8 * np.abs(g["tik_bcs"]).max())
    # evalEqNorm2 is a residual: the reference (and the NumPy oracle stored in the fixture) forms G b - v in double
    # precision and loses eps |v| / |G b - v| (8e-6 at lambda = 1e-9), the curvature derivative goes through two LU solves
    # with L.  The gate for both is therefore the SAME literal formulas evaluated in 50-digit arithmetic
    # (tests/hp_reference.py): north-star tolerance 1e-8, no allowance for the oracle's own noise.
    from hp_reference import mp_tikhonov
    for q in range(len(lam)):
        m = mp_tikhonov(g["G_spread"], g["vcs"], g["vcs_err"], g["tik_F"], lam[q])
        assert res["x"][q] == pytest.approx(m["x"], rel=1e-8), (lam[q], res["x"][q], m["x"])
        assert res["y"][q] == pytest.approx(m["y"], rel=1e-8)
        assert -res["c"][q] == pytest.approx(m["curv"], rel=1e-8)
        assert dc[q] == pytest.approx(m["dcurv"], rel=1e-8), (lam[q], dc[q], m["dcurv"])
        np.testing.assert_allclose(B[q], m["bcs"], rtol=1e-8, atol=1e-8 * np.abs(m["bcs"]).max())


def test_single_lambda_solve_and_cov():
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    s.reg_param = float(g["tik_lambda"][3])
    s.solve()
    np.testing.assert_allclose(s.bcs(), g["tik_bcs"][3], rtol=1e-8, atol=1e-8 * np.abs(g["tik_bcs"][3]).max())
    ref = g["tik_cov3"]
    assert np.abs(s.bcs_cov_matrix() - ref).max() <= 1e-8 * np.abs(ref).max()
    assert s.evalSmoothnessConstraintNorm2() == pytest.approx(float(g["tik_y"][3]), rel=1e-8)
    assert s.evalLCurveCurvature() == pytest.approx(float(g["tik_curv"][3]), rel=1e-8)


def test_diag_regulariser(oracle):
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    s.disableDerivNorm2Regularizator()
    F = oracle.tikhonov_F(g["D"], g["w"], False)
    np.testing.assert_allclose(s.regulariser(), F, rtol=1e-10, atol=1e-14)
    lam = 1e-3
    ref = oracle.tikhonov_lcurve(g["G_spread"], g["vcs"], g["vcs_err"], F, lam)
    s.reg_param = lam; s.solve()
    np.testing.assert_allclose(s.bcs(), ref["bcs"], rtol=1e-8, atol=1e-8 * np.abs(ref["bcs"]).max())
    assert s.evalLCurveCurvature() == pytest.approx(ref["curv"], rel=1e-8)


def test_toy_scan_vs_literal(oracle):
    """chi2 of (lambda, toy) pairs: O(N) diagonal rescaling vs the literal A+ / Cov^-1 matrices."""
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    lam = np.array([1e-4, 1e-2, 0.3])
    V, b0 = g["toys"][:48], g["bcs0"]
    from hp_reference import ld_tikhonov_chi2, mp_tikhonov
    got = s.toy_scan(lam, V, b0)
    for q, l in enumerate(lam):
        t = oracle.tikhonov_solve(g["G_spread"], g["vcs"], g["vcs_err"], g["tik_F"], l)
        D = V @ t["Ap"].T - b0[None, :]
        ref = np.einsum("ij,ij->i", D, np.linalg.solve(t["cov"], D.T).T)
        np.testing.assert_allclose(got[q], ref, rtol=1e-8)            # the NumPy oracle (literal A+ / Cov^-1 matrices)
        # ... and the literal formulas in 50-digit arithmetic, so that neither side's rounding is taken on trust
        m = mp_tikhonov(g["G_spread"], g["vcs"], g["vcs_err"], g["tik_F"], l, V[:12], b0)
        np.testing.assert_allclose(got[q][:12], m["chi2"], rtol=1e-8)
        # the long-double evaluation used at full size (test_gpu_full_configs) is itself pinned by the same numbers
        np.testing.assert_allclose(ld_tikhonov_chi2(g["G_spread"], g["vcs_err"], g["tik_F"], l, V[:12], b0)[0], m["chi2"], rtol=1e-12)
    # the selected-operator path (isr_tikhonov_select + generic toy batch) agrees with the scan
    s.reg_param = float(lam[1])
    r = s.toy_batch(V, b0, want_bcs=True)
    np.testing.assert_allclose(r["chi2"], got[1], rtol=1e-8)
    t = oracle.tikhonov_solve(g["G_spread"], g["vcs"], g["vcs_err"], g["tik_F"], lam[1])
    np.testing.assert_allclose(r["bcs"], V @ t["Ap"].T, rtol=1e-8, atol=1e-8 * np.abs(V).max())


def test_toy_scan_quadratic_form_equals_direct_kernel():
    """chi2(lambda, k) = A_k - 2 lambda B_k + lambda^2 C (default) against the direct O(N)-per-pair kernel over
    18 decades of lambda, including the bias-dominated end where lambda^2 C dwarfs A_k."""
    from isrsolver_b200 import _lib as L
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    lam = np.geomspace(1e-12, 1e6, 73)
    V, b0 = g["toys"], g["bcs0"]
    quad = s.toy_scan(lam, V, b0)
    L.check(L.load().isr_set_scan_kernel(s._p, 1))
    direct = s.toy_scan(lam, V, b0)
    L.check(L.load().isr_set_scan_kernel(s._p, 0))
    assert np.all(direct > 0)
    np.testing.assert_allclose(quad, direct, rtol=1e-10)
    np.testing.assert_array_equal(quad, s.toy_scan(lam, V, b0))     # deterministic


def test_toy_scan_hist_matches_host_binning(oracle):
    """Per-lambda TH1F(128, min, max) filled on the device == ROOT bin arithmetic on the chi2 rows."""
    from isrsolver_b200 import dist
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    lam = np.geomspace(1e-6, 1.0, 7)
    V, b0 = g["toys"], g["bcs0"]
    chi2 = s.toy_scan(lam, V, b0)
    r = s.toy_scan_hist(lam, V, b0)
    np.testing.assert_array_equal(r["min"], chi2.min(1))
    np.testing.assert_array_equal(r["max"], chi2.max(1))
    np.testing.assert_allclose(r["mean"], chi2.mean(1), rtol=1e-13)
    for q in range(len(lam)):
        np.testing.assert_array_equal(r["hist"][q], dist.th1_counts(chi2[q], 128, chi2[q].min(), chi2[q].max()).astype(np.float32))
        assert r["hist"][q].sum() == V.shape[0] and r["hist"][q][129] >= 1      # the maximum is in the overflow cell


def test_solve_lcurve_finds_curvature_maximum():
    g = load_golden("c4_tikhonov_n32_spread")
    s = tik_solver(g)
    res = s.solve_LCurve(1e-3)
    grid = np.geomspace(1e-12, 1.0, 4001)
    c = s.make_LCurve(grid)["c"]
    assert res["max_curv"] >= c.max() * (1 - 1e-6)
    assert res["lambda"] == pytest.approx(grid[np.argmax(c)], rel=2e-2)


def test_tsvd(oracle):
    import isrsolver_b200 as isr
    g = load_golden("c4_tikhonov_n32_spread")
    s = make_solver(g, cls=isr.ISRSolverTSVD, spread=True)
    s.set_matrix(g["G_spread"])
    np.testing.assert_allclose(s.singular_values(), g["sv"], rtol=1e-10)
    k = int(g["tsvd_k"])
    s.upper_TSVD_index = k
    s.solve()
    np.testing.assert_allclose(s.bcs(), g["tsvd_bcs"], rtol=1e-8, atol=1e-8 * np.abs(g["tsvd_bcs"]).max())
    s.enable_keepone(); s.solve()
    np.testing.assert_allclose(s.bcs(), g["tsvd_bcs_keepone"], rtol=1e-8, atol=1e-8 * np.abs(g["tsvd_bcs_keepone"]).max())
    s.disable_keepone(); s.upper_TSVD_index = len(g["ecm"]); s.solve()
    bo, covo = oracle.sle_solve(g["G_spread"], g["vcs"], g["vcs_err"])
    np.testing.assert_allclose(s.bcs(), bo, rtol=1e-8)
    assert np.abs(s.bcs_cov_matrix() - covo).max() <= 1e-8 * np.abs(covo).max()
    # toy batch with the truncated operator: b_k = K^+ v_k
    s.upper_TSVD_index = k
    r = s.toy_batch(g["toys"][:16], g["bcs0"], want_bcs=True)
    U, S, Vt = np.linalg.svd(g["G_spread"])
    Kp = Vt[:k].T @ np.diag(1 / S[:k]) @ U[:, :k].T
    np.testing.assert_allclose(r["bcs"], g["toys"][:16] @ Kp.T, rtol=1e-8, atol=1e-8 * np.abs(g["bcs0"]).max())


_LAMBDA_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(os.environ["RANK"])
os.environ["LOCAL_RANK"] = str(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())
import isrsolver_b200 as isr
dist.init_process_group("gloo")
g = np.load(sys.argv[2])
n = len(g["ecm"])
s = isr.ISRSolverTikhonov(n, g["ecm"], g["vcs"], g["ecm_err"], g["vcs_err"], 0.827, enableEnergySpread=True)
s.set_interp_settings([tuple(r) for r in g["settings"]])
lam = np.geomspace(1e-8, 1.0, 37)
r = s.toy_scan_hist_sharded(lam, g["toys"], g["bcs0"])
one = s.toy_scan_hist(lam, g["toys"], g["bcs0"])
for k in ("min", "max", "mean", "hist"):
    assert np.array_equal(r[k], one[k]), k
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_lambda_sharded_scan_equals_single_gpu(tmp_path):
    """C4's lambda sharding on real kernels: 3 ranks, 37 lambdas (ragged slices), one all_gather."""
    import os
    import subprocess
    import sys
    from conftest import GOLDEN, ROOT
    w = tmp_path / "worker.py"
    w.write_text(_LAMBDA_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
                          "--master-addr", "127.0.0.1", "--master-port", "29653", str(w), ROOT,
                          os.path.join(GOLDEN, "c4_tikhonov_n32_spread.npz")],
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 3
