"""BASELINE.json configs C2-C5 at their full sizes: toy-level comparisons with the oracle's linear-algebra layer (which
finishes thousands of toys at N = 100 ... 500 in seconds), matrix rows against the assembly oracle and the compiled
reference (a whole N = 500 matrix takes the CPU minutes), and size-independent properties."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _model(ecm, G):
    from bench import bw_born
    b0 = bw_born(ecm)
    vcs = G @ b0
    return b0, vcs, 0.05 * vcs


def test_c4_tikhonov_1024_lambdas_10k_toys():
    """C4: N = 200, cubic interpolation, 2 MeV energy spread, 1024 lambdas x 10k toys."""
    import isrsolver_b200 as isr
    n, T, L = 200, 10000, 1024
    ecm = np.linspace(1.18, 2.0, n)
    s = isr.ISRSolverTikhonov(n, ecm, np.ones(n), np.full(n, 0.002), np.ones(n), 0.827, enableEnergySpread=True)
    s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    s.eval_equation_matrix()
    G = s.intop_matrix().copy()
    b0, vcs, err = _model(ecm, G)
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    lam = 1e-9 * (1e9) ** (np.arange(L) / (L - 1))            # src/isrsolver-LCurve-plot.cpp:127-132
    V = isr.draw_toys(vcs, err, T)
    t0 = time.perf_counter()
    chi2 = s.toy_scan(lam, V, b0)
    dt = time.perf_counter() - t0
    assert chi2.shape == (L, T) and np.all(np.isfinite(chi2)) and np.all(chi2 >= 0)
    # (1) against the ORACLE at full size: 4 lambdas x 1024 toys.  The yardstick is the reference's formula
    # (b_k - b0)^T Cov_lambda^-1 (b_k - b0), b_k = T^-1 G^T W v_k, in 80-bit long double (tests/hp_reference.py, pinned by
    # 50-digit arithmetic at N = 32); the double-precision literal matrices of oracle.tikhonov_solve are compared too,
    # with the slack their explicit covariance inverse needs (reported, not assumed).
    from hp_reference import ld_tikhonov_chi2
    from oracle import isr_oracle as O
    F = s.regulariser()
    Fo = O.tikhonov_F(O.Interp(ecm, 0.827, [(False, 0, 0), (True, 1, n - 1)]).deriv_projector(),
                      O.Interp(ecm, 0.827, [(False, 0, 0), (True, 1, n - 1)]).integral_basis(), True)
    np.testing.assert_allclose(F, Fo, rtol=1e-8, atol=1e-10 * np.abs(Fo).max())
    worst_np = 0.0
    for q in (0, 400, 700, 1023):
        ref, Bref = ld_tikhonov_chi2(G, err, Fo, lam[q], V[:1024], b0)
        np.testing.assert_allclose(chi2[q, :1024], ref, rtol=RTOL)
        t = O.tikhonov_solve(G, vcs, err, Fo, lam[q])
        D = V[:1024] @ t["Ap"].T - b0[None, :]
        lit = np.einsum("ij,ij->i", D, np.linalg.solve(t["cov"], D.T).T)
        worst_np = max(worst_np, float(np.abs(lit / ref - 1).max()))
        np.testing.assert_allclose(chi2[q, :1024], lit, rtol=max(RTOL, 10 * worst_np))
        # (2) the O(N)-per-pair scan against the generic operator path (A+_lambda, Cov_lambda^-1 through the tile kernel)
        s.reg_param = float(lam[q])
        r = s.toy_batch(V[:2048], b0, want_bcs=True)
        np.testing.assert_allclose(chi2[q, :2048], r["chi2"], rtol=RTOL)
        np.testing.assert_allclose(r["bcs"][:1024], Bref, rtol=RTOL, atol=RTOL * np.abs(Bref).max())
    print(f"\nC4: double-precision literal oracle vs long double: worst relative chi2 deviation {worst_np:.2e}")
    # L-curve over the same grid: finite, and the residual norm grows / the smoothness norm falls with lambda
    lc = s.make_LCurve(lam)
    assert np.all(np.isfinite(lc["x"])) and np.all(np.isfinite(lc["y"])) and np.all(lc["c"] >= 0)
    assert np.all(np.diff(lc["x"]) >= -1e-9 * lc["x"][1:]) and np.all(np.diff(lc["y"]) <= 1e-9 * lc["y"][:-1])
    print(f"\nC4: {L} x {T} (lambda, toy) chi2 values through the host API in {dt * 1e3:.1f} ms")


def test_c5_sle_ratio_and_chi2_n500():
    """C5: N = 500, linear interpolation, 2^18 toys (the per-GPU share of 1M toys on 4 GPUs), ratio + chi2."""
    import isrsolver_b200 as isr
    n, T = 500, 1 << 18
    ecm = np.linspace(1.18, 2.0, n)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    s.eval_equation_matrix()
    G = s.intop_matrix().copy()
    assert np.all(np.triu(G, 1) == 0)
    b0, vcs, err = _model(ecm, G)
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    V = isr.draw_toys(vcs, err, T)
    r = s.toy_batch(V, b0, want_sum=True, want_ratio=True, want_hist=True)
    z = (V - vcs[None, :]) / err[None, :]
    np.testing.assert_allclose(r["chi2"], (z * z).sum(1), rtol=RTOL)          # survey A.7 identity
    assert r["chi2"].mean() == pytest.approx(n, rel=5e-3)
    np.testing.assert_allclose(r["sum_bcs"], np.linalg.solve(G, V.sum(0)), rtol=RTOL)   # linearity
    # ratio histograms: every toy lands in exactly one of the 1026 cells of its point; in-range count matches
    hist = r["ratio_hist"]
    assert np.all(hist.sum(1) == T)
    np.testing.assert_array_equal(hist[:, 1:1025].sum(1), r["ratio_stats"][:, 0].astype(np.int64))
    # means of (b - b0)/b0 are unbiased within 5 sigma of their own error
    cnt, sx, sx2 = r["ratio_stats"].T
    mean = sx / cnt
    sig = np.sqrt(np.abs(sx2 / cnt - mean ** 2) / cnt)
    assert np.all(np.abs(mean) < 6 * sig + 1e-12)
    # first rows against NumPy
    B = np.linalg.solve(G, V[:256].T).T
    rr = (B - b0[None, :]) / b0[None, :]
    s2 = s.toy_batch(V[:256], b0, want_ratio=True, want_bcs=True)
    np.testing.assert_allclose(s2["bcs"], B, rtol=RTOL, atol=1e-12)
    inr = (rr >= -2) & (rr < 2)
    np.testing.assert_allclose(s2["ratio_stats"][:, 1], np.where(inr, rr, 0).sum(0), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("n,cubic,T", [(100, True, 100000), (500, False, 1 << 17)])
def test_c2_c5_sizes_toy_level_vs_oracle(n, cubic, T, oracle):
    """BASELINE C2 (N = 100, cubic, 10^5 toys) and C5 (N = 500, linear, ratio + chi2) at their own N: the whole batch runs
    on the GPU; 8192 toys spread over it are compared toy by toy with the oracle -- b_k and chi2_k to 1e-8, the ratio
    histogram cells and the chi2 TH1F bit for bit."""
    import isrsolver_b200 as isr
    O = oracle
    ecm = np.linspace(1.18, 2.0, n)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    if cubic:
        s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    s.eval_equation_matrix()
    G = s.intop_matrix().copy()
    b0, vcs, err = _model(ecm, G)
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    V = isr.draw_toys(vcs, err, T)
    r = s.toy_batch(V, b0, want_sum=True, want_ratio=True)
    sel = np.unique(np.concatenate([np.arange(4096), np.linspace(0, T - 1, 4096).astype(int)]))
    c2, B = O.chi2_test_fair(G, V[sel], b0, err)
    np.testing.assert_allclose(r["chi2"][sel], c2, rtol=RTOL)
    # the reference's per-toy O(N^3) loop on a few of them (COD + G^T W G + inverse per toy, src/Chi2Test.cpp:38-60)
    c2l, _ = O.chi2_test_loop(G, V[sel[:6]], b0, err)
    np.testing.assert_allclose(r["chi2"][sel[:6]], c2l, rtol=RTOL)
    # a batch of exactly the selected toys: solutions, ratio cells, chi2 cells
    q = s.toy_batch(np.ascontiguousarray(V[sel]), b0, want_sum=True, want_ratio=True, want_hist=True, want_bcs=True)
    np.testing.assert_array_equal(q["chi2"], r["chi2"][sel])                      # a toy's chi2 does not depend on its batch
    np.testing.assert_allclose(q["bcs"], B, rtol=RTOL, atol=RTOL * np.abs(b0).max())
    np.testing.assert_allclose(q["sum_bcs"], B.sum(0), rtol=RTOL)
    means, errs, counts, stats = O.ratio_test(q["bcs"], b0)                       # ROOT bin arithmetic on the GPU's b_k
    np.testing.assert_array_equal(q["ratio_hist"], counts.astype(np.uint32))
    np.testing.assert_allclose(q["ratio_stats"], stats, rtol=1e-9, atol=1e-9)
    # cells from the oracle's OWN b_k: identical except where a ratio sits within rounding of a cell edge
    counts_o = O.ratio_test(B, b0)[2]
    assert np.abs(q["ratio_hist"].astype(np.int64) - counts_o.astype(np.int64)).sum() <= 4
    ref, _, lo, hi = O.chi2_histogram(q["chi2"])
    import ctypes as C
    from isrsolver_b200 import _lib as L
    cnt = np.zeros(130, dtype=np.uint32); clo, chi = C.c_double(), C.c_double()
    L.check(L.load().isr_last_chi2_minmax(s._p, C.byref(clo), C.byref(chi)))
    L.check(L.load().isr_last_chi2_histogram(s._p, 128, clo.value, chi.value, L.uptr(cnt)))
    assert (clo.value, chi.value) == (lo, hi)
    np.testing.assert_array_equal(cnt, ref.astype(np.uint32))


def test_c3_sle2_matrix_assembly_timing():
    """C3: N = 200 with a 512-bin efficiency histogram per point: assembly time and panel count."""
    import ctypes as C
    import isrsolver_b200 as isr
    from isrsolver_b200 import _lib as L
    n = 200
    ecm = np.linspace(1.18, 2.0, n)
    xc = (np.arange(512) + 0.5) / 512
    vals = np.zeros((n, 514))
    vals[:, 1:-1] = 0.25 + 0.5 * np.exp(-40.0 * xc)[None, :] * (1.0 - 0.1 * (ecm[:, None] - 1.5))
    s = isr.ISRSolverSLE2(n, ecm, np.ones(n), None, np.ones(n), 0.827, eff_hists=vals, xlo=0.0, xhi=1.0)
    s.eval_equation_matrix()
    t0 = time.perf_counter()
    for _ in range(5):
        s.eval_equation_matrix()
    dt = (time.perf_counter() - t0) / 5
    npan, nnode = C.c_int64(), C.c_int64()
    L.check(L.load().isr_assemble_stats(s._p, C.byref(npan), C.byref(nnode)))
    G = s.intop_matrix()
    assert np.all(np.triu(G, 1) == 0) and np.all(G[np.tril_indices(n)] > 0)
    # an efficiency <= 0.75 everywhere scales every entry below the eps = 1 matrix
    s1 = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    s1.eval_equation_matrix()
    G1 = s1.intop_matrix()
    tri = np.tril_indices(n)
    assert np.all(G[tri] < 0.7501 * G1[tri]) and np.all(G[tri] > 0.2499 * G1[tri])
    assert dt < 0.02
    print(f"\nC3: N=200 eps 512xN assembly {dt * 1e3:.3f} ms wall per call incl. D2H of G, {npan.value} panels, {nnode.value} nodes")


# ---- full-size matrix rows against the CPU side ---------------------------------------------------------
# The oracle (and the compiled reference, oracle/_ref, where that prebuilt library travelled) finish a few
# ROWS of the full-size matrices in seconds; the CUDA assembly is compared on those rows entry by entry.
FULL_CASES = {
    "bench_n200_cubic": (200, [(False, 0, 0), (True, 1, 199)], None, [3, 101, 199]),
    "c2_n100_cubic": (100, [(False, 0, 0), (True, 1, 99)], None, [0, 50, 99]),
    "c3_n200_eps512": (200, None, 512, [3, 101, 199]),
    "c5_n500_linear": (500, None, None, [7, 250, 499]),
}


@pytest.mark.parametrize("case", sorted(FULL_CASES))
def test_full_size_rows_vs_oracle_and_compiled_reference(case, oracle):
    import isrsolver_b200 as isr
    from conftest import assert_matrix_parity
    from oracle import isr_ref as R
    O = oracle
    n, settings, nbins, rows = FULL_CASES[case]
    ecm = np.linspace(1.18, 2.0, n)
    eff = O.row_table_efficiency(ecm, nbins) if nbins else None
    if eff is not None:
        s = isr.ISRSolverSLE2(n, ecm, np.ones(n), None, np.ones(n), 0.827, eff_hists=eff.values, xlo=0.0, xhi=1.0)
    else:
        s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    if settings:
        s.set_interp_settings(settings)
    s.eval_equation_matrix()
    G = s.intop_matrix()[rows]
    Gc, Ec = O.Interp(ecm, 0.827, settings).eval_rows(rows, eff=eff, mode="converged")
    assert_matrix_parity(G, Gc, Ec, what=case + " rows vs converged oracle")       # 1e-9 relative, entry by entry
    if R.available():
        Gr = R.RefInterp(ecm, 0.827, settings, v2=eff is not None).eval_eq_matrix(eff=eff, rows=rows)[rows]
        # the reference's own accuracy: 1e-9 of the row scale for smooth integrands; with the 512-bin table its
        # retry loop only converges at epsrel ~1e-4 (measured |ref - converged| up to 4e-8 here)
        assert_matrix_parity(G, Gr, None, row_rel=(2e-7 if nbins else 1e-9), what=case + " rows vs the compiled reference")
