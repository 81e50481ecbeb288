"""Edge cases of the hot path on the B200: minimal and large N, energies hugging the threshold (x_max below the
pair threshold x0), very unequal spacing, tiny and huge toy counts relative to the tile, odd N in every kernel."""
import numpy as np
import pytest

from conftest import assert_matrix_parity

pytestmark = pytest.mark.gpu
THR = 0.827


def _solver(ecm, settings=None, cls=None, **kw):
    import isrsolver_b200 as isr
    n = len(ecm)
    s = (cls or isr.ISRSolverSLE)(n, ecm, np.ones(n), None, np.ones(n), THR, **kw)
    if settings:
        s.set_interp_settings(settings)
    return s


@pytest.mark.parametrize("n", [2, 3, 5])
def test_minimal_sizes(oracle, n):
    ecm = np.linspace(1.0, 1.3, n)
    for st in (None, [(True, 0, n - 1)]):
        s = _solver(ecm, st)
        s.eval_equation_matrix()
        Gc, Ec, _ = oracle.Interp(ecm, THR, st).eval_eq_matrix(mode="converged", with_errors=True)
        assert_matrix_parity(s.intop_matrix(), Gc, Ec, what=f"N={n} {st}")
        b0 = np.linspace(1.0, 2.0, n); vcs = Gc @ b0; err = 0.05 * vcs
        s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
        s.solve()
        np.testing.assert_allclose(s.bcs(), np.linalg.solve(s.intop_matrix(), vcs), rtol=1e-10)
        V = oracle.draw_toys(vcs, err, 33)
        c2, B = oracle.chi2_test_fair(s.intop_matrix(), V, b0, err)
        r = s.toy_batch(V, b0, want_bcs=True)
        np.testing.assert_allclose(r["chi2"], c2, rtol=1e-8)
        np.testing.assert_allclose(r["bcs"], B, rtol=1e-8, atol=1e-12)


def test_energies_hugging_the_threshold(oracle):
    """x_max = 1 - (thr/E)^2 far below the pair threshold x0 = 4 m_e / E for the first points: only the
    t = x^beta panels exist there; the matrix must still match the adaptive oracle."""
    ecm = THR + np.array([1e-7, 1e-5, 5e-4, 2e-3, 0.05, 0.4])
    for st in (None, [(False, 0, 1), (True, 2, 5)]):
        s = _solver(ecm, st)
        s.eval_equation_matrix()
        Gc, Ec, _ = oracle.Interp(ecm, THR, st).eval_eq_matrix(mode="converged", with_errors=True)
        assert_matrix_parity(s.intop_matrix(), Gc, Ec, what=f"threshold-hugging {st}")
        assert np.all(np.isfinite(s.intop_matrix())) and np.all(np.diag(s.intop_matrix()) > 0)


def test_very_unequal_spacing_and_high_energy(oracle):
    rng = np.random.default_rng(17)
    ecm = np.sort(np.concatenate([THR + 10.0 ** rng.uniform(-4, 0.5, 20), [10.58, 91.2]]))
    st = [(False, 0, 4), (True, 5, len(ecm) - 1)]
    s = _solver(ecm, st)
    s.eval_equation_matrix()
    Gc, Ec, _ = oracle.Interp(ecm, THR, st).eval_eq_matrix(mode="converged", with_errors=True)
    assert_matrix_parity(s.intop_matrix(), Gc, Ec, rel=1e-9, what="unequal spacing up to 91 GeV")


@pytest.mark.parametrize("n", [999, 1000])
def test_large_n_rows_and_toys(oracle, n):
    """N = 1000 (5x the headline size; odd N exercises the padded leading dimension): sampled rows against the
    oracle, toys against NumPy."""
    ecm = np.linspace(1.0, 3.0, n)
    s = _solver(ecm)
    s.eval_equation_matrix()
    G = s.intop_matrix()
    rows = [0, 1, n // 2, n - 1]
    Gc, Ec = oracle.Interp(ecm, THR).eval_rows(rows, mode="converged")
    assert_matrix_parity(G[rows], Gc, Ec, what=f"N={n} rows")
    assert np.all(np.triu(G, 1) == 0)
    b0 = 1.0 + np.sin(ecm); vcs = G @ b0; err = 0.03 * vcs
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    V = oracle.draw_toys(vcs, err, 777)
    r = s.toy_batch(V, b0, want_sum=True, want_bcs=True)
    B = np.linalg.solve(G, V.T).T
    np.testing.assert_allclose(r["bcs"], B, rtol=1e-8, atol=1e-9 * np.abs(B).max())
    z = (V - vcs) / err
    np.testing.assert_allclose(r["chi2"], (z * z).sum(1), rtol=1e-8)
    np.testing.assert_allclose(r["sum_bcs"], B.sum(0), rtol=1e-8, atol=1e-8 * np.abs(B.sum(0)).max())
    # ratio statistics at this size need the smallest stage ring (the per-CTA accumulators take 128 KB)
    r4 = s.toy_batch(V[:300], b0, want_chi2=False, want_ratio=True, want_hist=True)
    rr = (B[:300] - b0[None, :]) / b0[None, :]
    inr = (rr >= -2) & (rr < 2)
    np.testing.assert_array_equal(r4["ratio_stats"][:, 0], inr.sum(0))
    np.testing.assert_allclose(r4["ratio_stats"][:, 1], np.where(inr, rr, 0).sum(0), rtol=1e-7, atol=1e-9)
    assert np.all(r4["ratio_hist"].sum(1) == 300)


@pytest.mark.parametrize("T", [1, 7, 63, 64, 65, 148 * 64 - 1, 148 * 64 + 1])
def test_toy_counts_around_the_tile_and_grid(oracle, T):
    """T below one tile, exactly one tile, one more than a tile, and around one full wave of 148 CTAs."""
    n = 31
    ecm = np.linspace(1.0, 2.0, n)
    s = _solver(ecm, [(False, 0, 0), (True, 1, n - 1)])
    s.eval_equation_matrix()
    G = s.intop_matrix()
    b0 = oracle.bw_born(ecm); vcs = G @ b0; err = 0.05 * vcs
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    V = oracle.draw_toys(vcs, err, T)
    r = s.toy_batch(V, b0, want_sum=True, want_ratio=True)
    c2, B = oracle.chi2_test_fair(G, V, b0, err)
    np.testing.assert_allclose(r["chi2"], c2, rtol=1e-8)
    np.testing.assert_allclose(r["sum_bcs"], B.sum(0), rtol=1e-8)
    assert r["ratio_stats"][:, 0].max() <= T


def test_tikhonov_and_tsvd_minimal_and_odd(oracle):
    import isrsolver_b200 as isr
    for n in (3, 17):
        ecm = np.linspace(1.0, 2.0, n)
        I = oracle.Interp(ecm, THR)
        G = I.eval_eq_matrix(mode="converged")
        b0 = oracle.bw_born(ecm); vcs = G @ b0; err = 0.05 * vcs
        t = isr.ISRSolverTikhonov(n, ecm, vcs, None, err, THR)
        F = oracle.tikhonov_F(I.deriv_projector(), I.integral_basis(), True)
        for lam in (1e-6, 1e-2):
            ref = oracle.tikhonov_lcurve(G, vcs, err, F, lam)
            t.reg_param = lam
            t.solve()
            np.testing.assert_allclose(t.bcs(), ref["bcs"], rtol=1e-7, atol=1e-9 * np.abs(ref["bcs"]).max())
        V = oracle.draw_toys(vcs, err, 50)
        assert t.toy_scan([1e-4, 1e-1], V, b0).shape == (2, 50)
        v = isr.ISRSolverTSVD(n, ecm, vcs, None, err, THR, upper_TSVD_index=max(1, n - 1))
        v.solve()
        np.testing.assert_allclose(v.bcs(), oracle.tsvd_solve(G, vcs, err, max(1, n - 1))[0], rtol=1e-6, atol=1e-9)


def test_factor_falls_back_to_pivoting_and_rejects_singular(oracle):
    """isr_factor tries an inverse without pivoting first (the cluster Gauss-Jordan kernel for N <= 240, cuSOLVER's
    LU above) and accepts it only if max |G X - I| is tiny; matrices
    that need row exchanges (zero or tiny leading pivots), or are badly conditioned, must come out right through
    the pivoted path, and a singular matrix must be reported."""
    import isrsolver_b200 as isr
    from isrsolver_b200 import IsrGpuError
    rng = np.random.default_rng(12)
    n = 23
    ecm = np.linspace(1.0, 2.0, n)
    cases = {
        "zero leading pivot": np.roll(np.eye(n), 1, axis=0) * 2.0 + 0.01 * rng.random((n, n)),
        "tiny pivots (growth without exchanges)": np.triu(rng.random((n, n)), 1) + np.diag(np.full(n, 1e-9)) + np.tril(rng.random((n, n)), -1),
        "ill-conditioned (cond 1e9)": (lambda u, v: u @ np.diag(np.geomspace(1, 1e-9, n)) @ v)(*[np.linalg.qr(rng.random((n, n)))[0] for _ in range(2)]),
        "well conditioned dense": np.eye(n) * 3 + rng.random((n, n)),
    }
    for name, G in cases.items():
        s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), THR)
        s.set_matrix(G)
        b0 = 1.0 + rng.random(n); vcs = G @ b0; err = 0.05 * np.abs(vcs) + 1e-3
        s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
        s.solve()
        cond = np.linalg.cond(G)
        np.testing.assert_allclose(s.bcs(), b0, rtol=1e-13 * cond + 1e-9, atol=1e-12 * cond, err_msg=name)
        V = oracle.draw_toys(vcs, err, 64)
        r = s.toy_batch(V, b0, want_bcs=True)
        np.testing.assert_allclose(r["bcs"], np.linalg.solve(G, V.T).T, rtol=1e-12 * cond + 1e-8, atol=1e-11 * cond, err_msg=name)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), THR)
    G = rng.random((n, n)); G[5] = G[3]               # exactly singular
    s.set_matrix(G)
    with pytest.raises(IsrGpuError):
        s.solve()


@pytest.mark.parametrize("n", [2, 4, 15, 16, 17, 63, 64, 65, 128, 129, 236, 239, 240, 241, 256])
def test_inverse_sizes_around_the_cluster_kernel_limits(n):
    """The register-resident Gauss-Jordan inverse (4 CTAs x warps of 4 rows, pivot blocks of 4, 64-column pair groups,
    N <= 240) against NumPy on dense and on lower-triangular matrices, at sizes around every one of those boundaries
    and across the hand-over to cuSOLVER; a lower-triangular matrix must come back exactly lower triangular
    (the toy kernel's triangular schedule relies on it)."""
    import isrsolver_b200 as isr
    rng = np.random.default_rng(3100 + n)
    ecm = np.linspace(1.0, 2.0, n) if n > 1 else np.array([1.5])
    for shape in ("dense", "lower"):
        G = np.eye(n) * 2.0 + np.tril(rng.random((n, n))) * 0.3
        if shape == "dense":
            G += 0.05 * rng.random((n, n))
        b0 = 1.0 + rng.random(n); vcs = G @ b0; err = 0.05 + 0.1 * rng.random(n)
        s = isr.ISRSolverSLE(n, ecm, vcs, None, err, THR)
        s.set_matrix(G)
        s.solve()
        np.testing.assert_allclose(s.bcs(), b0, rtol=1e-12)
        Ginv = np.linalg.inv(G)
        cov = (Ginv * err[None, :]) @ (Ginv * err[None, :]).T
        np.testing.assert_allclose(s.bcs_cov_matrix(), cov, rtol=1e-10, atol=1e-13 * np.abs(cov).max())
        V = vcs[None, :] + err[None, :] * rng.standard_normal((37, n))
        r = s.toy_batch(V, b0, want_bcs=True)
        np.testing.assert_allclose(r["bcs"], V @ Ginv.T, rtol=1e-11, atol=1e-12)
        if shape == "lower":
            # v = e_0 scaled: b must vanish exactly nowhere above ... equivalently a toy with only the LAST point excited
            # leaves every earlier b exactly untouched iff G^-1 is exactly lower triangular
            e = np.zeros((1, n)); e[0, -1] = 1.0
            be = s.toy_batch(e, np.zeros(n), want_bcs=True)["bcs"][0]
            assert np.all(be[:-1] == 0.0)


def test_inverse_every_size_through_the_cluster_kernel():
    """Every N from 2 to 260: the Gauss-Jordan kernel's slot / parity bookkeeping depends on N through the number of
    pivot blocks, the rows per CTA and which CTA owns which block -- a wrong phase would hang or corrupt, so each size
    is factored (dense and lower triangular) and checked against NumPy."""
    import isrsolver_b200 as isr
    rng = np.random.default_rng(77)
    for n in range(2, 261):
        ecm = np.linspace(1.0, 2.0, n)
        for lower in (False, True):
            G = np.eye(n) * 2.0 + np.tril(rng.random((n, n))) * 0.3
            if not lower:
                G += 0.05 * rng.random((n, n))
            b0 = 1.0 + rng.random(n); err = 0.05 + 0.1 * rng.random(n)
            s = isr.ISRSolverSLE(n, ecm, G @ b0, None, err, THR)
            s.set_matrix(G)
            s.solve()
            assert np.allclose(s.bcs(), b0, rtol=1e-12, atol=0), (n, lower)
            Ginv = np.linalg.inv(G)
            cov = (Ginv * err[None, :]) @ (Ginv * err[None, :]).T
            assert np.abs(s.bcs_cov_matrix() - cov).max() <= 1e-10 * np.abs(cov).max(), (n, lower)


@pytest.mark.parametrize("n", [2, 7, 64, 129, 200, 257, 300])
def test_batched_jacobi_singular_values(n):
    """The batched one-sided Jacobi kernel behind isr_cond_spread_scan (svd_jacobi.cu; the reference does a JacobiSVD per
    spread setting, src/isrsolver-condnum-test.cpp:88-104) on injected matrices -- sigma = 0 settings return the singular
    values of the matrix itself -- against LAPACK: well conditioned, graded (cond 1e10) and rank deficient."""
    import ctypes as C
    from isrsolver_b200 import _lib as L
    from isrsolver_b200.pyisr import _ctx
    lib = L.load()
    rng = np.random.default_rng(n)
    ecm = np.linspace(1.2, 2.0, n)
    st = np.ascontiguousarray([[0, 0, n - 1]], dtype=np.int32)
    p = C.c_void_p()
    L.check(lib.isr_problem_create(_ctx(), n, L.dptr(ecm), 0.827, 1, L.iptr(st), None, C.byref(p)))
    U, _ = np.linalg.qr(rng.standard_normal((n, n))); Vt, _ = np.linalg.qr(rng.standard_normal((n, n)))
    mats = [rng.standard_normal((n, n)) + 2 * np.eye(n),
            (U * np.geomspace(1.0, 1e-10, n)) @ Vt,
            np.tril(rng.uniform(0.1, 1.0, (n, n)))]
    if n > 2:
        R = rng.standard_normal((n, n)); R[:, -1] = R[:, 0] + R[:, 1]
        mats.append(R)
    for A in mats:
        L.check(lib.isr_set_matrix(p, L.dptr(np.ascontiguousarray(A.T))))
        K = 3
        cond, sv = np.zeros(K), np.zeros((K, n))
        L.check(lib.isr_cond_spread_scan(p, K, L.dptr(np.zeros(K)), 0, L.dptr(cond), L.dptr(sv)))
        ref = np.linalg.svd(A, compute_uv=False)
        for k in range(K):
            assert np.abs(sv[k] - ref).max() <= 1e-10 * ref[0], (n, np.abs(sv[k] - ref).max() / ref[0])
            assert np.all(np.diff(sv[k]) <= 0)
        np.testing.assert_array_equal(sv[0], sv[1])            # every CTA of the batch runs the same deterministic schedule
        if ref[-1] > 1e-9 * ref[0]:
            assert cond[0] == pytest.approx(ref[0] / ref[-1], rel=1e-6)
    lib.isr_problem_destroy(p)


def test_cond_spread_scan_100_settings_at_n200(oracle):
    """The notebook's scan (notebooks/condnum_enspread.ipynb: 100 spread settings) at N = 200 in one batched call."""
    import time
    import isrsolver_b200 as isr
    n = 200
    ecm = np.linspace(1.18, 2.0, n)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    sig = np.linspace(0.0, 0.005, 100)
    s.condnum_spread_scan(sig)
    t0 = time.perf_counter()
    cond, sv = s.condnum_spread_scan(sig, want_singular_values=True)
    dt = time.perf_counter() - t0
    s.eval_equation_matrix()
    G0 = s.intop_matrix()
    I = oracle.Interp(ecm, 0.827, [(False, 0, 0), (True, 1, n - 1)])
    for k in (0, 1, 37, 99):
        Gk = G0 if sig[k] == 0 else I.energy_spread_matrix(np.full(n, sig[k])) @ G0
        ref = np.linalg.svd(Gk, compute_uv=False)
        assert np.abs(sv[k] - ref).max() <= 1e-10 * ref[0]
        assert cond[k] == pytest.approx(ref[0] / ref[-1], rel=1e-8 * max(1.0, ref[0] / ref[-1] * 1e-4))
    assert np.all(np.diff(cond[:30]) > 0)         # (beyond a spread of about a third of the point spacing it fluctuates)
    print(f"\ncond-vs-spread scan, N = 200, 100 settings: {dt * 1e3:.1f} ms (round 1: 1.1 s through 100 x cuSOLVER gesvd)")
    assert dt < 0.5          # (39-47 ms measured; a loose bound: the first call of a process also pays allocations)
