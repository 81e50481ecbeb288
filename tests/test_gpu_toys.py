"""Solve / covariance / toy-batch / histogram parity on the B200 (through the C ABI).

Tolerances (north star): Born cross section, covariance and per-toy chi2 within 1e-8 relative; histogram
and bin indices bit-exact on identical chi2 inputs."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from test_gpu_assembly import make_solver

pytestmark = pytest.mark.gpu
RTOL = 1e-8


@pytest.mark.parametrize("name", ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c5_mixed3_n36"])
def test_solve_and_covariance(name):
    g = load_golden(name)
    s = make_solver(g)
    s.solve()
    np.testing.assert_allclose(s.bcs(), g["bcs"], rtol=RTOL)
    cov, ref = s.bcs_cov_matrix(), g["cov"]
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert np.abs(cov - ref).max() <= RTOL * scale.max()
    np.testing.assert_allclose(np.diag(cov), np.diag(ref), rtol=RTOL)
    assert np.array_equal(cov, cov.T)
    assert s.condnum_eval() == pytest.approx(float(g["cond"]), rel=1e-8)
    inv = s.bcs_inv_cov_matrix()
    np.testing.assert_allclose(inv @ ref, np.eye(len(ref)), atol=1e-7)
    # interp_eval: the solution at the nodes and the reference's conventions outside the range
    np.testing.assert_allclose(s.interp_eval(g["ecm"]), s.bcs(), rtol=1e-12)
    assert s.interp_eval(float(g["threshold"])) == 0.0
    assert s.interp_eval(g["ecm"][-1] + 0.5) == s.bcs()[-1]


@pytest.mark.parametrize("name", ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c3_sle2_n24_eps64"])
def test_toy_batch_vs_golden(name):
    import isrsolver_b200 as isr
    g = load_golden(name)
    eff = None
    if "eff_values" in g:
        eff = isr.Efficiency("row_table", values=g["eff_values"], nx=int(g["eff_nx"]), xlo=float(g["eff_xlo"]), xhi=float(g["eff_xhi"]))
    s = make_solver(g, eff=eff)
    # use the oracle's matrix so that this test isolates the solve path
    s.set_matrix(g["G_converged"])
    r = s.toy_batch(g["toys"], g["bcs0"], want_sum=True, want_ratio=True, want_hist=True, want_bcs=True)
    np.testing.assert_allclose(r["chi2"], g["chi2"], rtol=RTOL)
    np.testing.assert_allclose(r["chi2"], g["chi2_loop"], rtol=RTOL)        # the reference's per-toy O(N^3) loop
    np.testing.assert_allclose(r["bcs"], g["B"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(r["sum_bcs"], g["B"].sum(0), rtol=RTOL)
    np.testing.assert_allclose(r["ratio_stats"], g["ratio_stats"], rtol=RTOL, atol=1e-9)
    np.testing.assert_array_equal(r["ratio_hist"], g["ratio_counts"].astype(np.uint32))   # bit-exact bins


def test_toy_batch_with_assembled_matrix():
    """End to end on the device: assemble -> factor -> toys; chi2 still within 1e-8 of the oracle."""
    g = load_golden("c2_sle_n40_cspline")
    s = make_solver(g)
    r = s.toy_batch(g["toys"], g["bcs0"])
    np.testing.assert_allclose(r["chi2"], g["chi2"], rtol=RTOL)


def test_chi2_histogram_bit_exact(oracle):
    from isrsolver_b200 import _lib as L
    from isrsolver_b200.pyisr import _ctx
    rng = np.random.default_rng(11)
    chi2 = rng.chisquare(100, 200001)
    lo, hi = C.c_double(), C.c_double(); cnt = np.zeros(130, dtype=np.uint32)
    L.check(L.load().isr_chi2_histogram(_ctx(), len(chi2), L.dptr(chi2), 128, C.byref(lo), C.byref(hi), L.uptr(cnt)))
    ref, idx, rlo, rhi = oracle.chi2_histogram(chi2)
    assert (lo.value, hi.value) == (rlo, rhi)
    np.testing.assert_array_equal(cnt, ref.astype(np.uint32))
    assert cnt[129] >= 1 and cnt[0] == 0          # the maximum lands in the overflow cell


@pytest.mark.parametrize("n,T", [(3, 1), (7, 5), (33, 129), (101, 1000), (64, 64), (199, 333)])
def test_ragged_shapes(oracle, n, T):
    """Odd N (padded leading dimension), T smaller than a tile, T not a multiple of the tile."""
    import isrsolver_b200 as isr
    rng = np.random.default_rng(n * 1000 + T)
    ecm = np.sort(1.0 + rng.random(n)) ; ecm += np.arange(n) * 1e-3
    G = np.tril(rng.random((n, n)) * 0.1) + np.eye(n) + (0.01 * rng.random((n, n)) if n % 2 else 0)
    vcs_err = 0.05 + 0.1 * rng.random(n)
    b0 = 1.0 + rng.random(n)
    V = (G @ b0)[None, :] + vcs_err[None, :] * rng.standard_normal((T, n))
    s = isr.ISRSolverSLE(n, ecm, G @ b0, None, vcs_err, 0.9)
    s.set_matrix(G)
    r = s.toy_batch(V, b0, want_sum=True, want_ratio=True, want_bcs=True)
    c2, B = oracle.chi2_test_fair(G, V, b0, vcs_err)
    np.testing.assert_allclose(r["chi2"], c2, rtol=RTOL)
    np.testing.assert_allclose(r["bcs"], B, rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(r["sum_bcs"], B.sum(0), rtol=RTOL)
    s.solve()
    bo, covo = oracle.sle_solve(G, G @ b0, vcs_err)
    np.testing.assert_allclose(s.bcs(), bo, rtol=RTOL)


@pytest.mark.parametrize("n", [5, 16, 23, 100, 200, 257, 500])
@pytest.mark.parametrize("shape", ["lower", "dense"])
def test_triangular_schedule_equals_dense_schedule_and_oracle(oracle, n, shape):
    """Lower-triangular operators (linear interpolation, no spread) skip their structural zeros -- bit-identical
    to the dense schedule, the skipped products are exact zeros; under schedule 2 a dense chi2 factor is replaced by
    its QL factor (|L d| = |R d|): chi2 within rounding of the dense schedule; all within 1e-8 of the oracle.  N around the
    K chunk (24 / 40), the column tile (128 / 200 / 256) and two column blocks (257, 500)."""
    import isrsolver_b200 as isr
    rng = np.random.default_rng(977 + n)
    T = 777
    ecm = np.sort(1.0 + rng.random(n)); ecm += np.arange(n) * 1e-3
    G = np.tril(rng.random((n, n)) * 0.1) + np.eye(n)
    if shape == "dense":
        G += 0.02 * rng.random((n, n))
    vcs_err = 0.05 + 0.1 * rng.random(n)
    b0 = 1.0 + rng.random(n)
    V = (G @ b0)[None, :] + vcs_err[None, :] * rng.standard_normal((T, n))
    c2, B = oracle.chi2_test_fair(G, V, b0, vcs_err)
    res = {}
    for variant in (0, 1, 2):
        s = isr.ISRSolverSLE(n, ecm, G @ b0, None, vcs_err, 0.9)
        s.set_matrix(G)
        s.set_toy_schedule(variant)
        r = s.toy_batch(V, b0, want_sum=True, want_ratio=True, want_bcs=True)
        np.testing.assert_allclose(r["chi2"], c2, rtol=RTOL)
        np.testing.assert_allclose(r["bcs"], B, rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(r["sum_bcs"], B.sum(0), rtol=RTOL)
        r2 = s.toy_batch(V, b0, want_sum=True, want_ratio=True)   # second batch on the already triangularised factor
        np.testing.assert_array_equal(r2["chi2"], r["chi2"])
        res[variant] = r
    for v in (0, 2):
        np.testing.assert_array_equal(res[v]["bcs"], res[1]["bcs"])
        np.testing.assert_array_equal(res[v]["ratio_stats"], res[1]["ratio_stats"])
    np.testing.assert_array_equal(res[0]["chi2"], res[1]["chi2"])
    if shape == "lower":
        np.testing.assert_array_equal(res[2]["chi2"], res[1]["chi2"])
    else:
        np.testing.assert_allclose(res[2]["chi2"], res[1]["chi2"], rtol=1e-12)


@pytest.mark.parametrize("n,T", [(40, 3), (200, 70), (333, 1000)])
def test_triangular_schedule_without_chi2_pass(oracle, n, T):
    """Pass 1 alone (no chi2 requested: the mean-solution loop of chi2TestData, src/Chi2Test.cpp:177-183) on
    lower-triangular operators, including a single tile per CTA and two column blocks."""
    import isrsolver_b200 as isr
    rng = np.random.default_rng(4400 + n)
    ecm = np.linspace(1.0, 2.0, n)
    G = np.tril(rng.random((n, n)) * 0.1) + np.eye(n)
    vcs_err = 0.05 + 0.1 * rng.random(n)
    b0 = 1.0 + rng.random(n)
    V = (G @ b0)[None, :] + vcs_err[None, :] * rng.standard_normal((T, n))
    _, B = oracle.chi2_test_fair(G, V, b0, vcs_err)
    out = {}
    for variant in (0, 1):
        s = isr.ISRSolverSLE(n, ecm, G @ b0, None, vcs_err, 0.9)
        s.set_matrix(G); s.set_toy_schedule(variant)
        out[variant] = s.toy_batch(V, b0, want_chi2=False, want_sum=True, want_ratio=True, want_bcs=True)
        np.testing.assert_allclose(out[variant]["bcs"], B, rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(out[variant]["sum_bcs"], B.sum(0), rtol=RTOL)
    for key in ("bcs", "sum_bcs", "ratio_stats"):
        np.testing.assert_array_equal(out[0][key], out[1][key])


def test_triangular_schedule_every_size():
    """Every N from 2 to 300 (all tile shapes, chunk counts and odd / even paddings on the way): the triangular
    schedule must return the bits of the dense one on lower-triangular operators, chi2 and solutions alike."""
    import isrsolver_b200 as isr
    rng = np.random.default_rng(808)
    T = 70
    for n in range(2, 301):
        ecm = np.linspace(1.0, 2.0, n)
        G = np.tril(rng.random((n, n)) * 0.1) + np.eye(n)
        err = 0.05 + 0.1 * rng.random(n)
        b0 = 1.0 + rng.random(n)
        V = (G @ b0)[None, :] + err[None, :] * rng.standard_normal((T, n))
        out = []
        for variant in (0, 1):
            s = isr.ISRSolverSLE(n, ecm, G @ b0, None, err, 0.9)
            s.set_matrix(G); s.set_toy_schedule(variant)
            out.append(s.toy_batch(V, b0, want_sum=True, want_bcs=True))
        assert np.array_equal(out[0]["chi2"], out[1]["chi2"]) and np.array_equal(out[0]["bcs"], out[1]["bcs"]), n
        assert np.array_equal(out[0]["sum_bcs"], out[1]["sum_bcs"]), n
        ref = np.linalg.solve(G, V.T).T
        assert np.allclose(out[0]["bcs"], ref, rtol=1e-10, atol=1e-12), n


def test_qr_chi2_factor_of_rank_deficient_and_ill_conditioned_operators(oracle):
    """TSVD with k < N selects a rank-deficient chi2 factor, a tiny lambda a badly conditioned one: the Householder
    QL replacement must reproduce the dense schedule there too."""
    import isrsolver_b200 as isr
    n, T = 48, 300
    rng = np.random.default_rng(5)
    ecm = np.linspace(1.2, 2.0, n)
    G = np.tril(rng.random((n, n)) * 0.3) + np.eye(n) + 0.05 * rng.random((n, n))
    vcs_err = 0.05 + 0.1 * rng.random(n)
    b0 = 1.0 + rng.random(n)
    V = (G @ b0)[None, :] + vcs_err[None, :] * rng.standard_normal((T, n))
    out = {}
    for variant in (2, 1):
        t = isr.ISRSolverTSVD(n, ecm, G @ b0, None, vcs_err, 0.9)
        t.set_matrix(G); t.setUpperTSVDIndex(n - 7); t.set_toy_schedule(variant)
        k = isr.ISRSolverTikhonov(n, ecm, G @ b0, None, vcs_err, 0.9)
        k.set_matrix(G); k.setLambda(1e-10); k.set_toy_schedule(variant)
        out[variant] = (t.toy_batch(V, b0)["chi2"], k.toy_batch(V, b0)["chi2"])
    np.testing.assert_allclose(out[2][0], out[1][0], rtol=1e-10)
    np.testing.assert_allclose(out[2][1], out[1][1], rtol=1e-10)


def test_last_chi2_range_and_histogram_from_the_device(oracle):
    """After a host-fed batch the chi2 values are also still on the device: their range and TH1F must equal what the
    reference's loop derives from the returned array (src/Chi2Test.cpp:64-78), bit for bit, without re-uploading it."""
    from isrsolver_b200 import _lib as L
    g = load_golden("c2_sle_n40_cspline")
    s = make_solver(g)
    r = s.toy_batch(g["toys"], g["bcs0"])
    lo, hi = C.c_double(), C.c_double()
    L.check(L.load().isr_last_chi2_minmax(s._p, C.byref(lo), C.byref(hi)))
    assert (lo.value, hi.value) == (r["chi2"].min(), r["chi2"].max())
    cnt = np.zeros(130, dtype=np.uint32)
    L.check(L.load().isr_last_chi2_histogram(s._p, 128, lo.value, hi.value, L.uptr(cnt)))
    ref, _, rlo, rhi = oracle.chi2_histogram(r["chi2"])
    assert (rlo, rhi) == (lo.value, hi.value)
    np.testing.assert_array_equal(cnt, ref.astype(np.uint32))


def test_pinned_toys_are_the_same_toys(oracle):
    """draw_toys(pinned=True) fills page-locked memory (isr_host_alloc) with exactly the values of the pageable
    stream; a batch fed from it returns the same bits; the memory is released with the array."""
    import gc
    import isrsolver_b200 as isr
    from isrsolver_b200.pyisr import pinned_empty
    g = load_golden("c2_sle_n40_cspline")
    s = make_solver(g)
    vcs, err = g["G_converged"] @ g["bcs0"], np.full(len(g["bcs0"]), 0.03)
    Vp = isr.draw_toys(vcs, err, 300, seed=5, pinned=True)
    Vn = isr.draw_toys(vcs, err, 300, seed=5)
    np.testing.assert_allclose(Vp, Vn, rtol=0, atol=1e-15)      # (a * z + b in two steps vs one expression)
    r1, r2 = s.toy_batch(Vp, g["bcs0"]), s.toy_batch(np.array(Vp), g["bcs0"])
    np.testing.assert_array_equal(r1["chi2"], r2["chi2"])
    a = pinned_empty((3, 4)); a[:] = 7.0
    assert a.shape == (3, 4) and a.sum() == 84.0
    del a, Vp; gc.collect()
    assert pinned_empty(0).size == 0
    res = s.chi2_test(500, vcs, g["bcs0"], err, seed=11)          # draws its toys into pinned memory
    ref = s.toy_batch(isr.draw_toys(vcs, err, 500, seed=11), g["bcs0"])["chi2"]
    np.testing.assert_allclose(res["chi2"], ref, rtol=1e-12)


def test_empty_batch():
    import isrsolver_b200 as isr
    g = load_golden("c1_sle_n50_linear")
    s = make_solver(g)
    s.set_matrix(g["G_converged"])
    r = s.toy_batch(np.empty((0, 50)), g["bcs0"], want_sum=True, want_ratio=True)
    assert r["chi2"].shape == (0,) and np.all(r["sum_bcs"] == 0) and np.all(r["ratio_stats"] == 0)


def test_call_order_errors():
    import isrsolver_b200 as isr
    from isrsolver_b200 import _lib as L
    g = load_golden("c1_sle_n50_linear")
    s = make_solver(g)
    chi2 = np.empty(4)
    rc = L.load().isr_toy_batch(s._p, 4, L.dptr(g["toys"][:4].copy()), L.dptr(g["bcs0"]), L.dptr(chi2), None, None, None, None)
    assert rc == L.ISR_ESTATE
    with pytest.raises(ValueError):
        s2 = make_solver(g); s2._vcs_err[3] = 0.0; s2.solve()
    sing = make_solver(g); sing.set_matrix(np.ones((50, 50)))
    with pytest.raises(isr.IsrGpuError):
        sing.solve()


def test_deterministic_repeat():
    """Same inputs -> bit-identical outputs (fixed-order reductions, no floating-point atomics)."""
    g = load_golden("c2_sle_n40_cspline")
    s = make_solver(g); s.set_matrix(g["G_converged"])
    V = np.tile(g["toys"], (40, 1))
    a = s.toy_batch(V, g["bcs0"], want_sum=True, want_ratio=True)
    b = s.toy_batch(V, g["bcs0"], want_sum=True, want_ratio=True)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


def test_full_size_toy_properties(oracle):
    """BASELINE size (N = 200, 2^18 toys here; the bench runs 2^20) through size-independent properties:
    chi2_k = |W^(1/2)(v_k - G b0)|^2 (survey A.7), linearity of the solve, chi2 ~ chi2(N)."""
    import isrsolver_b200 as isr
    n, T = 200, 1 << 18
    ecm = np.linspace(1.18, 2.0, n)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    s.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    s.eval_equation_matrix()
    G = s.intop_matrix().copy()
    b0 = oracle.bw_born(ecm); vcs = G @ b0; err = 0.05 * vcs
    s.resetVisibleCS(vcs); s.resetVisibleCSErrors(err)
    V = isr.draw_toys(vcs, err, T)
    r = s.toy_batch(V, b0, want_sum=True)
    z = (V - vcs[None, :]) / err[None, :]
    np.testing.assert_allclose(r["chi2"], (z * z).sum(1), rtol=RTOL)
    assert r["chi2"].mean() == pytest.approx(n, rel=5e-3)
    # linearity: sum_k b_k = G^-1 sum_k v_k
    np.testing.assert_allclose(r["sum_bcs"], np.linalg.solve(G, V.sum(0)), rtol=RTOL)
    # first toys against the oracle loop
    c2, B = oracle.chi2_test_fair(G, V[:512], b0, err)
    np.testing.assert_allclose(r["chi2"][:512], c2, rtol=RTOL)


def test_chi2_test_model_and_ratio_api(tmp_path, oracle):
    """PyISR surface: chi2_test_model / ratio_test_model with .npz model files."""
    g = load_golden("c2_sle_n40_cspline")
    s = make_solver(g)
    model = tmp_path / "model.npz"
    np.savez(model, vcsBlured=np.stack([g["ecm"], g["vcs"], 0 * g["vcs"], g["vcs_err"]]),
             bcsSRC=np.stack([g["ecm"], g["bcs0"], 0 * g["ecm"], 0 * g["ecm"]]))
    out = tmp_path / "chi2.npz"
    res = s.chi2_test_model(n=len(g["toys"]), initial_ampl=1e4, model_path=str(model), model_visible_cs_name="vcsBlured",
                            model_born_cs_name="bcsSRC", output_path=str(out))
    np.testing.assert_allclose(res["chi2"], g["chi2"], rtol=1e-7)     # assembled (not injected) matrix
    h, idx, lo, hi = oracle.chi2_histogram(res["chi2"])
    np.testing.assert_array_equal(res["hist"], h)
    assert out.exists()
    rr = s.ratio_test_model(n=len(g["toys"]), model_path=str(model), model_visible_cs_name="vcsBlured", model_born_cs_name="bcsSRC")
    np.testing.assert_allclose(rr["means"], g["ratio_means"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(rr["mean_errors"], g["ratio_mean_errors"], rtol=1e-6, atol=1e-9)
    import isrsolver_b200 as isr
    np.savez(model, vcsBlured=g["vcs"][:-1], bcsSRC=g["bcs0"])
    with pytest.raises(isr.DifferentModelCrossSectionSizes):
        s.chi2_test_model(8, 1e4, str(model), "vcsBlured", "bcsSRC")


# ---- toy campaigns by count: toys drawn on the device (csrc/rng.cu) ---------------------------------------
def test_device_toys_are_standard_normal():
    import isrsolver_b200 as isr
    from scipy import stats
    n, T = 40, 1 << 15
    z = isr.draw_toys_device(np.zeros(n), np.ones(n), T, seed=12345)
    assert z.shape == (T, n) and np.all(np.isfinite(z))
    m = z.size
    assert abs(z.mean()) < 5 / np.sqrt(m) and abs(z.var() - 1) < 5 * np.sqrt(2 / m)
    assert abs(stats.skew(z.ravel())) < 5 * np.sqrt(6 / m) and abs(stats.kurtosis(z.ravel())) < 5 * np.sqrt(24 / m)
    assert stats.kstest(z.ravel()[:200000], "norm").pvalue > 1e-4
    # columns and toys are uncorrelated; even / odd columns (the cos / sin branches of a pair) both normal
    c = np.corrcoef(z[:, :8].T)
    assert np.abs(c - np.eye(8)).max() < 5 / np.sqrt(T)
    assert stats.kstest(z[:, 0], "norm").pvalue > 1e-4 and stats.kstest(z[:, 1], "norm").pvalue > 1e-4
    assert np.abs(np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]) < 5 / np.sqrt(T)
    # a different seed is a different stream; scale and shift are applied per column
    assert not np.array_equal(z, isr.draw_toys_device(np.zeros(n), np.ones(n), T, seed=12346))
    v = isr.draw_toys_device(np.arange(n, dtype=float), 0.5 * np.ones(n), 16, seed=12345)
    np.testing.assert_allclose(v, np.arange(n)[None, :] + 0.5 * z[:16], rtol=0, atol=1e-15 * n)


def test_device_toys_over_many_sizes():
    """The generator's thread mapping (column quads, several toys per block, quads walked in strides for N > 1024) at sizes
    around every boundary: finite, standard normal, independent columns, and identical however the campaign is chunked."""
    import isrsolver_b200 as isr
    for n in (1, 2, 3, 5, 7, 8, 63, 64, 65, 127, 128, 129, 255, 256, 257, 258, 500, 513, 1000, 1023, 1024, 1025, 1100, 2050):
        vcs, err = np.linspace(1, 2, n), np.linspace(0.1, 0.3, n)
        T = 4096 if n < 300 else 512
        full = isr.draw_toys_device(vcs, err, T, seed=99)
        parts = np.concatenate([isr.draw_toys_device(vcs, err, c, seed=99, first_toy=f) for f, c in ((0, 17), (17, 1), (18, T - 18))])
        np.testing.assert_array_equal(full, parts, err_msg=str(n))
        z = (full - vcs) / err
        assert np.all(np.isfinite(full)) and np.abs(z).max() < 6.8, n
        assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 6 * np.sqrt(2 / z.size), (n, z.mean(), z.var())
        assert len(np.unique(z[:, 0])) > 0.99 * T, n
        if n > 1:
            assert abs(np.corrcoef(z[:, 0], z[:, -1])[0, 1]) < 6 / np.sqrt(T), n
            assert abs(np.corrcoef(z[:, 0], z[:, n // 2])[0, 1]) < 6 / np.sqrt(T) or n // 2 == 0, n


@pytest.mark.parametrize("n", [36, 37])
def test_device_toys_depend_only_on_seed_and_index(n):
    """Sharding by first_toy reproduces the unsharded stream bit for bit (odd N included)."""
    import isrsolver_b200 as isr
    vcs, err = np.linspace(1, 2, n), np.linspace(0.1, 0.2, n)
    full = isr.draw_toys_device(vcs, err, 1000, seed=7)
    parts = [isr.draw_toys_device(vcs, err, c, seed=7, first_toy=f) for f, c in ((0, 400), (400, 1), (401, 599))]
    np.testing.assert_array_equal(full, np.concatenate(parts))
    big = isr.draw_toys_device(vcs, err, 3, seed=7, first_toy=(1 << 33) + 5)      # 64-bit toy indices
    assert np.all(np.isfinite(big)) and not np.array_equal(big, full[:3])


def test_toy_campaign_equals_batch_on_the_same_toys(oracle):
    """isr_toy_campaign (device toys) == isr_toy_batch on isr_draw_toys' vectors == the oracle on them."""
    import isrsolver_b200 as isr
    g = load_golden("c2_sle_n40_cspline")
    s = make_solver(g)
    s.set_matrix(g["G_converged"])           # the oracle's matrix: this test isolates the toy path
    T, seed = 70000, 99                      # more than one 2^16 chunk
    r = s.toy_campaign(T, g["vcs"], g["vcs_err"], g["bcs0"], seed, want_sum=True, want_ratio=True, want_hist=True)
    V = isr.draw_toys_device(g["vcs"], g["vcs_err"], T, seed)
    rb = s.toy_batch(V, g["bcs0"], want_sum=True, want_ratio=True, want_hist=True)
    np.testing.assert_array_equal(r["chi2"], rb["chi2"])
    np.testing.assert_allclose(r["sum_bcs"], rb["sum_bcs"], rtol=1e-12)
    np.testing.assert_array_equal(r["ratio_hist"], rb["ratio_hist"])
    np.testing.assert_allclose(r["ratio_stats"], rb["ratio_stats"], rtol=1e-10, atol=1e-9)
    c2, B = oracle.chi2_test_fair(g["G_converged"], V[:4096], g["bcs0"], g["vcs_err"])
    np.testing.assert_allclose(r["chi2"][:4096], c2, rtol=1e-8)
    hist, idx, lo, hi = oracle.chi2_histogram(r["chi2"])
    assert (r["lo"], r["hi"]) == (lo, hi)
    np.testing.assert_array_equal(r["hist"], hist)
    assert r["chi2"].mean() == pytest.approx(len(g["ecm"]), rel=0.02)
    # sharded campaign == unsharded campaign
    a = s.toy_campaign(30000, g["vcs"], g["vcs_err"], g["bcs0"], seed, want_chi2_hist=False)["chi2"]
    b = s.toy_campaign(40000, g["vcs"], g["vcs_err"], g["bcs0"], seed, first_toy=30000, want_chi2_hist=False)["chi2"]
    np.testing.assert_array_equal(np.concatenate([a, b]), r["chi2"])
    # the reference-style entry points with rng="device"
    d = s.chi2_test_data(4096, rng="device")
    assert d["chi2"].shape == (4096,) and d["hist"].sum() == 4096


_SHARDED_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
os.environ["LOCAL_RANK"] = str(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())   # ranks may share a GPU here
import isrsolver_b200 as isr
dist.init_process_group("gloo")
g = np.load(sys.argv[2])
n = len(g["ecm"])
s = isr.ISRSolverSLE(n, g["ecm"], g["vcs"], None, g["vcs_err"], 0.827)
s.set_interp_settings([tuple(r) for r in g["settings"]])
T = 100003
r = s.chi2_test_sharded(T, g["vcs"], g["bcs0"], g["vcs_err"], seed=5)
one = s.toy_campaign(T, g["vcs"], g["vcs_err"], g["bcs0"], seed=5, want_sum=True, want_ratio=True)
assert (r["lo"], r["hi"]) == (one["lo"], one["hi"]), (r["lo"], r["hi"], one["lo"], one["hi"])
assert np.array_equal(r["hist"], one["hist"]) and r["hist"].sum() == T
assert np.allclose(r["sum_bcs"], one["sum_bcs"], rtol=1e-11) and np.allclose(r["ratio_stats"], one["ratio_stats"], rtol=1e-9, atol=1e-9)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_chi2_test_equals_single_gpu(tmp_path, world):
    """(e) multi-GPU on real kernels: `world` ranks (gloo for the two small collectives; the ranks share the
    GPUs that exist) shard a device-drawn campaign; histogram bit-identical to the unsharded campaign."""
    import os
    import subprocess
    import sys
    from conftest import GOLDEN, ROOT
    w = tmp_path / "worker.py"
    w.write_text(_SHARDED_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29640 + world), str(w), ROOT,
                          os.path.join(GOLDEN, "c2_sle_n40_cspline.npz")],
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world


def test_device_resident_minmax_and_histogram(oracle):
    """isr_minmax_to_dev / isr_histogram_range_dev (the multi-GPU pipeline's forms) == the host-range forms."""
    import torch
    from isrsolver_b200 import _lib as L
    from isrsolver_b200.pyisr import _ctx
    lib, ctx = L.load(), _ctx()
    v = torch.from_numpy(np.random.default_rng(4).chisquare(40, 100003)).cuda()
    rng, cnt = torch.empty(2, dtype=torch.float64, device="cuda"), torch.full((130,), 7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    L.check(lib.isr_minmax_to_dev(ctx, v.numel(), v.data_ptr(), rng.data_ptr()))
    L.check(lib.isr_histogram_range_dev(ctx, v.numel(), v.data_ptr(), 128, rng.data_ptr(), cnt.data_ptr()))
    L.check(lib.isr_ctx_synchronize(ctx))
    lo, hi = float(rng[0]), -float(rng[1])
    assert (lo, hi) == (float(v.min()), float(v.max()))
    ref, *_ = oracle.chi2_histogram(v.cpu().numpy())
    np.testing.assert_array_equal(cnt.cpu().numpy(), ref.astype(np.int32))


def test_iterative_solver_matches_numpy():
    """IterISRInterpSolver::solve (src/IterISRInterpSolver.cpp:45-57) on the device against the same loop in NumPy,
    through the ctypes mirror and the compiled PyISR module."""
    import isrsolver_b200 as isr
    from isrsolver_b200 import PyISR
    g = load_golden("c2_sle_n40_cspline")
    n = len(g["ecm"])
    st = [tuple(int(v) for v in r) for r in g["settings"]]
    for n_iter in (0, 1, 10):
        s = isr.IterISRInterpSolver(n, g["ecm"], g["vcs"], None, g["vcs_err"], float(g["threshold"]), n_iter=n_iter)
        s.set_interp_settings(st)
        s.solve()
        G = s.intop_matrix()
        b, rad = g["vcs"].copy(), np.zeros(n)
        for _ in range(n_iter):
            rad = (G @ b) / b
            b = g["vcs"] / rad
        np.testing.assert_allclose(s.bcs(), b, rtol=1e-13)
        if n_iter:
            np.testing.assert_allclose(s.radcorr(), rad, rtol=1e-13)
            np.testing.assert_allclose(np.diag(s.bcs_cov_matrix()), (g["vcs_err"] / rad) ** 2, rtol=1e-13)
    # after 10 iterations the radiative-correction method agrees with the direct solution at the percent level
    np.testing.assert_allclose(s.bcs(), g["bcs"], rtol=0.05)
    c = PyISR.IterISRInterpSolver(n, g["ecm"], g["vcs"], None, g["vcs_err"], float(g["threshold"]))
    c.set_interp_settings([(bool(a), b_, c_) for a, b_, c_ in st])
    assert c.n_iter == 10
    c.solve()
    np.testing.assert_allclose(c.bcs(), s.bcs(), rtol=1e-15)
    np.testing.assert_allclose(c.radcorr(), s.radcorr(), rtol=1e-15)
