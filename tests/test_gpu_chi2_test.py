"""isr_chi2_test (the whole chi2 / ratio test in one stream-ordered call with a deferred certificate), the blocked
triangular inverse behind it for N > 240, the communicators and the single-process multi-GPU group -- through the C ABI.

Tolerances (north star): per-toy chi2, sums and covariance 1e-8 relative; histogram cells bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _lib():
    from isrsolver_b200 import _lib as L
    return L, L.load()


def _ctx():
    from isrsolver_b200.pyisr import _ctx
    return _ctx()


def _problem(L, lib, ctx, ecm, thr, settings, G=None):
    st = np.ascontiguousarray(settings, dtype=np.int32)
    p = C.c_void_p()
    L.check(lib.isr_problem_create(ctx, len(ecm), L.dptr(np.ascontiguousarray(ecm)), float(thr), len(st), L.iptr(st), None, C.byref(p)))
    if G is not None:
        L.check(lib.isr_set_matrix(p, L.dptr(np.ascontiguousarray(G.T))))     # column-major
    return p


def _run(L, lib, p, n, *, flags, err, bcs0, T, V=None, V_dev=None, vcs=None, seed=0, first=0, nbins=128, comm=None,
         want_chi2=True, ratio_hist=False, call=None):
    a = L.Chi2TestArgs()
    keep = []

    def arr(x, dt=np.float64):
        x = np.ascontiguousarray(x, dtype=dt); keep.append(x); return x
    a.flags = flags
    a.vcs_err = L.dptr(arr(err)); a.bcs0 = L.dptr(arr(bcs0)); a.n_toys = T
    if V is not None:
        a.V = L.dptr(arr(V))
    if V_dev is not None:
        a.V_dev = V_dev
    if vcs is not None:
        a.vcs = L.dptr(arr(vcs))
    a.seed, a.first_toy, a.nbins, a.comm = seed, first, nbins, comm
    out = {"chi2": np.zeros(max(T, 1)), "lohi": np.zeros(2), "counts": np.zeros(nbins + 2, dtype=np.int64), "sum_bcs": np.zeros(n),
           "ratio_stats": np.zeros(3 * n), "ratio_hist": np.zeros((n, 1026), dtype=np.uint32)}
    total, refac = C.c_int64(), C.c_int()
    if want_chi2:
        a.chi2_out = L.dptr(out["chi2"])
    a.chi2_lo_hi_out = L.dptr(out["lohi"]); a.chi2_counts_out = out["counts"].ctypes.data_as(C.POINTER(C.c_int64))
    a.sum_bcs_out = L.dptr(out["sum_bcs"])
    if flags & L.ISR_TEST_RATIO:
        a.ratio_stats_out = L.dptr(out["ratio_stats"])
        if ratio_hist:
            a.ratio_hist_out = L.uptr(out["ratio_hist"])
    a.total_toys_out = C.pointer(total); a.refactored_out = C.pointer(refac)
    L.check((call or lib.isr_chi2_test)(p, C.byref(a)))
    out["chi2"] = out["chi2"][:T]
    out["total"], out["refactored"] = total.value, refac.value
    return out


@pytest.mark.parametrize("name", ["c1_sle_n50_linear", "c2_sle_n40_cspline"])
def test_fused_test_equals_golden_and_separate_calls(name, oracle):
    """assemble + factor + toys + range + TH1F in one call == the oracle's chi2 / histogram and the call-by-call path."""
    L, lib = _lib()
    g = load_golden(name)
    n = len(g["ecm"])
    p = _problem(L, lib, _ctx(), g["ecm"], float(g["threshold"]), g["settings"])
    fl = L.ISR_TEST_ASSEMBLE | L.ISR_TEST_FACTOR | L.ISR_TEST_RATIO
    r = _run(L, lib, p, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=len(g["toys"]), V=g["toys"], ratio_hist=True)
    np.testing.assert_allclose(r["chi2"], g["chi2"], rtol=RTOL)
    ref, _, lo, hi = oracle.chi2_histogram(r["chi2"])
    assert (r["lohi"][0], r["lohi"][1]) == (lo, hi)
    np.testing.assert_array_equal(r["counts"], ref.astype(np.int64))
    np.testing.assert_allclose(r["sum_bcs"], g["B"].sum(0), rtol=RTOL)
    np.testing.assert_allclose(r["ratio_stats"].reshape(n, 3), g["ratio_stats"], rtol=RTOL, atol=1e-9)
    np.testing.assert_array_equal(r["ratio_hist"], g["ratio_counts"].astype(np.uint32))
    assert r["total"] == len(g["toys"]) and r["refactored"] == 0
    # call by call on the same problem: bit-identical chi2
    chi2 = np.zeros(len(g["toys"]))
    L.check(lib.isr_assemble(p, None, None)); L.check(lib.isr_factor(p, L.dptr(np.ascontiguousarray(g["vcs_err"]))))
    L.check(lib.isr_toy_batch(p, len(chi2), L.dptr(np.ascontiguousarray(g["toys"])), L.dptr(np.ascontiguousarray(g["bcs0"])), L.dptr(chi2), None, None, None, None))
    np.testing.assert_array_equal(chi2, r["chi2"])
    lib.isr_problem_destroy(p)


def test_fused_test_device_toys_and_drawn_toys(oracle):
    """V_dev and device-drawn toys take the same kernels: the drawn campaign equals isr_toy_campaign bit for bit, and the
    oracle on the vectors isr_draw_toys returns."""
    import torch
    L, lib = _lib()
    g = load_golden("c2_sle_n40_cspline")
    n, T = len(g["ecm"]), 70000            # more than one 2^16 block
    ctx = _ctx()
    p = _problem(L, lib, ctx, g["ecm"], float(g["threshold"]), g["settings"])
    fl = L.ISR_TEST_ASSEMBLE | L.ISR_TEST_FACTOR
    r = _run(L, lib, p, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=T, vcs=g["vcs"], seed=77, first=1000)
    V = np.zeros((T, n))
    L.check(lib.isr_draw_toys(ctx, T, n, 1000, 77, L.dptr(np.ascontiguousarray(g["vcs"])), L.dptr(np.ascontiguousarray(g["vcs_err"])), L.dptr(V)))
    c2, B = oracle.chi2_test_fair(g["G_converged"], V, g["bcs0"], g["vcs_err"])
    np.testing.assert_allclose(r["chi2"], c2, rtol=RTOL)
    np.testing.assert_allclose(r["sum_bcs"], B.sum(0), rtol=RTOL)
    ref, _, lo, hi = oracle.chi2_histogram(r["chi2"])
    np.testing.assert_array_equal(r["counts"], ref.astype(np.int64))
    # the campaign entry point of round 1 on the same stream
    chi2 = np.zeros(T); cnt = np.zeros(130, dtype=np.uint32); lohi = np.zeros(2)
    L.check(lib.isr_toy_campaign(p, T, 1000, 77, L.dptr(np.ascontiguousarray(g["vcs"])), L.dptr(np.ascontiguousarray(g["vcs_err"])),
                                 L.dptr(np.ascontiguousarray(g["bcs0"])), L.dptr(chi2), None, None, None, 128, L.dptr(lohi), L.uptr(cnt)))
    np.testing.assert_array_equal(chi2, r["chi2"])
    np.testing.assert_array_equal(cnt.astype(np.int64), r["counts"])
    # V_dev: the same vectors resident on the device, no assembly / factorisation in the call
    Vd = torch.from_numpy(V).cuda()
    r2 = _run(L, lib, p, n, flags=0, err=g["vcs_err"], bcs0=g["bcs0"], T=T, V_dev=Vd.data_ptr())
    np.testing.assert_array_equal(r2["chi2"], r["chi2"])
    np.testing.assert_array_equal(r2["counts"], r["counts"])
    # an empty shard is legal (fewer toys than ranks)
    r3 = _run(L, lib, p, n, flags=0, err=g["vcs_err"], bcs0=g["bcs0"], T=0, vcs=g["vcs"])
    assert r3["total"] == 0 and r3["counts"].sum() == 0 and r3["lohi"][0] == np.inf
    lib.isr_problem_destroy(p)


def test_deferred_certificate_falls_back_to_pivoting(oracle):
    """A matrix the unpivoted Gauss-Jordan cannot handle (zero leading pivot): the speculative toys are discarded, the
    factorisation is redone with partial pivoting inside the same call and the results match the oracle."""
    L, lib = _lib()
    rng = np.random.default_rng(5)
    n, T = 24, 300
    G = rng.standard_normal((n, n)) + 3 * np.eye(n)
    G[0, 0] = 0.0
    ecm = np.linspace(1.2, 2.0, n)
    p = _problem(L, lib, _ctx(), ecm, 0.827, [[0, 0, n - 1]], G=G)
    b0 = rng.uniform(1, 2, n); vcs = G @ b0; err = 0.05 * np.abs(vcs) + 0.01
    V = vcs + err * rng.standard_normal((T, n))
    r = _run(L, lib, p, n, flags=L.ISR_TEST_FACTOR, err=err, bcs0=b0, T=T, V=V)
    assert r["refactored"] == 1
    c2, B = oracle.chi2_test_fair(G, V, b0, err)
    np.testing.assert_allclose(r["chi2"], c2, rtol=1e-7)
    np.testing.assert_allclose(r["sum_bcs"], B.sum(0), rtol=1e-7)
    # a singular matrix is still an error
    G[:, 3] = 0.0
    L.check(lib.isr_set_matrix(p, L.dptr(np.ascontiguousarray(G.T))))
    with pytest.raises(L.IsrGpuError):
        _run(L, lib, p, n, flags=L.ISR_TEST_FACTOR, err=err, bcs0=b0, T=T, V=V)
    lib.isr_problem_destroy(p)


@pytest.mark.parametrize("n", [241, 256, 257, 300, 383, 449, 500, 512, 513, 640])
def test_blocked_triangular_inverse(n):
    """N > 240, lower-triangular matrix: the hand-written blocked inverse (tri_inverse.cu) behind isr_factor."""
    L, lib = _lib()
    rng = np.random.default_rng(n)
    G = np.tril(rng.uniform(-0.3, 0.3, (n, n))) + np.diag(rng.uniform(0.7, 1.3, n))
    ecm = np.linspace(1.2, 2.0, n)
    p = _problem(L, lib, _ctx(), ecm, 0.827, [[0, 0, n - 1]], G=G)
    err = rng.uniform(0.5, 1.5, n)
    L.check(lib.isr_factor(p, L.dptr(err)))
    X = np.zeros((n, n))
    L.check(lib.isr_get_inverse(p, L.dptr(X)))
    X = X.T
    ref = np.linalg.inv(G)
    assert np.abs(X - ref).max() <= 1e-10 * np.abs(ref).max()
    assert np.count_nonzero(np.triu(X, 1)) == 0                 # exact zeros above the diagonal
    lib.isr_problem_destroy(p)


@pytest.mark.parametrize("n", [241, 242, 255, 256, 257, 300, 333, 400, 479, 480, 481, 500, 600, 721, 960, 961, 1000])
def test_dense_inverse_beyond_240(n):
    """240 < N <= 960, dense matrix: block elimination on the cluster Gauss-Jordan kernel (two inverses of the parts + six
    products per level, linalg.cu::dense_inverse; two levels above 480); beyond 960 cuSOLVER's unpivoted LU.  Both behind
    the certificate."""
    L, lib = _lib()
    rng = np.random.default_rng(1000 + n)
    G = rng.uniform(-0.2, 0.2, (n, n)) / np.sqrt(n) * 4 + np.diag(rng.uniform(0.8, 1.2, n))     # cond ~ 10
    ecm = np.linspace(1.2, 2.0, n)
    p = _problem(L, lib, _ctx(), ecm, 0.827, [[0, 0, 0], [1, 1, n - 1]], G=G)
    err = rng.uniform(0.5, 1.5, n)
    L.check(lib.isr_factor(p, L.dptr(err)))
    X = np.zeros((n, n))
    L.check(lib.isr_get_inverse(p, L.dptr(X)))
    ref = np.linalg.inv(G)
    assert np.abs(X.T - ref).max() <= 1e-10 * np.abs(ref).max()
    vcs = G @ np.linspace(1.0, 2.0, n)
    b, cov = np.zeros(n), np.zeros((n, n))
    L.check(lib.isr_solve(p, L.dptr(vcs), L.dptr(b), L.dptr(cov)))
    np.testing.assert_allclose(b, np.linspace(1.0, 2.0, n), rtol=1e-10)
    lib.isr_problem_destroy(p)


def test_dense_inverse_beyond_240_falls_back_to_pivoting():
    """A matrix whose leading half is singular: the unpivoted block elimination cannot invert it, the certificate says so and
    the partially pivoted LU takes over."""
    L, lib = _lib()
    n = 300
    rng = np.random.default_rng(7)
    G = rng.uniform(-1, 1, (n, n))
    G[:150, :150] = 0.0                              # A11 = 0: no block elimination without pivoting
    G += 0.0
    ecm = np.linspace(1.2, 2.0, n)
    p = _problem(L, lib, _ctx(), ecm, 0.827, [[0, 0, 0], [1, 1, n - 1]], G=G)
    L.check(lib.isr_factor(p, L.dptr(np.ones(n))))
    X = np.zeros((n, n))
    L.check(lib.isr_get_inverse(p, L.dptr(X)))
    ref = np.linalg.inv(G)
    assert np.abs(X.T - ref).max() <= 1e-8 * np.abs(ref).max()
    lib.isr_problem_destroy(p)


def test_batch_dev_solve_batch_dev_sequence(oracle):
    """isr_solve / isr_tikhonov_solve / isr_tsvd_solve used to write their right-hand side into the buffer
    isr_toy_batch_dev caches the reference solution in; the second batch then computed d = b - vcs."""
    import torch
    L, lib = _lib()
    g = load_golden("c2_sle_n40_cspline")
    n, T = len(g["ecm"]), 64
    p = _problem(L, lib, _ctx(), g["ecm"], float(g["threshold"]), g["settings"], G=g["G_converged"])
    err = np.ascontiguousarray(g["vcs_err"]); b0 = np.ascontiguousarray(g["bcs0"]); vcs = np.ascontiguousarray(g["vcs"])
    L.check(lib.isr_factor(p, L.dptr(err)))
    Vd = torch.from_numpy(np.ascontiguousarray(g["toys"][:T])).cuda()
    c1, c2 = torch.zeros(T, dtype=torch.float64, device="cuda"), torch.zeros(T, dtype=torch.float64, device="cuda")
    L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(b0), c1.data_ptr(), None, None, None, None))
    bcs = np.zeros(n)
    L.check(lib.isr_solve(p, L.dptr(vcs), L.dptr(bcs), None))
    L.check(lib.isr_tsvd_solve(p, n, 0, L.dptr(vcs), L.dptr(err), L.dptr(bcs), None))
    L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(b0), c2.data_ptr(), None, None, None, None))
    L.check(lib.isr_ctx_synchronize(_ctx()))
    np.testing.assert_array_equal(c1.cpu().numpy(), c2.cpu().numpy())
    np.testing.assert_allclose(c1.cpu().numpy(), g["chi2"][:T], rtol=RTOL)
    lib.isr_problem_destroy(p)


def test_factor_invalidates_tikhonov_and_tsvd_cov_drops_selection():
    """isr_factor with new errors must invalidate the eigen-system prepared for the old ones, and
    isr_tsvd_solve(cov_out) must not silently replace a Tikhonov selection by the SLE operators."""
    L, lib = _lib()
    g = load_golden("c4_tikhonov_n32_spread")
    n = len(g["ecm"])
    p = _problem(L, lib, _ctx(), g["ecm"], float(g["threshold"]), g["settings"], G=g["G_spread"])
    err = np.ascontiguousarray(g["vcs_err"]); vcs = np.ascontiguousarray(g["vcs"])
    L.check(lib.isr_tikhonov_prepare(p, L.dptr(err), 1))
    lam = np.array([1e-3]); o = [np.zeros(1) for _ in range(4)]
    L.check(lib.isr_tikhonov_scan(p, 1, L.dptr(lam), L.dptr(vcs), *[L.dptr(x) for x in o], None))
    L.check(lib.isr_factor(p, L.dptr(err)))          # same weights: eigen-system stays valid
    L.check(lib.isr_tikhonov_scan(p, 1, L.dptr(lam), L.dptr(vcs), *[L.dptr(x) for x in o], None))
    L.check(lib.isr_factor(p, L.dptr(np.ascontiguousarray(2.0 * err))))
    with pytest.raises(L.IsrGpuError):
        L.check(lib.isr_tikhonov_scan(p, 1, L.dptr(lam), L.dptr(vcs), *[L.dptr(x) for x in o], None))
    L.check(lib.isr_tikhonov_prepare(p, L.dptr(err), 1))
    L.check(lib.isr_tikhonov_select(p, 1e-3))
    bcs, cov = np.zeros(n), np.zeros((n, n))
    L.check(lib.isr_tsvd_solve(p, n, 0, L.dptr(vcs), L.dptr(err), L.dptr(bcs), L.dptr(cov)))
    V = np.ascontiguousarray(np.tile(vcs, (4, 1))); chi2 = np.zeros(4)
    with pytest.raises(L.IsrGpuError):               # the selection was dropped, loudly
        L.check(lib.isr_toy_batch(p, 4, L.dptr(V), L.dptr(vcs), L.dptr(chi2), None, None, None, None))
    lib.isr_problem_destroy(p)


def _group(L, lib, ngpu):
    g = C.c_void_p()
    L.check(lib.isr_group_create(ngpu, None, C.byref(g)))
    return g


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ngpu", [1, 2, 3])
def test_group_chi2_test_equals_single_gpu(ngpu, oracle):
    """One process, ngpu devices (host thread + context + replicated problem each, NCCL all-reduces in stream): the
    histogram cells and the range are bit-identical to the single-GPU test, chi2 comes back in toy order."""
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    L, lib = _lib()
    g = load_golden("c2_sle_n40_cspline")
    n, T = len(g["ecm"]), 50001
    st = np.ascontiguousarray(g["settings"], dtype=np.int32)
    ecm = np.ascontiguousarray(g["ecm"])
    grp = _group(L, lib, ngpu)
    assert lib.isr_group_size(grp) == ngpu
    gp = C.c_void_p()
    L.check(lib.isr_group_problem_create(grp, n, L.dptr(ecm), float(g["threshold"]), len(st), L.iptr(st), None, C.byref(gp)))
    fl = L.ISR_TEST_ASSEMBLE | L.ISR_TEST_FACTOR | L.ISR_TEST_RATIO
    # toys in a NUMA-placed page-locked buffer of the group
    Vp = C.POINTER(C.c_double)()
    L.check(lib.isr_group_host_alloc(grp, T, n, C.byref(Vp)))
    V = np.ctypeslib.as_array(Vp, shape=(T, n))
    V[:] = oracle.draw_toys(g["vcs"], g["vcs_err"], T)
    rg = _run(L, lib, gp, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=T, V=V, ratio_hist=True, call=lib.isr_group_chi2_test)
    p1 = _problem(L, lib, _ctx(), ecm, float(g["threshold"]), st)
    r1 = _run(L, lib, p1, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=T, V=V, ratio_hist=True)
    np.testing.assert_array_equal(rg["chi2"], r1["chi2"])
    np.testing.assert_array_equal(rg["counts"], r1["counts"])
    np.testing.assert_array_equal(rg["lohi"], r1["lohi"])
    np.testing.assert_array_equal(rg["ratio_hist"], r1["ratio_hist"])
    np.testing.assert_allclose(rg["sum_bcs"], r1["sum_bcs"], rtol=1e-12)      # float sums are re-associated across ranks
    np.testing.assert_allclose(rg["ratio_stats"], r1["ratio_stats"], rtol=1e-12, atol=1e-12)
    assert rg["total"] == T
    # device-drawn campaign sharded by first_toy: same toys whatever the number of GPUs
    rd = _run(L, lib, gp, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=T, vcs=g["vcs"], seed=5, call=lib.isr_group_chi2_test)
    r1d = _run(L, lib, p1, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=T, vcs=g["vcs"], seed=5)
    np.testing.assert_array_equal(rd["chi2"], r1d["chi2"])
    np.testing.assert_array_equal(rd["counts"], r1d["counts"])
    # fewer toys than GPUs: empty shards
    re = _run(L, lib, gp, n, flags=fl, err=g["vcs_err"], bcs0=g["bcs0"], T=1, V=V[:1], call=lib.isr_group_chi2_test)
    assert re["total"] == 1 and re["counts"].sum() == 1
    lib.isr_group_host_free(grp, Vp)
    lib.isr_problem_destroy(p1)
    lib.isr_group_problem_destroy(gp)
    lib.isr_group_destroy(grp)


def test_comm_of_one_and_allreduce():
    import torch
    L, lib = _lib()
    ctx = _ctx()
    comm = C.c_void_p()
    L.check(lib.isr_comm_init_rank(ctx, 1, 0, None, C.byref(comm)))
    assert lib.isr_comm_size(comm) == 1 and lib.isr_comm_rank(comm) == 0
    x = torch.arange(5, dtype=torch.float64, device="cuda")
    L.check(lib.isr_comm_allreduce_dev(comm, x.data_ptr(), 5, 0, 0))
    L.check(lib.isr_ctx_synchronize(ctx))
    assert x.cpu().tolist() == [0, 1, 2, 3, 4]
    node = C.c_int(-2)
    L.check(lib.isr_ctx_bind_host_thread(ctx, C.byref(node)))
    assert node.value >= -1
    lib.isr_comm_destroy(comm)


@pytest.mark.parametrize("ngpu", [1, 2])
def test_tikhonov_chi2_scan_toys_sharded(ngpu):
    """BASELINE config C4 sharded on the toy axis: per-lambda {min, max} and TH1F cells from toys spread
    over `ngpu` contexts of one process (one host thread each, isr_comm_init_all) are bit-identical to the single-GPU
    isr_tikhonov_toy_scan_hist; the mean agrees to rounding (the sums are re-associated)."""
    import threading
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    L, lib = _lib()
    g = load_golden("c4_tikhonov_n32_spread")
    n, T, Lm = len(g["ecm"]), 5000, 37
    rng = np.random.default_rng(3)
    err = np.ascontiguousarray(g["vcs_err"]); b0 = np.ascontiguousarray(g["bcs0"])
    V = np.ascontiguousarray(g["vcs"][None, :] + err[None, :] * rng.standard_normal((T, n)))
    lam = np.ascontiguousarray(10.0 ** np.linspace(-8, 0, Lm))
    # reference: one GPU, round-1 entry point
    p1 = _problem(L, lib, _ctx(), g["ecm"], float(g["threshold"]), g["settings"], G=g["G_spread"])
    L.check(lib.isr_tikhonov_prepare(p1, L.dptr(err), 1))
    st1 = np.zeros((Lm, 3)); c1 = np.zeros((Lm, 130), dtype=np.uint32)
    L.check(lib.isr_tikhonov_toy_scan_hist(p1, Lm, L.dptr(lam), T, L.dptr(V), L.dptr(b0), 128, L.dptr(st1), L.uptr(c1)))
    # sharded
    ctxs = (C.c_void_p * ngpu)()
    for r in range(ngpu):
        h = C.c_void_p(); L.check(lib.isr_ctx_create(r, C.byref(h))); ctxs[r] = h
    comms = (C.c_void_p * ngpu)()
    L.check(lib.isr_comm_init_all(ngpu, ctxs, comms))
    res, errs = [None] * ngpu, []

    def work(r):
        try:
            p = _problem(L, lib, ctxs[r], g["ecm"], float(g["threshold"]), g["settings"], G=g["G_spread"])
            t0, t1 = r * T // ngpu, (r + 1) * T // ngpu
            Vr = np.ascontiguousarray(V[t0:t1])
            st = np.zeros((Lm, 3)); cn = np.zeros((Lm, 130), dtype=np.uint32); tot = C.c_int64()
            a = L.TikhonovScanArgs()
            a.flags, a.vcs_err, a.deriv_reg = L.ISR_TEST_FACTOR, L.dptr(err), 1
            a.n_lambda, a.lambda_, a.n_toys, a.V, a.bcs0, a.nbins, a.comm = Lm, L.dptr(lam), t1 - t0, L.dptr(Vr), L.dptr(b0), 128, comms[r]
            a.stats_out, a.counts_out, a.total_toys_out = L.dptr(st), L.uptr(cn), C.pointer(tot)
            L.check(lib.isr_tikhonov_chi2_scan(p, C.byref(a)))
            res[r] = (st, cn, tot.value)
            lib.isr_problem_destroy(p)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(r,)) for r in range(ngpu)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    for r in range(ngpu):
        st, cn, tot = res[r]
        assert tot == T
        np.testing.assert_array_equal(cn, c1)
        np.testing.assert_array_equal(st[:, :2], st1[:, :2])
        np.testing.assert_allclose(st[:, 2], st1[:, 2], rtol=1e-12)
    for r in range(ngpu):
        lib.isr_comm_destroy(comms[r]); lib.isr_ctx_destroy(ctxs[r])
    lib.isr_problem_destroy(p1)


@pytest.mark.parametrize("ngpu", [1, 2, 4])
def test_group_scans_equal_single_gpu(ngpu):
    """The two other sharded loops from ONE process: the condition-number-vs-spread scan (settings cut over the GPUs, no
    exchange) and BASELINE config C4 (per-lambda chi2 test, toys cut over the GPUs) -- identical to the single-GPU calls."""
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    L, lib = _lib()
    g = load_golden("c4_tikhonov_n32_spread")
    n = len(g["ecm"])
    st = np.ascontiguousarray(g["settings"], dtype=np.int32)
    ecm = np.ascontiguousarray(g["ecm"])
    grp = _group(L, lib, ngpu)
    gp = C.c_void_p()
    L.check(lib.isr_group_problem_create(grp, n, L.dptr(ecm), float(g["threshold"]), len(st), L.iptr(st), None, C.byref(gp)))
    p1 = _problem(L, lib, _ctx(), ecm, float(g["threshold"]), st)
    # --- cond scan: 11 settings (ragged slices), first one with the spread off
    K = 11
    sig = np.ascontiguousarray(np.linspace(0.0, 0.004, K))
    c1, s1, cg, sg = np.zeros(K), np.zeros((K, n)), np.zeros(K), np.zeros((K, n))
    L.check(lib.isr_cond_spread_scan(p1, K, L.dptr(sig), 0, L.dptr(c1), L.dptr(s1)))
    L.check(lib.isr_group_cond_spread_scan(gp, K, L.dptr(sig), 0, L.dptr(cg), L.dptr(sg)))
    np.testing.assert_array_equal(sg, s1)
    np.testing.assert_array_equal(cg, c1)
    # --- C4: assemble (with spread) + eigen-system + chi2 test of 3001 toys for 19 lambdas, in one call
    T, Lm = 3001, 19
    rng = np.random.default_rng(8)
    err = np.ascontiguousarray(g["vcs_err"]); b0 = np.ascontiguousarray(g["bcs0"]); eerr = np.ascontiguousarray(g["ecm_err"])
    V = np.ascontiguousarray(g["vcs"][None, :] + err[None, :] * rng.standard_normal((T, n)))
    lam = np.ascontiguousarray(10.0 ** np.linspace(-8, 0, Lm))

    def scan(call, prob):
        stt = np.zeros((Lm, 3)); cn = np.zeros((Lm, 130), dtype=np.uint32); tot = C.c_int64()
        a = L.TikhonovScanArgs()
        a.flags, a.ecm_err, a.vcs_err, a.deriv_reg = L.ISR_TEST_ASSEMBLE | L.ISR_TEST_FACTOR, L.dptr(eerr), L.dptr(err), 1
        a.n_lambda, a.lambda_, a.n_toys, a.V, a.bcs0, a.nbins = Lm, L.dptr(lam), T, L.dptr(V), L.dptr(b0), 128
        a.stats_out, a.counts_out, a.total_toys_out = L.dptr(stt), L.uptr(cn), C.pointer(tot)
        L.check(call(prob, C.byref(a)))
        return stt, cn, tot.value
    s_one, c_one, t_one = scan(lib.isr_tikhonov_chi2_scan, p1)
    s_grp, c_grp, t_grp = scan(lib.isr_group_tikhonov_chi2_scan, gp)
    assert t_one == T and t_grp == T and int(c_grp.sum()) == Lm * T
    np.testing.assert_array_equal(c_grp, c_one)
    np.testing.assert_array_equal(s_grp[:, :2], s_one[:, :2])
    np.testing.assert_allclose(s_grp[:, 2], s_one[:, 2], rtol=1e-12)
    lib.isr_problem_destroy(p1)
    lib.isr_group_problem_destroy(gp)
    lib.isr_group_destroy(grp)


@pytest.mark.parametrize("ngpu", [1, 2])
def test_python_mirror_fused_call_and_group(ngpu, oracle):
    """The ctypes mirror's chi2_test_call: one isr_chi2_test per test (and isr_group_chi2_test for ngpu > 1), same numbers as
    the call-by-call methods; Tikhonov solvers are refused for multi-GPU (their selected operator is not replicated)."""
    import isrsolver_b200 as isr
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    g = load_golden("c2_sle_n40_cspline")
    n = len(g["ecm"])
    s = isr.ISRSolverSLE(n, g["ecm"], g["vcs"], None, g["vcs_err"], float(g["threshold"]))
    s.set_interp_settings([tuple(r) for r in g["settings"]])
    V = oracle.draw_toys(g["vcs"], g["vcs_err"], 20011)
    r = s.chi2_test_call(g["bcs0"], V=V, want_sum=True, want_ratio=True, want_hist=True, ngpu=ngpu)
    ref = s.toy_batch(V, g["bcs0"], want_sum=True, want_ratio=True, want_hist=True)
    np.testing.assert_array_equal(r["chi2"], ref["chi2"])
    np.testing.assert_array_equal(r["ratio_hist"], ref["ratio_hist"])
    np.testing.assert_allclose(r["sum_bcs"], ref["sum_bcs"], rtol=1e-12)
    cnt, _, lo, hi = oracle.chi2_histogram(ref["chi2"])
    assert (r["lo"], r["hi"]) == (lo, hi) and r["n"] == len(V)
    np.testing.assert_array_equal(r["hist"], cnt)
    d = s.chi2_test_call(g["bcs0"], n=5000, vcs=g["vcs"], seed=3, ngpu=ngpu)
    one = s.toy_campaign(5000, g["vcs"], g["vcs_err"], g["bcs0"], seed=3)
    np.testing.assert_array_equal(d["chi2"], one["chi2"])
    np.testing.assert_array_equal(d["hist"], one["hist"])
    if ngpu > 1:
        t = isr.ISRSolverTikhonov(n, g["ecm"], g["vcs"], None, g["vcs_err"], float(g["threshold"]))
        with pytest.raises(ValueError):
            t.chi2_test_call(g["bcs0"], V=V[:10], ngpu=ngpu)
