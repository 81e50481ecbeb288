"""Matrix assembly parity on the B200 (through the C ABI via the PyISR mirror).

Tolerance (north star): matrix entries within 1e-9 relative.  The comparison target is the oracle's
CONVERGED back-end; for structurally tiny entries the oracle's own absolute tolerance (epsabs = 1e-12 per
adaptive call) is added as a floor.  Against the FAITHFUL back-end (the reference's call structure, whose
loosen-until-success loop ends at epsrel up to 1e-3 for tabulated efficiencies) the gate is the same plus
the faithful back-end's own reported error estimate.  The fixtures' G_ref / S_ref / vcs_ref arrays are outputs
of the reference's OWN compiled sources (oracle/_ref, see tests/golden/make_golden.py); the faithful back-end
reproduces them bit for bit (tests/test_ref_pin.py), so the faithful gate is the gate against the reference."""
import numpy as np
import pytest

from conftest import assert_matrix_parity, load_golden

pytestmark = pytest.mark.gpu


def make_solver(g, cls=None, eff=None, spread=False):
    import isrsolver_b200 as isr
    cls = cls or isr.ISRSolverSLE
    n = len(g["ecm"])
    s = cls(n, g["ecm"], g["vcs"], g.get("ecm_err") if spread else None, g["vcs_err"], float(g["threshold"]),
            efficiency=eff, enableEnergySpread=spread)
    s.set_interp_settings([tuple(r) for r in g["settings"]])
    return s


def test_kernel_values_on_device(oracle):
    import isrsolver_b200 as isr
    xs = np.array([1e-8, 1e-6, 1e-4, 0.00135, 0.00137, 0.002, 0.01, 0.3, 0.5, 0.8])
    for E in (1.18, 1.5, 2.0):
        got = isr.kernelKuraevFadin(xs, E * E)
        ref = np.array([oracle.kernel_kf(x, E * E) for x in xs])
        np.testing.assert_allclose(got, ref, rtol=5e-13)
    assert isr.kernelKuraevFadin(0.01, 2.25) == pytest.approx(5.283686571158812, rel=1e-13)


@pytest.mark.parametrize("name", ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c5_mixed3_n36"])
def test_matrix_vs_golden(name):
    g = load_golden(name)
    s = make_solver(g)
    s.eval_equation_matrix()
    G = s.intop_matrix()
    assert_matrix_parity(G, g["G_converged"], g["G_converged_err"], what=name + " vs converged oracle")
    assert_matrix_parity(G, g["G_faithful"], 3 * g["G_faithful_err"], row_rel=1e-9, what=name + " vs faithful oracle")
    assert_matrix_parity(G, g["G_ref"], 3 * g["G_faithful_err"], row_rel=1e-9, what=name + " vs the compiled reference")
    if "linear" in name:
        assert np.all(np.triu(G, 1) == 0)       # structural zeros are exact (F1)


def test_matrix_with_row_table_efficiency():
    import isrsolver_b200 as isr
    g = load_golden("c3_sle2_n24_eps64")
    eff = isr.Efficiency("row_table", values=g["eff_values"], nx=int(g["eff_nx"]), xlo=float(g["eff_xlo"]), xhi=float(g["eff_xhi"]))
    s = make_solver(g, eff=eff)
    s.eval_equation_matrix()
    assert_matrix_parity(s.intop_matrix(), g["G_converged"], g["G_converged_err"], what="eps table vs converged")
    assert_matrix_parity(s.intop_matrix(), g["G_faithful"], 3 * g["G_faithful_err"], row_rel=1e-9, what="eps table vs faithful")
    assert_matrix_parity(s.intop_matrix(), g["G_ref"], 3 * g["G_faithful_err"], row_rel=1e-9, what="eps table vs the compiled reference (Interpolator2)")
    # same through the v2 class (ISRSolverSLE2: histogram per energy point)
    s2 = isr.ISRSolverSLE2(len(g["ecm"]), g["ecm"], g["vcs"], None, g["vcs_err"], float(g["threshold"]),
                           eff_hists=g["eff_values"], xlo=0.0, xhi=1.0)
    s2.eval_equation_matrix()
    np.testing.assert_array_equal(s2.intop_matrix(), s.intop_matrix())


def test_matrix_with_energy_spread():
    g = load_golden("c4_tikhonov_n32_spread")
    s = make_solver(g, spread=True)
    s.eval_equation_matrix()
    assert_matrix_parity(s.intop_matrix(), g["G_spread"], 32 * g["G_converged_err"].max(), what="spread matrix")
    # the compiled reference's _energySpreadMatrix times its own matrix
    assert_matrix_parity(s.intop_matrix(), g["S_ref"] @ g["G_ref"], 32 * 3 * g["G_faithful_err"].max(), row_rel=1e-9,
                         what="spread matrix vs the compiled reference")


@pytest.mark.parametrize("kind", ["table_2d", "e_only", "row_var_edges"])
def test_efficiency_kinds_against_oracle(oracle, kind):
    """TEfficiency 2-D / 1-D and variable-width bins: bin lookups follow ROOT's arithmetic."""
    import isrsolver_b200 as isr
    n = 14
    ecm = np.linspace(1.18, 2.0, n); thr = 0.827
    rng = np.random.default_rng(5)
    if kind == "table_2d":
        kw = dict(values=0.2 + 0.8 * rng.random((6 + 2, 9 + 2)), nx=9, xlo=0.0, xhi=0.9, eedges=[1.0, 1.2, 1.3, 1.5, 1.7, 1.95, 2.2])
        k = "table_2d"
    elif kind == "e_only":
        kw = dict(values=0.2 + 0.8 * rng.random(5 + 2), ne=5, elo=1.1, ehi=1.9)   # last energies overflow
        k = "e_only"
    else:
        edges = np.array([0.0, 0.001, 0.004, 0.02, 0.1, 0.3, 0.5, 0.95])
        kw = dict(values=0.2 + 0.8 * rng.random((n, len(edges) + 1)), xedges=edges)
        k = "row_table"
    eo = oracle.Efficiency(k, **kw)
    ei = isr.Efficiency(k, **kw)
    Gc, Ec, _ = oracle.Interp(ecm, thr, [(0, 0, 5), (1, 6, n - 1)]).eval_eq_matrix(eff=eo, mode="converged", with_errors=True)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr, efficiency=ei)
    s.set_interp_settings([(0, 0, 5), (1, 6, n - 1)])
    s.eval_equation_matrix()
    if kind != "table_2d":
        assert_matrix_parity(s.intop_matrix(), Gc, Ec, what=kind)
    else:
        # the converged oracle splits at the x edges only for row tables; for the 2-D table build the
        # equivalent row table (row i = the E-bin of E_i) and compare with that
        rows = np.stack([eo.values.reshape(8, 11)[oracle.find_bin(0, 0, 0, kw["eedges"], e)] for e in ecm])
        eo2 = oracle.Efficiency("row_table", values=rows, nx=9, xlo=0.0, xhi=0.9)
        Gc2, Ec2, _ = oracle.Interp(ecm, thr, [(0, 0, 5), (1, 6, n - 1)]).eval_eq_matrix(eff=eo2, mode="converged", with_errors=True)
        assert_matrix_parity(s.intop_matrix(), Gc2, Ec2, what=kind)


def test_panel_lookups_bit_exact(oracle):
    """Index lookups must be bit-exact: every panel's efficiency value equals values[row][FindBin(x)] for
    x anywhere strictly inside the panel, and its (row, segment) equals the oracle's segment search."""
    import ctypes as C
    import isrsolver_b200 as isr
    from isrsolver_b200 import _lib as L
    n = 16; ecm = np.linspace(1.18, 2.0, n); thr = 0.827
    vals = oracle.row_table_efficiency(ecm, 37).values
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr,
                         efficiency=isr.Efficiency("row_table", values=vals, nx=37, xlo=0.0, xhi=1.0))
    s.eval_equation_matrix()
    lib = L.load()
    cnt = C.c_int64()
    L.check(lib.isr_debug_panels(s._p, 0, None, None, C.byref(cnt)))
    ab = np.empty((cnt.value, 4)); rs = np.empty((cnt.value, 2), dtype=np.int32)
    L.check(lib.isr_debug_panels(s._p, cnt.value, L.dptr(ab), L.iptr(rs), C.byref(cnt)))
    ext = np.concatenate([[thr], ecm])
    rng = np.random.default_rng(0)
    xs = ab[:, 0] + (ab[:, 1] - ab[:, 0]) * rng.uniform(0.02, 0.98, len(ab))
    for (a, b, ev, kind), (i, k), x in zip(ab, rs, xs):
        if b <= a:
            continue
        assert ev == vals[i, oracle.find_bin(37, 0.0, 1.0, None, x)]
        e = ecm[i] * np.sqrt(1 - x)
        kk = int(np.searchsorted(ext, e, side="left")) - 1       # ext[kk] < e <= ext[kk+1]
        assert kk == k
        x0 = 4 * oracle.ELECTRON_M / ecm[i]
        assert int(kind) == (1 if b <= x0 else (2 if b <= x0 * (1 + 1 / 16) else 0))
    # panels tile [0, xmax_i] exactly, row by row
    for i in range(n):
        m = rs[:, 0] == i
        a, b = ab[m, 0], ab[m, 1]
        o = np.argsort(a)
        assert a[o][0] == 0.0 and np.all(a[o][1:] == b[o][:-1])
        assert b[o][-1] == pytest.approx(1 - (thr / ecm[i]) ** 2, rel=1e-15)


def test_invalid_ranges_raise():
    import isrsolver_b200 as isr
    n = 8; ecm = np.linspace(1.2, 2.0, n)
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.9)
    for bad in ([(False, 1, 7)], [(False, 0, 6)], [(False, 0, 2), (True, 4, 7)], [(False, 0, 3), (True, 3, 7)]):
        with pytest.raises(isr.InterpRangeException):
            s.set_interp_settings(bad)
    s.set_interp_settings([(True, 4, 7), (False, 0, 3)])      # unsorted input is fine (sorted internally)


def test_input_sorting():
    """Points given out of order are sorted by energy (src/BaseISRSolver.cpp:105-128)."""
    import isrsolver_b200 as isr
    g = load_golden("c1_sle_n50_linear")
    perm = np.random.default_rng(0).permutation(len(g["ecm"]))
    s = isr.ISRSolverSLE(len(perm), g["ecm"][perm], g["vcs"][perm], None, g["vcs_err"][perm], float(g["threshold"]))
    np.testing.assert_array_equal(s.ecm(), g["ecm"]); np.testing.assert_array_equal(s.vcs(), g["vcs"])
    s.solve()
    np.testing.assert_allclose(s.bcs(), g["bcs"], rtol=1e-8)


@pytest.mark.parametrize("n,settings,rows", [(200, "linear", (0, 1, 57, 199)), (200, "mixed", (0, 3, 120, 199)),
                                             (500, "linear", (2, 499))])
def test_full_size_spot_rows(oracle, n, settings, rows):
    """BASELINE sizes: exact lower-triangularity for the linear basis, and selected full rows against the
    converged oracle (the whole N = 200 cubic matrix would take the oracle minutes; rows take seconds)."""
    import isrsolver_b200 as isr
    ecm = np.linspace(1.18, 2.0, n)
    st = [(False, 0, n - 1)] if settings == "linear" else [(False, 0, 0), (True, 1, n - 1)]
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), 0.827)
    s.set_interp_settings(st)
    s.eval_equation_matrix()
    G = s.intop_matrix()
    assert np.all(np.isfinite(G))
    if settings == "linear":
        assert np.all(np.triu(G, 1) == 0) and np.all(G[np.tril_indices(n)] > 0)
    Gr, Er = oracle.Interp(ecm, 0.827, st).eval_rows(rows)
    assert_matrix_parity(G[list(rows)], Gr, Er, what=f"N={n} {settings} rows {rows}")


def test_full_size_eps_table_spot_rows(oracle):
    """Config C3 shape: N = 200 with a 512-bin efficiency histogram per energy point."""
    import isrsolver_b200 as isr
    n = 200
    ecm = np.linspace(1.18, 2.0, n)
    eo = oracle.row_table_efficiency(ecm, 512)
    s = isr.ISRSolverSLE2(n, ecm, np.ones(n), None, np.ones(n), 0.827, eff_hists=eo.values, xlo=0.0, xhi=1.0)
    s.eval_equation_matrix()
    rows = (0, 101, 199)
    Gr, Er = oracle.Interp(ecm, 0.827).eval_rows(rows, eff=eo)
    assert_matrix_parity(s.intop_matrix()[list(rows)], Gr, Er, what="N=200 eps512 rows")


def test_condition_number_vs_spread_scan(oracle):
    """the loop of src/isrsolver-condnum-test.cpp:123-135 -- condition number of
    G_spread(sigma) G for a list of beam-energy spreads, from ONE assembly."""
    import isrsolver_b200 as isr
    g = load_golden("c4_tikhonov_n32_spread")
    n = len(g["ecm"])
    s = make_solver(g)
    sig = np.array([0.0, 0.0005, 0.002, 0.004])
    cond, sv = s.condnum_spread_scan(sig, want_singular_values=True)
    I = oracle.Interp(g["ecm"], float(g["threshold"]), [tuple(r) for r in g["settings"]])
    for k, sg in enumerate(sig):
        Gk = g["G_converged"] if sg == 0 else I.energy_spread_matrix(np.full(n, sg)) @ g["G_converged"]
        ref = np.linalg.svd(Gk, compute_uv=False)
        np.testing.assert_allclose(sv[k], ref, rtol=1e-8, atol=1e-12 * ref[0])
        assert cond[k] == pytest.approx(ref[0] / ref[-1], rel=1e-8 * max(1.0, ref[0] / ref[-1] * 1e-4))
    assert np.all(np.diff(cond) > 0)                 # smearing makes the problem worse conditioned
    # the same numbers through the single-point API of the reference class
    s2 = make_solver(g, spread=True)
    assert s2.condnum_eval() == pytest.approx(cond[2], rel=1e-9)
    s.eval_equation_matrix()
    assert s.condnum_eval() == pytest.approx(cond[0], rel=1e-12)
    # per-point spreads
    pp = np.tile(np.linspace(0.001, 0.003, n), (2, 1)); pp[1] *= 2
    c2 = s.condnum_spread_scan(pp)
    for k in range(2):
        ref = np.linalg.svd(I.energy_spread_matrix(pp[k]) @ g["G_converged"], compute_uv=False)
        assert c2[k] == pytest.approx(ref[0] / ref[-1], rel=1e-7)
    with pytest.raises(ValueError):
        s.condnum_spread_scan([-0.001])


def test_callable_efficiency_sampled_at_the_nodes(oracle):
    """SURVEY 8b `SAMPLED`: eps(x, E) as a Python callable, as the reference's PyISR constructor accepts it
    (include/PyISRSolver.hpp:189-197) -- the host samples it at the GPU's quadrature nodes.  Compared with the
    compiled reference running the SAME callable through its std::function, and with an equivalent fine table."""
    import isrsolver_b200 as isr
    from oracle import isr_ref as R
    n = 12
    ecm = np.linspace(1.18, 2.0, n); thr = 0.827
    eff = lambda x, e: 0.3 + 0.5 * np.exp(-30.0 * x) * (1.0 - 0.1 * (e - 1.5))
    st = [(False, 0, 2), (True, 3, n - 1)]
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr, efficiency=eff)
    s.set_interp_settings(st)
    s.eval_equation_matrix()
    G = s.intop_matrix()
    s1 = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr)
    s1.set_interp_settings(st)
    s1.eval_equation_matrix()
    assert np.all(G[np.tril_indices(n)] < s1.intop_matrix()[np.tril_indices(n)])       # eps < 1 everywhere
    if R.available():
        Gr = R.RefInterp(ecm, thr, st).eval_eq_matrix_callable(eff)
        assert_matrix_parity(G, Gr, None, rel=1e-9, abs_floor=2e-12, row_rel=1e-9, what="callable efficiency vs compiled reference")
    # a scalar-only callable takes the element-wise path and gives the same matrix
    s2 = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr, efficiency=lambda x, e: float(eff(float(x), float(e))))
    s2.set_interp_settings(st)
    s2.eval_equation_matrix()
    np.testing.assert_array_equal(s2.intop_matrix(), G)
    # with the energy spread on top, and the solve path downstream of a sampled matrix
    s3 = isr.ISRSolverSLE(n, ecm, G @ np.ones(n), np.full(n, 0.002), 0.05 * (G @ np.ones(n)), thr, efficiency=eff, enableEnergySpread=True)
    s3.set_interp_settings(st)
    s3.solve()
    S = oracle.Interp(ecm, thr, st).energy_spread_matrix(np.full(n, 0.002))
    np.testing.assert_allclose(s3.intop_matrix(), S @ G, rtol=1e-10, atol=1e-13)


def test_callable_efficiency_with_declared_jumps():
    """A step efficiency given as a callable plus its break points == the same efficiency given as a two-bin table."""
    import isrsolver_b200 as isr
    n = 14
    ecm = np.linspace(1.18, 2.0, n); thr = 0.827
    step = lambda x, e: np.where(x < 0.1, 0.8, 0.4) * (1.0 + 0.05 * (e - 1.5))
    st = [(False, 0, 3), (True, 4, n - 1)]
    s = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr, efficiency=step, efficiency_breaks=[0.1])
    s.set_interp_settings(st)
    s.eval_equation_matrix()
    vals = np.zeros((n, 4))
    vals[:, 1], vals[:, 2] = 0.8 * (1.0 + 0.05 * (ecm - 1.5)), 0.4 * (1.0 + 0.05 * (ecm - 1.5))
    t = isr.ISRSolverSLE2(n, ecm, np.ones(n), None, np.ones(n), thr, eff_hists=vals, xedges=[0.0, 0.1, 1.0])
    t.set_interp_settings(st)
    t.eval_equation_matrix()
    np.testing.assert_allclose(s.intop_matrix(), t.intop_matrix(), rtol=1e-13, atol=1e-16)
    # without the declared jump the panels straddle it and the result is visibly off
    u = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr, efficiency=step)
    u.set_interp_settings(st)
    u.eval_equation_matrix()
    assert np.abs(u.intop_matrix() - t.intop_matrix()).max() > 1e-6
