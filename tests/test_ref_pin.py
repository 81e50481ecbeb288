"""Pin the oracle against the REFERENCE'S OWN compiled code (no GPU).

oracle/_ref/libisr_ref.so is built by `make -C oracle` from the reference's unmodified assembly-path
sources (src/KuraevFadin.cpp, Integration.cpp, {Base,Linear,CSpline}RangeInterpolator{,2}.cpp,
Interpolator{,2}.cpp), with the absent GSL / Eigen resolved by oracle/ref_shim.  These tests run the same
inputs through the restatement (oracle/isr_oracle.c) and through that library and require BIT-IDENTICAL
results: formula order of kernelKuraevFadin, the x0 split and the retry loops of integrate/integrateS, the
range bookkeeping, the hat / cardinal-spline bases and the N^2 fill loop are then the reference's own.
(The restated GSL internals underneath both are pinned separately against SciPy in test_oracle.py.)
The library exists only where /root/reference does; elsewhere the golden fixtures' *_ref arrays stand in."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import isr_oracle as O
from oracle import isr_ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libisr_ref.so not built (needs /root/reference)")

N = 18
ECM = np.linspace(O.E_MIN, O.E_MAX, N)
SETTINGS = [None, [(True, 0, N - 1)], [(False, 0, 0), (True, 1, N - 1)],
            [(False, 0, 3), (True, 4, 10), (False, 11, N - 1)], [(True, 0, 5), (True, 6, 11), (False, 12, N - 1)]]


def test_kernel_bit_identical():
    xs = np.concatenate([10.0 ** np.linspace(-14, -0.001, 300), np.linspace(1e-3, 0.6, 200),
                         [0.00135, 0.001362663808, 0.00137]])
    for E in (0.85, 1.18, 1.5, 2.0, 3.1, 10.58):
        a = np.array([O.kernel_kf(x, E * E) for x in xs])
        b = np.array([R.kernel_kf(x, E * E) for x in xs])
        np.testing.assert_array_equal(a, b)
    # SURVEY 8c spot values through the reference's own function
    assert R.kernel_kf(1e-6, 2.25) == pytest.approx(28082.309274929197, rel=1e-13)
    assert R.kernel_kf(0.01, 2.25) == pytest.approx(5.283686571158812, rel=1e-13)
    assert R.kernel_kf(0.5, 2.25) == pytest.approx(0.08908917563287333, rel=1e-13)


@pytest.mark.parametrize("settings", SETTINGS)
def test_matrix_bit_identical(settings):
    Go = O.Interp(ECM, O.E_THR, settings).eval_eq_matrix(mode="faithful")
    Gr = R.RefInterp(ECM, O.E_THR, settings).eval_eq_matrix()
    np.testing.assert_array_equal(Go, Gr)


def test_matrix_unequal_spacing_bit_identical():
    rng = np.random.default_rng(11)
    ecm = np.sort(O.E_THR + 0.05 + 1.4 * rng.random(15))
    st = [(False, 0, 2), (True, 3, 14)]
    np.testing.assert_array_equal(O.Interp(ecm, O.E_THR, st).eval_eq_matrix(mode="faithful"),
                                  R.RefInterp(ecm, O.E_THR, st).eval_eq_matrix())


@pytest.mark.parametrize("kind", ["row_table", "table_2d", "e_only", "row_var_edges"])
def test_matrix_with_efficiency_bit_identical(kind):
    n = 12
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    rng = np.random.default_rng(3)
    v2 = kind.startswith("row")
    if kind == "row_table":
        eff = O.row_table_efficiency(ecm, 24)
    elif kind == "row_var_edges":
        edges = np.concatenate([[0.0], np.sort(rng.random(10)) * 0.7, [0.75]])
        eff = O.Efficiency("row_table", values=0.3 + 0.6 * rng.random((n, len(edges) + 1)), xedges=edges)
    elif kind == "table_2d":
        eff = O.Efficiency("table_2d", values=0.2 + 0.8 * rng.random((6 + 2, 9 + 2)), nx=9, xlo=0.0, xhi=0.9,
                           eedges=[1.0, 1.2, 1.3, 1.5, 1.7, 1.95, 2.2])
    else:
        eff = O.Efficiency("e_only", values=0.2 + 0.8 * rng.random(7 + 2), ne=7, elo=1.0, ehi=2.1)
    st = [(False, 0, 4), (True, 5, n - 1)]
    Go = O.Interp(ecm, O.E_THR, st).eval_eq_matrix(eff=eff, mode="faithful")
    Gr = R.RefInterp(ecm, O.E_THR, st, v2=v2).eval_eq_matrix(eff=eff)
    np.testing.assert_array_equal(Go, Gr)


@pytest.mark.parametrize("settings", SETTINGS)
def test_basis_queries_bit_identical(settings):
    oi, ri = O.Interp(ECM, O.E_THR, settings), R.RefInterp(ECM, O.E_THR, settings)
    es = np.concatenate([ECM, [O.E_THR, O.E_THR - 0.1, O.E_THR + 1e-9, ECM[-1] + 0.3],
                         np.random.default_rng(2).uniform(O.E_THR, ECM[-1], 40)])
    for j in range(N):
        for e in es:
            assert oi.basis_eval(j, e) == ri.basis_eval(j, e), (j, e)
            if O.E_THR < e <= ECM[-1]:
                assert oi.basis_deriv_eval(j, e) == ri.basis_deriv_eval(j, e), (j, e)
    y = O.bw_born(ECM)
    for e in es:
        assert oi.eval(y, e) == ri.eval(y, e)
    np.testing.assert_array_equal(oi.integral_basis(), ri.integral_basis())
    assert ri.min_energy() == O.E_THR and ri.max_energy() == ECM[-1]


def test_spread_matrix_bit_identical():
    st = [(False, 0, 0), (True, 1, N - 1)]
    err = np.linspace(0.001, 0.004, N)
    np.testing.assert_array_equal(O.Interp(ECM, O.E_THR, st).energy_spread_matrix(err),
                                  R.RefInterp(ECM, O.E_THR, st).energy_spread_matrix(err))


@pytest.mark.parametrize("settings,ok", [
    ([(False, 0, N - 1)], True), ([(True, 3, N - 1), (False, 0, 2)], True),          # unsorted input is sorted
    ([(False, 1, N - 1)], False), ([(False, 0, N - 2)], False),                       # must cover 0 .. N-1
    ([(False, 0, 5), (True, 5, N - 1)], False), ([(False, 0, 4), (True, 6, N - 1)], False),  # overlap / gap
    ([(False, 0, 5), (True, 9, 6), (False, 10, N - 1)], False),                       # reversed range
])
def test_range_validation_agrees(settings, ok):
    """checkRangeInterpSettings (src/Interpolator.cpp:90-147): both accept / reject the same settings."""
    for make in (lambda: O.Interp(ECM, O.E_THR, settings), lambda: R.RefInterp(ECM, O.E_THR, settings)):
        if ok:
            make()
        else:
            with pytest.raises(O.InterpRangeException):
                make()


def test_convolution_and_integrators_bit_identical():
    born = lambda e: float(O.bw_born(e)[0])
    for E in (1.2, 1.5, 1.97):
        ref = R.convolution_kf(E, born, 0.0, 1.0 - (O.E_THR / E) ** 2)
        assert ref == O.bw_visible(np.array([E]))[0]
    # known answers of SURVEY 8c through the reference's integrateS / integrate and its own kernel
    E = 1.5
    x0 = 4 * O.ELECTRON_M / E
    f = lambda x: R.kernel_kf(x, E * E)
    a, _ = R.integrate(f, 0.0, x0, singular=True)
    b, _ = R.integrate(f, x0, 1.0 - (0.827 / E) ** 2)
    assert a == pytest.approx(0.66700413602970, rel=2e-13)
    assert b == pytest.approx(0.32585579807259912, rel=2e-13)
    # gaussian_conv of a quadratic is exact for the 6-point rule: E^2 + sigma^2
    assert R.gaussian_conv(1.3, 0.002 ** 2, lambda e: e * e) == pytest.approx(1.3 ** 2 + 0.002 ** 2, rel=1e-14)


@pytest.mark.parametrize("name", ["c1_sle_n50_linear", "c2_sle_n40_cspline", "c3_sle2_n24_eps64",
                                  "c4_tikhonov_n32_spread", "c5_mixed3_n36"])
def test_golden_fixtures_are_reference_output(name):
    """The committed fixtures' G_ref is what the compiled reference produces today."""
    g = load_golden(name)
    eff = None
    if "eff_values" in g:
        eff = O.Efficiency("row_table", values=g["eff_values"], nx=int(g["eff_nx"]), xlo=float(g["eff_xlo"]), xhi=float(g["eff_xhi"]))
    ri = R.RefInterp(g["ecm"], float(g["threshold"]), [tuple(r) for r in g["settings"]], v2=eff is not None)
    rows = [0, len(g["ecm"]) // 2, len(g["ecm"]) - 1]
    G = ri.eval_eq_matrix(eff=eff, rows=rows)
    np.testing.assert_array_equal(G[rows], g["G_ref"][rows])
    np.testing.assert_array_equal(g["G_ref"], g["G_faithful"])
