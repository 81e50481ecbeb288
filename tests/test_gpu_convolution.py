"""Forward convolution service on the B200 against the CPU side.

convolutionKuraevFadin (src/KuraevFadin.cpp:56-73) through isr_conv_plan_*: tolerance 1e-8 relative
(north star: visible / Born cross sections within 1e-8), measured ~1e-12.  Where the prebuilt compiled
reference (oracle/_ref) is present it is the comparison target; the golden fixtures' vcs_ref array (its
output for the Breit-Wigner model) always is."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-8
THR = 0.827


def bw(e, m=1.5, g=0.25):
    e = np.asarray(e, dtype=np.float64)
    ps = np.clip(1.0 - THR ** 2 / np.maximum(e, 1e-300) ** 2, 0.0, None)
    return np.where(e > THR, 4.0 * m * m * g * g / ((e * e - m * m) ** 2 + m * m * g * g) * ps ** 1.5, 0.0)


def test_golden_visible_cross_section():
    """sigma_vis of the Breit-Wigner model at the C1 energies == the compiled reference's output."""
    import isrsolver_b200 as isr
    g = load_golden("c1_sle_n50_linear")
    plan = isr.ConvolutionPlan(g["ecm"], threshold=float(g["threshold"]))
    vis, err = plan.convolve(bw)
    np.testing.assert_allclose(vis, g["vcs_ref"], rtol=RTOL)
    assert np.all(err <= 1e-9 * np.abs(vis) + 1e-12)
    # scalar form of the PyISR function
    e = float(g["ecm"][17])
    assert isr.convolutionKuraevFadin(e, bw, 0.0, 1 - (THR / e) ** 2) == pytest.approx(g["vcs_ref"][17], rel=RTOL)


def test_convolution_of_interpolant_equals_matrix_row(oracle):
    """int F eps sum_j b_j phi_j = sum_j G_ij b_j: the convolution service and the matrix assembly integrate the
    same integrand with different panel rules (GK15 adaptive vs GL16 fixed)."""
    import isrsolver_b200 as isr
    g = load_golden("c5_mixed3_n36")
    n = len(g["ecm"])
    s = isr.ISRSolverSLE(n, g["ecm"], g["vcs"], None, g["vcs_err"], THR)
    s.set_interp_settings([tuple(r) for r in g["settings"]])
    s.solve()
    plan = isr.ConvolutionPlan(g["ecm"], threshold=THR, breaks=g["ecm"])       # cut at the spline knots
    direct = s.intop_matrix() @ s.bcs()
    # host-sampled interpolant (adaptive) and device-sampled interpolant (one pass on the refined grid)
    vis, _ = plan.convolve(lambda e: s.interp_eval(e))
    np.testing.assert_allclose(vis, direct, rtol=RTOL)
    vis2, _ = plan.convolve_interp(s)
    np.testing.assert_allclose(vis2, direct, rtol=RTOL)
    np.testing.assert_allclose(vis2, g["vcs"], rtol=1e-6)      # and the solution reproduces the data


@pytest.mark.parametrize("width", [0.25, 0.01, 0.0008])
def test_narrow_resonance_vs_compiled_reference(width):
    """A resonance much narrower than any fixed panel is resolved by the error-driven bisection."""
    import isrsolver_b200 as isr
    from oracle import isr_ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libisr_ref.so not present")
    f = lambda e: bw(e, 1.02, width)
    ens = np.array([1.0, 1.0195, 1.021, 1.05, 1.4, 2.0])
    got = isr.convolutionKuraevFadin(ens, f, 0.0, 1 - (THR / ens) ** 2, breaks=[1.02])
    ref = np.array([R.convolution_kf(e, lambda x: float(f(x)), 0.0, 1 - (THR / e) ** 2) for e in ens])
    np.testing.assert_allclose(got, ref, rtol=RTOL)
    # without the hint the refinement still finds the peak
    got2 = isr.convolutionKuraevFadin(ens, f, 0.0, 1 - (THR / ens) ** 2)
    np.testing.assert_allclose(got2, ref, rtol=RTOL)


def test_limits_and_efficiency_forms_vs_compiled_reference(oracle):
    import isrsolver_b200 as isr
    from oracle import isr_ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libisr_ref.so not present")
    O = oracle
    rng = np.random.default_rng(8)
    ens = np.array([1.19, 1.5, 1.93])
    f = lambda x: float(bw(x))
    # min_x above x0: a single integrate() call in the reference (src/KuraevFadin.cpp:70-72)
    lo, hi = np.array([0.01, 0.002, 0.1]), 1 - (THR / ens) ** 2
    got = isr.convolutionKuraevFadin(ens, bw, lo, hi)
    ref = np.array([R.convolution_kf(e, f, a, b) for e, a, b in zip(ens, lo, hi)])
    np.testing.assert_allclose(got, ref, rtol=RTOL)
    # TEfficiency 2-D table
    vals = 0.2 + 0.8 * rng.random((6 + 2, 9 + 2))
    kw = dict(values=vals, nx=9, xlo=0.0, xhi=0.9, eedges=[1.0, 1.2, 1.3, 1.5, 1.7, 1.95, 2.2])
    got = isr.convolutionKuraevFadin(ens, bw, 0.0, hi, efficiency=isr.Efficiency("table_2d", **kw))
    eo = O.Efficiency("table_2d", **kw)
    ref = np.array([R.convolution_kf(e, f, 0.0, b, eo) for e, b in zip(ens, hi)])
    np.testing.assert_allclose(got, ref, rtol=2e-6)     # the reference's retry loop stops at epsrel ~1e-6 on step functions
    # a smooth callable efficiency eps(x, E), as the reference's Python binding accepts
    eff = lambda x, e: 0.3 + 0.5 * np.exp(-30.0 * x) * (1 - 0.1 * (e - 1.5))
    got = isr.convolutionKuraevFadin(ens, bw, 0.0, hi, efficiency=eff)
    ref = []
    for e, b in zip(ens, hi):
        fe = lambda x, _e=e: float(bw(_e * np.sqrt(1 - x))) * R.kernel_kf(x, _e * _e) * float(eff(x, _e))
        x0 = 4 * O.ELECTRON_M / e
        ref.append(R.integrate(fe, 0.0, x0, singular=True)[0] + R.integrate(fe, x0, b)[0])
    np.testing.assert_allclose(got, ref, rtol=RTOL)


def test_blurred_convolution(oracle):
    import isrsolver_b200 as isr
    from oracle import isr_ref as R
    e, sg = 1.3, 0.004
    got = isr.convolutionKuraevFadinBlurred(e, bw, THR, sg)
    if R.available():
        vis = lambda en: R.convolution_kf(en, lambda x: float(bw(x)), 0.0, 1 - (THR / en) ** 2)
        assert got == pytest.approx(R.gaussian_conv(e, sg * sg, vis), rel=RTOL)
    plain = isr.convolutionKuraevFadin(e, bw, 0.0, 1 - (THR / e) ** 2)
    assert abs(got - plain) < 5e-3 * plain and got != plain


def test_plan_reuse_is_deterministic_and_validates():
    import isrsolver_b200 as isr
    ens = np.linspace(0.9, 3.0, 257)
    plan = isr.ConvolutionPlan(ens, threshold=THR)
    a, _ = plan.convolve(bw)
    npan = plan.size()[0]
    b, _ = plan.convolve(bw, refine=False)           # same refined grid, single pass
    np.testing.assert_array_equal(a, b)
    assert plan.size()[0] == npan
    c, _ = plan.convolve(lambda e: 2.0 * bw(e), refine=False)
    np.testing.assert_allclose(c, 2.0 * a, rtol=1e-14)   # linearity in f
    with pytest.raises(ValueError):
        isr.ConvolutionPlan([1.5], min_x=0.0, max_x=1.0)
    with pytest.raises(ValueError):
        isr.ConvolutionPlan([-1.0], threshold=THR)


def test_node_weights_reproduce_the_result():
    """sum_nodes w f(E') == the plan's own result: the weights are the convolution OPERATOR on the node set."""
    import ctypes as C
    import isrsolver_b200 as isr
    from isrsolver_b200 import _lib as L
    ens = np.linspace(1.0, 2.0, 33)
    plan = isr.ConvolutionPlan(ens, threshold=THR)
    res, _ = plan.convolve(bw)
    L.check(L.load().isr_conv_plan_reset(plan._p))
    e, x, r, w = plan.nodes(want_weights=True)
    op = np.zeros(len(ens))
    np.add.at(op, r, w * bw(e))
    np.testing.assert_allclose(op, res, rtol=1e-13)
    assert np.all((x >= 0) & (x < 1)) and np.allclose(e, ens[r] * np.sqrt(1 - x), rtol=1e-15)
