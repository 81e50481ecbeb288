"""The compiled PyISR module (isrsolver_b200/csrc/pyisr_module.cpp: pybind11 over include/isr_b200.hpp), i.e. the
reference's src/PyISR.cpp re-backed by the C ABI.  Surface checks need no GPU; the numerical checks are `gpu`."""
import os
import re

import numpy as np
import pytest

from conftest import load_golden

# method tables of the reference (include/PyISRSolverSLE.hpp:183-215, include/PyISRSolverTikhonov.hpp:111-146,
# include/PyISRSolverTSVD.hpp:30-81) and module functions (src/PyISR.cpp:197-207)
SLE = ["solve", "save", "bcs", "ecm", "ecm_err", "vcs", "vcs_err", "bcs_cov_matrix", "bcs_inv_cov_matrix", "intop_matrix",
       "condnum_eval", "interp_eval", "eval_equation_matrix", "reset_ecm_err", "set_interp_settings", "chi2_test_model",
       "ratio_test_model", "n", "energy_spread_enabled"]
TIK = SLE + ["make_LCurve", "solve_LCurve", "reg_param"]
TSVD = ["solve", "save", "bcs", "ecm", "ecm_err", "vcs", "vcs_err", "bcs_cov_matrix", "bcs_inv_cov_matrix", "intop_matrix",
        "condnum_eval", "interp_eval", "reset_ecm_err", "set_interp_settings", "enable_keepone", "disable_keepone", "n",
        "energy_spread_enabled", "upper_TSVD_index"]


def test_module_surface_matches_the_reference_tables():
    from isrsolver_b200 import PyISR
    for name in ("kernelKuraevFadin", "convolutionKuraevFadin", "convolutionKuraevFadinBlurred", "ISRSolverSLE",
                 "ISRSolverTikhonov", "ISRSolverTSVD"):
        assert hasattr(PyISR, name), name
    for cls, table in ((PyISR.ISRSolverSLE, SLE), (PyISR.ISRSolverTikhonov, TIK), (PyISR.ISRSolverTSVD, TSVD)):
        missing = [m for m in table if not hasattr(cls, m)]
        assert not missing, (cls.__name__, missing)
    # where the reference tree is present, the tables above are re-derived from its headers
    ref = "/root/reference/include"
    if os.path.isdir(ref):
        for hdr, cls in (("PyISRSolverSLE.hpp", PyISR.ISRSolverSLE), ("PyISRSolverTikhonov.hpp", PyISR.ISRSolverTikhonov),
                         ("PyISRSolverTSVD.hpp", PyISR.ISRSolverTSVD)):
            names = re.findall(r'^\s*\{"(\w+)",\s*\((?:PyCFunction|getter)\)', open(os.path.join(ref, hdr)).read(), re.M)
            assert names and all(hasattr(cls, n) for n in names), (hdr, [n for n in names if not hasattr(cls, n)])


@pytest.mark.gpu
def test_compiled_pyisr_matches_golden(tmp_path):
    from isrsolver_b200 import PyISR
    g = load_golden("c2_sle_n40_cspline")
    n = len(g["ecm"])
    perm = np.random.default_rng(1).permutation(n)            # the constructor sorts by energy
    s = PyISR.ISRSolverSLE(n=n, energy=g["ecm"][perm], vcs=g["vcs"][perm], energy_err=None, vcs_err=g["vcs_err"][perm],
                           threshold=float(g["threshold"]))
    with pytest.raises(ValueError):
        s.set_interp_settings([(False, 1, n - 1)])
    s.set_interp_settings([tuple(bool(r[0]) if k == 0 else int(r[k]) for k in range(3)) for r in g["settings"]])
    assert s.solve() == 0 and s.n == n and not s.energy_spread_enabled
    np.testing.assert_array_equal(s.ecm(), g["ecm"])
    G = s.intop_matrix()
    assert np.all(np.abs(G - g["G_ref"]) <= 1e-9 * np.abs(g["G_ref"]).max(1, keepdims=True) + 3 * g["G_faithful_err"] + 2e-12)
    np.testing.assert_allclose(s.bcs(), g["bcs"], rtol=1e-8)
    assert np.abs(s.bcs_cov_matrix() - g["cov"]).max() <= 1e-8 * np.abs(g["cov"]).max()
    np.testing.assert_allclose(s.bcs_inv_cov_matrix() @ g["cov"], np.eye(n), atol=1e-7)
    assert s.condnum_eval() == pytest.approx(float(g["cond"]), rel=1e-8)
    # against the reference's own compiled classes on the same inputs, nothing injected on either side
    r0 = load_golden("refrun_c2_sle_n40_cspline")
    assert np.abs(s.bcs() - r0["ref_bcs_native"]).max() <= 1e-8 * np.abs(r0["ref_bcs_native"]).max()
    assert np.abs(s.bcs_cov_matrix() - r0["ref_cov_native"]).max() <= 1e-8 * np.abs(r0["ref_cov_native"]).max()
    assert s.interp_eval(float(g["ecm"][7])) == pytest.approx(s.bcs()[7], rel=1e-12)
    # views keep the solver alive
    b = s.bcs()
    model = tmp_path / "model.npz"
    np.savez(model, vcsBlured=np.stack([g["ecm"], g["vcs"], 0 * g["ecm"], g["vcs_err"]]), bcsSRC=g["bcs0"])
    r = s.chi2_test_model(n=2000, initial_ampl=1e4, model_path=str(model), model_visible_cs_name="vcsBlured",
                          model_born_cs_name="bcsSRC", output_path=str(tmp_path / "chi2.npz"))
    assert r["chi2"].shape == (2000,) and r["hist"].sum() == 2000 and r["chi2"].mean() == pytest.approx(n, rel=0.1)
    rr = s.ratio_test_model(n=2000, model_path=str(model), model_visible_cs_name="vcsBlured", model_born_cs_name="bcsSRC")
    # the model visible cross section is the exact convolution, so the means carry the (percent-level)
    # interpolation bias the ratio test exists to expose; compare with the ctypes mirror on the same toys below
    assert np.all(np.abs(rr["means"]) < 0.05) and np.all(rr["mean_errors"] > 0) and rr["hist"].sum() == 2000 * n
    s.save(str(tmp_path / "out.npz"))
    with np.load(tmp_path / "out.npz") as f:
        np.testing.assert_array_equal(f["intergalOperatorMatrix"], G)
        assert set(f.files) >= {"vcs", "bcs", "covMatrixBornCS", "invCovMatrixBornCS"}
    del s
    assert np.all(np.isfinite(b))
    # module functions
    assert PyISR.kernelKuraevFadin(0.01, 2.25) == pytest.approx(5.283686571158812, rel=1e-13)
    from bench import bw_born
    e = float(g["ecm"][11])
    got = PyISR.convolutionKuraevFadin(e, lambda x: float(bw_born(x)), 0.0, 1 - (0.827 / e) ** 2)
    assert got == pytest.approx(g["vcs_ref"][11], rel=1e-8)
    got2 = PyISR.convolutionKuraevFadin(e, lambda x: float(bw_born(x)), 0.0, 1 - (0.827 / e) ** 2, lambda x, en: 0.5)
    assert got2 == pytest.approx(0.5 * g["vcs_ref"][11], rel=1e-8)
    bl = PyISR.convolutionKuraevFadinBlurred(e, lambda x: float(bw_born(x)), 0.827, 0.002)
    assert abs(bl - got) < 1e-3 * got


@pytest.mark.gpu
def test_compiled_pyisr_tikhonov_and_tsvd(oracle):
    from isrsolver_b200 import PyISR
    g = load_golden("c4_tikhonov_n32_spread")
    n = len(g["ecm"])
    t = PyISR.ISRSolverTikhonov(n, g["ecm"], g["vcs"], g["ecm_err"], g["vcs_err"], float(g["threshold"]), None, True)
    t.set_interp_settings([(bool(r[0]), int(r[1]), int(r[2])) for r in g["settings"]])
    lc = t.make_LCurve(g["tik_lambda"])
    # evalEqNorm2 is a residual: the reference forms G b - v (cancellation ~ eps |v| / |G b - v|)
    np.testing.assert_allclose(lc["x"], g["tik_x"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(lc["y"], g["tik_y"], rtol=1e-7)
    np.testing.assert_allclose(lc["c"], -g["tik_curv"], rtol=1e-6)
    assert t.reg_param == g["tik_lambda"][-1]
    t.reg_param = float(g["tik_lambda"][3])
    t.solve()
    np.testing.assert_allclose(t.bcs(), g["tik_bcs"][3], rtol=1e-7, atol=1e-9 * np.abs(g["tik_bcs"][3]).max())
    assert np.abs(t.bcs_cov_matrix() - g["tik_cov3"]).max() <= 1e-7 * np.abs(g["tik_cov3"]).max()
    r = t.solve_LCurve(1e-3)
    assert r["lambda"] > 0 and r["max_curv"] > 0
    v = PyISR.ISRSolverTSVD(n, g["ecm"], g["vcs"], g["ecm_err"], g["vcs_err"], float(g["threshold"]), None, True,
                            upper_TSVD_index=int(g["tsvd_k"]))
    v.set_interp_settings([(bool(r[0]), int(r[1]), int(r[2])) for r in g["settings"]])
    v.solve()
    np.testing.assert_allclose(v.bcs(), g["tsvd_bcs"], rtol=1e-6, atol=1e-8 * np.abs(g["tsvd_bcs"]).max())
    v.enable_keepone(); v.solve()
    np.testing.assert_allclose(v.bcs(), g["tsvd_bcs_keepone"], rtol=1e-6, atol=1e-8 * np.abs(g["tsvd_bcs"]).max())
    assert v.upper_TSVD_index == int(g["tsvd_k"])
    # ISRSolverTSVD of the reference itself (its own assembly, spread product, JacobiSVD, minimum-norm COD solve)
    rr = load_golden("refrun_c4_tikhonov_n32_spread")
    scale = np.abs(rr["ref_tsvd_bcs"]).max()
    assert np.abs(v.bcs() - rr["ref_tsvd_bcs_keepone"]).max() <= 1e-7 * scale
    v.disable_keepone(); v.solve()
    assert np.abs(v.bcs() - rr["ref_tsvd_bcs"]).max() <= 1e-7 * scale


@pytest.mark.gpu
def test_compiled_pyisr_accepts_a_callable_efficiency():
    """the reference's constructor takes efficiency as a Python callable (include/PyISRSolver.hpp:186-197);
    the compiled module wraps it into the std::function overload of the C++ mirror (sampled at the quadrature nodes).
    Same matrix as the ctypes mirror running the same callable, and as the compiled reference."""
    import math
    import isrsolver_b200 as isr
    from isrsolver_b200 import PyISR
    from oracle import isr_ref as R
    n = 12
    ecm = np.linspace(1.18, 2.0, n); thr = 0.827
    eff = lambda x, e: 0.3 + 0.5 * math.exp(-30.0 * x) * (1.0 - 0.1 * (e - 1.5))
    st = [(False, 0, 2), (True, 3, n - 1)]
    s = PyISR.ISRSolverSLE(n=n, energy=ecm, vcs=np.ones(n), energy_err=None, vcs_err=np.ones(n), threshold=thr, efficiency=eff)
    s.set_interp_settings(st)
    s.eval_equation_matrix()
    G = s.intop_matrix()
    m = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr, efficiency=eff)
    m.set_interp_settings(st)
    m.eval_equation_matrix()
    np.testing.assert_array_equal(G, m.intop_matrix())
    if R.available():
        Gr = R.RefInterp(ecm, thr, st).eval_eq_matrix_callable(eff)
        assert np.all(np.abs(G - Gr) <= 1e-9 * np.abs(Gr).max(1, keepdims=True) + 2e-12)
    # a step efficiency with its declared jump == the two-bin table
    class Step:
        breaks = [0.1]
        def __call__(self, x, e):
            return (0.8 if x < 0.1 else 0.4) * (1.0 + 0.05 * (e - 1.5))
    s2 = PyISR.ISRSolverSLE(n=n, energy=ecm, vcs=np.ones(n), energy_err=None, vcs_err=np.ones(n), threshold=thr, efficiency=Step())
    s2.set_interp_settings(st)
    s2.eval_equation_matrix()
    vals = np.zeros((n, 4)); vals[:, 1] = 0.8 * (1.0 + 0.05 * (ecm - 1.5)); vals[:, 2] = 0.4 * (1.0 + 0.05 * (ecm - 1.5))
    t = isr.ISRSolverSLE(n, ecm, np.ones(n), None, np.ones(n), thr,
                         efficiency=isr.Efficiency("row_table", values=vals, nx=2, xedges=np.array([0.0, 0.1, 1.0])))
    t.set_interp_settings(st)
    t.eval_equation_matrix()
    np.testing.assert_allclose(s2.intop_matrix(), t.intop_matrix(), rtol=1e-12, atol=1e-15)


@pytest.mark.gpu
def test_compiled_pyisr_releases_the_gil_during_long_calls(tmp_path):
    """long GPU calls of the compiled module run with the GIL released -- another Python thread makes progress
    while a toy campaign of 4 million toys is under way (the reference's PyISR holds the GIL throughout)."""
    import threading
    from isrsolver_b200 import PyISR
    g = load_golden("c2_sle_n40_cspline")
    n = len(g["ecm"])
    s = PyISR.ISRSolverSLE(n=n, energy=g["ecm"], vcs=g["vcs"], energy_err=None, vcs_err=g["vcs_err"], threshold=float(g["threshold"]))
    s.solve()
    model = tmp_path / "model.npz"
    np.savez(model, vcs=np.stack([g["ecm"], g["vcs"], 0 * g["ecm"], g["vcs_err"]]), bcs=g["bcs0"])
    done = threading.Event()
    out = {}

    def work():
        out["r"] = s.chi2_test_model(n=1 << 22, initial_ampl=1e4, model_path=str(model), model_visible_cs_name="vcs", model_born_cs_name="bcs")
        done.set()
    s.chi2_test_model(n=1024, initial_ampl=1e4, model_path=str(model), model_visible_cs_name="vcs", model_born_cs_name="bcs")   # warm-up
    th = threading.Thread(target=work)
    ticks = 0
    th.start()
    while not done.is_set():
        ticks += 1
    th.join()
    assert out["r"]["hist"].sum() == 1 << 22
    assert ticks > 1000, ticks
