"""The header-only C++ class mirror (include/isr_b200.hpp) compiled with g++ against libisr_b200.so."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _build(tmp_path, name="test_classes"):
    exe = tmp_path / name
    lib_dir = os.path.join(ROOT, "isrsolver_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-L", lib_dir, "-lisr_b200",
                           f"-Wl,-rpath,{lib_dir}", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", str(exe)])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    """No GPU needed: the class mirror compiles against the header and links against the library."""
    assert _build(tmp_path).exists()
    assert _build(tmp_path, "test_group").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("ngpu", [1, 2, 4])
def test_cpp_chi2_test_model_on_several_gpus(tmp_path, ngpu):
    """a C++ chi2TestModel caller spreads the test over the GPUs of the box from one process; results
    bit-identical to one GPU.  (ngpu = 1 runs the group machinery -- threads, communicator of one -- on any box.)"""
    import torch
    import isrsolver_b200 as isr
    if torch.cuda.device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    g = load_golden("c2_sle_n40_cspline")
    exe = _build(tmp_path, "test_group")
    for k in ("ecm", "vcs", "vcs_err", "bcs0"):
        np.ascontiguousarray(g[k], dtype=np.float64).tofile(tmp_path / f"{k}.bin")
    out = subprocess.run([str(exe), str(tmp_path), str(ngpu)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "group ok" in out.stdout and "group_hist_sum 30011" in out.stdout
    val = lambda key: [l for l in out.stdout.splitlines() if l.startswith(key + " ")][0].split()[1:]
    assert float(val("callable_half")[0]) < 1e-15
    n = len(g["ecm"])
    Gc = np.fromfile(tmp_path / "out_G_callable.bin").reshape(n, n).T
    m = isr.ISRSolverSLE(n, g["ecm"], g["vcs"], None, g["vcs_err"], 0.827,
                         efficiency=lambda x, e: 0.3 + 0.5 * np.exp(-30.0 * x) * (1.0 - 0.1 * (e - 1.5)))
    m.set_interp_settings([(False, 0, 0), (True, 1, n - 1)])
    m.eval_equation_matrix()
    np.testing.assert_allclose(Gc, m.intop_matrix(), rtol=1e-13, atol=1e-16)


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path, oracle):
    g = load_golden("c2_sle_n40_cspline")
    exe = _build(tmp_path)
    for k in ("ecm", "vcs", "vcs_err", "toys", "bcs0"):
        np.ascontiguousarray(g[k], dtype=np.float64).tofile(tmp_path / f"{k}.bin")
    out = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp ok" in out.stdout
    n = len(g["ecm"])
    rd = lambda name, shape=None: np.fromfile(tmp_path / name).reshape(shape) if shape else np.fromfile(tmp_path / name)
    G = rd("out_G.bin", (n, n)).T                      # column-major on disk
    assert np.all(np.abs(G - g["G_converged"]) <= 1e-9 * np.abs(g["G_converged"]) + 2e-12 + g["G_converged_err"])
    np.testing.assert_allclose(rd("out_bcs.bin"), g["bcs"], rtol=1e-8)
    cov = rd("out_cov.bin", (n, n)).T
    assert np.abs(cov - g["cov"]).max() <= 1e-8 * np.abs(g["cov"]).max()
    np.testing.assert_allclose(rd("out_chi2.bin"), g["chi2"], rtol=1e-7)
    np.testing.assert_allclose(rd("out_means.bin"), g["ratio_means"], rtol=1e-6, atol=1e-9)
    # the same against what the reference's own compiled classes returned (tests/golden/make_refrun_golden.py)
    r = load_golden("refrun_c2_sle_n40_cspline")
    assert np.abs(G - r["ref_G_native"]).max() <= 1e-9 * np.abs(r["ref_G_native"]).max()
    assert np.abs(rd("out_bcs.bin") - r["ref_bcs_native"]).max() <= 1e-8 * np.abs(r["ref_bcs_native"]).max()
    assert np.abs(cov - r["ref_cov_native"]).max() <= 1e-8 * np.abs(r["ref_cov_native"]).max()
    np.testing.assert_allclose(rd("out_chi2.bin"), r["ref_chi2"], rtol=1e-7)
    np.testing.assert_allclose(rd("out_means.bin"), r["ref_ratio_means"], rtol=1e-6, atol=1e-9)
    I = oracle.Interp(g["ecm"], 0.827)      # the Tikhonov/TSVD solvers in the C++ test keep the default linear basis
    Gl = I.eval_eq_matrix(mode="converged")
    F = oracle.tikhonov_F(I.deriv_projector(), I.integral_basis(), True)
    ref = oracle.tikhonov_lcurve(Gl, g["vcs"], g["vcs_err"], F, 1e-3)
    np.testing.assert_allclose(rd("out_tik_bcs.bin"), ref["bcs"], rtol=1e-7)
    line = [l for l in out.stdout.splitlines() if l.startswith("tik")][0].split()
    assert float(line[4]) == pytest.approx(ref["y"], rel=1e-7) and float(line[6]) == pytest.approx(ref["curv"], rel=1e-7)
    np.testing.assert_allclose(rd("out_tsvd_bcs.bin"), oracle.tsvd_solve(Gl, g["vcs"], g["vcs_err"], n // 2)[0], rtol=1e-7)
    assert f"hist_sum {len(g['toys'])}" in out.stdout
    # free functions, SLE2 and the scans added on top of the class mirror
    val = lambda key: [l for l in out.stdout.splitlines() if l.startswith(key + " ")][0].split()[1:]
    assert float(val("kernel")[0]) == pytest.approx(5.283686571158812, rel=1e-13)
    np.testing.assert_allclose(rd("out_vis.bin"), g["vcs_ref"], rtol=1e-8)       # the compiled reference's convolution
    assert float(val("conv_scalar")[0]) == pytest.approx(g["vcs_ref"][3], rel=1e-8)
    c0, c1 = map(float, val("condscan"))
    assert c0 == pytest.approx(np.linalg.cond(g["G_converged"]), rel=1e-8) and c1 > c0
    assert float(val("chi2model_mean")[0]) == pytest.approx(n, rel=0.15)
    assert int(val("chi2data_n")[0]) == 256
    assert float(val("sle2_half")[0]) < 1e-15
    assert int(val("toyscan")[0]) == 16 and float(val("lcurve_lambda")[0]) > 0
