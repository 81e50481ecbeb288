"""Host logic that needs no GPU: the C-ABI library loads and exports what include/isr_b200.h declares,
error behaviour without a device, sharding/packing, and the world_size-2 gloo path of the one collective."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from isrsolver_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "isr_b200.h")).read()
    declared = set(re.findall(r"\b(isr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.load()                     # dlopen + bind every symbol; raises if one is missing
    for name in declared:
        assert hasattr(lib, name)
    out = subprocess.check_output(["nm", "-D", L.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T isr_" in l}
    assert declared <= exported
    # the product library must not link the oracle
    assert "orc_" not in out


def test_header_cites_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "isr_b200.h")).read()
    for cite in ("src/ISRSolverSLE.cpp:138-149", "src/Chi2Test.cpp:38-60", "src/RatioTest.cpp:38-45",
                 "src/ISRSolverTikhonov.cpp:139-155", "src/ISRSolverTSVD.cpp:53-74", "src/BaseISRSolver.cpp:232-258"):
        assert cite in hdr


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode")
def test_cpp_class_mirror_compiles_against_the_c_abi(tmp_path):
    """include/isr_b200.hpp (the C++ mirror of BaseISRSolver / ISRSolverSLE / ISRSolverTikhonov / ISRSolverTSVD, the host
    side a maintainer of the reference would keep) must compile as plain C++17 against include/isr_b200.h alone -- no
    CUDA headers, no torch -- and link against the shared library."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.cpp"
    src.write_text('#include "isr_b200.hpp"\nint main() { return isr_last_error() == nullptr; }\n')
    lib = os.path.join(root, "isrsolver_b200")
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-I", os.path.join(root, "include"), str(src), "-o", str(tmp_path / "t"),
                           "-L", lib, "-lisr_b200", "-Wl,-rpath," + lib, "-Wl,-rpath,/usr/local/cuda/lib64"])


def test_toy_kernel_chunk_schedules_on_the_cpu(tmp_path):
    """tests/cpp/test_sched.cu walks the toy kernel's own chunk schedules (SchedDense / SchedTri of
    csrc/toy_gemm.cuh, __host__ __device__) on the host: every (tile, pass, column block, chunk) exactly once and in
    order, triangular chunk counts, the pipeline-safety distance between pass 1 and pass 2 of a tile, and -- for every shipped
    tile shape -- that the triangular schedule skips exactly the column tiles and chunks that hold only structural zeros."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_sched")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                           "--extended-lambda", "-I", os.path.join(root, "isrsolver_b200", "csrc"), "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_sched.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "schedule walk ok" in out.stdout, out.stdout + out.stderr


def test_integrand_source_on_the_cpu(tmp_path, oracle):
    """tests/cpp/test_kf_host.cu compiles the product's own integrand headers (kf_device.cuh, panel_device.cuh:
    __host__ __device__) for the host and compares kf_plain / kf_value / the ROOT bin arithmetic with the oracle's
    restatement of the reference (and SURVEY 8c's known values) -- the only product arithmetic that can run without a GPU."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    orc = os.path.join(root, "oracle")
    exe = str(tmp_path / "test_kf_host")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                           "-I", os.path.join(root, "isrsolver_b200", "csrc"), "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_kf_host.cu"), "-o", exe, "-L", orc, "-lisr_oracle",
                           "-Xlinker", "-rpath=" + orc])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "integrand source ok" in out.stdout, out.stdout + out.stderr


def test_panel_quadrature_on_the_cpu(tmp_path, oracle):
    """tests/cpp/test_panels_host.cu: the product's own panel enumeration, singularity-removing substitutions and
    integrand (csrc/panel_device.cuh, __host__ __device__), compiled for the host, integrate the four moments of every
    (row, segment) item of a 14-point problem -- eps = 1 and a tabulated efficiency -- and must agree with the oracle's
    adaptive QUADPACK integration of the reference integrand to 1e-10 of the item's zeroth moment.  The assembly's
    numerical core checked against the reference algorithm without a GPU."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    orc = os.path.join(root, "oracle")
    exe = str(tmp_path / "test_panels_host")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                           "--extended-lambda", "-I", os.path.join(root, "isrsolver_b200", "csrc"), "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_panels_host.cu"), "-o", exe, "-L", orc, "-lisr_oracle",
                           "-Xlinker", "-rpath=" + orc])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "panel quadrature ok" in out.stdout, out.stdout + out.stderr


def test_whole_matrices_from_the_product_source_on_the_cpu(tmp_path):
    """tests/cpp/test_matrix_host.cu assembles whole matrices on the CPU from the product's own source -- range
    normalisation + basis tables (csrc/basis_host.hpp, the host code behind isr_problem_set_ranges) and the panel
    quadrature the kernels run (csrc/panel_device.cuh compiled for the host).  The golden fixtures' problems must pass
    the SAME gates as on the GPU (tests/test_gpu_assembly.py::test_matrix_vs_golden): 1e-9 against the converged oracle
    and against the reference's own compiled code; invalid ranges must be refused."""
    import shutil
    import subprocess
    from conftest import assert_matrix_parity, load_golden
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_matrix_host")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                           "--extended-lambda", "-I", os.path.join(root, "isrsolver_b200", "csrc"), "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_matrix_host.cu"), "-o", exe])

    def run(ecm, thr, settings, eff=None, sigma=None):
        inp = (f"{len(ecm)} {float(thr)!r} {len(settings)}\n" + " ".join(repr(float(e)) for e in ecm) + "\n" +
               "\n".join(" ".join(str(int(v)) for v in r) for r in settings) + "\n")
        if eff is None:
            inp += "0 0.0 1.0\n"
        else:
            vals, nx, xlo, xhi = eff
            inp += f"{int(nx)} {float(xlo)!r} {float(xhi)!r}\n" + " ".join(repr(float(v)) for v in np.asarray(vals).ravel()) + "\n"
        inp += "0\n" if sigma is None else "1\n" + " ".join(repr(float(v)) for v in sigma) + "\n"
        return subprocess.run([exe], input=inp, capture_output=True, text=True)

    def matrix(out):
        assert out.returncode == 0, out.stderr
        return np.array([[float(v) for v in line.split()] for line in out.stdout.strip().splitlines()])
    for name in ("c1_sle_n50_linear", "c2_sle_n40_cspline", "c5_mixed3_n36"):
        g = load_golden(name)
        G = matrix(run(g["ecm"], g["threshold"], g["settings"]))
        assert_matrix_parity(G, g["G_converged"], g["G_converged_err"], what=name + " (CPU build of the product source) vs converged oracle")
        assert_matrix_parity(G, g["G_ref"], 3 * g["G_faithful_err"], row_rel=1e-9, what=name + " vs the compiled reference")
        if "linear" in name:
            assert np.all(np.triu(G, 1) == 0)
    # tabulated efficiency (ISRSolverSLE2, 64 bins per energy point) and beam-energy spread: gates of
    # test_matrix_with_row_table_efficiency / test_matrix_with_energy_spread
    g = load_golden("c3_sle2_n24_eps64")
    G = matrix(run(g["ecm"], g["threshold"], g["settings"], eff=(g["eff_values"], g["eff_nx"], g["eff_xlo"], g["eff_xhi"])))
    assert_matrix_parity(G, g["G_converged"], g["G_converged_err"], what="eps table vs converged")
    assert_matrix_parity(G, g["G_ref"], 3 * g["G_faithful_err"], row_rel=1e-9, what="eps table vs the compiled reference")
    g = load_golden("c4_tikhonov_n32_spread")
    G = matrix(run(g["ecm"], g["threshold"], g["settings"], sigma=g["ecm_err"]))
    assert_matrix_parity(G, g["G_spread"], 32 * g["G_converged_err"].max(), what="spread matrix")
    assert_matrix_parity(G, g["S_ref"] @ g["G_ref"], 32 * 3 * g["G_faithful_err"].max(), row_rel=1e-9, what="spread matrix vs the compiled reference")
    g = load_golden("c1_sle_n50_linear")
    n = len(g["ecm"])
    assert run(g["ecm"], g["threshold"], [(0, 0, 10), (1, 12, n - 1)]).returncode == 3      # segment 11 missed -> ISR_ERANGE
    assert run(g["ecm"], g["threshold"], [(0, 0, 20), (1, 15, n - 1)]).returncode == 3      # overlap


def test_tikhonov_lcurve_arithmetic_on_the_cpu(tmp_path):
    """tests/cpp/test_lcurve_host.cu runs the product's per-lambda Tikhonov arithmetic (csrc/tikhonov_device.cuh, the body
    of lcurve_kernel) on the host.  With (mu, U) from SciPy's generalised eigen-solver for the golden C4-like problem it
    must reproduce the literal matrix evaluation of src/ISRSolverTikhonov.cpp:61-128 stored in the fixture: solution,
    residual norm, smoothness norm, L-curve curvature and its derivative."""
    import shutil
    import subprocess
    import scipy.linalg as sla
    from conftest import load_golden
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_lcurve_host")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I", os.path.join(root, "isrsolver_b200", "csrc"),
                           os.path.join(root, "tests", "cpp", "test_lcurve_host.cu"), "-o", exe])
    g = load_golden("c4_tikhonov_n32_spread")
    G, F, v, err = g["G_spread"], g["tik_F"], g["vcs"], g["vcs_err"]
    W = np.diag(err ** -2.0)
    mu, U = sla.eigh(G.T @ W @ G, F)
    c = U.T @ G.T @ W @ v
    lam = np.asarray(g["tik_lambda"], dtype=float)
    inp = (f"{len(mu)} {len(lam)}\n" + " ".join(repr(float(x)) for x in mu) + "\n" + " ".join(repr(float(x)) for x in c) + "\n" +
           " ".join(repr(float(x)) for x in lam) + "\n")
    out = subprocess.run([exe], input=inp, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = np.array([[float(t) for t in line.split()] for line in out.stdout.strip().splitlines()])
    for q in range(len(lam)):
        x, y, curv, dcurv, z = rows[q, 0], rows[q, 1], rows[q, 2], rows[q, 3], rows[q, 4:]
        np.testing.assert_allclose(U @ z, g["tik_bcs"][q], rtol=1e-6)
        assert y == pytest.approx(g["tik_y"][q], rel=1e-6)
        assert x == pytest.approx(g["tik_x"][q], rel=1e-5, abs=1e-9)
        assert curv == pytest.approx(g["tik_curv"][q], rel=1e-5)
        if "tik_dcurv" in g:
            assert dcurv == pytest.approx(g["tik_dcurv"][q], rel=1e-4)


def test_toy_tile_chooser_without_a_device():
    """isr_toy_tile_name is pure host logic: every N the tests and BASELINE configs use gets a tile shape that fits the
    227 KB of shared memory, the measured winners are the ones chosen (DESIGN K5), and impossible requests return NULL."""
    from isrsolver_b200 import _lib as L
    lib = L.load()
    assert lib.isr_toy_tile_name(200, 1) == b"m64n200w20s2k40"      # bench / C4 size: 5 full 40-double chunks, 20 warps
    assert lib.isr_toy_tile_name(100, 1) == b"m128n104w16s2k40"     # C2
    assert lib.isr_toy_tile_name(500, 1) == b"m64n256w16s2k40"      # C5 chi2 only
    assert lib.isr_toy_tile_name(500, 4) == b"m64n128w16s2k40"      # C5 with ratio statistics: smaller column block
    for n in range(1, 1001):
        for nstat in (0, 1, 4):
            assert lib.isr_toy_tile_name(n, nstat) is not None, (n, nstat)
    assert lib.isr_toy_tile_name(0, 1) is None and lib.isr_toy_tile_name(200, 3) is None
    assert lib.isr_toy_tile_name(3000, 4) is None                   # ratio statistics of 3000 points do not fit


def test_gauss_jordan_slot_protocol_on_the_cpu(tmp_path):
    """tests/cpp/test_gj_phases.cu replays the mbarrier protocol of the cluster Gauss-Jordan inverse (four CTAs, four
    slots: empty / lfull / rfull / done, expect_tx re-arming, credit forwarding) for every N from 1 to 240 with the
    kernel's own phase arithmetic (csrc/gj_phases.cuh): each wait must be passed the parity of the phase that has just
    completed for that use, and nothing may stay armed at the end."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_gj_phases")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I", os.path.join(root, "isrsolver_b200", "csrc"),
                           os.path.join(root, "tests", "cpp", "test_gj_phases.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "slot protocol ok" in out.stdout, out.stdout + out.stderr


def test_philox_known_answers_on_the_cpu(tmp_path):
    """The bit source of the device toy generator (csrc/rng_device.cuh, __host__ __device__) reproduces the
    Philox4x32-10 known-answer vectors published with Random123 -- run on the host, no GPU."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_philox_host")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I", os.path.join(root, "isrsolver_b200", "csrc"),
                           os.path.join(root, "tests", "cpp", "test_philox_host.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "known answers ok" in out.stdout, out.stdout + out.stderr


def test_no_cpu_fallback_without_device():
    from isrsolver_b200 import _lib as L
    lib = L.load()
    h = C.c_void_p()
    rc = lib.isr_ctx_create(0, C.byref(h))
    assert rc == L.ISR_ECUDA and not h
    assert b"no CPU fallback" in lib.isr_last_error()
    import isrsolver_b200 as isr
    with pytest.raises(isr.IsrGpuError):
        isr.ISRSolverSLE(3, [1.0, 1.1, 1.2], [1, 1, 1], None, [0.1, 0.1, 0.1], 0.8)


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "isrsolver_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "isr_oracle" not in src and "libisr_oracle" not in src, f


def test_efficiency_descriptor_validation():
    import isrsolver_b200 as isr
    with pytest.raises(isr.EfficiencyDimensionException):
        isr.Efficiency("table_2d", dimension=3)
    e = isr.Efficiency("row_table", values=np.ones((4, 10)), nx=8, xlo=0, xhi=1)
    d = e.desc(4)
    assert d.kind == 1 and d.nx == 8
    with pytest.raises(ValueError):
        e.desc(5)
    e2 = isr.Efficiency("table_2d", values=np.ones((5, 6)), xedges=[0, .1, .3, .6, 1.0], eedges=[1.0, 1.5, 1.8, 2.5])
    assert (e2.nx, e2.ne, e2.xhi, e2.elo) == (4, 3, 1.0, 1.0)
    e2.desc(7)


def test_shard_bounds_cover_exactly():
    from isrsolver_b200.dist import shard_bounds
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    from isrsolver_b200.dist import pack, shard_bounds, unpack
    rng = np.random.default_rng(0)
    total, world, n = 11, 3, 4
    chi2 = rng.random(total)
    ms = max(shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world))
    rows, sb, rs = [], np.zeros(n), np.zeros((n, 3))
    for r in range(world):
        b, e = shard_bounds(total, world, r)
        s, t = rng.random(n), rng.random((n, 3))
        sb += s; rs += t
        rows.append(pack(chi2[b:e], s, t, ms))
    c, s2, t2 = unpack(np.stack(rows), total, world, n)
    np.testing.assert_array_equal(c, chi2)
    np.testing.assert_allclose(s2, sb); np.testing.assert_allclose(t2, rs)


def test_th1_counts_matches_oracle(oracle):
    from isrsolver_b200.dist import th1_counts
    rng = np.random.default_rng(3)
    v = rng.chisquare(50, 5000)
    lo, hi = v.min(), v.max()
    ref, idx, *_ = oracle.th1_fill(v, 128, lo, hi)
    np.testing.assert_array_equal(th1_counts(v, 128, lo, hi), ref.astype(np.uint32))


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from isrsolver_b200.dist import all_gather_results, histogram_sharded, shard_bounds, th1_counts
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total, n = 1001, 7
rng = np.random.default_rng(42)
chi2_all = rng.chisquare(n, total); B = rng.random((total, n))
b, e = shard_bounds(total, world, rank)
stats = np.stack([np.full(n, e - b), B[b:e].sum(0), (B[b:e] ** 2).sum(0)], 1)
chi2, sb, rs = all_gather_results(chi2_all[b:e], B[b:e].sum(0), stats, total)
assert np.array_equal(chi2, chi2_all)
assert np.allclose(sb, B.sum(0), rtol=1e-13) and np.allclose(rs[:, 0], total) and np.allclose(rs[:, 2], (B ** 2).sum(0), rtol=1e-13)
h = th1_counts(chi2, 128, chi2.min(), chi2.max())
assert np.array_equal(h, th1_counts(chi2_all, 128, chi2_all.min(), chi2_all.max()))   # bit-identical with 1 process
# the histogram-only path: no chi2 movement, two O(N)-byte collectives, same counters bit for bit
cnt, lo, hi, sb2, rs2 = histogram_sharded(chi2_all[b:e], B[b:e].sum(0), stats)
assert (lo, hi) == (chi2_all.min(), chi2_all.max()) and np.array_equal(cnt, h.astype(np.int64))
assert np.allclose(sb2, B.sum(0), rtol=1e-13) and np.allclose(rs2, rs, rtol=1e-13)
# fewer values than ranks: rank 1's shard is empty -- the global range must not become (-inf, +inf)  (ADVICE r1)
one = np.array([3.25])
mine = one if rank == 0 else np.empty(0)
cnt1, lo1, hi1, _, _ = histogram_sharded(mine, np.zeros(n), np.zeros((n, 3)))
assert (lo1, hi1) == (3.25, 3.25) and cnt1.sum() == 1
# ... and with the values on the "device": only the local range is passed
cnt2, lo2, hi2, _, _ = histogram_sharded(np.empty(0), np.zeros(n), np.zeros((n, 3)),
                                         count_fn=lambda lo, hi: th1_counts(mine, 128, lo, hi),
                                         local_range=(3.25, 3.25) if rank == 0 else (np.inf, -np.inf))
assert (lo2, hi2) == (3.25, 3.25) and np.array_equal(cnt2, cnt1)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_gather(tmp_path):
    """The N>1 path on CPU: two gloo ranks shard 1001 toys, one all_gather, identical histogram."""
    w = tmp_path / "worker.py"
    w.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(w), ROOT],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
