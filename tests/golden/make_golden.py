#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/.

The reference ships no golden vectors.  Its assembly-path SOURCES are compiled here into
oracle/_ref/libisr_ref.so (oracle/Makefile; GSL/Eigen resolved by oracle/ref_shim), and the fixtures
carry that library's outputs under *_ref keys:
    G_ref      evalEqMatrix fill loop over the reference's own Interpolator::evalKuraevFadinBasisIntegral
    S_ref      _energySpreadMatrix over the reference's gaussian_conv + Interpolator::basisEval
    vcs_ref    the reference's convolutionKuraevFadin of the Breit-Wigner Born model
    w_ref, D_ref   Interpolator::evalIntegralBasis / basisDerivEval
Everything else (solve, covariance, toys, Tikhonov, TSVD) is output of the restatement in oracle/
(NumPy/LAPACK in place of Eigen); since round 2 the reference's own solver classes run here as well, and their outputs on
these fixtures' inputs live in refrun_*.npz (make_refrun_golden.py), which tests/test_refrun_pin.py compares with these numbers.  This script
asserts that the restatement's faithful back-end reproduces the compiled reference BIT FOR BIT before
writing anything.  Run in the container that has /root/reference:
    python tests/golden/make_golden.py   (about a minute on 8 cores).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import isr_oracle as O  # noqa: E402
from oracle import isr_ref as R     # noqa: E402


def case(name, n, settings, eff=None, ecm_err=None, toys=64, tik=False):
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    I = O.Interp(ecm, O.E_THR, settings)
    Gf, Ef, Rf = I.eval_eq_matrix(eff=eff, with_errors=True)
    Gc, Ec, Rc = I.eval_eq_matrix(eff=eff, mode="converged", with_errors=True)
    out = dict(ecm=ecm, threshold=O.E_THR, settings=np.array(settings, dtype=np.int32),
               G_faithful=Gf, G_faithful_err=Ef, G_faithful_relerr=Rf, G_converged=Gc, G_converged_err=Ec)
    # ---- the reference's own compiled code on the same inputs
    v2 = eff is not None and eff.kind == "row_table"
    RI = R.RefInterp(ecm, O.E_THR, settings, v2=v2)
    G_ref = RI.eval_eq_matrix(eff=eff)
    assert np.array_equal(G_ref, Gf), name + ": faithful restatement differs from the compiled reference"
    w_ref = RI.integral_basis()
    D_ref = np.array([[RI.basis_deriv_eval(j, ecm[i]) for j in range(n)] for i in range(n)])
    born = lambda e: float(O.bw_born(e)[0])
    vcs_ref = np.array([R.convolution_kf(e, born, 0.0, 1.0 - (O.E_THR / e) ** 2,
                                         eff if (eff is not None and not v2) else None) for e in ecm]) if not v2 else None
    out.update(G_ref=G_ref, w_ref=w_ref, D_ref=D_ref)
    G = Gc
    if ecm_err is not None:
        S_ref = RI.energy_spread_matrix(ecm_err)
        out.update(S_ref=S_ref)
        S = I.energy_spread_matrix(ecm_err)
        G = S @ Gc
        out.update(ecm_err=ecm_err, S=S, G_spread=G)
    bcs0 = O.bw_born(ecm)
    vcs = O.bw_visible(ecm, eff=eff)
    if vcs_ref is not None:
        assert np.array_equal(vcs_ref, vcs), name + ": visible cross section differs from the compiled reference"
        out.update(vcs_ref=vcs_ref)
    vcs_err = 0.05 * vcs
    b, cov = O.sle_solve(G, vcs, vcs_err)
    V = O.draw_toys(vcs, vcs_err, toys)
    chi2_loop, B_loop = O.chi2_test_loop(G, V, bcs0, vcs_err)
    chi2_fair, B_fair = O.chi2_test_fair(G, V, bcs0, vcs_err)
    means, errs, counts, stats = O.ratio_test(B_fair, bcs0)
    hist, idx, lo, hi = O.chi2_histogram(chi2_fair)
    out.update(bcs0=bcs0, vcs=vcs, vcs_err=vcs_err, bcs=b, cov=cov, toys=V, chi2_loop=chi2_loop, chi2=chi2_fair,
               B=B_fair, ratio_means=means, ratio_mean_errors=errs, ratio_stats=stats,
               ratio_counts=counts.astype(np.uint16), chi2_hist=hist, chi2_lo=lo, chi2_hi=hi,
               cond=O.cond_number(G), w=I.integral_basis(), D=I.deriv_projector())
    if eff is not None:
        out.update(eff_values=eff.values, eff_nx=eff.nx, eff_xlo=eff.xlo, eff_xhi=eff.xhi)
    if tik:
        lam = np.geomspace(1e-9, 1.0, 12)
        F = O.tikhonov_F(out["D"], out["w"], True)
        rows = [O.tikhonov_lcurve(G, vcs, vcs_err, F, l) for l in lam]
        out.update(tik_lambda=lam, tik_F=F, tik_x=np.array([r["x"] for r in rows]), tik_y=np.array([r["y"] for r in rows]),
                   tik_curv=np.array([r["curv"] for r in rows]), tik_dcurv=np.array([r["dcurv"] for r in rows]),
                   tik_bcs=np.array([r["bcs"] for r in rows]), tik_cov3=rows[3]["cov"])
        k = n // 2
        out.update(tsvd_k=k, tsvd_bcs=O.tsvd_solve(G, vcs, vcs_err, k)[0], tsvd_bcs_keepone=O.tsvd_solve(G, vcs, vcs_err, k, True)[0],
                   sv=np.linalg.svd(G, compute_uv=False))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "done; faithful-vs-converged max abs", np.abs(Gf - Gc).max(), "loosest faithful epsrel", Rf.max())


if __name__ == "__main__":
    n = 50
    case("c1_sle_n50_linear", n, [(0, 0, n - 1)])                                     # BASELINE.json configs[0]
    n = 40
    case("c2_sle_n40_cspline", n, [(0, 0, 0), (1, 1, n - 1)], tik=False)              # configs[1] at test size
    n = 24
    ecm = np.linspace(O.E_MIN, O.E_MAX, n)
    case("c3_sle2_n24_eps64", n, [(0, 0, n - 1)], eff=O.row_table_efficiency(ecm, 64))  # configs[2] at test size
    n = 32
    case("c4_tikhonov_n32_spread", n, [(0, 0, 0), (1, 1, n - 1)], ecm_err=np.full(n, 0.002), tik=True)  # configs[3]
    n = 36
    case("c5_mixed3_n36", n, [(1, 0, 11), (0, 12, 20), (1, 21, n - 1)])
