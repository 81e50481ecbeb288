"""High-precision evaluations of the reference's Tikhonov / chi2 formulas -- test infrastructure, the yardstick for the
1e-8 parity gates where the reference's own double-precision route (two FullPivLU of T and L, an explicit inverse of an
ill-conditioned covariance: src/ISRSolverTikhonov.cpp:61-72, 139-155; src/Chi2Test.cpp:37, 51-55) carries cond * eps noise.

  mp_tikhonov   the literal matrix formulas of src/ISRSolverTikhonov.cpp:61-128, 139-155 and the per-toy chi2 of
                src/Chi2Test.cpp:51-55 in 50-digit mpmath arithmetic (small N: seconds at N = 32);
  ld_*          Gaussian elimination with partial pivoting in 80-bit long double (eps 1.1e-19) for the full sizes
                (N = 200), where mpmath would take minutes; chi2 is evaluated as |W^(-1/2) G^-T T d|^2, which equals
                d^T Cov^-1 d with Cov = A+ W^-1 A+^T, A+ = T^-1 G^T W exactly but never inverts the covariance."""
import numpy as np


def mp_tikhonov(G, vcs, err, F, lam, V=None, b0=None, dps=50):
    import mpmath as mp
    old = mp.mp.dps
    mp.mp.dps = dps
    try:
        n = len(vcs)
        f = lambda M: np.array(M.tolist(), dtype=float).reshape(M.rows, M.cols)
        mG = mp.matrix(np.asarray(G).tolist())
        mF = mp.matrix(np.asarray(F).tolist())
        mW = mp.diag([mp.mpf(float(e)) ** -2 for e in err])
        mv = mp.matrix(np.asarray(vcs).tolist())
        lam = mp.mpf(float(lam))
        mCovInv0 = mG.T * mW * mG
        mT = mCovInv0 + lam * mF                                   # :149-150
        mL = mp.inverse(mF) * mCovInv0 + lam * mp.eye(n)           # :151-152
        mAp = mp.inverse(mT) * mG.T * mW                           # :69
        mb = mAp * mv                                              # :70
        mCov = mAp * mp.inverse(mW) * mAp.T                        # :71
        dv = mG * mb - mv
        x, y = (dv.T * mW * dv)[0], (mb.T * mF * mb)[0]            # :106-113
        ds = -mp.lu_solve(mL, mb)
        d2s = -2 * mp.lu_solve(mL, ds)
        dksi = 2 * (mb.T * mF * ds)[0]
        d2ksi = 2 * (ds.T * mF * ds)[0] + 2 * (mb.T * mF * d2s)[0]
        curv = -abs(1 / dksi / (1 + lam * lam) ** mp.mpf("1.5"))   # :115-119
        dcurv = -d2ksi / dksi ** 2 * (1 + lam * lam) ** mp.mpf("-1.5") - 3 * lam / dksi * (1 + lam * lam) ** mp.mpf("-2.5")
        out = dict(x=float(x), y=float(y), curv=float(curv), dcurv=float(dcurv), bcs=f(mb).ravel(), cov=f(mCov))
        if V is not None:
            mCovInv = mp.inverse(mCov)                             # src/Chi2Test.cpp:37 (FullPivLU of Cov)
            mb0 = mp.matrix(np.asarray(b0).tolist())
            chi2, B = [], []
            for v in np.asarray(V):
                bk = mAp * mp.matrix(v.tolist())
                d = bk - mb0
                chi2.append(float((d.T * mCovInv * d)[0]))        # :51-55
                B.append(f(bk).ravel())
            out["chi2"], out["B"] = np.array(chi2), np.array(B)
        return out
    finally:
        mp.mp.dps = old


def ld_solve(A, B):
    """A^-1 B by Gaussian elimination with partial pivoting in long double."""
    A = np.array(A, dtype=np.longdouble)
    B = np.array(B, dtype=np.longdouble)
    one_d = B.ndim == 1
    if one_d:
        B = B[:, None]
    n = len(A)
    for k in range(n):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        if p != k:
            A[[k, p]] = A[[p, k]]
            B[[k, p]] = B[[p, k]]
        m = A[k + 1:, k] / A[k, k]
        A[k + 1:, k:] -= m[:, None] * A[k, k:][None, :]
        B[k + 1:] -= m[:, None] * B[k][None, :]
    X = np.empty_like(B)
    for k in range(n - 1, -1, -1):
        X[k] = (B[k] - A[k, k + 1:] @ X[k + 1:]) / A[k, k]
    return X[:, 0] if one_d else X


def ld_tikhonov_chi2(G, err, F, lam, V, b0):
    """(b_k - b0)^T Cov_lambda^-1 (b_k - b0) for every toy row of V, b_k = A+ v_k, in long double (see module docstring)."""
    ld = np.longdouble
    G = np.asarray(G, dtype=ld); F = np.asarray(F, dtype=ld); err = np.asarray(err, dtype=ld)
    W = err ** -2
    T = G.T @ (W[:, None] * G) + ld(lam) * F
    B = ld_solve(T, G.T @ (W[:, None] * np.asarray(V, dtype=ld).T)).T           # b_k = T^-1 G^T W v_k
    D = B - np.asarray(b0, dtype=ld)[None, :]
    Y = ld_solve(G.T, T @ D.T)                                                 # G^-T T d
    chi2 = ((Y * err[:, None]) ** 2).sum(0)                                    # |W^(-1/2) .|^2
    return chi2.astype(np.float64), B.astype(np.float64)
