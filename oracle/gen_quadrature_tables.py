#!/usr/bin/env python3
"""Generate the quadrature tables used by the oracle and by the CUDA kernels.

ORACLE / TEST INFRASTRUCTURE -- not product code.

Outputs (both committed, regenerate with `python oracle/gen_quadrature_tables.py`):
  oracle/gk_tables.h                      Gauss-Kronrod 21 and 61 point rules (the rules behind
                                          gsl_integration_qags / gsl_integration_qag key=6 that the
                                          reference calls at src/Integration.cpp:31,66)
  isrsolver_b200/csrc/gl_tables.cuh       Gauss-Legendre 16 and 32 point rules for the device panels
                                          and the 6 point Gauss-Hermite rule (src/Integration.cpp:92)

Everything is computed with mpmath at 80 digits: Gauss-Legendre nodes by Newton iteration on P_n,
Kronrod nodes as the roots of the Stieltjes polynomial E_{n+1} (orthogonal to all lower-degree
polynomials w.r.t. the sign-changing weight P_n), weights from the interpolatory moment equations.
The 21 point table is cross-checked against the copy scipy ships in scipy/integrate/_quad_vec.py.
"""
import os
import sys
import mpmath as mp

mp.mp.dps = 80
HERE = os.path.dirname(os.path.abspath(__file__))


def legendre_nodes(n):
    """Gauss-Legendre nodes (ascending) and weights on [-1,1]."""
    xs, ws = [], []
    for k in range(1, n + 1):
        x = mp.cos(mp.pi * (k - mp.mpf(1) / 4) / (n + mp.mpf(1) / 2))
        for _ in range(100):
            p0, p1 = mp.mpf(1), x
            for m in range(2, n + 1):
                p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
            dp = n * (x * p1 - p0) / (x * x - 1)
            dx = p1 / dp
            x -= dx
            if abs(dx) < mp.mpf(10) ** (-70):
                break
        p0, p1 = mp.mpf(1), x
        for m in range(2, n + 1):
            p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
        dp = n * (x * p1 - p0) / (x * x - 1)
        xs.append(x)
        ws.append(2 / ((1 - x * x) * dp * dp))
    order = sorted(range(n), key=lambda i: xs[i])
    return [xs[i] for i in order], [ws[i] for i in order]


def legendre_p(k, x):
    if k == 0:
        return mp.mpf(1)
    p0, p1 = mp.mpf(1), x
    for m in range(2, k + 1):
        p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
    return p1


def kronrod(n):
    """(2n+1)-point Kronrod extension of the n-point Gauss rule.
    Returns nodes (ascending), Kronrod weights, and Gauss weights (0 at the added nodes)."""
    gx, gw = legendre_nodes(n)
    # exact integration helper for polynomial integrands of degree <= 3n+2
    qx, qw = legendre_nodes(2 * n + 4)

    def integ(f):
        return sum(w * f(x) for x, w in zip(qx, qw))

    # E_{n+1} = P_{n+1} + sum_k c_k P_k, k same parity as n+1, k <= n-1 ... solve orthogonality to
    # P_n * P_m for m = 0..n (only m with parity of n+1+n, i.e. odd total, matter: m odd parity of 1)
    ks = [k for k in range(n + 1) if (k - (n + 1)) % 2 == 0]
    ms = [m for m in range(n + 1) if (m + n + (n + 1)) % 2 == 0]
    A = mp.matrix(len(ms), len(ks))
    rhs = mp.matrix(len(ms), 1)
    for r, m in enumerate(ms):
        for c, k in enumerate(ks):
            A[r, c] = integ(lambda x: legendre_p(n, x) * legendre_p(m, x) * legendre_p(k, x))
        rhs[r] = -integ(lambda x: legendre_p(n, x) * legendre_p(m, x) * legendre_p(n + 1, x))
    coef = mp.lu_solve(A, rhs)

    def E(x):
        return legendre_p(n + 1, x) + sum(coef[c] * legendre_p(k, x) for c, k in enumerate(ks))

    # the n+1 roots of E interlace with the Gauss nodes: one root in each gap of [-1, g_1, ..., g_n, 1]
    brk = [mp.mpf(-1)] + gx + [mp.mpf(1)]
    kx = []
    for lo, hi in zip(brk[:-1], brk[1:]):
        flo, fhi = E(lo), E(hi)
        assert flo * fhi < 0, "Stieltjes roots do not interlace"
        a, b = lo, hi
        for _ in range(400):
            mid = (a + b) / 2
            fm = E(mid)
            if fm == 0:
                a = b = mid
                break
            if flo * fm < 0:
                b, fhi = mid, fm
            else:
                a, flo = mid, fm
            if b - a < mp.mpf(10) ** (-72):
                break
        kx.append((a + b) / 2)
    nodes = sorted(gx + kx)
    m = len(nodes)
    # weights: sum_i w_i P_k(x_i) = 2 delta_k0, k = 0..m-1
    M = mp.matrix(m, m)
    r = mp.matrix(m, 1)
    for k in range(m):
        for i, x in enumerate(nodes):
            M[k, i] = legendre_p(k, x)
        r[k] = 2 if k == 0 else 0
    wk = mp.lu_solve(M, r)
    wg = []
    for x in nodes:
        hit = [w for g, w in zip(gx, gw) if abs(g - x) < mp.mpf(10) ** (-60)]
        wg.append(hit[0] if hit else mp.mpf(0))
    return nodes, [wk[i] for i in range(m)], wg


def hermite_nodes(n):
    """Physicists' Gauss-Hermite nodes/weights (weight exp(-t^2))."""
    # roots of H_n by Newton from asymptotic guesses via companion (use mp.polyroots on coefficients)
    coeffs = [mp.mpf(0)] * (n + 1)
    # H_n coefficients by recurrence H_{k+1} = 2x H_k - 2k H_{k-1}
    h0 = [mp.mpf(1)]
    h1 = [mp.mpf(0), mp.mpf(2)]
    for k in range(1, n):
        h2 = [mp.mpf(0)] + [2 * c for c in h1]
        for i, c in enumerate(h0):
            h2[i] -= 2 * k * c
        h0, h1 = h1, h2
    poly = list(reversed(h1))
    roots = sorted(mp.polyroots(poly, maxsteps=500, extraprec=400), key=lambda z: mp.re(z))
    xs = [mp.re(z) for z in roots]

    def H(k, x):
        a, b = mp.mpf(1), 2 * x
        if k == 0:
            return a
        for m in range(1, k):
            a, b = b, 2 * x * b - 2 * m * a
        return b

    ws = [mp.power(2, n - 1) * mp.factorial(n) * mp.sqrt(mp.pi) / (n * n * H(n - 1, x) ** 2) for x in xs]
    return xs, ws


def fmt(v):
    return mp.nstr(v, 25, min_fixed=-1, max_fixed=1) if v != 0 else "0.0"


def c_array(name, vals, ctype="static const double"):
    body = ",\n  ".join(fmt(v) for v in vals)
    return f"{ctype} {name}[{len(vals)}] = {{\n  {body}\n}};\n"


def main():
    # --- GK tables (QUADPACK layout: abscissae descending from the outermost, positive half only,
    #     last entry = centre 0; wg holds the Gauss weights of the embedded rule)
    out = ["// GENERATED by oracle/gen_quadrature_tables.py -- do not edit.\n",
           "// Gauss-Kronrod rules used by the QUADPACK restatement in oracle/isr_oracle.c.\n",
           "#ifndef ISR_ORACLE_GK_TABLES_H\n#define ISR_ORACLE_GK_TABLES_H\n\n"]
    for n in (10, 30):
        nodes, wk, wg = kronrod(n)
        m = len(nodes)            # 2n+1
        half = nodes[m // 2:]     # 0 ... largest
        halfk = wk[m // 2:]
        halfg = wg[m // 2:]
        xgk = list(reversed(half))     # n+1 entries, descending, last = 0
        wgk = list(reversed(halfk))
        wgg = list(reversed(halfg))
        out.append(c_array(f"xgk{2 * n + 1}", xgk))
        out.append(c_array(f"wgk{2 * n + 1}", wgk))
        out.append(c_array(f"wgf{2 * n + 1}", wgg))   # full-length Gauss weights (0 at Kronrod-only nodes)
        out.append("\n")
        assert abs(sum(wk) - 2) < mp.mpf(10) ** (-60)
    out.append("#endif\n")
    with open(os.path.join(HERE, "gk_tables.h"), "w") as f:
        f.write("".join(out))

    # --- device tables
    dev = ["// GENERATED by oracle/gen_quadrature_tables.py -- do not edit.\n",
           "// Gauss-Legendre panels (nodes ascending on [-1,1]) and the 6-point Gauss-Hermite rule\n",
           "// (reference: gsl_integration_fixed_hermite, n=6, src/Integration.cpp:92).\n",
           "#pragma once\n\n"]
    for n in (8, 16, 32):
        xs, ws = legendre_nodes(n)
        dev.append(c_array(f"kGL{n}_x", xs, "__device__ __constant__ double"))
        dev.append(c_array(f"kGL{n}_w", ws, "__device__ __constant__ double"))
    # 15-point Gauss-Kronrod rule (7-point Gauss embedded) for the forward-convolution service: the
    # Kronrod-minus-Gauss difference is reported back as the per-panel error estimate
    kx, kw, kg = kronrod(7)
    dev.append(c_array("kGK15_x", kx, "__device__ __constant__ double"))
    dev.append(c_array("kGK15_wk", kw, "__device__ __constant__ double"))
    dev.append(c_array("kGK15_wg", kg, "__device__ __constant__ double"))
    hx, hw = hermite_nodes(6)
    dev.append(c_array("kGH6_t", hx, "__device__ __constant__ double"))
    dev.append(c_array("kGH6_w_over_sqrtpi", [w / mp.sqrt(mp.pi) for w in hw], "__device__ __constant__ double"))
    with open(os.path.join(HERE, "..", "isrsolver_b200", "csrc", "gl_tables.cuh"), "w") as f:
        f.write("".join(dev))
    # oracle copy of GH6 (plain C)
    with open(os.path.join(HERE, "gh_tables.h"), "w") as f:
        f.write("// GENERATED by oracle/gen_quadrature_tables.py -- do not edit.\n")
        f.write("#ifndef ISR_ORACLE_GH_TABLES_H\n#define ISR_ORACLE_GH_TABLES_H\n")
        f.write(c_array("gh6_t", hx))
        f.write(c_array("gh6_w", hw))
        f.write("#endif\n")
    print("tables written")


if __name__ == "__main__":
    sys.exit(main())
