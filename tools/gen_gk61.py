#!/usr/bin/env python3
"""Generate the 61-point Gauss-Kronrod rule (QUADPACK qk61 layout) with mpmath.

TEST/ORACLE INFRASTRUCTURE.  GSL is not installed in this image and is not part
of /root/reference, so the oracle's mini-GSL shim needs the qk61 abscissae and
weights that gsl_integration_qag(key=6) uses (reference call sites:
src/AU_cosmological_parameters.h:865 and :957).  They are *generated* here from
the mathematical definition instead of being typed from memory:

  * Gauss nodes  = roots of the Legendre polynomial P_30.
  * Kronrod nodes = roots of the Stieltjes polynomial E_31, defined by
        int_{-1}^{1} P_30(x) E_31(x) x^j dx = 0,   j = 0..30.
  * Kronrod weights from exactness on P_0..P_60, Gauss weights from the
    classical formula.

Output layout follows QUADPACK: xgk[0..30] descending from the outermost
Kronrod node to 0, odd entries are the Gauss nodes; wgk[0..30] the Kronrod
weights; wg[0..14] the weights of the 30-point Gauss rule for xgk[1],xgk[3],...

Checks performed before writing: sum of weights == 2, exactness on x^90,
interlacing, and the three spot values recorded in SURVEY.md section 9 (V5).

Usage: python tools/gen_gk61.py > oracle/gsl_shim/gk61_table.h   (and redtime_b200/csrc/gk61_table.h)
"""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 80
N = 30


def legendre(n, x):
    return mp.legendre(n, x)


def main():
    # --- Gauss-Legendre nodes of P_30 (positive half) -----------------------
    def dlegendre(n, x):
        return n * (x * legendre(n, x) - legendre(n - 1, x)) / (x * x - 1)

    def gauss_nodes(n):
        # double-precision start values from numpy, polished by Newton
        xs = []
        for x0 in np.polynomial.legendre.leggauss(n)[0]:
            x = mp.mpf(float(x0))
            for _ in range(12):
                x = x - legendre(n, x) / dlegendre(n, x)
            xs.append(x)
        xs = sorted(xs)
        assert min(b - a for a, b in zip(xs[:-1], xs[1:])) > mp.mpf(10) ** -4
        return xs

    gnodes = gauss_nodes(N)
    # Gauss weights: 2 / ((1-x^2) P_n'(x)^2)
    gweights = [2 / ((1 - x * x) * dlegendre(N, x) ** 2) for x in gnodes]

    # --- Stieltjes polynomial E_{N+1} = P_{N+1} + sum_{k<=N} c_k P_k --------
    # orthogonality against P_j, j=0..N, with weight P_N.  Use an exact
    # high-order Gauss rule for the triple products (degree <= 3N+1 = 91).
    qn, qw = [], []
    M = 60
    for x in gauss_nodes(M):
        qn.append(x)
        qw.append(2 / ((1 - x * x) * dlegendre(M, x) ** 2))
    Pq = [[legendre(k, x) for x in qn] for k in range(N + 2)]

    def triple(j, k):
        return mp.fsum(w * Pq[N][i] * Pq[j][i] * Pq[k][i]
                       for i, w in enumerate(qw))

    # parity: E_{N+1} has the parity of N+1 (odd) -> only odd k contribute.
    ks = [k for k in range(N + 1) if (k % 2) == ((N + 1) % 2)]
    A = mp.matrix(len(ks), len(ks))
    rhs = mp.matrix(len(ks), 1)
    for r, j in enumerate(ks):
        for c, k in enumerate(ks):
            A[r, c] = triple(j, k)
        rhs[r] = -triple(j, N + 1)
    coef = mp.lu_solve(A, rhs)

    def E(x):
        return legendre(N + 1, x) + mp.fsum(coef[c] * legendre(k, x)
                                            for c, k in enumerate(ks))

    # --- Kronrod nodes interlace with the Gauss nodes -----------------------
    brackets = [-mp.mpf(1)] + gnodes + [mp.mpf(1)]
    knodes = []
    for lo, hi in zip(brackets[:-1], brackets[1:]):
        a, b = lo, hi
        fa, fb = E(a), E(b)
        assert fa * fb < 0, "Stieltjes roots must interlace"
        x = mp.findroot(E, (a, b), solver="anderson")
        assert lo < x < hi
        knodes.append(x)
    allnodes = sorted(gnodes + knodes)
    assert len(allnodes) == 2 * N + 1

    # --- Kronrod weights: exact on P_0..P_{2N} -------------------------------
    n_all = len(allnodes)
    V = mp.matrix(n_all, n_all)
    b = mp.matrix(n_all, 1)
    for k in range(n_all):
        for i, x in enumerate(allnodes):
            V[k, i] = legendre(k, x)
        b[k] = 2 if k == 0 else 0
    kw = mp.lu_solve(V, b)
    kweights = [kw[i] for i in range(n_all)]

    # --- checks ---------------------------------------------------------------
    assert abs(mp.fsum(kweights) - 2) < mp.mpf(10) ** -60
    assert abs(mp.fsum(gweights) - 2) < mp.mpf(10) ** -60
    for deg in (88, 90):  # exact through degree 3N+1 = 91
        exact = mp.mpf(2) / (deg + 1)
        got = mp.fsum(w * x ** deg for w, x in zip(kweights, allnodes))
        assert abs(got - exact) < mp.mpf(10) ** -55, (deg, got - exact)
    assert all(w > 0 for w in kweights)

    # QUADPACK layout (descending, non-negative half)
    pos = [(x, w) for x, w in zip(allnodes, kweights) if x > -mp.mpf(10) ** -70]
    pos.sort(key=lambda t: -t[0])
    assert len(pos) == N + 1
    xgk = [abs(x) if abs(x) > mp.mpf(10) ** -60 else mp.mpf(0) for x, _ in pos]
    wgk = [w for _, w in pos]
    gpos = sorted([(x, w) for x, w in zip(gnodes, gweights) if x > 0],
                  key=lambda t: -t[0])
    wg = [w for _, w in gpos]
    for j, (x, _) in enumerate(gpos):
        assert abs(xgk[2 * j + 1] - x) < mp.mpf(10) ** -60
    # spot values
    assert abs(xgk[0] - mp.mpf("0.99948441005049064")) < 1e-16
    assert abs(wgk[0] - mp.mpf("0.0013890136986770")) < 1e-15
    assert abs(wgk[30] - mp.mpf("0.05149472942945157")) < 1e-16

    out = sys.stdout
    out.write("/* GENERATED by tools/gen_gk61.py (mpmath, 80 digits) -- do not edit.\n"
              " * 61-point Gauss-Kronrod rule in QUADPACK qk61 layout. */\n")
    def arr(name, vals):
        out.write("static const double %s[%d] = {\n" % (name, len(vals)))
        for v in vals:
            out.write("  %s,\n" % mp.nstr(v, 25, strip_zeros=False))
        out.write("};\n")
    arr("gk61_xgk", xgk)
    arr("gk61_wgk", wgk)
    arr("gk61_wg", wg)


if __name__ == "__main__":
    main()
