#!/usr/bin/env python
"""Generate the Rys root/weight interpolation tables used by the sm_100a kernels.

Written from the published definition of the Rys quadrature (Dupuis, Rys, King
1976; Golub-Welsch 1969), NOT copied from the reference's
lib/gvhf-rys/rys_roots_dat.cu.  The reference evaluates the same quantities with
14-term Chebyshev series on 40 intervals of width 2.5 (rys_roots.cu:25-88); we
keep that interval scheme (it is what bounds the interpolation error at ~1e-14)
but the numbers, the memory layout and the large-x constants are regenerated
here with mpmath at 120 digits.

Measure:  d mu(u) = exp(-x u) / (2 sqrt(u)) du  on u = t^2 in [0,1];  moments
m_k = F_k(x) (Boys function).  Nodes u_i and weights W_i of the n-point Gauss
rule:  sum_i W_i u_i^k = F_k(x), k < 2n.

Layout written to rys_tables.h (double):
  RYS_CHEB[off(n) + ((it*2n + idx)*NCHEB + k)],  off(n) = NINT*NCHEB*n*(n-1)
      idx < n : root u_idx ;  idx >= n : weight W_{idx-n};  k-th Chebyshev coeff
      on interval it: x in [2.5 it, 2.5 (it+1)),  s = (x - 2.5 it)*0.8 - 1
  RYS_LARGE[off2(n) + idx], off2(n) = n*(n-1): idx<n: r_i (u_i -> r_i/x),
      idx>=n: w_i (W_i -> w_i/sqrt(x)); valid for x >= RYS_XLARGE(n) = 35 + 5 n
      (capped by the table end at x = 100).
"""
import sys
import os
import math
from multiprocessing import Pool
import mpmath as mp

NINT = 40
WIDTH = 2.5
NCHEB = 14
NMAX = int(os.environ.get('RYS_NMAX', '9'))
mp.mp.dps = 120


def boys(k, x):
    return mp.hyp1f1(k + mp.mpf(1) / 2, k + mp.mpf(3) / 2, -x) / (2 * k + 1)


def rys_rw(n, x):
    """n-point Gauss rule for the Rys measure at x; returns (roots u_i, weights W_i)."""
    x = mp.mpf(x)
    m = [boys(k, x) for k in range(2 * n + 1)]
    M = mp.matrix(n + 1, n + 1)
    for i in range(n + 1):
        for j in range(n + 1):
            M[i, j] = m[i + j]
    L = mp.cholesky(M)        # lower, M = L L^T ; R = L^T
    R = L.T
    alpha = []
    beta = []
    for j in range(n):
        a = R[j, j + 1] / R[j, j]
        if j > 0:
            a -= R[j - 1, j] / R[j - 1, j - 1]
        alpha.append(a)
        if j < n - 1:
            beta.append(R[j + 1, j + 1] / R[j, j])
    J = mp.matrix(n, n)
    for j in range(n):
        J[j, j] = alpha[j]
        if j < n - 1:
            J[j, j + 1] = beta[j]
            J[j + 1, j] = beta[j]
    E, V = mp.eigsy(J)
    order = sorted(range(n), key=lambda i: E[i])
    roots = [E[i] for i in order]
    weights = [m[0] * V[0, i] ** 2 for i in order]
    return roots, weights


def large_x(n):
    """Asymptotic constants: for x -> inf the measure becomes
    exp(-x u)/(2 sqrt u) on [0, inf): with u = v/x, nodes r_i/x and weights
    w_i/sqrt(x) where (r_i, w_i) is the Gauss rule for exp(-v)/(2 sqrt v) on
    [0,inf), moments Gamma(k+1/2)/2."""
    m = [mp.gamma(k + mp.mpf(1) / 2) / 2 for k in range(2 * n + 1)]
    M = mp.matrix(n + 1, n + 1)
    for i in range(n + 1):
        for j in range(n + 1):
            M[i, j] = m[i + j]
    R = mp.cholesky(M).T
    J = mp.matrix(n, n)
    for j in range(n):
        a = R[j, j + 1] / R[j, j]
        if j > 0:
            a -= R[j - 1, j] / R[j - 1, j - 1]
        J[j, j] = a
        if j < n - 1:
            b = R[j + 1, j + 1] / R[j, j]
            J[j, j + 1] = b
            J[j + 1, j] = b
    E, V = mp.eigsy(J)
    order = sorted(range(n), key=lambda i: E[i])
    return [E[i] for i in order], [m[0] * V[0, i] ** 2 for i in order]


def cheb_fit(vals):
    """vals[j] at nodes s_j = cos(pi (j+1/2)/N) -> N Chebyshev coefficients
    with f(s) = sum_k c_k T_k(s) (c_0 already halved)."""
    N = len(vals)
    c = []
    for k in range(N):
        s = mp.mpf(0)
        for j in range(N):
            s += vals[j] * mp.cos(mp.pi * k * (j + mp.mpf(1) / 2) / N)
        s *= mp.mpf(2) / N
        if k == 0:
            s /= 2
        c.append(s)
    return c


def cheb_eval(c, s):
    """double-precision Clenshaw, same recurrence as the device code"""
    b1 = b2 = 0.0
    for k in range(len(c) - 1, 0, -1):
        b1, b2 = 2 * s * b1 - b2 + c[k], b1
    return s * b1 - b2 + c[0]


def work(args):
    n, it = args
    x0 = mp.mpf(it) * WIDTH
    nodes = [x0 + WIDTH / 2 * (1 + mp.cos(mp.pi * (j + mp.mpf(1) / 2) / NCHEB)) for j in range(NCHEB)]
    vals = [rys_rw(n, xx) for xx in nodes]
    out = []
    for idx in range(2 * n):
        f = [(v[0][idx] if idx < n else v[1][idx - n]) for v in vals]
        out.append([float(ck) for ck in cheb_fit(f)])
    # interpolation error at 3 off-node points
    err = 0.0
    for frac in (0.07, 0.5, 0.93):
        xx = x0 + WIDTH * frac
        r, w = rys_rw(n, xx)
        s = float(frac * 2 - 1)
        for idx in range(2 * n):
            ref = r[idx] if idx < n else w[idx - n]
            got = cheb_eval(out[idx], s)
            err = max(err, abs(float((got - ref) / ref)))
    return n, it, out, err


def main(dst):
    tasks = [(n, it) for n in range(NMAX, 0, -1) for it in range(NINT)]
    with Pool(8) as pool:
        res = pool.map(work, tasks, chunksize=4)
    tab = {}
    maxerr = {}
    for n, it, out, err in res:
        tab[(n, it)] = out
        maxerr[n] = max(maxerr.get(n, 0), err)
    large = {n: large_x(n) for n in range(1, NMAX + 1)}
    # asymptotic error at the switch point
    lerr = {}
    for n in range(1, NMAX + 1):
        xs = min(35 + 5 * n, NINT * WIDTH)
        r, w = rys_rw(n, xs)
        e = 0.0
        for i in range(n):
            e = max(e, abs(float((large[n][0][i] / xs - r[i]) / r[i])),
                    abs(float((large[n][1][i] / mp.sqrt(xs) - w[i]) / w[i])))
        lerr[n] = e
    with open(dst, 'w') as f:
        f.write('// GENERATED by tools/gen_rys_tables.py (mpmath, %d digits) -- do not edit.\n' % mp.mp.dps)
        f.write('// Rys quadrature roots/weights: Chebyshev tables on %d intervals of width %.1f,\n' % (NINT, WIDTH))
        f.write('// %d coefficients each; layout documented in the generator.\n' % NCHEB)
        for n in range(1, NMAX + 1):
            f.write('// nroots=%d: max rel interpolation error %.2e, asymptotic error at switch %.2e\n'
                    % (n, maxerr[n], lerr[n]))
        f.write('#pragma once\n#define RYS_NMAX %d\n#define RYS_NINT %d\n#define RYS_NCHEB %d\n' % (NMAX, NINT, NCHEB))
        total = NINT * NCHEB * NMAX * (NMAX + 1)
        f.write('#define RYS_CHEB_SIZE %d\n#define RYS_LARGE_SIZE %d\n' % (total, NMAX * (NMAX + 1)))
        f.write('static const double RYS_LARGE_HOST[RYS_LARGE_SIZE] = {\n')
        for n in range(1, NMAX + 1):
            r, w = large[n]
            f.write(''.join('%.17e,\n' % float(v) for v in list(r) + list(w)))
        f.write('};\nstatic const double RYS_CHEB_HOST[RYS_CHEB_SIZE] = {\n')
        for n in range(1, NMAX + 1):
            for it in range(NINT):
                for idx in range(2 * n):
                    f.write(','.join('%.17e' % v for v in tab[(n, it)][idx]) + ',\n')
        f.write('};\n')
    print('max interpolation errors', maxerr)
    print('asymptotic errors', lerr)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'rys_tables.h')
