#!/usr/bin/env python
"""Coefficient tables of bfm::bessel_k01 (blacklight_b200/csrc/bf_math.cuh), from 40-digit arithmetic (mpmath):
  * Chebyshev coefficients (degree 22) of K_nu(x) e^x sqrt(x) in u = 4/x - 1 on x >= 2, nu = 0, 1;
  * coefficients in q = x^2/4 of the ascending series of K_0, K_1 for x <= 2 (A&S 9.6.11, 9.6.13).
Prints C initialisers; run once, paste into bf_math.cuh."""
import mpmath as mp

mp.mp.dps = 40


def cheb(nu, deg=22):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    f = lambda u: mp.besselk(nu, 4 / (u + 1)) * mp.e ** (4 / (u + 1)) * mp.sqrt(4 / (u + 1))
    fv = [f(u) for u in nodes]
    c = [mp.fsum(fv[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) * 2 / n for j in range(n)]
    c[0] /= 2
    return [float(v) for v in c]


def series(n=14):
    hk = [mp.mpf(0)]
    for k in range(1, n):
        hk.append(hk[-1] + mp.mpf(1) / k)
    fac = mp.factorial
    return {'SER_I0': [1 / fac(k) ** 2 for k in range(n)], 'SER_S0': [hk[k] / fac(k) ** 2 for k in range(n)],
            'SER_I1': [1 / (fac(k) * fac(k + 1)) for k in range(n)],
            'SER_S1': [(2 * (hk[k] - mp.euler) + mp.mpf(1) / (k + 1)) / (fac(k) * fac(k + 1)) for k in range(n)]}


if __name__ == '__main__':
    for nu in (0, 1):
        print('CHEB_K%d = {%s};' % (nu, ', '.join('%.17e' % v for v in cheb(nu))))
    for name, arr in series().items():
        print('%s = {%s};' % (name, ', '.join('%.17e' % float(v) for v in arr)))
    print('EULER = %.17e' % float(mp.euler))
