#!/usr/bin/env python3
"""Near-minimax coefficients for the range-specialised sincospi of csrc/ps_kernel.cuh.

   sin(pi r) = r*pi + r*z*S1(z),   cos(pi r) = 1 + z*C1(z),   z = r*r,  |r| <= 1/4

S1, C1 are fitted on z in [0, 1/16] by Chebyshev interpolation in 60-digit arithmetic (mpmath),
rounded to fp64, and the fp64 evaluation scheme is checked against mpmath on a dense grid."""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def S1(z):
    z = mp.mpf(z)
    if z == 0:
        return -mp.pi**3 / 6
    r = mp.sqrt(z)
    return (mp.sin(mp.pi * r) / r - mp.pi) / z


def C1(z):
    z = mp.mpf(z)
    if z == 0:
        return -mp.pi**2 / 2
    return (mp.cos(mp.pi * mp.sqrt(z)) - 1) / z


def fit(f, n):
    coeffs, err = mp.chebyfit(f, [0, mp.mpf(1) / 16], n, error=True)
    return [float(c) for c in coeffs[::-1]], float(err)  # lowest degree first


def check(s1, c1):
    rs = np.linspace(-0.25, 0.25, 200001)
    worst_s = worst_c = 0.0
    for r in rs[::20]:
        z = r * r
        p = 0.0
        for c in s1[::-1]:
            p = p * z + c
        s = r * z * p + r * float(np.pi)  # fma-free model; the kernel uses fma (more accurate)
        q = 0.0
        for c in c1[::-1]:
            q = q * z + c
        c_ = z * q + 1.0
        ts, tc = mp.sin(mp.pi * mp.mpf(float(r))), mp.cos(mp.pi * mp.mpf(float(r)))
        if r != 0:
            worst_s = max(worst_s, abs(float((mp.mpf(s) - ts) / ts)))
        worst_c = max(worst_c, abs(float((mp.mpf(c_) - tc) / tc)))
    return worst_s, worst_c


for n in (6, 7, 8):
    s1, es = fit(S1, n)
    c1, ec = fit(C1, n)
    ws, wc = check(s1, c1)
    print(f"n={n}: fit err S1={es:.2e} C1={ec:.2e}; fp64 eval rel err sin={ws:.2e} cos={wc:.2e}")
    if n == 7:
        print("S1 =", ", ".join(f"{c!r}" for c in s1))
        print("C1 =", ", ".join(f"{c!r}" for c in c1))
        print("S1 hex:", [float(c).hex() for c in s1])
        print("C1 hex:", [float(c).hex() for c in c1])
