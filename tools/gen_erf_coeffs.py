"""Coefficients of the erf approximation used by csrc/mle_fit.cu:

    erf(z) = sign(z) * (1 - exp(-z^2) * P(x)),  t = 1 / (1 + |z| / 2),  x = (2 t - (1 + t_lo)) / (1 - t_lo)

P = degree-16 polynomial (monomial basis in x on [-1, 1], from a Chebyshev interpolant of
erfcx(z(t)) on z in [0, 6]); exp(-z^2) is the Gaussian edge term the kernel needs anyway.
Prints the C array and the measured max abs error of a float64 Horner evaluation.
"""
import numpy as np
from numpy.polynomial import chebyshev as Ch
from scipy.special import erf, erfcx

ZMAX, C, DEG = 6.0, 0.5, 16
tlo, thi = 1 / (1 + C * ZMAX), 1.0
k = np.arange(8 * DEG)
x = np.cos(np.pi * (k + 0.5) / (8 * DEG))
t = 0.5 * (thi - tlo) * x + 0.5 * (thi + tlo)
z = (1 / t - 1) / C
cheb = Ch.chebfit(x, erfcx(z), DEG)
mono = Ch.cheb2poly(cheb)           # P(x) = sum mono[k] x^k


def erf_approx(zz):
    a = np.abs(zz)
    tt = 1.0 / (1.0 + C * a)
    xx = (2 * tt - (thi + tlo)) / (thi - tlo)
    p = np.zeros_like(xx) + mono[-1]
    for c in mono[-2::-1]:
        p = p * xx + c
    r = 1.0 - np.exp(-a * a) * p
    return np.sign(zz) * r


zz = np.linspace(-8, 8, 1_000_001)
err = np.abs(erf_approx(zz) - erf(zz)).max()
print(f"// max |erf_approx - erf| on [-8, 8] (float64 Horner): {err:.3e}")
print(f"// t_lo = {tlo!r};  x = t * {2 / (thi - tlo)!r} - {(thi + tlo) / (thi - tlo)!r}")
print("static __device__ const double kErfcxPoly[%d] = {" % (DEG + 1))
for c in mono:
    print(f"    {c!r},")
print("};")
