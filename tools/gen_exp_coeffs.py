"""Coefficients of the exp() used by csrc/mle_tps_core.cuh (exp_neg):

    exp(x), x <= 0:  n = rint(x * log2 e),  r = x - n ln2 (hi/lo split),  exp(r) ~ P(r),
    result = P(r) * 2^n   (exponent add on the high word)

P = degree-11 polynomial from a Chebyshev interpolant of exp on [-ln2/2, ln2/2] computed with
mpmath at 60 digits (near-minimax: max relative error ~3e-18 before rounding).
Prints the coefficients (monomial basis, r^0 .. r^11) as C hex-exact doubles.
"""
import mpmath as mp

mp.mp.dps = 60
DEG = 11
h = mp.log(2) / 2 * mp.mpf("1.0005")     # a little wider than needed (rint ties, lo-part error)
N = DEG + 1
nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / N) for k in range(N)]
fvals = [mp.exp(h * x) for x in nodes]
# Chebyshev coefficients by discrete orthogonality
cheb = []
for j in range(N):
    s = sum(fvals[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / N) for k in range(N))
    cheb.append(s * 2 / N)
cheb[0] /= 2
# Chebyshev -> monomial in x, then x = r / h
T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
for n in range(2, N):
    a = [mp.mpf(0)] + [2 * c for c in T[n - 1]]
    b = T[n - 2] + [mp.mpf(0)] * (len(a) - len(T[n - 2]))
    T.append([x - y for x, y in zip(a, b)])
mono = [mp.mpf(0)] * N
for j in range(N):
    for k, c in enumerate(T[j]):
        mono[k] += cheb[j] * c
mono = [c / h ** k for k, c in enumerate(mono)]
# error check at high precision
worst = mp.mpf(0)
for i in range(2001):
    r = -h + 2 * h * i / 2000
    p = sum(c * r ** k for k, c in enumerate(mono))
    worst = max(worst, abs(p / mp.exp(r) - 1))
print(f"// max relative error of P on [-ln2/2, ln2/2] (exact arithmetic): {mp.nstr(worst, 3)}")
for k, c in enumerate(mono):
    print(f"    {float(c)!r},   // r^{k}  {float(c).hex()}")
ln2 = mp.log(2)
hi = float(ln2)
# hi part with 11 trailing zero bits so n * hi is exact for |n| < 2^11
import struct
bits = struct.unpack("<Q", struct.pack("<d", hi))[0] & ~((1 << 11) - 1)
hi = struct.unpack("<d", struct.pack("<Q", bits))[0]
lo = float(ln2 - mp.mpf(hi))
print(f"// log2(e) = {float(1 / ln2)!r};  ln2_hi = {hi!r} ({hi.hex()});  ln2_lo = {lo!r}")
