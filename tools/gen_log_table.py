"""Table of the table-driven ln() used by the CRLB / log-likelihood pass (csrc/mle_tps_core.cuh,
log_tab):  x = m 2^e, m in [1, 2);  i = top 7 mantissa bits;  c_i = 1 + (i + 1/2) / 128;
    rc_i = double(1 / c_i)            r = fma(m, rc_i, -1)   (|r| < 2^-8 + 2^-53)
    lc_i = double(-ln(rc_i))          ln x = e ln2 + lc_i + log1p(r),  log1p by a degree-5 Taylor sum
lc is computed from the ROUNDED rc with mpmath (60 digits), so ln m = lc + log1p(r) holds exactly up
to the rounding of lc and the polynomial error r^6/6 < 6e-16.  Prints two C arrays."""
import mpmath as mp

mp.mp.dps = 60
rc, lc = [], []
for i in range(128):
    c = mp.mpf(1) + (mp.mpf(i) + mp.mpf(1) / 2) / 128
    r = float(1 / c)
    rc.append(r)
    lc.append(float(-mp.log(mp.mpf(r))))


def arr(name, v):
    out = [f"PB_TABLE double {name}[128] = {{"]
    for k in range(0, 128, 4):
        out.append("    " + ", ".join(float(x).hex() for x in v[k:k + 4]) + ",")
    out.append("};")
    return "\n".join(out)


print(arr("kLogRc", rc))
print(arr("kLogLc", lc))
# self-check against mpmath
import random
random.seed(1)
worst = 0
for _ in range(20000):
    x = random.uniform(0.01, 1e5)
    m, e = mp.frexp(x)
    mm, ee = float(m) * 2, int(e) - 1
    i = int((mm - 1) * 128)
    r = mm * rc[i] - 1     # (fma in the real code)
    p = r * (1 + r * (-0.5 + r * (1 / 3 + r * (-0.25 + r * 0.2))))
    got = ee * 0.6931471805599453 + lc[i] + p
    worst = max(worst, abs(got - float(mp.log(x))))
import sys
print("max abs error", worst, file=sys.stderr)
