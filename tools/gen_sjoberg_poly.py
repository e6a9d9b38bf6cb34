#!/usr/bin/env python
"""Coefficients of the device-side Sjoberg slot exponent: p(u) ~ t^2.4 on t in [0.98, 1.785],
u = (t - mid) / half, Chebyshev interpolant of degree 16 rewritten in powers of u (Horner + fma on
the device).  dw_sjoberg(t) = exp(-t^2.4) (dwflow.c:620-633) is evaluated on the device as
exp(-p(u)); measured error against the exact value <= 7 ulp (glibc's own exp(-pow(t, 2.4)): 3.8 ulp,
the exponent amplifies its rounding error by up to t^2.4 = 4).  Prints the C initialiser.
    python tools/gen_sjoberg_poly.py > /tmp/coeffs.txt"""
import mpmath as mp

mp.mp.dps = 60
N = 16
a, b = mp.mpf("0.98"), mp.mpf("1.785")
f = lambda t: mp.power(t, mp.mpf("2.4"))
ks = range(N + 1)
nodes = [mp.cos(mp.pi * (k + mp.mpf(0.5)) / (N + 1)) for k in ks]
fv = [f((a + b) / 2 + (b - a) / 2 * u) for u in nodes]
c = [mp.fsum(fv[k] * mp.cos(mp.pi * j * (k + mp.mpf(0.5)) / (N + 1)) for k in ks) * 2 / (N + 1) for j in range(N + 1)]
c[0] /= 2
T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
for n in range(2, N + 1):
    t = [mp.mpf(0)] + [2 * x for x in T[-1]]
    for i, x in enumerate(T[-2]):
        t[i] -= x
    T.append(t)
mono = [mp.mpf(0)] * (N + 1)
for j in range(N + 1):
    for i, x in enumerate(T[j]):
        mono[i] += c[j] * x
print("#define SWB_SJOBERG_MID  %s" % float((a + b) / 2).hex())
print("#define SWB_SJOBERG_RHALF %s   /* 1 / half-width */" % float(2 / (b - a)).hex())
print("#define SWB_SJOBERG_COEFFS { \\")
for i, x in enumerate(mono):
    print("    %s%s \\" % (float(x).hex(), "," if i < N else ""))
print("}")
