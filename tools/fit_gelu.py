"""Fit of the single-exp2 GELU used by gelu_erf() in vfmseg_b200/csrc/sm100_ptx.cuh, and its error scan (CPU, numpy/scipy).

gelu(x) = relu(x) - |x| * h(|x|), h(t) = 0.5 erfc(t / sqrt 2) = 2^q(t); q = degree-DEG polynomial in t on [0, 6 sqrt 2], weighted
least squares on Chebyshev nodes with weight t * h(t) (the absolute error of the GELU value). Prints the coefficients t^DEG..t^0
as they appear in the Horner chain, the maximum absolute error and the relative error above a few magnitude thresholds.
Usage: python tools/fit_gelu.py [DEG]   (shipped: 5; 6 -> 3e-7, 4 -> 7e-6)"""
import math
import sys

import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf, erfc

DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 5
T = 6.0 * math.sqrt(2)
n = 6000
t = np.cos(np.pi * (np.arange(n) + 0.5) / n) * T / 2 + T / 2
h = 0.5 * erfc(t / math.sqrt(2))
c = C.chebfit(2 * t / T - 1, np.log2(h), DEG, w=np.maximum(t * h, 1e-6))
pa = np.zeros(1)
for k, ck in enumerate(C.cheb2poly(c)):
    pa = P.polyadd(pa, ck * P.polypow([-1.0, 2.0 / T], k))
co = np.float32(pa)
print(f"coefficients t^{DEG}..t^0:", [repr(float(v)) for v in co[::-1]])
x = np.linspace(-9, 9, 2000001).astype(np.float32)
tt = np.minimum(np.abs(x), np.float32(T)).astype(np.float32)
q = np.full_like(tt, co[DEG])
for k in range(DEG - 1, -1, -1):
    q = (q * tt + co[k]).astype(np.float32)
e = np.exp2(q.astype(np.float64)).astype(np.float32)
g = (np.maximum(x, 0) - np.abs(x) * e).astype(np.float32)
gt = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / math.sqrt(2)))
err = np.abs(g - gt)
print("max abs error of gelu:", err.max(), "at x =", float(x[err.argmax()]))
for thr in (1e-2, 1e-3, 1e-4, 1e-5):
    m = np.abs(gt) > thr
    print(f"max rel error where |gelu| > {thr:g}:", (err[m] / np.abs(gt[m])).max())
