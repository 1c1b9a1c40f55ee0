"""Fit of the single-exp2 erfc used by gelu_erf() in vfmseg_b200/csrc/sm100_ptx.cuh, and its error scan (CPU, numpy/scipy)."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf, erfc

Z, DEG = 6.0, 6
z = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * Z / 2 + Z / 2
c = C.chebfit(2 * z / Z - 1, np.log2(erfc(z)), DEG, w=np.maximum(erfc(z), 1e-6))
pa = np.zeros(1)
for k, ck in enumerate(C.cheb2poly(c)):
    pa = P.polyadd(pa, ck * P.polypow([-1.0, 2.0 / Z], k))
co = np.float32(pa)
print("coefficients a^0..a^6:", [float(v) for v in co])
x = np.linspace(-9, 9, 400001).astype(np.float32)
a = np.minimum(np.abs(x) * np.float32(0.70710678), np.float32(Z))
q = np.full_like(a, co[DEG])
for k in range(DEG - 1, -1, -1):
    q = (q * a + co[k]).astype(np.float32)
e = np.exp2(q.astype(np.float64)).astype(np.float32)
g = (np.float32(0.5) * x * np.where(x < 0, e, np.float32(2) - e)).astype(np.float32)
gt = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("max abs error of gelu:", np.abs(g - gt).max())
m = (x > -4) & (x != 0)
print("max rel error for x > -4:", (np.abs(g - gt)[m] / np.abs(gt[m])).max())
