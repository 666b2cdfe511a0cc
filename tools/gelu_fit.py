#!/usr/bin/env python
"""Minimax fit of 0.5 x (1 + tanh(x (c1 + c3 x^2 + c5 x^4))) to the exact-erf GELU over |x| <= 8
(the constants of gelu2 in csrc/tc_gemm.cu)."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf
x = np.linspace(-8, 8, 200001)
gelu = x * 0.5 * (1 + erf(x / np.sqrt(2)))
def g(c):
    x2 = x * x
    return 0.5 * x * (1 + np.tanh(x * (c[0] + x2 * (c[1] + x2 * c[2]))))
c0 = np.array([0.7978845608, 0.044715 * 0.7978845608, 0.0])
print("textbook constants: max |err| =", np.abs(g(c0) - gelu).max())
r = minimize(lambda c: np.abs(g(c) - gelu).max(), c0, method="Nelder-Mead",
             options={"xatol": 1e-12, "fatol": 1e-12, "maxiter": 20000})
print("fitted:", r.x, "max |err| =", r.fun)
