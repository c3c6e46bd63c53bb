"""Fit of the GELU used by the conv_gemm epilogue: gelu(x) = x * Phi(x) ~= x / (1 + 2^(x * P(x^2))), P of degree 4 in x^2
(least squares on the output error, Levenberg-Marquardt from a logit(Phi) fit).  Prints the coefficients (already multiplied
by -log2 e) and the max abs error of the fp32 evaluation against the exact erf form."""
import numpy as np
from scipy.special import ndtr, log_ndtr
from scipy.optimize import least_squares

def q(c, x):
    x2 = x * x
    p = np.zeros_like(x)
    for a in c[::-1]:
        p = p * x2 + a
    return x * p

xs = np.linspace(-9, 9, 7201)
g = xs * ndtr(xs)
m = np.abs(xs) < 4
A = np.stack([xs[m] ** (2 * k + 1) for k in range(5)], 1)
c0 = np.linalg.lstsq(A, (log_ndtr(xs) - log_ndtr(-xs))[m], rcond=None)[0]
c = least_squares(lambda c: (xs / (1 + np.exp(-q(c, xs))) - g) * 1e4, c0, method="lm", max_nfev=20000).x
cc = (-c * np.log2(np.e)).astype(np.float32)
print("coefficients * -log2(e):", ", ".join(f"{v:.8e}f" for v in cc))
x = np.concatenate([np.linspace(-12, 12, 2_000_001), [-1e4, 1e4, -1e20, 1e20, 0.0]]).astype(np.float32)
t = x * x
pz = np.float32(cc[4])
with np.errstate(over="ignore"):
    for k in (3, 2, 1, 0):
        pz = (pz * t + cc[k]).astype(np.float32)
    u = (x * pz).astype(np.float32)
    y = (x * (np.float32(1) / (np.float32(1) + np.exp2(u).astype(np.float32)))).astype(np.float32)
ref = x.astype(np.float64) * ndtr(x.astype(np.float64))
err = np.abs(y - ref)
print("max abs err (fp32 eval)", err[:-5].max(), "at", x[err[:-5].argmax()], "| extremes:", y[-5:], ref[-5:])
