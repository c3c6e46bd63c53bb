"""The GELU of the conv_gemm epilogue: the five coefficients compiled into unicorn_b200/csrc/uc_epilogue.cuh (gelu2) are read from
the source and evaluated in fp32 exactly as the kernel does (x * rcp(1 + ex2(x * P(x^2)))); the result must stay within 4e-6 of
the exact erf GELU over the real line and saturate correctly — the bound DESIGN.md 4.1 states."""
import os
import re

import numpy as np
from scipy.special import ndtr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernel_gelu_constants_meet_the_stated_bound():
    src = open(os.path.join(ROOT, "unicorn_b200", "csrc", "uc_epilogue.cuh")).read()
    body = src[src.index("f32x2 gelu2(f32x2 x)"):]
    c = [np.float32(re.search(rf"c{i} = pk2\(([-+0-9.e]+)f", body).group(1)) for i in range(5)]
    assert abs(float(c[0]) + 1.59565837 * np.log2(np.e)) < 1e-6          # c0 = -log2(e) * first logit coefficient
    x = np.concatenate([np.linspace(-12, 12, 1_000_001), [-50.0, 50.0, -1e4, 1e4, 0.0]]).astype(np.float32)
    t = x * x
    pz = np.full_like(x, c[4])
    with np.errstate(over="ignore"):
        for k in (3, 2, 1, 0):
            pz = (pz * t + c[k]).astype(np.float32)
        u = (x * pz).astype(np.float32)
        y = (x * (np.float32(1) / (np.float32(1) + np.exp2(u).astype(np.float32)))).astype(np.float32)
    ref = x.astype(np.float64) * ndtr(x.astype(np.float64))
    assert np.abs(y[:-5] - ref[:-5]).max() < 4e-6
    assert np.array_equal(y[-5:], np.array([-0.0, 50.0, -0.0, 1e4, 0.0], dtype=np.float32))
