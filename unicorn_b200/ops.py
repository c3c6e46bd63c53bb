"""Python wrappers over the C ABI.  Tensors are torch CUDA tensors used purely as device memory:
activations are NHWC views (B, H, W, C) with unit channel stride and a constant pixel stride (so a channel
slice of a wider buffer is a valid operand, which is how concatenations are formed without copies)."""
import ctypes

import torch

from . import _lib
from ._lib import BF16, F32, F16, ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU, ACT_SIGMOID  # noqa: F401

_DT = {torch.bfloat16: BF16, torch.float32: F32, torch.float16: F16}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _nhwc_ld(t):
    """pixel stride of an NHWC view; validates the layout."""
    assert t.dim() == 4 and t.is_cuda, "expected a CUDA NHWC tensor"
    B, H, W, C = t.shape
    ld = t.stride(2) if W > 1 else (t.stride(1) // W if H > 1 else max(C, t.stride(2)))
    assert t.stride(3) == 1 or C == 1
    if W > 1:
        assert t.stride(2) == ld
    if H > 1:
        assert t.stride(1) == W * ld, (t.shape, t.stride())
    if B > 1:
        assert t.stride(0) == H * W * ld
    return ld


def pack_conv_weight(w, dtype=torch.bfloat16, cout_pad=8):
    """[Cout, Cin, KH, KW] (or [Cout, Cin] for Linear) fp32 -> [Cout_pad, KH*KW, Cin] 16-bit, K-major."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    Cout, Cin, KH, KW = w.shape
    Cp = (Cout + cout_pad - 1) // cout_pad * cout_pad
    out = torch.zeros(Cp, KH * KW, Cin, dtype=dtype, device=w.device)
    out[:Cout] = w.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin).to(dtype)
    return out.contiguous()


def conv2d(x, w_packed, KH, KW, stride=1, pad=0, bias=None, act=ACT_NONE, gamma=None, res=None, out=None,
           out_dtype=None, block_n=0, gn_stats=None, gn_groups=0):
    """x: NHWC view (B,H,W,Cin) bf16/f16.  w_packed: [Cout, KH*KW, Cin].  Returns NHWC (B,Ho,Wo,Cout)."""
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == KH * KW and w_packed.shape[2] == Cin and w_packed.is_contiguous()
    assert w_packed.dtype == x.dtype
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, Cout, dtype=out_dtype or x.dtype, device=x.device)
    assert out.shape == (B, Ho, Wo, Cout), (out.shape, (B, Ho, Wo, Cout))
    d = _lib.UcConv2d()
    d.x, d.x_dtype = _p(x), _DT[x.dtype]
    d.B, d.H, d.W, d.Cin, d.ldx = B, H, W, Cin, _nhwc_ld(x)
    d.w = _p(w_packed)
    d.Cout, d.KH, d.KW, d.stride, d.pad = Cout, KH, KW, stride, pad
    d.bias, d.act, d.gamma = _p(bias), act, _p(gamma)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= Cout
    if gamma is not None:
        assert gamma.dtype == torch.float32 and gamma.numel() >= Cout
    d.res, d.ldres = _p(res), (_nhwc_ld(res) if res is not None else 0)
    if res is not None:
        assert res.dtype == out.dtype and res.shape == out.shape
    d.y, d.ldy, d.y_dtype = _p(out), _nhwc_ld(out), _DT[out.dtype]
    d.block_n = block_n
    d.gn_stats, d.gn_groups = _p(gn_stats), gn_groups
    _lib.check(_lib.lib().uc_conv2d(ctypes.byref(d), _lib.stream_ptr()), "uc_conv2d")
    return out


def linear(x2d, w_packed, **kw):
    """x2d: [M, K] rows (unit inner stride).  Returns [M, N]."""
    M, K = x2d.shape
    assert x2d.stride(1) == 1
    x4 = x2d.as_strided((1, 1, M, K), (M * x2d.stride(0), M * x2d.stride(0), x2d.stride(0), 1))
    out = kw.pop("out", None)
    res = kw.pop("res", None)
    if out is not None:
        out = out.as_strided((1, 1, M, out.shape[1]), (M * out.stride(0), M * out.stride(0), out.stride(0), 1))
    if res is not None:
        res = res.as_strided((1, 1, M, res.shape[1]), (M * res.stride(0), M * res.stride(0), res.stride(0), 1))
    y = conv2d(x4, w_packed, 1, 1, out=out, res=res, **kw)
    return y.as_strided((M, y.shape[3]), (y.stride(2), 1))
