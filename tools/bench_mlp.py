"""Micro-benchmark of the ConvNeXt block back half (LayerNorm -> pwconv1 -> GELU -> pwconv2 -> layer scale -> residual): the fused launch
(uc_convnext_mlp) against the three separate kernels (uc_layernorm + 2 x uc_conv2d) on the stage-1 shapes, CUDA-graph replays timed
with CUDA events; the hidden map of the unfused path (M x 4C bf16) is what does not fit L2."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicorn_b200 import ops
dev = "cuda"
R = 10


def timed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (2 * R)


for name, M, C in [("L.s1 800x1280", 64000, 192), ("L.s1 1536x2048", 196608, 192), ("T.s1 800x1280", 64000, 96), ("T.s2 800x1280", 16000, 192), ("L.s2 800x1280", 16000, 384), ("L.s2 1536x2048", 49152, 384), ("head.l0 800x1280", 16000, 256), ("head.l1 800x1280", 4000, 256)]:
    t = torch.randn(M, C, device=dev).bfloat16()
    x = torch.randn(M, C, device=dev).bfloat16()
    lw, lb = torch.randn(C, device=dev), torch.randn(C, device=dev)
    w1 = torch.randn(4 * C, C, device=dev) / C ** 0.5
    b1, b2, gamma = torch.randn(4 * C, device=dev), torch.randn(C, device=dev), torch.randn(C, device=dev) * 0.1
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).bfloat16().contiguous()
    w1f = (w1 * lw[None]).bfloat16().contiguous()
    w1b = w1.bfloat16().contiguous()
    c1 = (w1 @ lb + b1).contiguous()
    tn, hid = torch.empty_like(t), torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    t_f = timed(lambda: ops.convnext_mlp(t, w1f, c1, w2, b2, gamma, x))

    def unfused():
        ops.layernorm(t, lw, lb, 1e-6, out=tn)
        ops.conv2d(tn.view(1, 1, M, C), w1b.view(4 * C, 1, C), 1, 1, bias=b1, act=ops.ACT_GELU, out=hid.view(1, 1, M, 4 * C))
        ops.conv2d(hid.view(1, 1, M, 4 * C), w2.view(C, 1, 4 * C), 1, 1, bias=b2, gamma=gamma, res=x.view(1, 1, M, C), out=x.view(1, 1, M, C))
    t_u = timed(unfused)
    fl = 2.0 * 2 * M * C * 4 * C
    print(f"{name:16s} M={M:6d} C={C:3d}  fused {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} TFLOP/s)   layernorm + pwconv1 + pwconv2 {t_u:7.1f} us ({fl / t_u / 1e6:6.1f} TFLOP/s)", flush=True)
