"""Micro-benchmark of uc_conv2d on the layer shapes of unicorn_track_large @ 800x1280 (CUDA events, L2-warm loop)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicorn_b200 import ops
dev = "cuda"
SHAPES = [
    ("s1.pw1", 1, 200, 320, 192, 768, 1, 1, dict(gelu=1)),
    ("s1.pw2", 1, 200, 320, 768, 192, 1, 1, dict(res=1)),
    ("s2.pw1", 1, 100, 160, 384, 1536, 1, 1, dict(gelu=1)),
    ("s2.pw2", 1, 100, 160, 1536, 384, 1, 1, dict(res=1)),
    ("s3.pw1", 1, 50, 80, 768, 3072, 1, 1, dict(gelu=1)),
    ("s3.pw2", 1, 50, 80, 3072, 768, 1, 1, dict(res=1)),
    ("s3.plain1", 1, 50, 80, 768, 3072, 1, 1, dict()),      # s3.pw1 without the GELU: what the epilogue math costs
    ("s4.pw1", 1, 25, 40, 1536, 6144, 1, 1, dict(gelu=1)),
    ("s4.pw2", 1, 25, 40, 6144, 1536, 1, 1, dict(res=1)),
    ("down2", 1, 100, 160, 384, 768, 2, 2, dict()),
    ("head3x3.l0", 1, 100, 160, 256, 256, 3, 1, dict(gn=16)),
    ("head3x3.l1", 1, 50, 80, 256, 256, 3, 1, dict(gn=16)),
    ("csp3x3.p4", 1, 50, 80, 384, 384, 3, 1, dict(gn=16)),
    ("csp1x1.c12", 1, 50, 80, 1536, 768, 1, 1, dict(gn=32)),
    ("up1", 1, 100, 160, 64, 256, 3, 1, dict()),
    ("up3", 1, 100, 160, 256, 128, 3, 1, dict()),
    ("null.s4", 1, 25, 40, 64, 6144, 1, 1, dict(gelu=1)),   # one K step per tile: launch + prologue + epilogue cost only
]
only = [a for a in sys.argv[1:] if "=" not in a or a.startswith("bn=")]
OPT = dict(a.split("=") for a in sys.argv[1:] if "=" in a and not a.startswith("bn="))
REPS = [int(r) for r in OPT.get("R", "20").split(",")]
ZERO = int(OPT.get("zero", 0))
for name, B, H, W, Cin, Cout, K, s, ex in SHAPES:
    if only and not any(o in name for o in only if not o.startswith("bn=")):
        continue
    bns = [int(o[3:]) for o in only if o.startswith("bn=")] or [0]
    bns = [b for b in bns if b < 1000 or not ex.get("gn") or (b - 1000) % (Cout // ex["gn"]) == 0]
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    if ZERO: x.zero_()
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, K, K, device=dev) / (Cin * K * K) ** 0.5)
    pad = (K - 1) // 2 if s == 1 else (1 if K == 3 else 0)
    Ho = (H + 2 * pad - K) // s + 1; Wo = (W + 2 * pad - K) // s + 1
    bias = torch.randn(Cout, device=dev)
    gamma = torch.randn(Cout, device=dev) if ex.get("res") else None
    res = torch.randn(B, Ho, Wo, Cout, device=dev).bfloat16() if ex.get("res") else None
    out = torch.empty(B, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(B, ex["gn"], 2, device=dev, dtype=torch.int64) if ex.get("gn") else None
    for bn in bns:
        def run():
            ops.conv2d(x, w, K, K, s, pad, bias=bias, act=ops.ACT_GELU if ex.get("gelu") else 0, gamma=gamma, res=res, out=out,
                       gn_stats=st, gn_groups=ex.get("gn", 0), block_n=bn)
        try:
            for _ in range(3): run()
        except Exception as e:
            print(name, "bn", bn, "ERR", str(e)[:100]); continue
        torch.cuda.synchronize()
        fl = 2.0 * B * Ho * Wo * Cout * Cin * K * K
        for R in REPS:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2_000_000 + 60_000 * R)  # the host queues all R launches behind it, so they run back to back
            a.record()
            for _ in range(R): run()
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / R
            print(f"{name:12s} bn={bn:3d} M={B*Ho*Wo:6d} N={Cout:5d} K={Cin*K*K:5d} R={R:3d} {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s", flush=True)
        if OPT.get("graph"):
            g = torch.cuda.CUDAGraph(); R = 50
            with torch.cuda.graph(g):
                for _ in range(R): run()
            g.replay(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / R
            print(f"{name:12s} bn={bn:3d} graph R={R} {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s", flush=True)
