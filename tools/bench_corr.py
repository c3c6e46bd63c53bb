"""Correlation kernel alone at the 800x1280 size (N = 16000, C = 128, 1 object), for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicorn_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
g = torch.Generator().manual_seed(0)
k = (torch.randn(N, 128, generator=g) * 0.5).cuda().half()
q = (torch.randn(N, 128, generator=g) * 0.5).cuda().half()
v = torch.rand(1, N, generator=g).cuda()
out = torch.empty(1, N, device="cuda")
for _ in range(3):
    ops.corr_propagate(k, q, v, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.corr_propagate(k, q, v, out=out)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) * 100
print(f"corr N={N}: {us:.1f} us  {2.0*N*N*129/us/1e6:.1f} TFLOP/s")
