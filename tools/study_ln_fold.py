"""Round-2 study (CPU, no GPU needed): numerical cost of folding the ConvNeXt block's LayerNorm into pwconv1.

current path : t = bf16(dwconv(x));  u = bf16(LN(t));            h = u @ bf16(W)^T + b1          (fp32 accumulate)
folded path  : t = bf16(dwconv(x));  W' = bf16(W * g);  h = rstd * (t @ W'^T) - rstd * mu * rowsum(W') + (W @ beta + b1)
Both are compared with the fp32 evaluation of the reference block on the SAME input x, over the blocks of a model with the
repo's seeded weights on a synthetic frame.  Prints per stage: max / rms error of the pre-GELU activations relative to their
rms, and the per-pixel |mu| / sigma ratio that governs the cancellation in the folded form."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unicorn_oracle as orc
from unicorn_b200.synthetic import make_video
from unicorn_b200.weights import make_state_dict

name = sys.argv[1] if len(sys.argv) > 1 else "unicorn_track_tiny"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (320, 320)
bf = lambda t: t.bfloat16().float()  # noqa: E731
sd = make_state_dict(name, 0)
cfg = orc.CONFIGS[name]
frames, _ = make_video(1, H, W, seed=0)
torch.set_num_threads(min(16, os.cpu_count() or 1))
p = "backbone.backbone."
x = frames[0:1]
rows = []
with torch.no_grad():
    for i in range(4):
        d = p + f"downsample_layers.{i}."
        if i == 0:
            x = orc.layernorm_cf(F.conv2d(x, sd[d + "0.weight"], sd[d + "0.bias"], stride=4), sd[d + "1.weight"], sd[d + "1.bias"])
        else:
            x = F.conv2d(orc.layernorm_cf(x, sd[d + "0.weight"], sd[d + "0.bias"]), sd[d + "1.weight"], sd[d + "1.bias"], stride=2)
        for j in range(cfg["depths"][i]):
            q = p + f"stages.{i}.{j}."
            C = x.shape[1]
            t = F.conv2d(x, sd[q + "dwconv.weight"], sd[q + "dwconv.bias"], padding=3, groups=C).permute(0, 2, 3, 1).reshape(-1, C)
            g, beta, W1, b1 = sd[q + "norm.weight"], sd[q + "norm.bias"], sd[q + "pwconv1.weight"], sd[q + "pwconv1.bias"]
            ref = F.linear(F.layer_norm(t, (C,), g, beta, 1e-6), W1, b1)            # fp32 reference
            tb = bf(t)
            cur = F.linear(bf(F.layer_norm(tb, (C,), g, beta, 1e-6)), bf(W1), b1)    # what the engine computes today
            mu = tb.mean(1, keepdim=True)
            rstd = torch.rsqrt(tb.var(1, unbiased=False, keepdim=True) + 1e-6)
            Wf = bf(W1 * g[None, :])
            fold = rstd * (tb @ Wf.t()) - rstd * mu * Wf.sum(1)[None, :] + (W1 @ beta + b1)[None, :]
            rms = ref.pow(2).mean().sqrt()
            e = lambda a: (((a - ref).abs().max() / rms).item(), ((a - ref).pow(2).mean().sqrt() / rms).item())  # noqa: E731
            ratio = (mu.abs() * rstd).flatten()
            rows.append((i, j, *e(cur), *e(fold), ratio.median().item(), ratio.max().item()))
            x = orc.convnext_block(x, sd, q)
print("stage blk | current max  rms | folded max  rms | |mu|/sigma median max")
for r in rows:
    print(f"  {r[0]}   {r[1]:2d}  | {r[2]:.2e} {r[3]:.2e} | {r[4]:.2e} {r[5]:.2e} | {r[6]:.2f} {r[7]:.2f}")

# ---- stress: add a per-pixel offset of k sigma to the dwconv output of the last block studied (LayerNorm is invariant to it
# in exact arithmetic); shows how the folded form degrades when trained weights produce |mu| >> sigma
print("stress (last block): offset k*sigma | current rms | folded rms")
with torch.no_grad():
    sig = t.std(1, keepdim=True)
    for k in (0, 1, 3, 10, 30, 100):
        ts = t + k * sig
        ref = F.linear(F.layer_norm(ts, (C,), g, beta, 1e-6), W1, b1)
        tb = bf(ts)
        cur = F.linear(bf(F.layer_norm(tb, (C,), g, beta, 1e-6)), bf(W1), b1)
        mu = tb.mean(1, keepdim=True)
        rstd = torch.rsqrt(tb.var(1, unbiased=False, keepdim=True) + 1e-6)
        fold = rstd * (tb @ Wf.t()) - rstd * mu * Wf.sum(1)[None, :] + (W1 @ beta + b1)[None, :]
        rms = ref.pow(2).mean().sqrt()
        print(f"  k={k:3d}  {((cur - ref).pow(2).mean().sqrt() / rms).item():.2e}  {((fold - ref).pow(2).mean().sqrt() / rms).item():.2e}")
