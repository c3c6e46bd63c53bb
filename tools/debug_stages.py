"""Developer aid: compare engine intermediates with the oracle at fine granularity (tiny config)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unicorn_oracle as orc
from unicorn_b200 import ops
from unicorn_b200.engine import UnicornEngine, _rows
from unicorn_b200.synthetic import make_video
from unicorn_b200.weights import make_state_dict

name = sys.argv[1] if len(sys.argv) > 1 else "unicorn_track_tiny"
H, W = (320, 320) if "tiny" in name else (800, 1280)
sd = make_state_dict(name, 0)
cfg = orc.CONFIGS[name]
frames, boxes = make_video(2, H, W, seed=0)
img = frames[1:2]
eng = UnicornEngine(sd, name)
P = eng.P

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
def nchw(t): return t.float().permute(0, 3, 1, 2).cpu()

p = "backbone.backbone."
with torch.no_grad():
    # stem
    xo = F.conv2d(img, sd[p + "downsample_layers.0.0.weight"], sd[p + "downsample_layers.0.0.bias"], stride=4)
    xo = orc.layernorm_cf(xo, sd[p + "downsample_layers.0.1.weight"], sd[p + "downsample_layers.0.1.bias"])
    imgd = img.cuda()
    eng.begin_frame()
    x = ops.stem_ln(imgd, *P["stem"])
    print("stem", rel(nchw(x), xo))
    for i in range(4):
        if i > 0:
            d = p + f"downsample_layers.{i}."
            xo = orc.layernorm_cf(xo, sd[d + "0.weight"], sd[d + "0.bias"])
            xo = F.conv2d(xo, sd[d + "1.weight"], sd[d + "1.bias"], stride=2)
            lw, lb, cw, cb = P[f"down{i}"]
            Bx, Hx, Wx, Cx = x.shape
            t = ops.layernorm(x.view(-1, Cx), lw, lb, 1e-6).view(1, Hx, Wx, Cx)
            x = ops.conv2d(t, cw, 2, 2, 2, 0, bias=cb)
            print(f"down{i}", rel(nchw(x), xo))
        for j in range(cfg["depths"][i]):
            bp = P["stages"][i][j]
            if i == 0 and j == 0:
                pp = p + "stages.0.0."
                C = xo.shape[1]
                a = F.conv2d(xo, sd[pp + "dwconv.weight"], sd[pp + "dwconv.bias"], padding=3, groups=C).permute(0, 2, 3, 1)
                a = F.layer_norm(a, (C,), sd[pp + "norm.weight"], sd[pp + "norm.bias"], 1e-6)
                t = ops.dwconv7_ln(x, bp["dw"], bp["dwb"], bp["lnw"], bp["lnb"], 1e-6)
                print("  blk0 dw+ln", rel(t, a))
                hid = ops.conv2d(t, bp["w1"], 1, 1, bias=bp["b1"], act=ops.ACT_GELU)
                a2 = F.gelu(F.linear(a, sd[pp + "pwconv1.weight"], sd[pp + "pwconv1.bias"]))
                print("  blk0 pw1", rel(hid, a2))
            xo = orc.convnext_block(xo, sd, p + f"stages.{i}.{j}.")
            eng.convnext_block(x, bp, f"dbg.s{i}")
            if j in (0, cfg["depths"][i] - 1):
                print(f"stage{i} block{j}", rel(nchw(x), xo), "max", xo.abs().max().item())
