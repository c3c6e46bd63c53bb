import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unicorn_oracle as orc
from unicorn_b200 import ops
from unicorn_b200.engine import UnicornEngine
from unicorn_b200.synthetic import make_video
from unicorn_b200.weights import make_state_dict
name = "unicorn_track_tiny"; H = W = 320
sd = make_state_dict(name, 0); cfg = orc.CONFIGS[name]
frames, boxes = make_video(2, H, W, seed=0)
img = frames[1:2]
eng = UnicornEngine(sd, name)
def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
def nchw(t): return t.float().permute(0, 3, 1, 2).cpu()
with torch.no_grad():
    feats = orc.convnext_features(img, sd, cfg)
    x2, x1, x0 = feats
    p = "backbone."
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    o = {}
    o["x2n"], o["x1n"], o["x0n"] = x2, x1, x0
    o["fpn_out0"] = orc.base_conv(x0, sd, p + "lateral_conv0.", 1)
    o["f_out0"] = orc.csp_layer(torch.cat([up(o["fpn_out0"]), x1], 1), sd, p + "C3_p4.")
    o["fpn_out1"] = orc.base_conv(o["f_out0"], sd, p + "reduce_conv1.", 1)
    o["pan_out2"] = orc.csp_layer(torch.cat([up(o["fpn_out1"]), x2], 1), sd, p + "C3_p3.")
    p_out1 = orc.base_conv(o["pan_out2"], sd, p + "bu_conv2.", 3, 2)
    o["pan_out1"] = orc.csp_layer(torch.cat([p_out1, o["fpn_out1"]], 1), sd, p + "C3_n3.")
    p_out0 = orc.base_conv(o["pan_out1"], sd, p + "bu_conv1.", 3, 2)
    o["pan_out0"] = orc.csp_layer(torch.cat([p_out0, o["fpn_out0"]], 1), sd, p + "C3_n4.")
    eng.begin_frame()
    eng.backbone(img.cuda(), tag="cur")
    torch.cuda.synchronize()
    for k in ("x2n", "x1n", "x0n", "fpn_out0", "f_out0", "fpn_out1", "pan_out2", "pan_out1", "pan_out0"):
        print(k, f"{rel(nchw(eng.dbg[k]), o[k]):.4f}", "max", f"{o[k].abs().max().item():.3f}")
    # CSP internals of C3_p4
    cp = eng.P["C3_p4"]
    xin = torch.cat([up(o["fpn_out0"]), x1], 1)
    a1 = orc.base_conv(xin, sd, p + "C3_p4.conv1.", 1); a2 = orc.base_conv(xin, sd, p + "C3_p4.conv2.", 1)
    xin_d = ops.nchw_to_nhwc(xin.cuda().contiguous())
    eng.begin_frame()
    cat = torch.empty(1, xin.shape[2], xin.shape[3], cp["c12"].cout, device="cuda", dtype=torch.bfloat16)
    eng.conv_gn(xin_d, cp["c12"], cat)
    hd = cp["c12"].cout // 2
    print("csp c12 first half", rel(nchw(cat[..., :hd]), a1), "second", rel(nchw(cat[..., hd:]), a2))
    c1, c2 = cp["m"][0]
    t = eng.conv_gn(cat[..., :hd], c1, torch.empty(1, xin.shape[2], xin.shape[3], hd, device="cuda", dtype=torch.bfloat16))
    b1 = orc.base_conv(a1, sd, p + "C3_p4.m.0.conv1.", 1)
    print("m0.conv1", rel(nchw(t), b1))
    u = eng.conv_gn(t, c2, torch.empty(1, xin.shape[2], xin.shape[3], hd, device="cuda", dtype=torch.bfloat16))
    b2 = orc.base_conv(b1, sd, p + "C3_p4.m.0.conv2.", 3)
    print("m0.conv2", rel(nchw(u), b2))
