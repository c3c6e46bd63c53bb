import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from unicorn_b200 import ops
from unicorn_b200.engine import UnicornEngine, _rows
from unicorn_b200.synthetic import make_video
from unicorn_b200.weights import make_state_dict
name = "unicorn_track_tiny"; H = W = 320
sd = make_state_dict(name, 0)
frames, boxes = make_video(2, H, W, seed=0)
eng = UnicornEngine(sd, name)
eng.begin_frame()
eng.backbone(frames[1:2].cuda(), tag="cur")
torch.cuda.synchronize()
def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
for i, key in ((1, "x2n"), (2, "x1n"), (3, "x0n")):
    xs = [v for k, v in eng._bufs.items() if k[0] == f"cur.x{i}"][0]
    nw, nb = eng.P[f"norm{i}"]
    C = xs.shape[-1]
    ref = F.layer_norm(xs.float(), (C,), nw, nb, 1e-6)
    print(key, "vs LN(stage buf)", rel(eng.dbg[key], ref), "shape", tuple(eng.dbg[key].shape), "strides", eng.dbg[key].stride(), "off", eng.dbg[key].storage_offset())
    # recompute into a fresh slice
    big = torch.zeros(1, xs.shape[1], xs.shape[2], 2 * C, device="cuda", dtype=torch.bfloat16)
    dst = big[..., C:]
    ops.layernorm(xs.view(-1, C), nw, nb, 1e-6, out=_rows(dst))
    torch.cuda.synchronize()
    print("   fresh slice", rel(dst, ref), "first half untouched", bool((big[..., :C] == 0).all()))
    r = _rows(dst)
    print("   rows view", tuple(r.shape), r.stride(), r.storage_offset(), r.data_ptr() - big.data_ptr())
