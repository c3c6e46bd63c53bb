import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicorn_b200.engine import UnicornEngine
from unicorn_b200.sot import UnicornSOTTrack
from unicorn_b200.synthetic import make_video
from unicorn_b200.weights import make_state_dict
name = "unicorn_track_tiny"
sd = make_state_dict(name, 0)
frames, boxes = make_video(3, 320, 320, seed=0)
eng = UnicornEngine(sd, name)
def snap(t):
    return {k: ([x.float().clone() for x in v] if isinstance(v, (tuple, list)) else v.float().clone()) for k, v in t.last.items()}
def diff(a, b):
    out = {}
    for k in a:
        if isinstance(a[k], list):
            out[k] = max((x - y).abs().max().item() for x, y in zip(a[k], b[k]))
        else:
            out[k] = (a[k] - b[k]).abs().max().item()
    return out
e = UnicornSOTTrack(eng, (320, 320), use_graph=False)
e.initialize_tensor(frames[0:1], boxes[0, 0])
d_a, n_a = e.track_tensor(frames[2:3]); s_a = snap(e)
d_b, n_b = e.track_tensor(frames[2:3]); s_b = snap(e)
print("eager vs eager", n_a, n_b, diff(s_a, s_b)); print(d_a[:2]); print(d_b[:2])
g = UnicornSOTTrack(eng, (320, 320), use_graph=True)
g.initialize_tensor(frames[0:1], boxes[0, 0])
g.track_tensor(frames[1:2])
d_c, n_c = g.track_tensor(frames[2:3]); s_c = snap(g)
print("eager vs graph", n_a, n_c, diff(s_a, s_c)); print(d_c[:2])
d_d, n_d = g.track_tensor(frames[2:3]); s_d = snap(g)
print("graph vs graph", n_c, n_d, diff(s_c, s_d))
