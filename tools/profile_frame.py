"""Run init + N eager (non-graph) SOT frames of a config; used under ncu to list per-kernel device times."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicorn_b200 import _lib, ops
from unicorn_b200.engine import UnicornEngine
from unicorn_b200.sot import UnicornSOTTrack
from unicorn_b200.synthetic import make_video
from unicorn_b200.weights import make_state_dict
name = sys.argv[1] if len(sys.argv) > 1 else "unicorn_track_large"
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W = (320, 320) if "tiny" in name else (800, 1280)
sd = make_state_dict(name, 0)
frames, boxes = make_video(nfr + 1, H, W, seed=0)
eng = UnicornEngine(sd, name)
trk = UnicornSOTTrack(eng, (H, W), use_graph=False)
trk.initialize_tensor(frames[0:1], boxes[0, 0])
print("launches after init", _lib.LAUNCHES)
for i in range(nfr):
    l0 = _lib.LAUNCHES
    ops.CONV_TRACE = [] if i == nfr - 1 else None
    trk.track_tensor(frames[1 + i:2 + i])
    print("frame", i, "launches", _lib.LAUNCHES - l0)
if os.environ.get("UC_CONV_TRACE"):
    import json
    json.dump(ops.CONV_TRACE, open(os.environ["UC_CONV_TRACE"], "w"))
if len(sys.argv) > 3:
    eng.save_tuning(sys.argv[3])
    print("saved tuning table", sys.argv[3], len(eng._bn_cache))
