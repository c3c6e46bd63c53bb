"""Golden ByteTrack outputs from the reference's BYTETracker.update (build container only).  The two third-party
pieces that are absent offline are emulated and stated: lap.lapjv -> scipy linear_sum_assignment on lap's own
extended cost matrix; cython_bbox.bbox_overlaps -> numpy restatement (inclusive-pixel IoU, float64).  Everything else
(STrack, Kalman filter, the three association stages, bookkeeping) is the reference's own code."""
import os
import sys
import types

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
np.float, np.int = float, int  # removed numpy aliases the reference still uses (byte_tracker.py:18, matching.py:61)
import ref_import  # noqa: E402
ref_import.install()


def lapjv(cost, extend_cost=True, cost_limit=np.inf):
    n, m = cost.shape
    ext = np.full((n + m, n + m), cost_limit / 2.0)
    ext[n:, m:] = 0
    ext[:n, :m] = cost
    r, c = linear_sum_assignment(ext)
    x, y = -np.ones(n, dtype=int), -np.ones(m, dtype=int)
    for i, j in zip(r, c):
        if i < n and j < m:
            x[i], y[j] = j, i
    return 0.0, x, y


def bbox_overlaps(a, b):
    out = np.zeros((a.shape[0], b.shape[0]))
    for k in range(b.shape[0]):
        area = (b[k, 2] - b[k, 0] + 1) * (b[k, 3] - b[k, 1] + 1)
        for n in range(a.shape[0]):
            iw = min(a[n, 2], b[k, 2]) - max(a[n, 0], b[k, 0]) + 1
            if iw > 0:
                ih = min(a[n, 3], b[k, 3]) - max(a[n, 1], b[k, 1]) + 1
                if ih > 0:
                    ua = (a[n, 2] - a[n, 0] + 1) * (a[n, 3] - a[n, 1] + 1) + area - iw * ih
                    out[n, k] = iw * ih / ua
    return out


sys.modules["lap"].lapjv = lapjv
sys.modules["cython_bbox"].bbox_overlaps = bbox_overlaps
from unicorn.tracker.byte_tracker import BYTETracker  # noqa: E402
from unicorn.tracker.basetrack import BaseTrack  # noqa: E402
from unicorn_b200.synthetic import make_detections  # noqa: E402

args = types.SimpleNamespace(track_thresh=0.6, track_buffer=30, match_thresh=0.9, mot20=False)
BaseTrack._count = 0
trk = BYTETracker(args)
frames = make_detections(n_frames=30, n_obj=12, seed=3)
res = {}
for i, (boxes, _) in enumerate(frames):
    out = trk.update(boxes.numpy().copy(), (800, 1280), (800, 1280))
    rows = np.array([[t.track_id, *t.tlwh, t.score] for t in out]).reshape(-1, 6)
    res[f"f{i}"] = rows[np.argsort(rows[:, 0])] if len(rows) else rows
print("frames", len(frames), "ids last frame", res[f"f{len(frames)-1}"][:, 0], "total ids", BaseTrack._count)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "byte_tracker.npz"), n_frames=30, n_obj=12, seed=3,
                    total_ids=BaseTrack._count, **res)
print("wrote byte_tracker.npz")
