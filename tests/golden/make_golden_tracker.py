"""Golden ids from the UNMODIFIED reference QuasiDenseEmbedTracker (build container only) on a seeded sequence of
synthetic detections; also checks oracle/tracker_oracle.py against it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
ref_import.install()
from unicorn.tracker.quasi_dense_embed_tracker import QuasiDenseEmbedTracker  # noqa: E402
import tracker_oracle as to  # noqa: E402
from unicorn_b200.synthetic import make_detections  # noqa: E402

frames = make_detections(n_frames=25, n_obj=14, seed=0)
ref, orc = QuasiDenseEmbedTracker(), to.QDTrackerOracle()
all_ids, all_boxes = [], []
for fid, (boxes, feats) in enumerate(frames, start=1):
    labels = torch.ones(boxes.size(0))
    rb, _, rid = ref.match(boxes.clone(), labels.clone(), feats.clone(), fid)
    ob, _, oid = orc.match(boxes.clone(), labels.clone(), feats.clone(), fid)
    assert torch.equal(rid, oid), (fid, rid, oid)
    assert torch.equal(rb, ob)
    all_ids.append(rid.numpy())
    all_boxes.append(rb.numpy())
print("frames", len(frames), "tracklets", ref.num_tracklets, "ids in last frame", all_ids[-1])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "qd_tracker.npz"),
                    n_frames=25, n_obj=14, seed=0, num_tracklets=int(ref.num_tracklets),
                    **{f"ids_{i}": a for i, a in enumerate(all_ids)}, **{f"boxes_{i}": a for i, a in enumerate(all_boxes)})
print("wrote qd_tracker.npz")
