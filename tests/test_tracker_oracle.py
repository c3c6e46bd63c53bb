"""oracle/tracker_oracle.py against the ids produced by the UNMODIFIED reference QuasiDenseEmbedTracker
(tests/golden/qd_tracker.npz, written by tests/golden/make_golden_tracker.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tracker_oracle as to  # noqa: E402
from unicorn_b200.synthetic import make_detections  # noqa: E402


def test_qd_oracle_reproduces_reference_ids():
    g = np.load(os.path.join(ROOT, "tests", "golden", "qd_tracker.npz"))
    frames = make_detections(int(g["n_frames"]), int(g["n_obj"]), int(g["seed"]))
    trk = to.QDTrackerOracle()
    for i, (boxes, feats) in enumerate(frames):
        b, _, ids = trk.match(boxes, torch.ones(boxes.size(0)), feats, i + 1)
        assert np.array_equal(ids.numpy(), g[f"ids_{i}"]), i
        assert np.allclose(b.numpy(), g[f"boxes_{i}"])
    assert trk.num_tracklets == int(g["num_tracklets"])


def test_sample_embeddings_shapes():
    emb = torch.randn(1, 16, 10, 12)
    boxes = torch.tensor([[0.0, 0.0, 16.0, 16.0], [40.0, 30.0, 90.0, 70.0], [200.0, 200.0, 300.0, 300.0]])
    f = to.sample_embeddings(emb, boxes, (80, 96))
    assert f.shape == (3, 16)
