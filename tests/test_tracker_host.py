"""Host logic of the trackers on CPU: BYTETracker.update (track life cycle, two-stage association, batched Kalman, LAP) against
the reference golden with the IoU KERNEL replaced by a numpy stand-in (the kernel itself is checked on the GPU in
tests/test_tracker_gpu.py), and the MOT driver's host half (collect) on hand-made result slots."""
import contextlib
import os
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _iou_np(a, b, plus_one=False):
    a, b, p = a.numpy().astype(np.float64), b.numpy().astype(np.float64), (1.0 if plus_one else 0.0)
    iw = np.clip(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]) + p, 0, None)
    ih = np.clip(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]) + p, 0, None)
    aa = (a[:, 2] - a[:, 0] + p) * (a[:, 3] - a[:, 1] + p)
    bb = (b[:, 2] - b[:, 0] + p) * (b[:, 3] - b[:, 1] + p)
    return torch.from_numpy((iw * ih / (aa[:, None] + bb[None] - iw * ih)).astype(np.float32))


def test_byte_tracker_host_logic_matches_reference_golden(monkeypatch):
    import unicorn_b200.tracker.byte_tracker as bt
    from unicorn_b200.synthetic import make_detections
    monkeypatch.setattr(bt.ops, "box_iou", _iou_np)
    monkeypatch.setattr(bt, "assoc_stream", lambda d: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    g = np.load(os.path.join(ROOT, "tests", "golden", "byte_tracker.npz"))
    args = types.SimpleNamespace(track_thresh=0.6, track_buffer=30, match_thresh=0.9, mot20=False)
    bt.STrack._count = 0
    trk = bt.BYTETracker(args, device="cpu")
    for i, (boxes, _) in enumerate(make_detections(int(g["n_frames"]), int(g["n_obj"]), int(g["seed"]))):
        out = trk.update(boxes.numpy().copy(), (800, 1280), (800, 1280))
        rows = np.array([[t.track_id, *t.tlwh, t.score] for t in out]).reshape(-1, 6)
        rows = rows[np.argsort(rows[:, 0])] if len(rows) else rows
        ref = g[f"f{i}"]
        assert rows.shape == ref.shape and np.array_equal(rows[:, 0], ref[:, 0]), (i, rows[:, 0], ref[:, 0])
        assert np.allclose(rows[:, 1:], ref[:, 1:], rtol=1e-4, atol=1e-2)
    assert bt.STrack._count == int(g["total_ids"])


def test_qd_tracker_host_logic_matches_reference_golden(monkeypatch):
    """QuasiDenseEmbedTracker.match / memo on CPU with numpy stand-ins for the two association kernels: the ids must
    bit-match the reference class's over the 25-frame golden sequence (tests/golden/qd_tracker.npz)."""
    import unicorn_b200.tracker.quasi_dense as qd
    from unicorn_b200.synthetic import make_detections

    def bisoftmax(e, m, ld=None, lm=None):
        f = e @ m.t()
        s = (f.softmax(1) + f.softmax(0)) / 2
        return s * (ld[:, None] == lm[None, :]).float() if ld is not None else s

    def qd_assign(scores, memo_ids, boxes5, match_thr, obj_thr, nms_conf_thr):  # the greedy loop uc_qd_assign runs on the device
        sc, ids = scores.clone(), torch.full((scores.shape[0],), -1, dtype=torch.long)
        for i in range(sc.shape[0]):
            conf, j = torch.max(sc[i], dim=0)
            if conf > match_thr and memo_ids[j] > -1:
                if boxes5[i, 4] > obj_thr:
                    ids[i] = memo_ids[j]
                    sc[:i, j] = 0
                    sc[i + 1:, j] = 0
                elif conf > nms_conf_thr:
                    ids[i] = -2
        return ids

    monkeypatch.setattr(qd.ops, "box_iou", lambda a, b, plus_one=False: _iou_np(a, b, plus_one))
    monkeypatch.setattr(qd.ops, "bisoftmax", bisoftmax)
    monkeypatch.setattr(qd.ops, "qd_assign", qd_assign)
    monkeypatch.setattr(qd, "assoc_stream", lambda d: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    g = np.load(os.path.join(ROOT, "tests", "golden", "qd_tracker.npz"))
    trk = qd.QuasiDenseEmbedTracker(device="cpu")
    for i, (boxes, feats) in enumerate(make_detections(int(g["n_frames"]), int(g["n_obj"]), int(g["seed"]))):
        b, _, ids = trk.match(boxes, torch.ones(boxes.size(0)), feats, i + 1)
        assert np.array_equal(ids.numpy(), g[f"ids_{i}"]), (i, ids, g[f"ids_{i}"])
        assert np.allclose(b.numpy(), g[f"boxes_{i}"])
    assert trk.num_tracklets == int(g["num_tracklets"])
