"""MOTS driver (unicorn_b200/mots.py) — written without GPU access at the end of round 1; first thing to run in round 2."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_mots_driver_masks_match_oracle_and_format():
    import unicorn_oracle as orc
    from unicorn_b200 import results as R
    from unicorn_b200.engine import UnicornEngine
    from unicorn_b200.mots import UnicornMOTSTracker
    from unicorn_b200.synthetic import make_video
    from unicorn_b200.tracker import QuasiDenseEmbedTracker
    from unicorn_b200.weights import make_state_dict
    name = "unicorn_track_tiny_mask"
    eng = UnicornEngine(make_state_dict(name, 0), name)
    frames, _ = make_video(4, 320, 320, seed=1, n_obj=3)
    trk = UnicornMOTSTracker(eng, (320, 320), conf=0.01, nms=0.7, score_thr=0.02, max_dets=16, min_box_area=0,
                             tracker=QuasiDenseEmbedTracker(init_score_thr=0.05, obj_score_thr=0.03))
    seen = []
    for t in range(4):
        fr = trk.step_tensor(frames[t:t + 1], 320, 320)
        assert fr[0] == t + 1 and fr[2:5] == (2, 320, 320) and len(fr[1]) == len(fr[5])
        assert fr[1] == sorted(fr[1]) and all(i >= 1 for i in fr[1])
        dec = [R.rle_decode(s, 320, 320) for s in fr[5]]
        if dec:
            assert np.stack(dec).sum(0).max() <= 1  # overlap free
        seen.append(fr[1])
        # the dynamic masks of the kept detections against the oracle's mask head on the engine's own head outputs
        last = trk.last
        head = last["head"].cpu()
        dyn = torch.cat([d[0, :, :, :169].reshape(-1, 169) for d in last["dyn"]], 0).cpu()[None]
        locs, lv = [], []
        for k, d in enumerate(last["dyn"]):
            a, b = d.shape[1:3]
            yv, xv = torch.meshgrid(torch.arange(a), torch.arange(b), indexing="ij")
            locs.append((torch.stack((xv, yv), 2).view(-1, 2).float() + 0.5) * (8, 16, 32)[k])
            lv.append(torch.full((1, a * b), k))
        n = min(last["dets"].shape[0], 4)
        od, om = orc.postprocess_inst(head, torch.cat(locs), dyn, torch.cat(lv, 1), last["mask_feats"].permute(0, 3, 1, 2).cpu(),
                                      last["up_masks"].permute(0, 3, 1, 2).cpu(), eng.ncls, 0.01, 0.7, d_rate=2, max_masks=n)
        assert torch.allclose(od[:n], last["dets"][:n], atol=1e-4)
        assert (om[:n, 0] - last["masks"][:n].cpu()).abs().max() < 1e-3
    assert any(seen), "no tracked instance in 4 frames"
