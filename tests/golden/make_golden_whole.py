"""Golden fixture for `mode="whole"` (MOT prediction set, zero priors — unicorn/models/unicorn.py:133-139,
unicorn_head.py:289-304) from the UNMODIFIED reference, for the plain and the mask model; checks the oracle against it.

    python tests/golden/make_golden_whole.py      (writes tests/golden/whole_tiny_320.npz; build container only)

Stored: the reference head output of `model(imgs, mode="whole")` (8 classes -> 13 columns), `postprocess` detections at the
track_omni CLI thresholds (conf 0.01, nms 0.7; tools/track_omni.py:100-101), and for unicorn_track_tiny_mask the whole-mode
tuple of UnicornHeadMask (outputs, dynamic params, mask_feats, up_masks) with `postprocess_inst` detections and masks.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
import unicorn_oracle as orc  # noqa: E402
from unicorn_b200.weights import make_state_dict  # noqa: E402
from unicorn_b200.synthetic import make_video  # noqa: E402

H = W = 320
CONF, NMS, KEEP = 0.01, 0.7, 4


def maxrel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def main():
    out = {}
    frames, _ = make_video(2, H, W, seed=1, n_obj=3)
    img = frames[1:2]
    # ---------------------------------------------------------------- plain model
    name = "unicorn_track_tiny"
    sd = make_state_dict(name, 0)
    _, model = ref_import.get_model(name)
    print(model.load_state_dict(sd, strict=True))
    from unicorn.utils.boxes import postprocess, postprocess_inst
    with torch.no_grad():
        head, seq = model(img, mode="whole")
        dets = postprocess(head.clone(), 8, CONF, NMS)[0]
    cfg = orc.CONFIGS[name]
    fpn, _ = orc.forward_backbone(img, sd, cfg)
    o_head = orc.whole_forward(img, sd, cfg)[0]
    e = maxrel(o_head, head)
    print("whole head oracle-vs-reference rel err", e)
    assert e < 1e-4
    o_dets = orc.postprocess(o_head, 8, CONF, NMS)[0]
    assert o_dets.shape == dets.shape, (o_dets.shape, dets.shape)
    d = torch.cdist(o_dets[:, :6], dets[:, :6], p=float("inf")).min(dim=0)[0].max().item() / dets[:, :6].abs().max().item()
    assert d < 1e-4 and torch.equal(o_dets[:, 6].sort()[0], dets[:, 6].sort()[0]), d
    out.update(head=head.numpy(), dets=dets.numpy(), feat_sub=seq["feat"][0, ::4].numpy())
    # ---------------------------------------------------------------- mask model (MOTS path, mot_evaluator.py:776-803)
    name_m = "unicorn_track_tiny_mask"
    sd_m = make_state_dict(name_m, 0)
    _, model_m = ref_import.get_model(name_m)
    print(model_m.load_state_dict(sd_m, strict=True))
    with torch.no_grad():
        (outs, locs, dyn, lvls, mf, um), _ = model_m(img, mode="whole")
        r = (outs.clone(), locs.clone(), dyn.clone(), lvls.clone(), mf.clone(), um.clone())
        mdets, mmasks = postprocess_inst(outs, locs, dyn, lvls, mf, model_m.head.mask_head, 8, CONF, NMS, d_rate=2, up_masks=um[0:1])
        mdets, mmasks = mdets[0], mmasks[0]
    cfg_m = orc.CONFIGS[name_m]
    o = orc.whole_forward(img, sd_m, cfg_m)[0]
    for a, b, n in zip(o, r, ("outputs", "locations", "dyn", "levels", "mask_feats", "up_masks")):
        e = maxrel(a, b)
        print(f"mask whole {n:10s} oracle-vs-reference rel err {e:.3e}")
        assert e < 1e-4
    od, om = orc.postprocess_inst(r[0], r[1], r[2], r[3], r[4], r[5], 8, CONF, NMS, d_rate=2, max_masks=KEEP)
    assert od.shape == mdets.shape
    assert (om - mmasks[:KEEP]).abs().max().item() < 1e-4
    out.update(m_head=r[0].numpy(), m_dyn_sub=r[2][0, ::16].numpy(), m_mask_feats=r[4].numpy(), m_up_masks_sub=r[5][0, :, ::4, ::4].numpy(),
               m_dets=mdets.numpy(), m_mask_area=(mmasks[:KEEP, 0] > 0.3).float().mean(dim=(1, 2)).numpy(),
               m_mask0_sub=mmasks[0, 0, ::2, ::2].numpy().astype(np.float16))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "whole_tiny_320.npz"), conf=CONF, nms=NMS, keep=KEEP,
                        seed_video=1, n_obj=3, frame=1, **out)
    print("wrote whole_tiny_320.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
