"""Golden fixture for the mask head (config 4) from the UNMODIFIED reference (unicorn_track_tiny_mask, 320x320,
VOS-style single object): reference UnicornHeadMask + postprocess_inst vs oracle; writes tests/golden/mask_tiny_320.npz."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
import unicorn_oracle as orc  # noqa: E402
from unicorn_b200.weights import make_state_dict  # noqa: E402
from unicorn_b200.synthetic import make_video  # noqa: E402

name, H, W = "unicorn_track_tiny_mask", 320, 320
sd = make_state_dict(name, 0)
_, model = ref_import.get_model(name)
print(model.load_state_dict(sd, strict=True))
from unicorn.utils.boxes import postprocess_inst  # noqa: E402
frames, boxes = make_video(2, H, W, seed=0)
cfg = orc.CONFIGS[name]
CONF, NMS, KEEP = 0.001, 0.65, 3
with torch.no_grad():
    _, pre = model(imgs=frames[0:1], mode="backbone")
    fpn, cur = model(imgs=frames[1:2], mode="backbone")
    f_pre, f_cur = model(seq_dict0=pre, seq_dict1=cur, mode="interaction")
    e_pre, e_cur = model(feat=f_pre, mode="upsample"), model(feat=f_cur, mode="upsample")
    lbs = F.interpolate(orc.get_label_map(boxes[0, 0], H, W), scale_factor=1 / 8, mode="bilinear", align_corners=False)[0].flatten(-2)
    trans = torch.softmax(torch.mm(e_pre.flatten(-2).squeeze().t(), e_cur.flatten(-2).squeeze()), dim=0)
    coarse = (lbs @ trans).view(1, -1, H // 8, W // 8)
    pri = (coarse, F.interpolate(coarse, scale_factor=1 / 2, mode="bilinear", align_corners=False),
           F.interpolate(coarse, scale_factor=1 / 4, mode="bilinear", align_corners=False))
    outs, locs, dyn, lvls, mf, um = model.head(fpn, pri, mode="sot")
    r_out = (outs.clone(), locs.clone(), dyn.clone(), lvls.clone(), mf.clone(), um.clone())
    dets, masks = postprocess_inst(outs, locs, dyn, lvls, mf, model.head.mask_head, 1, CONF, NMS, d_rate=2, up_masks=um[0:1])
    dets, masks = dets[0], masks[0]
    # oracle on the same inputs
    o_out = orc.head_forward_mask(fpn, pri, sd, cfg, "sot")
    for a, b, n in zip(o_out, r_out, ("outputs", "locations", "dyn", "levels", "mask_feats", "up_masks")):
        err = ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()
        print(f"{n:10s} oracle-vs-reference rel err {err:.3e}")
        assert err < 1e-4
    od, om = orc.postprocess_inst(r_out[0], r_out[1], r_out[2], r_out[3], r_out[4], r_out[5], 1, CONF, NMS, d_rate=2, max_masks=KEEP)
    assert od.shape == dets.shape
    print("dets", tuple(dets.shape), "masks", tuple(masks.shape), "mask err (first 3)", (om - masks[:KEEP]).abs().max().item())
    assert (om - masks[:KEEP]).abs().max().item() < 1e-4
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mask_tiny_320.npz"),
                    config=name, conf=CONF, nms=NMS, keep=KEEP, dets=dets.numpy(), mask0_sub=masks[0, 0, ::2, ::2].numpy().astype(np.float16),
                    mask_feats=mf.numpy(), up_masks_sub=um[0, :, ::4, ::4].numpy(), dyn_sub=dyn[0, ::16].numpy(),
                    mask_area=(masks[:KEEP, 0] > 0.5).float().mean(dim=(1, 2)).numpy())
print("wrote mask_tiny_320.npz")
