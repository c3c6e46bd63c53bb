"""Generate golden fixtures by running the UNMODIFIED reference (imported from /root/reference, build container
only) on seeded weights and synthetic frames, following external/lib/test/tracker/unicorn_sot.py:39-109 with
fp32 correlation, and check oracle/unicorn_oracle.py against it stage by stage.

Usage: python tests/golden/make_golden.py            (writes tests/golden/sot_tiny_320.npz)
"""
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


def run_reference_sot(model, frames, init_box, conf=0.001, nms=0.65):
    """unicorn_sot.py:39-56 (initialize) and :78-109 (get_det_results), fp32 throughout."""
    from unicorn.utils.boxes import postprocess
    H, W = frames.shape[-2:]
    with torch.no_grad():
        _, pre = model(imgs=frames[0:1], mode="backbone")
        dh, dw = pre["h"] * 2, pre["w"] * 2
        lbs = F.interpolate(orc.get_label_map(init_box, H, W), scale_factor=1 / 8, mode="bilinear", align_corners=False)[0].flatten(-2)
        out = []
        for t in range(1, frames.shape[0]):
            fpn, cur = model(imgs=frames[t:t + 1], mode="backbone")
            f_pre, f_cur = model(seq_dict0=pre, seq_dict1=cur, mode="interaction")
            e_pre = model(feat=f_pre, mode="upsample")
            e_cur = model(feat=f_cur, mode="upsample")
            simi = torch.mm(e_pre.flatten(-2).squeeze().transpose(1, 0), e_cur.flatten(-2).squeeze())
            trans = torch.softmax(simi, dim=0)
            coarse = (lbs @ trans).view(1, -1, dh, dw).float()
            pri = (coarse, F.interpolate(coarse, scale_factor=1 / 2, mode="bilinear", align_corners=False),
                   F.interpolate(coarse, scale_factor=1 / 4, mode="bilinear", align_corners=False))
            head = model.head(fpn, pri, mode="sot")
            head_keep = head.clone()
            dets = postprocess(head, 1, conf, nms)[0]
            out.append(dict(fpn=fpn, feat=cur["feat"], pos=cur["pos"], inter_pre=f_pre, inter_cur=f_cur, embed_pre=e_pre,
                            embed_cur=e_cur, coarse=coarse, head=head_keep, dets=dets))
    return out


def maxrel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def main():
    name, H, W, nf = "unicorn_track_tiny", 320, 320, 3
    sd = make_state_dict(name, seed=0)
    _, model = ref_import.get_model(name)
    missing = model.load_state_dict(sd, strict=True)
    print("load_state_dict strict:", missing)
    frames, boxes = make_video(nf, H, W, seed=0)
    ref = run_reference_sot(model, frames, boxes[0, 0])
    o = orc.SOTOracle(sd, name)
    o.initialize(frames[0:1], boxes[0, 0])
    worst = 0.0
    for t in range(1, nf):
        st = {}
        o.track(frames[t:t + 1], st)
        r = ref[t - 1]
        for k in ("feat", "pos", "inter_pre", "inter_cur", "embed_pre", "embed_cur", "coarse", "head"):
            e = maxrel(st[k], r[k]); worst = max(worst, e)
            print(f"frame {t} {k:10s} oracle-vs-reference max rel err {e:.3e}")
        for i in range(3):
            e = maxrel(st["fpn"][i], r["fpn"][i]); worst = max(worst, e)
            print(f"frame {t} fpn[{i}]     oracle-vs-reference max rel err {e:.3e}")
        assert r["dets"] is not None and st["dets"] is not None
        assert st["dets"].shape == r["dets"].shape, (st["dets"].shape, r["dets"].shape)
        # near-equal scores may swap order under 1e-7 perturbations: compare as sets (nearest row)
        d = torch.cdist(st["dets"][:, :6], r["dets"][:, :6], p=float("inf")).min(dim=0)[0].max().item() / r["dets"][:, :6].abs().max().item()
        worst = max(worst, d)
        print(f"frame {t} dets {tuple(r['dets'].shape)} set-distance rel {d:.3e}")
    assert worst < 1e-4, worst
    r = ref[-1]
    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.abspath(__file__)), "sot_tiny_320.npz"),
        config=name, seed=0, n_frames=nf, H=H, W=W, init_box=boxes[0, 0].numpy(),
        fpn0_sub=r["fpn"][0][0, :, ::4, ::4].numpy(), fpn1_sub=r["fpn"][1][0, :, ::2, ::2].numpy(), fpn2=r["fpn"][2][0, ::4].numpy(),
        feat_sub=r["feat"][0, ::4].numpy(), inter_cur_sub=r["inter_cur"][0, ::4].numpy(),
        embed_cur_sub=r["embed_cur"][0, :, ::4, ::4].numpy(), embed_pre_sub=r["embed_pre"][0, :, ::4, ::4].numpy(),
        coarse=r["coarse"].numpy(), head=r["head"].numpy(), dets=r["dets"].numpy(),
        dets_frame1=ref[0]["dets"].numpy(), head_frame1=ref[0]["head"].numpy())
    print("wrote sot_tiny_320.npz; worst oracle-vs-reference rel err", worst)


if __name__ == "__main__":
    main()
