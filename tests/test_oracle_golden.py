"""The CPU oracle (oracle/unicorn_oracle.py) against the golden outputs of the UNMODIFIED reference
(tests/golden/sot_tiny_320.npz, written by tests/golden/make_golden.py in the build container)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unicorn_oracle as orc  # noqa: E402
from unicorn_b200.synthetic import make_video  # noqa: E402
from unicorn_b200.weights import make_state_dict  # noqa: E402


@pytest.fixture(scope="module")
def run():
    g = np.load(os.path.join(ROOT, "tests", "golden", "sot_tiny_320.npz"))
    name = str(g["config"])
    sd = make_state_dict(name, int(g["seed"]))
    frames, boxes = make_video(int(g["n_frames"]), int(g["H"]), int(g["W"]), seed=0)
    assert np.allclose(boxes[0, 0].numpy(), g["init_box"])
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    o = orc.SOTOracle(sd, name)
    o.initialize(frames[0:1], boxes[0, 0])
    st1, st2 = {}, {}
    o.track(frames[1:2], st1)
    o.track(frames[2:3], st2)
    return g, st1, st2


def rel(a, b):
    b = torch.as_tensor(b)
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def test_stage_tensors(run):
    g, _, st = run
    assert rel(st["fpn"][0][0, :, ::4, ::4], g["fpn0_sub"]) < 1e-4
    assert rel(st["fpn"][1][0, :, ::2, ::2], g["fpn1_sub"]) < 1e-4
    assert rel(st["fpn"][2][0, ::4], g["fpn2"]) < 1e-4
    assert rel(st["feat"][0, ::4], g["feat_sub"]) < 1e-4
    assert rel(st["inter_cur"][0, ::4], g["inter_cur_sub"]) < 1e-4
    assert rel(st["embed_cur"][0, :, ::4, ::4], g["embed_cur_sub"]) < 1e-4
    assert rel(st["embed_pre"][0, :, ::4, ::4], g["embed_pre_sub"]) < 1e-4
    assert rel(st["coarse"], g["coarse"]) < 1e-4
    assert rel(st["head"], g["head"]) < 1e-4


def test_detections(run):
    g, st1, st2 = run
    for st, key in ((st1, "dets_frame1"), (st2, "dets")):
        ref = torch.from_numpy(g[key])
        assert st["dets"].shape == ref.shape
        d = torch.cdist(st["dets"][:, :6], ref[:, :6], p=float("inf")).min(dim=0)[0].max().item()
        assert d / ref[:, :6].abs().max().item() < 1e-4


def test_msda_core_matches_grid_sample_semantics():
    """Reference known-answer shapes (unicorn/models/ops/test.py:21-56, seed 3) through the reference's own
    pure-PyTorch formulation (ops/functions/ms_deform_attn_func.py:41-61)."""
    import torch.nn.functional as F
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    vals = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * loc - 1
    outs = []
    for lid, (h, w) in enumerate(shapes):
        v = vals[lid].flatten(2).transpose(1, 2).reshape(N * M, D, h, w)
        gr = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        outs.append(F.grid_sample(v, gr, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    ref = (torch.stack(outs, dim=-2).flatten(-2) * a).sum(-1).view(N, M * D, Lq).transpose(1, 2)
    assert torch.allclose(orc.msda_core(value, shapes, loc, attn), ref, rtol=1e-5, atol=1e-8)


def test_nms_restatement_matches_torchvision():
    import torchvision
    g = torch.Generator().manual_seed(0)
    n = 800
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 80 + 10
    boxes = torch.cat([xy, xy + wh], 1)
    scores = torch.rand(n, generator=g)
    cls = torch.randint(0, 3, (n,), generator=g)
    keep = torchvision.ops.batched_nms(boxes, scores, cls, 0.65)
    pred = torch.zeros(1, n, 8)
    pred[0, :, 0:2] = (boxes[:, :2] + boxes[:, 2:]) / 2
    pred[0, :, 2:4] = boxes[:, 2:] - boxes[:, :2]
    pred[0, :, 4] = scores
    pred[0, torch.arange(n), 5 + cls] = 1.0
    out = orc.postprocess(pred, 3, 0.0, 0.65)[0]
    assert out.shape[0] == keep.numel()
    assert torch.allclose(out[:, :4], boxes[keep], atol=1e-4)


def test_mask_path_matches_reference_golden():
    """Config 4 (UnicornHeadMask + CondInst dynamic masks): the oracle's backbone -> interaction -> correlation -> mask head ->
    postprocess_inst on the tiny mask model against tests/golden/mask_tiny_320.npz (outputs of the UNMODIFIED reference,
    tests/golden/make_golden_mask.py)."""
    import torch.nn.functional as F
    g = np.load(os.path.join(ROOT, "tests", "golden", "mask_tiny_320.npz"))
    name = str(g["config"])
    sd = make_state_dict(name, 0)
    cfg = orc.CONFIGS[name]
    frames, boxes = make_video(2, 320, 320, seed=0)
    with torch.no_grad():
        _, pre = orc.forward_backbone(frames[0:1], sd, cfg)
        fpn, cur = orc.forward_backbone(frames[1:2], sd, cfg)
        f_pre, f_cur = orc.deform_interaction(pre, cur, sd)
        e_pre, e_cur = orc.upsample_embed(f_pre, sd), orc.upsample_embed(f_cur, sd)
        pred = orc.corr_propagate(e_pre.flatten(-2)[0], e_cur.flatten(-2)[0], orc.label_map_s8(boxes[0, 0], 320, 320))
        pri = orc.prior_pyramid(pred.view(1, -1, 40, 40))
        outs, locs, dyn, lvls, mf, um = orc.head_forward_mask(fpn, pri, sd, cfg, "sot")
        dets, masks = orc.postprocess_inst(outs, locs, dyn, lvls, mf, um, 1, float(g["conf"]), float(g["nms"]), d_rate=2, max_masks=int(g["keep"]))
    assert rel(mf, g["mask_feats"]) < 1e-4
    assert rel(um[0, :, ::4, ::4], g["up_masks_sub"]) < 1e-4
    assert rel(dyn[0, ::16], g["dyn_sub"]) < 1e-4
    ref = torch.from_numpy(g["dets"])
    assert dets.shape == ref.shape and rel(dets[:, :6], ref[:, :6]) < 1e-4
    assert (masks[0, 0, ::2, ::2] - torch.from_numpy(g["mask0_sub"].astype(np.float32))).abs().max().item() < 2e-3  # fixture is fp16
    area = (masks[:, 0] > 0.5).float().mean(dim=(1, 2)).numpy()
    assert np.allclose(area, g["mask_area"], atol=1e-4)


def test_whole_mode_matches_reference_golden():
    """`mode="whole"` (unicorn.py:133-139: zero priors, MOT prediction set of 8 classes) for the plain and the mask model
    against tests/golden/whole_tiny_320.npz (outputs of the UNMODIFIED reference, tests/golden/make_golden_whole.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "whole_tiny_320.npz"))
    frames, _ = make_video(2, 320, 320, seed=int(g["seed_video"]), n_obj=int(g["n_obj"]))
    img = frames[int(g["frame"]):int(g["frame"]) + 1]
    conf, nms = float(g["conf"]), float(g["nms"])
    with torch.no_grad():
        name = "unicorn_track_tiny"
        head, seq = orc.whole_forward(img, make_state_dict(name, 0), orc.CONFIGS[name])
        assert head.shape == (1, 2100, 13) and rel(head, g["head"]) < 1e-4 and rel(seq["feat"][0, ::4], g["feat_sub"]) < 1e-4
        dets = orc.postprocess(head, 8, conf, nms)[0]
        ref = torch.from_numpy(g["dets"])
        assert dets.shape == ref.shape
        assert torch.cdist(dets[:, :6], ref[:, :6], p=float("inf")).min(dim=0)[0].max().item() / ref[:, :6].abs().max().item() < 1e-4
        name = "unicorn_track_tiny_mask"
        (outs, locs, dyn, lvls, mf, um), _ = orc.whole_forward(img, make_state_dict(name, 0), orc.CONFIGS[name])
        assert rel(outs, g["m_head"]) < 1e-4 and rel(dyn[0, ::16], g["m_dyn_sub"]) < 1e-4
        assert rel(mf, g["m_mask_feats"]) < 1e-4 and rel(um[0, :, ::4, ::4], g["m_up_masks_sub"]) < 1e-4
        md, mm = orc.postprocess_inst(outs, locs, dyn, lvls, mf, um, 8, conf, nms, d_rate=2, max_masks=int(g["keep"]))
        assert md.shape == g["m_dets"].shape and rel(md[:, :6], g["m_dets"][:, :6]) < 1e-4
        assert np.allclose((mm[:, 0] > 0.3).float().mean(dim=(1, 2)).numpy(), g["m_mask_area"], atol=1e-4)
        assert (mm[0, 0, ::2, ::2] - torch.from_numpy(g["m_mask0_sub"].astype(np.float32))).abs().max().item() < 2e-3


def test_vos_driver_matches_reference_class_golden():
    """oracle.VOSOracle (two first-frame objects, a third appearing in frame 2, soft aggregation, resize to the original frame)
    against the label maps produced by the UNMODIFIED reference class UnicornVOSTrack (tests/golden/make_golden_vos.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_vos_common import make_sequence, prep_frame, box_xyxy
    g = np.load(os.path.join(ROOT, "tests", "golden", "vos_tiny.npz"))
    name, size = str(g["config"]), tuple(int(v) for v in g["size"])
    H0, W0, new_at = int(g["H0"]), int(g["W0"]), int(g["new_at"])
    rgb, xywh, lab = make_sequence()
    r = min(size[0] / H0, size[1] / W0)
    o = orc.VOSOracle(make_state_dict(name, 0), name, half_corr=True)
    o.initialize(prep_frame(rgb[0], size), {"1": box_xyxy(xywh[0, 0], r), "2": box_xyxy(xywh[0, 1], r)}, orig_size=(H0, W0), r=r)
    for t in range(1, int(g["n_frames"])):
        new = {"3": box_xyxy(xywh[t, 2], r)} if t == new_at else None
        seg, _ = o.track(prep_frame(rgb[t], size), new, lab if t == new_at else None)
        assert (seg == g["segs"][t - 1]).mean() > 0.999
