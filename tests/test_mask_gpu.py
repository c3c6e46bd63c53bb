"""Mask path (config 4: UnicornHeadMask + CondInst dynamic masks) against the oracle and the reference golden."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_aligned_bilinear_add():
    import unicorn_oracle as orc
    from unicorn_b200 import ops
    g = torch.Generator().manual_seed(0)
    for f, (hs, ws) in ((2, (10, 13)), (4, (5, 7))):
        src = torch.randn(1, hs, ws, 64, generator=g).cuda().bfloat16()
        dst0 = torch.randn(1, hs * f, ws * f, 64, generator=g).cuda().bfloat16()
        dst = dst0.clone()
        ops.aligned_bilinear_add(src, dst, f)
        ref = dst0.float() + orc.aligned_bilinear(src.float().cpu().permute(0, 3, 1, 2), f).permute(0, 2, 3, 1).cuda()
        assert (dst.float() - ref).abs().max() < 3e-2


def test_dynamic_masks_exact_on_oracle_inputs():
    """kernels fed with the oracle's own head outputs must reproduce the oracle's masks (fp32 path, 1e-4)."""
    import unicorn_oracle as orc
    from unicorn_b200 import ops
    g = torch.Generator().manual_seed(1)
    h, w = 20, 24
    hw = [(h, w), (h // 2, w // 2), (h // 4, w // 4)]
    A = sum(a * b for a, b in hw)
    mf = torch.randn(1, 8, h, w, generator=g)
    um = torch.randn(1, 144, h, w, generator=g)
    dyn = torch.randn(1, A, 169, generator=g) * 0.5
    pred = torch.rand(1, A, 6, generator=g)
    pred[..., :2] *= torch.tensor([w * 8.0, h * 8.0])
    pred[..., 2:4] = pred[..., 2:4] * 60 + 10
    locs, lv = [], []
    for k, (a, b) in enumerate(hw):
        yv, xv = torch.meshgrid(torch.arange(a), torch.arange(b), indexing="ij")
        locs.append((torch.stack((xv, yv), 2).view(-1, 2).float() + 0.5) * (8, 16, 32)[k])
        lv.append(torch.full((1, a * b), k))
    locs, lv = torch.cat(locs), torch.cat(lv, 1)
    dets, masks = orc.postprocess_inst(pred, locs, dyn, lv, mf, um, 1, 0.3, 0.65, d_rate=2, max_masks=4)
    ws = ops.PostWorkspace(A, "cuda")
    ops.postprocess_device(pred[0].cuda().contiguous(), 1, 0.3, 0.65, ws)
    n = int(ws.count.item())
    assert n == dets.shape[0] and torch.allclose(ws.dets[:n].cpu(), dets, atol=1e-5)
    dl, off = [], 0
    for (a, b) in hw:
        t = torch.zeros(1, a, b, 176)
        t[..., :169] = dyn[0, off:off + a * b].view(1, a, b, 169)
        dl.append(t.cuda().contiguous())
        off += a * b
    got = ops.dynamic_masks(mf.permute(0, 2, 3, 1).contiguous().cuda(), um.permute(0, 2, 3, 1).contiguous().cuda(), dl, hw, ws, 4)
    assert (got.cpu() - masks[:, 0]).abs().max() < 1e-4


def test_vos_frame_vs_oracle_and_golden():
    import unicorn_oracle as orc
    from unicorn_b200.engine import UnicornEngine
    from unicorn_b200.synthetic import make_video
    from unicorn_b200.vos import UnicornVOSTrack
    from unicorn_b200.weights import make_state_dict
    g = np.load(os.path.join(ROOT, "tests", "golden", "mask_tiny_320.npz"))
    name = str(g["config"])
    sd = make_state_dict(name, 0)
    frames, boxes = make_video(2, 320, 320, seed=0)
    eng = UnicornEngine(sd, name)
    vos = UnicornVOSTrack(eng, (320, 320), conf=float(g["conf"]), nms=float(g["nms"]))
    vos.debug = True
    vos.initialize_tensor(frames[0:1], {"1": boxes[0, 0]})
    res = vos.track_tensor(frames[1:2])
    det, mask = res["objects"]["1"]
    assert det is not None
    # mask branch outputs vs the reference golden (bf16 path)
    mf = vos.last["mask_feats"].permute(0, 3, 1, 2).cpu()
    rel = lambda a, b: ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()  # noqa: E731
    assert rel(mf, torch.from_numpy(g["mask_feats"])) < 8e-2
    assert rel(vos.last["up_masks"].permute(0, 3, 1, 2).cpu()[0, :, ::4, ::4], torch.from_numpy(g["up_masks_sub"])) < 8e-2
    # the best instance's mask: same object as one of the reference's top detections, soft-mask IoU high
    ref_dets = torch.from_numpy(g["dets"])
    iou = orc.box_iou_np(det[None, :4].numpy(), ref_dets[:10, :4].numpy())
    assert iou.max() > 0.7
    # exactness of the mask kernels on the engine's own intermediates: re-run the oracle's dynamic mask head on them
    po = vos.last["per_obj"]["1"]
    head = po["head"].cpu()
    dyn = torch.cat([t[0, :, :, :169].reshape(-1, 169) for t in po["dyn"]], 0).cpu()[None]
    # controller output (169 dynamic-conv parameters per anchor, unicorn_head_mask.py:333-334) vs the reference golden
    assert rel(dyn[0, ::16], torch.from_numpy(g["dyn_sub"])) < 8e-2
    locs, lv = [], []
    for k, t in enumerate(po["dyn"]):
        a, b = t.shape[1:3]
        yv, xv = torch.meshgrid(torch.arange(a), torch.arange(b), indexing="ij")
        locs.append((torch.stack((xv, yv), 2).view(-1, 2).float() + 0.5) * (8, 16, 32)[k])
        lv.append(torch.full((1, a * b), k))
    od, om = orc.postprocess_inst(head, torch.cat(locs), dyn, torch.cat(lv, 1), mf, vos.last["up_masks"].permute(0, 3, 1, 2).cpu(),
                                  1, float(g["conf"]), float(g["nms"]), d_rate=2, max_masks=1)
    assert torch.allclose(od[0], det, atol=1e-4)
    assert (om[0, 0] - mask.cpu()).abs().max() < 1e-3
    m_bin, r_bin = (mask.cpu() > 0.5), (om[0, 0] > 0.5)
    inter, union = (m_bin & r_bin).sum().item(), (m_bin | r_bin).sum().item()
    assert union == 0 or inter / union >= 0.999  # mask IoU vs the oracle on identical head outputs
