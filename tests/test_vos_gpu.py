"""VOS driver (unicorn_b200/vos.py) against the label maps of the UNMODIFIED reference class UnicornVOSTrack
(tests/golden/vos_tiny.npz: two first-frame objects, a third appearing in frame 2, soft aggregation, mask resize to the original
frame) and the device-side result assembly (uc_vos_aggregate) against its torch / numpy definition."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_vos_aggregate_kernel_matches_definition():
    """unicorn_vos.py:139-155 (F.interpolate(scale_factor=1/r)[:H,:W] into a zero map) + :105-121 (float32 background product in
    list order, argmax with the lower channel winning ties) on random soft masks, one object given by its initial label map."""
    import ctypes
    from unicorn_b200 import _lib
    g = torch.Generator().manual_seed(3)
    Hin, Win, H0, W0 = 96, 160, 113, 187
    r = min(Hin / H0, Win / W0)
    ids = [4, 1, 7, 2]  # list order != id order: the product follows the list, ties go to the lower id
    masks = torch.rand(3, Hin, Win, generator=g)
    masks[0, :40] = 1.0   # saturated region: background product is exactly 0 there
    masks[1, :20] = 1.0   # ... and two objects tie at 1.0
    lab = torch.zeros(H0, W0, dtype=torch.uint8)
    lab[30:70, 50:120] = 2
    soft_ref = np.zeros((4, H0, W0), dtype=np.float32)
    m = F.interpolate(masks[:, None], scale_factor=1 / r, mode="bilinear", align_corners=False)[:, 0, :H0, :W0]
    soft_ref[:3, :m.shape[1], :m.shape[2]] = m.numpy()
    soft_ref[3] = (lab.numpy() == 2)
    merge = np.zeros((H0, W0, 8))
    for k, i in enumerate(ids):
        merge[:, :, i] = soft_ref[k]
    merge[:, :, 0] = np.prod(1 - np.stack(list(soft_ref), -1), axis=-1)
    seg_ref = np.argmax(merge, -1).astype(np.uint8)
    md, ld = masks.cuda().contiguous(), lab.cuda()
    soft = torch.zeros(4, H0, W0, device="cuda")
    seg = torch.zeros(H0, W0, dtype=torch.uint8, device="cuda")
    objs = (_lib.UcVosObject * 4)()
    for k, i in enumerate(ids):
        objs[k].id = i
        if k < 3:
            objs[k].mask = md[k].data_ptr()
        else:
            objs[k].init_mask = ld.data_ptr()
    _lib.check(_lib.lib().uc_vos_aggregate(objs, 4, Hin, Win, H0, W0, ctypes.c_float(r), ctypes.c_void_p(soft.data_ptr()),
                                           ctypes.c_void_p(seg.data_ptr()), _lib.stream_ptr()), "uc_vos_aggregate")
    assert np.abs(soft.cpu().numpy() - soft_ref).max() < 2e-6
    agree = (seg.cpu().numpy() == seg_ref).mean()
    assert agree > 0.9995, agree  # exact up to last-ulp differences of the bilinear weights at near ties


def _run_driver(use_graph):
    from make_golden_vos_common import make_sequence
    from unicorn_b200.engine import UnicornEngine
    from unicorn_b200.vos import UnicornVOSTrack
    from unicorn_b200.weights import make_state_dict
    g = np.load(os.path.join(ROOT, "tests", "golden", "vos_tiny.npz"))
    name, size, new_at = str(g["config"]), tuple(int(v) for v in g["size"]), int(g["new_at"])
    rgb, xywh, lab = make_sequence()
    trk = UnicornVOSTrack(UnicornEngine(make_state_dict(name, 0), name), size, use_graph=use_graph)
    trk.initialize(rgb[0], {"init_object_ids": ["1", "2"], "sequence_object_ids": ["1", "2", "3"],
                            "init_bbox": {"1": xywh[0, 0].tolist(), "2": xywh[0, 1].tolist()}})
    segs, states = [], []
    for t in range(1, int(g["n_frames"])):
        info = {"init_object_ids": ["3"], "init_bbox": {"3": xywh[t, 2].tolist()}, "init_mask": lab} if t == new_at else {}
        segs.append(trk.track(rgb[t], info)["segmentation"].copy())
        states.append([trk.state_pre_dict[o] for o in ("1", "2")])
    return g, segs, states, trk


def test_vos_driver_vs_reference_class_golden():
    g, segs, states, trk = _run_driver(False)
    agree = [float((s == r).mean()) for s, r in zip(segs, g["segs"])]
    iou = {}
    for t, (s, r) in enumerate(zip(segs, g["segs"])):
        for i in (1, 2, 3):
            u = ((s == i) | (r == i)).sum()
            if u > 500:
                iou[(t + 1, i)] = float(((s == i) & (r == i)).sum() / u)
    print("VOS label agreement with the reference class per frame:", agree, "per-object IoU:", iou)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        import json
        json.dump(dict(label_agreement=agree, iou={f"{k[0]}:{k[1]}": v for k, v in iou.items()}), open(os.path.join(out, "r2_vos_parity.json"), "w"))
    # bf16 features vs the reference's fp32: seeded random weights give noise-like soft masks, so labels near a 0.5 crossing flip
    assert min(agree) > 0.9, agree
    assert segs[int(g["new_at"]) - 1].max() == 3  # the new object's initial mask went through the aggregation
    ds = np.abs(np.array(states, dtype=np.float32) - g["states"]).max()
    print("max |state box - reference| (pixels):", ds)


def test_vos_graph_replay_matches_eager():
    _, segs_e, st_e, _ = _run_driver(False)
    _, segs_g, st_g, trk = _run_driver(True)
    assert trk.launches_per_frame > 0
    for a, b in zip(segs_e, segs_g):
        assert np.array_equal(a, b)
    assert st_e == st_g
