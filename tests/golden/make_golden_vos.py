"""Golden fixture for the VOS driver from the UNMODIFIED reference class external/lib/test/tracker/unicorn_vos.py
(UnicornVOSTrack.initialize / track): 5 RGB frames 256x400, two objects in the first frame, a third one appearing in frame 2
(reference groups, soft aggregation, mask resize to the original frame, PreprocessorX letterbox) on unicorn_track_tiny_mask with
seeded weights.  Build container only.  Writes tests/golden/vos_tiny.npz and checks oracle.VOSOracle against it.

The class is constructed through its own __init__ (get_exp, get_model, torch.load of a checkpoint written from the seeded
state_dict, load_state_dict, .cuda(), .eval()); the only changes are environmental: CPU redirection of "cuda" (oracle/ref_import),
`np.int` alias (removed from numpy >= 1.24; unicorn_vos.py:146 uses it) and the instance attribute `input_size` set to 320x320
(the exp's 800x1280 needs minutes per frame in fp16 on a CPU).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
import unicorn_oracle as orc  # noqa: E402
from unicorn_b200.weights import make_state_dict  # noqa: E402

NAME, H0, W0, SIZE, NF, NEW_AT = "unicorn_track_tiny_mask", 256, 400, (320, 320), 5, 2


sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_vos_common import make_sequence, prep_frame, box_xyxy  # noqa: E402


def main():
    ref_import.install()
    np.int = int  # noqa: unicorn_vos.py:146 (numpy < 1.24 alias)
    _to = torch.Tensor.to

    def to_cpu(self, *a, **k):  # .to("cuda") / .to(self.device) with device == "cuda"
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return _to(self, *a, **k)
    torch.Tensor.to = to_cpu
    sys.path.insert(0, os.path.join(ref_import.REF_ROOT, "external"))
    sd = make_state_dict(NAME, 0)
    ckpt = os.path.join(tempfile.mkdtemp(), "ckpt.pth")
    torch.save({"model": sd}, ckpt)
    cwd = os.getcwd()
    os.chdir(ref_import.REF_ROOT)  # get_exp("exps/default/<name>") is relative to the repository root
    try:
        from lib.test.tracker.unicorn_vos import UnicornVOSTrack
        params = types.SimpleNamespace(exp_name=NAME, checkpoint=ckpt)
        trk = UnicornVOSTrack(params, "dv2017_val")
    finally:
        os.chdir(cwd)
    trk.input_size = SIZE
    rgb, xywh, lab = make_sequence()
    info0 = {"init_object_ids": ["1", "2"], "sequence_object_ids": ["1", "2", "3"],
             "init_bbox": {"1": xywh[0, 0].tolist(), "2": xywh[0, 1].tolist()}}
    trk.initialize(rgb[0], info0)
    segs, states = [], []
    for t in range(1, NF):
        info = {}
        if t == NEW_AT:
            info = {"init_object_ids": ["3"], "init_bbox": {"3": xywh[t, 2].tolist()}, "init_mask": lab}
        out = trk.track(rgb[t], info)
        segs.append(out["segmentation"].copy())
        states.append(np.array([trk.state_pre_dict[o] for o in ("1", "2")], dtype=np.float32))
        print("frame", t, "labels", {int(v): int((segs[-1] == v).sum()) for v in np.unique(segs[-1])}, "state", states[-1].tolist())
    # ---- oracle restatement on the same frames (fp16 casts of the correlation mimicked)
    o = orc.VOSOracle(sd, NAME, half_corr=True)
    r = min(SIZE[0] / H0, SIZE[1] / W0)
    prep = lambda img: prep_frame(img, SIZE)  # noqa: E731
    box = lambda b: box_xyxy(b, r)  # noqa: E731
    o.initialize(prep(rgb[0]), {"1": box(xywh[0, 0].tolist()), "2": box(xywh[0, 1].tolist())}, orig_size=(H0, W0), r=r)
    agree = []
    for t in range(1, NF):
        new = {"3": box(xywh[t, 2].tolist())} if t == NEW_AT else None
        seg, _ = o.track(prep(rgb[t]), new, lab if t == NEW_AT else None)
        agree.append(float((seg == segs[t - 1]).mean()))
    print("oracle-vs-reference label agreement per frame:", agree)
    assert min(agree) > 0.999, agree
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vos_tiny.npz"), config=NAME, H0=H0, W0=W0, size=np.array(SIZE),
                        n_frames=NF, new_at=NEW_AT, seed_video=7, segs=np.stack(segs), states=np.stack(states), oracle_agreement=np.array(agree))
    print("wrote vos_tiny.npz")


if __name__ == "__main__":
    main()
